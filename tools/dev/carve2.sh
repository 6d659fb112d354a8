for c in none 65 75 85 100; do echo "carveout $c"; if [ $c = none ]; then unset IQO_CUDA_STREAM_CARVEOUT; else export IQO_CUDA_STREAM_CARVEOUT=$c; fi; bash tools/dev/runbench.sh; done
