cp libiqo_b200/lib/libiqo_cuda.so build/variants/_orig.so
for v in s16 s24; do cp build/variants/$v.so libiqo_b200/lib/libiqo_cuda.so; for c in none 50 65 75 85 100; do echo "variant $v carveout $c"; if [ $c = none ]; then unset IQO_CUDA_STREAM_CARVEOUT; else export IQO_CUDA_STREAM_CARVEOUT=$c; fi; bash tools/dev/runbench.sh; done; done
cp build/variants/_orig.so libiqo_b200/lib/libiqo_cuda.so
