for bp in 24 45 90 135 270; do echo "bandPairs $bp"; IQO_CUDA_STREAM_BAND_PAIRS=$bp bash tools/dev/runbench.sh; done
