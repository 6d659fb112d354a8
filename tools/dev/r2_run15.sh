cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
IQO_CUDA_MMA_EARLY=1 timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bands.py -q -x -m gpu --timeout 300 2>&1 | tail -2
for ea in 0 1; do for wl in cfg5s_lanczos4_8192_to_3000 cfg1_lanczos3_1080p_to_720p area_1080p_to_720p linear_720p_to_1080p; do
  IQO_CUDA_MMA_EARLY=$ea timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('early=$ea', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; IQO_CUDA_MMA_EARLY=$ea timeout 300 python tools/gigapixel.py --steps 2 | cut -c150-260; done
