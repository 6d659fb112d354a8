cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
IQO_CUDA_MMA_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bands.py -q -x -m gpu --timeout 300 2>&1 | tail -2
for di in 0 1; do for wl in cfg5s_lanczos4_8192_to_3000 cfg1_lanczos3_1080p_to_720p area_1080p_to_720p linear_720p_to_1080p cfg4_lanczos3_1080p_to_540p; do
  IQO_CUDA_MMA_DIRECT=$di timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('direct=$di', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; IQO_CUDA_MMA_DIRECT=$di timeout 300 python tools/gigapixel.py --steps 2 | cut -c150-260; done
