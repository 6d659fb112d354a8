set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bands.py tests/test_gpu_guards.py -q -x -m gpu --timeout 300 > gpurun_out/r2_mma_tests3.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_mma_tests3.log
for wc in 272; do for w in 4; do for wl in cfg4_lanczos3_1080p_to_540p cfg1_lanczos3_1080p_to_720p cfg5s_lanczos4_8192_to_3000 area_1080p_to_720p linear_720p_to_1080p; do
  IQO_CUDA_MMA_WCOLS=$wc IQO_CUDA_MMA_WARPS=$w timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wc $w', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; done; done
for bb in 4 8; do IQO_CUDA_MMA_BAND_BLOCKS=$bb timeout 300 python tools/gigapixel.py --steps 2 | cut -c1-260; done
