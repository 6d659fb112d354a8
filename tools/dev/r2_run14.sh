cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for wc in 272 304 336; do for tm in 2 4; do for wl in cfg5s_lanczos4_8192_to_3000 cfg1_lanczos3_1080p_to_720p area_1080p_to_720p; do
  IQO_CUDA_MMA_WCOLS=$wc IQO_CUDA_MMA_TILE_MULT=$tm timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wcols=$wc mult=$tm', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; done; done
