cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for rep in 1; do for p2 in 0; do for wl in cfg5s_lanczos4_8192_to_3000 cfg1_lanczos3_1080p_to_720p cfg4_lanczos3_1080p_to_540p; do
  IQO_CUDA_MMA_POW2=$p2 timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pow2=$p2', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'])"
done; done; done
