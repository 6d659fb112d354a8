set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for wc in 272 400; do for w in 4; do for wl in cfg4_lanczos3_1080p_to_540p cfg1_lanczos3_1080p_to_720p cfg5s_lanczos4_8192_to_3000 cfg3y_lanczos2_2160p_to_1080p; do
  IQO_CUDA_MMA_WCOLS=$wc IQO_CUDA_MMA_WARPS=$w timeout 300 python bench.py --workload $wl --path mma --no-extras --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wc $w', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; done; done
