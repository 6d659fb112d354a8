cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
one() { # name workload extra
  timeout 300 python bench.py --workload $2 $3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
}
for cfg in "4 272" "1 80" "1 112" "1 144" "2 112" "2 144" "2 208"; do set -- $cfg
  export IQO_CUDA_MMA_WARPS=$1 IQO_CUDA_MMA_WCOLS=$2
  one w$1-c$2 cfg5s_lanczos4_8192_to_3000 ""
  one w$1-c$2 area_1080p_to_720p ""
  one w$1-c$2 cfg1_lanczos3_1080p_to_720p "--path mma"
done
