cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
V=$PWD/libiqo_b200/lib/variants
one() { # name lib workload extra
  IQO_CUDA_LIB=$2 timeout 300 python bench.py --workload $3 $4 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
}
for wc in 272 208 144; do for n in base m12 m14 m16; do
  if [ $n = base ]; then L=""; else L=$V/libiqo_cuda_$n.so; fi
  export IQO_CUDA_MMA_WCOLS=$wc
  one $n-w$wc "$L" cfg5s_lanczos4_8192_to_3000 ""
  one $n-w$wc "$L" area_1080p_to_720p ""
  one $n-w$wc "$L" cfg1_lanczos3_1080p_to_720p "--path mma"
done; done
