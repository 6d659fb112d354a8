cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
V=libiqo_b200/lib/variants
one() { # name lib workload extra
  IQO_CUDA_LIB=$2 timeout 300 python bench.py --workload $3 $4 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
}
for n in hlate10 base12; do
  IQO_CUDA_LIB=$PWD/$V/libiqo_cuda_$n.so timeout 900 python -m pytest tests/test_gpu_mma.py -x -q 2>&1 | tail -2
done
for rep in 1 2; do for n in base vpair10 hlate10 both10 base12; do
  if [ $n = base ]; then L=""; else L=$PWD/$V/libiqo_cuda_$n.so; fi
  one $n "$L" cfg5s_lanczos4_8192_to_3000 ""
  one $n "$L" area_1080p_to_720p ""
  one $n "$L" cfg1_lanczos3_1080p_to_720p "--path mma"
done; done
