cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
V=$PWD/libiqo_b200/lib/variants
one() { # name lib workload extra
  IQO_CUDA_LIB=$2 timeout 300 python bench.py --workload $3 $4 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
}
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bands.py -x -q 2>&1 | tail -3
for rep in 1 2; do for n in ${VARIANTS:-head base new8}; do
  if [ $n = base ]; then L=""; else L=$V/libiqo_cuda_$n.so; fi
  one $n "$L" cfg5s_lanczos4_8192_to_3000 ""
  one $n "$L" area_1080p_to_720p ""
  one $n "$L" cfg1_lanczos3_1080p_to_720p "--path mma"
  one $n "$L" linear_720p_to_1080p ""
done; done
one base "" cfg1_lanczos3_1080p_to_720p "--path no_mma"
