cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for rep in 1 2; do for pr in 3 2 1 0; do
  IQO_CUDA_TMA_PROMO=$pr timeout 300 python bench.py --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('promo=$pr', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"
done; done
for pr in 3 2 0; do IQO_CUDA_TMA_PROMO=$pr timeout 300 python bench.py --workload cfg5s_lanczos4_8192_to_3000 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('promo=$pr', d['config']['workload'], d['ms_per_step'], d['roofline']['frac'], d['parity']['bit_exact'])"; done
