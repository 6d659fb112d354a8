set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for w in 4; do IQO_CUDA_MMA_WARPS=$w timeout 600 python -m pytest tests/test_gpu_mma.py -q -x -m gpu --timeout 300 2>&1 | tail -3; done
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
except Exception as e:
    print("no json", e); sys.exit(0)
print("HEAD", d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], 'parity', d['parity']['bit_exact'])
for w in d.get('workloads', []):
    if 'lanczos' in w.get('workload',''): print("  WL %-38s %-16s ms %-8s frac %-7s parity %s" % (w.get('workload'), w.get('kernel','')[:16], w.get('ms'), w.get('frac'), (w.get('parity') or {}).get('bit_exact'), ), w.get('error',''))
c=d.get('cfg5') or {}
print("  cfg5", c.get('kernel'), c.get('ms_kernel'), c.get('ms_e2e'), c.get('hash_ok'), c.get('error'))
P
}
for wc in 208 272; do for w in 2 4; do
  IQO_CUDA_MMA_AUTO=1 IQO_CUDA_MMA_WCOLS=$wc IQO_CUDA_MMA_WARPS=$w timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_mma_w${wc}_nw$w.json 2> gpurun_out/r2_bench_mma_w${wc}_nw$w.err; echo "rc=$? wcols=$wc warps=$w"; summ gpurun_out/r2_bench_mma_w${wc}_nw$w.json
done; done
