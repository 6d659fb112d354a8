set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=20 --timeout 600 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest3.log
timeout 600 python bench.py > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_v3.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_v3.json').read())
print("HEAD", d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], 'parity', d['parity']['bit_exact'], 'cpu', d['cpu_baseline']['value'])
print("e2e", json.dumps(d['e2e']))
for w in d.get('workloads', []):
    print("  WL %-38s %-16s ms %-8s frac %-7s parity %s" % (w.get('workload'), w.get('kernel','')[:16], w.get('ms'), w.get('frac'), (w.get('parity') or {}).get('bit_exact')), w.get('error',''))
c=d.get('cfg5') or {}
print("  cfg5", c.get('kernel'), c.get('ms_kernel'), c.get('ms_e2e'), c.get('hash_ok'), c.get('error'))
print(json.dumps(d['e2e_single']))
P
for t in 1 2 4 8; do IQO_CUDA_COPY_THREADS=$t timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('copy threads $t pageable', d['e2e']['pageable'], 'pinned', d['e2e']['value'])"; done
