set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=20 --timeout 600 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest2.log
timeout 600 python bench.py > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_v2.json').read())
print("HEAD", d['detail']['kernel'], d['ms_per_step'], d['roofline']['frac'], 'parity', d['parity']['bit_exact'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for w in d.get('workloads', []):
    print("  WL %-38s %-16s ms %-8s frac %-7s parity %s" % (w.get('workload'), w.get('kernel','')[:16], w.get('ms'), w.get('frac'), (w.get('parity') or {}).get('bit_exact')), w.get('error',''))
c=d.get('cfg5') or {}
print("  cfg5", c.get('kernel'), c.get('ms_kernel'), c.get('ms_e2e'), c.get('hash_ok'), c.get('error'))
P
for bb in 4 8 16; do IQO_CUDA_MMA_BAND_BLOCKS=$bb timeout 300 python tools/gigapixel.py --steps 2 | cut -c1-330; done
timeout 300 python tools/gigapixel.py --steps 2 --path stream | cut -c1-330
