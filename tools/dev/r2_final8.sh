# final 8-GPU evidence: 1-GPU line on the same box, the 8-GPU line (torchrun), multi-device tests, reference arm
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_final_topo_8gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_bands.py -q -m gpu -k "multi_device" > gpurun_out/r2_final_pytest_multi_8gpu.txt 2>&1; echo "multi test rc=$?"; tail -3 gpurun_out/r2_final_pytest_multi_8gpu.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2_final8_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2_final8_bench_1gpu.json 2> gpurun_out/r2_final8_bench_1gpu.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_final8_bench_8gpu.json 2> gpurun_out/r2_final8_bench_8gpu.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r2_final8_bench_8gpu.err
python - <<'P'
import json
for f in ("gpurun_out/r2_final8_bench_reference.json", "gpurun_out/r2_final8_bench_1gpu.json", "gpurun_out/r2_final8_bench_8gpu.json"):
    d=json.loads(open(f).read().strip().split('\n')[-1])
    print(f, d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
    if d.get('e2e'): print("  e2e", d['e2e'].get('value'), d['e2e'].get('pipeline_efficiency'), d['e2e'].get('copy_only'))
    if d.get('cfg5'): print("  cfg5", d['cfg5'].get('ms_kernel'), d['cfg5'].get('ms_e2e'), d['cfg5'].get('hash_ok'))
    if d.get('multi_device'): print("  multi", json.dumps(d['multi_device'])[:300])
P
