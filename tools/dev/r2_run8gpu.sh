set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_bands.py -q -m gpu -k "multi_device" > gpurun_out/r2_pytest_multi_8gpu.log 2>&1; echo "multi test rc=$?"; tail -3 gpurun_out/r2_pytest_multi_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r2_bench_8gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu.json').read().strip().split('\n')[-1])
print("HEAD", d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'])
print("e2e", json.dumps(d['e2e']))
print("cfg5", json.dumps(d.get('cfg5')))
print("multi", json.dumps(d.get('multi_device')))
P
