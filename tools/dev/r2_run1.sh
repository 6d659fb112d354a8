# round-2 GPU call 1: smoke (TMA-fed stream kernel first, under a short timeout), GPU test suite, bench both FIFO variants, reference arm
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "half_kernel and not extreme" -m gpu > gpurun_out/r2_half_tests.log 2>&1; echo "half rc=$?"; tail -5 gpurun_out/r2_half_tests.log
timeout 1500 python -m pytest tests -q -m gpu --maxfail=40 --timeout 600 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest.log
timeout 600 python bench.py > gpurun_out/r2_bench_tma1.json 2> gpurun_out/r2_bench_tma1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench_tma1.json
IQO_CUDA_STREAM_TMA=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench_tma0.json 2> gpurun_out/r2_bench_tma0.err; echo "bench0 rc=$?"; cut -c1-400 gpurun_out/r2_bench_tma0.json
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference.json
for st in 5 20 40; do timeout 200 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps $st | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps',d['steps'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],d['clocks'])"; done > gpurun_out/r2_steps_tma1.txt 2>&1
cat gpurun_out/r2_steps_tma1.txt
