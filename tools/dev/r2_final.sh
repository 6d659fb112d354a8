# final evidence run of round 2: GPU tests, both bench arms, smoke, launch list of the bench command
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_final_pytest.txt; cat gpurun_out/r2_final_pytest.txt
t0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_final_bench_cuda.json 2> gpurun_out/r2_final_bench_cuda.err; rc=$?; echo "cuda rc=$rc wall=$(( $(date +%s) - t0 ))s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
if [ $rc = 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_final_ncu.log 2>&1; echo "ncu rc=$?"
fi
python - <<PY
import json
for f in ("gpurun_out/r2_final_bench_reference.json", "gpurun_out/r2_final_bench_cuda.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("cpu_baseline"), d.get("clocks"))
PY
