cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tag=${1:-r2_final}
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.txt; cat gpurun_out/${tag}_pytest.txt; fi
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"])
for w in d.get("workloads", []): print(" ", w.get("workload"), w.get("kernel"), w.get("ms"), w.get("frac"), w.get("parity"))
print(" cfg5", d.get("cfg5")); print(" e2e_single", d.get("e2e_single"))
PY
