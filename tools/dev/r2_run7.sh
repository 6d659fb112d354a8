set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=20 --timeout 600 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest2.log
