set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mma.py -q -x -m gpu --timeout 300 > gpurun_out/r2_mma_tests.log 2>&1; echo "mma tests rc=$?"; tail -25 gpurun_out/r2_mma_tests.log
