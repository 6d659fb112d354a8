cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_float_mode.py -q -m gpu --timeout 300 2>&1 | tail -25
