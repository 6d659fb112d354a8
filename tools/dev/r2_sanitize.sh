# compute-sanitizer pass over a reduced GPU selection that reaches every kernel family (one tool per gpurun call)
# usage: r2_sanitize.sh memcheck|racecheck
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tool=$1
SEL="tests/test_gpu_parity.py::test_golden_small tests/test_gpu_parity.py::test_half_kernel tests/test_gpu_parity.py::test_ratio_stream_kernel tests/test_gpu_parity.py::test_area_2to1_streaming_kernel tests/test_gpu_parity.py::test_linear_integer_upsampling_kernel tests/test_gpu_parity.py::test_packed_kernel_wide_source_window tests/test_gpu_parity.py::test_yuv420_frames tests/test_gpu_mma.py::test_mma_kernel_matches_oracle tests/test_gpu_bands.py::test_stream_bands_every_split"
DESEL="--deselect tests/test_gpu_parity.py::test_linear_integer_upsampling_kernel[case0] --deselect tests/test_gpu_mma.py::test_mma_kernel_matches_oracle[case2] --deselect tests/test_gpu_parity.py::test_area_2to1_streaming_kernel[case0]"
timeout 2400 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest $SEL $DESEL -q -x -m gpu -p no:cacheprovider > gpurun_out/r2_sanitizer_$tool.log 2>&1
echo "sanitizer $tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -20
tail -5 gpurun_out/r2_sanitizer_$tool.log
