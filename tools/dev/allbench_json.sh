# Full bench.py JSON line of every BASELINE workload (device resident + e2e), then cfg3 frames and cfg5.
# Output: gpurun_out/all_workloads.jsonl, gpurun_out/cfg3_frames.txt, gpurun_out/cfg5_gigapixel.txt
mkdir -p gpurun_out
: > gpurun_out/all_workloads.jsonl
for w in cfg4_lanczos3_1080p_to_540p cfg3y_lanczos2_2160p_to_1080p cfg3uv_lanczos2_px2_1080p_to_540p cfg1_lanczos3_1080p_to_720p cfg2a_area_2160p_to_1080p cfg2b_linear_720p_to_2160p cfg5s_lanczos4_8192_to_3000; do
  timeout 120 python bench.py --workload $w --no-cpu-baseline --steps 10 --e2e-frames 256 2>/dev/null >> gpurun_out/all_workloads.jsonl
done
timeout 120 python tools/dev/cfg3.py > gpurun_out/cfg3_frames.txt 2>&1
timeout 200 python tools/gigapixel.py > gpurun_out/cfg5_gigapixel.txt 2>&1
