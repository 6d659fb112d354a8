# ncu --set full of the 3:2 streaming kernel on cfg1
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tag=r2_ratio_cfg1; px=$((1280*720*256))
CMD="python bench.py --workload cfg1_lanczos3_1080p_to_720p --frames 256 --no-e2e --no-cpu-baseline --no-extras --steps 2 --warmup 3 --path no_mma"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:resizeRatioStream -s 3 -c 1 -f -o gpurun_out/$tag $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv
python tools/ncu_summary.py gpurun_out/${tag}_raw.csv gpurun_out/${tag}_src.csv $px > gpurun_out/${tag}_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/${tag}_src.csv $px 0.3 >> gpurun_out/${tag}_summary.txt 2>&1
python tools/dev/stalls.py gpurun_out/${tag}_raw.csv >> gpurun_out/${tag}_summary.txt 2>&1
tail -5 gpurun_out/${tag}_plain.log | cut -c1-300
