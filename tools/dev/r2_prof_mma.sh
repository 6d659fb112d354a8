# one ncu --set full capture of the tensor-path kernel on cfg4 (512 frames)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD="env IQO_CUDA_MMA_WCOLS=${WCOLS:-144} python bench.py --frames 512 --no-e2e --no-cpu-baseline --no-extras --steps 2 --warmup 3 --path mma"
$CMD > gpurun_out/r2_mma_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:resizeLanczosMma -s 3 -c 1 -f -o gpurun_out/r2_mma_cfg4 $CMD > gpurun_out/r2_mma_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_mma_ncu.log
ncu -i gpurun_out/r2_mma_cfg4.ncu-rep --page raw --csv > gpurun_out/r2_mma_cfg4_raw.csv
ncu -i gpurun_out/r2_mma_cfg4.ncu-rep --page source --csv > gpurun_out/r2_mma_cfg4_src.csv
python tools/ncu_summary.py gpurun_out/r2_mma_cfg4_raw.csv gpurun_out/r2_mma_cfg4_src.csv $((512*960*540)) > gpurun_out/r2_mma_cfg4_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2_mma_cfg4_src.csv $((512*960*540)) 0.1 >> gpurun_out/r2_mma_cfg4_summary.txt 2>&1
cat gpurun_out/r2_mma_cfg4_summary.txt
