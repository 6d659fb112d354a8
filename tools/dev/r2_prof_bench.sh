# launch list of the default bench command + one full capture of the bench kernel (half_sym_stream, TMA-fed FIFO)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2_bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $CMD > gpurun_out/r2_bench_ncu_list.log 2>&1
echo "list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD2 > gpurun_out/r2_bench_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:resizeHalfStream -s 3 -c 1 -f -o gpurun_out/r2_cfg4_half_stream_tma $CMD2 > gpurun_out/r2_bench_ncu_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_cfg4_half_stream_tma.ncu-rep --page raw --csv > gpurun_out/r2_cfg4_half_stream_tma_raw.csv
ncu -i gpurun_out/r2_cfg4_half_stream_tma.ncu-rep --page source --csv > gpurun_out/r2_cfg4_half_stream_tma_src.csv
python tools/ncu_summary.py gpurun_out/r2_cfg4_half_stream_tma_raw.csv gpurun_out/r2_cfg4_half_stream_tma_src.csv $((4096*960*540)) > gpurun_out/r2_cfg4_half_stream_tma_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2_cfg4_half_stream_tma_src.csv $((4096*960*540)) 0.3 >> gpurun_out/r2_cfg4_half_stream_tma_summary.txt 2>&1
python tools/dev/stalls.py gpurun_out/r2_cfg4_half_stream_tma_raw.csv >> gpurun_out/r2_cfg4_half_stream_tma_summary.txt 2>&1
python tools/ncu_launch_list.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launch_list.txt 2>&1
head -45 gpurun_out/r2_cfg4_half_stream_tma_summary.txt; cat gpurun_out/r2_bench_launch_list.txt | head -40
