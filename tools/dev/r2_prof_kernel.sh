# ncu --set full of one kernel under bench.py: usage r2_prof_kernel.sh <tag> <kernel regex> <workload> <frames> <dst px per launch>
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tag=$1; rx=$2; wl=$3; fr=$4; px=$5
CMD="python bench.py --workload $wl --frames $fr --no-e2e --no-cpu-baseline --no-extras --steps 2 --warmup 3"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/$tag $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv
python tools/ncu_summary.py gpurun_out/${tag}_raw.csv gpurun_out/${tag}_src.csv $px > gpurun_out/${tag}_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/${tag}_src.csv $px 0.3 >> gpurun_out/${tag}_summary.txt 2>&1
python tools/dev/stalls.py gpurun_out/${tag}_raw.csv >> gpurun_out/${tag}_summary.txt 2>&1
