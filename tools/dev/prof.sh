# usage: prof.sh <tag> <kernel regex> [bench args...]
tag=$1; kr=$2; shift 2
python bench.py --frames 1024 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$kr -s 3 -c 1 -f -o gpurun_out/${tag} python bench.py --frames 1024 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv
tail -2 gpurun_out/${tag}_ncu.log
