# longer fuzz campaign over the dispatch paths (seeds differ from the pytest ones)
cd "$GRAFT_REPO_ROOT"
for seed in 11 12 13; do timeout 400 python tools/dev/fuzz.py $seed ${1:-60} 2>&1 | tail -3; done
FUZZ_STREAM=1 timeout 400 python tools/dev/fuzz.py 21 ${1:-60} 2>&1 | tail -3
FUZZ_MMA=1 timeout 400 python tools/dev/fuzz.py 31 ${1:-60} 2>&1 | tail -3
