"""Open-ended randomized parity fuzz of device-resident batches with random pitches / frame strides / base
offsets (bounded form: tests/test_gpu_fuzz.py).  usage: fuzz_batch.py [seed] [seconds]"""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libiqo_b200 as iqo
import fuzz_lib

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
PATH = iqo.PATH_STREAM if os.environ.get("FUZZ_STREAM") else iqo.PATH_MMA if os.environ.get("FUZZ_MMA") else iqo.PATH_AUTO
t0, stats, bad = time.time(), {}, 0
while time.time() - t0 < budget:
    case = fuzz_lib.batch_case(rng)
    res = fuzz_lib.run_batch(rng, case, PATH)
    if res is None:
        continue
    stats[res[0]] = stats.get(res[0], 0) + 1
    if not res[1]:
        bad += 1
        print("MISMATCH", case, res[0])
print("cases per kernel:", stats, "mismatching cases:", bad)
