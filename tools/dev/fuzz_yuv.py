"""Open-ended randomized parity fuzz of the YUV420 frame API (device and host frames) against the per-plane
oracle (bounded form: tests/test_gpu_fuzz.py).  usage: fuzz_yuv.py [seed] [seconds]"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import fuzz_lib

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
t0, cases, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    case = fuzz_lib.yuv_case(rng)
    res = fuzz_lib.run_yuv(rng, case)
    if res is None:
        continue
    cases += 1
    if not res[1]:
        bad += 1
        print("MISMATCH", case)
print("yuv cases:", cases, "mismatching:", bad)
