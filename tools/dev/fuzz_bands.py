"""Open-ended randomized parity fuzz of row bands (iqo_cuda_resize_band with srcRow0 != 0): random shapes of every
kind, random ragged band partitions, each band from a buffer that holds only the band + halo; the assembled image is
compared with the oracle's whole-image result (bounded form: tests/test_gpu_fuzz.py).
usage: fuzz_bands.py [seed] [seconds]; FUZZ_STREAM=1 / FUZZ_MMA=1 force those kernel families."""
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
    res = fuzz_lib.run_band_case(rng, PATH)
    if res is None:
        continue
    stats[res[0]] = stats.get(res[0], 0) + 1
    if not res[1]:
        bad += 1
        print("MISMATCH", res[2:], res[0], flush=True)
print("cases per kernel (last band's):", stats, "mismatching cases:", bad)
