"""Randomized parity of wide, short images (many column strips, few rows; ratios near the kernels' limits).
usage: fuzz_wide.py [seed] [seconds]; FUZZ_STREAM=1 / FUZZ_MMA=1 force those kernel families."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libiqo_b200 as iqo
import fuzz_lib
from oracle_lib import AREA, LANCZOS, LINEAR

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
PATH = iqo.PATH_STREAM if os.environ.get("FUZZ_STREAM") else iqo.PATH_MMA if os.environ.get("FUZZ_MMA") else iqo.PATH_AUTO
t0, stats, bad = time.time(), {}, 0
while time.time() - t0 < budget:
    kind = int(rng.choice([LANCZOS, LANCZOS, LANCZOS, AREA, LINEAR]))
    sw, sh = int(rng.randint(700, 6000)), int(rng.randint(16, 120))
    if rng.rand() < 0.5:
        sw = (sw + 15) & ~15
    if kind == LINEAR:
        dw, dh = int(rng.randint(sw // 2, 3 * sw)), int(rng.randint(max(8, sh // 2), 3 * sh))
    elif kind == AREA:
        dw, dh = int(rng.randint(max(8, sw // 6), sw + 1)), int(rng.randint(max(4, sh // 6), sh + 1))
    else:
        dw, dh = int(rng.randint(max(8, sw // 5), 2 * sw)), int(rng.randint(max(8, sh // 4), 2 * sh))
    case = (kind, int(rng.choice([1, 2, 3, 4, 5])) if kind == LANCZOS else 0, int(rng.choice([1, 1, 1, 2, 3])) if kind == LANCZOS else 1,
            sw, sh, dw, dh, int(rng.choice([0, 0, 16, 4])), int(rng.choice([0, 0, 8, 3])))
    res = fuzz_lib.run_single(rng, case, PATH)
    if res is None:
        continue
    stats[res[0]] = stats.get(res[0], 0) + 1
    if not res[1]:
        bad += 1
        print("MISMATCH", case, res[0], flush=True)
print("cases per kernel:", stats, "mismatching cases:", bad)
