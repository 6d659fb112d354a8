"""One-off randomized parity fuzz of the specialised kernels against the oracle (not part of the test suite)."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libiqo_b200 as iqo
from oracle_lib import oracle_resize, lcg_image, LANCZOS, AREA, LINEAR

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
VERBOSE = len(sys.argv) > 3
PATH = iqo.PATH_STREAM if os.environ.get("FUZZ_STREAM") else iqo.PATH_AUTO  # FUZZ_STREAM=1: streaming kernels on small launches too
t0 = time.time()
stats = {}
bad = 0


def run(kind, deg, px, sw, sh, dw, dh, spad=0, dpad=0):
    global bad
    src = lcg_image(sh, sw + spad, seed=int(rng.randint(1, 1 << 20)))
    if rng.rand() < 0.15:
        src[:] = rng.choice([0, 255])
    rc, want = oracle_resize(kind, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    if rc != 0:
        return
    dst = np.full((dh, dw + dpad), 0xA5, dtype=np.uint8)
    if VERBOSE:
        print("case", kind, deg, px, sw, sh, dw, dh, spad, dpad, iqo.plan_kernel(kind, deg, sw, sh, dw, dh, px)[0], flush=True)
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(PATH)
        r.resize(sw + spad, src, dw + dpad, dst)
        k = r.last_kernel()
    stats[k] = stats.get(k, 0) + 1
    if not np.array_equal(dst, want):
        bad += 1
        d = np.argwhere(dst != want)
        print("MISMATCH", kind, deg, px, sw, sh, dw, dh, spad, dpad, k, len(d), d[:4].tolist())


while time.time() - t0 < budget:
    fam = rng.randint(0, 6)
    pad = lambda: int(rng.choice([0, 0, 4, 8, 16, 3]))
    if fam == 0:      # 2:1 Lanczos
        dw, dh = int(rng.randint(2, 400)) * 2, int(rng.randint(8, 300))
        run(LANCZOS, int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 1, 2])), 2 * dw, 2 * dh, dw, dh, pad(), pad())
    elif fam == 1:    # ratio kernel families
        rs, rd = [(3, 2), (1, 2), (3, 4), (2, 1)][rng.randint(0, 4)]
        k = int(rng.randint(1, 60)) * 8
        sh = int(rng.randint(16, 500))
        dh = int(rng.randint(max(8, sh // 3), 2 * sh))
        run(LANCZOS, int(rng.choice([1, 2, 3, 4])), 1, rs * k, sh, rd * k, dh, pad(), pad())
    elif fam == 2:    # linear up-sampling
        kx = int(rng.choice([2, 3]))
        sw, sh = int(rng.randint(2, 200)) * 4, int(rng.randint(4, 200))
        dh = int(rng.randint(sh, 3 * sh + 1))
        run(LINEAR, 0, 1, sw, sh, kx * sw, dh, pad(), pad())
    elif fam == 3:    # area 2:1
        dw, dh = int(rng.randint(1, 200)) * 8, int(rng.randint(4, 300))
        run(AREA, 0, 1, 2 * dw, 2 * dh, dw, dh, pad(), pad())
    elif fam == 4:    # arbitrary Lanczos
        sw, sh = int(rng.randint(24, 700)), int(rng.randint(24, 500))
        dw, dh = int(rng.randint(12, 700)), int(rng.randint(12, 500))
        run(LANCZOS, int(rng.choice([1, 2, 3])), 1, sw, sh, dw, dh, pad(), pad())
    else:             # arbitrary area / linear
        sw, sh = int(rng.randint(8, 500)), int(rng.randint(8, 400))
        if rng.rand() < 0.5:
            run(AREA, 0, 1, sw, sh, int(rng.randint(4, sw + 1)), int(rng.randint(4, sh + 1)), pad(), pad())
        else:
            run(LINEAR, 0, 1, sw, sh, int(rng.randint(sw, 3 * sw + 1)), int(rng.randint(sh, 3 * sh + 1)), pad(), pad())
print("cases per kernel:", stats, "mismatching cases:", bad)
