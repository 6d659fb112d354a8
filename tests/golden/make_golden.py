#!/usr/bin/env python
"""Generate tests/golden/*.json|npz by RUNNING THE REFERENCE's own Generic implementation
(oracle/_ref/libiqo_ref_generic.so, compiled from /root/reference by oracle/Makefile).

Run in the development container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

Inputs are reproducible from the case description alone: LCG bytes (SURVEY 8c:
x=seed; x = x*1664525+1013904223 mod 2^32; byte = x>>24) laid out with the given
source stride.  Small cases store the full destination image; large cases store
the FNV-1a-64 of the destination.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import LANCZOS, AREA, LINEAR, lcg_image, fnv1a, ref_resize, ref_generic  # noqa: E402

# (kind, degree, pxScale, srcW, srcH, dstW, dstH, srcStridePad, dstStridePad, seed)
SMALL = [
    (LANCZOS, 3, 1, 64, 48, 40, 30, 0, 0, 1),
    (LANCZOS, 3, 1, 40, 30, 64, 48, 0, 0, 1),
    (LANCZOS, 2, 2, 64, 48, 32, 24, 0, 0, 1),
    (LANCZOS, 3, 1, 96, 54, 48, 27, 0, 0, 2),      # cfg4 shape / 20
    (LANCZOS, 2, 1, 96, 54, 48, 27, 2, 1, 3),      # cfg3-Y ratio, padded strides
    (LANCZOS, 2, 2, 48, 27, 24, 13, 1, 3, 4),
    (LANCZOS, 3, 1, 96, 54, 64, 36, 0, 0, 5),      # cfg1 ratio 3:2
    (LANCZOS, 4, 1, 128, 128, 47, 47, 0, 0, 6),    # coprime, many phases
    (LANCZOS, 1, 1, 50, 40, 25, 20, 0, 0, 7),
    (LANCZOS, 5, 1, 90, 70, 33, 21, 3, 0, 8),
    (LANCZOS, 9, 1, 120, 110, 60, 55, 0, 0, 9),
    (LANCZOS, 3, 1, 64, 48, 64, 30, 0, 0, 10),     # X pass-through
    (LANCZOS, 3, 1, 64, 48, 40, 48, 0, 0, 11),     # Y pass-through
    (LANCZOS, 3, 1, 31, 17, 31, 17, 1, 1, 12),     # identity
    (LANCZOS, 2, 1, 33, 21, 64, 47, 0, 0, 13),     # up, coprime
    (LANCZOS, 3, 1, 64, 30, 40, 48, 0, 0, 14),     # down in X, up in Y
    (LANCZOS, 4, 2, 100, 80, 50, 40, 0, 0, 15),
    (AREA, 0, 1, 64, 48, 40, 30, 0, 0, 1),
    (AREA, 0, 1, 96, 54, 48, 27, 0, 0, 2),         # cfg2a ratio 2:1
    (AREA, 0, 1, 100, 75, 33, 21, 2, 2, 3),
    (AREA, 0, 1, 64, 48, 64, 30, 0, 0, 4),
    (AREA, 0, 1, 64, 48, 40, 48, 0, 0, 5),
    (AREA, 0, 1, 40, 30, 64, 48, 0, 0, 6),         # "up" = nearest-floor
    (AREA, 0, 1, 77, 31, 5, 3, 0, 0, 7),
    (AREA, 0, 1, 29, 23, 29, 23, 0, 0, 8),
    (LINEAR, 0, 1, 40, 30, 100, 75, 0, 0, 1),
    (LINEAR, 0, 1, 32, 18, 96, 54, 0, 0, 2),       # cfg2b ratio 3x
    (LINEAR, 0, 1, 32, 18, 64, 36, 1, 2, 3),
    (LINEAR, 0, 1, 33, 21, 64, 47, 0, 0, 4),
    (LINEAR, 0, 1, 32, 18, 32, 54, 0, 0, 5),
    (LINEAR, 0, 1, 32, 18, 96, 18, 0, 0, 6),
    (LINEAR, 0, 1, 17, 9, 17, 9, 0, 0, 7),
    (LINEAR, 0, 1, 2, 2, 5, 6, 0, 0, 8),
]

# large cases: hashes only.  The first seven are SURVEY 8c's table (re-generated here and
# asserted equal to the values the survey recorded from the compiled reference).
LARGE = [
    (LANCZOS, 3, 1, 1920, 1080, 1280, 720, 0, 0, 1, "e8c30c28af8e71ba"),
    (AREA, 0, 1, 3840, 2160, 1920, 1080, 0, 0, 1, "df58152b2d7489f7"),
    (LINEAR, 0, 1, 1280, 720, 3840, 2160, 0, 0, 1, "aa300e10b8169c86"),
    (LANCZOS, 2, 1, 3840, 2160, 1920, 1080, 0, 0, 1, "b341a0860be0b77c"),
    (LANCZOS, 2, 2, 1920, 1080, 960, 540, 0, 0, 1, "2ae949c81e9c756c"),
    (LANCZOS, 3, 1, 1920, 1080, 960, 540, 0, 0, 1, "bc3ae031361c0774"),
    (LANCZOS, 4, 1, 2048, 2048, 750, 750, 0, 0, 1, "b50ce2ac904078d5"),
    (LANCZOS, 3, 1, 1920, 1080, 960, 540, 0, 0, 7, None),
    (LANCZOS, 3, 1, 1921, 1081, 960, 540, 3, 5, 2, None),
    (LANCZOS, 2, 1, 1280, 720, 1920, 1080, 0, 0, 3, None),
    (LANCZOS, 4, 1, 4096, 1024, 1500, 375, 0, 0, 4, None),   # cfg5 ratio 1024:375
    (AREA, 0, 1, 1920, 1080, 1280, 720, 0, 0, 5, None),
    (AREA, 0, 1, 1000, 1000, 333, 777, 8, 3, 6, None),
    (LINEAR, 0, 1, 960, 540, 1920, 1080, 0, 0, 7, None),
    (LINEAR, 0, 1, 640, 480, 1600, 1000, 0, 1, 8, None),
    (LANCZOS, 4, 1, 8192, 8192, 3000, 3000, 0, 0, 1, None),  # bench.py's cfg5s workload (cfg5's ratio, batch-sized)
    (AREA, 0, 1, 1920, 1080, 1280, 720, 0, 0, 1, None),      # bench.py: Area at a general ratio (3:2)
    (LINEAR, 0, 1, 1280, 720, 1920, 1080, 0, 0, 1, None),    # bench.py: Linear at a non-integer ratio (1.5x)
]

# BASELINE config 5 at full size (1 GiB source): hash only, kept out of LARGE so that the per-case tests do
# not iterate over it (tests/test_oracle.py and tests/test_gpu_bands.py have dedicated full-size tests)
HUGE = [
    (LANCZOS, 4, 1, 32768, 32768, 12000, 12000, 0, 0, 1, "0ec3dba9ab1194ca"),
]


def case_src(c):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = c[:10]
    return lcg_image(sh, sw + spad, seed=seed)


def main():
    assert ref_generic() is not None, "build oracle/_ref first (make -C oracle ref)"
    small_out = {}
    small_meta = []
    for i, c in enumerate(SMALL):
        kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = c
        rc, dst = ref_resize(kind, case_src(c), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert rc == 0
        small_out["dst%d" % i] = dst[:, :dw].copy()
        small_meta.append(list(c))
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small_out)
    large_meta = []
    for c in LARGE:
        kind, deg, px, sw, sh, dw, dh, spad, dpad, seed, expect = c
        rc, dst = ref_resize(kind, case_src(c), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert rc == 0
        h = "%016x" % fnv1a(dst, dw)
        if expect is not None:
            assert h == expect, (c, h)
        large_meta.append(list(c[:10]) + [h])
        print(c[:10], h)
    huge_meta = []
    for c in HUGE:
        kind, deg, px, sw, sh, dw, dh, spad, dpad, seed, expect = c
        rc, dst = ref_resize(kind, case_src(c), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert rc == 0
        h = "%016x" % fnv1a(dst, dw)
        assert h == expect, (c, h)
        huge_meta.append(list(c[:10]) + [h])
        print(c[:10], h)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (reference Generic via oracle/_ref)",
                   "fields": ["kind", "degree", "pxScale", "srcW", "srcH", "dstW", "dstH",
                              "srcStridePad", "dstStridePad", "seed", "(large only) fnv1a64 of dst"],
                   "small": small_meta, "large": large_meta, "huge": huge_meta}, f, indent=1)


if __name__ == "__main__":
    main()
