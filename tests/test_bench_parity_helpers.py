"""CPU tests of the parity leg of bench.py's CUDA arm: it plants the SURVEY 8c LCG frame in the timed batch and
checks the frames the kernel wrote against the committed hash of the reference's Generic output
(tests/golden/cases.json) without touching oracle/.  Here the oracle plays the kernel."""
import os
import sys

import numpy as np
import torch

from oracle_lib import LANCZOS, fnv1a, lcg_image, oracle_resize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from libiqo_b200 import vectors  # noqa: E402


def test_vectors_match_the_oracle_helpers():
    for h, w, seed in ((1080, 1920, 1), (7, 5, 9), (1, 1, 3), (30, 64, 77), (33, 1, 2)):
        assert np.array_equal(vectors.lcg_image(h, w, seed), lcg_image(h, w, seed))
    img = lcg_image(270, 480, 5)
    assert vectors.fnv1a64(img) == fnv1a(img)
    assert vectors.fnv1a64(np.zeros((0, 4), dtype=np.uint8)) == 0xCBF29CE484222325
    assert vectors.lcg_bytes(0).size == 0


def test_every_bench_workload_but_the_cfg5_stand_in_has_a_golden_hash():
    for name, (kind, deg, px, sw, sh, dw, dh, _) in bench.WORKLOADS.items():
        h = vectors.golden_hash(kind, deg, px, sw, sh, dw, dh)
        assert (h is None) == name.startswith("cfg5s"), name
    assert vectors.golden_hash(LANCZOS, 3, 1, 1920, 1080, 960, 540) == "bc3ae031361c0774"


def test_planted_frames_are_checked_against_the_golden_hash():
    work = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    kind, deg, px, sw, sh, dw, dh, _ = work
    for frames in (1, 2, 5):
        src = torch.randint(0, 256, (frames, sh, sw), dtype=torch.uint8)
        want = bench.plant_golden_frames(src, work)
        assert want == "bc3ae031361c0774"
        pos = bench.golden_positions(frames)
        assert pos[0] == 0 and pos[-1] == frames - 1
        dst = torch.zeros((frames, dh, dw), dtype=torch.uint8)
        for f in pos:   # the oracle stands in for the kernel
            rc, out = oracle_resize(kind, src[f].numpy(), dw, dh, deg, px)
            assert rc == 0
            dst[f] = torch.from_numpy(out)
        p = bench.check_golden_frames(dst, want)
        assert p["bit_exact"] and p["mismatches"] == 0 and p["max_abs_diff"] == 0 and p["frames_checked"] == len(pos)
        dst[pos[-1], dh - 1, dw - 1] ^= 1   # one wrong LSB anywhere must be seen
        p = bench.check_golden_frames(dst, want)
        assert not p["bit_exact"] and p["mismatches"] is None
    assert bench.plant_golden_frames(torch.zeros((0, sh, sw), dtype=torch.uint8), work) is None
    cfg5s = bench.WORKLOADS["cfg5s_lanczos4_8192_to_3000"]
    assert bench.plant_golden_frames(torch.zeros((1, 8, 8), dtype=torch.uint8), cfg5s) is None
