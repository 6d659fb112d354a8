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
        assert np.array_equal(vectors.lcg_bytes_numpy(h * w, seed).reshape(h, w), lcg_image(h, w, seed))
    img = lcg_image(270, 480, 5)
    assert vectors.fnv1a64(img) == fnv1a(img) == vectors.fnv1a64_python(img)
    assert vectors.fnv1a64(np.zeros((0, 4), dtype=np.uint8)) == 0xCBF29CE484222325
    assert vectors.lcg_bytes(0).size == 0
    # band-wise use (bench.py's cfg5 leg): rows [row0, row0 + h) of a taller image, hash chained band by band
    whole = lcg_image(90, 64, 7)
    for row0, h in ((0, 90), (1, 5), (37, 53), (89, 1)):
        assert np.array_equal(vectors.lcg_image(h, 64, 7, row0=row0), whole[row0:row0 + h])
        assert np.array_equal(vectors.lcg_bytes_numpy(h * 64, 7, offset=row0 * 64).reshape(h, 64), whole[row0:row0 + h])
    chained = vectors.fnv1a64(whole[40:], vectors.fnv1a64(whole[:40]))
    assert chained == fnv1a(whole)
    assert vectors._native() is not None, "libiqo_b200/lib/libiqo_vectors.so was not built (make -C libiqo_b200/csrc)"


def test_every_bench_workload_has_a_golden_hash():
    for name, (kind, deg, px, sw, sh, dw, dh, _) in bench.WORKLOADS.items():
        assert vectors.golden_hash(kind, deg, px, sw, sh, dw, dh) is not None, name
    assert vectors.golden_hash(LANCZOS, 3, 1, 1920, 1080, 960, 540) == "bc3ae031361c0774"
    assert vectors.golden_hash(*bench.CFG5) == "0ec3dba9ab1194ca"
    assert vectors.golden_hash(LANCZOS, 3, 1, 100, 100, 50, 50) is None


def test_both_arms_emit_the_same_config():
    a = bench.workload_config(bench.DEFAULT_WORKLOAD, 4096)
    assert a == bench.workload_config(bench.DEFAULT_WORKLOAD, 4096)
    assert a["workload"] == bench.DEFAULT_WORKLOAD and a["src"] == [1920, 1080] and a["dst"] == [960, 540]


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
    nogolden = (0, 3, 1, 100, 100, 50, 50, 1)
    assert bench.plant_golden_frames(torch.zeros((1, 100, 100), dtype=torch.uint8), nogolden) is None
