"""CPU tests: pin the oracle (oracle/iqo_oracle.c) against the reference.

(a) golden fixtures generated from the reference's own Generic code (tests/golden/),
(b) the FNV-1a hashes SURVEY.md 8c recorded from the compiled reference,
(c) when oracle/_ref is present (dev container and GPU box): a bit-for-bit sweep against it,
(d) known-answer tests from SURVEY.md section 4 (constant images, identity, coefficient KATs).
"""
import random

import numpy as np
import pytest

import golden_cases as G
from oracle_lib import (AREA, LANCZOS, LINEAR, fnv1a, lcg_image, oracle_resize, oracle_table,
                        ref_generic, ref_resize)


@pytest.mark.parametrize("i", range(len(G.SMALL)), ids=[G.case_id(c) for c in G.SMALL])
def test_oracle_matches_golden_small(i):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = G.SMALL[i]
    rc, dst = oracle_resize(kind, G.case_src(G.SMALL[i]), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    assert np.array_equal(dst[:, :dw], G.small_expected(i))
    # bytes between dstW and dstStride are never written (reference: ..._Generic.cpp:610)
    assert (dst[:, dw:] == 0xA5).all()


@pytest.mark.parametrize("c", G.LARGE, ids=[G.case_id(c) for c in G.LARGE])
def test_oracle_matches_golden_hash(c):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed, h = c
    rc, dst = oracle_resize(kind, G.case_src(c), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    assert "%016x" % fnv1a(dst, dw) == h


def test_oracle_matches_the_gigapixel_hash():
    """BASELINE config 5 at full size: 32768 x 32768 -> 12000 x 12000 Lanczos4 (22 taps, 375 phases per axis).
    SURVEY 8c recorded source hash 4c81a575b77c7ae5 and destination hash 0ec3dba9ab1194ca from the compiled
    reference; tools/gigapixel.py checks the GPU row bands against the same value.  ~15 s, 1.2 GB."""
    src = lcg_image(32768, 32768, seed=1)
    assert "%016x" % fnv1a(src) == "4c81a575b77c7ae5"
    rc, dst = oracle_resize(LANCZOS, src, 12000, 12000, 4, 1)
    assert rc == 0
    assert dst[0, :4].tolist() == [125, 123, 132, 113]
    assert "%016x" % fnv1a(dst) == "0ec3dba9ab1194ca"


@pytest.mark.parametrize("case", [(LANCZOS, 3, 1920, 1080, 1280, 720, "7625cae131094002"),
                                  (LANCZOS, 3, 1920, 1080, 960, 540, "9ccaee019da3eea2"),
                                  (AREA, 0, 3840, 2160, 1920, 1080, "d2e341d3c900d384"),
                                  (LINEAR, 0, 1280, 720, 3840, 2160, "3d4846274d6ae615")])
def test_oracle_matches_the_benchmark_fill_hashes(case):
    """SURVEY 8c also recorded the reference's Generic output for the source image its own benchmark makes
    (benchmark/benchmark.cpp:51-59: std::mt19937(0) + uniform_int_distribution<int>(0, 255), which libstdc++
    evaluates as the generator's 32-bit output >> 24).  numpy's legacy RandomState(0) is the same init_genrand(0)
    stream, so the input is reproduced here without the reference."""
    kind, deg, sw, sh, dw, dh, want = case
    raw = np.random.RandomState(0).randint(0, 2 ** 32, size=sh * sw, dtype=np.uint64)
    src = (raw >> 24).astype(np.uint8).reshape(sh, sw)
    rc, dst = oracle_resize(kind, src, dw, dh, deg, 1)
    assert rc == 0
    assert "%016x" % fnv1a(dst) == want


def test_coefficient_kats():
    # cfg1 (3:2) phases 0/1, cfg4 and cfg3 tables
    tx = oracle_table(LANCZOS, 0, 1920, 1280, 3)
    ty = oracle_table(LANCZOS, 1, 1080, 720, 3)
    assert tx.tolist() == [[266, -465, -1149, 6660, 10407, 1837, -1480, 272, 36, 0],
                           [36, 272, -1480, 1837, 10407, 6660, -1149, -465, 266, 0]]
    assert ty.tolist() == [[1, -2, -4, 26, 41, 7, -6, 1, 0, 0], [0, 1, -6, 7, 41, 26, -4, -2, 1, 0]]
    assert oracle_table(LANCZOS, 1, 1080, 540, 3).tolist() == [[0, 1, -2, -4, 9, 28, 28, 9, -4, -2, 1, 0]]
    assert oracle_table(LANCZOS, 0, 1920, 960, 3).tolist() == [
        [60, 247, -557, -1092, 2220, 7314, 7314, 2220, -1092, -557, 247, 60]]
    assert oracle_table(LANCZOS, 1, 2160, 1080, 2).tolist() == [[-1, -3, 7, 29, 29, 7, -3, -1]]
    assert oracle_table(LANCZOS, 0, 3840, 1920, 2).tolist() == [[-145, -687, 1909, 7115, 7115, 1909, -687, -145]]
    assert oracle_table(LANCZOS, 1, 1080, 540, 2, 2).tolist() == [[0, -4, 34, 34]]
    assert oracle_table(LANCZOS, 0, 1920, 960, 2, 2).tolist() == [[0, -964, 8674, 8674]]
    assert oracle_table(AREA, 1, 2160, 1080).tolist() == [[128, 128]]
    assert oracle_table(AREA, 0, 3840, 1920).tolist() == [[16384, 16384]]
    assert oracle_table(LINEAR, 1, 720, 2160).tolist() == [[85, 171], [256, 0], [171, 85]]
    assert oracle_table(LINEAR, 0, 1280, 3840).tolist() == [[10923, 21845], [32768, 0], [21845, 10923]]
    # every row sums to the fixed-point one
    for t, one in ((tx, 16384), (ty, 64), (oracle_table(LANCZOS, 0, 32768, 12000, 4), 16384),
                   (oracle_table(AREA, 0, 1000, 333), 32768)):
        assert (t.sum(axis=1) == one).all()
    assert oracle_table(LANCZOS, 0, 32768, 12000, 4).shape == (375, 22)


@pytest.mark.parametrize("v", [0, 1, 100, 254, 255])
def test_constant_image(v):
    src = np.full((54, 96), v, dtype=np.uint8)
    for kind, deg, dw, dh in ((AREA, 0, 48, 27), (AREA, 0, 64, 36), (LINEAR, 0, 192, 108), (LINEAR, 0, 240, 135)):
        rc, dst = oracle_resize(kind, src, dw, dh, deg)
        assert rc == 0 and (dst == v).all()
    # Lanczos: interior columns and all rows are exact; X-border columns may be v+1
    for deg, px, dw, dh in ((3, 1, 48, 27), (2, 1, 64, 36), (2, 2, 48, 27), (3, 1, 192, 108)):
        rc, dst = oracle_resize(LANCZOS, src, dw, dh, deg, px)
        assert rc == 0
        d = dst.astype(int) - v
        assert d.min() >= 0 and d.max() <= 1
        assert (d[:, 8:-8] == 0).all()


def test_constant_image_xborder_quirk():
    # Lanczos2 px=2 192x108 -> 96x54 gives v+1 in the right-most column
    src = np.full((108, 192), 100, dtype=np.uint8)
    rc, dst = oracle_resize(LANCZOS, src, 96, 54, 2, 2)
    assert rc == 0
    assert (dst[:, :-1] == 100).all() and (dst[:, -1] == 101).all()


def test_identity_and_rejects():
    src = lcg_image(17, 31)
    for kind, deg in ((LANCZOS, 3), (AREA, 0), (LINEAR, 0)):
        rc, dst = oracle_resize(kind, src, 31, 17, deg)
        assert rc == 0 and np.array_equal(dst, src)
    # image smaller than the kernel: the reference desynchronises its iterators
    rc, _ = oracle_resize(LANCZOS, lcg_image(8, 8), 5, 5, 3)
    assert rc == -2
    # all in-range taps quantise to 0 on a border row: reference divides by zero
    rc, _ = oracle_resize(LANCZOS, lcg_image(100, 100), 99, 99, 1)
    assert rc == -3
    rc, _ = oracle_resize(LANCZOS, lcg_image(4, 4), 0, 4, 3)
    assert rc == -1


@pytest.mark.skipif(ref_generic() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_sweep():
    rng = random.Random(1234)
    checked = 0
    for _ in range(1500):
        kind = rng.choice([LANCZOS, LANCZOS, AREA, LINEAR])
        sw, sh = rng.randint(1, 100), rng.randint(1, 100)
        if kind == LINEAR:
            dw, dh = rng.randint(sw, 3 * sw), rng.randint(sh, 3 * sh)
        elif kind == AREA:
            dw, dh = rng.randint(1, sw + 8), rng.randint(1, sh + 8)
        else:
            dw, dh = rng.randint(1, 160), rng.randint(1, 160)
        if rng.random() < 0.15:
            dw = sw
        if rng.random() < 0.15:
            dh = sh
        deg = rng.randint(1, 9) if kind == LANCZOS else 0
        px = rng.choice([1, 1, 2, 3]) if kind == LANCZOS else 1
        spad = rng.randint(0, 3)
        src = lcg_image(sh, sw + spad, seed=rng.randint(1, 1 << 30))
        rc, dst = oracle_resize(kind, src, dw, dh, deg, px, sw=sw)
        if rc != 0:
            continue  # reference behaviour undefined there
        rc2, ref = ref_resize(kind, src, dw, dh, deg, px, sw=sw)
        assert rc2 == 0
        assert np.array_equal(dst, ref), (kind, deg, px, sw, sh, dw, dh)
        checked += 1
    assert checked > 1000


@pytest.mark.skipif(ref_generic() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_narrow_sources_all_border_columns():
    """Source narrower than the horizontal kernel: mainBegin > mainEnd on X.  The reference is well defined there
    (resizeXborder re-seeds its iterator, src/IQOLanczosResizerImpl_Generic.cpp:547-549) as long as
    mainBegin <= dstW: every column takes the border formula.  The same condition on Y desynchronises the
    reference's iterators and stays rejected (-2)."""
    rng = random.Random(77)
    defined = 0
    for _ in range(400):
        deg = rng.randint(1, 6)
        sw, sh = rng.randint(2, 14), rng.randint(30, 60)
        dw, dh = rng.randint(1, sw), rng.choice([sh, rng.randint(20, 80)])
        src = lcg_image(sh, sw, seed=rng.randint(1, 1 << 30))
        rc, dst = oracle_resize(LANCZOS, src, dw, dh, deg, 1)
        if rc != 0:
            continue
        rc2, ref = ref_resize(LANCZOS, src, dw, dh, deg, 1)
        assert rc2 == 0 and np.array_equal(dst, ref), (deg, sw, sh, dw, dh)
        defined += 1
    assert defined > 100
    for sw, sh, dw, dh in ((10, 20, 3, 20), (10, 20, 4, 20), (10, 40, 3, 40)):   # the advisor's cases
        src = lcg_image(sh, sw, seed=3)
        rc, dst = oracle_resize(LANCZOS, src, dw, dh, 3)
        assert rc == 0 and np.array_equal(dst, ref_resize(LANCZOS, src, dw, dh, 3)[1])
    # mainBegin > dstW: the reference's first border loop writes past the row -> still rejected; so is the Y axis
    assert oracle_resize(LANCZOS, lcg_image(30, 7), 2, 30, 3)[0] == -2
    assert oracle_resize(LANCZOS, lcg_image(10, 30), 30, 3, 3)[0] == -2
