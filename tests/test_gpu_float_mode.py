"""Optional float ("SIMD-semantics") mode, SURVEY 8f-4 (iqo_cuda_set_arithmetic(IQO_CUDA_ARITH_SIMD_FLOAT)).

Not part of the parity contract: the reference's SIMD implementations compute in float and disagree with its
Generic path.  This mode reproduces their arithmetic -- normalised float tables, FMA accumulation in
tap order, round-to-nearest-even, saturation (src/IQOLanczosResizerImpl_AVX512.cpp:47-60,179-185,385-431,547-590)
-- with correctly masked border denominators.  Stated tolerance against the reference's own SIMD output
(oracle/_ref/libiqo_ref_full.so, CPUID dispatch): max |diff| <= 1 LSB on interior pixels (rows and columns of the
main range, 64 pixels away from the column borders, where the reference's vector loops hand over to its border code)."""
import ctypes as C

import numpy as np
import pytest

import libiqo_b200 as iqo
from oracle_lib import LANCZOS, lcg_image, oracle_resize, ref_full

pytestmark = pytest.mark.gpu


def float_resize(src, dw, dh, deg, px=1):
    sh, sw = src.shape
    dst = np.zeros((dh, dw), dtype=np.uint8)
    with iqo.LanczosResizer(deg, sw, sh, dw, dh, px) as r:
        r.set_arithmetic(iqo.ARITH_SIMD_FLOAT)
        r.resize(sw, src, dw, dst)
        assert r.last_kernel() == "float_simd_semantics"
        fixed = np.zeros((dh, dw), dtype=np.uint8)
        r.set_arithmetic(iqo.ARITH_FIXED)      # back to the contract arithmetic on the same handle
        r.resize(sw, src, dw, fixed)
    return dst, fixed


def simd_reference(src, dw, dh, deg, px=1):
    full = ref_full()
    sh, sw = src.shape
    dst = np.zeros((dh, dw), dtype=np.uint8)
    h = full.iqo_ref_public_new(LANCZOS, deg, sw, sh, dw, dh, px)
    full.iqo_ref_public_resize(h, sw, src.ctypes.data, dw, dst.ctypes.data)
    full.iqo_ref_public_delete(h)
    return dst


def emulate(src, dw, dh, deg, px=1):
    """The mode's definition in numpy (float32 values, products and sums rounded to float32 after every tap)."""
    sh, sw = src.shape
    qx = iqo.plan_query(LANCZOS, deg, sw, sh, dw, dh, px, 0)
    qy = iqo.plan_query(LANCZOS, deg, sw, sh, dw, dh, px, 1)
    # normalised float tables from the integer plan are not available: rebuild them from the Lanczos definition
    import math

    def weight(x):
        ax = abs(x)
        if math.fmod(ax, 1.0) < 1e-5:
            return 1.0 if ax < 1e-5 else 0.0
        if deg <= ax:
            return 0.0
        pi = 3.14159265358979
        u, v = pi * x, pi * (x / deg)
        return (math.sin(u) / u) * (math.sin(v) / v)

    def tables(S, D, N):
        g = math.gcd(S, D)
        rS, rD = S // g, D // g
        out = np.zeros((rD, N), dtype=np.float32)
        for t in range(rD):
            if rS > rD:
                widen = max(1, px // deg)
                sub = ((rD - (t * rS) % rD) * px) % rS
                origin = -deg * widen - 0.5 * px + 0.5 * rD * px / rS + sub / float(rS)
                num, den = rD * px, rS
            else:
                origin = -deg + 1.0 - math.fmod((t * rS) / float(rD), 1.0)
                num, den = rD, rD
            w = np.array([np.float32(weight(origin + (i * num) / float(den))) for i in range(N)], dtype=np.float32)
            s = np.float32(0)
            for v in w:
                s = np.float32(s + v)
            out[t] = w / s
        return out, rD

    def axis_pass(data, S, D, q):   # data: (S, n) float32 -> (D, n)
        N = q["numCoefs"]
        if S == D:
            return data.copy()
        tab, rD = tables(S, D, N)
        res = np.zeros((D, data.shape[1]), dtype=np.float32)
        for d in range(D):
            f = int(q["first"][d])
            acc = np.zeros(data.shape[1], dtype=np.float32)
            den = np.float32(0)
            border = q["row"][d] >= q["numTables"]
            for i in range(N):
                if 0 <= f + i < S:
                    c = tab[d % rD, i]
                    acc = (acc.astype(np.float64) + data[f + i].astype(np.float64) * np.float64(c)).astype(np.float32)
                    den = np.float32(den + c)
            res[d] = acc / den if border else acc
        return res

    work = axis_pass(src.astype(np.float32), sh, dh, qy)          # (dh, sw)
    out = axis_pass(work.T.copy(), sw, dw, qx).T                   # (dh, dw)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8), qx, qy


@pytest.mark.parametrize("case", [(3, 1, 1920, 1080, 960, 540), (3, 1, 1920, 1080, 1280, 720), (2, 1, 1280, 720, 1920, 1080),
                                  (4, 1, 2048, 1024, 750, 375), (2, 2, 960, 540, 480, 270)])
def test_float_mode_matches_the_reference_simd_paths_on_interior_pixels(case):
    if ref_full() is None:
        pytest.skip("oracle/_ref/libiqo_ref_full.so not built")
    deg, px, sw, sh, dw, dh = case
    src = lcg_image(sh, sw, seed=41)
    got, fixed = float_resize(src, dw, dh, deg, px)
    ref = simd_reference(src, dw, dh, deg, px)
    qx = iqo.plan_query(LANCZOS, deg, sw, sh, dw, dh, px, 0)
    qy = iqo.plan_query(LANCZOS, deg, sw, sh, dw, dh, px, 1)
    ys = slice(qy["mainBegin"], qy["mainEnd"])
    xs = slice(qx["mainBegin"] + 64, qx["mainEnd"] - 64)
    a, b = got[ys, xs].astype(int), ref[ys, xs].astype(int)
    assert a.size > 0.5 * dw * dh
    diff = np.abs(a - b)
    assert diff.max() <= 1, "stated tolerance: 1 LSB"
    assert (diff == 0).mean() > 0.99
    # the contract arithmetic on the same handle is untouched by the mode switch
    rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px)
    assert rc == 0 and np.array_equal(fixed, want)
    # ... while the float results differ from it: the 6-bit vertical / 14-bit horizontal tables of Generic and the float
    # tables are different filters (SURVEY 0.3 measured up to 83 levels between the reference's own two paths)
    assert np.abs(got[ys, xs].astype(int) - want[ys, xs].astype(int)).max() <= 16


@pytest.mark.parametrize("case", [(3, 1, 96, 54, 48, 27), (3, 1, 64, 48, 40, 30), (2, 1, 34, 21, 64, 47), (4, 2, 100, 80, 50, 40),
                                  (3, 1, 64, 48, 64, 30), (3, 1, 64, 48, 40, 48)])
def test_float_mode_definition_including_masked_borders(case):
    deg, px, sw, sh, dw, dh = case
    src = lcg_image(sh, sw, seed=43)
    got, _ = float_resize(src, dw, dh, deg, px)
    want, qx, qy = emulate(src, dw, dh, deg, px)
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1, np.argwhere(diff > 1)[:5].tolist()      # float32 FMA vs the float64-assisted emulation
    assert (diff == 0).mean() > 0.98
    for v in (0, 100, 255):                                          # normalised tables, masked denominators: constants stay constant
        flat = np.full((sh, sw), v, np.uint8)
        out, _ = float_resize(flat, dw, dh, deg, px)
        assert (out == v).all()


def test_float_mode_is_lanczos_only():
    with iqo.AreaResizer(64, 48, 32, 24) as r:
        with pytest.raises(iqo.IqoCudaError) as e:
            r.set_arithmetic(iqo.ARITH_SIMD_FLOAT)
        assert e.value.code == -2
