"""Tensor-path Lanczos kernel (resizeLanczosMmaKernel: both passes as integer mma.sync matrix products, source rows
fed by TMA) forced with IQO_CUDA_PATH_MMA and compared bit-for-bit with the oracle: the BASELINE shapes, odd sizes,
padded pitches, row bands, batches and a random sweep.  Replaces the reference's resizeYmain / resizeYborder /
resizeXmain / resizeXborder loops (src/IQOLanczosResizerImpl_Generic.cpp:464-612)."""
import random

import numpy as np
import pytest

import libiqo_b200 as iqo
from oracle_lib import LANCZOS, fnv1a, lcg_image, oracle_resize

pytestmark = pytest.mark.gpu


def mma_resize(src, dw, dh, deg, px=1, sw=None, dst_stride=None):
    sh, sst = src.shape
    sw = sw or sst
    dst_stride = dst_stride or dw
    dst = np.full((dh, dst_stride), 0xA5, dtype=np.uint8)
    with iqo.LanczosResizer(deg, sw, sh, dw, dh, px) as r:
        r.set_path(iqo.PATH_MMA)
        r.resize(sst, src, dst_stride, dst)
        kernel = r.last_kernel()
    return dst, kernel


CASES = [
    # (degree, pxScale, srcW, srcH, dstW, dstH, srcPad, dstPad)
    (3, 1, 1920, 1080, 960, 540, 0, 0),      # cfg4
    (3, 1, 1920, 1080, 1280, 720, 0, 0),     # cfg1
    (2, 1, 3840, 2160, 1920, 1080, 0, 0),    # cfg3 luma
    (2, 2, 1920, 1080, 960, 540, 0, 0),      # cfg3 chroma
    (4, 1, 4096, 1024, 1500, 375, 0, 0),     # cfg5's ratio
    (4, 1, 2048, 2048, 750, 750, 0, 0),
    (3, 1, 64, 48, 40, 30, 0, 0),
    (3, 1, 40, 30, 64, 48, 0, 0),            # up-sampling
    (3, 1, 320, 180, 640, 360, 0, 0),
    (2, 1, 34, 21, 64, 47, 0, 0),
    (3, 1, 96, 54, 64, 36, 16, 3),           # padded pitches, unaligned destination stride (byte stores)
    (3, 1, 488, 250, 244, 125, 8, 0),
    (3, 1, 1208, 98, 604, 49, 0, 0),
    (5, 1, 900, 700, 330, 211, 4, 1),
    (3, 1, 64, 48, 64, 30, 0, 0),            # X pass-through
    (3, 1, 64, 48, 40, 48, 0, 0),            # Y pass-through
    (1, 1, 50, 40, 24, 20, 14, 0),
    (4, 2, 100, 80, 50, 40, 12, 0),
    (3, 3, 300, 200, 150, 100, 4, 4),
    (3, 1, 18, 400, 9, 200, 14, 7),          # narrower than one tile
]


@pytest.mark.parametrize("case", CASES)
def test_mma_kernel_matches_oracle(case):
    deg, px, sw, sh, dw, dh, spad, dpad = case
    src = lcg_image(sh, sw + spad, seed=31)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    got, kernel = mma_resize(src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert kernel in ("lanczos_mma", "lanczos_mma_dp2a"), kernel
    bad = np.argwhere(got != want)
    assert bad.size == 0, (len(bad), bad[:8].tolist())


def test_mma_extreme_values():
    sw, sh = 480, 272
    yy, xx = np.mgrid[0:sh, 0:sw]
    for name, src in (("white", np.full((sh, sw), 255, np.uint8)), ("black", np.zeros((sh, sw), np.uint8)),
                      ("checker1", (((yy + xx) & 1) * 255).astype(np.uint8)),
                      ("checker2", ((((yy >> 1) + (xx >> 1)) & 1) * 255).astype(np.uint8)),
                      ("vstripes", ((xx & 1) * 255).astype(np.uint8)), ("hstripes", ((yy & 1) * 255).astype(np.uint8))):
        for deg, px, dw, dh in ((3, 1, 240, 136), (2, 1, 240, 136), (3, 1, 320, 181), (4, 1, 176, 100), (3, 1, 960, 544)):
            rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px)
            assert rc == 0
            got, kernel = mma_resize(src, dw, dh, deg, px)
            assert kernel in ("lanczos_mma", "lanczos_mma_dp2a")
            assert np.array_equal(got, want), (name, deg, px, dw, dh)


def test_mma_random_sweep():
    rng = random.Random(4242)
    ran = 0
    for _ in range(150):
        deg = rng.randint(1, 6)
        px = rng.choice([1, 1, 2, 3])
        sw, sh = 2 * rng.randint(4, 300), rng.randint(8, 300)          # even widths (the tensor map views rows as 16-bit pairs)
        dw, dh = rng.randint(4, 500), rng.randint(4, 400)
        if rng.random() < 0.1:
            dw = sw
        if rng.random() < 0.1:
            dh = sh
        spad = rng.choice([0, 16, 32]) + (-sw) % 16                     # staged pitch is 16-byte aligned anyway; vary the host stride
        dpad = rng.randint(0, 5)
        src = lcg_image(sh, sw + spad, seed=rng.randint(1, 1 << 30))
        rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        if rc != 0:
            continue
        got, kernel = mma_resize(src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert np.array_equal(got, want), (deg, px, sw, sh, dw, dh, kernel)
        ran += kernel in ("lanczos_mma", "lanczos_mma_dp2a")
    assert ran >= 80, ran


def test_mma_bands_batches_and_device_pitches():
    torch = pytest.importorskip("torch")
    # row bands with srcRow0 != 0 at cfg5's ratio
    sw, sh, dw, dh = 4096, 1024, 1500, 375
    src = lcg_image(sh, sw, seed=4)
    out = np.zeros((dh, dw), dtype=np.uint8)
    with iqo.LanczosResizer(4, sw, sh, dw, dh) as r:
        r.set_path(iqo.PATH_MMA)
        for y0, n in [(0, 5), (5, 123), (128, 200), (328, 47)]:
            s0, sn = r.band_src_rows(y0, n)
            dsrc = torch.from_numpy(src[s0:s0 + sn].copy()).cuda()
            ddst = torch.full((n, dw + 16), 0xA5, dtype=torch.uint8, device="cuda")
            r.resize_band(y0, n, s0, sn, sw, dsrc, dw + 16, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert r.last_kernel() in ("lanczos_mma", "lanczos_mma_dp2a")
            got = ddst.cpu().numpy()
            assert (got[:, dw:] == 0xA5).all()
            out[y0:y0 + n] = got[:, :dw]
    assert "%016x" % fnv1a(out) == "eb104bb1ff7eb33f"
    # device-resident batch, frame stride with slack, three ratios
    for deg, sw, sh, dw, dh in ((3, 960, 540, 480, 270), (3, 960, 540, 640, 360), (4, 1024, 512, 375, 188)):
        n = 5
        pitch = sw + 32
        host = np.stack([lcg_image(sh, pitch, seed=70 + f) for f in range(n)])
        host[3] = 255
        want = np.stack([oracle_resize(LANCZOS, host[f], dw, dh, deg, sw=sw)[1] for f in range(n)])
        dsrc = torch.from_numpy(host).cuda()
        ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
        with iqo.LanczosResizer(deg, sw, sh, dw, dh) as r:
            r.set_path(iqo.PATH_MMA)
            r.resize_batch(n, pitch, pitch * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert r.last_kernel() in ("lanczos_mma", "lanczos_mma_dp2a")
        assert np.array_equal(ddst.cpu().numpy(), want), (deg, sw, sh, dw, dh)
    # a pitch that is not 16-byte aligned is declined (TMA needs it): another kernel takes the launch
    host = lcg_image(270, 488, seed=3)
    dsrc = torch.from_numpy(host).cuda()
    ddst = torch.zeros((135, 240), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, 480, 270, 240, 135) as r:
        r.set_path(iqo.PATH_MMA)
        r.resize_batch(1, 488, 488 * 270, dsrc, 240, 240 * 135, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert not r.last_kernel().startswith("lanczos_mma")
    assert np.array_equal(ddst.cpu().numpy(), oracle_resize(LANCZOS, host, 240, 135, 3, sw=480)[1])


DP2A_CASES = [
    # (degree, pxScale, srcW, srcH, dstW, dstH, src pad, dst pad): Lanczos3 at 3:2 on X -- tensor-path vertical pass,
    # the 3:2 kernel's compile-time dp2a horizontal pass
    (3, 1, 1920, 1080, 1280, 720, 0, 0),     # cfg1
    (3, 1, 96, 60, 64, 40, 0, 0),            # one narrow strip holding both border sides
    (3, 1, 528, 333, 352, 222, 16, 8),       # partial last strip, padded rows
    (3, 1, 1920, 1080, 1280, 540, 0, 0),     # 3:2 on X, 2:1 on Y
    (3, 1, 960, 540, 640, 333, 0, 0),        # arbitrary Y ratio
    (3, 1, 480, 270, 320, 180, 0, 4),        # destination stride not a multiple of 8: byte stores
    (3, 1, 480, 100, 320, 250, 0, 0),        # up-sampling on Y
    (3, 1, 3840, 64, 2560, 43, 0, 0),        # many strips, few rows
]


@pytest.mark.parametrize("case", DP2A_CASES)
def test_mma_dp2a_kernel(case):
    deg, px, sw, sh, dw, dh, spad, dpad = case
    for seed, fill in ((61, None), (0, 255), (0, 0)):
        src = lcg_image(sh, sw + spad, seed=seed) if fill is None else np.full((sh, sw + spad), fill, np.uint8)
        rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert rc == 0
        got, kernel = mma_resize(src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        assert kernel == "lanczos_mma_dp2a", kernel
        bad = np.argwhere(got != want)
        assert bad.size == 0, (len(bad), bad[:8].tolist())


def test_mma_dp2a_bands_and_batches():
    torch = pytest.importorskip("torch")
    sw, sh, dw, dh = 1920, 1080, 1280, 720
    src = lcg_image(sh, sw, seed=8)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, 3)
    out = np.zeros((dh, dw), dtype=np.uint8)
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        r.set_path(iqo.PATH_MMA)
        for y0, n in [(0, 7), (7, 300), (307, 1), (308, 401), (709, 11)]:
            s0, sn = r.band_src_rows(y0, n)
            dsrc = torch.from_numpy(src[s0:s0 + sn].copy()).cuda()
            ddst = torch.full((n, dw + 16), 0xA5, dtype=torch.uint8, device="cuda")
            r.resize_band(y0, n, s0, sn, sw, dsrc, dw + 16, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert r.last_kernel() == "lanczos_mma_dp2a"
            got = ddst.cpu().numpy()
            assert (got[:, dw:] == 0xA5).all()
            out[y0:y0 + n] = got[:, :dw]
    assert np.array_equal(out, want)
    n = 7
    host = np.stack([lcg_image(540, 960 + 32, seed=90 + f) for f in range(n)])
    host[2] = 255
    wants = np.stack([oracle_resize(LANCZOS, host[f], 640, 360, 3, sw=960)[1] for f in range(n)])
    dsrc = torch.from_numpy(host).cuda()
    ddst = torch.zeros((n, 360, 640), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, 960, 540, 640, 360) as r:
        r.set_path(iqo.PATH_MMA)
        r.resize_batch(n, 992, 992 * 540, dsrc, 640, 640 * 360, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert r.last_kernel() == "lanczos_mma_dp2a"
    assert np.array_equal(ddst.cpu().numpy(), wants)


def test_auto_picks_the_kernel_by_ratio_taps_and_launch_size():
    """AUTO: big batches of 3:2-style reductions with ten or more taps run on the tensor path, Lanczos2 at 3:2,
    up-sampling and small launches stay on the 3:2 streaming / packed kernels (capi.cu, measured cross-overs in
    DESIGN 4.8) -- and whatever runs is bit-exact."""
    torch = pytest.importorskip("torch")
    for deg, sw, sh, dw, dh, n, want_kernel in ((3, 960, 540, 640, 360, 96, "lanczos_mma_dp2a"),   # 3:2 with the 10-tap pattern: dp2a horizontal pass
                                                (4, 960, 540, 640, 360, 96, "lanczos_mma"),
                                                (3, 960, 540, 480, 360, 96, "lanczos_mma"),   # 2:1 on X, 3:2 on Y
                                                (2, 960, 540, 640, 360, 96, "ratio_stream"),
                                                (3, 480, 270, 960, 540, 96, "ratio_stream"),
                                                (3, 960, 540, 640, 360, 2, None)):
        host = np.stack([lcg_image(sh, sw, seed=40 + f) for f in range(min(n, 3))])
        want = np.stack([oracle_resize(LANCZOS, host[f], dw, dh, deg)[1] for f in range(host.shape[0])])
        dsrc = torch.from_numpy(host).cuda()[torch.arange(n) % host.shape[0]].contiguous()
        ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
        with iqo.LanczosResizer(deg, sw, sh, dw, dh) as r:
            r.resize_batch(n, sw, sw * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            kernel = r.last_kernel()
        if want_kernel is not None:
            assert kernel == want_kernel, (deg, sw, sh, dw, dh, n, kernel)
        else:
            assert not kernel.startswith("lanczos_mma"), kernel
        got = ddst.cpu().numpy()
        for f in range(n):
            assert np.array_equal(got[f], want[f % host.shape[0]]), (deg, sw, sh, dw, dh, f, kernel)


# ---------------------------------------------------------------------------------------------
# Area / Linear on the same kernel (unsigned byte planes, 23-bit shift; a weight of 256 is split 255 + 1 over two
# k slots that read the same source row).  Replaces src/IQOAreaResizerImpl_Generic.cpp:303-368 and
# src/IQOLinearResizerImpl_Generic.cpp:290-407 for ratios without a dedicated streaming kernel.
# ---------------------------------------------------------------------------------------------
from oracle_lib import AREA, LINEAR  # noqa: E402

AL_CASES = [
    # (kind, srcW, srcH, dstW, dstH, srcPad, dstPad)
    (AREA, 1920, 1080, 1280, 720, 0, 0),
    (AREA, 1000, 1000, 333, 777, 8, 3),
    (AREA, 64, 48, 40, 30, 0, 0),
    (AREA, 40, 30, 64, 48, 0, 0),            # "up" = nearest-floor (single taps of weight 256 / 32768)
    (AREA, 64, 48, 64, 30, 0, 0),
    (AREA, 64, 48, 40, 48, 0, 0),
    (AREA, 78, 31, 5, 3, 2, 0),
    (AREA, 3840, 2160, 1920, 1080, 0, 0),    # cfg2a on the general kernel
    (LINEAR, 960, 540, 1920, 1080, 0, 0),
    (LINEAR, 640, 480, 1600, 1000, 0, 1),
    (LINEAR, 1280, 720, 3840, 2160, 0, 0),   # cfg2b on the general kernel
    (LINEAR, 40, 30, 100, 75, 0, 0),
    (LINEAR, 32, 18, 32, 54, 0, 0),
    (LINEAR, 32, 18, 96, 18, 0, 0),
    (LINEAR, 2, 2, 5, 6, 14, 0),
    (LINEAR, 200, 150, 500, 420, 8, 4),
]


@pytest.mark.parametrize("case", AL_CASES)
def test_mma_kernel_area_linear(case):
    kind, sw, sh, dw, dh, spad, dpad = case
    src = lcg_image(sh, sw + spad, seed=37)
    rc, want = oracle_resize(kind, src, dw, dh, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    dst = np.full((dh, dw + dpad), 0xA5, dtype=np.uint8)
    with iqo.make_resizer(kind, 0, sw, sh, dw, dh) as r:
        r.set_path(iqo.PATH_MMA)
        r.resize(sw + spad, src, dw + dpad, dst)
        kernel = r.last_kernel()
    assert kernel == ("area_mma" if kind == AREA else "linear_mma"), kernel
    bad = np.argwhere(dst != want)
    assert bad.size == 0, (len(bad), bad[:8].tolist())
    for v in (0, 255):
        flat = np.full((sh, sw), v, np.uint8)
        out = np.zeros((dh, dw), dtype=np.uint8)
        with iqo.make_resizer(kind, 0, sw, sh, dw, dh) as r:
            r.set_path(iqo.PATH_MMA)
            r.resize(sw, flat, dw, out)
        assert (out == v).all()


def test_mma_area_linear_random_sweep():
    rng = random.Random(777)
    ran = 0
    for _ in range(120):
        kind = rng.choice([AREA, LINEAR])
        sw, sh = 2 * rng.randint(2, 200), rng.randint(4, 200)
        if kind == LINEAR:
            dw, dh = rng.randint(sw, 3 * sw), rng.randint(sh, 3 * sh)
        else:
            dw, dh = rng.randint(1, sw + 8), rng.randint(1, sh + 8)
        src = lcg_image(sh, sw + (-sw) % 16, seed=rng.randint(1, 1 << 30))
        dpad = rng.randint(0, 5)
        rc, want = oracle_resize(kind, src, dw, dh, sw=sw, dst_stride=dw + dpad)
        assert rc == 0
        dst = np.full((dh, dw + dpad), 0xA5, dtype=np.uint8)
        with iqo.make_resizer(kind, 0, sw, sh, dw, dh) as r:
            r.set_path(iqo.PATH_MMA)
            r.resize(src.shape[1], src, dw + dpad, dst)
            ran += r.last_kernel().endswith("_mma")
        assert np.array_equal(dst, want), (kind, sw, sh, dw, dh)
    assert ran >= 100, ran
