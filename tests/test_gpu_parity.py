"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI
(libiqo_b200 -> libiqo_cuda.so) and is compared bit-for-bit with the oracle / golden fixtures."""
import random

import numpy as np
import pytest

import golden_cases as G
import libiqo_b200 as iqo
from oracle_lib import AREA, LANCZOS, LINEAR, fnv1a, lcg_image, oracle_resize

pytestmark = pytest.mark.gpu

PATHS = [iqo.PATH_AUTO, iqo.PATH_GENERIC]


def gpu_resize(kind, src, dw, dh, deg=0, px=1, sw=None, dst_stride=None, path=iqo.PATH_AUTO):
    sh, sst = src.shape
    sw = sw or sst
    dst_stride = dst_stride or dw
    dst = np.full((dh, dst_stride), 0xA5, dtype=np.uint8)
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(path)
        r.resize(sst, src, dst_stride, dst)
        kernel = r.last_kernel()
    return dst, kernel


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("i", range(len(G.SMALL)), ids=[G.case_id(c) for c in G.SMALL])
def test_golden_small(i, path):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = G.SMALL[i]
    dst, _ = gpu_resize(kind, G.case_src(G.SMALL[i]), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad, path=path)
    assert np.array_equal(dst[:, :dw], G.small_expected(i))
    assert (dst[:, dw:] == 0xA5).all(), "bytes beyond dstW must not be written"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("c", G.LARGE, ids=[G.case_id(c) for c in G.LARGE])
def test_golden_hash(c, path):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed, h = c
    dst, _ = gpu_resize(kind, G.case_src(c), dw, dh, deg, px, sw=sw, dst_stride=dw + dpad, path=path)
    assert "%016x" % fnv1a(dst, dw) == h


@pytest.mark.parametrize("path", PATHS)
def test_random_sweep_against_oracle(path):
    rng = random.Random(99)
    checked = 0
    for _ in range(160):
        kind = rng.choice([LANCZOS, LANCZOS, AREA, LINEAR])
        sw, sh = rng.randint(1, 300), rng.randint(1, 200)
        if kind == LINEAR:
            dw, dh = rng.randint(sw, 3 * sw), rng.randint(sh, 3 * sh)
        elif kind == AREA:
            dw, dh = rng.randint(1, sw + 8), rng.randint(1, sh + 8)
        else:
            dw, dh = rng.randint(1, 400), rng.randint(1, 300)
        if rng.random() < 0.15:
            dw = sw
        if rng.random() < 0.15:
            dh = sh
        deg = rng.randint(1, 9) if kind == LANCZOS else 0
        px = rng.choice([1, 1, 2, 3]) if kind == LANCZOS else 1
        src = lcg_image(sh, sw + rng.randint(0, 5), seed=rng.randint(1, 1 << 30))
        dpad = rng.randint(0, 5)
        rc, want = oracle_resize(kind, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
        if rc != 0:
            with pytest.raises(iqo.IqoCudaError) as e:
                iqo.make_resizer(kind, deg, sw, sh, dw, dh, px)
            assert e.value.code == rc
            continue
        got, _ = gpu_resize(kind, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad, path=path)
        assert np.array_equal(got, want), (kind, deg, px, sw, sh, dw, dh)
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("v", [0, 100, 255])
def test_constant_images(v):
    src = np.full((108, 192), v, dtype=np.uint8)
    for kind, deg, px, dw, dh in ((AREA, 0, 1, 96, 54), (LINEAR, 0, 1, 384, 216), (LANCZOS, 3, 1, 96, 54),
                                  (LANCZOS, 2, 2, 96, 54), (LANCZOS, 3, 1, 128, 72)):
        rc, want = oracle_resize(kind, src, dw, dh, deg, px)
        got, _ = gpu_resize(kind, src, dw, dh, deg, px)
        assert rc == 0 and np.array_equal(got, want)
        if kind != LANCZOS:
            assert (got == v).all()


def test_device_pointers_and_batch():
    torch = pytest.importorskip("torch")
    n, sw, sh, dw, dh = 5, 192, 108, 96, 54
    frames = np.stack([lcg_image(sh, sw, seed=10 + f) for f in range(n)])
    want = np.stack([oracle_resize(LANCZOS, frames[f], dw, dh, 3)[1] for f in range(n)])
    dsrc = torch.from_numpy(frames).cuda()
    ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        # single frame, device pointers, through resize()
        r.resize(sw, dsrc[2], dw, ddst[2])
        assert np.array_equal(ddst[2].cpu().numpy(), want[2])
        ddst.zero_()
        # one launch for the batch on torch's current stream
        r.resize_batch(n, sw, sw * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(ddst.cpu().numpy(), want)
        # host-resident batch through the pipelined path
        out = np.zeros_like(want)
        r.resize_batch_host(n, sw, sw * sh, frames, dw, dw * dh, out)
        assert np.array_equal(out, want)
        # host pointers are rejected by the device-batch entry point
        with pytest.raises(iqo.IqoCudaError):
            r.resize_batch(n, sw, sw * sh, frames, dw, dw * dh, out)


def test_batch_host_irregular_strides_and_many_chunks():
    # frame strides that are not rows*stride, and enough frames for several pipeline chunks
    n, sw, sh, dw, dh = 40, 1920, 1080, 960, 540
    sst, dst_st = sw + 16, dw + 8
    sfs, dfs = sst * sh + 64, dst_st * dh + 32
    src = np.zeros(n * sfs, dtype=np.uint8)
    base = lcg_image(sh, sst, seed=77)
    for f in range(n):
        src[f * sfs: f * sfs + sst * sh] = np.roll(base, f, axis=0).ravel()
    dst = np.full(n * dfs, 0xA5, dtype=np.uint8)
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        r.resize_batch_host(n, sst, sfs, src, dst_st, dfs, dst)
    for f in (0, 1, 17, 39):
        rc, want = oracle_resize(LANCZOS, np.roll(base, f, axis=0), dw, dh, 3, sw=sw, dst_stride=dst_st)
        got = dst[f * dfs: f * dfs + dst_st * dh].reshape(dh, dst_st)
        assert np.array_equal(got, want)
    assert (dst[dst_st * dh: dfs] == 0xA5).all()


@pytest.mark.parametrize("case", [(LANCZOS, 4, 1, 1024, 768, 375, 281), (AREA, 0, 1, 600, 500, 333, 211),
                                  (LINEAR, 0, 1, 200, 150, 500, 420), (LANCZOS, 2, 1, 300, 200, 450, 330)])
def test_row_bands_equal_whole_image(case):
    torch = pytest.importorskip("torch")
    kind, deg, px, sw, sh, dw, dh = case
    src = lcg_image(sh, sw, seed=5)
    rc, want = oracle_resize(kind, src, dw, dh, deg, px)
    assert rc == 0
    out = np.zeros((dh, dw), dtype=np.uint8)
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        bands = [(0, 7), (7, dh // 3), (7 + dh // 3, dh - 7 - dh // 3)]
        for y0, n in bands:
            s0, sn = r.band_src_rows(y0, n)
            dsrc = torch.from_numpy(src[s0:s0 + sn].copy()).cuda()   # band + halo only
            ddst = torch.zeros((n, dw), dtype=torch.uint8, device="cuda")
            r.resize_band(y0, n, s0, sn, sw, dsrc, dw, ddst)
            r.sync()
            out[y0:y0 + n] = ddst.cpu().numpy()
        # a buffer that misses halo rows is refused
        s0, sn = r.band_src_rows(7, dh // 3)
        if sn > 2:
            with pytest.raises(iqo.IqoCudaError):
                r.resize_band(7, dh // 3, s0 + 1, sn - 1, sw, dsrc, dw, ddst)
    assert np.array_equal(out, want)


def test_multi_device_drivers_on_available_devices():
    ndev = iqo.device_count()
    devices = list(range(min(ndev, 2))) * (2 if ndev == 1 else 1)   # 2 shards even on one GPU
    sw, sh, dw, dh = 640, 480, 400, 300
    src = lcg_image(sh, sw, seed=3)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, 3)
    out = np.zeros((dh, dw), dtype=np.uint8)
    iqo.resize_bands_multi(LANCZOS, 3, sw, sh, dw, dh, 1, sw, src, dw, out, devices)
    assert np.array_equal(out, want)
    n = 7
    frames = np.stack([lcg_image(sh, sw, seed=20 + f) for f in range(n)])
    outs = np.zeros((n, dh, dw), dtype=np.uint8)
    iqo.resize_batch_multi(LANCZOS, 3, sw, sh, dw, dh, 1, n, sw, sw * sh, frames, dw, dw * dh, outs, devices)
    for f in range(n):
        assert np.array_equal(outs[f], oracle_resize(LANCZOS, frames[f], dw, dh, 3)[1])


def test_full_size_batch_properties():
    """cfg4 shape at a batch size the oracle cannot check frame by frame: identical frames give
    identical results, and frame 0 hashes to the reference's golden value."""
    torch = pytest.importorskip("torch")
    n, sw, sh, dw, dh = 64, 1920, 1080, 960, 540
    base = torch.from_numpy(lcg_image(sh, sw, seed=1)).cuda()
    dsrc = base.unsqueeze(0).repeat(n, 1, 1).contiguous()
    ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        r.resize_batch(n, sw, sw * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    out = ddst.cpu().numpy()
    assert "%016x" % fnv1a(out[0]) == "bc3ae031361c0774"
    assert (out == out[0]).all()


def test_errors():
    with pytest.raises(iqo.IqoCudaError) as e:
        iqo.LanczosResizer(3, 8, 8, 5, 5)
    assert e.value.code == -2
    with iqo.AreaResizer(64, 48, 32, 24) as r:
        src = np.zeros((48, 64), dtype=np.uint8)
        dst = np.zeros((24, 32), dtype=np.uint8)
        with pytest.raises(iqo.IqoCudaError):
            r.resize(63, src, 32, dst)   # stride < width
        with pytest.raises(iqo.IqoCudaError):
            r.resize(64, 0, 32, dst)     # NULL
        r.resize(64, src, 32, dst)
        assert iqo.launch_count() > 0


HALF_CASES = [
    # (degree, pxScale, srcW, srcH, srcPad, dstPad, expected kernel)
    (3, 1, 1920, 1080, 0, 0, "half_sym"),     # cfg4
    (2, 1, 3840, 2160, 0, 0, "half_sym"),     # cfg3 luma
    (2, 2, 1920, 1080, 0, 0, "half_small"),   # cfg3 chroma (asymmetric 4-tap table): streaming kernel
    (4, 2, 480, 272, 0, 0, "half"),           # asymmetric 8-tap table: tiled kernel, general pair words
    (2, 2, 488, 250, 4, 0, "half"),           # width not a multiple of 16: the streaming kernel declines
    (3, 1, 480, 272, 0, 0, "half_sym"),       # several tiles, last tile row partial
    (3, 1, 488, 250, 4, 0, "half_sym"),       # width not a multiple of the tile, odd dstH, padded src
    (2, 1, 264, 100, 8, 3, "half_sym"),       # unaligned dst stride -> byte stores
    (1, 1, 256, 64, 0, 0, "half_small"),
    (3, 2, 960, 540, 0, 0, "generic"),        # negative border denominator: specialised kernel declines
    (2, 1, 64, 32, 0, 0, "half_sym"),         # image smaller than a tile
    (3, 1, 28, 26, 0, 0, "half_sym"),
    (3, 1, 1208, 98, 0, 0, "half_sym"),       # eleven strips (three CTAs, the last one partly empty), odd dstH
    (2, 1, 248, 1000, 8, 4, "half_sym"),      # tall: several row bands per strip
    (4, 2, 968, 520, 0, 0, "half"),
]


def half_variant(kname, path, sw):
    """Name of the 2:1 Lanczos kernel a single small host image (staged with a 16-byte aligned pitch) runs on."""
    if kname not in ("half", "half_sym"):
        return kname
    if path == iqo.PATH_STREAM and sw % 8 == 0 and sw // 2 >= 32:
        return kname + "_stream"     # a warp per column strip: needs whole 8-column words (AUTO: big launches only)
    if path in (iqo.PATH_AUTO, iqo.PATH_STREAM, iqo.PATH_NO_STREAM):
        return kname + "_tma"        # tiled kernel, source window staged by TMA
    return kname                     # tiled kernel, plain global loads


@pytest.mark.parametrize("path", [iqo.PATH_AUTO, iqo.PATH_STREAM, iqo.PATH_NO_STREAM, iqo.PATH_NO_TMA])
@pytest.mark.parametrize("case", HALF_CASES)
def test_half_kernel(case, path):
    deg, px, sw, sh, spad, dpad, kname = case
    dw, dh = sw // 2, sh // 2
    src = lcg_image(sh, sw + spad, seed=11)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    got, kernel = gpu_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad, path=path)
    assert kernel == half_variant(kname, path, sw)
    assert iqo.plan_kernel(LANCZOS, deg, sw, sh, dw, dh, px)[0] == kname
    bad = np.argwhere(got != want)
    assert bad.size == 0, (len(bad), bad[:8].tolist())


RATIO_CASES = [
    # (degree, pxScale, srcW, srcH, dstW, dstH, srcPad, dstPad, expected kernel)
    (3, 1, 1920, 1080, 1280, 720, 0, 0, "ratio_stream"),   # cfg1
    (3, 1, 96, 60, 64, 40, 0, 0, "ratio_stream"),          # one narrow strip holding both border sides
    (3, 1, 528, 333, 352, 222, 8, 8, "ratio_stream"),      # three strips (the last one partial), padded rows
    (2, 1, 480, 270, 320, 180, 0, 0, "ratio_stream"),      # Lanczos2 at 3:2 (6 taps)
    (3, 1, 320, 180, 640, 360, 0, 0, "ratio_stream"),      # Lanczos3 2x up-sampling
    (3, 1, 384, 216, 512, 288, 0, 0, "ratio_stream"),      # Lanczos3 3:4 up-sampling
    (3, 1, 1920, 1080, 1280, 540, 0, 0, "ratio_stream"),   # 3:2 on X, 2:1 on Y
    (3, 1, 480, 270, 320, 180, 0, 4, "ratio_stream"),      # destination stride not a multiple of 8: byte stores
    (3, 1, 1920, 1080, 960, 720, 0, 0, "ratio_stream"),    # 2:1 on X only (12 taps, first tap on an odd column)
    (2, 1, 960, 540, 480, 333, 0, 0, "ratio_stream"),      # Lanczos2, 2:1 on X, arbitrary Y
    (2, 1, 320, 180, 640, 360, 0, 0, "ratio_stream"),      # Lanczos2 2x up-sampling (4 taps, odd first column)
    (4, 1, 960, 540, 640, 360, 0, 0, "ratio_stream"),      # Lanczos4 at 3:2 (12 taps)
    (3, 2, 960, 540, 640, 360, 0, 0, "ratio_stream"),      # chroma plane of a 1080p -> 720p YUV420 frame (4 taps)
    (3, 1, 480, 270, 320, 180, 4, 0, "ratio_stream"),      # host rows are staged with an aligned pitch
    (3, 1, 492, 270, 328, 180, 0, 0, "lanczos_stream"),    # source width not a multiple of 8: general streaming kernel
]


@pytest.mark.parametrize("case", RATIO_CASES)
def test_ratio_stream_kernel(case):
    deg, px, sw, sh, dw, dh, spad, dpad, kname = case
    src = lcg_image(sh, sw + spad, seed=23)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert rc == 0
    got, kernel = gpu_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad, path=iqo.PATH_STREAM)
    assert kernel == kname
    bad = np.argwhere(got != want)
    assert bad.size == 0, (len(bad), bad[:8].tolist())
    # AUTO gives single small images to the packed kernel (a warp per strip cannot fill the GPU)
    got, kernel = gpu_resize(LANCZOS, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    assert kernel in (kname, "packed")
    assert np.array_equal(got, want)


def test_packed_kernel_wide_source_window():
    """Strong horizontal down-sampling: a 128-pixel tile spans the whole 566-column source row (71 eight-column
    units per row of the vertical pass; found by the randomized fuzz, the row/unit split used a 16-bit reciprocal)."""
    src = lcg_image(434, 582, seed=9)
    rc, want = oracle_resize(LANCZOS, src, 115, 392, 1, 1, sw=566)
    assert rc == 0
    got, kernel = gpu_resize(LANCZOS, src, 115, 392, 1, 1, sw=566, path=iqo.PATH_NO_STREAM)
    assert kernel == "packed"
    assert np.array_equal(got, want)
    got, kernel = gpu_resize(LANCZOS, src, 115, 392, 1, 1, sw=566, path=iqo.PATH_STREAM)
    assert kernel == "lanczos_stream"
    assert np.array_equal(got, want)


def test_ratio_stream_can_be_switched_off():
    src = lcg_image(270, 480, seed=3)
    rc, want = oracle_resize(LANCZOS, src, 320, 180, 3)
    got, kernel = gpu_resize(LANCZOS, src, 320, 180, 3, path=iqo.PATH_NO_STREAM)
    assert kernel == "packed"
    assert np.array_equal(got, want)


def test_ratio_stream_sweep():
    """X ratios the kernel is built for, arbitrary Y ratios (the vertical pass is record driven) and heights
    that end bands and 8-row turns at odd places."""
    rng = np.random.RandomState(5)
    cases = 0
    for (rs, rd, deg) in ((3, 2, 3), (3, 2, 2), (1, 2, 3), (3, 4, 3), (2, 1, 3), (2, 1, 2), (1, 2, 2), (3, 2, 4)):
        for _ in range(5):
            k = int(rng.randint(1, 12)) * 8
            sw, dw = rs * k, rd * k
            if dw < 8 or sw < 16:
                continue
            sh = int(rng.randint(24, 400))
            dh = int(rng.randint(max(8, sh // 3), min(2 * sh, 600)))
            src = lcg_image(sh, sw, seed=int(rng.randint(1, 1000)))
            rc, want = oracle_resize(LANCZOS, src, dw, dh, deg)
            if rc != 0:
                continue
            got, kernel = gpu_resize(LANCZOS, src, dw, dh, deg, path=iqo.PATH_STREAM)
            if iqo.plan_kernel(LANCZOS, deg, sw, sh, dw, dh, 1)[0] == "ratio_stream":
                assert kernel == "ratio_stream"
                cases += 1
            assert np.array_equal(got, want), (rs, rd, deg, sw, sh, dw, dh, kernel)
    assert cases >= 16


def test_ratio_stream_batch_and_extremes():
    torch = pytest.importorskip("torch")
    sw, sh, dw, dh, n = 960, 540, 640, 360, 5
    host = np.stack([lcg_image(sh, sw, seed=70 + f) for f in range(n)])
    host[3] = 255
    host[4] = ((np.indices((sh, sw)).sum(0) & 1) * 255).astype(np.uint8)
    want = np.stack([oracle_resize(LANCZOS, host[f], dw, dh, 3)[1] for f in range(n)])
    dsrc = torch.from_numpy(host).cuda()
    ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        r.set_path(iqo.PATH_STREAM)
        r.resize_batch(n, sw, sw * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert r.last_kernel() == "ratio_stream"
    assert np.array_equal(ddst.cpu().numpy(), want)


def test_half_kernel_extreme_values():
    # all-255 / all-0 / checkerboards maximise the intermediate range (bias and pair-sum headroom)
    sw, sh = 480, 272
    yy, xx = np.mgrid[0:sh, 0:sw]
    for name, src in (("white", np.full((sh, sw), 255, np.uint8)), ("black", np.zeros((sh, sw), np.uint8)),
                      ("checker1", (((yy + xx) & 1) * 255).astype(np.uint8)),
                      ("checker2", ((((yy >> 1) + (xx >> 1)) & 1) * 255).astype(np.uint8)),
                      ("vstripes", ((xx & 1) * 255).astype(np.uint8)), ("hstripes", ((yy & 1) * 255).astype(np.uint8)),
                      ("vstripes2", (((xx >> 1) & 1) * 255).astype(np.uint8)),
                      ("hstripes2", (((yy >> 1) & 1) * 255).astype(np.uint8))):
        for deg, px in ((3, 1), (2, 1), (2, 2), (1, 1)):
            rc, want = oracle_resize(LANCZOS, src, sw // 2, sh // 2, deg, px)
            got, kernel = gpu_resize(LANCZOS, src, sw // 2, sh // 2, deg, px)
            assert kernel.startswith("half")
            assert np.array_equal(got, want), (name, deg, px)
    # the tiled kernel on the same small-tap shapes (the streaming kernel needs 16-column multiples)
    sw = 488
    yy, xx = np.mgrid[0:sh, 0:sw]
    for src in (np.full((sh, sw), 255, np.uint8), (((yy + xx) & 1) * 255).astype(np.uint8)):
        for deg, px in ((2, 2), (1, 1)):
            rc, want = oracle_resize(LANCZOS, src, sw // 2, sh // 2, deg, px)
            got, kernel = gpu_resize(LANCZOS, src, sw // 2, sh // 2, deg, px)
            assert kernel.startswith("half") and kernel != "half_small"
            assert np.array_equal(got, want), (deg, px)


def test_half_kernel_device_pitches():
    """Device-resident frames: a 16-byte aligned pitch can take the streaming variant (AUTO: big launches only)
    or the TMA variant, one that is only 8- or 4-byte aligned the tiled global-load variant, an odd pitch
    falls back to the generic kernel."""
    torch = pytest.importorskip("torch")
    sw, sh, dw, dh, n = 488, 250, 244, 125, 3
    for pitch, path, expect in ((496, iqo.PATH_STREAM, "half_sym_stream"), (496, iqo.PATH_AUTO, "half_sym_tma"),
                                (496, iqo.PATH_NO_STREAM, "half_sym_tma"),
                                (504, iqo.PATH_NO_STREAM, "half_sym"), (504, iqo.PATH_STREAM, "half_sym"),
                                (492, iqo.PATH_AUTO, "half_sym"),
                                (489, iqo.PATH_AUTO, "generic")):
        host = np.stack([lcg_image(sh, pitch, seed=40 + f) for f in range(n)])
        want = np.stack([oracle_resize(LANCZOS, host[f], dw, dh, 3, sw=sw)[1] for f in range(n)])
        dsrc = torch.from_numpy(host).cuda()
        ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
        with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
            r.set_path(path)
            r.resize_batch(n, pitch, pitch * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert r.last_kernel() == expect, (pitch, path)
        assert np.array_equal(ddst.cpu().numpy(), want), (pitch, path)


# ---------------------------------------------------------------------------------------------
# YUV420 frames, command line tools, plan cache
# ---------------------------------------------------------------------------------------------
import os
import subprocess

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "bin")


def yuv_layout(w, h):
    stx, sty = w + w % 2, h + h % 2
    return stx, sty, stx * sty, stx * sty // 4


def oracle_yuv420(kind, deg, frame, sw, sh, dw, dh):
    """Per-plane oracle of one planar frame, exactly as sample/resize_yuv420p.cpp:125-163 calls the classes."""
    sx, sy, ssy, ssu = yuv_layout(sw, sh)
    dx, dy, dsy, dsu = yuv_layout(dw, dh)
    out = np.zeros(dsy + 2 * dsu, dtype=np.uint8)
    rc, y = oracle_resize(kind, frame[:ssy].reshape(sy, sx)[:sh], dw, dh, deg, 1, sw=sw, dst_stride=dx)
    assert rc == 0
    out[:dh * dx] = y.ravel()
    for p in range(2):
        plane = frame[ssy + p * ssu: ssy + (p + 1) * ssu].reshape(sy // 2, sx // 2)
        rc, c = oracle_resize(kind, plane, dx // 2, dy // 2, deg, 2, dst_stride=dx // 2)
        assert rc == 0
        out[dsy + p * dsu: dsy + (p + 1) * dsu] = c.ravel()
    return out


@pytest.mark.parametrize("case", [(LANCZOS, 2, 384, 216, 192, 108), (LANCZOS, 3, 386, 218, 258, 146),
                                  (AREA, 0, 384, 216, 192, 108), (LINEAR, 0, 128, 72, 384, 216)])
def test_yuv420_frames(case):
    torch = pytest.importorskip("torch")
    kind, deg, sw, sh, dw, dh = case
    n = 3
    with iqo.Yuv420Resizer(kind, deg, sw, sh, dw, dh) as r:
        src = lcg_image(n, r.src_frame_bytes, seed=9)
        want = np.stack([oracle_yuv420(kind, deg, src[f], sw, sh, dw, dh) for f in range(n)])
        dst = np.zeros((n, r.dst_frame_bytes), dtype=np.uint8)
        r.resize(n, src, dst)                       # host frames
        assert np.array_equal(dst, want)
        dsrc = torch.from_numpy(src).cuda()         # device frames, three launches on torch's stream
        ddst = torch.zeros((n, r.dst_frame_bytes), dtype=torch.uint8, device="cuda")
        r.resize(n, dsrc, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(ddst.cpu().numpy(), want)


@pytest.mark.parametrize("method,kind,deg", [("lanczos2", LANCZOS, 2), ("area", AREA, 0)])
def test_cli_resize_yuv420p(tmp_path, method, kind, deg):
    sw, sh, dw, dh = 384, 216, 192, 108
    sx, sy, ssy, ssu = yuv_layout(sw, sh)
    frame = lcg_image(1, ssy + 2 * ssu, seed=21)[0]
    want = oracle_yuv420(kind, deg, frame, sw, sh, dw, dh)
    fin = tmp_path / "in.yuv"
    frame.tofile(fin)
    tools = ["resize_yuv420p"]
    if os.path.exists(os.path.join(BIN, "ref_resize_yuv420p")):
        tools.append("ref_resize_yuv420p")   # the reference's own sample source, linked to this backend
    for tool in tools:
        fout = tmp_path / (tool + ".yuv")
        out = subprocess.run([os.path.join(BIN, tool), "-m", method, "-i", str(fin), "-iw", str(sw), "-ih", str(sh),
                              "-o", str(fout), "-ow", str(dw), "-oh", str(dh)], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "size: %dx%d" % (sw, sh) in out.stdout
        got = np.fromfile(fout, dtype=np.uint8)
        assert np.array_equal(got, want), tool


def test_cli_benchmark_runs():
    for tool in ("benchmark", "ref_benchmark"):
        exe = os.path.join(BIN, tool)
        if not os.path.exists(exe):
            continue
        out = subprocess.run([exe, "-m", "lanczos3", "-iw", "640", "-ih", "360", "-ow", "320", "-oh", "180"],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "elapsed time:" in out.stdout and "ms/cycle" in out.stdout


def test_plan_cache_and_workspace_pool():
    import time
    src = lcg_image(108, 192, seed=5)
    rc, want = oracle_resize(LANCZOS, src, 96, 54, 3)
    dst = np.zeros((54, 96), dtype=np.uint8)
    with iqo.LanczosResizer(3, 192, 108, 96, 54) as r:   # builds and caches the plan
        r.resize(192, src, 96, dst)
    t0 = time.perf_counter()
    for _ in range(200):                                  # the reference benchmark's pattern
        dst[:] = 0
        with iqo.LanczosResizer(3, 192, 108, 96, 54) as r:
            r.resize(192, src, 96, dst)
        assert np.array_equal(dst, want)
    per_cycle = (time.perf_counter() - t0) / 200
    assert per_cycle < 5e-3, per_cycle
    iqo.lib().iqo_cuda_clear_cache()
    with iqo.LanczosResizer(3, 192, 108, 96, 54) as r:
        dst[:] = 0
        r.resize(192, src, 96, dst)
    assert np.array_equal(dst, want)


@pytest.mark.parametrize("case", [(3840, 2160, 0, "area2"), (96, 54, 0, "area2"), (208, 50, 0, "area2"),
                                  (100, 54, 0, "packed")])   # width not a multiple of 16: general kernel
def test_area_2to1_streaming_kernel(case):
    sw, sh, spad, kname = case
    src = lcg_image(sh, sw + spad, seed=31)
    rc, want = oracle_resize(AREA, src, sw // 2, sh // 2, sw=sw)
    got, kernel = gpu_resize(AREA, src, sw // 2, sh // 2, sw=sw)
    assert kernel == kname
    assert np.array_equal(got, want)
    for v in (0, 255):
        flat = np.full((sh, sw), v, np.uint8)
        got, _ = gpu_resize(AREA, flat, sw // 2, sh // 2)
        assert (got == v).all()


@pytest.mark.parametrize("kind,deg,sw,sh,dw,dh", [(LANCZOS, 1, 32, 16, 16, 8), (AREA, 0, 32, 16, 16, 8),
                                                  (LINEAR, 0, 8, 4, 16, 8), (LANCZOS, 2, 24, 12, 16, 8)])
def test_more_than_65535_frames_in_one_call(kind, deg, sw, sh, dw, dh):
    """gridDim.z is limited to 65535: the batch entry point must chunk transparently."""
    torch = pytest.importorskip("torch")
    n = 70000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    dsrc = torch.randint(0, 256, (n, sh, sw), dtype=torch.uint8, device="cuda", generator=gen)
    ddst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh) as r:
        r.resize_batch(n, sw, sw * sh, dsrc, dw, dw * dh, ddst, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    for f in (0, 1, 65534, 65535, 65536, n - 1):
        rc, want = oracle_resize(kind, dsrc[f].cpu().numpy(), dw, dh, deg)
        assert rc == 0 and np.array_equal(ddst[f].cpu().numpy(), want), f


@pytest.mark.parametrize("case", [(1280, 720, 3840, 2160, "linear_up3"), (64, 36, 192, 108, "linear_up3"),
                                  (64, 36, 128, 72, "linear_up2"), (64, 36, 128, 100, "linear_up2"),
                                  (64, 36, 192, 36, "packed"),      # Y pass-through: general kernel
                                  (68, 10, 204, 23, "linear_up3"), (4, 4, 12, 12, "linear_up3"), (8, 3, 16, 7, "linear_up2"),
                                  # 2:3 and 3:4 on X (items of 8 / 12 source columns), any ratio on Y
                                  (1280, 720, 1920, 1080, "linear_up_2_3"), (64, 36, 96, 54, "linear_up_2_3"),
                                  (8, 3, 12, 7, "linear_up_2_3"), (72, 50, 108, 31, "linear_up_2_3"),
                                  (1440, 810, 1920, 1080, "linear_up_3_4"), (12, 5, 16, 9, "linear_up_3_4"),
                                  (84, 33, 112, 100, "linear_up_3_4"),
                                  (64, 36, 256, 144, "linear_up4"), (4, 4, 16, 9, "linear_up4"),
                                  (768, 432, 1920, 1080, "linear_up_2_5"), (8, 6, 20, 15, "linear_up_2_5"),
                                  (1536, 864, 1920, 1080, "linear_up_4_5"), (16, 9, 20, 11, "linear_up_4_5"),
                                  # mild reductions on X with the same item geometry
                                  (1920, 1080, 1280, 720, "linear_3_2"), (12, 9, 8, 6, "linear_3_2"), (96, 50, 64, 77, "linear_3_2"),
                                  (1920, 1080, 1440, 810, "linear_4_3"), (16, 12, 12, 9, "linear_4_3"),
                                  (1920, 1080, 960, 540, "linear_mma|packed"),   # from 2:1 on: other kernels
                                  (60, 36, 90, 54, "linear_mma|packed")])     # 60 is not a multiple of 8: other kernels
def test_linear_integer_upsampling_kernel(case):
    sw, sh, dw, dh, kname = case
    for seed, fill in ((41, None), (0, 255), (0, 0)):
        src = lcg_image(sh, sw, seed=seed) if fill is None else np.full((sh, sw), fill, np.uint8)
        rc, want = oracle_resize(LINEAR, src, dw, dh)
        got, kernel = gpu_resize(LINEAR, src, dw, dh)
        assert kernel in kname.split("|"), kernel
        bad = np.argwhere(got != want)
        assert bad.size == 0, (len(bad), bad[:6].tolist())


@pytest.mark.parametrize("case", [(1920, 1080, 1280, 720, "area_down"),      # 3:2 both axes
                                  (96, 60, 64, 40, "area_down"), (12, 9, 8, 6, "area_down"),
                                  (1920, 1080, 1440, 810, "area_down"),      # 4:3
                                  (64, 48, 48, 17, "area_down"),             # 4:3 on X, arbitrary Y
                                  (1920, 1080, 960, 720, "area_down"),       # 2:1 on X, 3:2 on Y
                                  (3840, 2160, 1536, 864, "area_down"),      # 5:2
                                  (40, 30, 16, 12, "area_down"),
                                  (3840, 2160, 1280, 720, "area_down"),      # 3:1
                                  (36, 21, 12, 50, "area_down"),             # 3:1 on X, up-sampling on Y
                                  (3840, 2160, 960, 540, "area_down"),       # 4:1
                                  (32, 20, 8, 5, "area_down"),
                                  (90, 60, 60, 40, "area_mma|packed"),       # 90 is not a multiple of 12: other kernels
                                  (1000, 700, 700, 400, "area_mma|packed")]) # 10:7
def test_area_reduction_kernel(case):
    sw, sh, dw, dh, kname = case
    for seed, fill in ((43, None), (0, 255), (0, 0)):
        src = lcg_image(sh, sw, seed=seed) if fill is None else np.full((sh, sw), fill, np.uint8)
        rc, want = oracle_resize(AREA, src, dw, dh)
        assert rc == 0
        got, kernel = gpu_resize(AREA, src, dw, dh)
        assert kernel in kname.split("|"), kernel
        bad = np.argwhere(got != want)
        assert bad.size == 0, (len(bad), bad[:6].tolist())
    # padded rows on both sides
    src = lcg_image(sh, sw + 8, seed=44)
    rc, want = oracle_resize(AREA, src, dw, dh, sw=sw, dst_stride=dw + 4)
    got, kernel = gpu_resize(AREA, src, dw, dh, sw=sw, dst_stride=dw + 4)
    assert kernel in kname.split("|"), kernel
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", [(5, 614, 411, 401, 342, 7), (5, 332, 289, 636, 285, 1)])
def test_border_row_that_wraps_int16_after_its_division(case):
    """Lanczos5 tables whose first / last rows keep few taps: `int16(int16(nume) * 64 / deno)` (resizeYborder) can exceed
    int16 for some content and wraps in the reference.  Found by the long fuzz runs (tools/dev/fuzz_paths.sh): the
    biased 16-bit window of the fast kernels cannot hold the wrapped value, so the planner hands such tables to the
    generic kernel whatever path is asked for."""
    deg, sw, sh, dw, dh, seed = case
    src = lcg_image(sh, sw, seed=seed)
    rc, want = oracle_resize(LANCZOS, src, dw, dh, deg)
    assert rc == 0
    assert iqo.plan_kernel(LANCZOS, deg, sw, sh, dw, dh, 1)[0] == "generic"
    for path in (iqo.PATH_AUTO, iqo.PATH_STREAM, iqo.PATH_MMA, iqo.PATH_NO_STREAM, iqo.PATH_GENERIC):
        got, kernel = gpu_resize(LANCZOS, src, dw, dh, deg, path=path)
        assert kernel == "generic", (path, kernel)
        assert np.array_equal(got, want), path


def test_resize_is_ordered_after_the_producer_of_a_device_source():
    """iqo_cuda_resize with device pointers runs on the legacy default stream: a source that torch is still
    producing on its (default) current stream must be complete before the kernel reads it (the call
    used to launch on a private non-blocking stream)."""
    torch = pytest.importorskip("torch")
    sw, sh, dw, dh = 1920, 1080, 960, 540
    base = torch.from_numpy(lcg_image(sh, sw, seed=1)).cuda()
    want = "bc3ae031361c0774"
    big = torch.zeros((64, sh, sw), dtype=torch.uint8, device="cuda")
    dst = torch.zeros((dh, dw), dtype=torch.uint8, device="cuda")
    with iqo.LanczosResizer(3, sw, sh, dw, dh) as r:
        for rep in range(5):
            src = torch.zeros((sh, sw), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            # a few milliseconds of queued work, then the producer of `src`, then resize() without any sync
            for _ in range(20):
                big.add_(1)
            src.copy_(base)
            r.resize(sw, src, dw, dst)
            assert "%016x" % fnv1a(dst.cpu().numpy()) == want, rep


def test_narrow_sources_all_border_columns():
    """Source narrower than the horizontal kernel (mainBegin > mainEnd on X, every column a border column):
    defined by the reference, used to be rejected with -2."""
    rng = random.Random(5)
    done = 0
    for _ in range(60):
        deg = rng.randint(1, 6)
        sw, sh = rng.randint(2, 14), rng.randint(30, 60)
        dw, dh = rng.randint(1, sw), rng.choice([sh, rng.randint(20, 80)])
        src = lcg_image(sh, sw, seed=rng.randint(1, 1 << 30))
        rc, want = oracle_resize(LANCZOS, src, dw, dh, deg, 1)
        if rc != 0:
            with pytest.raises(iqo.IqoCudaError) as e:
                iqo.LanczosResizer(deg, sw, sh, dw, dh)
            assert e.value.code == rc
            continue
        for path in PATHS:
            got, _ = gpu_resize(LANCZOS, src, dw, dh, deg, 1, path=path)
            assert np.array_equal(got, want), (deg, sw, sh, dw, dh, path)
        done += 1
    assert done > 15
    got, _ = gpu_resize(LANCZOS, lcg_image(20, 10, seed=3), 3, 20, 3)
    assert np.array_equal(got, oracle_resize(LANCZOS, lcg_image(20, 10, seed=3), 3, 20, 3)[1])


def test_extreme_downsampling_uses_a_smaller_generic_tile():
    """Area 16000 -> 8 wide: the source window of an 8-pixel tile is 16000 columns; the generic kernel shrinks its
    tile height instead of failing at launch."""
    src = lcg_image(40, 16000, seed=2)
    rc, want = oracle_resize(AREA, src, 8, 20)
    assert rc == 0
    got, kernel = gpu_resize(AREA, src, 8, 20)
    assert kernel == "generic"
    assert np.array_equal(got, want)
    src = lcg_image(64, 9000, seed=3)
    rc, want = oracle_resize(LANCZOS, src, 5, 32, 2)
    assert rc == 0
    got, kernel = gpu_resize(LANCZOS, src, 5, 32, 2)
    assert np.array_equal(got, want)
