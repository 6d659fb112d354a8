"""Randomized parity cases shared by the `-m gpu` fuzz tests (bounded case counts, fixed seeds) and the
open-ended tools under tools/dev/ (time budgets).  Every case goes through the C ABI and is compared
bit-for-bit with the oracle; a case returns (kernel name, ok) or None when the reference leaves the
shape undefined (the oracle reports an error code)."""
import numpy as np

import libiqo_b200 as iqo
from oracle_lib import AREA, LANCZOS, LINEAR, lcg_image, oracle_resize


def _pad(rng):
    return int(rng.choice([0, 0, 4, 8, 16, 3]))


def single_case(rng):
    """(kind, degree, pxScale, sw, sh, dw, dh, srcPad, dstPad) drawn from the kernel families."""
    fam = rng.randint(0, 10)
    if fam == 7:      # Area reductions with their own streaming kernel (items of 4 RS source columns), any Y ratio
        rs, rd = [(3, 2), (4, 3), (2, 1), (5, 2), (3, 1), (4, 1)][rng.randint(0, 6)]
        m = int(rng.randint(1, 60))
        sh = int(rng.randint(4, 400))
        dh = int(rng.randint(max(2, sh // 5), sh + 1)) if rng.rand() < 0.8 else int(rng.randint(sh, 2 * sh))
        return (AREA, 0, 1, 4 * rs * m, sh, 4 * rd * m, dh, _pad(rng), int(rng.choice([0, 0, 4, 8])))
    if fam == 8:      # Linear at the rational ratios of the streaming Linear kernel, any Y ratio
        rs, rd = [(1, 2), (1, 3), (1, 4), (2, 3), (2, 5), (3, 4), (4, 5), (3, 2), (4, 3)][rng.randint(0, 9)]
        m = int(rng.randint(1, 50))
        sh = int(rng.randint(4, 300))
        dh = int(rng.randint(max(2, sh // 2), 3 * sh + 1))
        return (LINEAR, 0, 1, 4 * rs * m, sh, 4 * rd * m, dh, int(rng.choice([0, 0, 4, 8])), int(rng.choice([0, 0, 4, 8])))
    if fam == 9:      # Lanczos3 at 3:2 on X (tensor-path vertical + dp2a horizontal when the tensor path is forced)
        k = int(rng.randint(1, 80)) * 8
        sh = int(rng.randint(16, 500))
        dh = int(rng.randint(max(8, sh // 3), 2 * sh))
        return (LANCZOS, 3, 1, 3 * k, sh, 2 * k, dh, int(rng.choice([0, 0, 16, 32])), _pad(rng))
    if fam == 0:      # 2:1 Lanczos
        dw, dh = int(rng.randint(2, 400)) * 2, int(rng.randint(8, 300))
        return (LANCZOS, int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 1, 2, 3])), 2 * dw, 2 * dh, dw, dh, _pad(rng), _pad(rng))
    if fam == 1:      # ratio kernel families
        rs, rd = [(3, 2), (1, 2), (3, 4), (2, 1)][rng.randint(0, 4)]
        k = int(rng.randint(1, 60)) * 8
        sh = int(rng.randint(16, 500))
        dh = int(rng.randint(max(8, sh // 3), 2 * sh))
        return (LANCZOS, int(rng.choice([1, 2, 3, 4])), 1, rs * k, sh, rd * k, dh, _pad(rng), _pad(rng))
    if fam == 2:      # linear integer up-sampling on X
        kx = int(rng.choice([2, 3]))
        sw, sh = int(rng.randint(2, 200)) * 4, int(rng.randint(4, 200))
        dh = int(rng.randint(sh, 3 * sh + 1))
        return (LINEAR, 0, 1, sw, sh, kx * sw, dh, _pad(rng), _pad(rng))
    if fam == 3:      # area 2:1
        dw, dh = int(rng.randint(1, 200)) * 8, int(rng.randint(4, 300))
        return (AREA, 0, 1, 2 * dw, 2 * dh, dw, dh, _pad(rng), _pad(rng))
    if fam == 4:      # arbitrary Lanczos
        sw, sh = int(rng.randint(24, 700)), int(rng.randint(24, 500))
        dw, dh = int(rng.randint(12, 700)), int(rng.randint(12, 500))
        return (LANCZOS, int(rng.choice([1, 2, 3, 4, 5])), int(rng.choice([1, 1, 2, 3])), sw, sh, dw, dh, _pad(rng), _pad(rng))
    if fam == 5:      # arbitrary area
        sw, sh = int(rng.randint(8, 500)), int(rng.randint(8, 400))
        return (AREA, 0, 1, sw, sh, int(rng.randint(4, sw + 1)), int(rng.randint(4, sh + 1)), _pad(rng), _pad(rng))
    sw, sh = int(rng.randint(8, 500)), int(rng.randint(8, 400))  # arbitrary linear (<= 3x: the reference's defined range)
    return (LINEAR, 0, 1, sw, sh, int(rng.randint(sw, 3 * sw + 1)), int(rng.randint(sh, 3 * sh + 1)), _pad(rng), _pad(rng))


def run_single(rng, case, path):
    kind, deg, px, sw, sh, dw, dh, spad, dpad = case
    src = lcg_image(sh, sw + spad, seed=int(rng.randint(1, 1 << 20)))
    if rng.rand() < 0.15:
        src[:] = rng.choice([0, 255])
    rc, want = oracle_resize(kind, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
    if rc != 0:
        return None
    dst = np.full((dh, dw + dpad), 0xA5, dtype=np.uint8)
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(path)
        r.resize(sw + spad, src, dw + dpad, dst)
        k = r.last_kernel()
    return k, bool(np.array_equal(dst, want))


def batch_case(rng):
    fam = rng.randint(0, 8)
    if fam == 5:      # Area reductions with their own streaming kernel
        rs, rd = [(3, 2), (4, 3), (2, 1), (5, 2), (3, 1), (4, 1)][rng.randint(0, 6)]
        m = int(rng.randint(1, 30))
        kind, deg, px, sw, sh, dw = AREA, 0, 1, 4 * rs * m, int(rng.randint(4, 150)), 4 * rd * m
        dh = int(rng.randint(max(2, sh // 4), sh + 1))
    elif fam == 6:    # Linear at the rational ratios of the streaming Linear kernel
        rs, rd = [(1, 2), (1, 4), (2, 3), (2, 5), (3, 4), (4, 5), (3, 2), (4, 3)][rng.randint(0, 8)]
        m = int(rng.randint(1, 30))
        kind, deg, px, sw, sh, dw = LINEAR, 0, 1, 4 * rs * m, int(rng.randint(4, 120)), 4 * rd * m
        dh = int(rng.randint(max(2, sh // 2), 3 * sh + 1))
    elif fam == 7:    # Lanczos3 / 4 at 3:2 and 2:1-on-X (tensor-path variants when that path is forced)
        k = int(rng.randint(1, 40)) * 8
        rs, rd = [(3, 2), (3, 2), (2, 1)][rng.randint(0, 3)]
        kind, deg, px, sw, sh, dw, dh = LANCZOS, int(rng.choice([3, 3, 4])), 1, rs * k, int(rng.randint(16, 200)), rd * k, int(rng.randint(16, 200))
    elif fam == 0:
        dw, dh = int(rng.randint(4, 200)) * 2, int(rng.randint(8, 120))
        kind, deg, px, sw, sh = LANCZOS, int(rng.choice([2, 3])), 1, 2 * dw, 2 * dh
    elif fam == 1:
        k = int(rng.randint(1, 30)) * 8
        kind, deg, px, sw, sh, dw, dh = LANCZOS, 3, 1, 3 * k, int(rng.randint(16, 200)), 2 * k, int(rng.randint(16, 200))
    elif fam == 2:
        sw, sh = int(rng.randint(2, 100)) * 4, int(rng.randint(4, 100))
        kind, deg, px, dw, dh = LINEAR, 0, 1, 3 * sw, int(rng.randint(sh, 3 * sh))
    elif fam == 3:
        dw, dh = int(rng.randint(1, 100)) * 8, int(rng.randint(4, 100))
        kind, deg, px, sw, sh = AREA, 0, 1, 2 * dw, 2 * dh
    else:             # arbitrary Lanczos ratio (general streaming / packed kernels)
        sw, sh = int(rng.randint(24, 400)), int(rng.randint(24, 200))
        dw, dh = int(rng.randint(12, 400)), int(rng.randint(12, 200))
        kind, deg, px = LANCZOS, int(rng.choice([2, 3, 4])), 1
    n = int(rng.randint(1, 4))
    spitch = sw + int(rng.choice([0, 4, 8, 16, 12, 1]))
    dpitch = dw + int(rng.choice([0, 4, 8, 3]))
    sfs = spitch * sh + int(rng.choice([0, 8, 16, 4]))
    dfs = dpitch * dh + int(rng.choice([0, 8, 5]))
    soff, doff = int(rng.choice([0, 16, 8, 4])), int(rng.choice([0, 8, 4, 1]))
    return (kind, deg, px, sw, sh, dw, dh, n, spitch, dpitch, sfs, dfs, soff, doff)


def run_batch(rng, case, path):
    """Device-resident frames with random pitches / frame strides / base offsets through iqo_cuda_resize_batch."""
    import torch
    kind, deg, px, sw, sh, dw, dh, n, spitch, dpitch, sfs, dfs, soff, doff = case
    host = np.zeros(soff + n * sfs + 64, dtype=np.uint8)
    want = []
    for f in range(n):
        img = lcg_image(sh, spitch, seed=int(rng.randint(1, 1 << 20)))
        host[soff + f * sfs: soff + f * sfs + sh * spitch] = img.reshape(-1)
        rc, w = oracle_resize(kind, img, dw, dh, deg, px, sw=sw)
        if rc != 0:
            return None
        want.append(w)
    dsrc = torch.from_numpy(host).cuda()
    ddst = torch.full((doff + n * dfs + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(path)
        r.resize_batch(n, spitch, sfs, dsrc[soff:], dpitch, dfs, ddst[doff:], torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = r.last_kernel()
    out = ddst.cpu().numpy()
    ok = True
    for f in range(n):
        got = out[doff + f * dfs: doff + f * dfs + dh * dpitch].reshape(dh, dpitch)
        ok = ok and np.array_equal(got[:, :dw], want[f]) and bool((got[:, dw:] == 0xA5).all())
    return k, ok


def run_band_case(rng, path):
    """A random shape of any kind in a random ragged partition of destination row bands, each band resized with
    iqo_cuda_resize_band from a device buffer that holds only the band + halo rows (srcRow0 != 0); the assembled image
    must equal the oracle's whole-image result.  Returns (kernel of the last band, ok) or None."""
    import torch
    kind, deg, px, sw, sh, dw, dh, _, _ = single_case(rng)
    spad = int(rng.choice([0, 0, 16, 32]))
    src = lcg_image(sh, sw, seed=int(rng.randint(1, 1 << 20)))
    rc, want = oracle_resize(kind, src, dw, dh, deg, px)
    if rc != 0:
        return None
    cuts = sorted(set([0, dh] + [int(v) for v in rng.randint(1, max(2, dh), size=int(rng.randint(0, 4)))]))
    out = np.full((dh, dw), 0xEE, dtype=np.uint8)
    k = "none"
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(path)
        for y0, y1 in zip(cuts[:-1], cuts[1:]):
            n = y1 - y0
            s0, sn = r.band_src_rows(y0, n)
            host = np.zeros((sn, sw + spad), dtype=np.uint8)
            host[:, :sw] = src[s0:s0 + sn]
            dsrc = torch.from_numpy(host).cuda()
            ddst = torch.full((n, dw + 8), 0xA5, dtype=torch.uint8, device="cuda")
            r.resize_band(y0, n, s0, sn, sw + spad, dsrc, dw + 8, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = ddst.cpu().numpy()
            if not (got[:, dw:] == 0xA5).all():
                return r.last_kernel(), False
            out[y0:y1] = got[:, :dw]
            k = r.last_kernel()
    return k, bool(np.array_equal(out, want)), (kind, deg, px, sw, sh, dw, dh, spad, cuts)


def yuv_layout(w, h):
    sx, sy = w + w % 2, h + h % 2
    return sx, sy, sx * sy, sx * sy // 4


def oracle_yuv(kind, deg, frame, sw, sh, dw, dh):
    """Per-plane oracle of one planar frame, as sample/resize_yuv420p.cpp:125-163 calls the classes."""
    sx, sy, ssy, ssu = yuv_layout(sw, sh)
    dx, dy, dsy, dsu = yuv_layout(dw, dh)
    out = np.zeros(dsy + 2 * dsu, dtype=np.uint8)
    rc, y = oracle_resize(kind, frame[:ssy].reshape(sy, sx)[:sh], dw, dh, deg, 1, sw=sw, dst_stride=dx)
    if rc:
        return None
    out[:dh * dx] = y.ravel()
    for p in range(2):
        plane = frame[ssy + p * ssu: ssy + (p + 1) * ssu].reshape(sy // 2, sx // 2)
        rc, c = oracle_resize(kind, plane, dx // 2, dy // 2, deg, 2, dst_stride=dx // 2)
        if rc:
            return None
        out[dsy + p * dsu: dsy + (p + 1) * dsu] = c.ravel()
    return out


def yuv_case(rng):
    fam = rng.randint(0, 4)
    if fam == 0:
        dw, dh = int(rng.randint(16, 300)) * 2, int(rng.randint(16, 200)) * 2
        return (LANCZOS, int(rng.choice([2, 3])), 2 * dw, 2 * dh, dw, dh)
    if fam == 1:
        sw, sh, dw, dh = [int(v) for v in rng.randint(40, 500, 4)]
        return (LANCZOS, int(rng.choice([1, 2, 3])), sw, sh, dw, dh)
    if fam == 2:
        sw, sh = int(rng.randint(16, 300)), int(rng.randint(16, 200))
        return (AREA, 0, sw, sh, int(rng.randint(8, sw + 1)), int(rng.randint(8, sh + 1)))
    sw, sh = int(rng.randint(8, 150)), int(rng.randint(8, 100))
    return (LINEAR, 0, sw, sh, int(rng.randint(sw, 3 * sw + 1)), int(rng.randint(sh, 3 * sh + 1)))


def run_yuv(rng, case):
    import torch
    kind, deg, sw, sh, dw, dh = case
    n = int(rng.randint(1, 4))
    try:
        r = iqo.Yuv420Resizer(kind, deg, sw, sh, dw, dh)
    except iqo.IqoCudaError:
        return None
    with r:
        src = np.frombuffer(rng.bytes(n * r.src_frame_bytes), dtype=np.uint8).reshape(n, r.src_frame_bytes).copy()
        want = [oracle_yuv(kind, deg, src[f], sw, sh, dw, dh) for f in range(n)]
        if any(w is None for w in want):
            return None
        if rng.rand() < 0.5:
            dsrc = torch.from_numpy(src).cuda()
            ddst = torch.zeros((n, r.dst_frame_bytes), dtype=torch.uint8, device="cuda")
            r.resize(n, dsrc, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = ddst.cpu().numpy()
        else:
            got = np.zeros((n, r.dst_frame_bytes), dtype=np.uint8)
            r.resize(n, src, got)
    dx, dy, dsy, dsu = yuv_layout(dw, dh)
    ok = True
    for f in range(n):
        # the even-rounded padding row / column of odd sizes is not written by the reference either: compare the planes
        g, w = got[f], want[f]
        ok = ok and np.array_equal(g[:dh * dx].reshape(dh, dx)[:, :dw], w[:dh * dx].reshape(dh, dx)[:, :dw])
        for p in range(2):
            ok = ok and np.array_equal(g[dsy + p * dsu: dsy + (p + 1) * dsu], w[dsy + p * dsu: dsy + (p + 1) * dsu])
    return "yuv420", bool(ok)
