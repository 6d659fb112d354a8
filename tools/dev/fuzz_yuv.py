"""One-off randomized parity fuzz of the YUV420 frame API (device and host frames) against the per-plane oracle."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import libiqo_b200 as iqo
from oracle_lib import oracle_resize, LANCZOS, AREA, LINEAR

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0


def layout(w, h):
    sx, sy = w + w % 2, h + h % 2
    return sx, sy, sx * sy, sx * sy // 4


def oracle(kind, deg, frame, sw, sh, dw, dh):
    sx, sy, ssy, ssu = layout(sw, sh)
    dx, dy, dsy, dsu = layout(dw, dh)
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


t0, cases, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    fam = rng.randint(0, 4)
    if fam == 0:
        dw, dh = int(rng.randint(16, 300)) * 2, int(rng.randint(16, 200)) * 2
        kind, deg, sw, sh = LANCZOS, int(rng.choice([2, 3])), 2 * dw, 2 * dh
    elif fam == 1:
        sw, sh, dw, dh = [int(v) for v in rng.randint(40, 500, 4)]
        kind, deg = LANCZOS, int(rng.choice([1, 2, 3]))
    elif fam == 2:
        sw, sh = int(rng.randint(16, 300)), int(rng.randint(16, 200))
        kind, deg, dw, dh = AREA, 0, int(rng.randint(8, sw + 1)), int(rng.randint(8, sh + 1))
    else:
        sw, sh = int(rng.randint(8, 150)), int(rng.randint(8, 100))
        kind, deg, dw, dh = LINEAR, 0, int(rng.randint(sw, 3 * sw + 1)), int(rng.randint(sh, 3 * sh + 1))
    n = int(rng.randint(1, 4))
    try:
        r = iqo.Yuv420Resizer(kind, deg, sw, sh, dw, dh)
    except iqo.IqoCudaError:
        continue
    with r:
        src = np.frombuffer(rng.bytes(n * r.src_frame_bytes), dtype=np.uint8).reshape(n, r.src_frame_bytes).copy()
        want = [oracle(kind, deg, src[f], sw, sh, dw, dh) for f in range(n)]
        if any(w is None for w in want):
            continue
        if rng.rand() < 0.5:
            dsrc = torch.from_numpy(src).cuda()
            ddst = torch.zeros((n, r.dst_frame_bytes), dtype=torch.uint8, device="cuda")
            r.resize(n, dsrc, ddst, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = ddst.cpu().numpy()
        else:
            got = np.zeros((n, r.dst_frame_bytes), dtype=np.uint8)
            r.resize(n, src, got)
    cases += 1
    dx, dy, dsy, dsu = layout(dw, dh)
    for f in range(n):
        # the even-rounded padding row / column of odd sizes is not written by the reference either: compare the planes
        g, w = got[f], want[f]
        ok = np.array_equal(g[:dh * dx].reshape(dh, dx)[:, :dw], w[:dh * dx].reshape(dh, dx)[:, :dw])
        for p in range(2):
            gp = g[dsy + p * dsu: dsy + (p + 1) * dsu]
            wp = w[dsy + p * dsu: dsy + (p + 1) * dsu]
            ok = ok and np.array_equal(gp, wp)
        if not ok:
            bad += 1
            print("MISMATCH", kind, deg, sw, sh, dw, dh, n)
            break
print("yuv cases:", cases, "mismatching:", bad)
