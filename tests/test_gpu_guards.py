"""Stand-ins for compute-sanitizer, which is closed on this GPU pool (profiles/r2_compute_sanitizer_closed.txt):
 * write guards (memcheck's job for stores): every kernel family writes into a destination that is surrounded by
   canary rows and canary columns; only the dstW x dstH pixels may change;
 * repeatability (racecheck's job): the warp-streaming and tensor-path kernels order their shared-memory traffic with
   cp.async groups, mbarriers, __syncwarp and CTA barriers instead of per-tile __syncthreads; a missing ordering shows
   up as run-to-run differences, so every family is launched repeatedly on a multi-CTA batch while other work runs on
   a second stream, and every result must equal the first one and the oracle."""
import numpy as np
import pytest

import libiqo_b200 as iqo
from oracle_lib import AREA, LANCZOS, LINEAR, lcg_image, oracle_resize

pytestmark = pytest.mark.gpu

FAMILIES = [
    # (kind, degree, pxScale, srcW, srcH, dstW, dstH, path, expected kernel)
    (LANCZOS, 3, 1, 960, 540, 480, 270, iqo.PATH_STREAM, "half_sym_stream"),
    (LANCZOS, 3, 1, 960, 540, 480, 270, iqo.PATH_NO_STREAM, "half_sym_tma"),
    (LANCZOS, 3, 1, 960, 540, 480, 270, iqo.PATH_NO_TMA, "half_sym"),
    (LANCZOS, 2, 2, 960, 540, 480, 270, iqo.PATH_AUTO, "half_small"),
    (LANCZOS, 3, 1, 960, 540, 640, 360, iqo.PATH_STREAM, "ratio_stream"),
    (LANCZOS, 4, 1, 1024, 512, 375, 188, iqo.PATH_STREAM, "lanczos_stream"),
    (LANCZOS, 4, 1, 1024, 512, 375, 188, iqo.PATH_MMA, "lanczos_mma"),
    (LANCZOS, 3, 1, 960, 540, 640, 360, iqo.PATH_MMA, "lanczos_mma_dp2a"),
    (LANCZOS, 3, 1, 960, 540, 480, 270, iqo.PATH_MMA, "lanczos_mma"),
    (LANCZOS, 4, 1, 1024, 512, 375, 188, iqo.PATH_NO_STREAM, "packed"),
    (LANCZOS, 3, 1, 333, 222, 200, 150, iqo.PATH_GENERIC, "generic"),
    (AREA, 0, 1, 960, 540, 480, 270, iqo.PATH_AUTO, "area2"),
    (AREA, 0, 1, 960, 540, 640, 360, iqo.PATH_AUTO, "area_down"),
    (AREA, 0, 1, 960, 540, 600, 360, iqo.PATH_AUTO, "packed"),
    (LINEAR, 0, 1, 320, 180, 960, 540, iqo.PATH_AUTO, "linear_up3"),
    (LINEAR, 0, 1, 320, 180, 480, 270, iqo.PATH_AUTO, "linear_up_2_3"),
    (LINEAR, 0, 1, 320, 180, 700, 400, iqo.PATH_AUTO, "packed"),
]


@pytest.mark.parametrize("fam", FAMILIES, ids=["%s-%d" % (f[8], i) for i, f in enumerate(FAMILIES)])
def test_write_guards_and_repeatability(fam):
    torch = pytest.importorskip("torch")
    kind, deg, px, sw, sh, dw, dh, path, kname = fam
    n, reps, guard, pad = 6, 12, 3, 16
    host = np.stack([lcg_image(sh, sw, seed=90 + f) for f in range(n)])
    want = np.stack([oracle_resize(kind, host[f], dw, dh, deg, px)[1] for f in range(n)])
    dsrc = torch.from_numpy(host).cuda()
    dpitch = dw + pad
    frame = (dh + 2 * guard) * dpitch
    side = torch.cuda.Stream()
    noise = torch.zeros((64, 1024, 1024), dtype=torch.uint8, device="cuda")
    first = None
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(path)
        for rep in range(reps):
            ddst = torch.full((n, dh + 2 * guard, dpitch), 0xC3, dtype=torch.uint8, device="cuda")
            with torch.cuda.stream(side):      # unrelated traffic on another stream while the kernel runs
                noise.add_(1)
            r.resize_batch(n, sw, sw * sh, dsrc, dpitch, frame, ddst[:, guard:], torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert r.last_kernel() == kname
            got = ddst.cpu().numpy()
            assert (got[:, :guard] == 0xC3).all() and (got[:, guard + dh:] == 0xC3).all(), "rows outside the image were written"
            assert (got[:, :, dw:] == 0xC3).all(), "bytes beyond dstW were written"
            body = got[:, guard:guard + dh, :dw]
            if first is None:
                first = body.copy()
                bad = np.argwhere(first != want)
                assert bad.size == 0, (len(bad), bad[:6].tolist())
            else:
                assert np.array_equal(body, first), "run %d differs from run 0 (shared-memory ordering?)" % rep
