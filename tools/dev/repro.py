"""Reproduce fuzz mismatches: case tuples on the command line are run under every path and compared with the oracle."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libiqo_b200 as iqo
from oracle_lib import lcg_image, oracle_resize

CASES = [(0, 5, 1, 614, 411, 401, 342, 0, 0), (0, 5, 1, 332, 289, 636, 285, 0, 0)]
PATHS = [("auto", iqo.PATH_AUTO), ("stream", iqo.PATH_STREAM), ("mma", iqo.PATH_MMA), ("no_stream", iqo.PATH_NO_STREAM), ("generic", iqo.PATH_GENERIC)]
for case in CASES:
    kind, deg, px, sw, sh, dw, dh, spad, dpad = case
    for seed in (1, 7, 12345):
        for fill in (None, 255, 0):
            src = lcg_image(sh, sw + spad, seed=seed) if fill is None else np.full((sh, sw + spad), fill, np.uint8)
            rc, want = oracle_resize(kind, src, dw, dh, deg, px, sw=sw, dst_stride=dw + dpad)
            line = "%s seed=%s fill=%s rc=%d:" % (case, seed, fill, rc)
            for name, path in PATHS:
                dst = np.full((dh, dw + dpad), 0xA5, dtype=np.uint8)
                with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
                    r.set_path(path)
                    r.resize(sw + spad, src, dw + dpad, dst)
                    k = r.last_kernel()
                bad = np.argwhere(dst != want)
                line += " %s[%s]=%d" % (name, k, len(bad))
                if len(bad):
                    y, x = bad[0]
                    line += "(first y=%d x=%d got=%d want=%d; rows %s cols %s)" % (y, x, dst[y, x], want[y, x], sorted(set(bad[:, 0].tolist()))[:6], sorted(set(bad[:, 1].tolist()))[:6])
            print(line, flush=True)
    for axis in (0, 1):
        q = iqo.plan_query(kind, deg, sw, sh, dw, dh, px, axis)
        co = q["coefs"]
        print(" axis", axis, "N", q["numCoefs"], "coef min/max", int(co.min()), int(co.max()), "pos sum max", int(np.clip(co, 0, None).sum(axis=1).max()), "neg sum min", int(np.clip(co, None, 0).sum(axis=1).min()))
