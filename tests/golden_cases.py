"""Access to the committed golden fixtures (tests/golden/, made by make_golden.py from the reference)."""
import json
import os

import numpy as np

from oracle_lib import lcg_image

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN_DIR, "cases.json")) as _f:
    _META = json.load(_f)
_SMALL = np.load(os.path.join(GOLDEN_DIR, "small_cases.npz"))

SMALL = [tuple(c) for c in _META["small"]]
LARGE = [tuple(c) for c in _META["large"]]


def case_id(c):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = c[:10]
    name = {0: "lanczos%d" % deg, 1: "area", 2: "linear"}[kind]
    return "%s-px%d-%dx%d-to-%dx%d-pad%d_%d-s%d" % (name, px, sw, sh, dw, dh, spad, dpad, seed)


def case_src(c):
    kind, deg, px, sw, sh, dw, dh, spad, dpad, seed = c[:10]
    return lcg_image(sh, sw + spad, seed=seed)


def small_expected(i):
    return _SMALL["dst%d" % i]
