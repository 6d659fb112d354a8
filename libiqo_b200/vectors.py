"""The test-vector conventions of SURVEY section 8c in numpy / plain Python: the LCG that fills
source images and the FNV-1a-64 hash the golden vectors (tests/golden/cases.json, recorded from
the reference's Generic code) are stated in.  Independent of everything under oracle/, so that
bench.py can check the frames it timed against the committed golden hashes."""
import json
import os

import numpy as np

_A, _C, _M32 = 1664525, 1013904223, 0xFFFFFFFF
_FNV_BASIS, _FNV_PRIME, _M64 = 0xCBF29CE484222325, 0x100000001B3, 0xFFFFFFFFFFFFFFFF


def lcg_bytes(n, seed=1):
    """x = seed; per byte: x = x * 1664525 + 1013904223 (mod 2^32), byte = x >> 24.  Built by doubling:
    elements [m, 2m) are the m-step map of elements [0, m)."""
    xs = np.empty(n, dtype=np.uint64)
    if n == 0:
        return xs.astype(np.uint8)
    xs[0] = (seed * _A + _C) & _M32
    am, cm, m = _A, _C, 1   # the m-step map x -> am * x + cm
    while m < n:
        k = min(m, n - m)
        xs[m:m + k] = (xs[:k] * np.uint64(am) + np.uint64(cm)) & np.uint64(_M32)
        am, cm = (am * am) & _M32, (am * cm + cm) & _M32
        m *= 2
    return (xs >> np.uint64(24)).astype(np.uint8)


def lcg_image(h, w, seed=1):
    """Contiguous (h, w) uint8 image of SURVEY 8c."""
    return lcg_bytes(h * w, seed).reshape(h, w)


def fnv1a64(img):
    """FNV-1a 64 over the bytes of a contiguous array (plain Python loop: ~4 MB/s, meant for single frames)."""
    h = _FNV_BASIS
    for b in np.ascontiguousarray(img).tobytes():
        h = ((h ^ b) * _FNV_PRIME) & _M64
    return h


def golden_hash(kind, degree, px_scale, sw, sh, dw, dh, seed=1):
    """Hash of the reference's Generic output for the seed's LCG source (contiguous), or None when
    tests/golden/cases.json holds no such case."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cases.json")
    for c in json.load(open(path))["large"]:
        if c[:10] == [kind, degree, px_scale, sw, sh, dw, dh, 0, 0, seed]:
            return c[10]
    return None
