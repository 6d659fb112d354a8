"""The test-vector conventions of SURVEY section 8c: the LCG that fills source images and the FNV-1a-64
hash the golden vectors (tests/golden/cases.json, recorded from the reference's Generic code) are stated in.
Independent of everything under oracle/, so that bench.py can check the frames it timed against the
committed golden hashes.  Two implementations of the same definitions: plain numpy / Python (always
available) and libiqo_b200/lib/libiqo_vectors.so (csrc/vectors.c), used when built because the
1 GiB source and the 144 MB result of BASELINE config 5 are out of reach of a Python byte loop."""
import ctypes as C
import json
import os

import numpy as np

_A, _C, _M32 = 1664525, 1013904223, 0xFFFFFFFF
_FNV_BASIS, _FNV_PRIME, _M64 = 0xCBF29CE484222325, 0x100000001B3, 0xFFFFFFFFFFFFFFFF

_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def _native():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "lib", "libiqo_vectors.so")
        if not os.path.exists(path):
            _clib = False
        else:
            lib = C.CDLL(path)
            lib.iqo_vec_fill_lcg.restype = None
            lib.iqo_vec_fill_lcg.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64]
            lib.iqo_vec_fnv1a64.restype = C.c_uint64
            lib.iqo_vec_fnv1a64.argtypes = [C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
            _clib = lib
    return _clib or None


def _jump(seed, offset):
    """State after `offset` steps of x -> A x + C (mod 2^32)."""
    a, c, x = _A, _C, seed & _M32
    k = offset
    while k:
        if k & 1:
            x = (a * x + c) & _M32
        c = ((a + 1) * c) & _M32
        a = (a * a) & _M32
        k >>= 1
    return x


def lcg_bytes_numpy(n, seed=1, offset=0):
    """x = seed; per byte: x = x * 1664525 + 1013904223 (mod 2^32), byte = x >> 24; bytes [offset, offset + n).
    Built by doubling: elements [m, 2m) are the m-step map of elements [0, m)."""
    xs = np.empty(n, dtype=np.uint64)
    if n == 0:
        return xs.astype(np.uint8)
    xs[0] = (_jump(seed, offset) * _A + _C) & _M32
    am, cm, m = _A, _C, 1   # the m-step map x -> am * x + cm
    while m < n:
        k = min(m, n - m)
        xs[m:m + k] = (xs[:k] * np.uint64(am) + np.uint64(cm)) & np.uint64(_M32)
        am, cm = (am * am) & _M32, (am * cm + cm) & _M32
        m *= 2
    return (xs >> np.uint64(24)).astype(np.uint8)


def lcg_fill(out, seed=1, offset=0):
    """Fill the contiguous uint8 numpy array `out` with bytes [offset, offset + out.size) of the stream."""
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    lib = _native()
    if lib is not None:
        lib.iqo_vec_fill_lcg(out.ctypes.data, out.size, seed & _M32, offset)
    else:
        out.reshape(-1)[:] = lcg_bytes_numpy(out.size, seed, offset)
    return out


def lcg_bytes(n, seed=1, offset=0):
    return lcg_fill(np.empty(n, dtype=np.uint8), seed, offset)


def lcg_image(h, w, seed=1, row0=0):
    """Contiguous (h, w) uint8 image of SURVEY 8c; row0 > 0 gives rows [row0, row0 + h) of a taller image."""
    return lcg_bytes(h * w, seed, row0 * w).reshape(h, w)


def fnv1a64_python(img, h=_FNV_BASIS):
    """FNV-1a 64 over the bytes of a contiguous array (plain Python loop: ~4 MB/s)."""
    for b in np.ascontiguousarray(img).tobytes():
        h = ((h ^ b) * _FNV_PRIME) & _M64
    return h


def fnv1a64(img, h=_FNV_BASIS):
    """FNV-1a 64 over the bytes of an array, continued from `h` (chain calls to hash an image band by band)."""
    img = np.ascontiguousarray(img)
    lib = _native()
    if lib is None:
        return fnv1a64_python(img, h)
    return int(lib.iqo_vec_fnv1a64(h, img.ctypes.data, img.size, 1, img.size))


def _cases():
    path = os.path.join(os.path.dirname(_HERE), "tests", "golden", "cases.json")
    with open(path) as f:
        return json.load(f)


def golden_hash(kind, degree, px_scale, sw, sh, dw, dh, seed=1):
    """Hash of the reference's Generic output for the seed's LCG source (contiguous), or None when
    tests/golden/cases.json holds no such case."""
    meta = _cases()
    for c in meta["large"] + meta.get("huge", []):
        if c[:10] == [kind, degree, px_scale, sw, sh, dw, dh, 0, 0, seed]:
            return c[10]
    return None
