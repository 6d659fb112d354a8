"""ctypes loaders for the CHECKERS (test infrastructure, never the product):

* ``oracle/liboracle.so``                -- the C restatement (oracle/iqo_oracle.c)
* ``oracle/_ref/libiqo_ref_generic.so``  -- the reference's own Generic code, compiled from
  /root/reference by oracle/Makefile (present when it was built in the dev container; it is
  git-ignored but travels to the GPU box)
* ``oracle/_ref/libiqo_ref_full.so``     -- the reference library with its SIMD dispatch + OpenMP
  (timed CPU baseline only)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

LANCZOS, AREA, LINEAR = 0, 1, 2
KIND_NAMES = {LANCZOS: "lanczos", AREA: "area", LINEAR: "linear"}

_sz = C.c_size_t
_u8p = C.c_void_p


def build_oracle():
    """(Re)build liboracle.so; also oracle/_ref when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


_oracle = None
_ref_generic = None
_ref_full = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "iqo_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        lib = C.CDLL(path)
        lib.iqo_oracle_resize.restype = C.c_int
        lib.iqo_oracle_resize.argtypes = [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, _sz, _u8p, _sz, _u8p]
        lib.iqo_oracle_resize_batch.restype = C.c_int
        lib.iqo_oracle_resize_batch.argtypes = [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz,
                                                _sz, _sz, _sz, _u8p, _sz, _sz, _u8p]
        lib.iqo_oracle_table.restype = C.c_int
        lib.iqo_oracle_table.argtypes = [C.c_int, C.c_int, C.c_uint, _sz, _sz, _sz,
                                         C.c_void_p, _sz, C.POINTER(_sz)]
        lib.iqo_oracle_fill_lcg.restype = None
        lib.iqo_oracle_fill_lcg.argtypes = [_u8p, _sz, C.c_uint32]
        lib.iqo_oracle_fnv1a.restype = C.c_uint64
        lib.iqo_oracle_fnv1a.argtypes = [_u8p, _sz, _sz, _sz]
        _oracle = lib
    return _oracle


def ref_generic():
    """The compiled reference Generic implementation, or None when not built."""
    global _ref_generic
    if _ref_generic is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "libiqo_ref_generic.so"))
        if lib is not None:
            lib.iqo_ref_generic_resize.restype = C.c_int
            lib.iqo_ref_generic_resize.argtypes = [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, _sz, _u8p, _sz, _u8p]
            lib.iqo_ref_num_coefs_lanczos.restype = _sz
            lib.iqo_ref_num_coefs_lanczos.argtypes = [C.c_int, _sz, _sz, _sz]
            lib.iqo_ref_num_coefs_area.restype = _sz
            lib.iqo_ref_num_coefs_area.argtypes = [_sz, _sz]
        _ref_generic = lib
    return _ref_generic


def ref_full():
    """The reference library with SIMD dispatch and OpenMP (CPU baseline timing), or None."""
    global _ref_full
    if _ref_full is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "libiqo_ref_full.so"))
        if lib is not None:
            lib.iqo_ref_public_new.restype = C.c_void_p
            lib.iqo_ref_public_new.argtypes = [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz]
            lib.iqo_ref_public_resize.restype = None
            lib.iqo_ref_public_resize.argtypes = [C.c_void_p, _sz, _u8p, _sz, _u8p]
            lib.iqo_ref_public_resize_batch.restype = None
            lib.iqo_ref_public_resize_batch.argtypes = [C.c_void_p, _sz, _sz, _sz, _u8p, _sz, _sz, _u8p]
            lib.iqo_ref_public_delete.restype = None
            lib.iqo_ref_public_delete.argtypes = [C.c_void_p]
            lib.iqo_ref_threads.restype = C.c_int
            lib.iqo_ref_set_threads.restype = None
            lib.iqo_ref_set_threads.argtypes = [C.c_int]
            lib.iqo_ref_generic_resize.restype = C.c_int
            lib.iqo_ref_generic_resize.argtypes = [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, _sz, _u8p, _sz, _u8p]
        _ref_full = lib
    return _ref_full


def lcg_image(h, w, seed=1, stride=None):
    """SURVEY 8c generator.  Returns a (h, stride) uint8 array whose rows hold w valid bytes."""
    stride = stride or w
    buf = np.empty(h * stride, dtype=np.uint8)
    oracle().iqo_oracle_fill_lcg(buf.ctypes.data, buf.size, seed)
    return buf.reshape(h, stride)


def fnv1a(img, w=None):
    img = np.ascontiguousarray(img)
    h, st = img.shape
    return int(oracle().iqo_oracle_fnv1a(img.ctypes.data, w or st, h, st))


def _run(fn, kind, degree, src, sw, dw, dh, px, dst_stride):
    src = np.ascontiguousarray(src)
    sh, sst = src.shape
    dst_stride = dst_stride or dw
    dst = np.full((dh, dst_stride), 0xA5, dtype=np.uint8)
    rc = fn(kind, degree, sw, sh, dw, dh, px, sst, src.ctypes.data, dst_stride, dst.ctypes.data)
    return rc, dst


def oracle_resize(kind, src, dw, dh, degree=0, px=1, sw=None, dst_stride=None):
    """Run the C restatement.  src is (srcH, srcStride) uint8; returns (rc, dst (dstH, dstStride))."""
    sw = sw or src.shape[1]
    return _run(oracle().iqo_oracle_resize, kind, degree, src, sw, dw, dh, px, dst_stride)


def ref_resize(kind, src, dw, dh, degree=0, px=1, sw=None, dst_stride=None):
    """Run the compiled reference Generic code (requires oracle/_ref)."""
    sw = sw or src.shape[1]
    return _run(ref_generic().iqo_ref_generic_resize, kind, degree, src, sw, dw, dh, px, dst_stride)


def oracle_table(kind, axis, S, D, degree=0, px=1):
    cap = 1 << 20
    buf = np.zeros(cap, dtype=np.int32)
    nt = _sz(0)
    n = oracle().iqo_oracle_table(kind, axis, degree, S, D, px, buf.ctypes.data, cap, C.byref(nt))
    assert n > 0, n
    return buf[: n * nt.value].reshape(nt.value, n).copy()
