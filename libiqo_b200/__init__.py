"""libiqo_b200 -- Python host mirror of libiqo's resizer classes over the CUDA C ABI.

The product is ``libiqo_b200/lib/libiqo_cuda.so`` (C ABI ``include/iqo_cuda.h`` + the C++
classes ``iqo::LanczosResizer`` / ``AreaResizer`` / ``LinearResizer``).  This module binds the
C ABI with ctypes and mirrors the reference's class interface
(reference include/libiqo/LanczosResizer.hpp:14-59, AreaResizer.hpp:14-56, LinearResizer.hpp:14-56):

    r = LanczosResizer(degree, srcW, srcH, dstW, dstH, pxScale=1)
    r.resize(srcSt, src, dstSt, dst)

``src`` / ``dst`` may be numpy arrays (host memory), torch tensors (host or CUDA), objects with
``__cuda_array_interface__`` or raw integer addresses.

There is NO CPU fallback: importing works anywhere, but the first use raises
``IqoCudaError`` if the shared library is missing or no CUDA device is usable.
"""
import ctypes as C
import os

__all__ = ["LanczosResizer", "AreaResizer", "LinearResizer", "Yuv420Resizer", "IqoCudaError", "lib", "build",
           "LANCZOS", "AREA", "LINEAR", "PATH_AUTO", "PATH_GENERIC", "PATH_NO_TMA", "PATH_NO_STREAM", "PATH_STREAM", "PATH_MMA", "PATH_NO_MMA", "ARITH_FIXED", "ARITH_SIMD_FLOAT", "exported_symbols"]

LANCZOS, AREA, LINEAR = 0, 1, 2
ARITH_FIXED, ARITH_SIMD_FLOAT = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_NO_TMA, PATH_NO_STREAM, PATH_STREAM, PATH_MMA, PATH_NO_MMA = 0, 1, 2, 3, 4, 5, 6

_HERE = os.path.dirname(os.path.abspath(__file__))
# IQO_CUDA_LIB: load another build of the library (A/B runs of two kernel variants in one gpurun call)
LIB_PATH = os.environ.get("IQO_CUDA_LIB") or os.path.join(_HERE, "lib", "libiqo_cuda.so")

_sz = C.c_size_t
_vp = C.c_void_p


class IqoCudaError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "iqo_cuda error %d: %s" % (code, message))
        self.code = code


_SIGNATURES = {
    "iqo_cuda_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz]),
    "iqo_cuda_create_on": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz]),
    "iqo_cuda_destroy": (None, [_vp]),
    "iqo_cuda_resize": (C.c_int, [_vp, _sz, _vp, _sz, _vp]),
    "iqo_cuda_resize_batch": (C.c_int, [_vp, _sz, _sz, _sz, _vp, _sz, _sz, _vp, _vp]),
    "iqo_cuda_resize_batch_host": (C.c_int, [_vp, _sz, _sz, _sz, _vp, _sz, _sz, _vp]),
    "iqo_cuda_resize_band": (C.c_int, [_vp, _sz, _sz, _sz, _sz, _sz, _vp, _sz, _vp, _vp]),
    "iqo_cuda_band_src_rows": (C.c_int, [_vp, _sz, _sz, C.POINTER(_sz), C.POINTER(_sz)]),
    "iqo_cuda_resize_bands_multi": (C.c_int, [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, _sz, _vp, _sz, _vp,
                                              C.c_int, C.POINTER(C.c_int)]),
    "iqo_cuda_resize_batch_multi": (C.c_int, [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, _sz,
                                              _sz, _sz, _vp, _sz, _sz, _vp, C.c_int, C.POINTER(C.c_int)]),
    "iqo_cuda_yuv420_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_uint, _sz, _sz, _sz, _sz]),
    "iqo_cuda_yuv420_destroy": (None, [_vp]),
    "iqo_cuda_yuv420_frame_bytes": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "iqo_cuda_yuv420_resize": (C.c_int, [_vp, _sz, _vp, _vp, _vp]),
    "iqo_cuda_get_table": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _vp, _sz]),
    "iqo_cuda_plan_query": (C.c_int, [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                      _vp, _sz, _vp, _vp, _sz]),
    "iqo_cuda_plan_kernel": (C.c_int, [C.c_int, C.c_uint, _sz, _sz, _sz, _sz, _sz, C.c_char_p, _sz, C.c_char_p, _sz]),
    "iqo_cuda_set_path": (C.c_int, [_vp, C.c_int]),
    "iqo_cuda_set_arithmetic": (C.c_int, [_vp, C.c_int]),
    "iqo_cuda_last_kernel": (C.c_char_p, [_vp]),
    "iqo_cuda_launch_count": (C.c_ulonglong, []),
    "iqo_cuda_sync": (C.c_int, [_vp]),
    "iqo_cuda_clear_cache": (None, []),
    "iqo_cuda_host_alloc": (_vp, [_sz]),
    "iqo_cuda_host_free": (None, [_vp]),
    "iqo_cuda_last_error": (C.c_char_p, []),
    "iqo_cuda_device_count": (C.c_int, []),
    "iqo_cuda_version": (C.c_char_p, []),
}


def exported_symbols():
    """Names include/iqo_cuda.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


_lib = None


def build(verbose=False):
    """Compile the shared library in-tree (nvcc, sm_100a).  No GPU needed."""
    import subprocess
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


def lib():
    """The loaded C ABI library.  Raises IqoCudaError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IqoCudaError(-5, "%s not found: build it with `make -C libiqo_b200/csrc` "
                               "(there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(rc):
    if rc != 0:
        raise IqoCudaError(rc, lib().iqo_cuda_last_error().decode())


def _address(buf):
    """Raw address of a numpy array / torch tensor / CUDA-array-interface object / int."""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):  # torch
        return buf.data_ptr()
    if hasattr(buf, "__cuda_array_interface__"):
        return buf.__cuda_array_interface__["data"][0]
    if hasattr(buf, "ctypes"):  # numpy
        return buf.ctypes.data
    if hasattr(buf, "__array_interface__"):
        return buf.__array_interface__["data"][0]
    return C.addressof(C.c_char.from_buffer(buf))


class _Resizer(object):
    _kind = None

    def __init__(self, degree, srcW, srcH, dstW, dstH, pxScale, device=None):
        self._h = _vp()
        self.srcW, self.srcH, self.dstW, self.dstH = srcW, srcH, dstW, dstH
        if device is None:
            rc = lib().iqo_cuda_create(C.byref(self._h), self._kind, degree, srcW, srcH, dstW, dstH, pxScale)
        else:
            rc = lib().iqo_cuda_create_on(C.byref(self._h), device, self._kind, degree, srcW, srcH, dstW, dstH, pxScale)
        _check(rc)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().iqo_cuda_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's interface --------------------------------------------------------
    def resize(self, srcSt, src, dstSt, dst):
        """resize(srcStride, src, dstStride, dst): strides in bytes, host or device buffers.

        Device buffers (torch CUDA tensors): the call runs on the legacy default stream, i.e. after
        everything torch has enqueued on its default current stream, and returns when dst is complete.
        If `src` is produced on a side stream created as non-blocking, synchronise it first or use
        resize_batch(1, ..., stream=that_stream)."""
        _check(lib().iqo_cuda_resize(self._h, srcSt, _address(src), dstSt, _address(dst)))

    # -- batched / sharded forms ----------------------------------------------------------
    def resize_batch(self, nFrames, srcSt, srcFrameStride, src, dstSt, dstFrameStride, dst, stream=None):
        """Device-resident frames, one asynchronous launch sequence on `stream` (int handle)."""
        _check(lib().iqo_cuda_resize_batch(self._h, nFrames, srcSt, srcFrameStride, _address(src),
                                           dstSt, dstFrameStride, _address(dst), stream))

    def resize_batch_host(self, nFrames, srcSt, srcFrameStride, src, dstSt, dstFrameStride, dst):
        """Host-resident frames through the pipelined H2D / kernel / D2H path (synchronous)."""
        _check(lib().iqo_cuda_resize_batch_host(self._h, nFrames, srcSt, srcFrameStride, _address(src),
                                                dstSt, dstFrameStride, _address(dst)))

    def band_src_rows(self, dstRow0, dstRows):
        a, b = _sz(0), _sz(0)
        _check(lib().iqo_cuda_band_src_rows(self._h, dstRow0, dstRows, C.byref(a), C.byref(b)))
        return a.value, b.value

    def resize_band(self, dstRow0, dstRows, srcRow0, srcRows, srcSt, src, dstSt, dst, stream=None):
        _check(lib().iqo_cuda_resize_band(self._h, dstRow0, dstRows, srcRow0, srcRows, srcSt, _address(src),
                                          dstSt, _address(dst), stream))

    # -- introspection --------------------------------------------------------------------
    def table(self, axis):
        """Integer coefficient table of axis 0 (X) / 1 (Y) as a list of rows."""
        n, t = C.c_int(0), C.c_int(0)
        _check(lib().iqo_cuda_get_table(self._h, axis, C.byref(n), C.byref(t), None, 0))
        buf = (C.c_int32 * (n.value * t.value))()
        _check(lib().iqo_cuda_get_table(self._h, axis, C.byref(n), C.byref(t), buf, len(buf)))
        return [list(buf[i * n.value:(i + 1) * n.value]) for i in range(t.value)]

    def set_path(self, path):
        _check(lib().iqo_cuda_set_path(self._h, path))

    def set_arithmetic(self, arithmetic):
        """ARITH_FIXED (default, the reference's Generic path, bit-exact) or ARITH_SIMD_FLOAT (optional float mode)."""
        _check(lib().iqo_cuda_set_arithmetic(self._h, arithmetic))

    def last_kernel(self):
        return lib().iqo_cuda_last_kernel(self._h).decode()

    def sync(self):
        _check(lib().iqo_cuda_sync(self._h))


class LanczosResizer(_Resizer):
    """iqo::LanczosResizer(degree, srcW, srcH, dstW, dstH, pxScale=1)"""
    _kind = LANCZOS

    def __init__(self, degree, srcW, srcH, dstW, dstH, pxScale=1, device=None):
        _Resizer.__init__(self, degree, srcW, srcH, dstW, dstH, pxScale, device)


class AreaResizer(_Resizer):
    """iqo::AreaResizer(srcW, srcH, dstW, dstH)"""
    _kind = AREA

    def __init__(self, srcW, srcH, dstW, dstH, device=None):
        _Resizer.__init__(self, 0, srcW, srcH, dstW, dstH, 1, device)


class LinearResizer(_Resizer):
    """iqo::LinearResizer(srcW, srcH, dstW, dstH)"""
    _kind = LINEAR

    def __init__(self, srcW, srcH, dstW, dstH, device=None):
        _Resizer.__init__(self, 0, srcW, srcH, dstW, dstH, 1, device)


class Yuv420Resizer(object):
    """Planar YUV420 frames in the layout of the reference's sample/resize_yuv420p.cpp:66-163
    (even-rounded strides; Y, U, V planes; chroma resized with pxScale 2)."""

    def __init__(self, kind, degree, srcW, srcH, dstW, dstH):
        self._h = _vp()
        _check(lib().iqo_cuda_yuv420_create(C.byref(self._h), kind, degree, srcW, srcH, dstW, dstH))
        a, b = _sz(0), _sz(0)
        _check(lib().iqo_cuda_yuv420_frame_bytes(self._h, C.byref(a), C.byref(b)))
        self.src_frame_bytes, self.dst_frame_bytes = a.value, b.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().iqo_cuda_yuv420_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def resize(self, nFrames, src, dst, stream=None):
        """Host buffers: synchronous pipelined path.  Device buffers: asynchronous on `stream`."""
        _check(lib().iqo_cuda_yuv420_resize(self._h, nFrames, _address(src), _address(dst), stream))


def make_resizer(kind, degree, srcW, srcH, dstW, dstH, pxScale=1, device=None):
    if kind == LANCZOS:
        return LanczosResizer(degree, srcW, srcH, dstW, dstH, pxScale, device)
    if kind == AREA:
        return AreaResizer(srcW, srcH, dstW, dstH, device)
    if kind == LINEAR:
        return LinearResizer(srcW, srcH, dstW, dstH, device)
    raise ValueError("unknown kind %r" % (kind,))


def resize_bands_multi(kind, degree, srcW, srcH, dstW, dstH, pxScale, srcSt, src, dstSt, dst, devices):
    """One host image, destination row bands (+ host-side halo) sharded over `devices`."""
    arr = (C.c_int * len(devices))(*devices)
    _check(lib().iqo_cuda_resize_bands_multi(kind, degree, srcW, srcH, dstW, dstH, pxScale,
                                             srcSt, _address(src), dstSt, _address(dst), len(devices), arr))


def resize_batch_multi(kind, degree, srcW, srcH, dstW, dstH, pxScale, nFrames,
                       srcSt, srcFrameStride, src, dstSt, dstFrameStride, dst, devices):
    """Host-resident frames sharded in contiguous blocks over `devices`."""
    arr = (C.c_int * len(devices))(*devices)
    _check(lib().iqo_cuda_resize_batch_multi(kind, degree, srcW, srcH, dstW, dstH, pxScale, nFrames,
                                             srcSt, srcFrameStride, _address(src),
                                             dstSt, dstFrameStride, _address(dst), len(devices), arr))


def plan_query(kind, degree, srcW, srcH, dstW, dstH, pxScale, axis):
    """Host-only planner view: dict with numCoefs, numTables, numRows, mainBegin, mainEnd,
    coefs (numRows x numCoefs), first[], row[] for axis 0 (X) / 1 (Y).  Needs no GPU."""
    import numpy as np
    n, t, rws = C.c_int(0), C.c_int(0), C.c_int(0)
    mb, me = C.c_longlong(0), C.c_longlong(0)
    args = (kind, degree, srcW, srcH, dstW, dstH, pxScale, axis)
    _check(lib().iqo_cuda_plan_query(*args, C.byref(n), C.byref(t), C.byref(rws), C.byref(mb), C.byref(me),
                                     None, 0, None, None, 0))
    D = dstH if axis else dstW
    coefs = np.zeros(rws.value * n.value, dtype=np.int32)
    first = np.zeros(D, dtype=np.int32)
    row = np.zeros(D, dtype=np.int32)
    _check(lib().iqo_cuda_plan_query(*args, None, None, None, None, None, coefs.ctypes.data, coefs.size,
                                     first.ctypes.data, row.ctypes.data, D))
    return dict(numCoefs=n.value, numTables=t.value, numRows=rws.value, mainBegin=mb.value, mainEnd=me.value,
                coefs=coefs.reshape(rws.value, n.value), first=first, row=row)


def plan_kernel(kind, degree, srcW, srcH, dstW, dstH, pxScale=1):
    """(kernel name AUTO would select for aligned buffers, reason when it is "generic").  Needs no GPU."""
    k, w = C.create_string_buffer(64), C.create_string_buffer(256)
    _check(lib().iqo_cuda_plan_kernel(kind, degree, srcW, srcH, dstW, dstH, pxScale, k, 64, w, 256))
    return k.value.decode(), w.value.decode()


def device_count():
    return lib().iqo_cuda_device_count()


def launch_count():
    return lib().iqo_cuda_launch_count()
