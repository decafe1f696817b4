"""ctypes binding of ``libmfbiclust_cuda.so`` (the C ABI in ``include/mfbiclust.h``).

This is the Python twin of the R ``.Call`` shim described in INTEGRATION.md: pointer
extraction and result allocation only -- no arithmetic, and NO CPU fallback.  If the
shared library is missing, or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (MFBICLUST_LIB: another build of the same library, for A/B runs during development)
LIB_PATH = os.environ.get("MFBICLUST_LIB") or os.path.join(_HERE, "lib", "libmfbiclust_cuda.so")

MFB_OK, MFB_SINGULAR, MFB_ZERO_ROW, MFB_NOT_CONVERGED = 0, 1, 2, 3
MFB_ERR_ARG, MFB_ERR_CUDA, MFB_ERR_ALLOC, MFB_ERR_NEGATIVE, MFB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> (restype, argtypes); mirrors include/mfbiclust.h declaration by declaration
PROTOTYPES = {
    "mfb_version": (C.c_int, []),
    "mfb_last_error": (C.c_char_p, []),
    "mfb_device_count": (C.c_int, []),
    "mfb_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mfb_context_destroy": (None, [_vp]),
    "mfb_context_sync": (C.c_int, [_vp]),
    "mfb_context_stream": (_vp, [_vp]),
    "mfb_context_trim": (C.c_int, [_vp]),
    "mfb_context_set_concurrency": (C.c_int, [_vp, C.c_int]),
    "mfb_matrix_upload": (C.c_int, [_vp, _dp, _i64, _i64, C.POINTER(_vp)]),
    "mfb_matrix_upload_f32": (C.c_int, [_vp, _dp, _i64, _i64, C.POINTER(_vp)]),
    "mfb_matrix_wrap_dev": (C.c_int, [_vp, _vp, _i64, _i64, C.POINTER(_vp)]),
    "mfb_matrix_view": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "mfb_matrix_free": (None, [_vp]),
    "mfb_matrix_stats": (C.c_int, [_vp, _dp]),
    "mfb_matrix_shift": (C.c_int, [_vp, C.c_double]),
    "mfb_matrix_dev_ptr": (_vp, [_vp]),
    "mfb_als_begin": (C.c_int, [_vp, _vp, C.c_int, _dp, _dp, C.c_double, C.POINTER(_vp)]),
    "mfb_als_begin_sharded": (C.c_int, [_vp, _vp, _i64, C.c_int, _dp, _dp, C.c_double, _vp, _vp, C.POINTER(_vp)]),
    "mfb_als_shard_export": (C.c_int, [_vp, _vp]),
    "mfb_als_shard_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "mfb_als_shard_disconnect": (C.c_int, [_vp]),
    "mfb_als_iterate": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, _ip, _ip]),
    "mfb_als_iterate_draws": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, _ip, _ip, C.POINTER(C.c_longlong)]),
    "mfb_als_restart": (C.c_int, [_vp, _dp]),
    "mfb_als_profile": (C.c_int, [_vp, C.c_int, _dp]),
    "mfb_als_result": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "mfb_als_end": (None, [_vp]),
    "mfb_als_run": (C.c_int, [_vp, _vp, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_double, C.c_int,
                              _dp, _dp, _dp, _ip, _ip, _dp]),
    "mfb_als_run_host": (C.c_int, [_vp, _dp, _i64, _i64, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_int,
                                   _dp, _dp, _dp, _ip, _ip, _dp]),
    "mfb_snmf_begin": (C.c_int, [_vp, _vp, C.c_int, _dp, C.c_double, C.c_double, C.POINTER(_vp)]),
    "mfb_snmf_iterate": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip, _dp]),
    "mfb_snmf_restart": (C.c_int, [_vp, _dp]),
    "mfb_nnls": (C.c_int, [_vp, _dp, _dp, _i64, C.c_int, _i64, _dp]),
    "mfb_nnls_draws": (C.c_int, [_vp, _dp, _dp, _i64, C.c_int, _i64, _dp, C.POINTER(C.c_longlong)]),
    "mfb_otsu": (C.c_int, [_vp, _dp, _i64, _i64, _dp]),
    "mfb_threshold": (C.c_int, [_vp, _dp, _i64, _i64, _dp, C.c_int, _i32p]),
    "mfb_bicluster_overlap": (C.c_int, [_vp, _i32p, _i64, _i32p, _i64, C.c_int, _dp, _dp]),
    "mfb_bicluster_union": (C.c_int, [_vp, _i32p, _i64, _i32p, _i64, C.c_int, _i32p, _dp]),
    "mfb_svd_pca": (C.c_int, [_vp, _vp, C.c_int, _dp, _dp, _ip]),
    "mfb_nipals": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                             _dp, _dp, _dp, _ip]),
    "mfb_bcv_cells": (C.c_int, [_vp, _vp, _i32p, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, _dp]),
    "mfb_bcv_cells_multi": (C.c_int, [_vp, _vp, C.c_int, _i32p, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, _dp]),
    "mfb_bcv_cells_na": (C.c_int, [_vp, _vp, C.c_int, _i32p, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, _dp]),
}


# int (*mfb_allreduce_fn)(void *user, double *dev_buf, int64_t count, void *stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, _i64, _vp)


class MfbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libmfbiclust_cuda status {status}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (built in-tree by ``mfbiclust_b200/build.py``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mfbiclust_b200.build` "
                "(nvcc, sm_100a). mfbiclust_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)       # raises AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> int:
    if status < 0:
        raise MfbError(status, lib().mfb_last_error().decode("utf-8", "replace"))
    return status


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def as_f(a, dtype=np.float64) -> np.ndarray:
    """Column-major (R layout) contiguous copy/view."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))


class Context:
    """``mfb_context``: one device, one stream.  Created lazily per process/device."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        check(lib().mfb_context_create(device, C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().mfb_context_sync(self._h))

    def stream(self) -> int:
        return int(lib().mfb_context_stream(self._h) or 0)

    def set_concurrency(self, n: int):
        """``n`` contexts work on this device at the same time (grid sizing hint; results do not depend on it)."""
        check(lib().mfb_context_set_concurrency(self._h, int(n)))

    def trim(self):
        """Return the cached device buffers of the workspace pool to the driver."""
        check(lib().mfb_context_trim(self._h))

    def close(self):
        if self._h:
            lib().mfb_context_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("MFBICLUST_DEVICE", "0")))
        n = lib().mfb_device_count()
        if n > 0:
            device %= n
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Matrix:
    """``mfb_matrix``: an FP64 matrix resident in HBM."""

    def __init__(self, ctx: Context, host: np.ndarray | None = None, dev_ptr: int | None = None, shape=None,
                 precision: str = "f64"):
        self.ctx = ctx
        self._h = _vp()
        self.precision = precision
        if host is not None:
            a = as_f(host)
            self.shape = a.shape
            up = lib().mfb_matrix_upload_f32 if precision == "f32" else lib().mfb_matrix_upload
            check(up(ctx.handle, dptr(a), a.shape[0], a.shape[1], C.byref(self._h)))
        else:
            self.shape = tuple(shape)
            check(lib().mfb_matrix_wrap_dev(ctx.handle, _vp(dev_ptr), shape[0], shape[1], C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def view(self, ctx: "Context") -> "Matrix":
        """Non-owning alias for another context on the same device (``mfb_matrix_view``)."""
        v = Matrix.__new__(Matrix)
        v.ctx, v.shape, v.precision, v._h = ctx, self.shape, self.precision, _vp()
        v._owner = self                       # keep the owner alive
        check(lib().mfb_matrix_view(ctx.handle, self._h, C.byref(v._h)))
        return v

    def stats(self):
        out = np.zeros(5)
        check(lib().mfb_matrix_stats(self._h, dptr(out)))
        return dict(min=out[0], max=out[1], sumsq=out[2], n_nan=int(out[3]), sum=out[4])

    def shift(self, s: float):
        check(lib().mfb_matrix_shift(self._h, float(s)))

    def free(self):
        if self._h:
            lib().mfb_matrix_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
