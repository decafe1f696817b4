"""Back-end functions with the reference's names and argument meaning.

These mirror the exported R functions of jonalim/mfBiclust that sit on the hot path
(``NAMESPACE:5-7,20-21,28-29,42``): ``als_nmf``, ``svd_pca``, ``nipals_pca``, ``bcv``,
``auto_bcv``, ``otsuHelper``, ``threshold``.  Bodies only marshal NumPy arrays into the C ABI
(``include/mfbiclust.h``); all arithmetic runs in ``libmfbiclust_cuda.so`` on the GPU.  Random
draws (``runif`` inits, ``sample`` folds, restarts) stay on the host exactly as they stay in R.
"""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass

import numpy as np

from . import _cabi
from ._cabi import MFB_OK, MFB_SINGULAR, MFB_ZERO_ROW, as_f, check, dptr, lib


# ---------------------------------------------------------------------------
# result carriers (R/helperClasses.R:10,19)
# ---------------------------------------------------------------------------
@dataclass
class genericFactorization:
    W: np.ndarray
    H: np.ndarray


@dataclass
class genericFit:
    fit: genericFactorization
    method: str


class HostRng:
    """Uniform source for inits and restarts: ``runif(n)``.  Any object with ``runif`` works
    (e.g. ``oracle.r_rng.RRng`` in the tests to replay an R session)."""

    def __init__(self, seed=None):
        self._g = np.random.default_rng(seed)

    def runif(self, n):
        return self._g.random(n)

    def permutation(self, n):
        return self._g.permutation(n)


def _normalised_random_W(rng, m, k):
    # R/bicluster.R:95-101
    W = np.asarray(rng.runif(m * k), dtype=np.float64).reshape((m, k), order="F")
    frob = np.sqrt((W ** 2).sum(axis=0))
    return np.asfortranarray(W / frob[None, :])


# ---------------------------------------------------------------------------
# als_nmf
# ---------------------------------------------------------------------------
class AlsSession:
    """One replicate of ``als_nmf``'s inner loop on the device (``mfb_als_*``)."""

    def __init__(self, A: _cabi.Matrix, k: int, W0, Hold0, resid_old0: float):
        self.A = A
        self.k = int(k)
        self.m, self.n = A.shape
        W0 = as_f(W0)
        Hold0 = as_f(Hold0)
        if W0.shape != (self.m, self.k) or Hold0.shape != (self.k, self.n):
            raise ValueError("W0 must be m x k and Hold0 k x n")
        self._h = C.c_void_p()
        check(lib().mfb_als_begin(A.ctx.handle, A.handle, self.k, dptr(W0), dptr(Hold0),
                                  float(resid_old0), C.byref(self._h)))
        self.iters = 0

    def iterate(self, max_new_iters, eps_conv=1e-4, fixed_iters=False):
        it, conv = C.c_int(0), C.c_int(0)
        st = check(lib().mfb_als_iterate(self._h, int(max_new_iters), float(eps_conv), int(bool(fixed_iters)),
                                         C.byref(it), C.byref(conv)))
        self.iters = it.value
        return st, it.value, bool(conv.value)

    def restart(self, Wnew):
        Wnew = as_f(Wnew)
        check(lib().mfb_als_restart(self._h, dptr(Wnew)))

    def result(self):
        W = np.zeros((self.m, self.k), order="F")
        H = np.zeros((self.k, self.n), order="F")
        resid = C.c_double(0.0)
        trace = np.full((max(self.iters, 1), 3), np.nan)
        check(lib().mfb_als_result(self._h, dptr(W), dptr(H), C.byref(resid), dptr(trace)))
        return W, H, resid.value, trace[:self.iters]

    def close(self):
        if self._h:
            lib().mfb_als_end(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def als_replicate(A, k, W0, Hold0, maxIter=100, eps_conv=1e-4, fixed_iters=False, rng=None, ctx=None,
                  verbose=False):
    """One replicate from a given start (``R/bicluster.R:118-203``), including R's ``restart()``
    protocol: on a failed ``.fcnnls`` or a zero row of H the host redraws W with ``rng`` and
    re-enters; after the 10th failure it stops with R's warning."""
    ctx = ctx or _cabi.default_context()
    own = not isinstance(A, _cabi.Matrix)
    Ad = _cabi.Matrix(ctx, host=A) if own else A
    try:
        st = Ad.stats()
        if st["min"] < 0:
            raise ValueError("Negative values are not allowed")                     # :76
        if eps_conv <= 0:
            raise ValueError("ALS-NMF: Invalid argument 'eps_conv' - value should be positive")   # :86
        m, n = Ad.shape
        sess = AlsSession(Ad, k, W0, Hold0, st["max"])                               # :105 residNormOld <- maxA
        try:
            nrestart = 0
            total = 0
            converged = False
            while total < maxIter:
                status, total, converged = sess.iterate(maxIter - total, eps_conv, fixed_iters)
                if status == MFB_OK:
                    break
                nrestart += 1                                                        # :124,138,153
                if nrestart >= 10:
                    warnings.warn("als_nmf: Factorization failed too many times "
                                  "[Computation stopped after the 9th restart]")
                    break
                if rng is None:
                    raise _cabi.MfbError(status, "ALS iteration failed (singular Gram / zero row of H) and no "
                                                 "rng was supplied for restart()")
                sess.restart(_normalised_random_W(rng, m, k))                        # :107-116
            W, H, resid, trace = sess.result()
        finally:
            sess.close()
    finally:
        if own:
            Ad.free()
    if verbose:
        for i, (r, dw, dh) in enumerate(trace, start=1):
            last = (i == len(trace))
            if (i % 10 == 0 or (last and converged)) and not np.isnan(r):
                print("Track:\tIter\tNorm\t\trms(dW)\t\trms(dH)")
                print("\t%d\t%f\t%f\t%f" % (i, r, dw, dh))
        if converged:
            print("Converged!")
    return dict(obj=resid, W=W, H=H, iters=total, converged=converged, trace=trace)


def als_nmf(A, k, reps=4, maxIter=100, eps_conv=1e-4, verbose=False, rng=None, inits=None, ctx=None, **_):
    """``als_nmf(A, k, reps = 4L, maxIter = 100L, eps_conv = 1e-4, verbose = FALSE, ...)``
    (``R/bicluster.R:70-216``) -> :class:`genericFit`.

    ``rng`` supplies ``runif`` (default: a fresh NumPy generator); ``inits`` optionally gives
    ``[(W0, Hold0), ...]`` per replicate ("identical inputs and initialisation")."""
    ctx = ctx or _cabi.default_context()
    A = as_f(A)
    if np.isnan(A).any():
        raise ValueError("als_nmf: NA values are not supported")
    if (A < 0).any():
        raise ValueError("Negative values are not allowed")
    if eps_conv <= 0:
        raise ValueError("ALS-NMF: Invalid argument 'eps_conv' - value should be positive")
    rng = rng or HostRng()
    m, n = A.shape
    Ad = _cabi.Matrix(ctx, host=A)
    try:
        sols = []
        for rep in range(int(reps)):
            if inits is not None:
                W0, Hold0 = inits[rep]
            else:
                W0 = _normalised_random_W(rng, m, k)                                               # :95-101
                Hold0 = np.asarray(rng.runif(k * n), dtype=np.float64).reshape((k, n), order="F")  # :104
            sols.append(als_replicate(Ad, k, W0, Hold0, maxIter, eps_conv, rng=rng, ctx=ctx, verbose=verbose))
        best = int(np.argmin([s["obj"] for s in sols]))                                            # :207
    finally:
        Ad.free()
    res = genericFit(genericFactorization(W=sols[best]["W"], H=sols[best]["H"]), "als-nmf")
    res.solutions = sols
    res.best = best
    return res


def nnls(Cmat, B, ctx=None):
    """``NMF::.fcnnls(x = C, y = B)[[1]]``: exact column-wise NNLS (unit-test entry point)."""
    ctx = ctx or _cabi.default_context()
    Cmat, B = as_f(Cmat), as_f(B)
    r, k = Cmat.shape
    p = B.shape[1]
    K = np.zeros((k, p), order="F")
    st = check(lib().mfb_nnls(ctx.handle, dptr(Cmat), dptr(B), r, k, p, dptr(K)))
    if st == MFB_SINGULAR:
        raise np.linalg.LinAlgError("system is computationally singular")
    return K


# ---------------------------------------------------------------------------
# otsuHelper / threshold
# ---------------------------------------------------------------------------
def otsuHelper(matrix, ctx=None):
    """``otsuHelper(matrix)`` (``R/BiclusterStrategy.R:442-456``): one threshold per column."""
    ctx = ctx or _cabi.default_context()
    M = as_f(matrix)
    if M.ndim == 1:
        M = as_f(M.reshape(-1, 1))
    th = np.zeros(M.shape[1])
    check(lib().mfb_otsu(ctx.handle, dptr(M), M.shape[0], M.shape[1], dptr(th)))
    return th


def threshold(m, th, MARGIN=2, ctx=None):
    """``threshold(m, th, MARGIN = 2)`` (``R/BiclusterStrategy.R:508-528``) -> boolean matrix."""
    ctx = ctx or _cabi.default_context()
    M = as_f(m)
    th = np.ascontiguousarray(np.asarray(th, dtype=np.float64).ravel())
    if MARGIN == 1:
        if th.size != M.shape[0]:
            raise ValueError("Length of th must equal nrow(m).")
    else:
        if th.size != M.shape[1]:
            raise ValueError("Length of th must equal ncol(m)")
    out = np.zeros(M.shape, dtype=np.int32, order="F")
    check(lib().mfb_threshold(ctx.handle, dptr(M), M.shape[0], M.shape[1], dptr(th), int(MARGIN),
                              out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out.astype(bool)
