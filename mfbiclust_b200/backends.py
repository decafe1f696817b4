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

    def iterate_draws(self, max_new_iters, eps_conv=1e-4, fixed_iters=False):
        """:meth:`iterate` plus the number of uniforms ``NMF::.fcnnls`` would have drawn through ``max.col`` during
        these iterations (``mfb_als_iterate_draws``; -1 when unknown)."""
        it, conv, draws = C.c_int(0), C.c_int(0), C.c_longlong(0)
        st = check(lib().mfb_als_iterate_draws(self._h, int(max_new_iters), float(eps_conv), int(bool(fixed_iters)),
                                               C.byref(it), C.byref(conv), C.byref(draws)))
        self.iters = it.value
        return st, it.value, bool(conv.value), int(draws.value)

    def profile(self, iters):
        """Mean duration (ms) of the four kernel launches of an iteration over ``iters`` further fixed
        iterations: ``[H stream, H solve, W stream, W solve]`` (``mfb_als_profile``); for a row-sharded session
        ``[H stream, fold, invert, H solve, W stream, W solve, finish]`` (callback exchange) or
        ``[H stream, fold + wait + invert, H solve, W stream, W solve + wait + finish]`` (peer memory)."""
        ms = np.zeros(8)
        check(lib().mfb_als_profile(self._h, int(iters), dptr(ms)))
        self.iters += int(iters)
        if isinstance(self, ShardedAlsSession):
            return ms[:5] if self.exchange == "peer" else ms[:7]
        return ms[:4]

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
                  verbose=False, follow_rng=False):
    """One replicate from a given start (``R/bicluster.R:118-203``), including R's ``restart()``
    protocol: on a failed ``.fcnnls`` or a zero row of H the host redraws W with ``rng`` and
    re-enters; after the 10th failure it stops with R's warning.

    ``follow_rng``: also consume from ``rng`` the uniforms ``NMF::.fcnnls`` would have consumed through
    ``max.col(ties.method = "random")`` (their number is reported by the device), so that a
    restart W -- and whatever the caller draws next -- comes from the stream position the reference would be at.
    The result carries ``maxcol_draws``."""
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
            draws_total = 0
            while total < maxIter:
                if follow_rng:
                    status, total, converged, draws = sess.iterate_draws(maxIter - total, eps_conv, fixed_iters)
                    if draws < 0:
                        raise _cabi.MfbError(_cabi.MFB_ERR_UNSUPPORTED, "max.col draw counts are not available for this rank")
                    draws_total += draws
                    if draws and rng is not None:
                        rng.runif(draws)                                             # the uniforms max.col consumed
                else:
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
    return dict(obj=resid, W=W, H=H, iters=total, converged=converged, trace=trace, maxcol_draws=draws_total)


def als_nmf(A, k, reps=4, maxIter=100, eps_conv=1e-4, verbose=False, rng=None, inits=None, ctx=None,
            precision="f64", follow_rng=False, **_):
    """``als_nmf(A, k, reps = 4L, maxIter = 100L, eps_conv = 1e-4, verbose = FALSE, ...)``
    (``R/bicluster.R:70-216``) -> :class:`genericFit`.

    ``rng`` supplies ``runif`` (default: a fresh NumPy generator); ``inits`` optionally gives
    ``[(W0, Hold0), ...]`` per replicate ("identical inputs and initialisation").  ``follow_rng = True`` (ranks up
    to 32) additionally consumes the uniforms the reference's ``.fcnnls`` calls would have consumed, so that with an
    R-compatible ``rng`` (``set.seed``) replicates 2.. and restarts start where they start in R.  ``precision = "f32"`` keeps
    the resident copy of A in single precision (the R side would read ``options(mfBiclust.precision=)``): half
    the HBM traffic per iteration, factors within ~1e-7 of the FP64 path (everything but A's storage stays FP64)."""
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
    Ad = _cabi.Matrix(ctx, host=A, precision=precision)
    try:
        sols = []
        for rep in range(int(reps)):
            if inits is not None:
                W0, Hold0 = inits[rep]
            else:
                W0 = _normalised_random_W(rng, m, k)                                               # :95-101
                Hold0 = np.asarray(rng.runif(k * n), dtype=np.float64).reshape((k, n), order="F")  # :104
            sols.append(als_replicate(Ad, k, W0, Hold0, maxIter, eps_conv, rng=rng, ctx=ctx, verbose=verbose,
                                      follow_rng=follow_rng))
        best = int(np.argmin([s["obj"] for s in sols]))                                            # :207
    finally:
        Ad.free()
    res = genericFit(genericFactorization(W=sols[best]["W"], H=sols[best]["H"]), "als-nmf")
    res.solutions = sols
    res.best = best
    return res


# ---------------------------------------------------------------------------
# snmf (R/bicluster.R:341-441) and the NMF::nmf(method = "snmf/r") run it delegates to
# ---------------------------------------------------------------------------
class SnmfSession(AlsSession):
    """One ``NMF::nmf(method = "snmf/r")`` run on the device (``mfb_snmf_*``; an ALS session with regularised Gram
    matrices and NMF's convergence test)."""

    def __init__(self, A: _cabi.Matrix, k: int, W0, beta: float, eta: float):
        self.A = A
        self.k = int(k)
        self.m, self.n = A.shape
        W0 = as_f(W0)
        if W0.shape != (self.m, self.k):
            raise ValueError("W0 must be m x k")
        self._h = C.c_void_p()
        check(lib().mfb_snmf_begin(A.ctx.handle, A.handle, self.k, dptr(W0), float(beta), float(eta), C.byref(self._h)))
        self.iters = 0

    def iterate(self, max_new_iters, bi_conv=(0, 10), eps_conv=1e-4):
        it, conv, obj = C.c_int(0), C.c_int(0), C.c_double(float("nan"))
        st = check(lib().mfb_snmf_iterate(self._h, int(max_new_iters), int(bi_conv[0]), int(bi_conv[1]), float(eps_conv),
                                          C.byref(it), C.byref(conv), C.byref(obj)))
        self.iters = it.value
        return st, it.value, bool(conv.value), obj.value

    def restart(self, Wnew):
        Wnew = as_f(Wnew)
        check(lib().mfb_snmf_restart(self._h, dptr(Wnew)))

    def result(self):
        W = np.zeros((self.m, self.k), order="F")
        H = np.zeros((self.k, self.n), order="F")
        check(lib().mfb_als_result(self._h, dptr(W), dptr(H), None, None))
        return W, H


class SnmfTooManyRestarts(UserWarning):
    """``NMF::snmf - Too many restarts due to too big 'beta' value`` (``snmf`` halves beta on it, R/bicluster.R:380-390)."""


def nmf_snmf(A, k, W0=None, beta=0.01, eta=-1.0, maxIter=20000, bi_conv=(0, 10), eps_conv=1e-4, rng=None, ctx=None):
    """What ``NMF::nmf(x = A, rank = k, method = "snmf/r", beta, eta)`` computes (the call at ``R/bicluster.R:363``):
    NMF 0.21.0's ``.nmf_snmf`` from the start ``W0`` (default: NMF's ``"random"`` seed drawn from ``rng``), including
    its restart protocol -- a zero row in H redraws ``W = matrix(runif(m * k), m, k)`` from ``rng``; the 10th such
    failure stops with NMF's warning (:class:`SnmfTooManyRestarts`).  A singular system raises ``MfbError`` with
    status ``MFB_SINGULAR`` (R: "system is computationally singular").  Returns ``dict(W, H, iters, converged,
    objective, restarts, beta, eta)``."""
    ctx = ctx or _cabi.default_context()
    own = not isinstance(A, _cabi.Matrix)
    Ad = _cabi.Matrix(ctx, host=A) if own else A
    try:
        m, n = Ad.shape
        if W0 is None:
            if rng is None:
                rng = HostRng()
            mx = Ad.stats()["max"]
            W0 = (np.asarray(rng.runif(m * k), dtype=np.float64) * mx).reshape((m, k), order="F")   # rnmf: basis ...
            rng.runif(k * n)                                                                        # ... then coef
        sess = SnmfSession(Ad, k, W0, beta, eta)
        try:
            nrestart, total, converged, objective = 0, 0, False, float("nan")
            stopped_by_restarts = False
            while total < maxIter:
                status, total, converged, objective = sess.iterate(maxIter - total, bi_conv, eps_conv)
                if status == MFB_OK:
                    break
                if status != _cabi.MFB_ZERO_ROW:
                    raise _cabi.MfbError(status, "snmf: system is computationally singular (solve() inside .fcnnls)")
                nrestart += 1
                if nrestart >= 10:
                    warnings.warn("NMF::snmf - Too many restarts due to too big 'beta' value "
                                  "[Computation stopped after the 9th restart]", SnmfTooManyRestarts)
                    stopped_by_restarts = True
                    break
                if rng is None:
                    raise _cabi.MfbError(status, "snmf: zero row in H and no rng was supplied for the restart")
                sess.restart(np.asarray(rng.runif(m * k), dtype=np.float64).reshape((m, k), order="F"))
            W, H = sess.result()
        finally:
            sess.close()
    finally:
        if own:
            Ad.free()
    return dict(W=W, H=H, iters=total, converged=converged, objective=objective, restarts=nrestart,
                too_many_restarts=stopped_by_restarts, beta=beta, eta=eta)


def snmf(A, k, verbose=True, rng=None, starts=None, ctx=None, beta=None, eta=None, **args):
    """``snmf(A, k, verbose = TRUE, ...)`` (``R/bicluster.R:341-441``) -> :class:`genericFit`.

    ``beta`` / ``eta`` default to ``mean(A)`` (:351-352); ``beta`` is halved when NMF gives up after too many
    restarts (:380-390), ``eta`` when a system is singular (:399-410); every attempt starts from the same seed
    (``rng = .Random.seed`` at :366: ``rng`` needs ``get_state`` / ``set_state`` for that, otherwise attempts continue
    the stream).  ``starts(attempt)`` -> iterator of start matrices replaces the RNG ("identical initialisation").
    An empty factorization is returned when the loop ends without a fit (:436-440)."""
    ctx = ctx or _cabi.default_context()
    A = as_f(A)
    eps = np.finfo(np.float64).eps
    pass_on = {key: args[key] for key in ("maxIter", "bi_conv", "eps_conv") if key in args}        # :345-346
    if beta is None:
        beta = float(A.mean())
    if eta is None:
        eta = float(A.mean())
    if eta < 0 or beta < eps:
        raise ValueError("eta must be >= 0 and beta must be > 0")
    m, n = A.shape
    Ad = _cabi.Matrix(ctx, host=A)
    state = rng.get_state() if hasattr(rng, "get_state") else None
    res = None
    attempt = 0
    try:
        while eta >= 0 and beta >= eps and res is None:
            attempt += 1
            if starts is not None:
                it = iter(starts(attempt))
                W0 = next(it)

                class _Replay:
                    def runif(self, cnt):
                        return np.asarray(next(it), dtype=np.float64).ravel(order="F")
                use_rng = _Replay()
            else:
                if state is not None:
                    rng.set_state(state)
                W0, use_rng = None, (rng or HostRng())
            try:
                with warnings.catch_warnings(record=True) as caught:
                    warnings.simplefilter("always")
                    out = nmf_snmf(Ad, k, W0=W0, beta=beta, eta=eta, rng=use_rng, ctx=ctx, **pass_on)
                for wmsg in caught:
                    if not issubclass(wmsg.category, SnmfTooManyRestarts):
                        warnings.warn(wmsg.message)
                if out["too_many_restarts"]:
                    beta = beta / 2
                    if verbose:
                        print(f"mfBiclust::snmf: Restarting with beta =  {beta}")
                    continue
                res = out
            except _cabi.MfbError as e:
                if e.status != _cabi.MFB_SINGULAR:
                    raise
                if eta == 0.0:
                    raise RuntimeError("Aborting snmf because dataset is singular") from e
                eta = eta / 2 if eta > eps else 0.0
                if verbose:
                    print(f"mfBiclust::snmf: Restarting with eta =  {eta}")
    finally:
        Ad.free()
    if res is None:
        fit = genericFit(genericFactorization(W=np.zeros((0, 0)), H=np.zeros((0, 0))), "snmf")
        return fit
    if verbose:
        print("method: snmf/r \nparameters:\nbeta: %g \neta: %g " % (res["beta"], res["eta"]))
    fit = genericFit(genericFactorization(W=res["W"], H=res["H"]), "snmf")
    fit.iters, fit.objective, fit.parameters, fit.attempts = res["iters"], res["objective"], dict(beta=beta, eta=eta), attempt
    return fit


# ---------------------------------------------------------------------------
# als_nmf on one very large matrix: rows of A and W sharded over the ranks
# ---------------------------------------------------------------------------
def row_range(m, rank, world):
    """Contiguous block of rows of A (and of W) owned by ``rank``."""
    return (m * rank) // world, (m * (rank + 1)) // world


class _DeviceDoubles:
    """``__cuda_array_interface__`` view of ``count`` doubles at a raw device pointer (no copy, not owned)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = dict(shape=(int(count),), typestr="<f8", data=(int(ptr), False),
                                             strides=None, version=3)


class HostAllReduce:
    """The sum-all-reduce the library asks its host for (``mfb_allreduce_fn``): in place on a device buffer of the
    library, ordered on the library's stream.  NCCL process groups reduce the buffer where it lies (the library's
    stream is made torch's current stream, so the collective is stream-ordered and the host does not block); any
    other back-end (gloo in the tests) stages through host memory.  World size 1 / no process group: nothing to do."""

    def __init__(self, ctx, group=None):
        self.rank, self.world, self.dist = _dist_info(group)
        self.group = group
        self.device = ctx.device
        self.calls = 0
        self.bytes = 0
        self.error = None
        self._views = {}
        self._stream = None
        self._ctx = ctx
        self.c_callback = _cabi.ALLREDUCE_FN(self._call)       # keep the thunk alive as long as the session

    def _view(self, ptr, count):
        import torch
        key = (ptr, count)
        if key not in self._views:
            self._views[key] = torch.as_tensor(_DeviceDoubles(ptr, count), device=torch.device("cuda", self.device))
        return self._views[key]

    def _call(self, user, ptr, count, stream):
        try:
            self.calls += 1
            self.bytes += 8 * int(count)
            if self.dist is None or self.world == 1:
                return 0
            import torch
            if self._stream is None:
                self._stream = torch.cuda.ExternalStream(int(stream or 0), device=torch.device("cuda", self.device))
            t = self._view(int(ptr), int(count))
            with torch.cuda.stream(self._stream):
                if self.dist.get_backend(self.group) == "nccl":
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
                else:
                    h = t.cpu()                                     # stream-ordered copy, then the host waits
                    self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM, group=self.group)
                    t.copy_(h)
                    self._stream.synchronize()
            return 0
        except BaseException as e:          # never unwind through the C frames
            self.error = e
            return 1


class ShardedAlsSession(AlsSession):
    """One ``als_nmf`` replicate on a matrix whose rows are split over the ranks of ``group``
    (``mfb_als_begin_sharded``): this rank holds ``A_rows`` and the matching rows of W; H is replicated."""

    def __init__(self, A_rows: _cabi.Matrix, m_global: int, k: int, W0_rows, Hold0, resid_old0: float, group=None,
                 exchange="auto"):
        """``exchange``: ``"peer"`` = one-shot all-reduce over cudaIpc peer memory fused into the consuming kernels
        (all ranks on one node), ``"callback"`` = the host's collective (NCCL / gloo through torch.distributed),
        ``"auto"`` = peer memory if every rank can map every other rank's block, else the callback."""
        self.A = A_rows
        self.k = int(k)
        self.m, self.n = A_rows.shape
        self.m_global = int(m_global)
        W0_rows = as_f(W0_rows)
        Hold0 = as_f(Hold0)
        if W0_rows.shape != (self.m, self.k) or Hold0.shape != (self.k, self.n):
            raise ValueError("W0_rows must be m_local x k and Hold0 k x n")
        self.reduce = HostAllReduce(A_rows.ctx, group)
        self._h = C.c_void_p()
        self._check(lib().mfb_als_begin_sharded(A_rows.ctx.handle, A_rows.handle, self.m_global, self.k, dptr(W0_rows),
                                                dptr(Hold0), float(resid_old0),
                                                C.cast(self.reduce.c_callback, C.c_void_p), None, C.byref(self._h)))
        self.iters = 0
        self.group = group
        self.exchange = "callback"
        if exchange in ("auto", "peer"):
            self._connect_peers(required=(exchange == "peer"))

    def _connect_peers(self, required):
        rank, world, dist = self.reduce.rank, self.reduce.world, self.reduce.dist
        mine = np.zeros(64, dtype=np.uint8)
        ok = lib().mfb_als_shard_export(self._h, mine.ctypes.data_as(C.c_void_p)) == MFB_OK
        table = np.zeros((world, 65))
        table[rank, :64] = mine
        table[rank, 64] = 0.0 if ok else 1.0
        table = _allreduce_host(table, "SUM", self.A.ctx.device, self.group)       # exact: one contributor per row
        ok = table[:, 64].sum() == 0
        if ok:
            handles = np.ascontiguousarray(table[:, :64].astype(np.uint8))
            ok = lib().mfb_als_shard_connect(self._h, rank, world, handles.ctypes.data_as(C.c_void_p)) == MFB_OK
        # the ranks must agree: if one of them could not map a peer, all of them keep the callback
        bad, = _allreduce_host([0.0 if ok else 1.0], "SUM", self.A.ctx.device, self.group)
        if bad == 0:
            self.exchange = "peer"
        else:
            if ok:
                check(lib().mfb_als_shard_disconnect(self._h))      # back to the callback, like the ranks that failed
            if required:
                raise _cabi.MfbError(_cabi.MFB_ERR_UNSUPPORTED, "peer-memory exchange is not available on every "
                                     "rank: " + lib().mfb_last_error().decode("utf-8", "replace"))

    def close(self):
        if self._h and self.exchange == "peer" and self.reduce.dist is not None and self.reduce.world > 1:
            self.reduce.dist.barrier(group=self.group)      # peers may still be reading this rank's exchange block
        super().close()

    def _check(self, status):
        if self.reduce.error is not None:
            e, self.reduce.error = self.reduce.error, None
            raise e
        return check(status)

    def iterate(self, max_new_iters, eps_conv=1e-4, fixed_iters=False):
        it, conv = C.c_int(0), C.c_int(0)
        st = self._check(lib().mfb_als_iterate(self._h, int(max_new_iters), float(eps_conv), int(bool(fixed_iters)),
                                               C.byref(it), C.byref(conv)))
        self.iters = it.value
        return st, it.value, bool(conv.value)


def _allreduce_host(values, op, ctx_device, group):
    """Tiny host-side all-reduce (global max / min of A) over the process group."""
    rank, world, dist = _dist_info(group)
    v = np.asarray(values, dtype=np.float64)
    if dist is None or world == 1:
        return v
    import torch
    dev = torch.device("cuda", ctx_device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(v.copy()).to(dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op), group=group)
    return t.cpu().numpy()


def gather_rows(W_rows, m_global, ctx_device=0, group=None):
    """All ranks' row blocks of W, in rank order, as one ``m_global x k`` matrix on every rank (one all-reduce of the
    zero-padded matrix: every entry has exactly one non-zero contributor, so the result is exact)."""
    rank, world, dist = _dist_info(group)
    if dist is None or world == 1:
        return np.asfortranarray(W_rows)
    import torch
    r0, r1 = row_range(m_global, rank, world)
    full = np.zeros((m_global, W_rows.shape[1]))
    full[r0:r1] = W_rows
    dev = torch.device("cuda", ctx_device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(full).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return np.asfortranarray(t.cpu().numpy())


def als_replicate_sharded(A_rows, m_global, k, W0, Hold0, maxIter=100, eps_conv=1e-4, fixed_iters=False, rng=None,
                          ctx=None, group=None, gather=True, exchange="auto"):
    """:func:`als_replicate` with the rows of A split over the ranks of ``group``.

    ``A_rows`` is THIS rank's block ``A[row_range(m_global, rank, world), :]`` (host array or resident
    :class:`Matrix`); ``W0`` (``m_global x k``), ``Hold0`` and ``rng`` are REPLICATED: every rank holds the same
    start and draws the same restart matrices (same seed), and keeps its own rows -- exactly what one R process
    would draw.  Status codes, restarts, traces and the convergence decision are identical on every rank.  Returns
    the full W when ``gather`` (default), else this rank's rows."""
    ctx = ctx or _cabi.default_context()
    rank, world, _ = _dist_info(group)
    r0, r1 = row_range(m_global, rank, world)
    own = not isinstance(A_rows, _cabi.Matrix)
    Ad = _cabi.Matrix(ctx, host=A_rows) if own else A_rows
    try:
        if Ad.shape[0] != r1 - r0:
            raise ValueError(f"rank {rank} of {world} owns rows [{r0}, {r1}) but A_rows has {Ad.shape[0]} rows")
        st = Ad.stats()
        amin, = _allreduce_host([st["min"]], "MIN", ctx.device, group)
        amax, = _allreduce_host([st["max"]], "MAX", ctx.device, group)
        if amin < 0:
            raise ValueError("Negative values are not allowed")                     # :76
        if eps_conv <= 0:
            raise ValueError("ALS-NMF: Invalid argument 'eps_conv' - value should be positive")   # :86
        W0 = np.asarray(W0)
        if W0.shape != (m_global, k):
            raise ValueError("W0 must be the full m_global x k start")
        sess = ShardedAlsSession(Ad, m_global, k, W0[r0:r1], Hold0, float(amax), group, exchange=exchange)
        try:
            nrestart = 0
            total = 0
            converged = False
            while total < maxIter:
                status, total, converged = sess.iterate(maxIter - total, eps_conv, fixed_iters)
                if status == MFB_OK:
                    break
                nrestart += 1
                if nrestart >= 10:
                    warnings.warn("als_nmf: Factorization failed too many times "
                                  "[Computation stopped after the 9th restart]")
                    break
                if rng is None:
                    raise _cabi.MfbError(status, "ALS iteration failed (singular Gram / zero row of H) and no "
                                                 "rng was supplied for restart()")
                sess.restart(_normalised_random_W(rng, m_global, k)[r0:r1])          # same draw on every rank
            W, H, resid, trace = sess.result()
            stats = dict(allreduce_calls=sess.reduce.calls, allreduce_bytes=sess.reduce.bytes, exchange=sess.exchange)
        finally:
            sess.close()
    finally:
        if own:
            Ad.free()
    if gather:
        W = gather_rows(W, m_global, ctx.device, group)
    return dict(obj=resid, W=W, H=H, iters=total, converged=converged, trace=trace, rows=(r0, r1), **stats)


def als_nmf_sharded(A_rows, m_global, k, reps=4, maxIter=100, eps_conv=1e-4, rng=None, inits=None, ctx=None,
                    group=None, exchange="auto", **_):
    """``als_nmf`` for a matrix too large for (or simply spread over) one device: same replicates, same winner
    (``which.min`` of the residuals, which are identical on every rank) -> :class:`genericFit` with the full W, H."""
    ctx = ctx or _cabi.default_context()
    A_rows = as_f(A_rows)
    rng = rng or HostRng(0)
    n = A_rows.shape[1]
    Ad = _cabi.Matrix(ctx, host=A_rows)
    try:
        sols = []
        for rep in range(int(reps)):
            if inits is not None:
                W0, Hold0 = inits[rep]
            else:
                W0 = _normalised_random_W(rng, m_global, k)
                Hold0 = np.asarray(rng.runif(k * n), dtype=np.float64).reshape((k, n), order="F")
            sols.append(als_replicate_sharded(Ad, m_global, k, W0, Hold0, maxIter, eps_conv, rng=rng, ctx=ctx,
                                              group=group, exchange=exchange))
        best = int(np.argmin([s["obj"] for s in sols]))
    finally:
        Ad.free()
    res = genericFit(genericFactorization(W=sols[best]["W"], H=sols[best]["H"]), "als-nmf")
    res.solutions = sols
    res.best = best
    return res


def nnls(Cmat, B, ctx=None, return_draws=False):
    """``NMF::.fcnnls(x = C, y = B)[[1]]``: exact column-wise NNLS (unit-test entry point).  ``return_draws``: also
    the number of uniforms the call would consume through ``max.col`` in R (-1 for k > 32)."""
    ctx = ctx or _cabi.default_context()
    Cmat, B = as_f(Cmat), as_f(B)
    r, k = Cmat.shape
    p = B.shape[1]
    K = np.zeros((k, p), order="F")
    draws = C.c_longlong(0)
    st = check(lib().mfb_nnls_draws(ctx.handle, dptr(Cmat), dptr(B), r, k, p, dptr(K), C.byref(draws)))
    if st == MFB_SINGULAR:
        raise np.linalg.LinAlgError("system is computationally singular")
    return (K, int(draws.value)) if return_draws else K


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


# ---------------------------------------------------------------------------
# post-processing of memberships: sizes / overlap / union.biclust / filter.biclust
# ---------------------------------------------------------------------------
def _memberships(rowxBicluster, biclusterxCol):
    R = np.asfortranarray(np.asarray(rowxBicluster).astype(np.int32))
    Cm = np.asfortranarray(np.asarray(biclusterxCol).astype(np.int32))
    if R.ndim != 2 or Cm.ndim != 2 or R.shape[1] != Cm.shape[0]:
        raise ValueError("RowxBicluster must have the same number of columns asBiclusterxCol")   # R/helperFunctions.R:157-160
    return R, Cm


def biclusterMatrix2List(rowxBicluster, biclusterxCol):
    """``biclusterMatrix2List`` (``R/helperFunctions.R:25-31``): 1-based row / column indices of every bicluster."""
    R, Cm = _memberships(rowxBicluster, biclusterxCol)
    return ([np.flatnonzero(R[:, b]) + 1 for b in range(R.shape[1])],
            [np.flatnonzero(Cm[b, :]) + 1 for b in range(Cm.shape[0])])


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def overlap_sizes(rowxBicluster, biclusterxCol, ctx=None):
    """``sizes()`` and ``overlap(..., matrix = TRUE)`` (``R/helperFunctions.R:303-308, 226-247``) of all biclusters in
    one device call (``mfb_bicluster_overlap``): returns ``(sizes[k], overlaps[k, k])``, ``overlaps[i, j]`` = elements
    shared by biclusters i and j / size of bicluster i (NaN for an empty bicluster i, as ``0 / 0`` in R)."""
    ctx = ctx or _cabi.default_context()
    R, Cm = _memberships(rowxBicluster, biclusterxCol)
    k = R.shape[1]
    if k == 0:
        return np.zeros(0), np.zeros((0, 0))
    sz = np.zeros(k)
    ov = np.zeros((k, k), order="F")
    check(lib().mfb_bicluster_overlap(ctx.handle, _i32p(R), R.shape[0], _i32p(Cm), Cm.shape[1], k, dptr(sz), dptr(ov)))
    return sz, ov


def sizes(rowxBicluster, biclusterxCol, ctx=None):
    return overlap_sizes(rowxBicluster, biclusterxCol, ctx)[0]


def overlap(rowxBicluster, biclusterxCol, matrix=False, ctx=None):
    ov = overlap_sizes(rowxBicluster, biclusterxCol, ctx)[1]
    return ov if matrix else [ov[i, :].copy() for i in range(ov.shape[0])]


def union_biclust(RowxBiclust, BiclustxCol, choose=None, ctx=None):
    """``union.biclust(RowxBiclust, BiclustxCol, choose)`` (``R/helperFunctions.R:336-345``): m x n matrix of 0 / 1."""
    ctx = ctx or _cabi.default_context()
    R, Cm = _memberships(RowxBiclust, BiclustxCol)
    k = R.shape[1]
    out = np.zeros((R.shape[0], Cm.shape[1]), order="F")
    if k == 0 or out.size == 0:
        return out
    ch = None
    if choose is not None:
        ch = np.ascontiguousarray(np.asarray(choose).astype(np.int32))
        if ch.size != k:
            raise ValueError("choose must have one element per bicluster")
    check(lib().mfb_bicluster_union(ctx.handle, _i32p(R), R.shape[0], _i32p(Cm), Cm.shape[1], k,
                                    _i32p(ch) if ch is not None else None, dptr(out)))
    return out


def filter_biclust(rowxBicluster, biclusterxCol, max=None, overlap=0.25, ctx=None):
    """``filter.biclust(rowxBicluster, biclusterxCol, max = NULL, overlap = 0.25)`` (``R/helperFunctions.R:150-206``):
    biclusters are taken largest first; one that shares more than ``overlap`` of ITS elements with a chosen one, an
    empty one and one that covers the whole matrix are never taken.  The pairwise intersections and the union come
    from the device; the greedy loop over k numbers runs here, statement for statement as in R."""
    R, Cm = _memberships(rowxBicluster, biclusterxCol)
    k = R.shape[1]
    if k == 0:
        chosen = np.zeros(0, dtype=bool)
    elif k == 1:
        chosen = np.ones(1, dtype=bool)                       # no filtering needed (:166-168)
    else:
        sz, ov = overlap_sizes(R, Cm, ctx)
        chosen = np.zeros(k, dtype=bool)
        pool = np.ones(k, dtype=bool)
        pool[(sz == 0) | (sz == R.shape[0] * Cm.shape[1])] = False          # :181
        # `while(all(sum(chosen) < max) && sum(pool) > 0)`: with max = NULL the comparison is logical(0), all() of it TRUE
        while (max is None or chosen.sum() < max) and pool.sum() > 0:
            idx = np.flatnonzero(pool)
            choose_me = idx[np.argmax(sz[idx])]                             # which.max: first maximum in index order
            chosen[choose_me] = True
            with np.errstate(invalid="ignore"):
                pool[ov[:, choose_me] > overlap] = False                    # NaN > x is NA in R: `pool[NA] <- FALSE` only
            pool[choose_me] = False                                         # touches nothing for NA subscripts of an empty one
    biclustered = union_biclust(R, Cm, chosen, ctx) if k > 0 else np.zeros((R.shape[0], Cm.shape[1]))
    return {"RowxBicluster": np.asarray(rowxBicluster)[:, chosen], "BiclusterxCol": np.asarray(biclusterxCol)[chosen, :],
            "biclustered": biclustered, "chosen": chosen}


# ---------------------------------------------------------------------------
# svd_pca / nipals_pca
# ---------------------------------------------------------------------------
def svd_pca(A, k, ctx=None, **_):
    """``svd_pca(A, k, ...)`` (``R/bicluster.R:528-539``): ``prcomp(t(A), rank. = k, retx = TRUE,
    center = FALSE)``; W = rotation (m x k'), H = t(x) (k' x n), k' = min(k, dim(A)).

    The sign of a component is the one the device eigen-solver returns (LAPACK's is equally arbitrary);
    W[:, j] and H[j, :] always flip together."""
    ctx = ctx or _cabi.default_context()
    own = not isinstance(A, _cabi.Matrix)
    Ad = _cabi.Matrix(ctx, host=A) if own else A
    try:
        m, n = Ad.shape
        kk = min(int(k), m, n)
        W = np.zeros((m, kk), order="F")
        H = np.zeros((kk, n), order="F")
        kout = C.c_int(0)
        check(lib().mfb_svd_pca(ctx.handle, Ad.handle, int(k), dptr(W), dptr(H), C.byref(kout)))
        assert kout.value == kk
    finally:
        if own:
            Ad.free()
    return genericFit(genericFactorization(W=W, H=H), "svd-pca")


def clean(obj, cleanParam=0.0, dimsRemain=True):
    """``clean(object, cleanParam, dimsRemain)`` (``R/helperFunctions.R:89-109``): host-side pre-step."""
    if not (0 <= cleanParam <= 1):
        raise ValueError('Arg "cleanParam" must be in the range of 0 to 1.')
    obj = np.asarray(obj, dtype=np.float64)
    max_na_row = np.floor((1 - cleanParam) * obj.shape[1] + 0.5)
    max_na_col = np.floor((1 - cleanParam) * obj.shape[0] + 0.5)
    bad = np.isnan(obj) | (obj == 0)
    good_rows = bad.sum(axis=1) < max_na_row
    good_cols = bad.sum(axis=0) < max_na_col
    out = obj[np.ix_(good_rows, good_cols)]
    return (out, [good_rows, good_cols]) if dimsRemain else out


def nipals(x, ncomp, center=True, scale=True, maxiter=500, tol=1e-6, gramschmidt=True, ctx=None):
    """``nipals::nipals(x, ncomp, center, scale, maxiter, tol, gramschmidt)`` for complete data ->
    ``dict(eig, scores, loadings, iter)``; the call site is ``R/bicluster.R:226``."""
    ctx = ctx or _cabi.default_context()
    own = not isinstance(x, _cabi.Matrix)
    xd = _cabi.Matrix(ctx, host=x) if own else x
    try:
        m, n = xd.shape
        ncomp = int(ncomp)
        scores = np.zeros((m, ncomp), order="F")
        loadings = np.zeros((n, ncomp), order="F")
        eig = np.zeros(ncomp)
        iters = np.zeros(ncomp, dtype=np.int32)
        st = check(lib().mfb_nipals(ctx.handle, xd.handle, ncomp, int(bool(center)), int(bool(scale)), int(maxiter),
                                    float(tol), int(bool(gramschmidt)), dptr(scores), dptr(loadings), dptr(eig),
                                    iters.ctypes.data_as(C.POINTER(C.c_int))))
        if st == _cabi.MFB_NOT_CONVERGED:
            warnings.warn("Stopping after %d iterations for a principal component" % int(maxiter))
    finally:
        if own:
            xd.free()
    return dict(eig=eig, scores=scores, loadings=loadings, iter=iters)


def nipals_pca_nocatch(A, k, ctx=None, **kwargs):
    """``nipals_pca_nocatch(A, k, ...)`` (``R/bicluster.R:218-242``): forwards scale / center / maxiter /
    gramschmidt when given (``tol`` is fixed at 1e-6 by the call) -> genericFit."""
    fwd = {key: kwargs[key] for key in ("scale", "center", "maxiter", "gramschmidt") if kwargs.get(key) is not None}
    np_res = nipals(A, k, tol=1e-6, ctx=ctx, **fwd)
    return genericFit(genericFactorization(W=np_res["scores"], H=np.asfortranarray(np_res["loadings"].T)),
                      "nipals-pca")


def nipals_pca(A, k, cleanParam=0, verbose=True, ctx=None, **kwargs):
    """``nipals_pca(A, k, cleanParam = 0, verbose = TRUE, ...)`` (``R/bicluster.R:269-289``) ->
    ``dict(m, genericFit, indexRemaining)``."""
    mClean, indexRem = clean(A, cleanParam, dimsRemain=True)
    miss = np.isnan(mClean)
    if miss.any() and (miss.all(axis=0).any() or miss.all(axis=1).any()):
        raise ValueError("At least one column or row is all NAs")          # nipals::nipals' own check
    return dict(m=mClean, genericFit=nipals_pca_nocatch(mClean, k, ctx=ctx, **kwargs), indexRemaining=indexRem)


# ---------------------------------------------------------------------------
# bcv / auto_bcv
# ---------------------------------------------------------------------------
def kLimiter(holdouts, dim, ks):
    """``kLimiter`` (``R/bcv.R:241-251``)."""
    ks = np.asarray(ks)
    kLimit = np.floor((holdouts - 1) / holdouts * dim)
    if (ks < kLimit).sum() < len(ks):
        warnings.warn("Due to the number of holdouts, not all ks can be tested. A higher number of holdouts can "
                      "be specified to test higher values of k, but the author does not recommend doing so.\n")
    return ks[ks < kLimit]


def draw_folds(rng, n, p, holdouts):
    """``sample((seq_len(n) %% holdouts) + 1L)`` for rows and columns (``R/bcv.R:170-174``)."""
    rf = ((np.arange(1, n + 1) % holdouts) + 1)[np.asarray(rng.permutation(n))]
    cf = ((np.arange(1, p + 1) % holdouts) + 1)[np.asarray(rng.permutation(p))]
    return rf.astype(np.int32), cf.astype(np.int32)


def _dist_info(group):
    """(rank, world, dist module or None) of the process group the cells are sharded over."""
    try:
        import torch.distributed as dist
    except Exception:
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, None
    return dist.get_rank(group), dist.get_world_size(group), dist


def cell_range(ncell, rank, world):
    """Contiguous block of (row-fold, col-fold) cells owned by ``rank``."""
    return (ncell * rank) // world, (ncell * (rank + 1)) // world


def _bcast_folds_many(folds_list, n, p, ctx_device, dist, group):
    """Rank 0's fold vectors of several draws in ONE broadcast."""
    import torch
    arr = np.concatenate([np.concatenate([np.asarray(rf, dtype=np.int32), np.asarray(cf, dtype=np.int32)])
                          for rf, cf in folds_list]).astype(np.int32)
    dev = torch.device("cuda", ctx_device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(arr).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    a = t.cpu().numpy().reshape(len(folds_list), n + p)
    return [(np.ascontiguousarray(a[i, :n]), np.ascontiguousarray(a[i, n:])) for i in range(len(folds_list))]


def shard_cells(ncell, nks, compute, ctx_device=0, group=None):
    """Run ``compute(cell_begin, cell_end) -> (cells x nks) array`` on this rank's contiguous block of the
    ``ncell`` bcv cells and exchange the results: one all-reduce of a zero-padded ``ncell x nks`` array.  Each
    entry has exactly one non-zero contributor, so the gathered array is exact and does not depend on the
    number of ranks.  Without ``torch.distributed`` (or world size 1) all cells run here."""
    rank, world, dist = _dist_info(group)
    c0, c1 = cell_range(ncell, rank, world)
    err = np.zeros((ncell, nks))
    if c1 > c0:
        err[c0:c1] = compute(c0, c1)
    if dist is not None and world > 1:
        import torch
        dev = torch.device("cuda", ctx_device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(err).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        err = t.cpu().numpy()
    return err


def bcvGivenKs(Y, ks, holdouts=10, folds=None, rng=None, ctx=None, group=None, cell_errors=False, streams=None):
    """``bcvGivenKs(Y, ks, holdouts)`` (``R/bcv.R:164-239``) on a resident matrix.

    The ``holdouts^2`` cells are independent given the fold vectors: under ``torch.distributed`` every rank
    computes a contiguous block of cells on its own GPU and only the per-cell error vectors are exchanged
    (:func:`shard_cells`); within a rank the cells are spread over ``streams`` concurrent CUDA streams."""
    ctx = ctx or _cabi.default_context()
    own = not isinstance(Y, _cabi.Matrix)
    Yd = _cabi.Matrix(ctx, host=Y) if own else Y
    try:
        n, p = Yd.shape
        if folds is None:
            folds = draw_folds(rng or HostRng(), n, p, holdouts)
        err = _bcv_cells_all(Yd, ks, holdouts, [folds], ctx, group, streams)[0]
    finally:
        if own:
            Yd.free()
    rcvs = err.sum(axis=0)            # rowSums over row folds of rowSums over column folds (:231-233)
    return (rcvs, err) if cell_errors else rcvs


def bcv(Y, ks=None, holdouts=10, interactive=True, rng=None, folds=None, ctx=None, group=None):
    """``bcv(Y, ks = seq_len(min(dim(Y)) - 1), holdouts = 10L, interactive = TRUE)`` (``R/bcv.R:124-159``)
    -> dict ``{k: bcv value}`` (R: named numeric vector)."""
    shape = Y.shape
    if not isinstance(Y, _cabi.Matrix):
        Y = np.asarray(Y, dtype=np.float64)
        if np.isnan(Y).any() and interactive:                                  # :128-135 (R asks with menu())
            print("Since some elements of the matrix are NA, BCV will run a much slower iterative algorithm.")
    minDim = min(shape)
    if ks is None:
        ks = np.arange(1, minDim)
    ks = kLimiter(holdouts, minDim, ks)
    ks = np.asarray(ks)
    if ks.size == 0 or (ks < 1).any() or (ks != np.floor(ks)).any():
        raise ValueError("ks must be a vector of positive whole numbers")
    if holdouts > minDim:
        holdouts = minDim
        warnings.warn(f"Using {holdouts} holdouts to ensure non-zero holdout size.")
    res = bcvGivenKs(Y, ks, holdouts, folds=folds, rng=rng, ctx=ctx, group=group)
    return {int(k): float(v) for k, v in zip(ks, res)}


def _bcv_cells_all(Yd, ks, holdouts, folds_list, ctx=None, group=None, streams=None, na=None):
    """Per-cell errors ``(len(folds_list), holdouts^2, len(ks))`` for several fold draws at once: the
    ``len(folds_list) * holdouts^2`` cells are one flat work list sharded over the ranks (:func:`shard_cells`); a
    rank hands its whole share to ONE ``mfb_bcv_cells_multi`` call, which batches the cells on the device (one
    cooperative launch of the eigen-solver per batch, the thin products of the cells side by side on auxiliary
    streams).  Every cell is computed by the same arithmetic whatever the batching: the values do not depend on
    the number of ranks or on the batch size.  ``streams`` is accepted for compatibility and ignored."""
    ctx = ctx or _cabi.default_context()
    rank, world, dist = _dist_info(group)
    ks = np.ascontiguousarray(np.asarray(ks, dtype=np.int32))
    h2 = holdouts * holdouts
    n, p = Yd.shape
    if dist is not None and world > 1:
        flat = _bcast_folds_many(folds_list, n, p, ctx.device, dist, group)
    else:
        flat = [(np.ascontiguousarray(rf, dtype=np.int32), np.ascontiguousarray(cf, dtype=np.int32))
                for rf, cf in folds_list]
    rfs = np.ascontiguousarray(np.stack([f[0] for f in flat]), dtype=np.int32)
    cfs = np.ascontiguousarray(np.stack([f[1] for f in flat]), dtype=np.int32)
    if rfs.shape != (len(flat), n) or cfs.shape != (len(flat), p):
        raise ValueError("fold vectors must have one entry per row / column of Y")

    if na is None:                                     # R/bcv.R:177: if (any(is.na(Y))) the NIPALS-based pca is used
        na = Yd.stats()["n_nan"] > 0
    entry = lib().mfb_bcv_cells_na if na else lib().mfb_bcv_cells_multi

    def compute(c0, c1):
        out = np.zeros((c1 - c0, len(ks)))
        ctx.set_concurrency(1)
        check(entry(ctx.handle, Yd.handle, len(flat), rfs.ctypes.data_as(C.POINTER(C.c_int32)),
                                        cfs.ctypes.data_as(C.POINTER(C.c_int32)), int(holdouts),
                                        ks.ctypes.data_as(C.POINTER(C.c_int32)), len(ks), int(c0), int(c1), dptr(out)))
        return out
    err = shard_cells(len(flat) * h2, len(ks), compute, ctx.device, group)
    return err.reshape(len(flat), h2, len(ks))


def bcv_batch(Yd, ks, holdouts, folds_list, ctx=None, group=None, streams=None):
    """BCV values ``(len(folds_list), len(ks))`` of several fold draws (see :func:`_bcv_cells_all`)."""
    return _bcv_cells_all(Yd, ks, holdouts, folds_list, ctx, group, streams).sum(axis=1)


def _iters_until_stable(distr, distrOld, tol, winner, keys, limit):
    """Number of further outer iterations after which ``auto_bcv``'s stopping rule fires if ``winner`` wins every
    one of them (at least 1, at most ``limit``): used to size the speculative batches."""
    if winner is None:
        return limit
    d, dold = dict(distr), dict(distrOld)
    for t in range(1, limit + 1):
        d[winner] += 1
        sd, so = sum(d.values()), sum(dold.values())
        if all((d[kk] / sd - dold[kk] / so) ** 2 < tol for kk in keys):
            return t
        dold = dict(d)
    return limit


def _earliest_stop(distr, tol, limit):
    """Lower bound on the number of further outer iterations before ``auto_bcv``'s stopping rule can fire on ANY
    sequence of winners.  The rule needs, for the winner w of the deciding iteration,
    ``((d_w + 1) / (S + 1) - d_w / S)^2 < tol`` (S = sum of the counts before it); the left side decreases with
    d_w, and after t - 1 further iterations no count exceeds ``max(distr) + t - 1``: the path on which the current
    leader wins every time fires first.  Iterations up to this bound can be computed as ONE batch without waste."""
    S, D = sum(distr.values()), max(distr.values())
    for t in range(1, limit + 1):
        if ((D + t) / (S + t) - (D + t - 1) / (S + t - 1)) ** 2 < tol:
            return t
    return limit


def auto_bcv(Y, ks=None, holdouts=3, maxIter=100, tol=1e-4, bestOnly=False, verbose=False, interactive=True,
             rng=None, ctx=None, group=None, speculate=None):
    """``auto_bcv(Y, ks, holdouts = 3L, maxIter = 100L, tol = 1e-4, bestOnly = FALSE, verbose = FALSE,
    interactive = TRUE)`` (``R/bcv.R:31-98``).  Y is uploaded once and stays resident for every ``bcv`` call.

    The outer loop is sequential only through its stopping rule; the fold draws do not depend on the data.  With
    ``speculate = B`` (default: the number of ranks, i.e. 1 on a single GPU) the folds of the next B outer
    iterations are drawn in order, their ``B * holdouts^2`` cells are computed as one sharded batch, and the
    stopping rule is then applied iteration by iteration exactly as in R; iterations computed past the one that
    converges are discarded.  Independently of B, the iterations the rule provably cannot do without
    (:func:`_earliest_stop`: 22 on an 11-candidate grid, 25 on a 20-candidate grid from a fresh start) are always
    computed as one batch, so their cells fill all streams and all ranks.  ``best``, ``counts`` and the iteration count are therefore identical for every B and
    every number of GPUs; only the number of uniforms consumed from ``rng`` after convergence differs."""
    ctx = ctx or _cabi.default_context()
    rng = rng or HostRng()
    rank, world, _ = _dist_info(group)
    B = max(1, int(speculate if speculate is not None else world))
    own = not isinstance(Y, _cabi.Matrix)
    if own:
        Y = np.asarray(Y, dtype=np.float64)
        if np.isnan(Y).any() and interactive:                                  # :39-46 (R asks with menu())
            print("Since some elements of the matrix are NA, BCV will run a much slower iterative algorithm.")
    Yd = _cabi.Matrix(ctx, host=Y) if own else Y
    try:
        n, p = Yd.shape
        minDim = min(n, p)
        if ks is None:
            ks = np.arange(1, minDim)
        ks = kLimiter(holdouts, minDim, ks)                                   # :48
        ks = np.asarray(ks)
        if ks.size == 0 or (ks < 1).any() or (ks != np.floor(ks)).any():
            raise ValueError("ks must be a vector of positive whole numbers")
        if holdouts > minDim:                                                  # bcv(): :146-150
            holdouts = minDim
            warnings.warn(f"Using {holdouts} holdouts to ensure non-zero holdout size.")
        keys = [int(k) for k in ks]
        distr = {k: 1 for k in keys}
        distrOld = dict(distr)
        i = 0
        converged = False
        last_best = None
        while not converged and i < maxIter:                                   # :56
            # batch size: every iteration the stopping rule provably cannot do without (no waste, any number of
            # ranks); beyond that, speculation: never more than B, and no more than the rule can use if the current
            # winner keeps winning (it cannot fire earlier on that path; any other path is continued next batch)
            import os as _os
            pre = 1 if _os.environ.get("MFBICLUST_BCV_NO_PREBATCH") else _earliest_stop(distr, tol, maxIter - i)
            nb = max(pre, min(B, _iters_until_stable(distr, distrOld, tol, last_best, keys, B)))
            nb = min(nb, maxIter - i)
            folds = [draw_folds(rng, n, p, holdouts) for _ in range(nb)]       # in order: the stream R would consume
            vals = bcv_batch(Yd, keys, holdouts, folds, ctx=ctx, group=group)  # :58, nb calls of bcv()
            for j in range(nb):
                res = dict(zip(keys, vals[j]))
                best = min(keys, key=lambda kk: (res[kk], keys.index(kk)))     # names(which.min(res))
                last_best = best
                distr[best] += 1
                sd, so = sum(distr.values()), sum(distrOld.values())
                resid = [(distr[kk] / sd - distrOld[kk] / so) ** 2 for kk in keys]   # :69-71
                if verbose:
                    cur = max(keys, key=lambda kk: (distr[kk], -keys.index(kk)))
                    print(f"Iteration {i + 1} ")
                    print(f"Current best: {cur} ")
                    print("BCV result distribution:")
                    print("\t".join(str(kk) for kk in keys) + "\n" + "\t".join(str(distr[kk] - 1) for kk in keys))
                converged = all(r < tol for r in resid)                        # :84
                distrOld = dict(distr)
                i += 1
                if converged:
                    break
        if i == maxIter:
            warnings.warn(f"BCV results did not converge after{maxIter}iterations")
    finally:
        if own:
            Yd.free()
    counts = {kk: v - 1 for kk, v in distr.items()}                            # :93
    med = max(keys, key=lambda kk: (counts[kk], -keys.index(kk)))              # which.max: first maximum
    if bestOnly:
        return med
    return dict(best=med, counts=counts, iterations=i)


# ---------------------------------------------------------------------------
# BiclusterStrategy / addStrat (the callers either side of the path; R/BiclusterStrategy.R:61-163,
# 475-502 and R/addStrat.R:49-97) -- glue only, no arithmetic of their own
# ---------------------------------------------------------------------------
def pseudovalues(m):
    """``pseudovalues`` (``R/helperFunctions.R:270-278``)."""
    m = np.asarray(m, dtype=np.float64)
    if (m < 0).any():
        return m + abs(m.min())
    return m


def validateKM(k, m, method):
    """``validateKM`` (``R/helperFunctions.R:400-412``) including its ``"method" == "svd-pca"`` literal."""
    if not float(k) > 0:
        raise ValueError('Argument "k" must be positive.')
    k = int(np.floor(float(k) + 0.5))
    if method in ("als-nmf", "nipals-pca", "snmf") and k > min(m.shape):
        warnings.warn("Initializing k to the size of the smaller matrix dimension.")
        return min(m.shape)
    return k


@dataclass
class BiclustResult:
    """The slots of ``biclust::BiclustResult`` the strategy fills (``R/BiclusterStrategy.R:150``)."""
    RowxNumber: np.ndarray
    NumberxCol: np.ndarray
    Number: int


@dataclass
class BiclusterStrategyResult:
    factors: genericFit
    featureThresh: np.ndarray
    sampleThresh: np.ndarray
    threshAlgo: str
    k: int
    biclust: BiclustResult
    name: str


def thresholdHelper(W, H, k, featureThresh="otsu", sampleThresh="otsu", threshAlgo="otsu", ctx=None):
    """``thresholdHelper`` (``R/BiclusterStrategy.R:475-502``) -> (threshAlgo, featureThresh, sampleThresh,
    RowxNumber, NumberxCol)."""
    if k <= 0:
        warnings.warn("Biclustering did not find valid results. No samples or features are biclustered.")
        return threshAlgo, featureThresh, sampleThresh, np.zeros((W.shape[0], 0), bool), np.zeros((0, H.shape[1]), bool)
    if not isinstance(featureThresh, str) and not isinstance(sampleThresh, str):
        featureThresh, sampleThresh = np.asarray(featureThresh, float), np.asarray(sampleThresh, float)
        if featureThresh.size != k or sampleThresh.size != k:
            raise ValueError('Length of "featureThresh" and "sampleThresh" must equal "k"')
        threshAlgo = "user"
    elif threshAlgo == "otsu":
        featureThresh = otsuHelper(W, ctx=ctx)                        # :485
        sampleThresh = otsuHelper(np.asfortranarray(H.T), ctx=ctx)    # :486
    else:
        raise ValueError('Currently "otsu" is the only implemented thresholding algorithm')
    return (threshAlgo, featureThresh, sampleThresh, threshold(W, featureThresh, 2, ctx=ctx),
            threshold(H, sampleThresh, 1, ctx=ctx))


_METHODS = {"als-nmf": "als_nmf", "svd-pca": "svd_pca", "nipals-pca": "nipals_pca", "snmf": "snmf"}


def BiclusterStrategy(obj, k, method="als-nmf", threshAlgo="otsu", featureThresh=None, sampleThresh=None, ctx=None,
                      **kwargs):
    """``BiclusterStrategy(obj, k, method, ...)`` for a matrix or a genericFit
    (``R/BiclusterStrategy.R:61-100, 134-163``)."""
    featureThresh = threshAlgo if featureThresh is None else featureThresh
    sampleThresh = threshAlgo if sampleThresh is None else sampleThresh
    if not isinstance(obj, genericFit):
        if method not in _METHODS:
            raise ValueError(f"method '{method}' is outside the GPU hot path (als-nmf | svd-pca | nipals-pca | snmf)")
        obj = np.asarray(obj, dtype=np.float64)
        if method in ("als-nmf", "snmf"):
            obj = pseudovalues(obj)                                    # :79
        try:
            if method == "nipals-pca":
                fit = nipals_pca(obj, k, ctx=ctx, **kwargs)["genericFit"]
            else:
                fit = globals()[_METHODS[method]](obj, k, ctx=ctx, **kwargs)   # do.call(get(method), ...) :82-87
        except (_cabi.MfbError, np.linalg.LinAlgError, ValueError) as e:       # :88-94
            warnings.warn(str(e))
            warnings.warn(f"{_METHODS[method]} failed, switching to PCA.")
            fit = svd_pca(obj, k, ctx=ctx)
            method = "svd-pca"
    else:
        fit = obj
    W, H = fit.fit.W, fit.fit.H
    k = min(int(k), W.shape[1])                                        # :138
    algo, fth, sth, rn, nc = thresholdHelper(W, H, W.shape[1], featureThresh, sampleThresh, threshAlgo, ctx=ctx)
    pretty = {"als-nmf": "ALS-NMF", "svd-pca": "SVD-PCA", "nipals-pca": "NIPALS-PCA", "snmf": "SNMF"}.get(method, method)
    name = " | ".join([pretty, algo.capitalize() if algo != "user" else "User", str(k)])
    return BiclusterStrategyResult(factors=fit, featureThresh=np.asarray(fth), sampleThresh=np.asarray(sth),
                                   threshAlgo=algo, k=k, biclust=BiclustResult(rn, nc, W.shape[1]), name=name)


class BiclusterExperiment:
    """Minimal carrier of the ``abund`` assay matrix and the strategies list
    (``R/BiclusterExperiment.R:36-38, 118-145``)."""

    def __init__(self, m):
        self.abund = np.asfortranarray(np.asarray(m, dtype=np.float64))
        self.strategies = {}

    def as_matrix(self):
        return self.abund


def addStrat(bce, k, method="als-nmf", silent=False, ctx=None, **kwargs):
    """``addStrat(bce, k, method, silent = FALSE, ...)`` (``R/addStrat.R:49-97``)."""
    m = bce.as_matrix()
    k = validateKM(k, m, method)
    if method == "nipals-pca" or np.isnan(m).any():
        if method != "nipals-pca":
            warnings.warn("Since some data is NA, the NIPALS-PCA algorithm must be used.")
        res = nipals_pca(m, k, cleanParam=0, center=False, ctx=ctx)                  # :72
        bcs = BiclusterStrategy(res["genericFit"], k, method="nipals-pca", ctx=ctx)
        rows, cols = res["indexRemaining"]
        if not (rows.all() and cols.all()):
            warnings.warn("Some samples or features with too much missing data were removed.")
            bce.abund = np.asfortranarray(m[np.ix_(rows, cols)])
    else:
        bcs = BiclusterStrategy(m, k, method=method, ctx=ctx, verbose=not silent, **kwargs)
    bce.strategies[bcs.name] = bcs
    if not silent:
        print(f"Added BiclusterStrategy named {bcs.name}")
    return bce
