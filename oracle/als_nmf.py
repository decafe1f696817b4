"""Restatement of ``als_nmf`` (``R/bicluster.R:70-216``).

Test infrastructure (see ``oracle/__init__.py``).

Two entry points:

* :func:`als_nmf` -- the whole R function under an R RNG stream
  (``reps`` random restarts, ``runif`` inits, ``max.col`` draws inside
  ``.fcnnls``, ``which.min`` over replicates).  Pinned end to end by the
  rendered vignette (``vignettes/my-vignette.html:224-399``).
* :func:`als_replicate` -- one replicate from a GIVEN ``W0`` / ``Hold0``
  ("identical inputs and initialisation" in BASELINE.json's north_star); this
  is what the CUDA path is compared against.  ``literal=True`` mirrors the
  reference's costs (materialised ``t(A)``, full ``A - W %*% H`` temporary) and
  is what ``bench.py`` times as the CPU baseline.
"""
from __future__ import annotations

import numpy as np

from .fcnnls import SingularError, fcnnls
from .r_rng import RRng


def _normalise_cols(W: np.ndarray) -> np.ndarray:
    # R/bicluster.R:98-101  frob <- apply(W, 2, function(x) sqrt(sum(x^2))); sweep(W, 2, frob, "/")
    frob = np.sqrt((W ** 2).sum(axis=0))
    return W / frob[None, :]


def als_replicate(A, W0, Hold0, max_iter=100, eps_conv=1e-4, fixed_iters=False,
                  rng: RRng | None = None, redraw=None, trace=None, literal=False):
    """One replicate of the ``while (i < maxIter)`` loop (``R/bicluster.R:118-203``).

    ``W0`` must already be column-normalised (``R/bicluster.R:98-101``).
    ``fixed_iters=True`` disables the convergence ``break`` (:179) so exactly
    ``max_iter`` iterations run (throughput / fixed-iteration parity mode).
    ``redraw()`` supplies a fresh normalised W on ``restart()`` (:107-116);
    without it a restart raises.  Returns ``dict(obj, W, H, iters, status)``.
    """
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    W = np.array(W0, dtype=np.float64)
    Wold = W.copy()
    Hold = np.array(Hold0, dtype=np.float64)
    resid_old = A.max()            # :105  residNormOld <- maxA
    resid = A.max()                # :91   residNorm <- max(A)
    H = None
    nrestart = 0
    At = None
    i = 0
    status = "maxiter"

    def _restart():
        nonlocal W, nrestart
        nrestart += 1
        if nrestart >= 10:                       # :126-131
            return False
        if redraw is None:
            raise SingularError("restart requested but no redraw() supplied")
        W = redraw()
        return True

    while i < max_iter:
        i += 1
        try:
            H_new, _ = fcnnls(W, A, rng)          # :123
        except SingularError:
            if not _restart():
                status = "failed"
                break
            continue
        H = H_new
        if (H.sum(axis=1) == 0).any():            # :137
            if not _restart():
                status = "failed"
                break
            continue
        try:
            if At is None or literal:
                At = np.ascontiguousarray(A.T)    # :150 t(A); R re-materialises it every iteration (literal=True)
            Wt, _ = fcnnls(H.T, At, rng)          # :150
        except SingularError:
            if not _restart():
                status = "failed"
                break
            continue
        W = Wt.T
        resid = float(np.sqrt(np.mean((A - W @ H) ** 2)))     # :173
        dW = float(np.sqrt(np.mean((W - Wold) ** 2)))         # :174-175
        dH = float(np.sqrt(np.mean((H - Hold) ** 2)))         # :176-177
        if trace is not None:
            trace.append((i, resid, dW, dH))
        if not fixed_iters and (abs(resid_old - resid) <= eps_conv or
                                (dW <= eps_conv and dH <= eps_conv)):   # :179
            status = "converged"
            break
        resid_old = resid
        Hold = H
        Wold = W
    return dict(obj=resid, W=W, H=H, iters=i, status=status)


def als_nmf(A, k, reps=4, max_iter=100, eps_conv=1e-4, rng: RRng | None = None, traces=None):
    """``als_nmf(A, k, reps, maxIter, eps_conv)`` under the R RNG ``rng``."""
    A = np.asarray(A, dtype=np.float64)
    if (A < 0).any():
        raise ValueError("Negative values are not allowed")           # :76
    if eps_conv <= 0:
        raise ValueError("ALS-NMF: Invalid argument 'eps_conv' - value should be positive")  # :86
    if rng is None:
        rng = RRng(0)
    m, n = A.shape
    sols = []
    for _ in range(reps):
        W0 = _normalise_cols(rng.runif(m * k).reshape((m, k), order="F"))     # :95-101
        Hold0 = rng.runif(k * n).reshape((k, n), order="F")                    # :104
        tr = [] if traces is not None else None
        sol = als_replicate(
            A, W0, Hold0, max_iter, eps_conv, rng=rng,
            redraw=lambda: _normalise_cols(rng.runif(m * k).reshape((m, k), order="F")),
            trace=tr)
        if traces is not None:
            traces.append(tr)
        sols.append(sol)
    best = int(np.argmin([s["obj"] for s in sols]))                             # :207
    return dict(W=sols[best]["W"], H=sols[best]["H"], best=best, solutions=sols)


