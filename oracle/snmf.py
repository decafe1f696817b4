"""Restatement of ``snmf`` (``R/bicluster.R:341-441``) and of the algorithm it delegates to,
``NMF::nmf(method = "snmf/r")`` (NMF 0.21.0, ``NMF/R/algorithms-snmf.R``: ``.nmf_snmf``, version "R").

Test infrastructure (see ``oracle/__init__.py``).

NMF 0.21.0 is pinned by ``packrat/packrat.lock:137-139`` but NOT vendored in the reference.  The algorithm is Kim &
Park's sparse NMF by alternating non-negativity-constrained least squares (Bioinformatics 23, 2007), right-sparse
variant::

    min_{W,H >= 0}  1/2 ( ||A - W H||_F^2  +  eta ||W||_F^2  +  beta sum_j (sum_c H_cj)^2 )

as shipped in NMF: columns of the start W scaled to unit length, then per iteration

    H  <- .fcnnls(rbind(W, sqrt(beta) 1'), rbind(A, 0))          i.e.  (W'W + beta 11') H = W'A,  H >= 0
    zero row in H  ->  redraw W (at most 9 times), `next`
    W' <- .fcnnls(rbind(t(H), eta I_k), rbind(t(A), 0))          i.e.  (HH' + eta^2 I) W' = H A', W >= 0

and, on the first iteration after a (re)start and on every fifth, the convergence test: the arg-max memberships of the
rows of W and columns of H (``max.col(ties.method = "first")``) unchanged for ``bi_conv[2] = 10`` consecutive tests
AND the mean absolute KKT residual at most ``eps_conv = 1e-4`` of its first value.

Pinned by ``bce.snmf`` of ``tests/testdata/ref_bces.rda`` (``tests/golden/ref_bces.json``): the stored NMFfit holds the
``.Random.seed`` the run started from, ``beta`` / ``eta``, the factors, the iteration count (50) and the final
objective value; :func:`nmf_snmf` reproduces all of them from the seed (``tests/test_oracle_golden.py``).  The seeding
method is NMF's ``"random"`` (``rnmf``: W then H from ``runif(max = max(A))``; only W is used).
"""
from __future__ import annotations

import numpy as np

from .fcnnls import SingularError, fcnnls
from .r_rng import RRng

_EPS = np.finfo(np.float64).eps


class TooManyRestarts(Warning):
    """``NMF::snmf - Too many restarts due to too big 'beta' value`` (a warning in R)."""


def _unit_columns(W):
    # W = apply(W, 2, function(x) x / sqrt(sum(x^2)))
    return W / np.sqrt((W ** 2).sum(axis=0))[None, :]


def max_col_first(M):
    """``max.col(M, ties.method = "first")``, 1-based like R."""
    return np.argmax(M, axis=1) + 1


def snmf_objective(A, W, H, eta, beta):
    """``.snmf.objective``: 1/2 (sum((A - WH)^2) + eta sum(W^2) + beta sum(colSums(H)^2))."""
    return 0.5 * (((A - W @ H) ** 2).sum() + eta * (W ** 2).sum() + beta * (H.sum(axis=0) ** 2).sum())


def kkt_residual(A, W, H, eta, beta):
    """The two ``pmin`` blocks of the convergence test: (sum |res|, #{|res| > 0})."""
    k = W.shape[1]
    res_h = np.minimum(H, (W.T @ W) @ H - W.T @ A + np.full((k, k), beta) @ H)
    res_w = np.minimum(W, W @ (H @ H.T) - A @ H.T + eta * eta * W)
    resvec = np.concatenate([res_h.ravel(order="F"), res_w.ravel(order="F")])
    return np.abs(resvec).sum(), int((np.abs(resvec) > 0).sum())


def nmf_snmf(A, W0, eta=-1.0, beta=0.01, max_iter=20000, bi_conv=(0, 10), eps_conv=1e-4, redraw=None, log=None):
    """``.nmf_snmf(A, x, version = "R")`` from a GIVEN start ``W0`` (m x k, not yet normalised).

    ``redraw()`` supplies the ``matrix(runif(m * k), m, k)`` of a restart.  Returns ``dict(W, H, iters, restarts,
    objective, warning)``; raises :class:`SingularError` where ``solve()`` inside ``.fcnnls`` would.
    """
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    if eta < 0:
        eta = A.max()
    eta2 = eta * eta
    wminchange, iconv = bi_conv
    W = _unit_columns(np.array(W0, dtype=np.float64))
    k = W.shape[1]
    idxWold = np.zeros(m, dtype=np.int64)
    idxHold = np.zeros(n, dtype=np.int64)
    inc = 0
    erravg1 = None
    I_k = np.diag(np.full(k, eta))
    betavec = np.full((1, k), np.sqrt(beta))
    nrestart = 0
    H = None
    warning = None
    i = 0
    while i < max_iter:
        i += 1
        H, _ = fcnnls(np.vstack([W, betavec]), np.vstack([A, np.zeros((1, n))]))
        if np.any(H.sum(axis=1) == 0):
            nrestart += 1
            if nrestart >= 10:
                warning = ("NMF::snmf - Too many restarts due to too big 'beta' value "
                           "[Computation stopped after the 9th restart]")
                break
            idxWold[:] = 0
            idxHold[:] = 0
            inc = 0
            erravg1 = None
            if redraw is None:
                raise RuntimeError("restart requested but no redraw() supplied")
            W = _unit_columns(np.asarray(redraw(), dtype=np.float64))
            continue
        Wt, _ = fcnnls(np.vstack([H.T, I_k]), np.vstack([A.T, np.zeros((k, m))]))
        W = Wt.T
        if i % 5 == 0 or erravg1 is None:
            idxW = max_col_first(W)
            idxH = max_col_first(H.T)
            changedW = int((idxW != idxWold).sum())
            changedH = int((idxH != idxHold).sum())
            inc = inc + 1 if (changedW <= wminchange and changedH == 0) else 0
            conv, convnum = kkt_residual(A, W, H, eta, beta)
            erravg = conv / convnum if convnum else np.nan
            if erravg1 is None:
                erravg1 = erravg
            if log is not None:
                log.append(dict(i=i, inc=inc, changedW=changedW, changedH=changedH, conv=conv, convnum=convnum))
            if inc >= iconv and erravg <= eps_conv * erravg1:
                break
            idxWold, idxHold = idxW, idxH
    return dict(W=W, H=H, iters=i, restarts=nrestart, warning=warning, eta=eta, beta=beta,
                objective=snmf_objective(A, W, H, eta, beta))


def snmf(A, k, rng: RRng | None = None, beta=None, eta=None, starts=None, **args):
    """``snmf(A, k, ...)`` (``R/bicluster.R:341-441``): ``beta`` / ``eta`` default to ``mean(A)``; ``beta`` is halved
    on the too-many-restarts warning (:380-390), ``eta`` on "computationally singular" (:399-410); every attempt
    re-seeds from the SAME ``.Random.seed`` (``rng = .Random.seed`` at :366 is evaluated once per ``do.call`` from the
    unchanged session state, because ``NMF::nmf`` restores the RNG on exit).

    ``starts(attempt)`` -> iterator of start matrices (first = seed W, later = restart draws) replaces the RNG.
    Returns ``dict(W, H, ..., beta, eta, attempts)`` or ``None`` for the empty ``genericFit`` (:436-440).
    """
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    if beta is None:
        beta = A.mean()
    if eta is None:
        eta = A.mean()
    if eta < 0 or beta < _EPS:
        raise ValueError("eta must be >= 0 and beta must be > 0")
    attempts = 0
    state = None if rng is None else rng.get_state()
    while eta >= 0 and beta >= _EPS:
        attempts += 1
        if starts is not None:
            it = iter(starts(attempts))
            W0 = next(it)
            redraw = lambda: next(it)     # noqa: E731
        else:
            rng.set_state(state)
            mx = A.max()
            W0 = (rng.runif(m * k) * mx).reshape((m, k), order="F")     # rnmf: basis first ...
            rng.runif(k * n)                                            # ... then coef (unused by snmf)
            redraw = lambda: rng.runif(m * k).reshape((m, k), order="F")   # noqa: E731
        try:
            res = nmf_snmf(A, W0, eta=eta, beta=beta, redraw=redraw, **args)
        except SingularError:
            eta = eta / 2 if eta > _EPS else 0.0
            if eta == 0.0 and attempts > 200:
                break
            continue
        if res["warning"] is not None:
            beta = beta / 2
            continue
        res["attempts"] = attempts
        return res
    return None
