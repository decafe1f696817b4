"""Restatement of ``auto_bcv`` / ``bcv`` / ``bcvGivenKs`` / ``kLimiter``
(``R/bcv.R:31-98, 124-159, 164-239, 241-251``) with ``MASS::ginv``
(MASS 7.3-51.1, ``packrat/packrat.lock:118-120``, NOT vendored; call site
``R/bcv.R:222``) and ``stats::prcomp`` (call site ``R/bcv.R:184``).

Test infrastructure (see ``oracle/__init__.py``).  Spec: SURVEY.md App. A.4.

Pinned (weakly): the vignette's ``auto_bcv(yeast_benchmark[[1]])`` -> best = 11
with all 22 counts at k = 11 (``vignettes/my-vignette.html:408-413``).  The
per-k BCV values, the ``ginv`` tolerance rule and ``prcomp`` centring have no
numeric artefact: parity unpinned; they follow the published sources.  Folds
are passed in (the R wrapper draws them with ``sample``), or drawn with
``RRng.sample`` (unpinned).
"""
from __future__ import annotations

import numpy as np

from .pca import prcomp

_SQRT_EPS = np.sqrt(np.finfo(np.float64).eps)


def ginv(X: np.ndarray, tol: float = _SQRT_EPS) -> np.ndarray:
    """``MASS::ginv``: SVD pseudo-inverse keeping ``d > max(tol * d[1], 0)``."""
    u, d, vt = np.linalg.svd(X, full_matrices=False)
    pos = d > max(tol * d[0], 0.0)
    if pos.all():
        return (vt.T * (1.0 / d)[None, :]) @ u.T
    if not pos.any():
        return np.zeros((X.shape[1], X.shape[0]))
    return (vt[pos].T * (1.0 / d[pos])[None, :]) @ u[:, pos].T


def k_limiter(holdouts: int, dim: int, ks):
    """``kLimiter`` (``R/bcv.R:241-251``): keep ``ks < floor((h-1)/h * dim)``."""
    k_limit = np.floor((holdouts - 1) / holdouts * dim)
    ks = np.asarray(ks)
    return ks[ks < k_limit]


def make_folds(n: int, holdouts: int, perm: np.ndarray) -> np.ndarray:
    """``sample((seq_len(n) %% holdouts) + 1L)`` given the permutation (1-based fold ids)."""
    base = (np.arange(1, n + 1) % holdouts) + 1
    return base[perm]


def bcv_given_ks_literal(Y, ks, row_fold, col_fold, holdouts):
    """``bcvGivenKs`` exactly as written: one ``ginv`` (full SVD) per k per cell
    (``R/bcv.R:195-233``).  The dead whole-matrix PCA (:189-191) is skipped."""
    Y = np.asarray(Y, dtype=np.float64)
    ks = list(ks)
    out = np.zeros(len(ks))
    kmax = max(ks)
    for x in range(1, holdouts + 1):
        r = row_fold == x
        for y in range(1, holdouts + 1):
            s = col_fold == y
            A = Y[np.ix_(r, s)]
            D = Y[np.ix_(~r, ~s)]
            tcv, rot, _ = prcomp(D, kmax, center=True)      # scores, loadings = t(rotation)
            pcv = rot.T
            for i, k in enumerate(ks):
                estD = tcv[:, :k] @ pcv[:k, :]
                estA = Y[np.ix_(r, ~s)] @ ginv(estD) @ Y[np.ix_(~r, s)]
                out[i] += float(((A - estA) ** 2).sum())
    return out


def bcv_given_ks_na_literal(Y, ks, row_fold, col_fold, holdouts):
    """``bcvGivenKs`` when ``any(is.na(Y))`` (``R/bcv.R:177-181``): ``pca`` is then
    ``nipals_pca(., k, center = TRUE, scale = FALSE)$genericFit`` -- scores ``W`` (unit columns: nipals normalises
    them and ``nipals_pca`` keeps no singular values) and loadings ``H`` -- and the rest is unchanged: per k a
    ``MASS::ginv`` of ``tcv[,1:k] %*% pcv[1:k,]``, ``%*%`` products that propagate NA along whole rows / columns of
    ``estA`` and ``sum(resid^2, na.rm = TRUE)``.  Parity unpinned (no artefact of the reference exercises this branch)."""
    from .pca import nipals
    from .helpers import clean
    Y = np.asarray(Y, dtype=np.float64)
    ks = list(ks)
    kmax = max(ks)
    out = np.zeros(len(ks))
    for x in range(1, holdouts + 1):
        r = row_fold == x
        for y in range(1, holdouts + 1):
            s = col_fold == y
            A = Y[np.ix_(r, s)]
            D = Y[np.ix_(~r, ~s)]
            Dc, good_rows, good_cols = clean(D, 0.0)
            if not (good_rows.all() and good_cols.all()):
                raise ValueError("clean() would drop rows / columns of a hold-in block: non-conformable in the reference")
            res = nipals(D, kmax, center=True, scale=False, tol=1e-6)
            tcv, pcv = res["scores"], res["loadings"].T
            Y1, Y2 = Y[np.ix_(r, ~s)], Y[np.ix_(~r, s)]
            bad_rows, bad_cols = np.isnan(Y1).any(axis=1), np.isnan(Y2).any(axis=0)
            Y1z, Y2z = np.where(np.isnan(Y1), 0.0, Y1), np.where(np.isnan(Y2), 0.0, Y2)
            for i, k in enumerate(ks):
                estD = tcv[:, :k] @ pcv[:k, :]
                estA = Y1z @ ginv(estD) @ Y2z
                estA[bad_rows, :] = np.nan                 # NA %*% anything is NA for the whole row ...
                estA[:, bad_cols] = np.nan                 # ... / column
                out[i] += float(np.nansum((A - estA) ** 2))
    return out


def bcv_cell_errors(Y, ks, r, s):
    """One (row-fold, col-fold) cell, one SVD: algebraically identical to the
    literal per-k ``ginv`` path (verified to ~1e-16 relative
    in ``tests/test_oracle_bcv.py``)."""
    A = Y[np.ix_(r, s)]
    D = Y[np.ix_(~r, ~s)]
    Dc = D - D.mean(axis=0, keepdims=True)
    u, d, vt = np.linalg.svd(Dc, full_matrices=False)
    kmax = max(ks)
    kk = min(kmax, d.size)
    B = (Y[np.ix_(r, ~s)] @ vt[:kk].T)            # |r| x kk   (times 1/sigma below)
    C = u[:, :kk].T @ Y[np.ix_(~r, s)]            # kk x |s|
    errs = np.zeros(len(ks))
    for i, k in enumerate(ks):
        # ginv of the rank-k approximation keeps sigma_j > sqrt(eps) * sigma_1, j <= k
        keep = np.flatnonzero(d[:min(k, kk)] > max(_SQRT_EPS * d[0], 0.0))
        est = (B[:, keep] / d[keep][None, :]) @ C[keep, :]
        errs[i] = float(((A - est) ** 2).sum())
    return errs


def bcv_given_ks(Y, ks, row_fold, col_fold, holdouts):
    """``bcvGivenKs`` via one SVD per cell (fast restatement used as the checker)."""
    Y = np.asarray(Y, dtype=np.float64)
    out = np.zeros(len(ks))
    for x in range(1, holdouts + 1):
        r = row_fold == x
        for y in range(1, holdouts + 1):
            s = col_fold == y
            out += bcv_cell_errors(Y, list(ks), r, s)
    return out


def auto_bcv(Y, ks, holdouts, draw_folds, max_iter=100, tol=1e-4, bcv_fn=bcv_given_ks):
    """``auto_bcv`` (``R/bcv.R:31-98``).  ``draw_folds()`` -> ``(row_fold, col_fold)``.

    Returns ``(best, counts, iterations)``."""
    Y = np.asarray(Y, dtype=np.float64)
    ks = k_limiter(holdouts, min(Y.shape), ks)               # :48
    distr = np.ones(len(ks), dtype=np.int64)
    distr_old = distr.copy()
    i = 0
    converged = False
    while not converged and i < max_iter:                    # :56
        rf, cf = draw_folds()
        res = bcv_fn(Y, ks, rf, cf, holdouts)                # :58
        distr[int(np.argmin(res))] += 1                      # :60-61
        resid = (distr / distr.sum() - distr_old / distr_old.sum()) ** 2   # :69-71
        converged = bool((resid < tol).all())                # :84
        distr_old = distr.copy()
        i += 1
    counts = distr - 1                                       # :93
    best = int(ks[int(np.argmax(counts))])                   # :95
    return best, counts, i
