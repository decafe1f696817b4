"""Restatement of ``svd_pca`` (``R/bicluster.R:528-539``), ``nipals_pca`` /
``nipals_pca_nocatch`` (``R/bicluster.R:218-289``), ``stats::prcomp`` (R 3.6.0)
and ``nipals::nipals`` (nipals 0.5, ``packrat/packrat.lock:613-615``, NOT
vendored; call site ``R/bicluster.R:226``).

Test infrastructure (see ``oracle/__init__.py``).  Specs: SURVEY.md App. A.2/A.3.

Pinned: ``svd_pca`` and ``nipals_pca`` at k=1 against
``tests/testdata/ref_bces.rda`` (W/H to ~1e-16, thresholds bit-exact).
``nipals`` with more than one component (deflation, Gram-Schmidt) and the
NA branch: parity unpinned -- follows the published nipals 0.5 source.
"""
from __future__ import annotations

import numpy as np

from .helpers import clean


def prcomp(x: np.ndarray, rank: int, center: bool):
    """``prcomp(x, rank. = rank, retx = TRUE, center = center)`` -> (x_scores, rotation, sdev)."""
    x = np.asarray(x, dtype=np.float64)
    if center:
        x = x - x.mean(axis=0, keepdims=True)
    k = min(int(rank), x.shape[0], x.shape[1])
    # La.svd -> LAPACK dgesdd (numpy.linalg.svd uses gesdd as well)
    _, d, vt = np.linalg.svd(x, full_matrices=False)
    rotation = vt[:k].T
    scores = x @ rotation
    sdev = d / np.sqrt(max(1, x.shape[0] - 1))
    return scores, rotation, sdev


def svd_pca(A: np.ndarray, k: int):
    """``svd_pca(A, k)``: ``prcomp(t(A), rank. = k, retx = TRUE, center = FALSE)``;
    W = rotation (m x k), H = t(x) (k x n)."""
    scores, rotation, _ = prcomp(np.asarray(A, dtype=np.float64).T, k, center=False)
    return rotation, scores.T


def nipals(x: np.ndarray, ncomp: int, center=True, scale=True, maxiter=500, tol=1e-6,
           gramschmidt=True):
    """``nipals::nipals`` (``startcol = 0``, ``fitted = FALSE``), complete data and the ``has.na`` branch
    (NaN = NA; SURVEY.md App. A.3: ``ph = x0't / colSums(T2)``, ``th = x0 p / colSums(P2)`` with the squared
    vectors zeroed at the missing cells).  The NA branch and ncomp > 1 are parity unpinned.

    Returns ``dict(eig, scores, loadings, iters)``; scores have unit-norm columns.
    """
    x = np.array(x, dtype=np.float64)
    nobs, nvar = x.shape
    miss = np.isnan(x)
    has_na = bool(miss.any())
    if has_na and (miss.all(axis=0).any() or miss.all(axis=1).any()):
        raise ValueError("At least one column or row is all NAs")
    if center:
        x = x - np.nanmean(x, axis=0, keepdims=True)
    if scale:
        x = x / np.nanstd(x, axis=0, ddof=1, keepdims=True)      # apply(x, 2, sd, na.rm = TRUE)
    PPp = np.zeros((nvar, nvar))
    TTp = np.zeros((nobs, nobs)) if nobs <= 4096 else None
    Ts = []
    eig = np.zeros(ncomp)
    loadings = np.zeros((nvar, ncomp))
    scores = np.zeros((nobs, ncomp))
    iters = np.zeros(ncomp, dtype=np.int64)
    obs = (~miss).astype(np.float64)
    for h in range(ncomp):
        scol = int(np.argmax(np.nansum(np.abs(x), axis=0)))
        x0 = np.where(miss, 0.0, x)
        th = x0[:, scol].copy()
        pciter = 1
        while True:
            if has_na:
                ph = (x0.T @ th) / (obs.T @ (th * th))
            else:
                ph = x.T @ th / float(th @ th)
            if gramschmidt and h > 0:
                ph = ph - PPp @ ph
            ph = ph / np.sqrt(float(ph @ ph))
            th_old = th
            if has_na:
                th = (x0 @ ph) / (obs @ (ph * ph))
            else:
                th = x @ ph / float(ph @ ph)
            if gramschmidt and h > 0:
                if TTp is not None:
                    th = th - TTp @ th
                else:  # same operator without the nobs x nobs matrix
                    for (t_prev, e_prev) in Ts:
                        th = th - t_prev * (float(t_prev @ th) / e_prev)
            if float(((th - th_old) ** 2).sum()) < tol:
                break
            pciter += 1
            if pciter == maxiter:
                break
        x = x - np.outer(th, ph)
        loadings[:, h] = ph
        scores[:, h] = th
        eig[h] = float(th @ th)
        iters[h] = pciter
        if gramschmidt:
            PPp += np.outer(ph, ph)
            if TTp is not None:
                TTp += np.outer(th, th) / eig[h]
            else:
                Ts.append((th.copy(), eig[h]))
    eig = np.sqrt(eig)
    scores = scores / eig[None, :]
    return dict(eig=eig, scores=scores, loadings=loadings, iters=iters)


def nipals_pca(A: np.ndarray, k: int, clean_param: float = 0.0, center=False, scale=True,
               maxiter=500):
    """``nipals_pca(A, k, cleanParam = 0, center = FALSE)`` as ``addStrat`` calls it
    (``R/addStrat.R:72``): ``clean`` then ``nipals::nipals(x, ncomp = k, tol = 1e-6,
    center = FALSE)`` -- ``scale`` is not forwarded, so nipals' default
    ``scale = TRUE`` applies.  Returns ``(W, H, good_rows, good_cols)``."""
    m_clean, good_rows, good_cols = clean(A, clean_param)
    res = nipals(m_clean, k, center=center, scale=scale, tol=1e-6, maxiter=maxiter)
    return res["scores"], res["loadings"].T, good_rows, good_cols
