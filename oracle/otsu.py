"""Restatement of ``otsuHelper`` / ``thresholdHelper`` / ``threshold``
(``R/BiclusterStrategy.R:442-456, 475-502, 508-528``) and of
``EBImage::otsu`` (EBImage 4.24.0, ``packrat/packrat.lock:81-83``, NOT vendored;
call site ``R/BiclusterStrategy.R:448``) on top of ``graphics::hist.default``.

Test infrastructure (see ``oracle/__init__.py``).

Spec: SURVEY.md App. A.5.  All histogram moments are dyadic rationals, so the
restatement is bit-exact when the operation order below is kept (no FMA).
Pinned by all six thresholds and both membership matrices of
``tests/testdata/ref_bces.rda`` (difference exactly 0.0) and by the vignette's
3x18 sample-membership matrix and feature-membership counts.
"""
from __future__ import annotations

import numpy as np

LEVELS = 256


def hist_counts_unit(x: np.ndarray) -> np.ndarray:
    """``hist(x, breaks = seq(0, 1, length.out = 257), plot = FALSE)$counts``.

    ``hist.default`` with ``right = TRUE, include.lowest = TRUE`` fuzzes the
    breaks by ``diddle = 1e-7 * median(diff(breaks))`` and bins with
    ``C_BinCount``: bin ``lo`` such that ``fuzzy[lo] < x <= fuzzy[lo+1]``
    (values equal to ``fuzzy[0]`` land in bin 0).
    """
    breaks = np.arange(LEVELS + 1, dtype=np.float64) / LEVELS
    diddle = 1e-7 * (1.0 / LEVELS)
    fuzzy = breaks + diddle
    fuzzy[0] = breaks[0] - diddle
    x = np.asarray(x, dtype=np.float64).ravel()
    # right-closed: index of first fuzzy[i] >= x, minus one
    idx = np.searchsorted(fuzzy, x, side="left") - 1
    idx = np.where(x == fuzzy[0], 0, idx)
    ok = (idx >= 0) & (idx < LEVELS)
    return np.bincount(idx[ok], minlength=LEVELS).astype(np.float64)


def otsu_from_counts(counts: np.ndarray) -> float:
    """``EBImage::otsu`` on a 256-bin histogram over ``range = c(0, 1)``."""
    mids = (2.0 * np.arange(LEVELS) + 1.0) / (2.0 * LEVELS)
    w1 = np.cumsum(counts)
    w2 = w1[-1] + counts - w1
    cm = counts * mids
    m1 = np.cumsum(cm)
    m2 = m1[-1] + cm - m1
    with np.errstate(divide="ignore", invalid="ignore"):
        d = m2 / w2 - m1 / w1
        var = (w1 * w2) * (d * d)          # R: w1 * w2 * (...)^2 -- `^` first, then left-to-right `*`
    finite = ~np.isnan(var)
    mx = var[finite].max()
    maxi = np.flatnonzero(finite & (var == mx))
    return float((mids[maxi[0]] + mids[maxi[-1]]) / 2.0)


def otsu_helper(mat: np.ndarray) -> np.ndarray:
    """``otsuHelper(matrix)``: one threshold per COLUMN (``R/BiclusterStrategy.R:442-456``)."""
    mat = np.asarray(mat, dtype=np.float64)
    out = np.empty(mat.shape[1])
    for j in range(mat.shape[1]):
        x = mat[:, j]
        mx, mn = x.max(), x.min()
        if mx != mn:
            rescaled = (x - mn) / (mx - mn)
            t = otsu_from_counts(hist_counts_unit(rescaled))
            out[j] = t * (mx - mn) + mn
        else:
            out[j] = x[0]
    return out


def threshold(m: np.ndarray, th: np.ndarray, margin: int = 2) -> np.ndarray:
    """``threshold(m, th, MARGIN)`` (``R/BiclusterStrategy.R:508-528``): strict compare,
    direction flips to ``<`` when the threshold is negative."""
    m = np.asarray(m, dtype=np.float64)
    th = np.asarray(th, dtype=np.float64)
    if margin == 1:
        if th.size != m.shape[0]:
            raise ValueError("Length of th must equal nrow(m).")
        t = th[:, None]
    else:
        if th.size != m.shape[1]:
            raise ValueError("Length of th must equal ncol(m)")
        t = th[None, :]
    return np.where(t < 0, m < t, m > t)


def threshold_helper(W: np.ndarray, H: np.ndarray):
    """``thresholdHelper`` with ``threshAlgo = "otsu"`` (``R/BiclusterStrategy.R:475-502``).

    Returns ``(featureThresh, sampleThresh, RowxNumber, NumberxCol)``."""
    ft = otsu_helper(W)              # :485 columns of W
    st = otsu_helper(H.T)            # :486 rows of H
    return ft, st, threshold(W, ft, 2), threshold(H, st, 1)
