"""Restatement of the host-side pre-steps around the hot path
(``R/helperFunctions.R:89-109`` ``clean``, ``:270-278`` ``pseudovalues``,
``:367-412`` ``validateK`` / ``validateKM``) and of the post-processing after it
(``:25-31`` ``biclusterMatrix2List``, ``:150-206`` ``filter.biclust``, ``:226-247`` ``overlap``,
``:303-308`` ``sizes``, ``:336-345`` ``union.biclust``).  Test infrastructure
(see ``oracle/__init__.py``).  Parity unpinned (no artefact exercises them
beyond the vignette's ``+ abs(min)`` shift, which is pinned end to end).
"""
from __future__ import annotations

import numpy as np


def pseudovalues(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    if (m < 0).any():
        return m + abs(m.min())
    return m


def r_round(x: float) -> float:
    # R's round(): IEC 60559 half-even in 4.x; 3.6.0 rounds half away from zero
    # for exactly representable .5 -- only integers*(1-cleanParam) occur here.
    return float(np.floor(x + 0.5))


def clean(obj: np.ndarray, clean_param: float = 0.0):
    """``clean(object, cleanParam, dimsRemain = TRUE)``."""
    if not (0 <= clean_param <= 1):
        raise ValueError('Arg "cleanParam" must be in the range of 0 to 1.')
    obj = np.asarray(obj, dtype=np.float64)
    max_na_row = r_round((1 - clean_param) * obj.shape[1])
    max_na_col = r_round((1 - clean_param) * obj.shape[0])
    bad = np.isnan(obj) | (obj == 0)
    good_rows = bad.sum(axis=1) < max_na_row
    good_cols = bad.sum(axis=0) < max_na_col
    return obj[np.ix_(good_rows, good_cols)], good_rows, good_cols


def validate_km(k, m: np.ndarray, method: str) -> int:
    """``validateKM`` incl. its quirk: the literal ``"method" == "svd-pca"``
    (``R/helperFunctions.R:402``) means svd-pca is never clamped here."""
    k = int(round(float(k)))
    if not k > 0:
        raise ValueError('Argument "k" must be positive.')
    if method in ("als-nmf", "nipals-pca", "snmf"):
        if k > min(m.shape):
            return min(m.shape)
    return k


# ---------------------------------------------------------------------------
# post-processing of membership matrices (literal restatement: index lists and set intersections, as the R code)
# ---------------------------------------------------------------------------
def bicluster_matrix2list(rowx, bicx):
    rowx, bicx = np.asarray(rowx, dtype=bool), np.asarray(bicx, dtype=bool)
    return ([np.flatnonzero(rowx[:, b]) + 1 for b in range(rowx.shape[1])],          # which(): 1-based
            [np.flatnonzero(bicx[b, :]) + 1 for b in range(bicx.shape[0])])


def sizes(bic_rows, bic_cols):
    return np.array([float(len(r) * len(c)) for r, c in zip(bic_rows, bic_cols)])     # :303-308


def overlap(bic_rows, bic_cols):
    """``overlap(BiclusterRows, BiclusterCols, matrix = TRUE)`` (:226-247): row i = overlaps of bicluster i."""
    sz = sizes(bic_rows, bic_cols)
    k = len(bic_rows)
    out = np.zeros((k, k))
    for i in range(k):
        for j in range(k):
            i1 = len(set(bic_rows[i].tolist()) & set(bic_rows[j].tolist()))
            i2 = len(set(bic_cols[i].tolist()) & set(bic_cols[j].tolist()))
            with np.errstate(invalid="ignore", divide="ignore"):
                out[i, j] = np.float64(i1 * i2) / sz[i]                               # 0 / 0 = NaN as in R
    return out


def union_biclust(rowx, bicx, choose=None):
    rowx, bicx = np.asarray(rowx, dtype=bool), np.asarray(bicx, dtype=bool)
    k = rowx.shape[1]
    choose = np.ones(k, dtype=bool) if choose is None else np.asarray(choose, dtype=bool)
    out = np.zeros((rowx.shape[0], bicx.shape[1]))
    for b in np.flatnonzero(choose):                                                  # lapply(which(choose), ...)
        out[np.ix_(rowx[:, b], bicx[b, :])] = 1.0
    return out


def filter_biclust(rowx, bicx, max=None, overlap_limit=0.25):
    """``filter.biclust`` (:150-206)."""
    rowx, bicx = np.asarray(rowx, dtype=bool), np.asarray(bicx, dtype=bool)
    if rowx.shape[1] != bicx.shape[0]:
        raise ValueError("RowxBicluster must have the same number of columns asBiclusterxCol")
    k = rowx.shape[1]
    if k == 0:
        chosen = np.zeros(0, dtype=bool)
    elif k == 1:
        chosen = np.ones(1, dtype=bool)
    else:
        rows, cols = bicluster_matrix2list(rowx, bicx)
        chosen = np.zeros(k, dtype=bool)
        pool = np.ones(k, dtype=bool)
        sz = sizes(rows, cols)
        pool[(sz == 0) | (sz == rowx.shape[0] * bicx.shape[1])] = False
        ov = overlap(rows, cols)
        while (max is None or chosen.sum() < max) and pool.sum() > 0:
            idx = np.flatnonzero(pool)
            choose_me = idx[int(np.argmax(sz[idx]))]
            chosen[choose_me] = True
            with np.errstate(invalid="ignore"):
                pool[ov[:, choose_me] > overlap_limit] = False
            pool[choose_me] = False
    return {"RowxBicluster": rowx[:, chosen], "BiclusterxCol": bicx[chosen, :],
            "biclustered": union_biclust(rowx, bicx, chosen), "chosen": chosen}
