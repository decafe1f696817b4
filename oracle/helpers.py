"""Restatement of the host-side pre-steps around the hot path
(``R/helperFunctions.R:89-109`` ``clean``, ``:270-278`` ``pseudovalues``,
``:367-412`` ``validateK`` / ``validateKM``).  Test infrastructure
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
