"""``genSimData``-shaped synthetic matrices (``R/genBiclusters.R:74-195``).

Test / benchmark infrastructure (see ``oracle/__init__.py``).  The order of
effects follows ``genSimDataHelper`` (:147-192) and the block extents follow
``genSimData`` (:100-106, including the height/width swap at :101,:103); the
normal draws come from NumPy's ``default_rng`` (R's ``rnorm`` stream is not
reproduced -- SURVEY.md section 8d), so this defines the input SHAPE of
BASELINE.json's configs 3-5, not a bit-level restatement.
"""
from __future__ import annotations

import numpy as np


def gen_sim_data(n=1, cluster_height=20, cluster_width=20, dimx=80, dimy=80,
                 overlap_rows=0, overlap_cols=0, bg_const=0.0, bg_norm=0.0,
                 bicluster_constant=None, bicluster_shift=0.0, row_base=0.0,
                 row_shift=0.0, row_scale=None, col_shift=0.0, col_scale=None,
                 shuffle=True, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x_starts = (cluster_height - overlap_rows) * np.arange(n)
    x_stops = x_starts + cluster_width            # :101 (width used for row extent)
    y_starts = (cluster_width - overlap_cols) * np.arange(n)
    y_stops = y_starts + cluster_height           # :103
    res = np.full((dimx, dimy), float(bg_const))
    if row_base:
        res += rng.normal(0.0, row_base, size=(dimx, 1))          # :150-152
    else:
        res += 0.0
    if bg_norm:
        res += bg_norm * rng.standard_normal((dimx, dimy))         # :155
    if bicluster_constant is not None:
        for a, b, c, d in zip(x_starts, x_stops, y_starts, y_stops):
            res[a:b, c:d] = bicluster_constant                     # :158-164
    for a, b, c, d in zip(x_starts, x_stops, y_starts, y_stops):
        res[a:b, c:d] += rng.normal(0.0, bicluster_shift) if bicluster_shift else 0.0   # :168
        res[a:b, c:d] += (rng.normal(0.0, col_shift, size=d - c) if col_shift
                          else np.zeros(d - c))[None, :]           # :172-173
        if col_scale is not None:
            res[a:b, c:d] *= rng.normal(0.0, col_scale, size=d - c)[None, :]   # :177-181
        res[a:b, c:d] += (rng.normal(0.0, row_shift, size=b - a) if row_shift
                          else np.zeros(b - a))[:, None]           # :184-185
        if row_scale is not None:
            res[a:b, c:d] *= rng.normal(0.0, row_scale, size=b - a)[:, None]   # :187-190
    if shuffle:
        res = res[rng.permutation(dimx)][:, rng.permutation(dimy)]  # :118-121
    return np.asfortranarray(res.astype(dtype))
