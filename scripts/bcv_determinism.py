"""Run-to-run determinism of the bcv cells: the same folds must give bit-identical errors for any number of streams,
any batch composition and any history of the workspace pools."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mfbiclust_b200 as mfb
from mfbiclust_b200.backends import _bcv_cells_all
from bench import make_d3
Y = make_d3(); ks = np.arange(1, 21); ctx = mfb.default_context()
Yd = mfb.Matrix(ctx, host=Y)
rng = mfb.HostRng(51)
folds = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(12)]
ref = None
for s in (1, 1, 3, 3, 6, 3, 1):
    g = _bcv_cells_all(Yd, [int(k) for k in ks], 3, folds, ctx, None, s)
    if ref is None:
        ref = g
    d = np.abs(g - ref) / np.abs(ref)
    bad = np.argwhere(d.max(axis=-1) > 0)
    print("streams", s, "max rel diff vs first", d.max(), "cells differing", len(bad), bad[:6].tolist(), flush=True)
# single fold sets one at a time (fresh call each) against the batch
one = np.stack([_bcv_cells_all(Yd, [int(k) for k in ks], 3, [f], ctx, None, 1)[0] for f in folds])
d = np.abs(one - ref) / np.abs(ref)
print("one-at-a-time vs batch", d.max(), np.argwhere(d.max(axis=-1) > 0)[:6].tolist())
