"""Small-shape run of every entry point for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
rs = np.random.default_rng(0)
for (m, n, k) in [(130, 33, 7), (300, 200, 16), (257, 140, 40)]:
    A = np.asfortranarray(rs.random((m, n)) + 0.05)
    W0 = rs.random((m, k)); W0 /= np.sqrt((W0 ** 2).sum(0))[None, :]
    H0 = rs.random((k, n))
    r = mfb.als_replicate(A, k, W0, H0, maxIter=3, fixed_iters=True)
    print("als", m, n, k, r["obj"])
A = rs.standard_normal((300, 40)) + 1.0
print("nipals", mfb.nipals(A, 3, center=False, scale=True)["iter"])
print("svd", mfb.svd_pca(A, 5).fit.W.shape, mfb.svd_pca(A.T.copy(), 5).fit.W.shape)
rf, cf = mfb.draw_folds(mfb.HostRng(1), 300, 40, 3)
print("bcv", mfb.bcvGivenKs(A, np.arange(1, 9), 3, folds=(rf, cf))[:3])
th = mfb.otsuHelper(A); print("otsu", th[:3], mfb.threshold(A, th, 2).sum())
print("nnls", mfb.nnls(rs.random((50, 6)), rs.standard_normal((50, 9))).shape)
A = np.asfortranarray(rs.random((150, 70)) + 0.05)
r = mfb.nmf_snmf(A, 5, W0=rs.random((150, 5)), beta=float(A.mean()), eta=float(A.mean()), maxIter=12)
print("snmf", r["iters"], r["objective"])
