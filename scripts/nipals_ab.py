"""Development aid: time mfb_nipals on D4 (and a few odd shapes) and print a digest of the results; run once with
MFB_NIPALS_OLD=1 and once without to compare the two complete-data kernels."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfbiclust_b200 as mfb
from bench import make_workload
ctx = mfb.default_context()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    rs = np.random.default_rng(3)
    for (mm, nn, k) in [(1000, 24, 5), (1001, 7, 3), (333, 700, 4), (5000, 300, 6), (70000, 40, 3), (20000, 64, 2), (2000, 5000, 2)]:
        x = rs.standard_normal((mm, nn)) + np.outer(rs.standard_normal(mm), rs.standard_normal(nn)) * 2
        t0 = time.perf_counter()
        r = mfb.nipals(np.asfortranarray(x), k, center=True, scale=False, tol=1e-9, ctx=ctx)
        dt = time.perf_counter() - t0
        print((mm, nn, k), "%.2f ms" % (dt * 1e3), [int(i) for i in r["iter"]], np.round(r["eig"], 9), flush=True)
    A, _, _ = make_workload()
    Ad = mfb.Matrix(ctx, host=A)
    m, n = A.shape
    for rep in range(3):
        t0 = time.perf_counter()
        res = mfb.nipals(Ad, 3, center=False, scale=True, maxiter=40, tol=1e-6, ctx=ctx)
        dt = time.perf_counter() - t0
    it = [int(i) for i in res["iter"]]
    passes = 1 + 3 + 1 + sum(2 * i for i in it) + 2 * (len(it) - 1)
    print("D4 nipals %.2f ms iters %s  %.0f GB/s" % (dt * 1e3, it, passes * m * n * 8 / dt / 1e9))
    print("  eig", res["eig"], "scores[0,:]", res["scores"][0], "load[0,:]", res["loadings"][0])
