"""Development aid: svd_pca(D4, k = 16) and a short NIPALS run on the resident matrix, for an ncu launch list."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfbiclust_b200 as mfb
from bench import make_workload
A, _, _ = make_workload()
ctx = mfb.default_context()
Ad = mfb.Matrix(ctx, host=A)
for _ in range(2):
    r = mfb.svd_pca(Ad, 16, ctx=ctx)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    n = mfb.nipals(Ad, 3, center=False, scale=True, maxiter=40, tol=1e-6, ctx=ctx)
print(r.fit.W.shape, n["iter"])
