"""Development aid (run under gpurun): the Lanczos eigen-solver behind svd_pca / bcv -- parity on small cases, then
wall times of svd_pca(D4), one bcv call on D3, a 36-cell batch and auto_bcv(D3) for a few grid caps (MFB_LZ_BLOCKS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from oracle import bcv as B, pca as P
from bench import make_d3, make_workload

ctx = mfb.default_context()
rs = np.random.default_rng(0)


def folds(rs, n, p, h):
    return (B.make_folds(n, h, rs.permutation(n)).astype(np.int32), B.make_folds(p, h, rs.permutation(p)).astype(np.int32))


# ---- parity, small ----
for (m, n, k) in [(300, 40, 5), (40, 300, 5), (500, 200, 30), (1000, 17, 10), (64, 64, 63), (300, 120, 100)]:
    A = np.asfortranarray(rs.standard_normal((m, 6)) @ rs.standard_normal((6, n)) + rs.standard_normal((m, n)))
    t0 = time.perf_counter()
    fit = mfb.svd_pca(A, k)
    dt = time.perf_counter() - t0
    Wr, Hr = P.svd_pca(A, k)
    sgn = np.sign((fit.fit.W * Wr).sum(axis=0))
    print("svd_pca", (m, n, k), "W err", float(np.abs(fit.fit.W * sgn[None, :] - Wr).max()),
          "H rel", float(np.abs(fit.fit.H * sgn[:, None] - Hr).max() / np.abs(Hr).max()), "t", round(dt, 4), flush=True)
for (n, p, h, kmax) in [(900, 300, 3, 20), (40, 150, 4, 8), (1000, 200, 3, 132), (120, 45, 3, 12)]:
    if (n, p) == (120, 45):
        Y = rs.standard_normal((n, 3)) @ rs.standard_normal((3, p))
    else:
        Y = rs.standard_normal((n, 5)) @ rs.standard_normal((5, p)) + rs.standard_normal((n, p))
    Y = np.asfortranarray(Y)
    ks = np.arange(1, kmax + 1)
    rf, cf = folds(rs, n, p, h)
    t0 = time.perf_counter()
    got = mfb.bcvGivenKs(Y, ks, h, folds=(rf, cf))
    dt = time.perf_counter() - t0
    ref = B.bcv_given_ks(Y, ks, rf, cf, h)
    print("bcv", (n, p, h, kmax), "curve rel err", float(np.abs(got - ref).max() / np.abs(ref).max()), "t", round(dt, 4), flush=True)

if "--small" in sys.argv:
    sys.exit(0)

# ---- D3 ----
Y = make_d3()
Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21)
rf, cf = mfb.draw_folds(mfb.HostRng(50), *Y.shape, 3)
got = mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx)
t0 = time.perf_counter()
ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
print("D3 oracle bcv call", round(time.perf_counter() - t0, 2), "s; curve rel err",
      float(np.abs(got / ref - 1).max()), "argmin", int(np.argmin(got)), int(np.argmin(ref)), flush=True)
rng = mfb.HostRng(50)
fl = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(4)]
def timed(label):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx)
        best = min(best, time.perf_counter() - t0)
    b36 = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        mfb.bcv_batch(Yd, ks, 3, fl, ctx=ctx)
        b36 = min(b36, time.perf_counter() - t0)
    print(label, "bcv call (9 cells)", round(best, 4), "36 cells", round(b36, 4), flush=True)


for batch in ("16",):
    os.environ["MFBICLUST_BCV_BATCH"] = batch
    for aux in ("4",):
        os.environ["MFBICLUST_BCV_AUX_STREAMS"] = aux
        timed("batch %s aux %s" % (batch, aux))
os.environ["MFBICLUST_BCV_AUX_STREAMS"] = "4"
for batch in ():
    os.environ["MFBICLUST_BCV_BATCH"] = batch
    for cap in ("8", "12", "16", "24", "32"):
        os.environ["MFB_LZ_BLOCKS"] = cap
        timed("batch %s lz_blocks %s" % (batch, cap))
    os.environ.pop("MFB_LZ_BLOCKS", None)
os.environ.pop("MFBICLUST_BCV_BATCH", None)
os.environ.pop("MFBICLUST_BCV_AUX_STREAMS", None)
os.environ.pop("MFB_LZ_BLOCKS", None)
for _ in range(2):
    t0 = time.perf_counter()
    res = mfb.auto_bcv(Yd, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx)
    print("auto_bcv D3", round(time.perf_counter() - t0, 4), res["best"], res["iterations"], flush=True)
Yd.free()

# ---- D4 svd_pca ----
A, _, _ = make_workload()
Ad = mfb.Matrix(ctx, host=A)
for _ in range(3):
    t0 = time.perf_counter()
    r = mfb.svd_pca(Ad, 16, ctx=ctx)
    print("svd_pca D4 k=16", round(time.perf_counter() - t0, 4), flush=True)
u, s, vt = np.linalg.svd(A[:, :], full_matrices=False) if "--svd-check" in sys.argv else (None, None, None)
if u is not None:
    sgn = np.sign((r.fit.W * u[:, :16]).sum(axis=0))
    print("svd_pca D4 W err", float(np.abs(r.fit.W * sgn[None, :] - u[:, :16]).max()))
Ad.free()
