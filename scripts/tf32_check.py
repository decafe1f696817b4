"""Development aid: the TF32 tensor-core path of als_nmf (A resident in FP32) against the FP64 oracle, small then D4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from oracle.als_nmf import als_replicate as orep
from oracle.gen_sim_data import gen_sim_data
from oracle.helpers import pseudovalues


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


ctx = mfb.default_context()
for (m, n, k) in [(640, 384, 8), (2000, 500, 16), (1037, 333, 10), (70, 45, 3), (6, 6, 1)]:
    A = pseudovalues(gen_sim_data(n=4, cluster_height=max(1, min(m, n) // 5), cluster_width=max(1, min(m, n) // 5), dimx=m, dimy=n,
                                  bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    rs = np.random.default_rng(44)
    W0 = rs.random((m, k)); W0 /= np.sqrt((W0 ** 2).sum(0))[None, :]; H0 = rs.random((k, n))
    ref = orep(A, W0, H0, max_iter=8, fixed_iters=True)
    for mode in ("tf32", "tf32_nocomp", "dmma"):
        os.environ.pop("MFB_F32_DMMA", None); os.environ.pop("MFB_TF32_NOCOMP", None)
        if mode == "dmma":
            os.environ["MFB_F32_DMMA"] = "1"
        if mode == "tf32_nocomp":
            os.environ["MFB_TF32_NOCOMP"] = "1"
        Ad = mfb.Matrix(ctx, host=A, precision="f32")
        try:
            got = mfb.als_replicate(Ad, k, W0, H0, maxIter=8, fixed_iters=True)
            print((m, n, k), mode, "W", rel(got["W"], ref["W"]), "H", rel(got["H"], ref["H"]), "resid", got["obj"], ref["obj"], flush=True)
        except Exception as e:
            print((m, n, k), mode, "FAILED", e, flush=True)
        Ad.free()
os.environ.pop("MFB_F32_DMMA", None)
if "--small" in sys.argv:
    sys.exit(0)
from bench import make_workload, peaks
A, W0, H0 = make_workload()
m, n = A.shape
for mode in ("tf32", "tf32_nocomp", "dmma"):
    os.environ.pop("MFB_F32_DMMA", None); os.environ.pop("MFB_TF32_NOCOMP", None)
    if mode == "dmma":
        os.environ["MFB_F32_DMMA"] = "1"
    if mode == "tf32_nocomp":
        os.environ["MFB_TF32_NOCOMP"] = "1"
    Af = mfb.Matrix(ctx, host=A, precision="f32")
    sess = mfb.AlsSession(Af, 16, W0, H0, Af.stats()["max"])
    st, _, _ = sess.iterate(5, fixed_iters=True)
    ctx.sync()
    t0 = time.perf_counter()
    st, _, _ = sess.iterate(40, fixed_iters=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    kern = [float(v) for v in sess.profile(10)]
    Wg, Hg, resid, _ = sess.result()
    sess.close(); Af.free()
    gbs = 2.0 * m * n * 4 * 40 / dt / 1e9
    print("D4", mode, "iters/s", 40 / dt, "frac of FP32 roofline", gbs / peaks()[0], "kernel_ms", kern, "resid", resid, flush=True)
