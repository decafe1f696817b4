"""Development aid: stage times of a bcv batch on D3 (MFB_BCV_TIMING) + event-timed eigen-solver kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3(); ctx = mfb.default_context(); Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21); rng = mfb.HostRng(50)
fl = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(4)]
mfb.bcv_batch(Yd, ks, 3, fl, ctx=ctx)
os.environ["MFB_BCV_TIMING"] = "1"; os.environ["MFB_LZ_VERBOSE"] = "1"; os.environ["MFB_LZ_PHASES"] = "1"
for batch in ("9", "12", "16"):
    os.environ["MFBICLUST_BCV_BATCH"] = batch
    print("---- batch", batch, flush=True); sys.stderr.flush()
    t0 = time.perf_counter(); mfb.bcv_batch(Yd, ks, 3, fl, ctx=ctx); print("36 cells", time.perf_counter() - t0, flush=True)
for k in ("MFB_BCV_TIMING", "MFB_LZ_VERBOSE", "MFB_LZ_PHASES"):
    os.environ.pop(k, None)
for batch in ("6", "8", "9", "12", "16", "18"):
    os.environ["MFBICLUST_BCV_BATCH"] = batch
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); mfb.bcv_batch(Yd, ks, 3, fl, ctx=ctx); best = min(best, time.perf_counter() - t0)
    print("batch", batch, "36 cells", round(best, 4), flush=True)
os.environ.pop("MFBICLUST_BCV_BATCH", None)
for _ in range(2):
    t0 = time.perf_counter()
    res = mfb.auto_bcv(Yd, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx)
    print("auto_bcv D3", round(time.perf_counter() - t0, 4), res["best"], res["iterations"], flush=True)
