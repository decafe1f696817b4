import sys, time, os; sys.path.insert(0, ".")
import numpy as np, mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3(); ctx = mfb.default_context(); Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21); rng = mfb.HostRng(50)
folds = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(4)]
mfb.bcv_batch(Yd, ks, 3, folds, ctx=ctx)
ref = None
for cap in ("", "112", "96", "72", "48", "32", ""):
    if cap: os.environ["MFB_TRI_BLOCKS"] = cap
    else: os.environ.pop("MFB_TRI_BLOCKS", None)
    for streams in (3, 1):
        t0 = time.perf_counter(); g = mfb.bcv_batch(Yd, ks, 3, folds, ctx=ctx, streams=streams); dt = time.perf_counter() - t0
        if ref is None: ref = g
        print("cap", cap or "148", "streams", streams, "36 cells", round(dt, 4), "max rel diff", float(np.abs(g - ref).max() / np.abs(ref).max()), flush=True)
