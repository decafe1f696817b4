"""Development aid: bcv batch time on D3 for several caps of the tridiagonalisation grid (MFB_TRI_BLOCKS) and stream
counts.  36 cells per batch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3(); ctx = mfb.default_context(); Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21); rng = mfb.HostRng(50)
folds = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(4)]
mfb.bcv_batch(Yd, ks, 3, folds, ctx=ctx, streams=6)
for cap in ("", "72", "48"):
    if cap: os.environ["MFB_TRI_BLOCKS"] = cap
    else: os.environ.pop("MFB_TRI_BLOCKS", None)
    for streams in (3, 4, 5, 6):
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter(); g = mfb.bcv_batch(Yd, ks, 3, folds, ctx=ctx, streams=streams); best = min(best, time.perf_counter() - t0)
        print("cap", cap or "default", "streams", streams, "36 cells", round(best, 4), flush=True)
