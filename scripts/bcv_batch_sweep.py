"""Development aid: auto_bcv on D3 for several eigen-solver batch sizes (MFBICLUST_BCV_BATCH)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3(); ks = np.arange(1, 21); ctx = mfb.default_context()
Yd = mfb.Matrix(ctx, host=Y)
for B in (16, 6, 8, 10, 12, 16, 24, 32):
    os.environ["MFBICLUST_BCV_BATCH"] = str(B)
    best = None
    for rep in range(2):
        t0 = time.perf_counter(); r = mfb.auto_bcv(Yd, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx); ctx.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print("batch", B, round(best, 4), r["best"], r["iterations"], flush=True)
