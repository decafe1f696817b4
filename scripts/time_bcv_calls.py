import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3()
ctx = mfb.default_context()
Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21)
rf, cf = mfb.draw_folds(mfb.HostRng(50), *Y.shape, 3)
for i in range(3):
    t0 = time.perf_counter()
    r = mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx)
    print("bcv call", i, round(time.perf_counter() - t0, 3), int(np.argmin(r)) + 1, flush=True)
if "--check" in sys.argv:
    from oracle import bcv as B
    ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
    print("curve rel err %.2e" % (np.abs(r - ref).max() / np.abs(ref).max()))
