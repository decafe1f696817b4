"""Development aid: one bcv call on D3 (9 cells, single stream) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3()
ctx = mfb.default_context()
Yd = mfb.Matrix(ctx, host=Y)
ks = np.arange(1, 21)
rf, cf = mfb.draw_folds(mfb.HostRng(50), *Y.shape, 3)
for _ in range(2):
    got = mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx, streams=1)
print(int(np.argmin(got)))
