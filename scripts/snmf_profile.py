"""Development aid: ten snmf iterations (two convergence tests) at D4, k = 16 for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_workload
A, W0, H0 = make_workload()
m, n = A.shape
ctx = mfb.default_context()
Ad = mfb.Matrix(ctx, host=A)
mean = float(Ad.stats()["sum"]) / (float(m) * n)
sess = mfb.SnmfSession(Ad, 16, np.asfortranarray(np.random.default_rng(43).random((m, 16))), mean / 256.0, mean)
print(sess.iterate(10, bi_conv=(0, 10 ** 6)))
sess.close()
