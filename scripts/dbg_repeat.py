import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
import mfbiclust_b200 as mfb
from mfbiclust_b200._cabi import lib
from oracle.gen_sim_data import gen_sim_data
from oracle.helpers import pseudovalues
m, n, k = 20000, 5000, 16
side = min(m, n)//5
A = np.asfortranarray(pseudovalues(gen_sim_data(n=5, cluster_height=side, cluster_width=side, dimx=m, dimy=n, bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=4)))
rs = np.random.default_rng(40)
W0 = rs.random((m, k)); W0 /= np.sqrt((W0**2).sum(0))[None, :]
H0 = rs.random((k, n))
ctx = mfb.default_context()
Ad = mfb.Matrix(ctx, host=A)
sess = mfb.AlsSession(Ad, k, W0, H0, Ad.stats()["max"])
L = lib(); f = L.mfb_als_debug_repeat; f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_int, C.c_int]
import os
which = int(os.environ.get("DBG_KIND", "1"))
t0 = time.time(); r = f(sess._h, which, 3)
print("repeat kind", which, "->", r, L.mfb_last_error() if r else "", "%.3f s" % (time.time()-t0), flush=True)
