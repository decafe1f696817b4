import sys, time, numpy as np
sys.path.insert(0, '.')
import mfbiclust_b200 as mfb
from oracle.gen_sim_data import gen_sim_data
from oracle.helpers import pseudovalues
from oracle.als_nmf import als_replicate as orep
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
side = min(m, n)//5
A = np.asfortranarray(pseudovalues(gen_sim_data(n=5, cluster_height=side, cluster_width=side, dimx=m, dimy=n, bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=4)))
rs = np.random.default_rng(40)
W0 = rs.random((m, k)); W0 /= np.sqrt((W0**2).sum(0))[None, :]
H0 = rs.random((k, n))
ctx = mfb.default_context()
Ad = mfb.Matrix(ctx, host=A)
sess = mfb.AlsSession(Ad, k, W0, H0, Ad.stats()["max"])
for i in range(int(sys.argv[4])):
    t0 = time.time()
    st, tot, conv = sess.iterate(1, fixed_iters=True)
    print("iter", i, "status", st, "total", tot, "dt %.4f" % (time.time()-t0), flush=True)
W, H, resid, tr = sess.result()
print("resid", resid, flush=True)
if m*n <= 4e6:
    ref = orep(A, W0, H0, max_iter=int(sys.argv[4]), fixed_iters=True)
    print("rel W %.2e rel H %.2e" % (np.linalg.norm(W-ref["W"])/np.linalg.norm(ref["W"]), np.linalg.norm(H-ref["H"])/np.linalg.norm(ref["H"])))
