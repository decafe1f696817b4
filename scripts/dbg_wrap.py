import sys, numpy as np
sys.path.insert(0, '.')
import mfbiclust_b200 as mfb
from oracle.als_nmf import als_replicate as orep
for (m, n, k) in [(19200,128,16),(19200,256,16),(3200,4096,16),(19200,256,8)]:
    rs = np.random.default_rng(1)
    A = np.asfortranarray(rs.random((m, n)) + 0.1)
    W0 = rs.random((m, k)); W0 /= np.sqrt((W0**2).sum(0))[None, :]
    H0 = rs.random((k, n))
    got = mfb.als_replicate(A, k, W0, H0, maxIter=1, fixed_iters=True)
    got2 = mfb.als_replicate(A, k, W0, H0, maxIter=1, fixed_iters=True)
    ref = orep(A, W0, H0, max_iter=1, fixed_iters=True)
    eH = np.abs(got["H"]-ref["H"]); eW = np.abs(got["W"]-ref["W"])
    print(m,n,k,"rel H %.2e rel W %.2e"%(np.linalg.norm(eH)/np.linalg.norm(ref["H"]), np.linalg.norm(eW)/np.linalg.norm(ref["W"])),
          "deterministic:", np.array_equal(got["H"],got2["H"]), np.array_equal(got["W"],got2["W"]),
          "bad H cols", np.flatnonzero(eH.max(0)>1e-9)[:12], "n bad", (eH.max(0)>1e-9).sum(), flush=True)
