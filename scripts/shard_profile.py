"""Per-kernel timings of a row-sharded ALS session at world size 1 (no peers to wait for): what the extra kernels of
the sharded path cost by themselves.  python scripts/shard_profile.py [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mfbiclust_b200 as mfb
from bench import make_workload
k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
A, W0, H0 = make_workload(k=k)
ctx = mfb.default_context()
Ad = mfb.Matrix(ctx, host=A)
amax = Ad.stats()["max"]
for exchange in ("peer", "callback"):
    sess = mfb.ShardedAlsSession(Ad, A.shape[0], k, W0, H0, amax, exchange=exchange)
    sess.iterate(5, fixed_iters=True)
    print(exchange, [round(float(v) * 1e3, 1) for v in sess.profile(10)], flush=True)
    sess.close()
sess = mfb.AlsSession(Ad, k, W0, H0, amax)
sess.iterate(5, fixed_iters=True)
print("one-device", [round(float(v) * 1e3, 1) for v in sess.profile(10)], flush=True)
import ctypes as C
sess = mfb.ShardedAlsSession(Ad, A.shape[0], k, W0, H0, amax, exchange="peer")
sess.iterate(5, fixed_iters=True)
st = (C.c_ulonglong * 8)()
L = mfb._cabi.lib()
L.mfb_als_shard_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.mfb_als_shard_stamps(sess._h, st, 1)
sess.iterate(1, fixed_iters=True)
L.mfb_als_shard_stamps(sess._h, st, 0)
t = [int(v) for v in st]
print("stamps (us rel. to first block done): elected %.1f waited %.1f" %
      tuple((t[i] - t[4]) / 1e3 for i in range(2)))
