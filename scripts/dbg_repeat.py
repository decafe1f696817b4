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
print(sess.iterate(3, fixed_iters=True))
for which in (16*16+0, 16*16+1):
    r = f(sess._h, which, 1)
    buf = (C.c_ulonglong * (148*8+24))()
    L.mfb_als_debug_times(buf)
    marks = np.array(buf[148*8:148*8+6], dtype=np.float64); print('   epilogue marks (us):', np.diff(marks)/1e3); nm = np.array(buf[148*8+8:148*8+16], dtype=np.float64); print('   nnls round starts (us):', (nm-nm[0])/1e3); fm = np.array(buf[148*8+16:148*8+20], dtype=np.float64); print('   finalize phases (us): gram-reduce, gauss-jordan, ctrl:', np.diff(fm)/1e3, 'finalize start rel. kernel start %.1f' % ((fm[0]-np.array(buf[:148*8:8], dtype=np.float64).min())/1e3))
    t = np.array(buf[:148*8], dtype=np.float64).reshape(148, 8)
    t0 = t[:,0].min()
    for name, col in (("start",0),("mid-flush begin",1),("mid-flush end",3),("stream end",2),("last flush end",4)):
        v = t[:,col]; v = v[v>0]
        if len(v): print("kind", which & 15, "%-16s n=%3d min %7.1f med %7.1f max %7.1f us" % (name, len(v), (v.min()-t0)/1e3, (np.median(v)-t0)/1e3, (v.max()-t0)/1e3))
    d = (t[:,4]-t[:,2])/1e3
    print("   final flush duration: med %.1f max %.1f us; top5:" % (np.median(d), d.max()), np.sort(d)[-5:])
    dm = (t[:,3]-t[:,1])/1e3; dm = dm[t[:,1]>0]
    if len(dm): print("   mid flush duration: med %.1f max %.1f us; top5:" % (np.median(dm), dm.max()), np.sort(dm)[-5:])
