import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mfbiclust_b200 as mfb
from bench import make_workload, peaks
ctx = mfb.default_context()
A, W0, H0 = make_workload()
m, n = A.shape
os.environ["MFB_TF32_TIMING"] = "1"
Af = mfb.Matrix(ctx, host=A, precision="f32")
sess = mfb.AlsSession(Af, 16, W0, H0, Af.stats()["max"])
st, _, _ = sess.iterate(5, fixed_iters=True)
ctx.sync()
t0 = time.perf_counter()
st, _, _ = sess.iterate(40, fixed_iters=True)
ctx.sync()
dt = time.perf_counter() - t0
kern = [float(v) for v in sess.profile(10)]
sess.close(); Af.free()
gbs = 2.0 * m * n * 4 * 40 / dt / 1e9
print("D4 tf32 iters/s", 40 / dt, "frac of FP32 roofline", gbs / peaks()[0], "kernel_ms", kern, flush=True)
