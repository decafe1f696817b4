"""Development aid: a few ALS iterations at D4, k = 16 with A resident in FP32 (TF32 tensor-core stream kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfbiclust_b200 as mfb
from bench import make_workload
A, W0, H0 = make_workload()
ctx = mfb.default_context()
Af = mfb.Matrix(ctx, host=A, precision="f32")
sess = mfb.AlsSession(Af, 16, W0, H0, Af.stats()["max"])
st, _, _ = sess.iterate(6, fixed_iters=True)
ctx.sync()
print(st, sess.result()[2])
