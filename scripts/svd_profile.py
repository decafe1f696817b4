import sys; sys.path.insert(0, '.')
import numpy as np, mfbiclust_b200 as mfb
from bench import make_workload
A, _, _ = make_workload()
ctx = mfb.default_context(); Ad = mfb.Matrix(ctx, host=A)
r = mfb.svd_pca(Ad, 16, ctx=ctx)
