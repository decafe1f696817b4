"""Development aid: the convergence-test log of an snmf run on the device (MFB_SNMF_DEBUG=1, stderr) next to the oracle's."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["MFB_SNMF_DEBUG"] = "1"
from test_gpu_snmf import planted, relfro  # noqa: E402
from mfbiclust_b200 import backends as B  # noqa: E402
from oracle.snmf import kkt_residual, nmf_snmf  # noqa: E402

m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (300, 200, 5)
A = planted(m, n, k, 3)
W0 = np.random.default_rng(1).random((m, k))
log = []
ref = nmf_snmf(A, W0, eta=A.mean(), beta=A.mean(), log=log)
for e in log:
    if e["i"] < 20 or e["i"] > ref["iters"] - 40:
        print("oracle", e, e["conv"] / e["convnum"])
out = B.nmf_snmf(A, k, W0=W0, beta=A.mean(), eta=A.mean())
print(out["iters"], ref["iters"], relfro(out["W"], ref["W"]), relfro(out["H"], ref["H"]))
for it in (1, 5, 100):
    r = nmf_snmf(A, W0, eta=A.mean(), beta=A.mean(), max_iter=it)
    o = B.nmf_snmf(A, k, W0=W0, beta=A.mean(), eta=A.mean(), maxIter=it)
    print(it, relfro(o["W"], r["W"]), relfro(o["H"], r["H"]), kkt_residual(A, o["W"], o["H"], A.mean(), A.mean()))
