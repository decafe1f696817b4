#!/bin/bash
for c in "2560 8192" "20000 640" "20000 1280" "20000 2560" "10000 5000" "5120 5000" "20000 5000"; do
  set -- $c
  timeout 60 python - <<PY 2>&1 | tail -1
import sys, numpy as np
sys.path.insert(0, '.')
import mfbiclust_b200 as mfb
m, n, k = $1, $2, 16
rs = np.random.default_rng(1)
A = np.asfortranarray(rs.random((m, n)) + 0.1)
W0 = rs.random((m, k)); W0 /= np.sqrt((W0**2).sum(0))[None, :]
H0 = rs.random((k, n))
try:
    got = mfb.als_replicate(A, k, W0, H0, maxIter=2, fixed_iters=True)
    print(m, n, "ok", got["obj"])
except Exception as e:
    print(m, n, "FAIL", str(e)[:100])
PY
done
