#!/bin/bash
# A/B of several builds of the library (development aid): bash scripts/ab.sh lib1.so lib2.so ...
for v in "$@"; do
  cp mfbiclust_b200/lib/$v /tmp/cur.so
  python - <<PY 2>&1 | tail -2
import mfbiclust_b200._cabi as c
c.LIB_PATH = "/tmp/cur.so"
import sys, json
sys.argv = ["bench.py", "--no-cpu-baseline", "--no-e2e", "--no-bcv", "--no-secondary", "--steps", "40"]
import runpy, io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    runpy.run_path("bench.py", run_name="__main__")
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print("$v", round(d["value"], 1), {k: round(v * 1e3, 1) for k, v in d["roofline"]["kernel_ms"].items()}, round(d["roofline"]["step_frac"], 3))
PY
done
