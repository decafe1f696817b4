#!/bin/bash
# A/B of several builds of the library (development aid): [K=64] bash scripts/ab.sh lib1.so lib2.so ...
for v in "$@"; do
  MFBICLUST_LIB=$PWD/mfbiclust_b200/lib/$v python bench.py --k ${K:-16} --no-cpu-baseline --no-e2e --no-bcv --no-d5 --no-secondary --steps ${STEPS:-40} 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$v', round(d['value'], 1), {k: round(v * 1e3, 1) for k, v in d['roofline']['kernel_ms'].items()}, round(d['roofline']['step_frac'], 3))"
done
