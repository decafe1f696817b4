#!/bin/bash
# bench line, launch list and one full ncu capture of the two ALS half-step kernels (run under gpurun)
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || { tail -20 gpurun_out/bench_${TAG}.err; exit 1; }
cat gpurun_out/bench_${TAG}.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-bcv --no-secondary > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-bcv --no-secondary > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:als_s -s 12 -c 4 -f \
    -o gpurun_out/prof_${TAG} python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-bcv --no-secondary > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page details --csv > gpurun_out/prof_${TAG}_details.csv 2>/dev/null
tail -2 gpurun_out/ncu1.log; tail -2 gpurun_out/ncu2.log
# row-sharded session (world size 1, both exchange modes) and NIPALS: launch lists
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_shard_${TAG}.csv python scripts/shard_profile.py > gpurun_out/ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:nipals -c 2 --csv \
    --log-file gpurun_out/launches_nipals_${TAG}.csv python -c "
import sys; sys.path.insert(0, '.')
import numpy as np, mfbiclust_b200 as mfb
from bench import make_workload
A, _, _ = make_workload()
ctx = mfb.default_context(); Ad = mfb.Matrix(ctx, host=A)
import warnings; warnings.simplefilter('ignore')
r = mfb.nipals(Ad, 3, center=False, scale=True, maxiter=40, tol=1e-6, ctx=ctx); print(r['iter'])
" > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu3.log; tail -2 gpurun_out/ncu4.log
