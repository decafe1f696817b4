#!/bin/bash
# launch list + one full capture of the two ALS half-step kernels (run under gpurun)
set -e
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:als_half -c 14 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:als_half -s 6 -c 2 \
    -o gpurun_out/prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/plain.log
