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
