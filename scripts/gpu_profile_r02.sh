#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU): launch lists and ncu --set full captures of the kernels DESIGN.md
# (the .ncu-rep files stay in /tmp on the box: gpurun brings back at most 64 MiB; the raw metric pages come back as csv)
# quotes.  Every ncu command follows a plain run of the same command line that exited 0.
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-bcv --no-d5 --no-secondary"
$B > gpurun_out/plain_als.log 2>&1 || { tail -5 gpurun_out/plain_als.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_als.csv $B > gpurun_out/ncu_als1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:als_stream_kernel -s 6 -c 2 -f -o /tmp/r02_prof_als $B > gpurun_out/ncu_als2.log 2>&1
ncu -i /tmp/r02_prof_als.ncu-rep --page raw --csv > gpurun_out/r02_prof_als_raw.csv 2>/dev/null
# FP32 / TF32 tensor-core stream kernel
F="python scripts/f32_profile.py"
$F > gpurun_out/plain_f32.log 2>&1 || { tail -5 gpurun_out/plain_f32.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:als_stream_tf32 -s 6 -c 2 -f -o /tmp/r02_prof_tf32 $F > gpurun_out/ncu_f32.log 2>&1
ncu -i /tmp/r02_prof_tf32.ncu-rep --page raw --csv > gpurun_out/r02_prof_tf32_raw.csv 2>/dev/null
# bcv cells on D3 (gather, Gram, Lanczos, top-k, products, errors): launch list + full capture of Lanczos and Gram
C="python scripts/bcv_profile.py"
$C > gpurun_out/plain_bcv.log 2>&1 || { tail -5 gpurun_out/plain_bcv.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bcv.csv $C > gpurun_out/ncu_bcv1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lanczos_kernel|gemm_tn_kernel" -s 20 -c 3 -f -o /tmp/r02_prof_bcv $C > gpurun_out/ncu_bcv2.log 2>&1
ncu -i /tmp/r02_prof_bcv.ncu-rep --page raw --csv > gpurun_out/r02_prof_bcv_raw.csv 2>/dev/null
# svd_pca on D4 (Gram 5000 x 5000 over 20000 rows + Lanczos N = 5000) and NIPALS
S="python scripts/svd_profile.py"
$S > gpurun_out/plain_svd.log 2>&1 || { tail -5 gpurun_out/plain_svd.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_svd.csv $S > gpurun_out/ncu_svd.log 2>&1
# NIPALS streaming kernel (one cooperative launch = the whole mfb_nipals call on D4): full capture
ncu --set full --clock-control none --import-source on -k regex:nipals_stream_kernel -c 1 -f -o /tmp/r02_prof_nipals $S > gpurun_out/ncu_nipals.log 2>&1
ncu -i /tmp/r02_prof_nipals.ncu-rep --page raw --csv > gpurun_out/r02_prof_nipals_raw.csv 2>/dev/null
# k = 64 (BASELINE configs[3]'s upper rank): stream, inversion and solve kernels of one iteration
K="python bench.py --k 64 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-bcv --no-d5 --no-secondary"
$K > gpurun_out/plain_k64.log 2>&1 || { tail -5 gpurun_out/plain_k64.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_k64.csv $K > gpurun_out/ncu_k64a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"als_stream_kernel|als_solve_kernel|als_shard_invert" -s 10 -c 6 -f -o /tmp/r02_prof_k64 $K > gpurun_out/ncu_k64b.log 2>&1
ncu -i /tmp/r02_prof_k64.ncu-rep --page raw --csv > gpurun_out/r02_prof_k64_raw.csv 2>/dev/null
for f in gpurun_out/ncu_*.log; do echo "== $f"; tail -2 $f; done
ls -la gpurun_out/*.csv
