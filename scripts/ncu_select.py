"""Condense the raw metric pages of the round's ncu --set full captures (gpurun_out/r02_prof_*_raw.csv, written by
scripts/gpu_profile_r02.sh) into profiles/r02_ncu_full_selected.csv: one row per (capture, kernel launch, metric)."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAPTURES = [("als FP64", "r02_prof_als_raw.csv"), ("als FP32 / TF32 tensor cores", "r02_prof_tf32_raw.csv"),
            ("bcv cell on D3", "r02_prof_bcv_raw.csv"), ("nipals on D4 (one cooperative launch)", "r02_prof_nipals_raw.csv"),
            ("als FP64 k = 64", "r02_prof_k64_raw.csv")]
METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_active.max",
           "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
out = csv.writer(open(os.path.join(ROOT, "profiles", "r02_ncu_full_selected.csv"), "w", newline=""))
out.writerow(["capture", "kernel", "metric", "value", "unit"])
for name, fn in CAPTURES:
    path = os.path.join(ROOT, "gpurun_out", fn)
    if not os.path.exists(path):
        print("missing", path, file=sys.stderr)
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        kern = r[hdr.index("Kernel Name")]
        short = kern.split("(")[0]
        out.writerow([name, short, "Kernel Name", kern, ""])
        for m in METRICS:
            if m in hdr:
                out.writerow([name, short, m, r[hdr.index(m)], units[hdr.index(m)]])
