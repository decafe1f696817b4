#!/usr/bin/env python
"""Headline benchmark: ALS-NMF iterations/s at 20 000 x 5 000, k = 16, FP64 (BASELINE.json configs[3]).

A "step" is one ALS iteration of ``als_nmf``'s inner loop (H-step + W-step + residual /
convergence scalars, ``R/bicluster.R:119-203``) on the synthetic genSimData-shaped matrix D4 of
SURVEY.md section 8(d), convergence break disabled (fixed-iteration mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* ``value``  : iterations/s with A resident in HBM, timed with CUDA events on the library's stream.
               N = 1: one D4.  N > 1 (torchrun): ONE (20000 N) x 5000 matrix whose rows are sharded over the ranks
               (every rank holds a D4-sized row block; per iteration the k x n and k x k partial products and three
               scalars are all-reduced over NVLink peer memory inside the kernels): ``value`` = N x the iterations/s of
               that matrix = D4-sized units per second over all ranks, ``scaling`` = weak.
* ``e2e``    : the same metric through the public call with HOST buffers (upload of A from pinned memory, K
               iterations, download of W / H inside the timed region).
* ``roofline``: achieved HBM GB/s of the streaming kernel (2 launches per iteration, each reads A once: algorithmic
               bytes = m*n*8 per launch) against MEASURED_PEAKS.json ``hbm_gbs``; ``step_*`` = the whole iteration.
* ``cpu_baseline`` (N = 1): the NumPy oracle (literal restatement of the R loop incl. t(A) copy and the m x n residual
               temporary) on the host cores, 3 iterations -- whose factors are also the parity check of the GPU run.
* ``config.legs`` / ``config.multi_gpu``: compact results of the other legs of BASELINE.json's metric (auto_bcv wall
               time on D3, the 100 000 x 20 000 bcv sweep, FP32 / k = 64 variants, NIPALS, svd_pca); the full records
               go to stderr (one ``DETAIL {...}`` line) so that the stdout line stays short.
* ``--impl reference``: R is not installed and the reference's arithmetic lives in un-vendored packages, so the
  reference arm times the oracle port on all host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS, N_COLS, RANK = 20000, 5000, 16
METRIC = "ALS-NMF iters/s @20kx5k k=16 (FP64)"
UNIT = "iter/s"
WORKLOAD = "ALS-NMF 20000x5000 k=16 FP64, genSimData-shaped D4 (BASELINE configs[3]), fixed iterations"
FP64_PEAK_TFLOPS = 37.1          # DMMA.8x8x4, measured (scripts/microbench/fp64_rates.cu)


def make_workload(m=M_ROWS, n=N_COLS, k=RANK, seed=4):
    """D4: genSimData(n=5, dimx=20000, dimy=5000, clusterHeight=clusterWidth=1000, bgNorm=1,
    biclusterShift=3, rowShift=1, colShift=1), then pseudovalues(); W0 ~ U(0,1) column-normalised,
    Hold0 ~ U(0,1)."""
    from oracle.gen_sim_data import gen_sim_data
    from oracle.helpers import pseudovalues
    side = min(m, n) // 5
    A = gen_sim_data(n=5, cluster_height=side, cluster_width=side, dimx=m, dimy=n, bg_norm=1.0,
                     bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=seed)
    A = np.asfortranarray(pseudovalues(A))
    rs = np.random.default_rng(40)
    W0 = rs.random((m, k))
    W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
    H0 = rs.random((k, n))
    return A, np.asfortranarray(W0), np.asfortranarray(H0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def stream_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per stream-kernel launch from the committed ncu --set full
    capture (mean over the H-step and W-step launches listed in profiles/)."""
    path = os.path.join(ROOT, "profiles", "als_stream_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return float(d["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_d3(dimx=10000, dimy=2000, seed=3):
    """D3: genSimData(n=5, dimx=10000, dimy=2000, clusterHeight=clusterWidth=400, bgNorm=1,
    biclusterShift=3, rowShift=1, colShift=1)."""
    from oracle.gen_sim_data import gen_sim_data
    side = min(dimx, dimy) // 5
    return np.asfortranarray(gen_sim_data(n=5, cluster_height=side, cluster_width=side, dimx=dimx, dimy=dimy,
                                          bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=seed))


def make_d5_device(torch, n, p, device, seed=5, nblocks=5):
    """D5 generated ON THE DEVICE (16 GB at 100 000 x 20 000): the recipe of oracle/gen_sim_data.py --
    N(0,1) background, 5 planted blocks with a block shift N(0,3) and per-column / per-row N(0,1) shifts, rows and
    columns shuffled -- from seeded torch generators.  Returns the (p, n) row-major tensor == column-major n x p."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Yt = torch.empty((p, n), dtype=torch.float64, device=device)
    step = max(1, (1 << 27) // n)
    for j0 in range(0, p, step):
        Yt[j0:j0 + step].normal_(generator=g)
    side_r, side_c = n // (nblocks * 5), p // nblocks                   # 4000 x 4000 at 100000 x 20000
    cg = torch.Generator(device="cpu")
    cg.manual_seed(seed + 1)
    perm_r = torch.randperm(n, generator=cg).to(device)
    perm_c = torch.randperm(p, generator=cg).to(device)
    small = torch.Generator(device="cpu")
    small.manual_seed(seed + 2)
    for b in range(nblocks):
        rows = perm_r[b * side_r:(b + 1) * side_r]
        cols = perm_c[b * side_c:(b + 1) * side_c]
        shift = float(torch.randn(1, generator=small).item()) * 3.0
        colshift = torch.randn(side_c, generator=small, dtype=torch.float64).to(device)
        rowshift = torch.randn(side_r, generator=small, dtype=torch.float64).to(device)
        Yt[cols[:, None], rows[None, :]] += shift + colshift[:, None] + rowshift[None, :]
    return Yt


def run_secondary(mfb, ctx, Ad, m, n):
    """The other kernels of the path on the same resident D4 matrix (wall time of the C-ABI call, second run):
    NIPALS (2 passes over x per inner iteration + one read/write deflation pass per component), svd_pca k = 16,
    Otsu + threshold on a 20000 x 16 factor."""
    import warnings
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):
            t0 = time.perf_counter()
            res = mfb.nipals(Ad, 3, center=False, scale=True, maxiter=40, tol=1e-6, ctx=ctx)
            dt = time.perf_counter() - t0
    # copy, sd scaling (3), |.| sums, 2 per inner iteration, deflation r+w after every component but the last
    passes = 1 + 3 + 1 + sum(int(i) * 2 for i in res["iter"]) + 2 * (len(res["iter"]) - 1)
    peak, _ = peaks()
    out["nipals"] = {"workload": "nipals(x = D4, ncomp = 3, center = FALSE, scale = TRUE, maxiter = 40)",
                     "wall_s": dt, "inner_iterations": [int(i) for i in res["iter"]], "passes_over_x": passes,
                     "achieved_GBps": passes * m * n * 8 / dt / 1e9, "frac_of_hbm": passes * m * n * 8 / dt / 1e9 / peak,
                     "note": "wall time of mfb_nipals incl. the working copy of x and result download"}
    for _ in range(3):
        t0 = time.perf_counter()
        mfb.svd_pca(Ad, 16, ctx=ctx)
        dt = time.perf_counter() - t0
    out["svd_pca"] = {"workload": "svd_pca(D4, k = 16): symmetric Gram (DMMA) + Lanczos + thin products", "wall_s": dt,
                      "gram_flops": 1.0 * m * n * n, "note": "the Gram product alone is m*n*n flop (half the tiles)"}
    Wf = np.asfortranarray(np.abs(np.random.default_rng(1).standard_normal((m, 16))))
    for _ in range(2):
        t0 = time.perf_counter()
        th = mfb.otsuHelper(Wf, ctx=ctx)
        mfb.threshold(Wf, th, 2, ctx=ctx)
        dt = time.perf_counter() - t0
    out["otsu_threshold"] = {"workload": "otsuHelper + threshold on a 20000 x 16 factor (host buffers)", "wall_s": dt}
    return out


def run_k64(mfb, ctx, Ad, m, n, k=64, iters=10):
    """BASELINE configs[3]'s compute-bound variant (k = 64, 25.7 GFLOP per iteration) on the same resident matrix:
    reported against the measured FP64 tensor-pipe peak, not the HBM roofline."""
    rs = np.random.default_rng(41)
    W0 = rs.random((m, k))
    W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
    H0 = rs.random((k, n))
    sess = mfb.AlsSession(Ad, k, np.asfortranarray(W0), np.asfortranarray(H0), Ad.stats()["max"])
    st, _, _ = sess.iterate(3, fixed_iters=True)
    assert st == 0, st
    t0 = time.perf_counter()
    st, _, _ = sess.iterate(iters, fixed_iters=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    kern = [float(v) for v in sess.profile(5)]
    sess.close()
    flops = 4.0 * m * n * k + 2.0 * (m + n) * k * k
    return {"workload": "ALS-NMF 20000x5000 k=64 FP64, fixed iterations", "iters_per_s": iters / dt,
            "fp64_tflops_whole_iteration": flops * iters / dt / 1e12,
            "fp64_tflops_stream_kernels": 4.0 * m * n * k / ((kern[0] + kern[2]) * 1e-3) / 1e12,
            "fp64_peak_tflops_measured": FP64_PEAK_TFLOPS,
            "kernel_ms": {"h_stream": kern[0], "h_solve": kern[1], "w_stream": kern[2], "w_solve": kern[3]}}


def run_snmf(mfb, ctx, Ad, A, m, n, k=RANK, iters=40):
    """snmf (R/bicluster.R:341-441; NMF's snmf/r) on the same resident matrix: the ALS kernels with regularised Gram
    matrices plus NMF's convergence test on every fifth iteration (the next H-step's stream kernel runs ahead for it
    and is not repeated): 2*m*n*8 algorithmic bytes per iteration as for als_nmf."""
    W0 = np.asfortranarray(np.random.default_rng(43).random((m, k)))
    mean = float(Ad.stats()["sum"]) / (float(m) * n)
    dt = beta = None
    # mfBiclust starts at beta = mean(A) and halves it while NMF reports a zero row in H (R/bicluster.R:380-390): time
    # the first beta of that sequence that runs through
    for halvings in range(0, 40, 4):
        beta = mean / 2.0 ** halvings
        sess = mfb.SnmfSession(Ad, k, W0, beta, mean)
        st, _, _, _ = sess.iterate(5, bi_conv=(0, 10 ** 6))      # (no stop: iconv out of reach)
        if st == 0:
            ctx.sync()
            t0 = time.perf_counter()
            st, it, conv, obj = sess.iterate(iters, bi_conv=(0, 10 ** 6))
            ctx.sync()
            if st == 0:
                dt = time.perf_counter() - t0
        sess.close()
        if dt is not None:
            break
    if dt is None:
        return {"unavailable": "zero row in H for every beta tried"}
    peak, _ = peaks()
    gbs = 2.0 * m * n * 8 * iters / dt / 1e9
    return {"workload": "snmf/r 20000x5000 k=16 FP64, eta = mean(A), beta = mean(A) / 2^%d, %d iterations incl. %d convergence tests" % (halvings, iters, iters // 5),
            "iters_per_s": iters / dt, "step_achieved_GBps": gbs, "step_frac_of_hbm": gbs / peak, "objective": obj}


def run_f32(mfb, ctx, A, W0, H0, m, n, k=RANK, iters=40):
    """FP32 leg of BASELINE configs[3]: the same workload with A resident in single precision (half the
    algorithmic bytes: 2*m*n*4 per iteration); roofline fraction against the same HBM peak."""
    Af = mfb.Matrix(ctx, host=A, precision="f32")
    sess = mfb.AlsSession(Af, k, W0, H0, Af.stats()["max"])
    st, _, _ = sess.iterate(5, fixed_iters=True)
    assert st == 0, st
    ctx.sync()
    t0 = time.perf_counter()
    st, _, _ = sess.iterate(iters, fixed_iters=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    kern = [float(v) for v in sess.profile(10)]
    sess.close()
    Af.free()
    peak, _ = peaks()
    gbs = 2.0 * m * n * 4 * iters / dt / 1e9
    return {"workload": "ALS-NMF 20000x5000 k=16, A stored FP32, fixed iterations",
            "iters_per_s": iters / dt, "step_achieved_GBps": gbs, "step_frac_of_hbm": gbs / peak,
            "kernel_ms": {"h_stream": kern[0], "h_solve": kern[1], "w_stream": kern[2], "w_solve": kern[3]}}


def run_auto_bcv(mfb, ctx, rank, world, with_cpu):
    """auto_bcv(Y, ks = 1:20, holdouts = 3) on D3 to convergence, the bcv cells sharded over the ranks: wall time with
    Y resident (``wall_s``) and through the public call with a HOST matrix (upload inside, ``wall_e2e_s``); plus one
    bcv call beside the oracle's."""
    import torch
    import torch.distributed as dist
    Y = make_d3()
    ks = np.arange(1, 21)
    rf, cf = mfb.draw_folds(mfb.HostRng(50), *Y.shape, 3)
    Yd = mfb.Matrix(ctx, host=Y)
    mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx)            # warm-up (workspace pools, module load)
    wrng = mfb.HostRng(49)
    mfb.bcv_batch(Yd, ks, 3, [mfb.draw_folds(wrng, *Y.shape, 3) for _ in range(max(2, 2 * world))], ctx=ctx)
    torch.cuda.synchronize()

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())

    got, t_call = timed(lambda: mfb.bcvGivenKs(Yd, ks, 3, folds=(rf, cf), ctx=ctx))
    best_t, res = 1e9, None
    for _ in range(2):
        res, t = timed(lambda: mfb.auto_bcv(Yd, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx))
        best_t = min(best_t, t)
    Yd.free()
    res_h, t_e2e = timed(lambda: mfb.auto_bcv(Y, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx))
    assert res_h == res
    info = {"workload": "auto_bcv(Y, ks=1:20, holdouts=3) on genSimData-shaped 10000x2000 (D3), FP64",
            "wall_s": best_t, "wall_e2e_s": t_e2e, "best": int(res["best"]), "iterations": int(res["iterations"]),
            "bcv_call_s": t_call, "bcv_argmin_k": int(ks[int(np.argmin(got))]), "cells_per_call": 9, "n_gpus": world}
    if with_cpu and rank == 0:
        from oracle import bcv as OB
        t0 = time.perf_counter()
        ref = OB.bcv_given_ks(Y, ks, rf, cf, 3)
        info["cpu_bcv_call_s"] = time.perf_counter() - t0
        info["cpu_kind"] = ("oracle one-SVD-per-cell restatement (NumPy/LAPACK, %d cores); the reference's literal path "
                            "does one more full SVD per k per cell" % os.cpu_count())
        info["curve_max_rel_err"] = float(np.abs(got / ref - 1.0).max())
        info["cpu_argmin_k"] = int(ks[int(np.argmin(ref))])
    return info


def run_d5(mfb, ctx, rank, world, local_rank, with_cpu, n=100000, p=20000, h=4, kmax=50):
    """BASELINE.json configs[4]: one bcv sweep ks = 1..50 on a genSimData-shaped 100 000 x 20 000 FP64 matrix (16 GB,
    a replica per rank, generated on the device), holdout cells sharded over the ranks; holdouts = 4 -> 16 cells
    (2 per rank on 8 GPUs).  Parity and the CPU baseline use a D5-shaped matrix subsampled to 1/5 per side, one
    cell (the literal bcv at full size is infeasible on the host)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    Yt = make_d5_device(torch, n, p, dev)
    torch.cuda.synchronize()
    Yd = mfb.Matrix(ctx, dev_ptr=Yt.data_ptr(), shape=(n, p))
    ks = np.arange(1, kmax + 1)
    rs = np.random.default_rng(50)
    rf = ((np.arange(1, n + 1) % h) + 1)[rs.permutation(n)].astype(np.int32)
    cf = ((np.arange(1, p + 1) % h) + 1)[rs.permutation(p)].astype(np.int32)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vals = mfb.bcvGivenKs(Yd, ks, h, folds=(rf, cf), ctx=ctx)
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    Yd.free()
    del Yt
    torch.cuda.empty_cache()
    ctx.trim()
    info = {"workload": f"bcv(Y {n}x{p} FP64 genSimData-shaped, ks=1:{kmax}, holdouts={h}), cells sharded over the ranks",
            "n_gpus": world, "cells": h * h, "wall_s": float(tt.item()), "argmin_k": int(ks[int(np.argmin(vals))]),
            "curve_checksum": float(np.sum(vals * np.arange(1, len(vals) + 1)))}
    if with_cpu and rank == 0:
        from oracle import bcv as OB
        ns, ps = n // 5, p // 5
        Ys = make_d5_device(torch, ns, ps, dev, seed=6).cpu().numpy().T        # F-contiguous (ns, ps) view
        rfs = ((np.arange(1, ns + 1) % 3) + 1)[rs.permutation(ns)].astype(np.int32)
        cfs = ((np.arange(1, ps + 1) % 3) + 1)[rs.permutation(ps)].astype(np.int32)
        _, cells = mfb.bcvGivenKs(Ys, ks, 3, folds=(rfs, cfs), ctx=ctx, cell_errors=True)
        t0 = time.perf_counter()
        ref = OB.bcv_cell_errors(Ys, list(ks), rfs == 1, cfs == 1)
        info["cpu_one_cell_s"] = time.perf_counter() - t0
        info["cpu_sample"] = f"one cell of a D5-shaped {ns}x{ps} matrix (1/5 per side), oracle one-SVD-per-cell, {os.cpu_count()} cores"
        info["subsample_cell_max_rel_err"] = float(np.abs(cells[0] / ref - 1.0).max())
    return info


def run_sharded(mfb, ctx, Ad_full, A, W0, H0, rank, world, local_rank, K, warm, k=RANK):
    """Row-sharded als_nmf (mfb_als_begin_sharded) over the ranks: (weak) every rank holds a full D4 as
    its row block, i.e. one (20000 N) x 5000 matrix -- the N > 1 headline; (strong) D4 itself split by rows.
    Device-timed (CUDA events on the library's stream), barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    m, n = A.shape
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    amax = Ad_full.stats()["max"]
    out = {"n_gpus": world, "exchange_doubles_per_iteration": ((n + 127) // 128) * 128 * 16 + 256 + 4,
           "modes": "peer_memory = cudaIpc-mapped exchange blocks, one-shot all-reduce over NVLink fused into the "
                    "invert / H-solve / finish kernels; nccl_callback = NCCL all-reduce (torch.distributed) on the "
                    "library's exchange buffer between the kernels, stream-ordered"}

    def timed(sess, iters):
        st, _, _ = sess.iterate(max(3, warm), fixed_iters=True)
        assert st == 0, st
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st, _, _ = sess.iterate(iters, fixed_iters=True)
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, _ = peaks()
    r0, r1 = mfb.row_range(m, rank, world)
    Ar = mfb.Matrix(ctx, host=np.asfortranarray(A[r0:r1]))
    Wst = np.asfortranarray(W0 / np.sqrt(world))
    for exchange in ("peer", "callback"):
        key = "peer_memory" if exchange == "peer" else "nccl_callback"
        try:
            sess = mfb.ShardedAlsSession(Ad_full, m * world, k, Wst, H0, amax, exchange=exchange)
        except mfb.MfbError as e:          # peer memory cannot be mapped on this box: all ranks agree (see _connect_peers)
            out[key] = {"unavailable": str(e)}
            continue
        # weak: every rank's row block is a whole D4 -> one (m * world) x n matrix
        ms = timed(sess, K)
        _, _, resid, _ = sess.result()
        kern = [float(v) for v in sess.profile(10)] if exchange == "peer" else None
        sess.close()
        gbs = 2.0 * m * world * n * 8 * K / (ms * 1e-3) / 1e9
        res = {"weak": {"workload": "%dx5000 k=16 FP64 (one D4 per GPU as its row block) over %d GPUs" % (m * world, world),
                        "iters_per_s": K / (ms * 1e-3), "ms_per_iter": ms / K, "ms_total": ms, "final_resid": resid,
                        "aggregate_GBps": gbs, "frac_of_aggregate_hbm": gbs / (peak * world)}}
        if kern:
            res["weak"]["kernel_ms"] = dict(zip(("h_stream", "fold_wait_invert", "h_solve", "w_stream",
                                                 "w_solve_wait_finish"), kern))
        # strong: D4 split by rows
        sess = mfb.ShardedAlsSession(Ar, m, k, W0[r0:r1], H0, amax, exchange=exchange)
        ms = timed(sess, K)
        _, _, resid, _ = sess.result()
        sess.close()
        res["strong"] = {"workload": "D4 20000x5000 k=16 FP64 row-sharded over %d GPUs" % world,
                         "iters_per_s": K / (ms * 1e-3), "ms_per_iter": ms / K, "final_resid": resid,
                         "aggregate_GBps": 2.0 * m * n * 8 * K / (ms * 1e-3) / 1e9}
        out[key] = res
    Ar.free()
    return out


def run_sharded_e2e(mfb, ctx, A, W0, H0, rank, world, local_rank, K, k=RANK):
    """The N > 1 headline through the public call with HOST buffers: every rank uploads its row block (a D4) from
    pinned memory, runs K iterations of the (20000 N) x 5000 factorisation and downloads its rows of W and H."""
    import torch
    import torch.distributed as dist
    m, n = A.shape
    pin = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
    pin.numpy()[...] = A.T
    Ah = pin.numpy().T
    Wst = np.asfortranarray(np.tile(W0 / np.sqrt(world), (world, 1)))
    mfb.als_replicate_sharded(Ah, m * world, k, Wst, H0, maxIter=2, fixed_iters=True, ctx=ctx, gather=False)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = mfb.als_replicate_sharded(Ah, m * world, k, Wst, H0, maxIter=K, fixed_iters=True, ctx=ctx, gather=False)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    h2d = (A.nbytes + W0.nbytes + H0.nbytes) / K
    d2h = (r["W"].nbytes + r["H"].nbytes) / K
    return {"value": _r(world * K / dt), "unit": UNIT, "h2d_bytes_per_step": int(round(h2d)), "d2h_bytes_per_step": int(round(d2h)),
            "call": "als_replicate_sharded (%s)" % r["exchange"]}


def run_reference(args, rank, world):
    """Reference arm: the CPU oracle port (literal R loop) on all host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle.als_nmf import als_replicate
    A, W0, H0 = make_workload()
    cores = os.cpu_count()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bounded sample: each timed "step" is one literal iteration of the full-size workload
    steps = min(steps, 4)
    warm = min(warm, 1)
    als_replicate(A, W0, H0, max_iter=max(1, warm), fixed_iters=True, literal=True)
    t0 = time.perf_counter()
    als_replicate(A, W0, H0, max_iter=steps, fixed_iters=True, literal=True)
    dt = time.perf_counter() - t0
    val = steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD,
                      "note": "R absent and NMF::.fcnnls un-vendored: oracle port (NumPy/OpenBLAS, literal R loop)"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{steps} literal ALS iterations of the full 20000x5000 k=16 workload "
                                      f"(after {warm} warm-up); steps capped at 4 to bound the run"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def _r(x, nd=4):
    return None if x is None else float(f"{x:.{nd}g}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bcv", action="store_true", help="skip the auto_bcv wall-time leg (second half of the metric)")
    ap.add_argument("--no-d5", action="store_true", help="skip the 100 000 x 20 000 bcv sweep (BASELINE configs[4])")
    ap.add_argument("--no-secondary", action="store_true", help="skip the NIPALS / svd_pca / Otsu / k=64 / FP32 timings")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: keep the replicate leg as the headline")
    ap.add_argument("--k", type=int, default=RANK, help="rank (default 16 = the headline; 64 = BASELINE configs[3]'s "
                                                        "compute-bound variant, reported against the FP64 pipe)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mfbiclust_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import mfbiclust_b200 as mfb
    ctx = mfb.Context(local_rank)
    K, W = max(1, args.steps), max(3, args.warmup)
    with_cpu = not args.no_cpu_baseline

    RANK_ = args.k
    A, W0, H0 = make_workload(k=RANK_)
    m, n = A.shape
    peak, peak_kind = peaks()
    detail = {}
    # ---- resident-input timing: one D4 per GPU (N = 1: the headline; N > 1: als_nmf's replicate axis) -------------
    Ad = mfb.Matrix(ctx, host=A)
    stats = Ad.stats()
    sess = mfb.AlsSession(Ad, RANK_, W0, H0, stats["max"])
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    st, _, _ = sess.iterate(W, fixed_iters=True)          # warm-up iterations (untimed)
    assert st == 0, f"warm-up failed with status {st}"
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    st, total, _ = sess.iterate(K, fixed_iters=True)      # EXACTLY K timed iterations, 4K kernel launches
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert st == 0 and total == W + K, (st, total)
    _, _, resid, _ = sess.result()
    # per-kernel durations (CUDA events after every launch, on the library's stream) over further iterations
    kern_ms = [float(v) for v in sess.profile(min(K, 20))]
    sess.close()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_rep = float(t.item())
    value_rep = world * K / (ms_rep * 1e-3)

    # ---- N > 1 headline: rows of ONE (20000 N) x 5000 matrix sharded over the ranks ---------------------------------
    sharded = None
    if world > 1 and not args.no_sharded and RANK_ == RANK:
        sharded = run_sharded(mfb, ctx, Ad, A, W0, H0, rank, world, local_rank, K, W)
        detail["als_row_sharded"] = sharded
    clocks = sampler.stop() if rank == 0 else None
    head = None
    if sharded:
        for key in ("peer_memory", "nccl_callback"):
            if "weak" in sharded.get(key, {}):
                head = (key, sharded[key])
                break

    # ---- parity of the timed workload against the oracle + the CPU baseline (N = 1, rank 0) -------------------------
    cpu = parity = None
    if with_cpu and world == 1:
        from oracle.als_nmf import als_replicate
        cores = os.cpu_count()
        nit = 3
        t0 = time.perf_counter()
        ref = als_replicate(A, W0, H0, max_iter=nit, fixed_iters=True, literal=True)
        dt = time.perf_counter() - t0
        cpu = {"value": _r(nit / dt), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nit} oracle iterations of this workload"}
        got = mfb.als_replicate(Ad, RANK_, W0, H0, maxIter=nit, fixed_iters=True, ctx=ctx)
        parity = {"iters": nit,
                  "rel_fro_W": _r(float(np.linalg.norm(got["W"] - ref["W"]) / np.linalg.norm(ref["W"])), 3),
                  "rel_fro_H": _r(float(np.linalg.norm(got["H"] - ref["H"]) / np.linalg.norm(ref["H"])), 3),
                  "zero_pattern_equal": bool(np.array_equal(got["W"] == 0, ref["W"] == 0)
                                             and np.array_equal(got["H"] == 0, ref["H"] == 0))}

    # ---- end to end through the host-buffer API -----------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        if head is not None:
            e2e = run_sharded_e2e(mfb, ctx, A, W0, H0, rank, world, local_rank, K)
        else:
            pin = torch.empty((n, m), dtype=torch.float64, pin_memory=True)   # column-major m x n == row-major n x m
            pin.numpy()[...] = A.T
            Ah = pin.numpy().T                                              # F-contiguous view on pinned memory
            assert Ah.flags["F_CONTIGUOUS"]
            import ctypes as C
            from mfbiclust_b200._cabi import dptr, lib, check
            Wout = np.zeros((m, RANK_), order="F")
            Hout = np.zeros((RANK_, n), order="F")
            residc, itc, convc = C.c_double(0), C.c_int(0), C.c_int(0)
            tr = np.zeros((K, 3))

            def one_call():
                check(lib().mfb_als_run_host(ctx.handle, dptr(Ah), m, n, RANK_, dptr(W0), dptr(H0), K, 1e-4, 1,
                                             dptr(Wout), dptr(Hout), C.byref(residc), C.byref(itc), C.byref(convc),
                                             dptr(tr)))
            one_call()                                  # warm-up call
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            reps = 2
            t0 = time.perf_counter()
            for _ in range(reps):
                one_call()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            td = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
            # per step (= per iteration) byte counts, from the buffers actually copied by the call
            h2d = (A.nbytes + W0.nbytes + H0.nbytes) / K
            d2h = (Wout.nbytes + Hout.nbytes + tr.nbytes) / K
            e2e = {"value": _r(world * K / dt), "unit": UNIT, "h2d_bytes_per_step": int(round(h2d)), "d2h_bytes_per_step": int(round(d2h)),
                   "call": "mfb_als_run_host"}
            # the host-buffer call must return what the resident session computes from the same start
            chk = mfb.als_replicate(Ad, RANK_, W0, H0, maxIter=K, fixed_iters=True, ctx=ctx)
            assert itc.value == K and np.array_equal(Wout, chk["W"]) and np.array_equal(Hout, chk["H"]), \
                "mfb_als_run_host and the resident session disagree"
            e2e["eq_resident"] = True

    if not args.no_secondary and rank == 0:
        sec = run_secondary(mfb, ctx, Ad, m, n)
        if RANK_ == RANK:
            sec["als_k64"] = run_k64(mfb, ctx, Ad, m, n)
            sec["snmf"] = run_snmf(mfb, ctx, Ad, A, m, n)
            sec["als_fp32"] = run_f32(mfb, ctx, A, W0, H0, m, n)
        detail["secondary"] = sec
    Ad.free()
    ctx.trim()

    # ---- auto_bcv wall time (second half of BASELINE.json's metric): D3, cells sharded over the ranks ----------------
    bcv_info = d5_info = None
    if not args.no_bcv:
        bcv_info = run_auto_bcv(mfb, ctx, rank, world, with_cpu and world == 1)
        detail["auto_bcv"] = bcv_info
    if not args.no_d5:
        d5_info = run_d5(mfb, ctx, rank, world, local_rank, with_cpu and world == 1)
        detail["bcv_d5"] = d5_info

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    bytes_per_launch = float(m) * n * 8           # one pass over A per stream launch (2*m*n*s per iteration)
    stream_ms = 0.5 * (kern_ms[0] + kern_ms[2])   # mean launch duration of the dominant (stream) kernel
    achieved = bytes_per_launch / (stream_ms * 1e-3) / 1e9
    legs = {}
    if bcv_info:
        legs.update({"auto_bcv_s": _r(bcv_info["wall_s"]), "auto_bcv_e2e_s": _r(bcv_info["wall_e2e_s"]),
                     "auto_bcv_best": bcv_info["best"], "auto_bcv_iters": bcv_info["iterations"]})
        if "curve_max_rel_err" in bcv_info:
            legs["bcv_curve_err"] = _r(bcv_info["curve_max_rel_err"], 2)
    if d5_info:
        legs["d5_bcv_s"] = _r(d5_info["wall_s"])
        if "subsample_cell_max_rel_err" in d5_info:
            legs["d5_sub_err"] = _r(d5_info["subsample_cell_max_rel_err"], 2)
    sec = detail.get("secondary") or {}
    if "als_fp32" in sec:
        legs.update({"f32_iters_s": _r(sec["als_fp32"]["iters_per_s"]), "f32_frac": _r(sec["als_fp32"]["step_frac_of_hbm"], 3)})
    if "als_k64" in sec:
        legs.update({"k64_iters_s": _r(sec["als_k64"]["iters_per_s"]), "k64_tflops": _r(sec["als_k64"]["fp64_tflops_stream_kernels"], 3)})
    if "iters_per_s" in sec.get("snmf", {}):
        legs["snmf_iters_s"] = _r(sec["snmf"]["iters_per_s"])
    if "nipals" in sec:
        legs.update({"nipals_frac": _r(sec["nipals"]["frac_of_hbm"], 3), "svd_pca_s": _r(sec["svd_pca"]["wall_s"], 3)})

    config = {"workload": WORKLOAD if RANK_ == RANK else WORKLOAD.replace("k=16", f"k={RANK_}"),
              "l2": "input 800 MB/GPU > 126 MB L2: no flush"}
    if head is not None:
        key, res = head
        wk = res["weak"]
        value, ms_step = world * wk["iters_per_s"], wk["ms_per_iter"]
        step_achieved = wk["aggregate_GBps"] / world       # per GPU
        config["parallelism"] = f"rows of one {m * world}x{n} matrix sharded over {world} GPUs, {key} all-reduce"
        config["multi_gpu"] = {"weak_iters_s": _r(wk["iters_per_s"]), "weak_frac": _r(wk["frac_of_aggregate_hbm"], 3),
                               "strong_iters_s": _r(res["strong"]["iters_per_s"]),
                               "one_gpu_iters_s": _r(value_rep / world), "replicates_iters_s": _r(value_rep)}
        other = sharded.get("nccl_callback" if key == "peer_memory" else "peer_memory", {})
        if "weak" in other:
            config["multi_gpu"]["weak_iters_s_nccl"] = _r(other["weak"]["iters_per_s"])
        launches = 5 * K
    else:
        value, ms_step = value_rep, ms_rep / K
        step_achieved = 2 * bytes_per_launch * K / (ms_rep * 1e-3) / 1e9
        config["parallelism"] = "1 GPU" if world == 1 else f"{world} independent replicates (als_nmf reps axis)"
        launches = 4 * K
    if legs:
        config["legs"] = legs
    flops_per_iter = 4.0 * m * n * RANK_ + 2.0 * (m + n) * RANK_ * RANK_
    out = {"metric": METRIC if RANK_ == RANK else METRIC.replace("k=16", f"k={RANK_}"), "value": _r(value, 6), "unit": UNIT,
           "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": _r(ms_step, 5), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "roofline": {"bound": "hbm", "achieved": _r(achieved, 5), "peak": peak, "unit": "GB/s",
                        "frac": _r(achieved / peak, 4), "traffic": stream_traffic_bytes(), "peak_src": peak_kind,
                        "kernel": "als_stream_kernel",
                        "kernel_ms": {"h_stream": _r(kern_ms[0]), "h_solve": _r(kern_ms[1]), "w_stream": _r(kern_ms[2]),
                                      "w_solve": _r(kern_ms[3])},
                        "step_achieved": _r(step_achieved, 5), "step_frac": _r(step_achieved / peak, 4),
                        "fp64_tflops_step": _r(flops_per_iter / (ms_step * 1e-3) / 1e12 * (world if head else 1), 4)},
           "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
           "final_resid": _r(resid, 8)}
    detail["line"] = dict(out)
    sys.stderr.write("DETAIL " + json.dumps(detail) + "\n")
    sys.stderr.flush()
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"bench_detail_n{world}.json"), "w") as fh:
            json.dump(detail, fh, indent=1)
    except OSError:
        pass
    line = json.dumps(out, separators=(",", ":"))
    if len(line) > 1400:                           # the driver keeps a 1500-character tail: drop what the DETAIL line repeats
        out.pop("final_resid", None)
        out["roofline"].pop("fp64_tflops_step", None)
        (out.get("clocks") or {}).pop("samples", None)
        line = json.dumps(out, separators=(",", ":"))
    print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
