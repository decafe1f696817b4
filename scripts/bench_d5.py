#!/usr/bin/env python
"""BASELINE.json configs[4]: bcv sweep ks = 1..50 on a genSimData-shaped 100 000 x 20 000 FP64 matrix (16 GB),
holdout cells sharded over the ranks.  The matrix is generated ON THE DEVICE (torch is plumbing here: a seeded
device generator + the same recipe as oracle/gen_sim_data.py: N(0,1) background, 5 planted 4000 x 4000 blocks with a
block shift N(0,3), per-column N(0,1) and per-row N(0,1) shifts, rows/columns shuffled), wrapped with
mfb_matrix_wrap_dev, and every rank holds a replica.

    python scripts/bench_d5.py [--rows 100000 --cols 20000 --holdouts 3]        (1 GPU)
    torchrun --nproc-per-node 8 ... scripts/bench_d5.py                           (8 GPUs)

Prints one JSON line: wall time of one bcv call, the per-k curve checksum and argmin (identical for every number of
ranks: each cell is computed by exactly one rank with the same kernels)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_d5_device(torch, n, p, device, seed=5, nblocks=5):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Yt = torch.empty((p, n), dtype=torch.float64, device=device)      # row-major (p, n) == column-major n x p
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--cols", type=int, default=20000)
    ap.add_argument("--holdouts", type=int, default=3)
    ap.add_argument("--kmax", type=int, default=50)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mfbiclust_b200 as mfb
    ctx = mfb.Context(local)
    n, p, h = args.rows, args.cols, args.holdouts
    t0 = time.perf_counter()
    Yt = make_d5_device(torch, n, p, dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    Yd = mfb.Matrix(ctx, dev_ptr=Yt.data_ptr(), shape=(n, p))
    ks = np.arange(1, args.kmax + 1)
    rs = np.random.default_rng(50)
    rf = ((np.arange(1, n + 1) % h) + 1)[rs.permutation(n)].astype(np.int32)
    cf = ((np.arange(1, p + 1) % h) + 1)[rs.permutation(p)].astype(np.int32)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vals, cells = mfb.bcvGivenKs(Yd, ks, h, folds=(rf, cf), ctx=ctx, cell_errors=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"bcv(Y {n}x{p} FP64 genSimData-shaped, ks=1:{args.kmax}, holdouts={h})",
                          "n_gpus": world, "cells": h * h, "wall_s": float(tt.item()), "generate_s": t_gen,
                          "argmin_k": int(ks[int(np.argmin(vals))]),
                          "curve_head": [float(v) for v in vals[:6]],
                          "curve_checksum": float(np.sum(vals * np.arange(1, len(vals) + 1)))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
