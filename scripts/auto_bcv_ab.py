"""auto_bcv on D3 under torchrun: wall time with / without the provable first batch and for several stream counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import mfbiclust_b200 as mfb
from bench import make_d3
Y = make_d3(); ks = np.arange(1, 21); ctx = mfb.Context(lr)
for pre, streams in (("", 3), ("1", 3), ("", 3), ("", 2), ("", 1), ("1", 2), ("1", 1)):
    os.environ["MFBICLUST_BCV_STREAMS"] = str(streams)
    if pre: os.environ["MFBICLUST_BCV_NO_PREBATCH"] = "1"
    else: os.environ.pop("MFBICLUST_BCV_NO_PREBATCH", None)
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); r = mfb.auto_bcv(Y, ks, holdouts=3, rng=mfb.HostRng(51), ctx=ctx); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0: print("world", world, "no_prebatch" if pre else "prebatch", "streams", streams, round(dt, 3), r["best"], r["iterations"], flush=True)
if world > 1: dist.destroy_process_group()
