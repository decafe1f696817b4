"""Development aid (torchrun, one rank per GPU, or plain python for world 1): when the kernels of one row-sharded ALS
iteration start (block 0 scheduled / dependency satisfied), relative to the H-step stream kernel, per rank."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
import torch.distributed as dist
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import mfbiclust_b200 as mfb
from bench import make_workload
A, W0, H0 = make_workload()
ctx = mfb.Context(lr)
Ad = mfb.Matrix(ctx, host=A)
amax = Ad.stats()["max"]
L = mfb._cabi.lib()
L.mfb_als_timeline.argtypes = [C.c_void_p, C.c_int]
names = ["Hstream", "close", "fold", "Hsolve", "Wstream", "Wsolve"]
for mode in (["one-device"] if world == 1 else []) + ["sharded"]:
    if mode == "one-device":
        sess = mfb.AlsSession(Ad, 16, W0, H0, amax)
    else:
        Wst = np.asfortranarray(W0 / np.sqrt(world))
        sess = mfb.ShardedAlsSession(Ad, A.shape[0] * world, 16, Wst, H0, amax, exchange="peer" if world > 1 else "auto")
    sess.iterate(10, fixed_iters=True)
    L.mfb_als_timeline(None, 1)
    sess.iterate(10, fixed_iters=True)
    ctx.sync()
    st = (C.c_ulonglong * 32)()
    L.mfb_als_timeline(st, 0)
    t = [int(v) for v in st]
    t0 = t[1]
    line = "rank %d %s period %.1f us |" % (rank, mode, (t[1] - t[25]) / 1e3)
    for i, nm in enumerate(names):
        if t[2 * i + 1] >= t0:
            line += " %s sched %+.1f ready %+.1f |" % (nm, (t[2 * i] - t0) / 1e3, (t[2 * i + 1] - t0) / 1e3)
    print(line, flush=True)
    if world > 1:
        dist.barrier()
    sess.close()
if world > 1:
    dist.destroy_process_group()
