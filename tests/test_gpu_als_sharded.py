"""GPU parity tests for the row-sharded ALS-NMF path (mfb_als_begin_sharded): rows of A and
W split over the ranks, [W'A | W'W] and the iteration scalars all-reduced once per iteration.  World size 1 runs in
process; world size 2 runs as two gloo processes that share cuda:0 (the collective is the host-staged branch of
HostAllReduce; the NCCL branch is exercised by bench.py --gpus N and scripts/bench_sharded.py on a multi-GPU box)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle.als_nmf import als_replicate as oracle_replicate
from oracle.gen_sim_data import gen_sim_data
from oracle.helpers import pseudovalues

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_fro(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_init(m, n, k, seed):
    rs = np.random.default_rng(seed)
    W0 = rs.random((m, k))
    W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
    return np.asfortranarray(W0), np.asfortranarray(rs.random((k, n)))


@pytest.mark.parametrize("exchange", ["callback", "peer"])
@pytest.mark.parametrize("m,n,k", [(2000, 500, 16), (640, 1030, 24), (600, 400, 64), (37, 23, 3)])
def test_sharded_world1_matches_oracle_and_the_one_device_path(m, n, k, exchange):
    import mfbiclust_b200 as mfb
    A = pseudovalues(gen_sim_data(n=4, cluster_height=min(m, n) // 5, cluster_width=min(m, n) // 5, dimx=m, dimy=n,
                                  bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    W0, H0 = make_init(m, n, k, 43)
    got = mfb.als_replicate_sharded(A, m, k, W0, H0, maxIter=10, fixed_iters=True, exchange=exchange)
    assert got["exchange"] == exchange
    one = mfb.als_replicate(A, k, W0, H0, maxIter=10, fixed_iters=True)
    tr = []
    ref = oracle_replicate(A, W0, H0, max_iter=10, fixed_iters=True, trace=tr)
    assert got["iters"] == 10 and got["allreduce_calls"] == (1 + 2 * 10 if exchange == "callback" else 1)
    assert rel_fro(got["W"], ref["W"]) < 1e-8 and rel_fro(got["H"], ref["H"]) < 1e-8        # north-star FP64 bar
    assert np.array_equal(got["W"] == 0, ref["W"] == 0) and np.array_equal(got["H"] == 0, ref["H"] == 0)
    assert np.allclose(got["trace"], np.array([t[1:] for t in tr]), rtol=1e-9, atol=1e-12)
    assert rel_fro(got["W"], one["W"]) < 1e-11 and rel_fro(got["H"], one["H"]) < 1e-11


WORKER = textwrap.dedent("""
    import os, sys, warnings
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    rank, world, exchange = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import mfbiclust_b200 as mfb
    from oracle.als_nmf import als_replicate as oracle_replicate
    from oracle.gen_sim_data import gen_sim_data
    from oracle.helpers import pseudovalues

    def rel(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    def init(m, n, k, seed):
        rs = np.random.default_rng(seed)
        W0 = rs.random((m, k)); W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
        return np.asfortranarray(W0), np.asfortranarray(rs.random((k, n)))

    # (1) fixed iterations, ragged row blocks (m not a multiple of anything), three ranks' worth of shapes
    for (m, n, k) in [(2001, 500, 16), (643, 1030, 24), (601, 400, 40)]:
        A = pseudovalues(gen_sim_data(n=4, cluster_height=min(m, n) // 5, cluster_width=min(m, n) // 5, dimx=m, dimy=n,
                                      bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
        W0, H0 = init(m, n, k, 43)
        r0, r1 = mfb.row_range(m, rank, world)
        got = mfb.als_replicate_sharded(np.asfortranarray(A[r0:r1]), m, k, W0, H0, maxIter=10, fixed_iters=True,
                                        exchange=exchange)
        assert got["exchange"] == exchange
        tr = []
        ref = oracle_replicate(A, W0, H0, max_iter=10, fixed_iters=True, trace=tr)
        eW, eH = rel(got["W"], ref["W"]), rel(got["H"], ref["H"])
        assert eW < 1e-8 and eH < 1e-8, (m, n, k, eW, eH)
        assert np.array_equal(got["W"] == 0, ref["W"] == 0) and np.array_equal(got["H"] == 0, ref["H"] == 0)
        assert np.allclose(got["trace"], np.array([t[1:] for t in tr]), rtol=1e-9, atol=1e-12)
        assert got["rows"] == (r0, r1) and got["allreduce_calls"] == (21 if exchange == "callback" else 1)
        print("rank", rank, "fixed", (m, n, k), eW, eH, flush=True)

    # (2) convergence break: same iteration count and factors as the oracle, decided identically on every rank
    z = np.load(os.path.join({root!r}, "tests", "golden", "yeast_benchmark.npz"))
    A = pseudovalues(np.asfortranarray(z["y01"]))
    m, n = A.shape
    W0, H0 = init(m, n, 3, 7)
    r0, r1 = mfb.row_range(m, rank, world)
    got = mfb.als_replicate_sharded(np.asfortranarray(A[r0:r1]), m, 3, W0, H0, maxIter=100, eps_conv=1e-4,
                                    exchange=exchange)
    ref = oracle_replicate(A, W0, H0, max_iter=100, eps_conv=1e-4)
    assert got["converged"] and ref["status"] == "converged" and got["iters"] == ref["iters"], (got["iters"], ref["iters"])
    assert rel(got["W"], ref["W"]) < 1e-8 and rel(got["H"], ref["H"]) < 1e-8 and abs(got["obj"] - ref["obj"]) < 1e-10
    print("rank", rank, "converged after", got["iters"], flush=True)

    # (3) restart protocol on the rank-deficient simdata matrix: every rank sees the same failures and redraws the
    #     same W from its replicated generator; gives up with R's warning after the 9th restart
    s = np.load(os.path.join({root!r}, "tests", "golden", "simdata.npz"))
    A = pseudovalues(np.asfortranarray(s["fiveSquareRowShift"]))
    W0, H0 = init(80, 80, 5, 41)
    r0, r1 = mfb.row_range(80, rank, world)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        got = mfb.als_replicate_sharded(np.asfortranarray(A[r0:r1]), 80, 5, W0, H0, maxIter=10, fixed_iters=True,
                                        rng=mfb.HostRng(5), exchange=exchange)
    assert any("Factorization failed too many times" in str(w.message) for w in wlist)
    assert np.isfinite(got["W"]).all() and got["W"].shape == (80, 5)

    # (4) a negative entry on ONE rank is an error on ALL ranks (no rank is left waiting in a collective)
    A = np.abs(np.random.default_rng(3).standard_normal((64, 48))) + 0.1
    A[60, 5] = -1.0                        # last rank's block
    r0, r1 = mfb.row_range(64, rank, world)
    W0, H0 = init(64, 48, 4, 1)
    try:
        mfb.als_replicate_sharded(np.asfortranarray(A[r0:r1]), 64, 4, W0, H0, maxIter=3, fixed_iters=True,
                                  exchange=exchange)
        raise SystemExit("negative entry not rejected")
    except ValueError as e:
        assert "Negative values are not allowed" in str(e)

    # (5) als_nmf_sharded: replicates + winner identical on every rank, full W / H returned
    m, n, k = 400, 300, 6
    A = pseudovalues(gen_sim_data(n=3, cluster_height=60, cluster_width=60, dimx=m, dimy=n, bg_norm=1.0,
                                  bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=9))
    r0, r1 = mfb.row_range(m, rank, world)
    inits = [init(m, n, k, 100 + r) for r in range(3)]
    fit = mfb.als_nmf_sharded(np.asfortranarray(A[r0:r1]), m, k, reps=3, maxIter=30, inits=inits, exchange=exchange)
    objs = [oracle_replicate(A, W0, H0, max_iter=30, eps_conv=1e-4)["obj"] for (W0, H0) in inits]
    assert fit.best == int(np.argmin(objs)) and fit.fit.W.shape == (m, k) and fit.fit.H.shape == (k, n)
    assert np.allclose([s_["obj"] for s_ in fit.solutions], objs, rtol=1e-9)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


@pytest.mark.parametrize("world,exchange", [(2, "callback"), (3, "callback"), (2, "peer"), (3, "peer")])
def test_sharded_gloo_processes_on_one_gpu(tmp_path, world, exchange):
    """exchange = "callback": the host's collective between the kernels; "peer": cudaIpc-mapped exchange blocks, the
    one-shot all-reduce fused into the invert / solve / finish kernels (here the ranks share one GPU; over NVLink on a
    multi-GPU box, bench.py --gpus N)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    env = dict(os.environ, MFBICLUST_DEVICE="0")
    env.pop("LOCAL_RANK", None)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), exchange], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-4000:]
        assert f"rank {r} ok" in o


@pytest.mark.parametrize("exchange", ["callback", "peer"])
def test_sharded_fp32_storage_matches_the_one_device_fp32_path(exchange):
    """The row-sharded session on a single-precision resident block (mfb_matrix_upload_f32): same kernels, same
    exchange; compared with the one-device FP32-storage path (1e-11) and the FP64 oracle (north star: 1e-4)."""
    import mfbiclust_b200 as mfb
    m, n, k = 1500, 700, 10
    A = pseudovalues(gen_sim_data(n=4, cluster_height=140, cluster_width=140, dimx=m, dimy=n, bg_norm=1.0,
                                  bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    W0, H0 = make_init(m, n, k, 43)
    ctx = mfb.default_context()
    Af = mfb.Matrix(ctx, host=A, precision="f32")
    try:
        got = mfb.als_replicate_sharded(Af, m, k, W0, H0, maxIter=10, fixed_iters=True, exchange=exchange)
        one = mfb.als_replicate(Af, k, W0, H0, maxIter=10, fixed_iters=True)
    finally:
        Af.free()
    ref = oracle_replicate(A, W0, H0, max_iter=10, fixed_iters=True)
    assert rel_fro(got["W"], one["W"]) < 1e-11 and rel_fro(got["H"], one["H"]) < 1e-11
    assert rel_fro(got["W"], ref["W"]) < 1e-4 and rel_fro(got["H"], ref["H"]) < 1e-4
