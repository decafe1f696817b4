"""CPU-only checks: the C-ABI library loads and exports every symbol include/mfbiclust.h declares, the
host-side mirror of the reference's pre-steps, and the multi-rank cell sharding over gloo (world size 2)."""
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import mfbiclust_b200._cabi as cabi
    L = cabi.lib()                       # raises if the .so is missing or a prototype is not exported
    hdr = open(os.path.join(ROOT, "include", "mfbiclust.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/mfbiclust.h but not exported"
    assert declared == set(cabi.PROTOTYPES), declared ^ set(cabi.PROTOTYPES)
    assert L.mfb_version() >= 100


def test_ctypes_prototypes_match_the_header_arity():
    """The Python twin of the R shim must pass exactly the arguments include/mfbiclust.h declares (an arity drift is
    silent with ctypes): parameter counts of every declaration against _cabi.PROTOTYPES."""
    import mfbiclust_b200._cabi as cabi
    hdr = open(os.path.join(ROOT, "include", "mfbiclust.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"^\s*typedef[^;]*;", "", hdr, flags=re.M)
    seen = 0
    for m in re.finditer(r"\b(mfb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        depth, n = 0, (0 if params in ("", "void") else 1)
        for ch in params:
            depth += ch == "("
            depth -= ch == ")"
            n += (ch == "," and depth == 0)
        assert name in cabi.PROTOTYPES, name
        assert n == len(cabi.PROTOTYPES[name][1]), (name, n, len(cabi.PROTOTYPES[name][1]))
        seen += 1
    assert seen == len(cabi.PROTOTYPES), (seen, len(cabi.PROTOTYPES))


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import mfbiclust_b200 as mfb
    with pytest.raises(mfb.MfbError, match="no CUDA device|no CPU fallback"):
        mfb.Context(0)
    with pytest.raises(mfb.MfbError):
        mfb.otsuHelper(np.arange(12.0).reshape(4, 3))


def test_klimiter_and_cell_ranges():
    import mfbiclust_b200 as mfb
    with pytest.warns(UserWarning):
        assert list(mfb.kLimiter(3, 18, np.arange(1, 18))) == list(range(1, 12))      # R/bcv.R:241-251
    for ncell in (9, 100, 4):
        for world in (1, 2, 3, 8):
            got = [mfb.cell_range(ncell, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == ncell
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_draw_folds_balanced():
    import mfbiclust_b200 as mfb
    rf, cf = mfb.draw_folds(mfb.HostRng(1), 1099, 18, 3)
    assert sorted(np.bincount(rf)[1:]) == [366, 366, 367] and sorted(np.bincount(cf)[1:]) == [6, 6, 6]


def test_validate_km_clean_pseudovalues():
    import mfbiclust_b200 as mfb
    from oracle import helpers as Hh
    m = np.arange(12.0).reshape(4, 3) - 3
    with pytest.warns(UserWarning):
        assert mfb.validateKM(7, m, "als-nmf") == 3
    assert mfb.validateKM(7, m, "svd-pca") == 7          # R/helperFunctions.R:402 quirk
    assert np.array_equal(mfb.pseudovalues(m), Hh.pseudovalues(m))
    x = m.copy()
    x[1, :] = 0
    x[:, 2] = np.nan
    got, idx = mfb.clean(x, 0)
    ref, gr, gc = Hh.clean(x, 0)
    assert np.array_equal(idx[0], gr) and np.array_equal(idx[1], gc) and got.shape == ref.shape


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    import mfbiclust_b200 as mfb
    from oracle import bcv as B
    z = np.load(os.path.join({root!r}, "tests", "golden", "yeast_benchmark.npz"))
    Y = np.asfortranarray(z["y06"])
    h = 3
    ks = B.k_limiter(h, min(Y.shape), np.arange(1, 7))
    rs = np.random.default_rng(5)
    rf = B.make_folds(Y.shape[0], h, rs.permutation(Y.shape[0]))
    cf = B.make_folds(Y.shape[1], h, rs.permutation(Y.shape[1]))
    calls = []
    def compute(c0, c1):      # stands in for mfb_bcv_cells (no GPU here): same cell numbering, cell = (x-1)*h + (y-1)
        calls.append((c0, c1))
        return np.stack([B.bcv_cell_errors(Y, list(ks), rf == (c // h + 1), cf == (c % h + 1)) for c in range(c0, c1)])
    err = mfb.shard_cells(h * h, len(ks), compute)
    full = np.stack([B.bcv_cell_errors(Y, list(ks), rf == (c // h + 1), cf == (c % h + 1)) for c in range(h * h)])
    assert calls == [mfb.cell_range(h * h, dist.get_rank(), 2)], calls
    assert np.array_equal(err, full), np.abs(err - full).max()
    assert np.array_equal(err.sum(axis=0), full.sum(axis=0))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", sys.argv[1], "ok")
""")


def test_bcv_cell_sharding_gloo_world2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


SHARD_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    rank = int(sys.argv[1])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
    import mfbiclust_b200 as mfb
    from mfbiclust_b200.backends import _allreduce_host
    m, k = 101, 3
    W = np.random.default_rng(1).random((m, k))
    r0, r1 = mfb.row_range(m, rank, 2)
    assert (r0, r1) == ((0, 50) if rank == 0 else (50, 101))
    full = mfb.gather_rows(np.asfortranarray(W[r0:r1]), m)
    assert np.array_equal(full, W)                      # exact: one non-zero contributor per entry
    lo, = _allreduce_host([W[r0:r1].min()], "MIN", 0, None)
    hi, = _allreduce_host([W[r0:r1].max()], "MAX", 0, None)
    assert lo == W.min() and hi == W.max()
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_als_row_sharding_host_side_gloo_world2(tmp_path):
    """Host logic of the row-sharded als_nmf (mfb_als_begin_sharded): row blocks, exact gather of W, global max/min."""
    import socket
    import mfbiclust_b200 as mfb
    for m in (1, 7, 128, 20000):
        for world in (1, 2, 3, 8):
            blocks = [mfb.row_range(m, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_auto_bcv_batch_sizing_matches_the_stopping_rule():
    # R/bcv.R:56-87: with one k winning every time the rule fires at iteration 22 on an 11-candidate grid
    # (vignettes/my-vignette.html:408-413: counts = 22); the batch sizing must predict exactly that
    from mfbiclust_b200.backends import _iters_until_stable
    keys = list(range(1, 12))
    distr = {k: 1 for k in keys}
    old = dict(distr)
    done = 0
    while True:
        nb = _iters_until_stable(distr, old, 1e-4, 11 if done else None, keys, 8)
        assert 1 <= nb <= 8
        conv = False
        for _ in range(nb):
            distr[11] += 1
            sd, so = sum(distr.values()), sum(old.values())
            conv = all((distr[k] / sd - old[k] / so) ** 2 < 1e-4 for k in keys)
            old = dict(distr)
            done += 1
            if conv:
                break
        if conv:
            break
    assert done == 22 and distr[11] - 1 == 22
    # the last batch must not overshoot: after 19 identical winners exactly 3 more iterations are needed
    d = {k: 1 for k in keys}
    o = dict(d)
    d[11] += 19
    o[11] += 18
    assert _iters_until_stable(d, o, 1e-4, 11, keys, 8) == 3


def test_auto_bcv_earliest_stop_is_a_lower_bound_on_every_path():
    """The first batch of auto_bcv covers every iteration the stopping rule (R/bcv.R:69-84) cannot do without: on no
    sequence of winners may the rule fire before ``_earliest_stop`` iterations, and on the leader-wins-all path it
    fires exactly there (vignette: 22 on 11 candidates)."""
    from mfbiclust_b200.backends import _earliest_stop
    rs = np.random.default_rng(0)
    for nkeys in (2, 5, 11, 20, 50):
        keys = list(range(1, nkeys + 1))
        for trial in range(40):
            distr = {k: 1 for k in keys}
            old = dict(distr)
            # a random history first (the bound must hold from any state)
            for _ in range(int(rs.integers(0, 12))):
                old = dict(distr)
                distr[int(rs.choice(keys))] += 1
            L = _earliest_stop(distr, 1e-4, 200)
            favourite = int(rs.choice(keys))
            pfav = rs.random()
            d, o = dict(distr), dict(distr)
            for t in range(1, L):
                w = favourite if rs.random() < pfav else int(rs.choice(keys))
                d[w] += 1
                sd, so = sum(d.values()), sum(o.values())
                assert not all((d[k] / sd - o[k] / so) ** 2 < 1e-4 for k in keys), (nkeys, trial, t, L)
                o = dict(d)
        # leader wins all: fires exactly at L
        distr = {k: 1 for k in keys}
        L = _earliest_stop(distr, 1e-4, 500)
        d, o = dict(distr), dict(distr)
        fired = None
        for t in range(1, L + 1):
            d[keys[-1]] += 1
            sd, so = sum(d.values()), sum(o.values())
            if all((d[k] / sd - o[k] / so) ** 2 < 1e-4 for k in keys):
                fired = t
                break
            o = dict(d)
        assert fired == L, (nkeys, fired, L)
    assert _earliest_stop({k: 1 for k in range(11)}, 1e-4, 100) == 22
    assert _earliest_stop({k: 1 for k in range(20)}, 1e-4, 100) == 25


def test_auto_bcv_batching_replays_the_sequential_loop(yeast, monkeypatch):
    """auto_bcv's host loop (provable first batch + speculative batches) against the oracle's plain sequential loop
    (R/bcv.R:56-87) with the SAME bcv values: the device call is replaced by the oracle's bcv, so only the batching,
    the fold-draw order and the replay of the stopping rule are under test -- on the CPU."""
    import mfbiclust_b200 as mfb
    import mfbiclust_b200.backends as BK
    from mfbiclust_b200 import _cabi
    from oracle import bcv as OB
    Y = yeast[0]
    ks = list(range(1, 12))
    batches = []

    def fake_batch(Yd, keys, holdouts, folds_list, ctx=None, group=None, streams=None):
        batches.append(len(folds_list))
        return np.stack([OB.bcv_given_ks(Y, np.asarray(keys), rf, cf, holdouts) for rf, cf in folds_list])
    monkeypatch.setattr(BK, "bcv_batch", fake_batch)
    Yd = _cabi.Matrix.__new__(_cabi.Matrix)             # a resident-matrix handle that is never dereferenced
    Yd.shape, Yd._h, Yd.ctx, Yd.precision = Y.shape, None, None, "f64"
    ref_rng = mfb.HostRng(7)
    best, counts, iters = OB.auto_bcv(Y, np.asarray(ks), 3, lambda: mfb.draw_folds(ref_rng, *Y.shape, 3))
    for B in (1, 4):
        batches.clear()
        got = mfb.auto_bcv(Yd, ks, holdouts=3, rng=mfb.HostRng(7), ctx=object(), speculate=B)
        assert got["best"] == best and got["iterations"] == iters
        assert [got["counts"][k] for k in ks] == [int(c) for c in counts]
        assert batches[0] == 22                       # 11 candidates: the rule cannot fire before iteration 22
        assert sum(batches) >= iters and all(b <= max(B, 1) for b in batches[1:])


def test_filter_biclust_host_loop_replays_the_oracle(monkeypatch):
    """filter.biclust above the C ABI (R/helperFunctions.R:150-206): with the two device calls stubbed by the oracle's
    set arithmetic the greedy loop must make the oracle's choices -- ties in size go to the first bicluster, an empty
    bicluster (overlap NaN) and a whole-matrix one are never chosen, `max` stops the loop."""
    import mfbiclust_b200.backends as B
    from oracle import helpers as H

    def fake_overlap_sizes(rowx, bicx, ctx=None):
        rows, cols = H.bicluster_matrix2list(rowx, bicx)
        return H.sizes(rows, cols), H.overlap(rows, cols)

    monkeypatch.setattr(B, "overlap_sizes", fake_overlap_sizes)
    monkeypatch.setattr(B, "union_biclust", lambda r, c, choose=None, ctx=None: H.union_biclust(r, c, choose))
    rs = np.random.default_rng(11)
    for trial in range(40):
        m, n, k = int(rs.integers(4, 60)), int(rs.integers(4, 40)), int(rs.integers(0, 9))
        rowx, bicx = rs.random((m, k)) < 0.4, rs.random((k, n)) < 0.4
        if k >= 3 and trial % 3 == 0:
            rowx[:, 0] = False
            rowx[:, 1] = rowx[:, 2]
            bicx[1, :] = bicx[2, :]
        mx = None if trial % 4 else 2
        lim = [0.25, 0.0, 0.5, 1.0][trial % 4]
        got = B.filter_biclust(rowx, bicx, max=mx, overlap=lim)
        ref = H.filter_biclust(rowx, bicx, max=mx, overlap_limit=lim)
        assert np.array_equal(got["chosen"], ref["chosen"]), trial
        assert np.array_equal(got["biclustered"], ref["biclustered"]), trial


def test_snmf_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import mfbiclust_b200 as mfb
    with pytest.raises(mfb.MfbError):
        mfb.snmf(np.arange(1.0, 37.0).reshape(6, 6), 2, verbose=False)


def test_snmf_wrapper_search_mirrors_the_reference(monkeypatch):
    """R/bicluster.R:357-431 with the device run stubbed: beta is halved while NMF gives up after too many restarts
    (:380-390), eta while a system is singular (:399-410), every attempt restarts from the same RNG state (:366), the
    defaults are mean(A) (:351-352) and the validation message is the reference's (:353-355)."""
    from mfbiclust_b200 import _cabi, backends as B

    class FakeMatrix:
        def __init__(self, ctx, host=None, **_):
            self.shape = host.shape

        def free(self):
            pass
    monkeypatch.setattr(_cabi, "Matrix", FakeMatrix)
    monkeypatch.setattr(_cabi, "default_context", lambda *a, **k: object())
    calls = []

    class Rng:
        def __init__(self):
            self.state = 7

        def get_state(self):
            return self.state

        def set_state(self, s):
            self.state = s

        def runif(self, n):
            self.state += 1
            return np.zeros(n)

    def fake_run(Ad, k, W0=None, beta=0.01, eta=-1.0, rng=None, ctx=None, **kw):
        calls.append((beta, eta, rng.state, kw))
        rng.runif(3)                                   # the run consumes draws; the next attempt must not see that
        if len(calls) <= 2:
            return dict(W=None, H=None, iters=20, converged=False, objective=0.0, restarts=10, too_many_restarts=True,
                        beta=beta, eta=eta)
        if len(calls) == 3:
            raise _cabi.MfbError(_cabi.MFB_SINGULAR, "system is computationally singular")
        return dict(W=np.ones((4, 2)), H=np.ones((2, 3)), iters=55, converged=True, objective=1.5, restarts=0,
                    too_many_restarts=False, beta=beta, eta=eta)
    monkeypatch.setattr(B, "nmf_snmf", fake_run)
    A = np.arange(1.0, 13.0).reshape(4, 3)
    fit = B.snmf(A, 2, verbose=False, rng=Rng(), maxIter=77, unknown_option=1)
    mean = A.mean()
    assert [c[0] for c in calls] == [mean, mean / 2, mean / 4, mean / 4]
    assert [c[1] for c in calls] == [mean, mean, mean, mean / 2]
    assert all(c[2] == 7 for c in calls) and all(c[3] == {"maxIter": 77} for c in calls)
    assert fit.method == "snmf" and fit.attempts == 4 and fit.iters == 55
    assert fit.parameters == dict(beta=mean / 4, eta=mean / 2)
    with pytest.raises(ValueError, match="eta must be >= 0 and beta must be > 0"):
        B.snmf(A, 2, verbose=False, beta=0.0)


def test_built_library_uses_the_blackwell_paths():
    """The hot kernels are compiled for sm_100a with the hardware paths DESIGN.md claims (cuobjdump on the in-tree
    objects, no GPU needed): tcgen05.mma + tensor memory + TMA in the FP32 stream kernels, FP64 tensor cores + TMA +
    mbarriers in the FP64 ones, TMA bulk loads in the NIPALS stream kernel, and no legacy mma.sync tiles."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    obj = os.path.join(ROOT, "mfbiclust_b200", "lib", "als.o")
    if not os.path.exists(obj):
        pytest.skip("objects not built in-tree")
    head = subprocess.run(["cuobjdump", "-lelf", obj], capture_output=True, text=True).stdout
    assert "sm_100a" in head, head
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = set()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur].add(m.group(1))
    tf32 = [ops for k, ops in per.items() if "als_stream_tf32_kernel" in k]
    f64 = [ops for k, ops in per.items() if "als_stream_kernel" in k]
    assert tf32 and all({"UTCHMMA", "LDTM", "STTM", "UTMALDG", "SYNCS"} <= ops for ops in tf32)
    assert f64 and all({"DMMA", "UTMALDG", "SYNCS"} <= ops for ops in f64)
    assert not any({"HMMA", "IMMA"} & ops for ops in per.values())
    nip = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mfbiclust_b200", "lib", "nipals.o")],
                         capture_output=True, text=True).stdout
    assert "UTMALDG" in nip
