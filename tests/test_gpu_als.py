"""GPU parity tests for the ALS-NMF path (C ABI -> CUDA) against the CPU oracle."""
import numpy as np
import pytest

from oracle import otsu as O
from oracle.als_nmf import als_replicate as oracle_replicate
from oracle.fcnnls import fcnnls
from oracle.helpers import pseudovalues
from oracle.gen_sim_data import gen_sim_data

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_init(m, n, k, seed):
    rs = np.random.default_rng(seed)
    W0 = rs.random((m, k))
    W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
    return np.asfortranarray(W0), np.asfortranarray(rs.random((k, n)))


@pytest.mark.parametrize("k,r,p", [(1, 50, 7), (3, 200, 33), (8, 300, 500), (16, 257, 129), (20, 100, 40), (33, 150, 90),
                                   (48, 200, 130), (64, 300, 257)])
def test_nnls_matches_fcnnls(k, r, p):
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(k * 100 + p)
    Cm = rs.random((r, k))
    B = rs.standard_normal((r, p)) + 0.25
    K = mfb.nnls(Cm, B)
    Kref, _ = fcnnls(Cm, B)
    assert (K >= 0).all()
    assert np.array_equal(K == 0, Kref == 0)          # identical active sets
    assert rel_fro(K, Kref) < 1e-11


def _check_replicate(A, k, iters, seed, tol=1e-8):
    import mfbiclust_b200 as mfb
    m, n = A.shape
    W0, H0 = make_init(m, n, k, seed)
    got = mfb.als_replicate(A, k, W0, H0, maxIter=iters, fixed_iters=True)
    tr = []
    ref = oracle_replicate(A, W0, H0, max_iter=iters, fixed_iters=True, trace=tr)
    assert got["iters"] == iters
    assert rel_fro(got["W"], ref["W"]) < tol, rel_fro(got["W"], ref["W"])
    assert rel_fro(got["H"], ref["H"]) < tol, rel_fro(got["H"], ref["H"])
    assert np.array_equal(got["W"] == 0, ref["W"] == 0)
    assert np.array_equal(got["H"] == 0, ref["H"] == 0)
    tr = np.array([t[1:] for t in tr])
    assert np.allclose(got["trace"], tr, rtol=1e-9, atol=1e-12)
    return got, ref


def test_als_simdata_k5(simdata):
    # BASELINE.json configs[0]: addStrat(BiclusterExperiment(simdata), k=5, method='als-nmf')
    import mfbiclust_b200 as mfb
    for name in ("fiveSquareConstant", "plaid"):
        A = pseudovalues(simdata[name])
        got, ref = _check_replicate(A, 5, 10, seed=41)
        # Otsu / threshold on identical inputs: bit-exact thresholds, identical memberships
        ft, st, rn, nc = O.threshold_helper(got["W"], got["H"])
        gft, gst = mfb.otsuHelper(got["W"]), mfb.otsuHelper(got["H"].T)
        assert np.array_equal(gft, ft), name
        assert np.array_equal(gst, st), name
        assert np.array_equal(mfb.threshold(got["W"], gft, 2), rn), name
        assert np.array_equal(mfb.threshold(got["H"], gst, 1), nc), name
        # and the memberships agree with the ones the oracle derives from ITS factors
        _, _, rn_ref, nc_ref = O.threshold_helper(ref["W"], ref["H"])
        assert np.array_equal(rn, rn_ref) and np.array_equal(nc, nc_ref), name


def test_als_simdata_degenerate_restarts(simdata):
    """fiveSquareRowShift is rank deficient for k = 5: in R every other .fcnnls call dies in
    solve() ("computationally singular", cond(HH') ~ 1e16) and als_nmf gives up after the 9th
    restart (R/bicluster.R:126-131).  The singular/regular decision sits at rcond ~ eps, so only
    the behaviour is compared: the GPU path reports the same failures and the same warning."""
    import mfbiclust_b200 as mfb
    from oracle.als_nmf import als_replicate as orep
    A = pseudovalues(simdata["fiveSquareRowShift"])
    W0, H0 = make_init(80, 80, 5, 41)
    g = np.random.default_rng(5)

    def redraw():
        W = g.random((80, 5))
        return W / np.sqrt((W ** 2).sum(0))[None, :]
    ref = orep(A, W0, H0, max_iter=10, fixed_iters=True, redraw=redraw)
    assert ref["status"] == "failed"
    with pytest.warns(UserWarning, match="Factorization failed too many times"):
        got = mfb.als_replicate(A, 5, W0, H0, maxIter=10, fixed_iters=True, rng=mfb.HostRng(5))
    assert np.isfinite(got["W"]).all()


def test_als_simdata_overlapped_restart_protocol(simdata):
    """The fourth simdata matrix (BASELINE.json configs[0] runs all four): fiveSquareOverlapped has rank 4, so at
    k = 5 .fcnnls dies in solve() ("computationally singular") on some iterations and als_nmf's restart() redraws W
    (R/bicluster.R:107-116, :123-161).  Whether a given Gram matrix counts as singular is decided at rcond ~ eps --
    it differs between BLAS builds of the ORACLE itself (the same start survives after 2 restarts on one host and
    fails 10 times on another), so the trajectory is only compared when both sides took the same number of
    restarts; the protocol is checked always: same start, same redraw stream, R's warning after the 9th restart,
    finite non-negative factors, thresholds and memberships bit-exact on the factors returned."""
    import warnings
    import mfbiclust_b200 as mfb
    from oracle.als_nmf import als_replicate as orep
    A = pseudovalues(simdata["fiveSquareOverlapped"])
    W0, H0 = make_init(80, 80, 5, 41)

    class Stream:                                   # runif() source shared by both sides
        def __init__(self):
            self.g = np.random.default_rng(5)
            self.calls = 0

        def runif(self, n):
            self.calls += 1
            return self.g.random(n)

    so, sg = Stream(), Stream()

    def redraw():
        W = so.runif(80 * 5).reshape((80, 5), order="F")
        return W / np.sqrt((W ** 2).sum(0))[None, :]
    tr = []
    ref = orep(A, W0, H0, max_iter=12, fixed_iters=True, redraw=redraw, trace=tr)
    with warnings.catch_warnings(record=True) as wrn:
        warnings.simplefilter("always")
        got = mfb.als_replicate(A, 5, W0, H0, maxIter=12, fixed_iters=True, rng=sg)
    gave_up = any("Factorization failed too many times" in str(w.message) for w in wrn)
    assert sg.calls <= 9 and gave_up == (sg.calls == 9)
    assert np.isfinite(got["W"]).all() and np.isfinite(got["H"]).all()
    assert (got["W"] >= 0).all() and (got["H"] >= 0).all()
    if ref["status"] == "maxiter" and sg.calls == so.calls and not gave_up:
        assert got["iters"] == ref["iters"] == 12
        assert rel_fro(got["W"], ref["W"]) < 1e-8 and rel_fro(got["H"], ref["H"]) < 1e-8
        assert np.array_equal(got["W"] == 0, ref["W"] == 0) and np.array_equal(got["H"] == 0, ref["H"] == 0)
    if not gave_up:
        ft, st, rn, nc = O.threshold_helper(got["W"], got["H"])
        assert np.array_equal(mfb.otsuHelper(got["W"]), ft) and np.array_equal(mfb.otsuHelper(got["H"].T), st)
        assert np.array_equal(mfb.threshold(got["W"], ft, 2), rn)
        assert np.array_equal(mfb.threshold(got["H"], st, 1), nc)


def test_als_yeast_k3(yeast):
    A = pseudovalues(yeast[0])
    _check_replicate(A, 3, 10, seed=42)


@pytest.mark.parametrize("m,n,k", [(2000, 500, 16), (1500, 700, 10), (640, 1030, 24), (900, 600, 40), (600, 400, 64)])
def test_als_synthetic(m, n, k):
    A = pseudovalues(gen_sim_data(n=4, cluster_height=min(m, n) // 5, cluster_width=min(m, n) // 5, dimx=m, dimy=n,
                                  bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    _check_replicate(A, k, 10, seed=43)


def test_als_convergence_matches_oracle(yeast):
    import mfbiclust_b200 as mfb
    A = pseudovalues(yeast[0])
    W0, H0 = make_init(*A.shape, 3, 7)
    got = mfb.als_replicate(A, 3, W0, H0, maxIter=100, eps_conv=1e-4)
    ref = oracle_replicate(A, W0, H0, max_iter=100, eps_conv=1e-4)
    assert got["converged"] and ref["status"] == "converged"
    assert got["iters"] == ref["iters"]
    assert rel_fro(got["W"], ref["W"]) < 1e-8 and rel_fro(got["H"], ref["H"]) < 1e-8
    assert abs(got["obj"] - ref["obj"]) < 1e-10


def test_nnls_is_optimal_where_the_reference_hits_its_sweep_cap():
    """NMF::.fcnnls caps its inner loop with a sweep counter that is cumulative over ALL columns (maxiter = 3k,
    SURVEY.md App. A.1): on a large enough problem at k = 64 the reference (and the literal oracle) stops early and
    returns a K that violates the KKT conditions.  The GPU solver has no such cap across problems: it must return
    the optimum, i.e. agree with an independent exact NNLS and satisfy KKT.  (Knowing divergence, DESIGN.md.)"""
    import scipy.optimize as so
    import mfbiclust_b200 as mfb
    m, n, k = 1200, 800, 64
    A = pseudovalues(gen_sim_data(n=4, cluster_height=160, cluster_width=160, dimx=m, dimy=n, bg_norm=1.0,
                                  bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    W0, _ = make_init(m, n, k, 43)
    K = mfb.nnls(W0, A)
    G, R = W0.T @ W0, W0.T @ A
    w = R - G @ K
    assert (K >= 0).all()
    assert (w[K == 0] <= 1e-8 * np.abs(R).max()).all()            # no positive gradient left on the active set
    assert np.abs(w[K > 0]).max() <= 1e-8 * np.abs(R).max()       # stationarity on the passive set
    for j in range(0, n, 80):
        x, _ = so.nnls(W0, A[:, j])
        assert np.abs(x - K[:, j]).max() < 1e-8 * max(1.0, np.abs(x).max())


def test_fixture_w_halfstep_k1(ref_bces):
    # tests/testdata/ref_bces.rda pins the W half-step arithmetic: stored W == NNLS(t(H), t(A))  (R/bicluster.R:150)
    import mfbiclust_b200 as mfb
    g = ref_bces["bce.als_nmf"]
    A = np.array(ref_bces["input"])
    W, H = np.array(g["W"]), np.array(g["H"])
    Wt = mfb.nnls(H.T, A.T)
    assert np.abs(Wt.T - W).max() < 1e-14
    bcs = mfb.BiclusterStrategy(mfb.genericFit(mfb.genericFactorization(W, H), "als-nmf"), 1, method="als-nmf")
    assert np.array_equal(bcs.featureThresh, np.array(g["featureThresh"]))
    assert np.array_equal(bcs.sampleThresh, np.array(g["sampleThresh"]))
    assert np.array_equal(bcs.biclust.RowxNumber.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(bcs.biclust.NumberxCol.astype(int), np.array(g["NumberxCol"]))


@pytest.mark.parametrize("m,n,k", [(6, 6, 1), (17, 5, 2), (5, 40, 3), (130, 33, 7), (33, 300, 4)])
def test_als_tiny_and_ragged(m, n, k):
    # fewer work units than CTAs, single tiles, dimensions that are not multiples of any tile size
    rs = np.random.default_rng(m * 1000 + n)
    A = np.asfortranarray(rs.random((m, n)) + 0.05)
    _check_replicate(A, k, 6, seed=m + n)


def test_als_rejects_negative_and_bad_eps():
    import mfbiclust_b200 as mfb
    A = np.ones((8, 8))
    A[2, 3] = -1.0
    W0, H0 = make_init(8, 8, 2, 1)
    with pytest.raises(ValueError, match="Negative values are not allowed"):          # R/bicluster.R:76
        mfb.als_nmf(A, 2)
    with pytest.raises(ValueError, match="eps_conv"):                                 # R/bicluster.R:86
        mfb.als_nmf(np.abs(A), 2, eps_conv=0.0)
    with pytest.raises(mfb.MfbError, match="Negative values are not allowed"):
        mfb.AlsSession(mfb.Matrix(mfb.default_context(), host=A), 2, W0, H0, 1.0)


def _d4(k, seed_init):
    m, n = 20000, 5000
    A = np.asfortranarray(pseudovalues(gen_sim_data(n=5, cluster_height=1000, cluster_width=1000, dimx=m, dimy=n,
                                                    bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0,
                                                    seed=4)))
    W0, H0 = make_init(m, n, k, seed_init)
    return A, W0, H0


def test_als_full_size_matches_oracle():
    """BASELINE.json configs[3] itself (20000 x 5000, k = 16, the bench workload): three literal oracle iterations
    take a few seconds on the host; the GPU trajectory must equal them -- factors to 1e-8 relative Frobenius error,
    identical zero patterns, identical traces -- not merely end in a KKT point."""
    import mfbiclust_b200 as mfb
    A, W0, H0 = _d4(16, 40)
    tr = []
    ref = oracle_replicate(A, W0, H0, max_iter=3, fixed_iters=True, trace=tr)
    got = mfb.als_replicate(A, 16, W0, H0, maxIter=3, fixed_iters=True)
    assert rel_fro(got["W"], ref["W"]) < 1e-8, rel_fro(got["W"], ref["W"])
    assert rel_fro(got["H"], ref["H"]) < 1e-8, rel_fro(got["H"], ref["H"])
    assert np.array_equal(got["W"] == 0, ref["W"] == 0) and np.array_equal(got["H"] == 0, ref["H"] == 0)
    assert np.allclose(got["trace"], np.array([t[1:] for t in tr]), rtol=1e-9, atol=1e-12)


def test_als_full_size_k64_properties():
    """BASELINE.json configs[3]'s k = 64 variant at full size.  The literal oracle cannot be the checker here: at
    k = 64 NMF::.fcnnls hits its cumulative sweep cap (see test_nnls_is_optimal_where_the_reference_hits_its_sweep_cap)
    and returns non-optimal columns.  Size-independent properties instead: the W-step ends in the KKT point of its
    NNLS problem, the reported residual equals ||A - WH|| computed independently, reruns are bit-identical."""
    import mfbiclust_b200 as mfb
    A, W0, H0 = _d4(64, 41)
    ctx = mfb.default_context()
    Ad = mfb.Matrix(ctx, host=A)
    try:
        a = mfb.als_replicate(Ad, 64, W0, H0, maxIter=3, fixed_iters=True)
        b = mfb.als_replicate(Ad, 64, W0, H0, maxIter=3, fixed_iters=True)
    finally:
        Ad.free()
    W, H = a["W"], a["H"]
    assert np.array_equal(W, b["W"]) and np.array_equal(H, b["H"])
    assert (W >= 0).all() and (H >= 0).all()
    resid = np.sqrt(np.mean((A - W @ H) ** 2))
    assert abs(a["trace"][-1, 0] - resid) <= 1e-10 * resid
    G, R = H @ H.T, A @ H.T
    grad = R - W @ G
    scale = np.abs(R).max()
    assert np.abs(grad[W > 0]).max() <= 1e-9 * scale
    assert grad[W == 0].max() <= 1e-9 * scale


def test_als_full_size_properties():
    """BASELINE.json configs[3] at full size (20000 x 5000, k = 16) beyond the three iterations the oracle is run
    for in test_als_full_size_matches_oracle: size-independent properties -- the reported residual equals ||A - WH|| computed independently, the
    final W satisfies the KKT conditions of its NNLS problem given the final H, factors are non-negative, the
    residual trace never increases, and two runs are bit-identical."""
    import mfbiclust_b200 as mfb
    m, n, k = 20000, 5000, 16
    A = np.asfortranarray(pseudovalues(gen_sim_data(n=5, cluster_height=1000, cluster_width=1000, dimx=m, dimy=n,
                                                    bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0,
                                                    seed=4)))
    W0, H0 = make_init(m, n, k, 40)
    ctx = mfb.default_context()
    Ad = mfb.Matrix(ctx, host=A)
    try:
        a = mfb.als_replicate(Ad, k, W0, H0, maxIter=6, fixed_iters=True)
        b = mfb.als_replicate(Ad, k, W0, H0, maxIter=6, fixed_iters=True)
    finally:
        Ad.free()
    W, H = a["W"], a["H"]
    assert np.array_equal(W, b["W"]) and np.array_equal(H, b["H"]) and np.array_equal(a["trace"], b["trace"])
    assert (W >= 0).all() and (H >= 0).all()
    resid = np.sqrt(np.mean((A - W @ H) ** 2))
    assert abs(a["trace"][-1, 0] - resid) <= 1e-10 * resid            # no third pass over A, same number
    assert (np.diff(a["trace"][:, 0]) <= 1e-12).all()
    G, R = H @ H.T, A @ H.T                                           # W-step problem of the last iteration
    grad = R - W @ G
    scale = np.abs(R).max()
    assert np.abs(grad[W > 0]).max() <= 1e-9 * scale                  # stationarity on the passive set
    assert grad[W == 0].max() <= 1e-9 * scale                         # no positive gradient left on the active set


def _fp32_case(m, n, k, seed=44):
    A = pseudovalues(gen_sim_data(n=4, cluster_height=min(m, n) // 5, cluster_width=min(m, n) // 5, dimx=m, dimy=n,
                                  bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=m + n))
    W0, H0 = make_init(m, n, k, seed)
    return A, W0, H0


@pytest.mark.parametrize("m,n,k", [(2000, 500, 16), (1037, 333, 10), (70, 45, 3), (640, 384, 8), (6, 6, 1), (1500, 3000, 16)])
def test_als_fp32_tensor_core_path(m, n, k, monkeypatch):
    """FP32 path of BASELINE.json configs[3] (north star: factors within 1e-4 relative Frobenius error of the FP64
    reference after a fixed iteration count): A resident in single precision, the two products of an iteration on the
    TF32 tensor cores (tcgen05.mma, 3-term split, accumulators flushed per 32-deep stage; csrc/als_tf32.cuh), Gram
    matrices, NNLS and scalars in FP64.  Ranks up to 16 take this path."""
    import mfbiclust_b200 as mfb
    monkeypatch.delenv("MFB_F32_DMMA", raising=False)
    A, W0, H0 = _fp32_case(m, n, k)
    ctx = mfb.default_context()
    Ad = mfb.Matrix(ctx, host=A, precision="f32")
    try:
        got = mfb.als_replicate(Ad, k, W0, H0, maxIter=8, fixed_iters=True)
        again = mfb.als_replicate(Ad, k, W0, H0, maxIter=8, fixed_iters=True)
    finally:
        Ad.free()
    ref64 = oracle_replicate(A, W0, H0, max_iter=8, fixed_iters=True)
    assert rel_fro(got["W"], ref64["W"]) < 1e-4, rel_fro(got["W"], ref64["W"])          # tolerance: north star, FP32 path
    assert rel_fro(got["H"], ref64["H"]) < 1e-4, rel_fro(got["H"], ref64["H"])
    assert abs(got["obj"] - ref64["obj"]) < 1e-4 * ref64["obj"]
    assert np.array_equal(got["W"], again["W"]) and np.array_equal(got["H"], again["H"])  # deterministic


@pytest.mark.parametrize("m,n,k", [(2000, 500, 16), (1037, 333, 10), (70, 45, 3), (640, 1030, 24), (600, 400, 64)])
def test_als_fp32_storage(m, n, k, monkeypatch):
    """FP32 STORAGE with FP64 tensor-core accumulation (ranks above 16, or MFB_F32_DMMA=1): against the oracle on the
    float32-rounded matrix it is the same computation (1e-8); against the oracle on the original matrix it stays far
    inside the north-star FP32 tolerance of 1e-4 relative Frobenius error."""
    import mfbiclust_b200 as mfb
    monkeypatch.setenv("MFB_F32_DMMA", "1")
    A, W0, H0 = _fp32_case(m, n, k)
    ctx = mfb.default_context()
    Ad = mfb.Matrix(ctx, host=A, precision="f32")
    try:
        st = Ad.stats()
        A32 = A.astype(np.float32).astype(np.float64)
        assert st["max"] == A32.max() and st["min"] == A32.min()
        got = mfb.als_replicate(Ad, k, W0, H0, maxIter=8, fixed_iters=True)
    finally:
        Ad.free()
    ref32 = oracle_replicate(A32, W0, H0, max_iter=8, fixed_iters=True)
    ref64 = oracle_replicate(A, W0, H0, max_iter=8, fixed_iters=True)
    assert rel_fro(got["W"], ref32["W"]) < 1e-8 and rel_fro(got["H"], ref32["H"]) < 1e-8
    assert rel_fro(got["W"], ref64["W"]) < 1e-4 and rel_fro(got["H"], ref64["H"]) < 1e-4


def test_handles_may_outlive_their_context():
    # finalisers run in arbitrary order under a garbage collector (R's too): destroying the context first must be safe
    import mfbiclust_b200 as mfb
    ctx = mfb.Context(0)
    A = np.asfortranarray(np.random.default_rng(0).random((64, 40)) + 0.1)
    Ad = mfb.Matrix(ctx, host=A)
    W0, H0 = make_init(64, 40, 3, 1)
    sess = mfb.AlsSession(Ad, 3, W0, H0, 1.0)
    sess.iterate(2, fixed_iters=True)
    ctx.close()
    sess.close()
    Ad.free()


@pytest.mark.parametrize("k,r,p", [(1, 40, 9), (3, 200, 33), (8, 300, 500), (16, 257, 129), (20, 100, 40), (32, 150, 90)])
def test_nnls_reports_the_max_col_draws_of_fcnnls(k, r, p):
    """NMF::.fcnnls consumes R's RNG through max.col(ties.method = "random"): the number of
    uniforms is a deterministic function of the values and is reported by the device, so that an R wrapper can
    advance .Random.seed exactly as the reference would have."""
    import mfbiclust_b200 as mfb
    from oracle.r_rng import RRng
    rs = np.random.default_rng(k * 100 + p)
    Cm = rs.random((r, k))
    Bm = rs.standard_normal((r, p)) + 0.25
    rng = RRng(1)
    Kref, _ = fcnnls(Cm, Bm, rng)
    K, draws = mfb.nnls(Cm, Bm, return_draws=True)
    assert rel_fro(K, Kref) < 1e-11
    assert draws == rng.draws, (draws, rng.draws)
    if k == 1:
        assert draws == 0                                   # no ties with a single variable


def test_als_nmf_follows_the_r_session_of_the_vignette(yeast, vignette):
    """set.seed(12345); als_nmf(yeast_benchmark[[1]] + 2.71, k = 3) as rendered in vignettes/my-vignette.html:224-399,
    WITHOUT passing initial factors: the device reports the uniforms the reference's .fcnnls calls consume, the host
    skips them, and replicates 2..4 therefore start from the stream position they start from in R -- same four
    objectives, same winner, same factors as the oracle's run of the whole R function under the R generator."""
    import mfbiclust_b200 as mfb
    from oracle.als_nmf import als_nmf as oracle_als_nmf
    from oracle.r_rng import RRng
    A = pseudovalues(yeast[0])
    r_ref = RRng(12345)
    ref = oracle_als_nmf(A, 3, reps=4, max_iter=100, eps_conv=1e-4, rng=r_ref)
    r_gpu = RRng(12345)
    got = mfb.als_nmf(A, 3, reps=4, maxIter=100, eps_conv=1e-4, rng=r_gpu, follow_rng=True)
    assert r_gpu.draws == r_ref.draws                        # the session ends at the same stream position
    m, n = A.shape
    per_rep = [s["maxcol_draws"] for s in got.solutions]
    assert sum(per_rep) == r_ref.draws - 4 * (m * 3 + 3 * n)
    assert per_rep == [143, 218, 130, 178]                  # the draws that make the vignette reproducible
    assert got.best == ref["best"]
    for gs, rsol in zip(got.solutions, ref["solutions"]):
        assert abs(gs["obj"] - rsol["obj"]) < 1e-9
        assert gs["iters"] == rsol["iters"]
    assert rel_fro(got.fit.W, ref["W"]) < 1e-8 and rel_fro(got.fit.H, ref["H"]) < 1e-8
