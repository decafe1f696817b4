"""GPU parity tests for bcv / auto_bcv (C ABI -> CUDA) against the CPU oracle."""
import numpy as np
import pytest

from oracle import bcv as B
from oracle.gen_sim_data import gen_sim_data

pytestmark = pytest.mark.gpu


def _folds(rs, n, p, h):
    return (B.make_folds(n, h, rs.permutation(n)).astype(np.int32), B.make_folds(p, h, rs.permutation(p)).astype(np.int32))


def test_bcv_cells_match_oracle_yeast(yeast):
    import mfbiclust_b200 as mfb
    for idx, h in ((0, 3), (1, 3), (5, 3), (0, 5)):
        Y = yeast[idx]
        rs = np.random.default_rng(100 + idx + h)
        ks = B.k_limiter(h, min(Y.shape), np.arange(1, min(Y.shape)))
        rf, cf = _folds(rs, *Y.shape, h)
        got, cells = mfb.bcvGivenKs(Y, ks, h, folds=(rf, cf), cell_errors=True)
        ref = B.bcv_given_ks(Y, ks, rf, cf, h)
        assert cells.shape == (h * h, len(ks))
        # FP64 end to end; the curve, not only its argmin
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8, (idx, h, got, ref)
        assert int(np.argmin(got)) == int(np.argmin(ref))


def test_bcv_literal_ginv_path_small(yeast):
    # the literal per-k MASS::ginv path of R/bcv.R:214-227 on a small case
    import mfbiclust_b200 as mfb
    Y = yeast[5]
    rs = np.random.default_rng(3)
    ks = B.k_limiter(3, min(Y.shape), np.arange(1, 7))
    rf, cf = _folds(rs, *Y.shape, 3)
    got = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf))
    lit = B.bcv_given_ks_literal(Y, ks, rf, cf, 3)
    assert np.abs(got - lit).max() / np.abs(lit).max() < 1e-8


def test_bcv_synthetic_planted(yeast):
    import mfbiclust_b200 as mfb
    Y = gen_sim_data(n=5, cluster_height=60, cluster_width=60, dimx=900, dimy=300, bg_norm=1.0, bicluster_shift=3.0,
                     row_shift=1.0, col_shift=1.0, seed=3)
    rs = np.random.default_rng(50)
    ks = np.arange(1, 21)
    rf, cf = _folds(rs, *Y.shape, 3)
    got = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf))
    ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8
    assert int(np.argmin(got)) == int(np.argmin(ref))


def test_bcv_wide_holdin(yeast):
    # hold-in block with fewer rows than columns (row-side Gram)
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(9)
    Y = rs.standard_normal((40, 3)) @ rs.standard_normal((3, 150)) + 0.3 * rs.standard_normal((40, 150))
    ks = np.arange(1, 9)
    rf, cf = _folds(rs, *Y.shape, 4)
    got = mfb.bcvGivenKs(Y, ks, 4, folds=(rf, cf))
    ref = B.bcv_given_ks(Y, ks, rf, cf, 4)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8


def test_auto_bcv_vignette(yeast, vignette):
    # vignettes/my-vignette.html:408-413: auto_bcv(yeast_benchmark[[1]]) -> best = 11, all 22 counts at k = 11
    import mfbiclust_b200 as mfb
    Y = yeast[0]
    res = mfb.auto_bcv(Y, rng=mfb.HostRng(11))
    assert res["best"] == vignette["auto_bcv_best"] == 11
    assert res["iterations"] == 22
    assert res["counts"][11] == 22 and sum(res["counts"].values()) == 22
    # same folds -> same decisions as the oracle's auto_bcv
    rs = mfb.HostRng(11)

    def draw():
        return mfb.draw_folds(rs, *Y.shape, 3)
    best, counts, iters = B.auto_bcv(Y, np.arange(1, min(Y.shape)), 3, draw)
    assert best == res["best"] and iters == res["iterations"]
    assert list(counts) == [res["counts"][k] for k in sorted(res["counts"])]


def test_auto_bcv_speculation_is_invisible(yeast):
    # computing several outer iterations per batch must not change best / counts / iteration count
    import mfbiclust_b200 as mfb
    Y = yeast[2]
    a = mfb.auto_bcv(Y, rng=mfb.HostRng(7), speculate=1)
    b = mfb.auto_bcv(Y, rng=mfb.HostRng(7), speculate=5)
    assert a == b


def test_bcv_streams_do_not_change_values(yeast):
    # cells spread over concurrent streams (worker contexts on views of the resident matrix): same values bit for bit
    import mfbiclust_b200 as mfb
    Y = yeast[1]
    rs = np.random.default_rng(77)
    ks = B.k_limiter(3, min(Y.shape), np.arange(1, min(Y.shape)))
    rf, cf = _folds(rs, *Y.shape, 3)
    a, ca = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf), cell_errors=True, streams=1)
    b, cb = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf), cell_errors=True, streams=4)
    assert np.array_equal(ca, cb) and np.array_equal(a, b)


def test_bcv_batches_are_bit_identical_for_any_stream_count():
    """Cells of several fold draws computed as one batch on concurrent streams (auto_bcv's first batch) must equal,
    bit for bit, the same cells computed one draw at a time on one stream, run after run.  (A grid-barrier-per-column
    eigen-solver that reuses a buffer one barrier too early passes every single-stream test and fails this one.)"""
    import mfbiclust_b200 as mfb
    from mfbiclust_b200.backends import _bcv_cells_all
    from oracle.gen_sim_data import gen_sim_data
    Y = np.asfortranarray(gen_sim_data(n=4, cluster_height=300, cluster_width=150, dimx=4000, dimy=900, bg_norm=1.0,
                                       bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=11))
    ctx = mfb.default_context()
    Yd = mfb.Matrix(ctx, host=Y)
    try:
        ks = list(range(1, 13))
        rng = mfb.HostRng(3)
        folds = [mfb.draw_folds(rng, *Y.shape, 3) for _ in range(8)]
        one = np.stack([_bcv_cells_all(Yd, ks, 3, [f], ctx, None, 1)[0] for f in folds])
        for streams in (4, 2, 6, 4):
            got = _bcv_cells_all(Yd, ks, 3, folds, ctx, None, streams)
            assert np.array_equal(got, one), (streams, float(np.abs(got - one).max()))
    finally:
        Yd.free()


def _d3(dimx=10000, dimy=2000, seed=3):
    side = min(dimx, dimy) // 5
    return np.asfortranarray(gen_sim_data(n=5, cluster_height=side, cluster_width=side, dimx=dimx, dimy=dimy,
                                          bg_norm=1.0, bicluster_shift=3.0, row_shift=1.0, col_shift=1.0, seed=seed))


def test_bcv_d3_curve_matches_oracle():
    """BASELINE.json configs[2] at full size (genSimData-shaped 10000 x 2000, ks = 1:20, holdouts = 3): the whole BCV
    curve of one fold draw against the oracle (one LAPACK SVD per cell, a few seconds on the host), not only the
    argmin -- the curve is flat around its minimum."""
    import mfbiclust_b200 as mfb
    Y = _d3()
    ks = np.arange(1, 21)
    rf, cf = mfb.draw_folds(mfb.HostRng(50), *Y.shape, 3)
    got = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf))
    ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8
    assert np.abs(got / ref - 1.0).max() < 1e-8                  # relative, entry by entry
    assert int(np.argmin(got)) == int(np.argmin(ref))


def test_auto_bcv_matches_oracle_planted():
    """auto_bcv against the oracle's auto_bcv on the same fold stream: a D3-shaped matrix at a quarter of the size
    (2500 x 500, 5 planted biclusters, ks = 1:20, holdouts = 3) so that the oracle's 25+ sequential bcv calls finish
    in seconds.  Same best, same counts, same number of outer iterations (R/bcv.R:56-97)."""
    import mfbiclust_b200 as mfb
    Y = _d3(2500, 500, seed=3)
    ks = np.arange(1, 21)
    res = mfb.auto_bcv(Y, ks, holdouts=3, rng=mfb.HostRng(51))
    rs = mfb.HostRng(51)
    best, counts, iters = B.auto_bcv(Y, ks, 3, lambda: mfb.draw_folds(rs, *Y.shape, 3))
    assert res["best"] == best and res["iterations"] == iters
    assert [res["counts"][int(k)] for k in ks] == list(counts)


@pytest.mark.parametrize("case", ["exact_rank3", "duplicated_columns", "rank3_tiny_noise", "wide_centred_full_k"])
def test_bcv_rank_deficient_holdin(case):
    """Hold-in blocks whose trailing singular values are (numerically) zero: MASS::ginv drops every component with
    s_j <= sqrt(eps) s_1 (R/bcv.R:222), LAPACK finds them at ~1e-16 s_1.  A Gram-matrix eigenvalue of a null direction
    is only ~eps * lambda_1, i.e. sqrt() of it sits right at the 1.5e-8 cut: the keep/drop decision must come from
    s_j = ||D v_j|| computed directly, and ks beyond the rank must reproduce the reference's plateau."""
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(123)
    if case == "exact_rank3":
        Y = rs.standard_normal((120, 3)) @ rs.standard_normal((3, 45))
        h, ks = 3, np.arange(1, 13)
    elif case == "duplicated_columns":
        base = rs.standard_normal((90, 6)) @ rs.standard_normal((6, 20)) + 0.5 * rs.standard_normal((90, 20))
        Y = np.concatenate([base, base, base[:, :8]], axis=1)          # 48 columns, rank 20
        h, ks = 3, np.arange(1, 30)
    elif case == "rank3_tiny_noise":
        Y = rs.standard_normal((150, 3)) @ rs.standard_normal((3, 60)) + 1e-11 * rs.standard_normal((150, 60))
        h, ks = 3, np.arange(1, 16)
    else:
        # 12 rows: a hold-in block has 8 rows and 40 columns; centring leaves rank 7 and prcomp(rank. = 8) still
        # returns 8 components, the last one with s ~ 1e-16 s_1
        Y = rs.standard_normal((12, 60))
        h, ks = 3, np.arange(1, 8)
        ks = np.arange(1, 8)
    Y = np.asfortranarray(Y)
    rf, cf = _folds(rs, *Y.shape, h)
    got, cells = mfb.bcvGivenKs(Y, ks, h, folds=(rf, cf), cell_errors=True)
    ref = B.bcv_given_ks(Y, ks, rf, cf, h)
    scale = np.abs(ref).max()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() / scale < 1e-8, (case, got, ref)


def test_bcv_wide_centred_keeps_k_equal_rows():
    # ks reaching the number of hold-in rows: R's kLimiter admits k < floor((h-1)/h * min(dim)); with 12 rows, h = 3
    # that is k <= 7 while the centred 8-row block has rank 7 -- and with min(dim) on the column side the same
    # happens through prcomp's k <- min(rank., dim(D)).  Explicit ks beyond kLimiter go through bcvGivenKs directly.
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(5)
    Y = np.asfortranarray(rs.standard_normal((12, 60)))
    rf, cf = _folds(rs, 12, 60, 3)
    ks = np.arange(1, 11)                     # 8, 9, 10 exceed the block's 8 rows: clamped to 8 components
    got = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf))
    ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-8
    assert got[7] == got[8] == got[9]


def test_bcv_default_ks_beyond_64_components():
    """bcv(Y) / auto_bcv(Y) with the reference's default ks = seq_len(min(dim) - 1) clipped by kLimiter
    (R/bcv.R:31, :124, :241-251): on a 1000 x 200 matrix that is ks = 1..132 -- far beyond any fixed component cap."""
    import mfbiclust_b200 as mfb
    Y = _d3(1000, 200, seed=8)
    rs = np.random.default_rng(2)
    rf, cf = _folds(rs, *Y.shape, 3)
    with pytest.warns(UserWarning, match="not all ks can be tested"):
        got = mfb.bcv(Y, holdouts=3, folds=(rf, cf))
    ks = B.k_limiter(3, 200, np.arange(1, 200))
    assert list(got) == [int(k) for k in ks] and len(ks) == 132
    ref = B.bcv_given_ks(Y, ks, rf, cf, 3)
    g = np.array([got[int(k)] for k in ks])
    assert np.abs(g - ref).max() / np.abs(ref).max() < 1e-8
    assert int(np.argmin(g)) == int(np.argmin(ref))
    with pytest.warns(UserWarning):
        res = mfb.auto_bcv(Y, holdouts=3, rng=mfb.HostRng(4), maxIter=4)      # default ks; a few outer iterations
    assert res["best"] in got


def test_svd_pca_more_than_64_components():
    import mfbiclust_b200 as mfb
    from oracle import pca as P
    rs = np.random.default_rng(6)
    A = np.asfortranarray(rs.standard_normal((300, 7)) @ rs.standard_normal((7, 120)) + rs.standard_normal((300, 120)))
    fit = mfb.svd_pca(A, 100)
    Wr, Hr = P.svd_pca(A, 100)
    sgn = np.sign((fit.fit.W * Wr).sum(axis=0))
    assert np.abs(fit.fit.W * sgn[None, :] - Wr).max() < 1e-8
    assert np.abs(fit.fit.H * sgn[:, None] - Hr).max() < 1e-8 * np.abs(Hr).max()


def test_bcv_with_missing_entries_uses_the_nipals_branch():
    """Y with NA entries (R/bcv.R:177-181): every hold-in block is factored by nipals_pca(center = TRUE, scale = FALSE)
    -- scores and loadings without singular values -- and NA entries of the products are dropped by na.rm.  Against
    the literal restatement (per-k MASS::ginv, explicit NA propagation); parity unpinned, tolerance 1e-6 relative on
    the curve (NIPALS stops at tol = 1e-6 on both sides; the iteration counts agree)."""
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(31)
    Y = rs.standard_normal((240, 4)) @ rs.standard_normal((4, 90)) + 0.4 * rs.standard_normal((240, 90))
    miss = rs.random(Y.shape) < 0.01
    Y[miss] = np.nan
    Y = np.asfortranarray(Y)
    ks = np.arange(1, 7)
    rf, cf = _folds(rs, *Y.shape, 3)
    got = mfb.bcvGivenKs(Y, ks, 3, folds=(rf, cf))
    ref = B.bcv_given_ks_na_literal(Y, ks, rf, cf, 3)
    assert np.isfinite(got).all()
    assert np.abs(got / ref - 1.0).max() < 1e-6, (got, ref)
    assert int(np.argmin(got)) == int(np.argmin(ref))
    res = mfb.bcv(Y, ks, holdouts=3, interactive=False, folds=(rf, cf))
    assert list(res) == [int(k) for k in ks] and np.allclose([res[int(k)] for k in ks], got)
    out = mfb.auto_bcv(Y, ks, holdouts=3, rng=mfb.HostRng(3), interactive=False, maxIter=30)
    assert out["best"] in [int(k) for k in ks]
