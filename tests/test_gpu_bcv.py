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
