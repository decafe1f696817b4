"""GPU parity tests for otsuHelper / threshold (R/BiclusterStrategy.R:442-456, 508-528) and the resident-matrix
helpers (max(A), any(A < 0), pseudovalues: R/bicluster.R:76,81; R/helperFunctions.R:270-278)."""
import numpy as np
import pytest

from oracle import otsu as O

pytestmark = pytest.mark.gpu


def test_otsu_bit_exact_on_many_shapes(yeast):
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(12)
    mats = [rs.standard_normal((1000, 7)), rs.random((37, 3)) ** 3, np.abs(rs.standard_normal((5000, 16))) * 1e-3,
            rs.standard_normal((257, 1)) * 1e6 - 3e6, yeast[0], yeast[3].T.copy()]
    for M in mats:
        got = mfb.otsuHelper(M)
        ref = O.otsu_helper(np.asarray(M, dtype=np.float64))
        assert np.array_equal(got, ref)                      # bit-exact thresholds
        assert np.array_equal(mfb.threshold(M, got, 2), O.threshold(M, ref, 2))


def test_otsu_constant_column_and_negative_thresholds():
    import mfbiclust_b200 as mfb
    M = np.zeros((50, 4))
    M[:, 0] = 3.25                                   # constant column: threshold is x[1] itself (:451-453)
    M[:, 1] = np.linspace(-5, -1, 50)                # all negative: th < 0 flips the comparison to `<` (:514,:520)
    M[:, 2] = np.r_[np.zeros(25), np.ones(25)]
    M[:, 3] = np.linspace(-2, 2, 50)
    th = mfb.otsuHelper(M)
    ref = O.otsu_helper(M)
    assert np.array_equal(th, ref) and th[0] == 3.25 and th[1] < 0
    assert np.array_equal(mfb.threshold(M, th, 2), O.threshold(M, ref, 2))
    # MARGIN = 1: one threshold per row
    T = np.asfortranarray(M.T)
    assert np.array_equal(mfb.threshold(T, th, 1), O.threshold(T, ref, 1))
    with pytest.raises(ValueError, match="Length of th"):
        mfb.threshold(M, th[:2], 2)


def test_matrix_stats_shift_and_trim():
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(3)
    A = rs.standard_normal((333, 77))
    A[5, 6] = np.nan
    ctx = mfb.default_context()
    Ad = mfb.Matrix(ctx, host=A)
    st = Ad.stats()
    assert st["n_nan"] == 1 and st["min"] == np.nanmin(A) and st["max"] == np.nanmax(A)
    assert abs(st["sumsq"] - np.nansum(A ** 2)) <= 1e-12 * np.nansum(A ** 2)
    Ad.shift(abs(st["min"]))                         # pseudovalues(): A + abs(min(A))
    st2 = Ad.stats()
    assert st2["min"] == 0.0 and abs(st2["max"] - (np.nanmax(A) + abs(np.nanmin(A)))) < 1e-15 * 10
    Ad.free()
    ctx.trim()                                       # cached device buffers go back to the driver; context stays usable
    assert np.array_equal(mfb.otsuHelper(np.arange(20.0).reshape(10, 2)), O.otsu_helper(np.arange(20.0).reshape(10, 2)))
