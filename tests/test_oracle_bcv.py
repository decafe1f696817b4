"""Oracle checks for bcv / auto_bcv."""
import numpy as np

from oracle import bcv as B
from oracle.r_rng import RRng


def _folds(rng, n, p, h):
    return (B.make_folds(n, h, rng.permutation(n)), B.make_folds(p, h, rng.permutation(p)))


def test_one_svd_per_cell_equals_literal_ginv_path(yeast):
    Y = yeast[5]            # 1030 x 7
    rs = np.random.default_rng(3)
    ks = B.k_limiter(3, min(Y.shape), np.arange(1, 7))
    rf, cf = _folds(rs, *Y.shape, 3)
    lit = B.bcv_given_ks_literal(Y, ks, rf, cf, 3)
    fast = B.bcv_given_ks(Y, ks, rf, cf, 3)
    assert np.abs(lit - fast).max() / np.abs(lit).max() < 1e-12
    assert int(np.argmin(lit)) == int(np.argmin(fast))


def test_k_limiter():
    # R/bcv.R:241-251 : ks < floor((h-1)/h * dim)
    assert list(B.k_limiter(3, 18, np.arange(1, 18))) == list(range(1, 12))


def test_vignette_auto_bcv(yeast, vignette):
    # vignettes/my-vignette.html:408-413: best = 11, all 22 counts at k = 11.
    Y = yeast[0]
    rs = np.random.default_rng(11)
    best, counts, iters = B.auto_bcv(Y, np.arange(1, min(Y.shape)), 3,
                                     lambda: _folds(rs, *Y.shape, 3))
    assert best == vignette["auto_bcv_best"]
    assert iters == 22
    assert list(counts) == [vignette["auto_bcv_counts"][str(k)] if isinstance(
        next(iter(vignette["auto_bcv_counts"])), str) else vignette["auto_bcv_counts"][k]
        for k in range(1, 12)]


def test_r_sample_is_a_permutation():
    rng = RRng(1)
    p = rng.sample(range(1, 101))
    assert sorted(p) == list(range(1, 101))
