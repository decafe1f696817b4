"""GPU parity tests for the bicluster post-processing (C ABI -> CUDA: mfb_bicluster_overlap / mfb_bicluster_union and
the host loop of filter.biclust above them) against the literal oracle restatement of R/helperFunctions.R:150-346.
Integer / bit work: the bar is exact equality."""
import numpy as np
import pytest

from oracle import helpers as H

pytestmark = pytest.mark.gpu


def _random_memberships(m, n, k, seed, density=0.3):
    rs = np.random.default_rng(seed)
    rowx = rs.random((m, k)) < density
    bicx = rs.random((k, n)) < density
    return rowx, bicx


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 2), (33, 64, 3), (100, 31, 5), (1000, 24, 10), (257, 130, 64),
                                   (64, 4100, 65), (5000, 33, 130)])
def test_overlap_and_sizes_match_the_oracle(m, n, k):
    import mfbiclust_b200 as mfb
    rowx, bicx = _random_memberships(m, n, k, seed=m * 7 + n + k)
    if k >= 3:
        rowx[:, 1] = False                      # an empty bicluster: its overlaps are 0 / 0 = NaN, as in R
        rowx[:, 2] = True
        bicx[2, :] = True                       # one that covers the whole matrix
    rows, cols = H.bicluster_matrix2list(rowx, bicx)
    sz, ov = mfb.overlap_sizes(rowx, bicx)
    assert np.array_equal(sz, H.sizes(rows, cols))
    ref = H.overlap(rows, cols)
    assert np.array_equal(np.isnan(ov), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(ov, nan=-1.0), np.nan_to_num(ref, nan=-1.0))      # same integers, same division
    got_rows, got_cols = mfb.biclusterMatrix2List(rowx, bicx)
    assert all(np.array_equal(a, b) for a, b in zip(got_rows, rows)) and all(np.array_equal(a, b) for a, b in zip(got_cols, cols))
    lst = mfb.overlap(rowx, bicx)               # matrix = FALSE: list of rows
    assert len(lst) == k and np.array_equal(np.nan_to_num(lst[0], nan=-1.0), np.nan_to_num(ref[0], nan=-1.0))


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 2), (100, 31, 5), (1000, 24, 10), (257, 130, 64), (300, 70, 65),
                                   (2000, 1500, 16)])
def test_union_matches_the_oracle(m, n, k):
    import mfbiclust_b200 as mfb
    rowx, bicx = _random_memberships(m, n, k, seed=m + 3 * n + k, density=0.15)
    assert np.array_equal(mfb.union_biclust(rowx, bicx), H.union_biclust(rowx, bicx))
    choose = np.random.default_rng(k).random(k) < 0.5
    assert np.array_equal(mfb.union_biclust(rowx, bicx, choose), H.union_biclust(rowx, bicx, choose))
    none = np.zeros(k, dtype=bool)
    assert not mfb.union_biclust(rowx, bicx, none).any()


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("limit,mx", [(0.25, None), (0.0, None), (0.6, 2), (1.0, None)])
def test_filter_biclust_matches_the_oracle(seed, limit, mx):
    import mfbiclust_b200 as mfb
    m, n, k = 120 + 13 * seed, 60 + 7 * seed, 3 + seed
    rowx, bicx = _random_memberships(m, n, k, seed=100 + seed, density=0.35)
    if seed % 2:
        rowx[:, 0] = False                      # empty
        rowx[:, k - 1] = True
        bicx[k - 1, :] = True                   # whole matrix
    if seed >= 4:
        rowx[:, 2] = rowx[:, 1]
        bicx[2, :] = bicx[1, :]                 # a duplicate: overlap 1 with its twin
    got = mfb.filter_biclust(rowx, bicx, max=mx, overlap=limit)
    ref = H.filter_biclust(rowx, bicx, max=mx, overlap_limit=limit)
    assert np.array_equal(got["chosen"], ref["chosen"])
    assert np.array_equal(got["RowxBicluster"], ref["RowxBicluster"])
    assert np.array_equal(got["BiclusterxCol"], ref["BiclusterxCol"])
    assert np.array_equal(got["biclustered"], ref["biclustered"])


def test_filter_biclust_edge_cases():
    import mfbiclust_b200 as mfb
    rowx = np.zeros((5, 0), dtype=bool)
    bicx = np.zeros((0, 4), dtype=bool)
    r = mfb.filter_biclust(rowx, bicx)
    assert r["chosen"].size == 0 and r["biclustered"].shape == (5, 4) and not r["biclustered"].any()
    one = mfb.filter_biclust(np.ones((5, 1), dtype=bool), np.ones((1, 4), dtype=bool))      # k = 1: kept without a test
    assert one["chosen"].tolist() == [True] and one["biclustered"].all()
    with pytest.raises(ValueError, match="same number of columns"):
        mfb.filter_biclust(np.ones((5, 2), dtype=bool), np.ones((3, 4), dtype=bool))


def test_filter_biclust_on_a_strategy_of_the_vignette_shape():
    """The doc example of filter.biclust (R/helperFunctions.R:139-147): als-nmf memberships of a planted matrix."""
    import mfbiclust_b200 as mfb
    from oracle.gen_sim_data import gen_sim_data
    A = H.pseudovalues(gen_sim_data(n=4, cluster_height=30, cluster_width=30, dimx=200, dimy=160, bg_norm=1.0,
                                    bicluster_shift=4.0, overlap_rows=10, overlap_cols=10, seed=5))
    bce = mfb.addStrat(mfb.BiclusterExperiment(A), 4, method="svd-pca", silent=True)
    bcs = list(bce.strategies.values())[-1]
    rowx, bicx = np.asarray(bcs.biclust.RowxNumber, dtype=bool), np.asarray(bcs.biclust.NumberxCol, dtype=bool)
    got = mfb.filter_biclust(rowx, bicx, overlap=0.25)
    ref = H.filter_biclust(rowx, bicx, overlap_limit=0.25)
    assert np.array_equal(got["chosen"], ref["chosen"]) and np.array_equal(got["biclustered"], ref["biclustered"])
