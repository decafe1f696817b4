"""GPU parity tests for svd_pca / nipals_pca (C ABI -> CUDA) against the CPU oracle and the reference fixture."""
import numpy as np
import pytest

from oracle import otsu as O
from oracle import pca as P

pytestmark = pytest.mark.gpu


def rel_fro(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def align_signs(W, H, Wref):
    """LAPACK's and the device solver's component signs are both arbitrary:
    flip W[:, j] / H[j, :] together onto the oracle's orientation before comparing."""
    s = np.sign((W * Wref).sum(axis=0))
    s[s == 0] = 1.0
    return W * s[None, :], H * s[:, None]


def test_svd_pca_fixture_k1(ref_bces):
    # tests/testthat/test_BiclusterExperiment.R:39-41 against tests/testdata/ref_bces.rda
    import mfbiclust_b200 as mfb
    g = ref_bces["bce.svd_pca"]
    A = np.array(ref_bces["input"])
    fit = mfb.svd_pca(A, 1)
    Wg, Hg = np.array(g["W"]), np.array(g["H"])
    W, H = align_signs(fit.fit.W, fit.fit.H, Wg)
    assert np.abs(W - Wg).max() < 1e-13
    assert np.abs(H - Hg).max() < 1e-12
    bcs = mfb.BiclusterStrategy(mfb.genericFit(mfb.genericFactorization(W, H), "svd-pca"), 1, method="svd-pca")
    assert np.allclose(bcs.featureThresh, g["featureThresh"], rtol=1e-12)
    assert np.array_equal(bcs.biclust.RowxNumber.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(bcs.biclust.NumberxCol.astype(int), np.array(g["NumberxCol"]))


def test_nipals_pca_fixture_k1(ref_bces):
    # tests/testthat/test_BiclusterExperiment.R:61-65
    import mfbiclust_b200 as mfb
    g = ref_bces["bce.nipals_pca"]
    A = np.array(ref_bces["input"])
    bce = mfb.addStrat(mfb.BiclusterExperiment(A), 1, method="nipals-pca", silent=True)
    bcs = bce.strategies["NIPALS-PCA | Otsu | 1"]
    assert np.abs(bcs.factors.fit.W - np.array(g["W"])).max() < 1e-14
    assert np.abs(bcs.factors.fit.H - np.array(g["H"])).max() < 1e-14
    assert np.array_equal(bcs.biclust.RowxNumber.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(bcs.biclust.NumberxCol.astype(int), np.array(g["NumberxCol"]))
    assert np.allclose(bcs.featureThresh, g["featureThresh"], rtol=1e-12)
    assert np.allclose(bcs.sampleThresh, g["sampleThresh"], rtol=1e-12)


def _membership_check(W, H, Wref, Href):
    import mfbiclust_b200 as mfb
    _, _, rn_ref, nc_ref = O.threshold_helper(Wref, Href)
    ft, st = mfb.otsuHelper(W), mfb.otsuHelper(np.asfortranarray(H.T))
    rn, nc = mfb.threshold(W, ft, 2), mfb.threshold(H, st, 1)
    return int((rn != rn_ref).sum() + (nc != nc_ref).sum())


def test_svd_pca_yeast_k10(yeast):
    # BASELINE.json configs[1]: svd_pca k=10 on yeast_benchmark, membership parity after Otsu threshold
    import mfbiclust_b200 as mfb
    flips = 0
    for i, A in enumerate(yeast):
        k = min(10, A.shape[1])
        fit = mfb.svd_pca(A, k)
        Wr, Hr = P.svd_pca(A, k)
        W, H = align_signs(fit.fit.W, fit.fit.H, Wr)
        assert W.shape == Wr.shape and H.shape == Hr.shape
        assert rel_fro(W, Wr) < 1e-8, (i, rel_fro(W, Wr))
        assert rel_fro(H, Hr) < 1e-8, (i, rel_fro(H, Hr))
        flips += _membership_check(W, H, Wr, Hr)
    assert flips == 0


def test_svd_pca_wide_and_tall():
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(8)
    for (m, n, k) in [(60, 300, 5), (700, 90, 12), (129, 129, 7)]:
        A = rs.standard_normal((m, 4)) @ rs.standard_normal((4, n)) * 3 + rs.standard_normal((m, n))
        fit = mfb.svd_pca(A, k)
        Wr, Hr = P.svd_pca(A, k)
        W, H = align_signs(fit.fit.W, fit.fit.H, Wr)
        assert rel_fro(W, Wr) < 1e-8, ((m, n, k), rel_fro(W, Wr))
        assert rel_fro(H, Hr) < 1e-8, ((m, n, k), rel_fro(H, Hr))
        assert np.abs(W.T @ W - np.eye(k)).max() < 1e-10


def test_nipals_pca_yeast_k10(yeast):
    # BASELINE.json configs[1]: nipals_pca k=10 on yeast_benchmark (scale = TRUE default, center = FALSE, tol = 1e-6)
    import mfbiclust_b200 as mfb
    flips = 0
    for i, A in enumerate(yeast):
        k = min(10, A.shape[1])
        res = mfb.nipals_pca(A, k, cleanParam=0, center=False)
        W, H = res["genericFit"].fit.W, res["genericFit"].fit.H
        Wr, Hr, gr, gc = P.nipals_pca(A, k)
        assert np.array_equal(res["indexRemaining"][0], gr) and np.array_equal(res["indexRemaining"][1], gc)
        assert rel_fro(W, Wr) < 1e-8, (i, rel_fro(W, Wr))
        assert rel_fro(H, Hr) < 1e-8, (i, rel_fro(H, Hr))
        flips += _membership_check(W, H, Wr, Hr)
    assert flips == 0


@pytest.mark.parametrize("center,scale,gs", [(True, True, True), (True, False, True), (False, True, False)])
def test_nipals_options(center, scale, gs):
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(21)
    x = rs.standard_normal((500, 3)) @ rs.standard_normal((3, 40)) * 2 + rs.standard_normal((500, 40)) + 1.5
    got = mfb.nipals(x, 4, center=center, scale=scale, gramschmidt=gs, tol=1e-9)
    ref = P.nipals(x, 4, center=center, scale=scale, gramschmidt=gs, tol=1e-9)
    assert np.array_equal(got["iter"], ref["iters"])
    assert rel_fro(got["scores"], ref["scores"]) < 1e-8
    assert rel_fro(got["loadings"], ref["loadings"]) < 1e-8
    assert np.allclose(got["eig"], ref["eig"], rtol=1e-10)


def test_addstrat_svd_and_als(simdata):
    import mfbiclust_b200 as mfb
    A = simdata["plaid"]
    bce = mfb.addStrat(mfb.BiclusterExperiment(A), 5, method="svd-pca", silent=True)
    assert "SVD-PCA | Otsu | 5" in bce.strategies
    bce = mfb.addStrat(bce, 5, method="als-nmf", silent=True, reps=1, rng=mfb.HostRng(3))
    s = bce.strategies["ALS-NMF | Otsu | 5"]
    assert s.biclust.RowxNumber.shape == (80, 5) and s.biclust.NumberxCol.shape == (5, 80)
    assert (s.factors.fit.W >= 0).all() and (s.factors.fit.H >= 0).all()


def test_nipals_missing_values():
    # nipals' has.na branch (NaN = NA), as addStrat routes any matrix with NA (R/addStrat.R:67-84)
    import mfbiclust_b200 as mfb
    rs = np.random.default_rng(33)
    x = rs.standard_normal((400, 3)) @ rs.standard_normal((3, 30)) * 2 + rs.standard_normal((400, 30)) + 1.0
    x[rs.random(x.shape) < 0.05] = np.nan
    for center, scale in ((False, True), (True, False)):
        got = mfb.nipals(x, 3, center=center, scale=scale, tol=1e-9)
        ref = P.nipals(x, 3, center=center, scale=scale, tol=1e-9)
        assert np.array_equal(got["iter"], ref["iters"])
        assert rel_fro(got["scores"], ref["scores"]) < 1e-8
        assert rel_fro(got["loadings"], ref["loadings"]) < 1e-8
    with pytest.warns(UserWarning, match="NIPALS-PCA algorithm must be used"):
        bce = mfb.addStrat(mfb.BiclusterExperiment(x), 2, method="svd-pca", silent=True)
    assert "NIPALS-PCA | Otsu | 2" in bce.strategies
