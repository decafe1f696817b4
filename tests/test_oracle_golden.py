"""Pin the CPU oracle against the reference's own artefacts:
``tests/testdata/ref_bces.rda`` (via tests/golden/ref_bces.json) and the
rendered vignette (tests/golden/vignette.json)."""
import numpy as np
import pytest

from oracle import otsu as O
from oracle.als_nmf import als_nmf
from oracle.fcnnls import fcnnls, nnls_reference
from oracle.helpers import pseudovalues
from oracle.pca import nipals_pca, svd_pca
from oracle.r_rng import RRng
from oracle.snmf import nmf_snmf, snmf


def test_runif_reproduces_fixture_input(ref_bces):
    # tests/testthat/test_BiclusterExperiment.R:5-7
    rng = RRng(1261534)
    inp = rng.runif(36).reshape((6, 6), order="F")
    inp[0:2, 0:2] = 20
    assert np.array_equal(inp, np.array(ref_bces["input"]))


def test_fixture_als_w_halfstep_and_thresholds(ref_bces):
    g = ref_bces["bce.als_nmf"]
    A = np.array(ref_bces["input"])
    W, H = np.array(g["W"]), np.array(g["H"])
    # the stored pair is internally consistent: W = NNLS(t(H), t(A))  (R/bicluster.R:150)
    Wt, _ = fcnnls(H.T, A.T)
    assert np.abs(Wt.T - W).max() < 1e-15
    ft, st, rn, nc = O.threshold_helper(W, H)
    assert np.array_equal(ft, np.array(g["featureThresh"]))
    assert np.array_equal(st, np.array(g["sampleThresh"]))
    assert np.array_equal(rn.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(nc.astype(int), np.array(g["NumberxCol"]))


def test_fixture_svd_pca(ref_bces):
    g = ref_bces["bce.svd_pca"]
    A = np.array(ref_bces["input"])
    W, H = svd_pca(A, 1)
    Wg, Hg = np.array(g["W"]), np.array(g["H"])
    if np.sign(W[0, 0]) != np.sign(Wg[0, 0]):   # LAPACK sign convention
        W, H = -W, -H
    assert np.abs(W - Wg).max() < 1e-14
    assert np.abs(H - Hg).max() < 1e-13
    ft, st, rn, nc = O.threshold_helper(Wg, Hg)
    assert np.array_equal(ft, np.array(g["featureThresh"]))
    assert np.array_equal(st, np.array(g["sampleThresh"]))
    assert np.array_equal(rn.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(nc.astype(int), np.array(g["NumberxCol"]))


def test_fixture_nipals_pca(ref_bces):
    g = ref_bces["bce.nipals_pca"]
    A = np.array(ref_bces["input"])
    W, H, gr, gc = nipals_pca(A, 1)
    assert gr.all() and gc.all()
    assert np.abs(W - np.array(g["W"])).max() < 1e-15
    assert np.abs(H - np.array(g["H"])).max() < 1e-15
    ft, st, rn, nc = O.threshold_helper(np.array(g["W"]), np.array(g["H"]))
    assert np.array_equal(ft, np.array(g["featureThresh"]))
    assert np.array_equal(st, np.array(g["sampleThresh"]))
    assert np.array_equal(rn.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(nc.astype(int), np.array(g["NumberxCol"]))


@pytest.fixture(scope="module")
def vignette_run(yeast):
    # vignettes/my-vignette.html: set.seed(12345); addStrat(bce, k = 3, method = "als-nmf")
    A = pseudovalues(yeast[0])
    rng = RRng(12345)
    traces = []
    res = als_nmf(A, 3, reps=4, rng=rng, traces=traces)
    return A, res, traces


def test_vignette_traces_and_winner(vignette, vignette_run):
    _, res, traces = vignette_run
    for rep in range(4):
        tr = {t[0]: t[1:] for t in traces[rep]}
        for (it, resid, dw, dh) in vignette["traces"][rep]:
            got = tr[it]
            assert f"{got[0]:f}" == f"{resid:f}"
            assert f"{got[1]:f}" == f"{dw:f}"
            assert f"{got[2]:f}" == f"{dh:f}"
        assert traces[rep][-1][0] == 21
    assert res["best"] == 1


def test_vignette_factors_and_memberships(vignette, vignette_run, yeast_names):
    _, res, _ = vignette_run
    W, H = res["W"], res["H"]
    for i in range(3):
        want = np.array(vignette["sampleFactor"][f"Bicluster.{i + 1}"])
        # printed with 7 significant digits
        assert np.allclose(H[i], want, rtol=2e-7, atol=5e-6)
    assert H[2, 0] == 0.0
    rows, _ = yeast_names
    for row in vignette["featureFactor_head"]:
        i = rows.index(row[0])
        assert np.allclose(W[i], row[1:], rtol=0, atol=6e-11)
    ft, st, rn, nc = O.threshold_helper(W, H)
    for i in range(3):
        assert list(nc[i]) == vignette["clusteredSamples"][f"Bicluster.{i + 1}"]
    for row in vignette["clusteredFeatures_head"]:
        assert list(rn[rows.index(row[0])]) == row[1:]
    assert int(rn[:, 0].sum()) == vignette["b1_feature_count"]
    assert int(rn[:, 1].sum()) == vignette["b2_feature_count"]
    assert [rows[i] for i in np.flatnonzero(rn[:, 0])[:80]] == vignette["b1_feature_names_head"]
    assert [rows[i] for i in np.flatnonzero(rn[:, 1])[:80]] == vignette["b2_feature_names_head"]


def test_fcnnls_matches_per_column_nnls():
    rs = np.random.default_rng(7)
    C = rs.random((60, 8))
    A = rs.standard_normal((60, 40)) + 0.3
    K, _ = fcnnls(C, A)
    G, R = C.T @ C, C.T @ A
    for j in range(A.shape[1]):
        x = nnls_reference(G, R[:, j])
        assert np.abs(x - K[:, j]).max() < 1e-12
    assert (K == 0).any() and (K >= 0).all()


def test_fixture_snmf_from_the_stored_seed(ref_bces):
    """bce.snmf (test_BiclusterExperiment.R:16,43-54): the NMFfit stores the .Random.seed the run started from, beta /
    eta, the factors, the iteration count and the final objective -- NMF::nmf(method = "snmf/r") restated in
    oracle/snmf.py reproduces all of them, which pins the algorithm (regularisers, restart / convergence protocol,
    "random" seeding order) and mfBiclust's wrapper (beta = eta = mean(A), R/bicluster.R:351-352)."""
    g = ref_bces["bce.snmf"]
    A = np.array(ref_bces["input"])
    assert g["beta"] == A.mean() and g["eta"] == A.mean()
    res = snmf(A, 1, rng=RRng.from_random_seed(g["random_seed"]))
    assert res["attempts"] == 1 and res["restarts"] == 0
    assert res["iters"] == g["iterations"] == 50
    assert np.abs(res["W"] - np.array(g["W"])).max() < 1e-14
    assert np.abs(res["H"] - np.array(g["H"])).max() < 1e-14
    assert abs(res["objective"] - g["residual"]) < 1e-12 * g["residual"]
    ft, st, rn, nc = O.threshold_helper(res["W"], res["H"])
    assert np.allclose(ft, g["featureThresh"], rtol=1e-14) and np.allclose(st, g["sampleThresh"], rtol=1e-14)
    assert np.array_equal(rn.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(nc.astype(int), np.array(g["NumberxCol"]))
    # the seeding order matters at the 1e-10 level even at k = 1: coef-first draws miss the stored factors
    rng = RRng.from_random_seed(g["random_seed"])
    rng.runif(6)
    other = nmf_snmf(A, rng.runif(6).reshape(6, 1), eta=g["eta"], beta=g["beta"])
    assert np.abs(other["W"] - np.array(g["W"])).max() > 1e-12
