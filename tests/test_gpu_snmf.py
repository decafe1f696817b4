"""snmf (R/bicluster.R:341-441) on the device against the oracle's restatement of NMF::nmf(method = "snmf/r") and
against the reference's own artefact (bce.snmf of tests/testdata/ref_bces.rda: start seed, factors, iteration count,
objective, memberships; test_BiclusterExperiment.R:43-54).  Tolerance: factors within 1e-8 (relative Frobenius), the
FP64 bar of BASELINE.json; iteration / restart counts and memberships exact."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def planted(m, n, k, seed):
    r = np.random.default_rng(seed)
    W = np.zeros((m, k))
    H = np.zeros((k, n))
    for c in range(k):
        W[r.choice(m, m // (k + 1), replace=False), c] = r.uniform(1, 2, m // (k + 1))
        H[c, r.choice(n, n // (k + 1), replace=False)] = r.uniform(1, 2, n // (k + 1))
    return np.asfortranarray(W @ H + 0.1 * r.random((m, n)))


def relfro(X, Y):
    return np.linalg.norm(X - Y) / max(np.linalg.norm(Y), 1e-300)


def test_snmf_fixture_from_the_stored_seed(ref_bces):
    from mfbiclust_b200 import backends as B
    from oracle.r_rng import RRng
    g = ref_bces["bce.snmf"]
    A = np.array(ref_bces["input"])
    fit = B.snmf(A, 1, verbose=False, rng=RRng.from_random_seed(g["random_seed"]))
    assert fit.method == "snmf" and fit.iters == g["iterations"] == 50 and fit.attempts == 1
    assert fit.parameters == dict(beta=g["beta"], eta=g["eta"])
    assert relfro(fit.fit.W, np.array(g["W"])) < 1e-12 and relfro(fit.fit.H, np.array(g["H"])) < 1e-12
    assert abs(fit.objective - g["residual"]) < 1e-10 * g["residual"]
    # the whole strategy, as tests/testthat/test_BiclusterExperiment.R:43-54 compares it
    bce = B.addStrat(B.BiclusterExperiment(A), 1, method="snmf", silent=True, rng=RRng.from_random_seed(g["random_seed"]))
    (name, bcs), = bce.strategies.items()
    assert name == g["name"] == "SNMF | Otsu | 1"
    assert np.array_equal(bcs.biclust.RowxNumber.astype(int), np.array(g["RowxNumber"]))
    assert np.array_equal(bcs.biclust.NumberxCol.astype(int), np.array(g["NumberxCol"]))
    assert np.allclose(bcs.featureThresh, g["featureThresh"], rtol=1e-10)
    assert np.allclose(bcs.sampleThresh, g["sampleThresh"], rtol=1e-10)


@pytest.mark.parametrize("m,n,k,max_iter", [(60, 40, 4, 100), (300, 200, 5, 100), (500, 260, 16, 60), (260, 300, 20, 40),
                                            (333, 129, 32, 25), (130, 257, 9, 150)])
def test_snmf_run_equals_the_oracle(m, n, k, max_iter):
    """Same start, same iteration budget (with 2 + budget / 5 convergence tests on the way): same factors."""
    from mfbiclust_b200 import backends as B
    from oracle.snmf import nmf_snmf
    A = planted(m, n, k, 3)
    W0 = np.random.default_rng(1).random((m, k))
    ref = nmf_snmf(A, W0, eta=A.mean(), beta=A.mean(), max_iter=max_iter)
    out = B.nmf_snmf(A, k, W0=W0, beta=A.mean(), eta=A.mean(), maxIter=max_iter)
    assert out["iters"] == ref["iters"] == max_iter and out["restarts"] == ref["restarts"] == 0 and not out["converged"]
    assert relfro(out["W"], ref["W"]) < 1e-8 and relfro(out["H"], ref["H"]) < 1e-8
    assert np.array_equal(out["W"] == 0, ref["W"] == 0) and np.array_equal(out["H"] == 0, ref["H"] == 0)
    assert abs(out["objective"] - ref["objective"]) < 1e-9 * ref["objective"]


@pytest.mark.parametrize("m,n,k", [(60, 40, 4), (300, 200, 5)])
def test_snmf_stops_where_the_oracle_stops_up_to_rounding_noise(m, n, k):
    """NMF's stopping rule divides the KKT residual by the NUMBER of its non-zero entries, and the W block of that
    residual is pure rounding noise (W has just been solved exactly): hundreds of its entries are or are not exactly
    zero depending on the summation order of whoever evaluates it (observed: the count moves between 700 and 1100 from
    test to test on one run).  The iteration at which `erravg <= eps_conv * erravg1` first holds is therefore not a
    function of the inputs alone -- not in R either (it depends on the BLAS) -- and two implementations agree on it
    only to a few tests.  The factors at a fixed iteration count are equal (test above); here: same stop within 5 % of
    the run, and a fixed point equal to the oracle's at the accuracy the rule asks for."""
    from mfbiclust_b200 import backends as B
    from oracle.snmf import nmf_snmf
    A = planted(m, n, k, 3)
    W0 = np.random.default_rng(1).random((m, k))
    ref = nmf_snmf(A, W0, eta=A.mean(), beta=A.mean())
    out = B.nmf_snmf(A, k, W0=W0, beta=A.mean(), eta=A.mean())
    assert out["converged"] and out["iters"] % 5 == 0
    assert abs(out["iters"] - ref["iters"]) <= max(10, 0.05 * ref["iters"])
    assert relfro(out["W"], ref["W"]) < 1e-3 and relfro(out["H"], ref["H"]) < 1e-3
    again = nmf_snmf(A, W0, eta=A.mean(), beta=A.mean(), max_iter=out["iters"], eps_conv=0.0)
    assert relfro(out["W"], again["W"]) < 1e-8 and relfro(out["H"], again["H"]) < 1e-8


def test_snmf_eta_default_is_max_and_iteration_budget_splits():
    """eta < 0 means max(A) (NMF's default); a run split into several mfb_snmf_iterate calls (any budget, on or off the
    test iterations) is the same run."""
    from mfbiclust_b200 import _cabi, backends as B
    from oracle.snmf import nmf_snmf
    A = planted(120, 90, 3, 11)
    W0 = np.random.default_rng(2).random((120, 3))
    ref = nmf_snmf(A, W0, eta=-1.0, beta=0.01, max_iter=37)
    out = B.nmf_snmf(A, 3, W0=W0, beta=0.01, eta=-1.0, maxIter=37)
    assert out["iters"] == ref["iters"] == 37 and not out["converged"]
    assert relfro(out["W"], ref["W"]) < 1e-8 and relfro(out["H"], ref["H"]) < 1e-8
    assert abs(out["objective"] - ref["objective"]) < 1e-9 * ref["objective"]
    ctx = _cabi.default_context()
    Ad = _cabi.Matrix(ctx, host=A)
    sess = B.SnmfSession(Ad, 3, W0, 0.01, -1.0)
    total = 0
    for budget in (1, 3, 1, 7, 5, 20):
        st, total, conv, obj = sess.iterate(budget)
        assert st == 0 and not conv
    assert total == 37
    W, H = sess.result()
    sess.close()
    Ad.free()
    assert np.array_equal(W, out["W"]) and np.array_equal(H, out["H"]) and obj == out["objective"]


def test_snmf_restart_protocol_follows_the_oracle():
    """A beta that empties a row of H: .nmf_snmf redraws W (the membership state and the first KKT residual are reset,
    the iteration counter keeps counting) -- six restarts, then a converged run."""
    from mfbiclust_b200 import backends as B
    from oracle.snmf import nmf_snmf
    A = planted(60, 40, 4, 1)
    rr = np.random.default_rng(8)
    draws = [rr.random((60, 4)) for _ in range(12)]
    it = iter(draws)
    ref = nmf_snmf(A, next(it), eta=A.mean(), beta=34 * A.mean(), redraw=lambda: next(it), max_iter=40)
    assert ref["restarts"] == 6 and ref["warning"] is None and ref["iters"] == 40

    class Replay:
        def __init__(self):
            self.it = iter(draws[1:])

        def runif(self, cnt):
            return next(self.it).ravel(order="F")
    out = B.nmf_snmf(A, 4, W0=draws[0], beta=34 * A.mean(), eta=A.mean(), rng=Replay(), maxIter=40)
    assert out["iters"] == 40 and out["restarts"] == 6 and not out["converged"]
    assert relfro(out["W"], ref["W"]) < 1e-8 and relfro(out["H"], ref["H"]) < 1e-8
    # to convergence (the stop itself is rounding-noise dependent, see above)
    it = iter(draws)
    full = nmf_snmf(A, next(it), eta=A.mean(), beta=34 * A.mean(), redraw=lambda: next(it))
    out = B.nmf_snmf(A, 4, W0=draws[0], beta=34 * A.mean(), eta=A.mean(), rng=Replay())
    assert out["converged"] and out["restarts"] == 6 and abs(out["iters"] - full["iters"]) <= 10


def test_snmf_wrapper_halves_beta_like_the_reference():
    """R/bicluster.R:380-390: NMF's too-many-restarts warning halves beta and the run starts over from the same seed."""
    from mfbiclust_b200 import backends as B
    from oracle.snmf import snmf as oracle_snmf
    A = planted(60, 40, 4, 1)

    def starts(attempt):
        rr = np.random.default_rng(8)          # every attempt re-seeds from the same .Random.seed (:366)
        return iter([rr.random((60, 4)) for _ in range(12)])
    ref = oracle_snmf(A, 4, beta=400 * A.mean(), eta=A.mean(), starts=starts, max_iter=60)
    assert ref["attempts"] == 3 and ref["iters"] == 60
    fit = B.snmf(A, 4, verbose=False, beta=400 * A.mean(), eta=A.mean(), starts=starts, maxIter=60)
    assert fit.attempts == ref["attempts"] and fit.parameters["beta"] == ref["beta"] and fit.iters == ref["iters"]
    assert relfro(fit.fit.W, ref["W"]) < 1e-8 and relfro(fit.fit.H, ref["H"]) < 1e-8


def test_snmf_singular_system_is_reported_and_eta_is_halved():
    """eta = 0 on a rank-one matrix: HH' is singular, solve() inside .fcnnls fails ("computationally singular") -- the
    device reports MFB_SINGULAR where the oracle raises."""
    from mfbiclust_b200 import _cabi, backends as B
    from oracle.fcnnls import SingularError
    from oracle.snmf import nmf_snmf
    rs = np.random.default_rng(5)
    A = np.asfortranarray(np.outer(rs.random(30) + 1, rs.random(20) + 1))
    W0 = rs.random((30, 3))
    with pytest.raises(SingularError):
        nmf_snmf(A, W0, eta=0.0, beta=1e-3)
    with pytest.raises(_cabi.MfbError) as ei:
        B.nmf_snmf(A, 3, W0=W0, beta=1e-3, eta=0.0)
    assert ei.value.status == _cabi.MFB_SINGULAR


def test_snmf_large_rank_properties():
    """k = 40 (block-pivoting NNLS; the reference's .fcnnls does not terminate on this shape -- cumulative sweep cap --
    so there is no oracle run): both factors are KKT points of their regularised half-steps, and the objective equals
    its definition."""
    from mfbiclust_b200 import backends as B
    m, n, k = 400, 300, 40
    A = planted(m, n, k, 3)
    beta = eta = float(A.mean())
    out = B.nmf_snmf(A, k, W0=np.random.default_rng(1).random((m, k)), beta=beta, eta=eta, maxIter=10)
    W, H = out["W"], out["H"]
    assert out["iters"] == 10 and (W >= 0).all() and (H >= 0).all()
    gW = W @ (H @ H.T + eta * eta * np.eye(k)) - A @ H.T          # W-step optimality of the final W
    scale = np.abs(A @ H.T).max()
    assert np.abs(gW[W > 0]).max() < 1e-9 * scale and gW[W == 0].min() > -1e-9 * scale
    obj = 0.5 * (((A - W @ H) ** 2).sum() + eta * (W ** 2).sum() + beta * (H.sum(axis=0) ** 2).sum())
    assert abs(out["objective"] - obj) < 1e-9 * obj


def test_snmf_strategy_shifts_negative_input():
    from mfbiclust_b200 import backends as B
    A = planted(80, 50, 2, 4) - 0.05
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        bcs = B.BiclusterStrategy(A, 2, method="snmf", verbose=False, rng=B.HostRng(3))
    assert bcs.name == "SNMF | Otsu | 2" and bcs.biclust.RowxNumber.shape == (80, 2)


def test_snmf_full_size_matches_oracle():
    """The bench workload (20000 x 5000, k = 16; eta = mean(A), beta = mean(A) / 256 -- the first beta of mfBiclust's
    halving sequence that keeps every row of H alive on it): two literal oracle iterations, with the convergence test
    of iteration 1 on the way, against the device -- factors to 1e-8, identical zero patterns, same objective."""
    from mfbiclust_b200 import backends as B
    from oracle.snmf import nmf_snmf
    from bench import make_workload
    A, _, _ = make_workload()
    m, n = A.shape
    W0 = np.random.default_rng(43).random((m, 16))
    mean = float(A.mean())
    ref = nmf_snmf(A, W0, eta=mean, beta=mean / 256, max_iter=2)
    out = B.nmf_snmf(A, 16, W0=W0, beta=mean / 256, eta=mean, maxIter=2)
    assert out["iters"] == ref["iters"] == 2 and out["restarts"] == ref["restarts"] == 0
    assert relfro(out["W"], ref["W"]) < 1e-8 and relfro(out["H"], ref["H"]) < 1e-8
    assert np.array_equal(out["W"] == 0, ref["W"] == 0) and np.array_equal(out["H"] == 0, ref["H"] == 0)
    assert abs(out["objective"] - ref["objective"]) < 1e-9 * ref["objective"]
