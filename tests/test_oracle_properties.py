"""Independent mathematical checks of the oracle paths that no artefact of the reference pins numerically (DESIGN.md
section 4: nipals with more than one component, the missing-value branch, Otsu beyond the fixture's six thresholds).
They do not replace a golden vector -- they show that the restatement computes what the published algorithm defines."""
import numpy as np

from oracle import otsu as O
from oracle import pca as P


def _planted(m, n, r, seed, noise=0.05):
    rs = np.random.default_rng(seed)
    U = rs.standard_normal((m, r)) * np.array([9.0, 5.0, 3.0, 2.0, 1.5][:r])[None, :]
    V = rs.standard_normal((n, r))
    return U @ V.T + noise * rs.standard_normal((m, n))


def test_nipals_components_are_the_leading_singular_triplets():
    """Deflated power iterations with Gram-Schmidt converge to the top singular vectors of the pre-processed matrix
    (nipals' tol is absolute on the un-normalised scores, so the agreement is ~sqrt(tol) / gap, not machine precision)."""
    x = _planted(300, 40, 4, 1)
    res = P.nipals(x, 4, center=True, scale=True, tol=1e-12, maxiter=2000)
    z = x - x.mean(axis=0, keepdims=True)
    z = z / z.std(axis=0, ddof=1, keepdims=True)
    U, s, Vt = np.linalg.svd(z, full_matrices=False)
    for h in range(4):
        sgn = np.sign(res["loadings"][:, h] @ Vt[h])
        assert np.allclose(sgn * res["loadings"][:, h], Vt[h], atol=1e-5), h
        assert np.allclose(sgn * res["scores"][:, h], U[:, h], atol=1e-5), h
        assert abs(res["eig"][h] - s[h]) < 1e-6 * s[0], h
    # orthonormal scores and loadings (the purpose of the Gram-Schmidt steps)
    assert np.allclose(res["scores"].T @ res["scores"], np.eye(4), atol=1e-9)
    assert np.allclose(res["loadings"].T @ res["loadings"], np.eye(4), atol=1e-9)


def test_nipals_missing_value_branch_recovers_an_exactly_low_rank_matrix():
    """On a rank-1 matrix with a few cells missing the has.na branch (sums over observed cells only, per-column /
    per-row denominators) recovers the complete-data factors: t p' reproduces every cell, also the missing ones."""
    rs = np.random.default_rng(2)
    t = rs.standard_normal(60) + 3.0
    p = rs.standard_normal(12) + 2.0
    full = np.outer(t, p)
    x = full.copy()
    miss = rs.random(x.shape) < 0.08
    miss[:, 0] = False
    miss[0, :] = False
    x[miss] = np.nan
    res = P.nipals(x, 1, center=False, scale=False, tol=1e-20, maxiter=500)
    rec = np.outer(res["scores"][:, 0] * res["eig"][0], res["loadings"][:, 0])
    assert np.allclose(rec, full, rtol=1e-9, atol=1e-9)
    # and with no missing cell the branch is never taken: identical to the complete-data path
    a = P.nipals(full, 1, center=False, scale=False)
    b = P.nipals(np.where(miss, full, full), 1, center=False, scale=False)
    assert np.array_equal(a["scores"], b["scores"])


def test_otsu_maximises_the_between_class_variance_of_the_histogram():
    """EBImage::otsu: the returned level is the (mean of the) bin centre(s) maximising w1 w2 (m2/w2 - m1/w1)^2, computed
    here by brute force from the definition for every split, on bimodal, skewed and tied histograms."""
    rs = np.random.default_rng(3)
    cases = [np.concatenate([rs.normal(0.25, 0.05, 500), rs.normal(0.7, 0.08, 300)]).clip(0, 1),
             rs.beta(2.0, 7.0, 1000), rs.random(64),
             np.concatenate([np.full(10, 0.1), np.full(10, 0.9)])]
    mids = (2.0 * np.arange(256) + 1.0) / 512.0
    for x in cases:
        counts = O.hist_counts_unit(x)
        assert counts.sum() == x.size
        best, arg = -1.0, []
        for i in range(256):
            w1, w2 = counts[:i + 1].sum(), counts[i:].sum()          # EBImage's classes overlap in bin i
            if w1 == 0 or w2 == 0:
                continue
            m1 = (counts[:i + 1] * mids[:i + 1]).sum() / w1
            m2 = (counts[i:] * mids[i:]).sum() / w2
            v = w1 * w2 * (m2 - m1) ** 2
            if v > best * (1 + 1e-12):
                best, arg = v, [i]
            elif abs(v - best) <= 1e-12 * best:
                arg.append(i)
        got = O.otsu_from_counts(counts)
        assert abs(got - (mids[arg[0]] + mids[arg[-1]]) / 2.0) < 1e-12, (got, arg)


def test_snmf_oracle_is_block_coordinate_descent_on_its_objective():
    """oracle/snmf.py beyond its k = 1 fixture (ref_bces.rda): each half-step of NMF's snmf/r is the exact minimiser of
    1/2 (||A - WH||^2 + eta ||W||^2 + beta sum_j (sum_c H_cj)^2) over one factor, so the objective never increases from
    one iteration to the next, and at the iteration where the run stops both KKT blocks are small against their
    first values; memberships are max.col(ties = "first")."""
    from oracle.snmf import kkt_residual, max_col_first, nmf_snmf, snmf_objective
    rs = np.random.default_rng(12)
    m, n, k = 70, 50, 4
    Wt = np.where(rs.random((m, k)) < 0.3, rs.random((m, k)) + 0.5, 0.0)
    Ht = np.where(rs.random((k, n)) < 0.3, rs.random((k, n)) + 0.5, 0.0)
    A = Wt @ Ht + 0.05 * rs.random((m, n))
    W0 = rs.random((m, k))
    beta = eta = float(A.mean())
    objs = [nmf_snmf(A, W0, eta=eta, beta=beta, max_iter=t)["objective"] for t in range(1, 13)]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(objs, objs[1:])), objs
    first = nmf_snmf(A, W0, eta=eta, beta=beta, max_iter=1)
    done = nmf_snmf(A, W0, eta=eta, beta=beta)
    assert done["iters"] % 5 == 0 and done["iters"] >= 50 and done["restarts"] == 0      # 10 stable tests, every fifth iteration
    c1, n1 = kkt_residual(A, first["W"], first["H"], eta, beta)
    c2, n2 = kkt_residual(A, done["W"], done["H"], eta, beta)
    assert c2 / n2 <= 1e-4 * c1 / n1
    assert abs(done["objective"] - snmf_objective(A, done["W"], done["H"], eta, beta)) < 1e-12 * done["objective"]
    assert list(max_col_first(np.array([[1.0, 3.0, 3.0], [2.0, 2.0, 1.0], [0.0, 0.0, 0.0]]))) == [2, 1, 1]
