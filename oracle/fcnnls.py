"""Restatement of ``NMF::.fcnnls`` / ``.cssls`` (NMF 0.21.0) and ``base::max.col``.

Test infrastructure (see ``oracle/__init__.py``).

NMF 0.21.0 is pinned by ``packrat/packrat.lock:137-139`` but NOT vendored in
the reference; the call sites are ``R/bicluster.R:123`` (``.fcnnls(W, A)``)
and ``R/bicluster.R:150`` (``.fcnnls(t(H), t(A))``).  The algorithm is
Van Benthem & Keenan's fast combinatorial NNLS (J. Chemometrics 18, 2004) as
shipped in ``NMF/R/algorithms-snmf.R``; the statement-level spec used here is
SURVEY.md App. A.1.

Two R-isms matter for bit-level parity with a ``set.seed()`` session and are
reproduced literally:

* ``solve(CtC) %*% CtA`` -- explicit LU inverse (LAPACK dgesv against the
  identity) followed by a product, with R's "computationally singular"
  error when ``rcond < .Machine$double.eps``;
* ``max.col(..., ties.method = "random")`` draws from the session RNG at
  every (near-)tie, including structural ties among leading ``-Inf`` / zero
  entries, which shifts the ``runif`` stream of later replicates.

Pinned by: the rendered vignette (traces of four replicates incl. the RNG
offsets, H with an exact zero) and the k=1 W half-step of
``tests/testdata/ref_bces.rda`` (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import numpy as np

from .r_rng import RRng

_EPS = np.finfo(np.float64).eps


class SingularError(Exception):
    """R's ``solve()`` error (exactly or computationally singular)."""


def r_solve_inverse(a: np.ndarray) -> np.ndarray:
    """``solve(a)``: inverse via dgesv, error when rcond(1-norm) < eps."""
    n = a.shape[0]
    if n == 0:
        return np.zeros((0, 0))
    try:
        inv = np.linalg.inv(a)
    except np.linalg.LinAlgError as exc:  # exact zero pivot
        raise SingularError(str(exc))
    anorm = np.abs(a).sum(axis=0).max()
    inorm = np.abs(inv).sum(axis=0).max()
    if not np.isfinite(inorm) or anorm * inorm * _EPS > 1.0:
        raise SingularError("system is computationally singular")
    return inv


def max_col(mat: np.ndarray, rng: RRng | None) -> np.ndarray:
    """``max.col(mat, ties.method="random")`` (0-based result), R_max_col."""
    nr, nc = mat.shape
    out = np.zeros(nr, dtype=np.int64)
    for r in range(nr):
        row = mat[r]
        fin = np.isfinite(row)
        large = np.abs(row[fin]).max() if fin.any() else 0.0
        tol = 1e-5 * large
        a = row[0]
        m = 0
        ntie = 1
        for c in range(1, nc):
            b = row[c]
            if b > a + tol:
                a = b
                m = c
                ntie = 1
            elif b >= a - tol:
                ntie += 1
                u = rng.unif_rand() if rng is not None else 1.0
                if ntie * u < 1.0:
                    m = c
        out[r] = m
    return out


def cssls(CtC: np.ndarray, CtA: np.ndarray, Pset: np.ndarray | None = None) -> np.ndarray:
    """``.cssls``: solve the normal equations on each column's passive set."""
    K = np.zeros_like(CtA)
    if Pset is None or Pset.size == 0 or Pset.all():
        return r_solve_inverse(CtC) @ CtA
    lVar, pRHS = Pset.shape
    weights = 2.0 ** np.arange(lVar - 1, -1, -1)
    coded = weights @ Pset.astype(np.float64)
    order = np.argsort(coded, kind="stable")
    sorted_codes = coded[order]
    breaks = np.flatnonzero(np.diff(sorted_codes) > 0) + 1
    bounds = np.concatenate(([0], breaks, [pRHS]))
    for b in range(len(bounds) - 1):
        cols = order[bounds[b]:bounds[b + 1]]
        if cols.size == 0:
            continue
        vars_ = Pset[:, cols[0]]
        if not vars_.any():
            continue
        sub = CtC[np.ix_(vars_, vars_)]
        K[np.ix_(vars_, cols)] = r_solve_inverse(sub) @ CtA[np.ix_(vars_, cols)]
    return K


def fcnnls(C: np.ndarray, A: np.ndarray, rng: RRng | None = None, eps: float = 0.0):
    """``.fcnnls(x=C, y=A)`` -> ``(K, Pset)``; ``min ||C K - A||_F, K >= 0``.

    ``rng`` is the R session RNG consumed by ``max.col``; ``None`` resolves
    every tie deterministically to the first maximum (no draws).
    """
    C = np.asarray(C, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    lVar = C.shape[1]
    pRHS = A.shape[1]
    W = np.zeros((lVar, pRHS))
    it = 0
    maxiter = 3 * lVar
    CtC = C.T @ C
    CtA = C.T @ A

    K = cssls(CtC, CtA)
    Pset = K > 0
    K[~Pset] = 0.0
    D = K.copy()
    Fset = np.flatnonzero(Pset.sum(axis=0) != lVar)

    while Fset.size > 0:
        K[:, Fset] = cssls(CtC, CtA[:, Fset], Pset[:, Fset])
        Hset = Fset[(K[:, Fset] < eps).sum(axis=0) > 0]
        if Hset.size > 0:
            nH = Hset.size
            while nH > 0 and it < maxiter:
                it += 1
                alpha = np.full((lVar, nH), np.inf)
                neg = Pset[:, Hset] & (K[:, Hset] < eps)
                if not neg.any():
                    break
                ii, jj = np.nonzero(neg)
                cols = Hset[jj]
                alpha[ii, jj] = D[ii, cols] / (D[ii, cols] - K[ii, cols])
                minIdx = max_col(-alpha.T, rng)
                alphaMin = alpha[minIdx, np.arange(nH)]
                D[:, Hset] = D[:, Hset] - alphaMin[None, :] * (D[:, Hset] - K[:, Hset])
                D[minIdx, Hset] = 0.0
                Pset[minIdx, Hset] = False
                K[:, Hset] = cssls(CtC, CtA[:, Hset], Pset[:, Hset])
                Hset = np.flatnonzero((K < eps).sum(axis=0) > 0)
                nH = Hset.size
        W[:, Fset] = CtA[:, Fset] - CtC @ K[:, Fset]
        notP_W = np.where(~Pset[:, Fset], 1.0, 0.0) * W[:, Fset]
        J = (notP_W > eps).sum(axis=0) == 0
        Fset = Fset[~J]
        if Fset.size > 0:
            notP_W = np.where(~Pset[:, Fset], 1.0, 0.0) * W[:, Fset]
            mx = max_col(notP_W.T, rng)
            Pset[mx, Fset] = True
            D[:, Fset] = K[:, Fset]
    return K, Pset


def nnls_reference(G: np.ndarray, r: np.ndarray, max_pivots: int | None = None) -> np.ndarray:
    """Per-column exact NNLS on normal equations (``min 1/2 x'Gx - r'x, x>=0``).

    Same pivot rules as one column of ``fcnnls`` (initial passive set = positive
    part of the unconstrained solution; Lawson-Hanson inner alpha step; add the
    most positive gradient), with Cholesky-free ``np.linalg.solve`` on the
    passive block.  Used by tests as an independent check of the fast
    combinatorial grouping above and of the CUDA per-column solver.
    """
    k = G.shape[0]
    x = np.linalg.solve(G, r)
    P = x > 0
    x = np.where(P, x, 0.0)
    d = x.copy()
    cap = max_pivots if max_pivots is not None else 6 * k + 10
    for _ in range(cap):
        if P.all():
            break
        x = np.zeros(k)
        if P.any():
            x[P] = np.linalg.solve(G[np.ix_(P, P)], r[P])
        while (x[P] < 0).any():
            neg = P & (x < 0)
            alpha = np.full(k, np.inf)
            alpha[neg] = d[neg] / (d[neg] - x[neg])
            i = int(np.argmin(alpha))
            d = d - alpha[i] * (d - x)
            d[i] = 0.0
            P[i] = False
            x = np.zeros(k)
            if P.any():
                x[P] = np.linalg.solve(G[np.ix_(P, P)], r[P])
        w = r - G @ x
        w[P] = 0.0
        if not (w > 0).any():
            break
        P[int(np.argmax(w))] = True
        d = x.copy()
    return x
