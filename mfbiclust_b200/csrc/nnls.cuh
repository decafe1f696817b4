// Per-problem exact NNLS on the normal equations and the k x k Gram inverse it shares.
//
// Replaces NMF::.fcnnls / .cssls (NMF 0.21.0; call sites R/bicluster.R:123,150).  The R code
// groups columns by passive set and calls solve() per group; here every column (H-step) or
// row (W-step) of the factor is one thread running the same pivot rules on its own problem
//     min_x 1/2 x'Gx - r'x,  x >= 0        (G = C'C, r = C'a)
//   * first guess: unconstrained solution xu = solve(G) %*% r, passive set P = {xu > 0}
//   * inner loop (Lawson-Hanson): while the solution on P has a negative entry, step from the
//     last feasible point d towards it until the first entry hits zero, drop that entry
//   * outer loop: gradient w = r - G x on the active set; optimal when no w > 0, else add
//     argmax w to P.
// The solve on a passive set P uses the SHARED inverse M = G^-1 (one k x k inversion per
// half-step for all columns) through the Schur complement on the ACTIVE set N = ~P:
//     x_P = xu_P - M_PN y,   y = (M_NN)^-1 xu_N,   x_N = 0
// which costs O(|N|^3 + k|N|) per pivot instead of O(|P|^3), and |N| is small in practice.
#pragma once
#include "common.cuh"

// In-place inversion of the SPD k x k matrix `a` (shared memory, leading dimension ld) by
// Gauss-Jordan without pivoting; all `nthreads` threads of the calling group participate and
// `bar()` must synchronise exactly that group.  Returns (to all threads) 0 on success, 1 if a
// pivot is not safely positive (R's solve(): "system is computationally singular").
template <typename Bar>
__device__ int gauss_jordan_inverse(double *a, int k, int ld, int tid, int nthreads, double *scratch /* >= 2 doubles */,
                                    Bar bar) {
    // R's solve() refuses the system when a pivot vanishes or rcond (1-norm) < .Machine$double.eps
    if (tid == 0) {
        double an = 0.0;
        for (int j = 0; j < k; ++j) {
            double cs = 0.0;
            for (int i = 0; i < k; ++i) cs += fabs(a[i * ld + j]);
            an = fmax(an, cs);
        }
        scratch[0] = an;
        scratch[1] = 0.0;  // singular flag
    }
    bar();
    for (int p = 0; p < k; ++p) {
        double piv = a[p * ld + p];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0) scratch[1] = 1.0;
            piv = 1.0;
        }
        bar();
        const double ipiv = 1.0 / piv;
        // eliminate column p from all other rows (uses the unscaled pivot row)
        for (int e = tid; e < k * k; e += nthreads) {
            int i = e / k, j = e - i * k;
            if (i != p && j != p) a[i * ld + j] -= a[i * ld + p] * ipiv * a[p * ld + j];
        }
        bar();
        for (int e = tid; e < k; e += nthreads) {
            if (e != p) {
                a[e * ld + p] = -a[e * ld + p] * ipiv;   // column p
                a[p * ld + e] = a[p * ld + e] * ipiv;    // row p
            }
        }
        if (tid == 0) a[p * ld + p] = ipiv;
        bar();
    }
    if (tid == 0) {
        double in = 0.0;
        for (int j = 0; j < k; ++j) {
            double cs = 0.0;
            for (int i = 0; i < k; ++i) cs += fabs(a[i * ld + j]);
            in = fmax(in, cs);
        }
        if (!isfinite(in) || scratch[0] * in * 2.220446049250313e-16 > 1.0) scratch[1] = 1.0;
    }
    bar();
    return scratch[1] != 0.0;
}

// One NNLS problem per thread.  G, M: shared memory, leading dimension KP (KP = padded rank).
// r: right-hand side (k entries used).  x: result.  Returns the number of pivots taken.
template <int KP>
__device__ int nnls_thread(const double *__restrict__ G, const double *__restrict__ M, int k, const double *r,
                           double *x) {
    double xu[KP], d[KP], y[KP];
    int idx[KP];
    double L[KP * (KP + 1) / 2];
    uint64_t P = 0;
    const uint64_t full = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += M[i * KP + j] * r[j];
        xu[i] = s;
        if (s > 0.0) P |= (1ull << i);
    }
    if (P == full) {
        for (int i = 0; i < k; ++i) x[i] = xu[i];
        return 0;
    }
    for (int i = 0; i < k; ++i) {
        x[i] = (xu[i] > 0.0) ? xu[i] : 0.0;
        d[i] = x[i];
    }
    int pivots = 0;
    const int cap = 4 * k + 8;

    auto solve_on_P = [&]() {
        // active list
        int nn = 0;
        for (int i = 0; i < k; ++i)
            if (!((P >> i) & 1ull)) idx[nn++] = i;
        if (nn == 0) {
            for (int i = 0; i < k; ++i) x[i] = xu[i];
            return;
        }
        if (nn == k) {
            for (int i = 0; i < k; ++i) x[i] = 0.0;
            return;
        }
        // Cholesky of M[N,N] (row-packed lower triangle), forward/back substitution
        for (int a = 0; a < nn; ++a) {
            for (int b = 0; b <= a; ++b) {
                double s = M[idx[a] * KP + idx[b]];
                for (int c = 0; c < b; ++c) s -= L[a * (a + 1) / 2 + c] * L[b * (b + 1) / 2 + c];
                if (a == b)
                    L[a * (a + 1) / 2 + a] = sqrt(fmax(s, 1e-300));
                else
                    L[a * (a + 1) / 2 + b] = s / L[b * (b + 1) / 2 + b];
            }
        }
        for (int a = 0; a < nn; ++a) {
            double s = xu[idx[a]];
            for (int c = 0; c < a; ++c) s -= L[a * (a + 1) / 2 + c] * y[c];
            y[a] = s / L[a * (a + 1) / 2 + a];
        }
        for (int a = nn - 1; a >= 0; --a) {
            double s = y[a];
            for (int c = a + 1; c < nn; ++c) s -= L[c * (c + 1) / 2 + a] * y[c];
            y[a] = s / L[a * (a + 1) / 2 + a];
        }
        for (int i = 0; i < k; ++i) {
            if ((P >> i) & 1ull) {
                double s = xu[i];
                for (int a = 0; a < nn; ++a) s -= M[i * KP + idx[a]] * y[a];
                x[i] = s;
            } else {
                x[i] = 0.0;
            }
        }
    };

    while (pivots < cap) {
        solve_on_P();
        // inner loop: restore feasibility
        while (pivots < cap) {
            double amin = INFINITY;
            int imin = -1;
            for (int i = 0; i < k; ++i) {
                if (((P >> i) & 1ull) && x[i] < 0.0) {
                    double al = d[i] / (d[i] - x[i]);
                    if (al < amin) {
                        amin = al;
                        imin = i;
                    }
                }
            }
            if (imin < 0) break;
            for (int i = 0; i < k; ++i) d[i] = d[i] - amin * (d[i] - x[i]);
            d[imin] = 0.0;
            P &= ~(1ull << imin);
            ++pivots;
            solve_on_P();
        }
        // optimality on the active set: w = r - G x
        double wmax = 0.0;
        int imax = -1;
        for (int i = 0; i < k; ++i) {
            if (!((P >> i) & 1ull)) {
                double w = r[i];
                for (int j = 0; j < k; ++j) w -= G[i * KP + j] * x[j];
                if (w > wmax) {
                    wmax = w;
                    imax = i;
                }
            }
        }
        if (imax < 0) break;
        P |= (1ull << imax);
        for (int i = 0; i < k; ++i) d[i] = x[i];
        ++pivots;
    }
    // guard: whatever the path, never return a negative entry
    for (int i = 0; i < k; ++i)
        if (!(x[i] > 0.0)) x[i] = 0.0;
    return pivots;
}


// ---------------------------------------------------------------------------
// Warp-level in-place inversion of the SPD k x k matrix `a` (shared memory, ld, k <= 32): lane i owns row i in
// registers, Gauss-Jordan without pivoting, pivot rows broadcast by shuffles (no barriers).  Same acceptance rule
// as gauss_jordan_inverse.  Must be called by one full warp; returns 1 (to all lanes) if singular.
// ---------------------------------------------------------------------------
template <int KP>
__device__ int warp_gauss_jordan_inverse(double *a, int k, int ld, int lane) {
    static_assert(KP <= 32, "one lane per row");
    const unsigned FULL = 0xffffffffu;
    double r[KP];
    double an = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        r[j] = (lane < k && j < k) ? a[lane * ld + j] : ((lane == j) ? 1.0 : 0.0);   // identity padding
        if (lane < k && j < k) an += fabs(r[j]);
    }
    an = warp_max(an);                       // 1-norm (symmetric: max absolute row sum)
    int sing = 0;
#pragma unroll
    for (int p = 0; p < KP; ++p) {
        if (p < k) {
            double piv = __shfl_sync(FULL, r[p], p);
            if (!(piv > 0.0) || !isfinite(piv)) { sing = 1; piv = 1.0; }
            const double ipiv = 1.0 / piv;
            const double f = r[p] * ipiv;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const double rp = __shfl_sync(FULL, r[j], p);
                if (j != p) r[j] = (lane == p) ? rp * ipiv : r[j] - f * rp;
            }
            r[p] = (lane == p) ? ipiv : -f;
        }
    }
    double in = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        if (lane < k && j < k) {
            in += fabs(r[j]);
            a[lane * ld + j] = r[j];
        }
    }
    in = warp_max(in);
    if (!isfinite(in) || an * in * 2.220446049250313e-16 > 1.0) sing = 1;
    return sing;
}

// ---------------------------------------------------------------------------
// Warp-cooperative NNLS (rank padded to KP <= 32).
//
// Same pivot rules as nnls_thread / one column of .fcnnls, but carried as a principal-pivoting
// tableau  z_B = q + T z_NB  (basic variables: x_i on the passive set S, the gradient
// w_i = (r - Gx)_i on the active set; start S = all, q = xu = G^-1 r, T = -G^-1), so that one
// exchange x_i <-> w_i is a rank-1 update of a KP x KP matrix that lives in REGISTERS: lane i of
// an LP-lane group owns row i.  Both the primal values (q on S) and the KKT gradient (q on ~S) come
// for free after every pivot: no refactorisation, no local memory, no per-pivot division chain
// beyond one reciprocal.  LP = 16 for KP <= 16 (two problems per warp), 32 otherwise.
//
// xs: shared [KP][XS] -- right-hand sides r on entry, solutions x on exit (column = problem).  The warp solves
// the problems [first, first + count) of xs, 32 / LP at a time; problems >= nvalid are skipped (x = 0).
// sM: shared G^-1 (ld KP, symmetric).  Returns this lane's partial of sum_i x_i r_i.
// ---------------------------------------------------------------------------
template <int KP, int XS>
__device__ double nnls_lanes(const double *__restrict__ sM, int k, double *xs, int first, int count, int nvalid,
                             int lane) {   // (each lane leaves x_i of its problem in xs; see the callers)
    constexpr int LP = (KP <= 16) ? 16 : 32;
    constexpr int PPW = 32 / LP;
    const unsigned FULL = 0xffffffffu;
    const int sub = lane / LP, i = lane % LP, base = sub * LP;
    const bool rowvalid = i < k;
    const unsigned gmask = (LP == 32) ? 0xffffffffu : (0xffffu << base);
    double t1 = 0.0;
    for (int rd = 0; rd < count; rd += PPW) {
        const int unit = first + rd + sub;
        const bool uvalid = (rd + sub < count) && (unit < nvalid);
        if (!__any_sync(FULL, uvalid)) {
            if (rowvalid && rd + sub < count) xs[i * XS + unit] = 0.0;
            continue;
        }
        const double ri = (rowvalid && uvalid) ? xs[i * XS + unit] : 0.0;
        double T[KP];
        double qv = 0.0;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const double rj = __shfl_sync(FULL, ri, base + (j % LP));
            const double mij = (rowvalid && j < k) ? sM[j * KP + i] : 0.0;
            T[j] = -mij;
            qv += mij * rj;
        }
        bool inS = rowvalid;
        double dv = qv > 0.0 ? qv : 0.0;
        const unsigned negs = __ballot_sync(FULL, rowvalid && !(qv > 0.0));
        unsigned pend = uvalid ? ((negs & gmask) >> base) : 0u;
        bool done = !uvalid;
        int iters = 0;
        const int cap = 6 * k + 16;
        while (true) {
            int piv = -1;
            if (!done) {
                if (pend) {
                    piv = __ffs(pend) - 1;
                    pend &= pend - 1;
                } else {
                    const double x = inS ? qv : 0.0;
                    const bool isneg = inS && x < 0.0;
                    if (__ballot_sync(gmask, isneg) != 0u) {
                        // Lawson-Hanson step length: min over passive entries that went negative
                        double al = isneg ? dv / (dv - x) : INFINITY;
                        int ai = i;
#pragma unroll
                        for (int o = LP / 2; o > 0; o >>= 1) {
                            const double oa = __shfl_xor_sync(gmask, al, o);
                            const int oi = __shfl_xor_sync(gmask, ai, o);
                            if (oa < al || (oa == al && oi < ai)) { al = oa; ai = oi; }
                        }
                        dv = dv - al * (dv - x);
                        if (i == ai) dv = 0.0;
                        piv = ai;
                    } else {
                        const bool ispos = (!inS && rowvalid) && qv > 0.0;
                        if (__ballot_sync(gmask, ispos) != 0u) {
                            double w = ispos ? qv : -1.0;
                            int wi = i;
#pragma unroll
                            for (int o = LP / 2; o > 0; o >>= 1) {
                                const double ow = __shfl_xor_sync(gmask, w, o);
                                const int oi = __shfl_xor_sync(gmask, wi, o);
                                if (ow > w || (ow == w && oi < wi)) { w = ow; wi = oi; }
                            }
                            dv = x;
                            piv = wi;
                        } else {
                            done = true;
                        }
                    }
                }
                if (++iters > cap) { done = true; piv = -1; }
            }
            if (!__any_sync(FULL, piv >= 0)) break;
            // exchange x_piv <-> w_piv: every row becomes a * row + b * (pivot row), then column piv is patched;
            // a problem that does not pivot in this pass runs with (a, b) = (1, 0)
            const int pq = piv >= 0 ? piv : 0;
            const int src = base + pq;
            double tiq = 0.0;
#pragma unroll
            for (int j = 0; j < KP; ++j) tiq = (j == pq) ? T[j] : tiq;
            const double qq = __shfl_sync(FULL, qv, src);
            const double pp = __shfl_sync(FULL, tiq, src);
            const double ip = 1.0 / pp;
            const bool act = piv >= 0, isrow = act && (i == piv);
            const double f = tiq * ip;
            const double a = isrow ? 0.0 : 1.0;
            const double b = act ? (isrow ? -ip : -f) : 0.0;
            const double fix = isrow ? ip : f;
            const int pfix = act ? pq : -1;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const double rq = __shfl_sync(FULL, T[j], src);
                const double t = fma(b, rq, a * T[j]);
                T[j] = (j == pfix) ? fix : t;
            }
            qv = fma(b, qq, a * qv);
            if (isrow) inS = !inS;
        }
        const double x = (inS && rowvalid && uvalid && qv > 0.0) ? qv : 0.0;
        t1 += x * ri;
        if (rowvalid && rd + sub < count) xs[i * XS + unit] = x;
    }
    return t1;
}
