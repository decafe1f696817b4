// Exact NNLS on the normal equations and the k x k Gram inverse it shares.
//
// Replaces NMF::.fcnnls / .cssls (NMF 0.21.0; call sites R/bicluster.R:123,150).  The R code groups columns by
// passive set and calls solve() per group; here every column (H-step) or row (W-step) of the factor is solved by
// a group of 16 or 32 lanes running the same pivot rules on its own problem
//     min_x 1/2 x'Gx - r'x,  x >= 0        (G = C'C, r = C'a)
// against the SHARED inverse M = G^-1 (one k x k inversion per half-step for all problems): see nnls_warp below.
#pragma once
#include "common.cuh"

// In-place inversion of the SPD k x k matrix `a` (shared memory, leading dimension ld) by
// Gauss-Jordan without pivoting; all `nthreads` threads of the calling group participate and
// `bar()` must synchronise exactly that group.  Returns (to all threads) 0 on success, 1 if a
// pivot is not safely positive (R's solve(): "system is computationally singular").
template <typename Bar>
__device__ int gauss_jordan_inverse(double *a, int k, int ld, int tid, int nthreads, double *scratch /* >= 2 doubles */,
                                    Bar bar) {
    // R's solve() refuses the system when a pivot vanishes or rcond (1-norm) < .Machine$double.eps.
    // 1-norms: thread j sums column j (rows in order), the maximum is formed with an integer atomicMax on the bit
    // pattern (non-negative doubles order like their bit patterns; NaN patterns order above +inf and fail isfinite)
    unsigned long long *nrm = (unsigned long long *)scratch;
    if (tid == 0) {
        nrm[0] = 0ull;
        scratch[1] = 0.0;  // singular flag
    }
    bar();
    for (int j = tid; j < k; j += nthreads) {
        double cs = 0.0;
        for (int i = 0; i < k; ++i) cs += fabs(a[i * ld + j]);
        atomicMax(&nrm[0], (unsigned long long)__double_as_longlong(cs));
    }
    bar();
    const double an = __longlong_as_double((long long)nrm[0]);
    bar();
    if (tid == 0) nrm[0] = 0ull;
    for (int p = 0; p < k; ++p) {
        double piv = a[p * ld + p];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0) scratch[1] = 1.0;
            piv = 1.0;
        }
        const double ipiv = 1.0 / piv;
        // eliminate column p from all other rows (uses the unscaled pivot row; row p and column p are only read).
        // Warp w takes every nw-th row, a lane the columns lane, lane + 32, ...: a[i][p] is a broadcast read, the row
        // segments are conflict-free, and there is no index arithmetic beyond adds (a flat loop over the k * k entries
        // with a division per entry made this the tail of the stream kernel at k = 64: ~600 000 cycles on the side-job
        // warps).  Per entry the arithmetic is unchanged: a_ij - ((a_ip * ipiv) * a_pj).
        {
            const int w = tid >> 5, l = tid & 31, nw = nthreads >> 5;
            if (nw > 0) {
                if (w < nw) {
                    for (int j0 = 0; j0 < k; j0 += 64) {
                        const int ja = j0 + l, jb = j0 + l + 32;
                        const bool oka = ja < k && ja != p, okb = jb < k && jb != p;
                        const double pa = oka ? a[p * ld + ja] : 0.0, pb = okb ? a[p * ld + jb] : 0.0;
                        // four rows at a time: all their loads before the first store (a store to a[i][ja] could alias
                        // the next row's a[i'][p] as far as the compiler knows, which would serialise the rows)
                        for (int i0 = w; i0 < k; i0 += 4 * nw) {
                            double f[4], va[4], vb[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int i = i0 + u * nw;
                                const bool rok = i < k && i != p;
                                f[u] = rok ? a[i * ld + p] * ipiv : 0.0;
                                va[u] = (rok && oka) ? a[i * ld + ja] : 0.0;
                                vb[u] = (rok && okb) ? a[i * ld + jb] : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int i = i0 + u * nw;
                                const bool rok = i < k && i != p;
                                if (rok && oka) a[i * ld + ja] = va[u] - f[u] * pa;
                                if (rok && okb) a[i * ld + jb] = vb[u] - f[u] * pb;
                            }
                        }
                    }
                }
            } else {
                for (int e = tid; e < k * k; e += nthreads) {
                    int i = e / k, j = e - i * k;
                    if (i != p && j != p) a[i * ld + j] -= a[i * ld + p] * ipiv * a[p * ld + j];
                }
            }
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
    for (int j = tid; j < k; j += nthreads) {
        double cs = 0.0;
        for (int i = 0; i < k; ++i) cs += fabs(a[i * ld + j]);
        atomicMax(&nrm[0], (unsigned long long)__double_as_longlong(cs));
    }
    bar();
    if (tid == 0) {
        const double in = __longlong_as_double((long long)nrm[0]);
        if (!isfinite(in) || an * in * 2.220446049250313e-16 > 1.0) scratch[1] = 1.0;
    }
    bar();
    return scratch[1] != 0.0;
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
// sM: shared G^-1 (ld KP, symmetric); sG: shared G (same layout; nullptr: always start from S = all).
// Returns this lane's partial of sum_i x_i r_i.
// ---------------------------------------------------------------------------
// Number of unif_rand() draws base::max.col(ties.method = "random") makes on the row (v_0 .. v_{k-1}) held one entry
// per lane by an LP-lane group (R_max_col: `large` = max |finite entry|, tol = 1e-5 large, running maximum `a`; every
// entry with b <= a + tol and b >= a - tol costs one draw -- including the structural ties among leading -Inf / zero
// entries).  The count depends on the values only, never on the uniforms drawn.
template <int LP>
__device__ __forceinline__ int r_max_col_draws(double v, int k, unsigned gmask, int base) {
    double large = 0.0;
#pragma unroll
    for (int j = 0; j < LP; ++j) {
        const double vj = __shfl_sync(gmask, v, base + j);
        if (j < k && isfinite(vj)) large = fmax(large, fabs(vj));
    }
    const double tol = 1e-5 * large;
    double a = __shfl_sync(gmask, v, base);
    int cnt = 0;
#pragma unroll
    for (int j = 1; j < LP; ++j) {
        const double b = __shfl_sync(gmask, v, base + j);
        if (j < k) {
            if (b > a + tol) a = b;
            else if (b >= a - tol) ++cnt;
        }
    }
    return cnt;
}

// `draws` (optional): the lane with i == 0 of every group adds the number of uniforms NMF::.fcnnls would have drawn
// for its problems through max.col -- one call per Lawson-Hanson step back (row = -alpha, -Inf off the candidates) and
// one per variable entering the passive set (row = (!Pset) * W).
template <int KP, int XS>
__device__ double nnls_lanes(const double *__restrict__ sG, const double *__restrict__ sM, int k, double *xs, int first,
                             int count, int nvalid, int lane, double *rowbuf /* this warp's, >= (32 / LP) * (KP + 2) doubles, 16-byte aligned */,
                             unsigned long long *draws = nullptr) {   // (each lane leaves x_i of its problem in xs)
    constexpr int LP = (KP <= 16) ? 16 : 32;
    constexpr int PPW = 32 / LP;
    const unsigned FULL = 0xffffffffu;
    const int sub = lane / LP, i = lane % LP, base = sub * LP;
    const bool rowvalid = i < k;
    const unsigned gmask = (LP == 32) ? 0xffffffffu : (0xffffu << base);
    double t1 = 0.0;
    int ndraw = 0;
    for (int rd = 0; rd < count; rd += PPW) {
        const int unit = first + rd + sub;
        const bool uvalid = (rd + sub < count) && (unit < nvalid);
        if (!__any_sync(FULL, uvalid)) {
            if (rowvalid && rd + sub < count) xs[i * XS + unit] = 0.0;
            continue;
        }
        const double ri = (rowvalid && uvalid) ? xs[i * XS + unit] : 0.0;
        // unconstrained solution xu = G^-1 r: .fcnnls' first passive set is P = {xu > 0}
        double T[KP];
        double xu = 0.0;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const double rj = __shfl_sync(FULL, ri, base + (j % LP));
            const double mij = (rowvalid && j < k) ? sM[j * KP + i] : 0.0;
            T[j] = -mij;
            xu += mij * rj;
        }
        const unsigned negs = (__ballot_sync(FULL, rowvalid && !(xu > 0.0)) & gmask) >> base;
        const unsigned poss = (__ballot_sync(FULL, rowvalid && xu > 0.0) & gmask) >> base;
        // The tableau of the basis P is reached from whichever end is closer: from S = all (q = xu, T = -G^-1) by
        // exchanging the variables of ~P out, or from S = {} (q = r, T = -G) by exchanging those of P in -- the same
        // principal pivots either way, min(|P|, k - |P|) of them (sparse factors: a handful instead of almost k).
        const bool fromG = sG != nullptr && __popc(poss) < __popc(negs);
        if (__any_sync(FULL, fromG)) {
#pragma unroll
            for (int j = 0; j < KP; ++j)
                if (fromG) T[j] = (rowvalid && j < k) ? -sG[j * KP + i] : 0.0;
        }
        double qv = fromG ? ri : xu;
        bool inS = rowvalid && !fromG;
        double dv = xu > 0.0 ? xu : 0.0;
        unsigned pend = uvalid ? (fromG ? poss : negs) : 0u;
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
                        if (draws) ndraw += r_max_col_draws<LP>(-al, k, gmask, base);
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
                            if (draws) ndraw += r_max_col_draws<LP>((!inS && rowvalid) ? qv : 0.0, k, gmask, base);
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
            // The pivot row goes through shared memory: its owner stores it (8 x 16 bytes), every lane reads it back
            // with broadcast loads -- instead of 2 KP shuffles.  The pivot COLUMN entry of row i is taken from the same
            // copy: a principal-pivot tableau of a symmetric matrix stays sign-symmetric, T[i][p] = +-T[p][i] with
            // + when i and p are in the same class (both in S or both out), so the KP-way select of a dynamically
            // indexed register (3 KP instructions) becomes one load.  (Equal up to rounding, not bit for bit.)
            const int pq = piv >= 0 ? piv : 0;
            const int src = base + pq;
            double *rb = rowbuf + sub * (KP + 2);
            if (i == pq) {
#pragma unroll
                for (int j = 0; j < KP; j += 2) *reinterpret_cast<double2 *>(rb + j) = make_double2(T[j], T[j + 1]);
                rb[KP] = inS ? 1.0 : 0.0;
            }
            __syncwarp();
            const double qq = __shfl_sync(FULL, qv, src);
            const double pp = rb[pq];
            const bool sameclass = (rb[KP] != 0.0) == inS;
            double tiq = (i < KP) ? rb[i] : 0.0;
            tiq = sameclass ? tiq : -tiq;
            const double ip = 1.0 / pp;
            const bool act = piv >= 0, isrow = act && (i == piv);
            const double f = tiq * ip;
            const double a = isrow ? 0.0 : 1.0;
            const double b = act ? (isrow ? -ip : -f) : 0.0;
            const double fix = isrow ? ip : f;
            const int pfix = act ? pq : -1;
#pragma unroll
            for (int j = 0; j < KP; j += 2) {
                const double2 rq = *reinterpret_cast<const double2 *>(rb + j);
                const double t0 = fma(b, rq.x, a * T[j]), t1 = fma(b, rq.y, a * T[j + 1]);
                T[j] = (j == pfix) ? fix : t0;
                T[j + 1] = (j + 1 == pfix) ? fix : t1;
            }
            qv = fma(b, qq, a * qv);
            if (isrow) inS = !inS;
            __syncwarp();                                    // the row buffer is rewritten by the next exchange
        }
        const double x = (inS && rowvalid && uvalid && qv > 0.0) ? qv : 0.0;
        t1 += x * ri;
        if (rowvalid && rd + sub < count) xs[i * XS + unit] = x;
    }
    if (draws && i == 0 && ndraw) atomicAdd(draws, (unsigned long long)ndraw);
    return t1;
}

// ---------------------------------------------------------------------------
// Warp-cooperative exact NNLS for any rank up to 64 (used for k > 32, where a tableau row no longer fits the
// registers of one lane): active-set iteration with a COMPRESSED solve.
//
// First passive set = {xu > 0} as in .fcnnls; then BLOCK principal pivoting (Judice & Pires / Kim & Park): every
// iteration solves the current partition exactly and exchanges ALL infeasible variables at once (negative primal
// values leave the passive set, positive gradients enter it); when the number of infeasibilities has not
// decreased for three iterations only the infeasible variable with the largest index is exchanged (Murty's rule,
// which cannot cycle).  The optimum of the strictly convex problem is unique, so this reaches the same solution as
// the reference's one-variable-at-a-time Lawson-Hanson path in ~4 solves instead of one per exchanged variable.
// A group of LP lanes owns one problem, lane i owning the variables i (+ LP); every iteration solves the current
// partition directly on its SMALLER side S (|S| <= k/2):
//   S = N (active set):   y = (M_NN)^-1 xu_N,  x_P = xu_P - M_PN y,  w_N = y          (M = G^-1, xu = M r)
//   S = P (passive set):  y = (G_PP)^-1 r_P,   x_P = y,              w_N = r_N - G_NP y
// The |S| x |S| system is gathered into registers (lane l = compressed row l), eliminated without pivoting (SPD)
// with the pivot rows broadcast by shuffles, and back-substituted so that every lane ends up holding all of y.
// The initial clamp guess therefore costs ONE small solve instead of one tableau exchange per clamped variable.
//
// xs: shared [KP][XS] -- right-hand sides r on entry, solutions x on exit (column = problem).  The warp solves the
// problems [first, first + count) of xs, 32 / LP at a time; problems >= nvalid are skipped (x = 0).
// scr: per-warp shared scratch of nnls_scratch_doubles<KP>() doubles.  Returns this lane's partial of sum x_i r_i.
// ---------------------------------------------------------------------------
// 1/x to full double precision from the single-precision estimate and two Newton steps (x well inside the float
// range, as Gram pivots are); falls back to the division otherwise
__device__ __forceinline__ double fast_rcp(double x) {
    const float xf = (float)x;
    double r = (double)(1.0f / xf);
    if (!(fabsf(xf) > 1e-30f && fabsf(xf) < 1e30f)) return 1.0 / x;
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

template <int KP>
struct NnlsCfg {
    static constexpr int LP = (KP <= 16) ? 16 : 32;
    static constexpr int R = (KP + LP - 1) / LP;           // variables per lane
    static constexpr int NS = (KP / 2 < 1) ? 1 : KP / 2;   // largest compressed system
    static constexpr int PPW = 32 / LP;                    // problems per warp and round
    static constexpr int SCR = PPW * (KP + NS + (NS + 1) / 2);   // doubles: xu[KP], y[NS], idx[NS] (ints) per problem
};

// One exact solve of the current partition on its smaller side S (see nnls_warp below), ns <= ns_max <= NSB: gathers the
// ns x ns system into registers (lane l = compressed row l), eliminates, back-substitutes so that every lane holds all
// of y, and forms the primal values on P and the gradient on N.
template <int KP, int XS, int NSB>
__device__ __forceinline__ void nnls_partition_solve(const double *Msrc, const double *xu_s, const double *xs, int ucol, const int *idx_s,
                                                     double *y_s, int i, int base, int ns, int ns_max, bool useN, bool done,
                                                     unsigned long long smask, const int (&pos_)[NnlsCfg<KP>::R],
                                                     const bool (&val_)[NnlsCfg<KP>::R], const double (&xu_)[NnlsCfg<KP>::R],
                                                     const double (&r_)[NnlsCfg<KP>::R], double (&x_)[NnlsCfg<KP>::R],
                                                     double (&w_)[NnlsCfg<KP>::R]) {
    constexpr int LP = NnlsCfg<KP>::LP, R = NnlsCfg<KP>::R;
    const unsigned FULL = 0xffffffffu;
    // ---- gather the compressed system: lane i < ns owns row i ----
    double A[NSB], yv[NSB];
    const bool rowok = i < ns;
    const int my = rowok ? idx_s[i] : 0;
#pragma unroll
    for (int b = 0; b < NSB; ++b) {
        A[b] = (b == i) ? 1.0 : 0.0;                 // identity padding outside the ns x ns block
        yv[b] = 0.0;
        if (b < ns_max) {
            const int ib = (b < ns) ? idx_s[b] : 0;
            if (rowok && b < ns) A[b] = Msrc[ib * KP + my];
        }
    }
    double rhs = rowok ? (useN ? xu_s[my] : xs[my * XS + ucol]) : 0.0;
    // ---- forward elimination (no pivoting: principal sub-matrix of an SPD matrix) ----
#pragma unroll
    for (int a = 0; a < NSB; ++a) {
        if (a < ns_max) {
            const double pa = __shfl_sync(FULL, A[a], base + a);
            const double ip = fast_rcp(pa);
            yv[a] = ip;
            const double f = (i > a) ? A[a] * ip : 0.0;
#pragma unroll
            for (int b = a + 1; b < NSB; ++b)
                if (b < ns_max) A[b] -= f * __shfl_sync(FULL, A[b], base + a);
            rhs -= f * __shfl_sync(FULL, rhs, base + a);
        }
    }
    // ---- back substitution: every lane learns every y_a ----
#pragma unroll
    for (int a = NSB - 1; a >= 0; --a) {
        if (a < ns_max) {
            const double ya = __shfl_sync(FULL, rhs, base + a) * yv[a];
            yv[a] = ya;
            if (i < a) rhs -= A[a] * ya;
            if (i == 0 && a < ns) y_s[a] = ya;
        }
    }
    __syncwarp();
    // ---- primal values on P, gradient on N ----
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int v = i + LP * q;
        if (val_[q] && !done) {
            const bool inS = (smask >> v) & 1ull;
            if (inS) {
                const double ys = y_s[pos_[q]];
                x_[q] = useN ? 0.0 : ys;
                w_[q] = useN ? ys : 0.0;
            } else {
                double acc = useN ? xu_[q] : r_[q];
#pragma unroll
                for (int b = 0; b < NSB; ++b)
                    if (b < ns) acc -= Msrc[idx_s[b] * KP + v] * yv[b];
                x_[q] = useN ? acc : 0.0;
                w_[q] = useN ? 0.0 : acc;
            }
        }
    }
    __syncwarp();
}

template <int KP, int XS>
__device__ double nnls_warp(const double *__restrict__ sG, const double *__restrict__ sM, int k, double *xs,
                            double *scr, int first, int count, int nvalid, int lane) {
    using C = NnlsCfg<KP>;
    constexpr int LP = C::LP, R = C::R, NS = C::NS, PPW = C::PPW;
    const unsigned FULL = 0xffffffffu;
    const int grp = lane / LP, i = lane % LP, base = grp * LP;
    const unsigned long long fullk = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    double *xu_s = scr + grp * (KP + NS + (NS + 1) / 2);
    double *y_s = xu_s + KP;
    int *idx_s = (int *)(y_s + NS);
    double t1 = 0.0;
    for (int rd = 0; rd < count; rd += PPW) {
        const int unit = first + rd + grp;
        const bool inrange = rd + grp < count;
        const bool uvalid = inrange && (unit < nvalid);
        if (!__any_sync(FULL, uvalid)) {
            if (inrange)
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (i + LP * q < k) xs[(i + LP * q) * XS + unit] = 0.0;
            continue;
        }
        const int ucol = inrange ? unit : first;           // a readable column for lanes without a problem
        double r_[R], xu_[R], x_[R], w_[R];
        bool inN_[R], val_[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int v = i + LP * q;
            val_[q] = uvalid && (v < k);
            r_[q] = val_[q] ? xs[v * XS + ucol] : 0.0;
            double acc = 0.0;
            if (v < k)
                for (int j = 0; j < k; ++j) acc += sM[j * KP + v] * xs[j * XS + ucol];
            xu_[q] = val_[q] ? acc : 0.0;
            if (v < KP) xu_s[v] = xu_[q];
            inN_[q] = val_[q] && !(xu_[q] > 0.0);
            x_[q] = xu_[q] > 0.0 ? xu_[q] : 0.0;
            w_[q] = 0.0;
        }
        bool done = !uvalid;
        int iters = 0, best_inf = k + 1, patience = 3;
        const int cap = 6 * k + 16;
        while (true) {
            // ---- partition of this group's problem ----
            unsigned long long maskN = 0ull;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const unsigned bal = __ballot_sync(FULL, inN_[q]);
                const unsigned long long gb = (LP == 32) ? (unsigned long long)bal : (unsigned long long)((bal >> base) & 0xffffu);
                maskN |= gb << (LP * q);
            }
            const int na = __popcll(maskN), np = k - na;
            const bool useN = na <= np;
            const unsigned long long smask = useN ? maskN : (~maskN & fullk);
            const int ns = done ? 0 : (useN ? na : np);
            int ns_max = ns;
#pragma unroll
            for (int o = 16; o >= LP; o >>= 1) ns_max = max(ns_max, __shfl_xor_sync(FULL, ns_max, o));
            if (!__any_sync(FULL, !done)) break;
            const double *Msrc = useN ? sM : sG;
            int pos_[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int v = i + LP * q;
                pos_[q] = __popcll(smask & ((1ull << v) - 1ull));
                if (!done && ((smask >> v) & 1ull)) idx_s[pos_[q]] = v;
            }
            __syncwarp();
            // ---- solve the partition on its smaller side with loops unrolled to the smallest bucket that holds it ----
            // (the systems of one launch are mostly far smaller than KP / 2: unrolled to NS, every elimination step
            // would issue NS predicated-off updates whatever ns is)
            if (ns_max <= 4)
                nnls_partition_solve<KP, XS, (NS < 4 ? NS : 4)>(Msrc, xu_s, xs, ucol, idx_s, y_s, i, base, ns, ns_max, useN, done, smask, pos_, val_, xu_, r_, x_, w_);
            else if (ns_max <= 8)
                nnls_partition_solve<KP, XS, (NS < 8 ? NS : 8)>(Msrc, xu_s, xs, ucol, idx_s, y_s, i, base, ns, ns_max, useN, done, smask, pos_, val_, xu_, r_, x_, w_);
            else if (ns_max <= 16)
                nnls_partition_solve<KP, XS, (NS < 16 ? NS : 16)>(Msrc, xu_s, xs, ucol, idx_s, y_s, i, base, ns, ns_max, useN, done, smask, pos_, val_, xu_, r_, x_, w_);
            else
                nnls_partition_solve<KP, XS, NS>(Msrc, xu_s, xs, ucol, idx_s, y_s, i, base, ns, ns_max, useN, done, smask, pos_, val_, xu_, r_, x_, w_);
            // ---- block principal pivoting ----
            if (!done) {
                const unsigned gm = (LP == 32) ? FULL : (0xffffu << base);
                bool bad_[R];
                int ninf = 0, top = -1;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    bad_[q] = val_[q] && (inN_[q] ? (w_[q] > 0.0) : (x_[q] < 0.0));
                    const unsigned bal = __ballot_sync(gm, bad_[q]);
                    const unsigned gb = (LP == 32) ? bal : ((bal >> base) & 0xffffu);
                    ninf += __popc(gb);
                    if (gb) top = max(top, LP * q + 31 - __clz(gb));     // largest infeasible variable index
                }
                if (ninf == 0) {
                    done = true;
                } else {
                    bool all;
                    if (ninf < best_inf) { best_inf = ninf; patience = 3; all = true; }
                    else if (patience > 0) { --patience; all = true; }
                    else all = false;
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (bad_[q] && (all || i + LP * q == top)) inN_[q] = !inN_[q];
                }
                if (++iters > cap) done = true;
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int v = i + LP * q;
            const double x = (val_[q] && !inN_[q] && x_[q] > 0.0) ? x_[q] : 0.0;
            t1 += x * r_[q];
            if (inrange && v < k) xs[v * XS + unit] = x;
        }
        __syncwarp();
    }
    return t1;
}
