// NIPALS PCA on the GPU: replaces nipals::nipals (nipals 0.5; call site R/bicluster.R:226), complete data and the
// missing-value branch (NaN = NA: every sum skips missing cells and the regression denominators t't / p'p become
// per-column / per-row sums over the observed cells, nipals' `has.na` path).
//
// The whole algorithm -- optional centring / sd scaling, and for every component the start-column search, the
// deflated power iteration (two passes over x per iteration: p = x't / t't and t = x p / p'p), Gram-Schmidt
// against the previous loadings / scores, the convergence test and the fused deflation + column |.|-sum pass --
// runs inside ONE persistent cooperative kernel; phases are separated by grid-wide barriers, every reduction is
// a fixed-order two-stage sum (per block, then over blocks) so results are run-to-run identical, and no value
// makes a round trip to the host until the last component is done.  The n x n and m x m Gram-Schmidt matrices
// the R code materialises (PPp, TTp) are never formed: P(P'p) and T(T't / eig) are applied as two thin products.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#define NIP_THREADS 256

struct NipalsParams {
    double *x;            // working copy, m x n, column-major, leading dimension ld (deflated in place)
    int64_t m, n, ld;
    int ncomp, center, scale, maxiter, gramschmidt, has_na;
    double tol;
    double *T;            // m x ncomp unnormalised scores
    double *P;            // n x ncomp loadings
    double *eig;          // ncomp: sum(th^2)
    int *iters;           // ncomp
    double *th, *th_old;  // m
    double *ph;           // n  (raw x'th, then the normalised loading)
    double *colabs;       // n
    double *phn;          // [nblocks][n] block-private normalised loading
    double *part;         // [nblocks][ncomp + 4] per-block partials
    double *spart;        // [nblocks][m] per-block partial scores (complete-data S3: blocks own COLUMNS)
    int chunk;            // rows of the shared-memory score accumulator per pass
    int *status;          // 0 ok, 1 maxiter reached on some component, 2 constant column under scale
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NIP_THREADS / 32; ++i) s += sh[i];
    return s;
}

// ---- warp-level column helpers: a warp owns whole columns and keeps 4 columns x 4 rows of loads in flight ----
#define NIP_WARPS (NIP_THREADS / 32)

// dot products of up to four columns with the vector t (lane partials; caller reduces)
__device__ __forceinline__ void cols_dot4(const double *c0, const double *c1, const double *c2, const double *c3,
                                          const double *t, int64_t m, int lane, double s[4]) {
    s[0] = s[1] = s[2] = s[3] = 0.0;
    int64_t i = lane;
    for (; i + 96 < m; i += 128) {
        double tv[4], a[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tv[u] = __ldcg(&t[i + 32 * u]);
            a[0][u] = c0[i + 32 * u]; a[1][u] = c1[i + 32 * u]; a[2][u] = c2[i + 32 * u]; a[3][u] = c3[i + 32 * u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s[0] += a[0][u] * tv[u]; s[1] += a[1][u] * tv[u]; s[2] += a[2][u] * tv[u]; s[3] += a[3][u] * tv[u];
        }
    }
    for (; i < m; i += 32) {
        const double tv = __ldcg(&t[i]);
        s[0] += c0[i] * tv; s[1] += c1[i] * tv; s[2] += c2[i] * tv; s[3] += c3[i] * tv;
    }
}


// consumer-only barrier of the streaming kernel (its producer warp never joins) or the whole block
template <bool CONS>
__device__ __forceinline__ void nip_bar() {
    if (CONS) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncthreads();
}

// ---- pre-processing: per column mean / sd (about the mean, n-1 denominator), one warp per column; leaves the
// column |.| sums of the starting matrix in colabs ----
__device__ void nip_preprocess(const NipalsParams &p, int64_t gw, int64_t tw, int lane) {
    const int64_t m = p.m, n = p.n, ld = p.ld;
    if (p.center || p.scale) {
        for (int64_t j = gw; j < n; j += tw) {
            double *col = p.x + j * ld;
            if (p.has_na) {      // colMeans(na.rm = TRUE) / sd(na.rm = TRUE): plain loops, missing cells skipped
                double s = 0.0, cnt = 0.0;
                for (int64_t i = lane; i < m; i += 32) { const double v = col[i]; if (v == v) { s += v; cnt += 1.0; } }
                s = warp_sum(s); cnt = warp_sum(cnt);
                const double mean = s / cnt;
                double sd = 1.0;
                if (p.scale) {
                    double q = 0.0;
                    for (int64_t i = lane; i < m; i += 32) { const double v = col[i]; if (v == v) q += (v - mean) * (v - mean); }
                    sd = sqrt(warp_sum(q) / (cnt - 1.0));
                    if (!(sd > 0.0) && lane == 0) atomicExch(p.status, 2);
                }
                double ab = 0.0;
                for (int64_t i = lane; i < m; i += 32) {
                    double v = col[i];
                    if (p.center) v -= mean;
                    if (p.scale) v /= sd;
                    col[i] = v;
                    if (v == v) ab += fabs(v);
                }
                ab = warp_sum(ab);
                if (lane == 0) p.colabs[j] = ab;
                continue;
            }
            // (eight rows of loads in flight per lane; the column stays in L2 for the second and third pass)
            double sa[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) sa[u] = 0.0;
            int64_t i = lane;
            for (; i + 224 < m; i += 256) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = col[i + 32 * u];
#pragma unroll
                for (int u = 0; u < 8; ++u) sa[u] += v[u];
            }
            for (; i < m; i += 32) sa[0] += col[i];
            const double mean = warp_sum(((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]))) / (double)m;
            double sd = 1.0;
            if (p.scale) {
#pragma unroll
                for (int u = 0; u < 8; ++u) sa[u] = 0.0;
                i = lane;
                for (; i + 224 < m; i += 256) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = col[i + 32 * u] - mean;
#pragma unroll
                    for (int u = 0; u < 8; ++u) sa[u] += v[u] * v[u];
                }
                for (; i < m; i += 32) { const double d0 = col[i] - mean; sa[0] += d0 * d0; }
                sd = sqrt(warp_sum(((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]))) / (double)(m - 1));
                if (!(sd > 0.0) && lane == 0) atomicExch(p.status, 2);
            }
            double ab = 0.0;
            i = lane;
            for (; i + 224 < m; i += 256) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = col[i + 32 * u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (p.center) v[u] -= mean;
                    if (p.scale) v[u] /= sd;
                    col[i + 32 * u] = v[u];
                    ab += fabs(v[u]);
                }
            }
            for (; i < m; i += 32) {
                double v = col[i];
                if (p.center) v -= mean;
                if (p.scale) v /= sd;
                col[i] = v;
                ab += fabs(v);
            }
            ab = warp_sum(ab);
            if (lane == 0) p.colabs[j] = ab;
        }
    } else {
        // column |.| sums of the starting matrix
        for (int64_t j = gw; j < n; j += tw) {
            const double *col = p.x + j * ld;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t i = lane;
            if (p.has_na) {
                for (; i < m; i += 32) { const double v = col[i]; if (v == v) s0 += fabs(v); }
            }
            for (; i + 96 < m; i += 128) { s0 += fabs(col[i]); s1 += fabs(col[i + 32]); s2 += fabs(col[i + 64]); s3 += fabs(col[i + 96]); }
            for (; i < m; i += 32) s0 += fabs(col[i]);
            const double ab = warp_sum((s0 + s1) + (s2 + s3));
            if (lane == 0) p.colabs[j] = ab;
        }
    }
}

// start column of a component: which.max(colSums(abs(x))) -- first maximum; every block finds it redundantly
template <bool CONS>
__device__ int64_t nip_start_col(const NipalsParams &p, int tid) {
    __shared__ double sbv[NIP_THREADS];
    __shared__ long long sbj[NIP_THREADS];
    double best = -1.0;
    int64_t bj = 0;
    for (int64_t j = tid; j < p.n; j += NIP_THREADS) {
        const double v = __ldcg(&p.colabs[j]);
        if (v > best) { best = v; bj = j; }
    }
    // block arg-max, ties -> smallest index
    sbv[tid] = best; sbj[tid] = bj;
    nip_bar<CONS>();
    for (int o = NIP_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) {
            const double ov = sbv[tid + o];
            const long long oj = sbj[tid + o];
            if (ov > sbv[tid] || (ov == sbv[tid] && oj < sbj[tid])) { sbv[tid] = ov; sbj[tid] = oj; }
        }
        nip_bar<CONS>();
    }
    const int64_t scol = sbj[0];
    nip_bar<CONS>();
    return scol;
}

// deflation x -= t ph' fused with the column |.| sums the next start-column search needs; a warp owns whole columns
__device__ void nip_deflate(const NipalsParams &p, const double *tvec, const double *phn, int64_t gw, int64_t tw, int lane) {
    const int64_t m = p.m, n = p.n, ld = p.ld;
    for (int64_t j = gw; j < n; j += tw) {
        double *col = p.x + j * ld;
        const double pj = phn[j];
        double s0 = 0.0, s1 = 0.0;
        int64_t i = lane;
        for (; i + 224 < m; i += 256) {
            double tq[8], v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { tq[u] = __ldcg(&tvec[i + 32 * u]); v[u] = col[i + 32 * u]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v[u] -= tq[u] * pj;
                col[i + 32 * u] = v[u];
                const double a = (v[u] == v[u]) ? fabs(v[u]) : 0.0;
                if (u & 1) s1 += a; else s0 += a;
            }
        }
        for (; i < m; i += 32) {
            const double v0 = col[i] - __ldcg(&tvec[i]) * pj;
            col[i] = v0;
            s0 += (v0 == v0) ? fabs(v0) : 0.0;
        }
        const double ab = warp_sum(s0 + s1);
        if (lane == 0) p.colabs[j] = ab;
    }
}

__global__ void __launch_bounds__(NIP_THREADS) nipals_kernel(NipalsParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double nsm[];          // [2][NIP_WARPS][256] partial scores (and denominators) of a 256-row pass
    __shared__ double sh[NIP_THREADS / 32];
    __shared__ double scoef[64];
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t m = p.m, n = p.n, ld = p.ld;
    const int PW = p.ncomp + 4;      // width of a block's partial record
    const int64_t gw = (int64_t)b * NIP_WARPS + warp, tw = (int64_t)nb * NIP_WARPS;   // global warp index / count

    nip_preprocess(p, gw, tw, lane);
    grid.sync();

    // rows of the scores owned by this block: a multiple of 32 so that a warp always reads 256 contiguous bytes
    const int64_t rows_per = ((m + nb - 1) / nb + 31) / 32 * 32;
    const int64_t r0 = (b * rows_per < m) ? b * rows_per : m, r1 = (r0 + rows_per < m) ? r0 + rows_per : m;

    for (int h = 0; h < p.ncomp; ++h) {
        // start column: which.max(colSums(abs(x))) -- first maximum; every block finds it redundantly
        const int64_t scol = nip_start_col<false>(p, tid);
        // th = x[, scol]; tt = sum(th^2)
        double tts = 0.0;
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) {
            double v = __ldcg(&p.x[scol * ld + i]);
            if (!(v == v)) v = 0.0;                      // x0: missing cells count as zero
            p.th[i] = v;
            tts += v * v;
        }
        tts = block_sum(tts, sh);
        if (tid == 0) p.part[(size_t)b * PW + 0] = tts;
        grid.sync();
        double tt = 0.0;
        for (int bb = 0; bb < nb; ++bb) tt += __ldcg(&p.part[(size_t)bb * PW + 0]);

        int pciter = 1;
        const double *phn = p.phn + (size_t)b * n;
        while (true) {
            // S1: ph_raw[j] = sum_i x[i,j] th[i]: a warp owns columns, four at a time
            for (int64_t j0 = gw * 4; j0 < n; j0 += tw * 4) {
                const double *c0 = p.x + j0 * ld;
                const double *c1 = p.x + ((j0 + 1 < n) ? j0 + 1 : j0) * ld;
                const double *c2 = p.x + ((j0 + 2 < n) ? j0 + 2 : j0) * ld;
                const double *c3 = p.x + ((j0 + 3 < n) ? j0 + 3 : j0) * ld;
                double s4[4], den4[4] = {tt, tt, tt, tt};
                if (p.has_na) {      // crossprod(x0, th) / colSums(T2): numerator and denominator over observed cells
                    const double *cc[4] = {c0, c1, c2, c3};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double sn = 0.0, sd2 = 0.0;
                        for (int64_t i = lane; i < m; i += 32) {
                            const double v = cc[q][i], t = __ldcg(&p.th[i]);
                            if (v == v) { sn += v * t; sd2 += t * t; }
                        }
                        s4[q] = sn;
                        den4[q] = warp_sum(sd2);
                    }
                } else {
                    cols_dot4(c0, c1, c2, c3, p.th, m, lane, s4);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) s4[q] = warp_sum(s4[q]) / den4[q];
                if (lane < 4 && j0 + lane < n) p.ph[j0 + lane] = s4[lane];
            }
            grid.sync();
            // S2 (redundant per block): ph = ph_raw / tt; Gram-Schmidt against P[, 0..h); normalise
            double pp_den;   // ph'ph after normalisation
            const int nh = (p.gramschmidt && h > 0) ? h : 0;
            {
                for (int l = 0; l < nh; ++l) {
                    double s = 0.0;
                    for (int64_t j = tid; j < n; j += NIP_THREADS) s += p.P[(size_t)l * n + j] * __ldcg(&p.ph[j]);
                    s = block_sum(s, sh);
                    if (tid == 0) scoef[l] = s;
                }
                __syncthreads();
                double nrm = 0.0;
                for (int64_t j = tid; j < n; j += NIP_THREADS) {
                    double v = __ldcg(&p.ph[j]);
                    double corr = 0.0;
                    for (int l = 0; l < nh; ++l) corr += p.P[(size_t)l * n + j] * scoef[l];
                    v -= corr;
                    nrm += v * v;
                }
                nrm = sqrt(block_sum(nrm, sh));
                double p2 = 0.0;
                for (int64_t j = tid; j < n; j += NIP_THREADS) {
                    double v = __ldcg(&p.ph[j]);
                    double corr = 0.0;
                    for (int l = 0; l < nh; ++l) corr += p.P[(size_t)l * n + j] * scoef[l];
                    v = (v - corr) / nrm;
                    p2 += v * v;
                    p.phn[(size_t)b * n + j] = v;   // block-private copy of the normalised loading (no extra barrier)
                }
                p2 = block_sum(p2, sh);
                pp_den = p2;
                __syncthreads();
            }
            // S3, complete data: th_new = x ph / (ph'ph) with the blocks owning COLUMNS, like S1: a block streams whole
            // column segments (contiguous, DRAM-page friendly; four columns x four rows of loads in flight per thread)
            // and accumulates its partial scores for `chunk` rows in shared memory (thread t only ever touches the
            // entries t, t + 256, ...: no barriers), then the blocks' partials are added in block order by the owner of
            // each row.  (Row-owning blocks -- the missing-value branch below -- read 768-byte pieces 160 KB apart.)
            if (!p.has_na) {
                const int nbc = (n < (int64_t)nb) ? (int)n : nb;            // blocks that own columns
                if (b < nbc) {
                    const int64_t jlo = n * b / nbc, jhi = n * (b + 1) / nbc;
                    for (int64_t c0 = 0; c0 < m; c0 += p.chunk) {
                        const int len = (int)((m - c0 < p.chunk) ? m - c0 : p.chunk);
                        for (int i = tid; i < len; i += NIP_THREADS) nsm[i] = 0.0;
                        int64_t j = jlo;
                        for (; j + 3 < jhi; j += 4) {
                            const double q0 = phn[j], q1 = phn[j + 1], q2 = phn[j + 2], q3 = phn[j + 3];
                            const double *ca = p.x + j * ld + c0;
                            int i = tid;
                            for (; i + 3 * NIP_THREADS < len; i += 4 * NIP_THREADS) {
                                double v[4][4];
#pragma unroll
                                for (int u = 0; u < 4; ++u)
#pragma unroll
                                    for (int c = 0; c < 4; ++c) v[c][u] = __ldcg(&ca[c * ld + i + u * NIP_THREADS]);
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    nsm[i + u * NIP_THREADS] += (v[0][u] * q0 + v[1][u] * q1) + (v[2][u] * q2 + v[3][u] * q3);
                            }
                            for (; i < len; i += NIP_THREADS)
                                nsm[i] += (__ldcg(&ca[i]) * q0 + __ldcg(&ca[ld + i]) * q1) +
                                          (__ldcg(&ca[2 * ld + i]) * q2 + __ldcg(&ca[3 * ld + i]) * q3);
                        }
                        for (; j < jhi; ++j) {
                            const double q0 = phn[j];
                            const double *ca = p.x + j * ld + c0;
                            for (int i = tid; i < len; i += NIP_THREADS) nsm[i] += __ldcg(&ca[i]) * q0;
                        }
                        double *dst = p.spart + (size_t)b * m + c0;
                        for (int i = tid; i < len; i += NIP_THREADS) dst[i] = nsm[i];
                    }
                }
                grid.sync();
                // add the blocks' partials for this block's rows in block order: 256 rows per pass, warp w takes its
                // eighth of the blocks for all eight 32-row groups (4 blocks x 8 groups of loads in flight), then the
                // warps' sums are added in warp order
                for (int64_t rb = r0; rb < r1; rb += 256) {
                    const int blo = nbc * warp / NIP_WARPS, bhi = nbc * (warp + 1) / NIP_WARPS;
                    double acc[8];
                    bool ok[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) { acc[g] = 0.0; ok[g] = rb + 32 * g + lane < r1; }
                    int bb = blo;
                    for (; bb + 3 < bhi; bb += 4) {
                        double v[4][8];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                v[u][g] = ok[g] ? __ldcg(&p.spart[(size_t)(bb + u) * m + rb + 32 * g + lane]) : 0.0;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int g = 0; g < 8; ++g) acc[g] += v[u][g];
                    }
                    for (; bb < bhi; ++bb)
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (ok[g]) acc[g] += __ldcg(&p.spart[(size_t)bb * m + rb + 32 * g + lane]);
#pragma unroll
                    for (int g = 0; g < 8; ++g) nsm[warp * 256 + 32 * g + lane] = acc[g];
                    __syncthreads();
                    const int64_t i = rb + tid;
                    if (tid < 256 && i < r1) {
                        double sacc = 0.0;
#pragma unroll
                        for (int w = 0; w < NIP_WARPS; ++w) sacc += nsm[w * 256 + tid];
                        p.th_old[i] = p.th[i];
                        p.th[i] = sacc / pp_den;
                    }
                    __syncthreads();
                }
            }
            // S3, missing values: this block's rows, 256 rows per pass: warp w sums its eighth of the columns for all
            // eight 32-row groups of the pass (two columns x eight row groups of loads in flight)
            for (int64_t rb = r0; p.has_na && rb < r1; rb += 256) {
                const int64_t jlo = n * warp / NIP_WARPS, jhi = n * (warp + 1) / NIP_WARPS;
                double acc[8], dacc[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) acc[g] = dacc[g] = 0.0;
                const int ngr = (int)((r1 - rb + 31) / 32 < 8 ? (r1 - rb + 31) / 32 : 8);
                int64_t j = jlo;
                if (p.has_na) {      // x0 %*% ph / colSums(P2): observed cells only
                    for (; j < jhi; ++j) {
                        const double pj0 = phn[j];
                        const double *ca = p.x + j * ld + rb + lane;
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (g < ngr && rb + 32 * g + lane < r1) {
                                const double v = __ldcg(&ca[32 * g]);
                                if (v == v) { acc[g] += v * pj0; dacc[g] += pj0 * pj0; }
                            }
                    }
                }
                for (; j + 1 < jhi; j += 2) {
                    const double pj0 = phn[j], pj1 = phn[j + 1];
                    const double *ca = p.x + j * ld + rb + lane, *cb2 = ca + ld;
                    double va[8], vb[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const bool ok = g < ngr && rb + 32 * g + lane < r1;
                        va[g] = ok ? __ldcg(&ca[32 * g]) : 0.0;      // (x is deflated by other blocks: read through L2)
                        vb[g] = ok ? __ldcg(&cb2[32 * g]) : 0.0;
                    }
#pragma unroll
                    for (int g = 0; g < 8; ++g) acc[g] += va[g] * pj0 + vb[g] * pj1;
                }
                for (; j < jhi; ++j) {
                    const double pj0 = phn[j];
                    const double *ca = p.x + j * ld + rb + lane;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < ngr && rb + 32 * g + lane < r1) acc[g] += __ldcg(&ca[32 * g]) * pj0;
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    nsm[warp * 256 + 32 * g + lane] = acc[g];
                    nsm[(NIP_WARPS + warp) * 256 + 32 * g + lane] = dacc[g];
                }
                __syncthreads();
                {
                    const int64_t i = rb + tid;
                    if (i < r1) {
                        double sacc = 0.0, sden = 0.0;
#pragma unroll
                        for (int w = 0; w < NIP_WARPS; ++w) { sacc += nsm[w * 256 + tid]; sden += nsm[(NIP_WARPS + w) * 256 + tid]; }
                        p.th_old[i] = p.th[i];
                        p.th[i] = sacc / (p.has_na ? sden : pp_den);
                    }
                }
                __syncthreads();
            }
            for (int l = 0; l < nh; ++l) {
                double s = 0.0;
                for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) s += p.T[(size_t)l * m + i] * p.th[i];
                s = block_sum(s, sh);
                if (tid == 0) p.part[(size_t)b * PW + 4 + l] = s;
            }
            if (nh > 0) {
                grid.sync();
                for (int l = tid; l < nh; l += NIP_THREADS) {
                    double s = 0.0;
                    for (int bb = 0; bb < nb; ++bb) s += __ldcg(&p.part[(size_t)bb * PW + 4 + l]);
                    scoef[l] = s / __ldcg(&p.eig[l]);
                }
                __syncthreads();
            }
            // th -= T[, 0..h) (T' th / eig); convergence sums
            double dsum = 0.0, tsum = 0.0;
            for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) {
                double v = p.th[i];
                double corr = 0.0;
                for (int l = 0; l < nh; ++l) corr += p.T[(size_t)l * m + i] * scoef[l];
                v -= corr;
                p.th[i] = v;
                const double d = v - p.th_old[i];
                dsum += d * d;
                tsum += v * v;
            }
            dsum = block_sum(dsum, sh);
            tsum = block_sum(tsum, sh);
            if (tid == 0) {
                p.part[(size_t)b * PW + 1] = dsum;
                p.part[(size_t)b * PW + 2] = tsum;
            }
            grid.sync();
            double diff = 0.0;
            tt = 0.0;
            for (int bb = 0; bb < nb; ++bb) {
                diff += __ldcg(&p.part[(size_t)bb * PW + 1]);
                tt += __ldcg(&p.part[(size_t)bb * PW + 2]);
            }
            if (diff < p.tol) break;
            ++pciter;
            if (pciter == p.maxiter) {
                if (b == 0 && tid == 0) atomicMax(p.status, 1);
                break;
            }
        }
        // component done: record, then deflate x -= th ph' fused with the next start-column search
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) p.T[(size_t)h * m + i] = p.th[i];
        if (b == 0) {
            for (int64_t j = tid; j < n; j += NIP_THREADS) p.P[(size_t)h * n + j] = phn[j];
            if (tid == 0) { p.eig[h] = tt; p.iters[h] = pciter; }
        }
        nip_deflate(p, p.th, phn, gw, tw, lane);
        grid.sync();
    }
    // eig <- sqrt(eig); scores <- sweep(scores, 2, eig, "/")  (unit-norm score columns)
    for (int h = 0; h < p.ncomp; ++h) {
        const double e = sqrt(__ldcg(&p.eig[h]));
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) p.T[(size_t)h * m + i] /= e;
        if (b == 0 && tid == 0) p.part[h] = e;     // part[] is free now: reuse for the sqrt(eig) output
    }
}


// ---------------------------------------------------------------------------
// Complete data: the inner iteration as two TMA-fed streams over x and two grid barriers
// ---------------------------------------------------------------------------
// One block per SM = eight consumer warps + one producer warp that keeps a shared-memory ring of boxes of x full
// (bulk copies signalled on mbarriers, ~130-190 KB in flight per SM whatever the consumers are doing):
//   loadings pass: the (8-column group, 256-row chunk) boxes of x are cut stream-K style into equal contiguous ranges,
//                  one per block (every box is a full 16 KB TMA box whatever n / blocks is); thread = row of the box,
//                  one accumulator per column; a block leaves one partial of x[, j]'t per group it touched (a group
//                  is shared by two or three blocks at most)                      -> grid barrier 1
//   (every block, redundantly: Gram-Schmidt against the previous loadings, normalisation; its own copy of p)
//   scores pass:   a block owns ROWS (tiles of RB <= 256 rows; one stage = one 2-D box of RB rows x CS columns);
//                  a warp takes whole stages (lane = pairs of rows, 128-bit shared loads), the warps' sums are added in
//                  warp order at the end of the tile; t_new[i] = x[i, ]p / p'p is complete inside the block
//                  (no per-block partial vectors, no fold), with this block's share of T't -> grid barrier 2
//   (first component / no Gram-Schmidt: the owners have already written t and their shares of |dt|^2, |t|^2; otherwise
//    the owners apply t -= T (T't / eig) to their rows and a third barrier follows; every block folds the shares in
//    block order -- identical arithmetic everywhere, so all blocks take the same branch of the convergence test)
// The boxes of the scores pass do not depend on barrier 1, so the producer runs ahead into it while the consumers
// are still reducing / waiting.  Every sum has a fixed order: results are run-to-run identical.
#define NS_CONS 256
#define NS_THREADS (NS_CONS + 32)
#define NS_CB 8                   // columns per box of the loadings pass
#define NS_MAXST 12
#define NS_CHUNK 256              // rows per box of the loadings pass

struct NipStreamParams {
    NipalsParams q;
    int RB, CS, G;                // row tile height (even, <= 256), columns per scores stage, ceil(RB / 64)
    int maxs;                     // partial slots per column group of the loadings pass
    double *ppart;                // [groups][maxs][NS_CB] partial loadings
    int64_t mt;                   // stride of tv (m rounded up to an even number: bulk copies move 16-byte units)
    int nst;                      // ring slots
    unsigned slot_bytes;
    int64_t ntiles;               // row tiles
    unsigned long long *bar;      // grid barrier counter (monotone)
    int *abort_flag;              // set when a bounded wait gave up: the call fails instead of hanging the GPU
    double *th2;                  // m: scores of the current iteration before their Gram-Schmidt step (row owners)
    double *tv;                   // mt: the current score vector (rows written by their owners)
    unsigned long long *dbg;      // optional [8]: nanoseconds block 0 spent per phase (MFB_NIPALS_TIMING)
};

__device__ __forceinline__ uint32_t ns_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool ns_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(ns_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ns_wait(uint64_t *bar, uint32_t parity, int *abort_flag) {
    if (ns_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!ns_try_wait(bar, parity)) {
        if (*(volatile int *)abort_flag) return;
        if (clock64() - t0 > 4000000000ll) { atomicExch(abort_flag, 1); return; }     // ~2 s
    }
}
__device__ __forceinline__ void ns_cons_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ double ns_cons_sum(double v, double *sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    ns_cons_bar();
    if (l == 0) sh[w] = v;
    ns_cons_bar();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NS_CONS / 32; ++i) s += sh[i];
    return s;
}
// sum of the blocks' partials part[bb * PW + idx] by one warp: lane partials in block order, then the fixed shuffle tree
__device__ __forceinline__ double ns_fold(const double *part, int PW, int idx, int nb, int lane) {
    double s = 0.0;
    for (int bb = lane; bb < nb; bb += 32) s += __ldcg(&part[(size_t)bb * PW + idx]);
    return warp_sum(s);
}
// grid-wide barrier of the consumer threads (the launch is cooperative: all blocks are resident)
__device__ __forceinline__ void ns_grid_bar(const NipStreamParams &sp, unsigned long long &target) {
    ns_cons_bar();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(sp.bar, 1ull);
        const long long t0 = clock64();
        while (true) {
            unsigned long long v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(sp.bar) : "memory");
            if (v >= target) break;
            if (*(volatile int *)sp.abort_flag) break;
            if (clock64() - t0 > 4000000000ll) { atomicExch(sp.abort_flag, 1); break; }
        }
        __threadfence();
    }
    ns_cons_bar();
}

// block that owns unit u when U units are cut into G contiguous ranges [floor(c U / G), floor((c + 1) U / G))
__device__ __forceinline__ int ns_owner(long long u, long long U, int G) {
    return (int)(((u + 1) * (long long)G + U - 1) / U) - 1;
}

__global__ void __launch_bounds__(NS_THREADS, 1) nipals_stream_kernel(const __grid_constant__ NipStreamParams sp,
                                                                      const __grid_constant__ CUtensorMap tm1,
                                                                      const __grid_constant__ CUtensorMap tmx) {
    extern __shared__ __align__(128) unsigned char ns_smem[];
    const NipalsParams &p = sp.q;
    const int nst = sp.nst;
    unsigned char *ring = ns_smem;
    double *red = (double *)(ns_smem + (size_t)nst * sp.slot_bytes);          // [8][256]
    uint64_t *full_bar = (uint64_t *)(red + 8 * 256);
    uint64_t *empty_bar = full_bar + NS_MAXST;
    __shared__ double sh[NS_CONS / 32];
    __shared__ double scoef[64];
    __shared__ int sflag;
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t m = p.m, n = p.n, ld = p.ld;
    const int PW = p.ncomp + 4;
    const int RB = sp.RB, CS = sp.CS, G = sp.G;
    // loadings pass: this block's range of (column group, row chunk) boxes, group-major
    const int nrc = (int)((m + NS_CHUNK - 1) / NS_CHUNK);
    const long long U1 = (long long)((n + NS_CB - 1) / NS_CB) * nrc;
    const long long u0 = U1 * b / nb, u1 = U1 * (b + 1) / nb;
    const int nsc = (int)((n + CS - 1) / CS);                                 // stages per row tile

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ns_smem_u32(&full_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ns_smem_u32(&empty_bar[s])), "r"(NS_CONS / 32));   // 8 arrivals
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sflag = 0;
    }
    __syncthreads();

    if (warp == NS_CONS / 32) {
        // ===================== producer warp =====================
        int slot = 0;                                      // ring position of the next stage and its phase parity
        uint32_t par = 0;
        __syncthreads();                                   // (A) pre-processing done
        const double *tvb = sp.tv;
        for (int h = 0; h < p.ncomp; ++h) {
            __syncthreads();                               // (D) the start vector of the component is in tv
            while (true) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");      // x and tv were written with ordinary stores
                    // loadings pass
                    {
                        int g = (int)(u0 / nrc), rc = (int)(u0 - (long long)g * nrc);
                        for (long long u = u0; u < u1; ++u) {
                            ns_wait(&empty_bar[slot], par ^ 1u, sp.abort_flag);
                            // the box of x and the matching 256 entries of t land in the same stage
                            const int64_t row0 = (int64_t)rc * NS_CHUNK;
                            const uint32_t tbytes = (uint32_t)((sp.mt - row0 < NS_CHUNK) ? sp.mt - row0 : NS_CHUNK) * 8u;
                            unsigned char *dst = ring + (size_t)slot * sp.slot_bytes;
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                         ::"r"(ns_smem_u32(&full_bar[slot])), "r"((uint32_t)(NS_CHUNK * NS_CB * 8) + tbytes) : "memory");
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                ::"r"(ns_smem_u32(dst)), "l"(&tm1), "r"(ns_smem_u32(&full_bar[slot])),
                                  "r"(rc * NS_CHUNK), "r"(g * NS_CB) : "memory");
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                ::"r"(ns_smem_u32(dst + NS_CHUNK * NS_CB * 8)), "l"(tvb + row0), "r"(tbytes),
                                  "r"(ns_smem_u32(&full_bar[slot])) : "memory");
                            if (++slot == nst) { slot = 0; par ^= 1u; }
                            if (++rc == nrc) { rc = 0; ++g; }
                        }
                    }
                    // scores pass
                    for (int64_t t = b; t < sp.ntiles; t += nb)
                        for (int sc = 0; sc < nsc; ++sc) {
                            ns_wait(&empty_bar[slot], par ^ 1u, sp.abort_flag);
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                         ::"r"(ns_smem_u32(&full_bar[slot])), "r"((uint32_t)(RB * CS * 8)) : "memory");
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                ::"r"(ns_smem_u32(ring + (size_t)slot * sp.slot_bytes)), "l"(&tmx), "r"(ns_smem_u32(&full_bar[slot])),
                                  "r"((int)(t * RB)), "r"(sc * CS) : "memory");
                            if (++slot == nst) { slot = 0; par ^= 1u; }
                        }
                }
                __syncwarp();
                __syncthreads();                           // (B) the consumers have decided
                if (sflag) break;
            }
            __syncthreads();                               // (C) deflation done everywhere
        }
        return;
    }

    // ===================== consumers =====================
    unsigned long long tmark = 0ull;
#define NS_MARK(slot)                                                                             \
    do {                                                                                          \
        if (sp.dbg && b == 0 && tid == 0) {                                                       \
            unsigned long long now_;                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                              \
            if ((slot) >= 0) sp.dbg[(slot)] += now_ - tmark;                                      \
            tmark = now_;                                                                         \
        }                                                                                         \
    } while (0)
    const int64_t gw = (int64_t)b * NIP_WARPS + warp, tw = (int64_t)nb * NIP_WARPS;
    unsigned long long target = 0ull;
    int slot = 0;                                          // ring position of the next stage and its phase parity
    uint32_t par = 0;
    double *tv = sp.tv;                                    // the current score vector (rows written by their owners)
    double *phn = p.phn + (size_t)b * n;
    __shared__ double sfold[2];
    // (the t region of a stage is only partly rewritten by the last chunk of a column: start from finite values)
    for (int e = tid; e < nst * NS_CHUNK; e += NS_CONS)
        ((double *)(ring + (size_t)(e / NS_CHUNK) * sp.slot_bytes + NS_CHUNK * NS_CB * 8))[e % NS_CHUNK] = 0.0;
    nip_preprocess(p, gw, tw, lane);
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    ns_grid_bar(sp, target);
    __syncthreads();                                       // (A)

    for (int h = 0; h < p.ncomp; ++h) {
        const int64_t scol = nip_start_col<true>(p, tid);
        // t = x[, scol], every row by its owner; tt = t't
        {
            double tts = 0.0;
            for (int64_t t = b; t < sp.ntiles; t += nb) {
                const int64_t row = t * RB + tid;
                if (tid < RB && row < m) {
                    const double v = __ldcg(&p.x[scol * ld + row]);
                    tv[row] = v;
                    tts += v * v;
                }
            }
            tts = ns_cons_sum(tts, sh);
            if (tid == 0) p.part[(size_t)b * PW + 0] = tts;
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        ns_grid_bar(sp, target);
        if (warp == 0) { const double f = ns_fold(p.part, PW, 0, nb, lane); if (lane == 0) sfold[0] = f; }
        __syncthreads();                                   // (D)
        double tt = sfold[0];
        const int nh = (p.gramschmidt && h > 0) ? h : 0;
        int pciter = 1;
        while (true) {
            NS_MARK(-1);
            // ---- loadings pass ----
            if (u0 < u1) {
                int g = (int)(u0 / nrc), rc = (int)(u0 - (long long)g * nrc);
                double acc[NS_CB];
#pragma unroll
                for (int cc = 0; cc < NS_CB; ++cc) acc[cc] = 0.0;
                for (long long u = u0; u < u1; ++u) {
                    const int rcn = (rc + 1 == nrc) ? 0 : rc + 1;
                    ns_wait(&full_bar[slot], par, sp.abort_flag);
                    const double *xs = (const double *)(ring + (size_t)slot * sp.slot_bytes) + tid;
#ifndef NS_NOCOMPUTE
                    const double tcur = xs[NS_CB * NS_CHUNK];
#pragma unroll
                    for (int cc = 0; cc < NS_CB; ++cc) acc[cc] += xs[cc * NS_CHUNK] * tcur;      // rows >= m arrive as zeros
#endif
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ns_smem_u32(&empty_bar[slot])) : "memory");
                    if (++slot == nst) { slot = 0; par ^= 1u; }
                    if (rcn == 0 || u + 1 == u1) {
                        // last box of this block in group g: leave the block's partial in its slot of the group
#pragma unroll
                        for (int cc = 0; cc < NS_CB; ++cc) acc[cc] = warp_sum(acc[cc]);
                        ns_cons_bar();
                        if (lane == 0) {
#pragma unroll
                            for (int cc = 0; cc < NS_CB; ++cc) red[warp * NS_CB + cc] = acc[cc];
                        }
                        ns_cons_bar();
                        if (tid < NS_CB) {
                            double sres = 0.0;
#pragma unroll
                            for (int w = 0; w < NS_CONS / 32; ++w) sres += red[w * NS_CB + tid];
                            const int sl = b - ns_owner((long long)g * nrc, U1, nb);
                            sp.ppart[((size_t)g * sp.maxs + sl) * NS_CB + tid] = sres;
                        }
#pragma unroll
                        for (int cc = 0; cc < NS_CB; ++cc) acc[cc] = 0.0;
                        ++g;
                    }
                    rc = rcn;
                }
            }
            NS_MARK(0);
            __threadfence();
            ns_grid_bar(sp, target);
            NS_MARK(1);
            // ---- every block: Gram-Schmidt against P[, 0..h), normalise, keep its own copy of the loading ----
            double pp_den;
            {
                // p_raw[j] = (the blocks' partials of x[, j]'t, in block order; unused slots hold zeros) / t't
                const double *__restrict__ Pm = p.P;
                const double *__restrict__ ppart = sp.ppart;
                const int maxs = sp.maxs;
                {
                    int64_t j = tid;
                    for (; j + 3 * NS_CONS < n; j += 4 * NS_CONS) {
                        double v[4] = {0.0, 0.0, 0.0, 0.0};
                        const double *pp[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int64_t jj = j + u * NS_CONS, g = jj / NS_CB;
                            pp[u] = ppart + ((size_t)g * maxs) * NS_CB + (jj - g * NS_CB);
                        }
                        for (int q = 0; q < maxs; ++q) {
                            double a[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) a[u] = __ldcg(&pp[u][(size_t)q * NS_CB]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] += a[u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) phn[j + u * NS_CONS] = v[u] / tt;
                    }
                    for (; j < n; j += NS_CONS) {
                        const int64_t g = j / NS_CB;
                        const double *pp1 = ppart + ((size_t)g * maxs) * NS_CB + (j - g * NS_CB);
                        double v = 0.0;
                        for (int q = 0; q < maxs; ++q) v += __ldcg(&pp1[(size_t)q * NS_CB]);
                        phn[j] = v / tt;
                    }
                }
                for (int l = 0; l < nh; ++l) {
                    double sdot = 0.0;
                    int64_t j = tid;
                    for (; j + 3 * NS_CONS < n; j += 4 * NS_CONS) {
                        double a[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) a[u] = __ldcg(&Pm[(size_t)l * n + j + u * NS_CONS]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) sdot += a[u] * phn[j + u * NS_CONS];
                    }
                    for (; j < n; j += NS_CONS) sdot += __ldcg(&Pm[(size_t)l * n + j]) * phn[j];
                    sdot = ns_cons_sum(sdot, sh);
                    if (tid == 0) scoef[l] = sdot;
                }
                ns_cons_bar();
                // corrected loading, then its norm, then the scaling (a thread only ever touches its own entries of phn)
                double nrm = 0.0;
                {
                    int64_t j = tid;
                    for (; j + 3 * NS_CONS < n; j += 4 * NS_CONS) {
                        double v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = phn[j + u * NS_CONS];
                        double corr[4] = {0.0, 0.0, 0.0, 0.0};
                        for (int l = 0; l < nh; ++l) {
                            const double cf = scoef[l];
                            double pl[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) pl[u] = __ldcg(&Pm[(size_t)l * n + j + u * NS_CONS]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) corr[u] += pl[u] * cf;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) { v[u] -= corr[u]; phn[j + u * NS_CONS] = v[u]; nrm += v[u] * v[u]; }
                    }
                    for (; j < n; j += NS_CONS) {
                        double v = phn[j];
                        double corr = 0.0;
                        for (int l = 0; l < nh; ++l) corr += __ldcg(&Pm[(size_t)l * n + j]) * scoef[l];
                        v -= corr;
                        phn[j] = v;
                        nrm += v * v;
                    }
                }
                nrm = sqrt(ns_cons_sum(nrm, sh));
                double p2 = 0.0;
                for (int64_t j = tid; j < n; j += NS_CONS) {      // (a thread re-reads only what it wrote itself)
                    const double v = phn[j] / nrm;
                    p2 += v * v;
                    phn[j] = v;
                }
                pp_den = ns_cons_sum(p2, sh);               // (its barriers also publish phn to the other warps)
            }
            NS_MARK(2);
            // ---- scores pass ----
            double tdot = 0.0;                              // thread l < nh: this block's share of T[, l]'t
            double dsum = 0.0, tsum = 0.0;                  // this thread's share of |t - t_old|^2 and |t|^2
            for (int64_t t = b; t < sp.ntiles; t += nb) {
                // a warp takes the stages that land in ITS ring slots (slot mod 8): it then sees every phase of those slots'
                // barriers in order -- a parity wait by a warp that skipped a phase would succeed one revolution early;
                // lane = pairs of rows (2 q, 2 q + 1), q = lane + 32 g
                double2 acc[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[g] = make_double2(0.0, 0.0);
                for (int sc = 0; sc < nsc; ++sc) {
                    if ((slot & 7) == warp) {
                        const int64_t j0 = (int64_t)sc * CS;
                        const double pjl = (lane < CS && j0 + lane < n) ? phn[j0 + lane] : 0.0;
                        ns_wait(&full_bar[slot], par, sp.abort_flag);
#ifndef NS_NOCOMPUTE
                        const uint32_t xa = ns_smem_u32(ring + (size_t)slot * sp.slot_bytes) + (uint32_t)lane * 16u;
                        for (int cc = 0; cc < CS; ++cc) {
                            const double pj = __shfl_sync(0xffffffffu, pjl, cc);
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                if (g < G && 2 * (lane + 32 * g) < RB) {
                                    double2 v;
                                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y)
                                                 : "r"(xa + (uint32_t)((cc * RB + 64 * g) * 8)));
                                    acc[g].x += v.x * pj;
                                    acc[g].y += v.y * pj;
                                }
                        }
#endif
                        __syncwarp();
                        if (lane < NS_CONS / 32)
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ns_smem_u32(&empty_bar[slot])) : "memory");
                    }
                    if (++slot == nst) { slot = 0; par ^= 1u; }
                }
                ns_cons_bar();
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (g < G && 2 * (lane + 32 * g) < RB) {
                        red[warp * 256 + 2 * (lane + 32 * g)] = acc[g].x;
                        red[warp * 256 + 2 * (lane + 32 * g) + 1] = acc[g].y;
                    }
                ns_cons_bar();
                const int64_t row = t * RB + tid;
                const bool mine = tid < RB && row < m;
                double tn = 0.0;
                if (mine) {
#pragma unroll
                    for (int w = 0; w < NS_CONS / 32; ++w) tn += red[w * 256 + tid];
                    tn /= pp_den;
                    if (nh == 0) {                          // no Gram-Schmidt on the scores: the row is final
                        const double d = tn - tv[row];
                        tv[row] = tn;
                        dsum += d * d;
                        tsum += tn * tn;
                    } else {
                        sp.th2[row] = tn;
                    }
                }
                for (int l = 0; l < nh; ++l) {
                    const double s = ns_cons_sum(mine ? __ldcg(&p.T[(size_t)l * m + row]) * tn : 0.0, sh);
                    if (tid == l) tdot += s;
                }
            }
            if (tid < nh) p.part[(size_t)b * PW + 4 + tid] = tdot;
            if (nh == 0) {
                dsum = ns_cons_sum(dsum, sh);
                tsum = ns_cons_sum(tsum, sh);
                if (tid == 0) { p.part[(size_t)b * PW + 1] = dsum; p.part[(size_t)b * PW + 2] = tsum; }
            }
            NS_MARK(3);
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");      // (the next loadings pass reads tv through TMA)
            ns_grid_bar(sp, target);
            NS_MARK(4);
            if (nh > 0) {
                // ---- t -= T (T't / eig), every row by its owner; then |t - t_old|^2, |t|^2 ----
                for (int l = warp; l < nh; l += NS_CONS / 32) {
                    const double f = ns_fold(p.part, PW, 4 + l, nb, lane);
                    if (lane == 0) scoef[l] = f / __ldcg(&p.eig[l]);
                }
                ns_cons_bar();
                dsum = tsum = 0.0;
                for (int64_t t = b; t < sp.ntiles; t += nb) {
                    const int64_t row = t * RB + tid;
                    if (tid < RB && row < m) {
                        double v = sp.th2[row];
                        double corr = 0.0;
                        for (int l = 0; l < nh; ++l) corr += __ldcg(&p.T[(size_t)l * m + row]) * scoef[l];
                        v -= corr;
                        const double d = v - tv[row];
                        tv[row] = v;
                        dsum += d * d;
                        tsum += v * v;
                    }
                }
                dsum = ns_cons_sum(dsum, sh);
                tsum = ns_cons_sum(tsum, sh);
                if (tid == 0) { p.part[(size_t)b * PW + 1] = dsum; p.part[(size_t)b * PW + 2] = tsum; }
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                ns_grid_bar(sp, target);
            }
            if (warp < 2) { const double f = ns_fold(p.part, PW, 1 + warp, nb, lane); if (lane == 0) sfold[warp] = f; }
            ns_cons_bar();
            const double diff = sfold[0];
            tt = sfold[1];
            NS_MARK(5);
            bool stop = diff < p.tol;
            if (!stop) {
                ++pciter;
                if (pciter == p.maxiter) {
                    if (b == 0 && tid == 0) atomicMax(p.status, 1);
                    stop = true;
                }
            }
            if (*(volatile int *)sp.abort_flag) stop = true;
            if (tid == 0) sflag = stop ? 1 : 0;
            __syncthreads();                               // (B)  (also orders the reads of sfold before its next writes)
            if (stop) break;
        }
        // component done: record, then deflate x -= t p' fused with the next start-column search
        for (int64_t t = b; t < sp.ntiles; t += nb) {
            const int64_t row = t * RB + tid;
            if (tid < RB && row < m) p.T[(size_t)h * m + row] = tv[row];
        }
        if (b == 0) {
            for (int64_t j = tid; j < n; j += NS_CONS) p.P[(size_t)h * n + j] = phn[j];
            if (tid == 0) { p.eig[h] = tt; p.iters[h] = pciter; }
        }
        if (h + 1 < p.ncomp) nip_deflate(p, tv, phn, gw, tw, lane);
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        ns_grid_bar(sp, target);
        __syncthreads();                                   // (C)
    }
    // eig <- sqrt(eig); scores <- sweep(scores, 2, eig, "/")  (unit-norm score columns)
    for (int h = 0; h < p.ncomp; ++h) {
        const double e = sqrt(__ldcg(&p.eig[h]));
        for (int64_t t = b; t < sp.ntiles; t += nb) {
            const int64_t row = t * RB + tid;
            if (tid < RB && row < m) p.T[(size_t)h * m + row] /= e;
        }
        if (b == 0 && tid == 0) p.part[h] = e;
    }
}

// NIPALS on a device matrix xd (m x n, leading dimension ld; untouched: a working copy is deflated).  Results stay
// on the device: T (m x ncomp, ld = m, unit columns), P (n x ncomp, ld = n), eig / iters / status through the pinned
// mailbox of the context (eig_host, iters_host: ncomp entries; may be NULL).  Returns MFB_OK, MFB_NOT_CONVERGED or an error.
static int nipals_collect(mfb_context *ctx, const NipalsParams &p, int ncomp, int *ibuf, double *T_out, double *P_out,
                          double *eig_host, int *iters_host) {
    cudaStream_t st = ctx->stream;
    const int64_t m = p.m, n = p.n;
    std::vector<double> hE(ncomp);
    std::vector<int> hI(ncomp + 2);
    MFB_CUDA(cudaMemcpyAsync(T_out, p.T, (size_t)m * ncomp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(P_out, p.P, (size_t)n * ncomp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(hE.data(), p.part, ncomp * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(hI.data(), ibuf, (ncomp + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    for (int h = 0; h < ncomp; ++h) {
        if (eig_host) eig_host[h] = hE[h];
        if (iters_host) iters_host[h] = hI[h];
    }
    if (hI[ncomp + 1] != 0) {
        mfb_set_error("mfb_nipals: a wait inside nipals_stream_kernel timed out (protocol error)");
        return MFB_ERR_CUDA;
    }
    if (hI[ncomp] == 2) {
        mfb_set_error("mfb_nipals: a column has zero variance and scale = TRUE");
        return MFB_ERR_ARG;
    }
    return hI[ncomp] == 1 ? MFB_NOT_CONVERGED : MFB_OK;
}

// complete data: nipals_stream_kernel on a working copy with an even leading dimension
static int nipals_stream_on_device(mfb_context *ctx, const double *xd, int64_t m, int64_t n, int64_t ld, int ncomp, int center,
                                   int scale, int maxiter, double tol, int gramschmidt, double *T_out, double *P_out,
                                   double *eig_host, int *iters_host) {
    cudaStream_t st = ctx->stream;
    const int nb = ctx->sm_count;
    const int64_t ldw = round_up(m, 2);
    // row tiles of the scores pass: an even height <= 256, as close to one (or a few) tiles per block as m allows
    const int64_t tpb = (m + (int64_t)nb * 256 - 1) / ((int64_t)nb * 256);
    int64_t RB = round_up((m + nb * tpb - 1) / (nb * tpb), 2);
    if (RB < 2) RB = 2;
    if (RB > 256) RB = 256;
    const int64_t ntiles = (m + RB - 1) / RB;
    const int CS = (RB <= 128) ? 32 : 16;
    // loadings pass: boxes of 256 rows x 8 columns; partial slots per column group under the stream-K cut
    const int64_t nrc = (m + NS_CHUNK - 1) / NS_CHUNK, ng1 = (n + NS_CB - 1) / NS_CB;
    const long long U1 = (long long)ng1 * nrc;
    auto owner = [&](long long u) { return (int)(((u + 1) * (long long)nb + U1 - 1) / U1) - 1; };
    int maxs = 1;
    for (int64_t g = 0; g < ng1; ++g) {
        const int ns_ = owner(g * nrc + nrc - 1) - owner(g * nrc) + 1;
        if (ns_ > maxs) maxs = ns_;
    }
    const int64_t mt = round_up(m, 2);
    size_t slot = (size_t)NS_CHUNK * NS_CB * 8 + NS_CHUNK * 8;      // a box of x + 256 entries of t
    if ((size_t)RB * CS * 8 > slot) slot = (size_t)RB * CS * 8;
    slot = (size_t)round_up((int64_t)slot, 128);
    const size_t fixed = 8 * 256 * sizeof(double) + 2 * NS_MAXST * sizeof(uint64_t);
    int nst = (int)((216 * 1024 - fixed) / slot);
    if (nst > NS_MAXST) nst = NS_MAXST;
    MFB_ARG(nst >= 2, "nipals: shared-memory ring does not fit");
    const size_t smem = (size_t)nst * slot + fixed;
    MFB_CUDA(cudaFuncSetAttribute(nipals_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nipals_stream_kernel, NS_THREADS, smem));
    MFB_ARG(per_sm >= 1, "nipals: the streaming kernel does not fit an SM");
    const size_t PW = ncomp + 4;
    const size_t nd = (size_t)ldw * n + (size_t)m * ncomp + (size_t)n * ncomp + ncomp + (size_t)m + 2 * (size_t)n +
                      (size_t)nb * PW + (size_t)nb * n + (size_t)mt + 2 + 8 + (size_t)ng1 * maxs * NS_CB;
    DevBuf pool(ctx);
    double *buf = nullptr;
    int *ibuf = nullptr;
    {
        int rc2 = pool.alloc(&buf, nd);
        if (rc2 == MFB_OK) rc2 = pool.alloc(&ibuf, (size_t)ncomp + 2);
        if (rc2 != MFB_OK) return rc2;
    }
    if (ldw != m) MFB_CUDA(cudaMemsetAsync(buf, 0, (size_t)ldw * n * sizeof(double), st));      // the padding row
    MFB_CUDA(cudaMemsetAsync(buf + (size_t)ldw * n, 0, (nd - (size_t)ldw * n) * sizeof(double), st));
    MFB_CUDA(cudaMemsetAsync(ibuf, 0, (ncomp + 2) * sizeof(int), st));
    NipStreamParams sp;
    NipalsParams &p = sp.q;
    p.x = buf;
    sp.tv = p.x + (size_t)ldw * n;                 // (ldw is even: tv stays 16-byte aligned for the bulk copies)
    sp.mt = mt;
    p.T = sp.tv + (size_t)mt;
    p.P = p.T + (size_t)m * ncomp;
    p.eig = p.P + (size_t)n * ncomp;
    sp.th2 = p.eig + ncomp;
    p.ph = sp.th2 + m;
    p.colabs = p.ph + n;
    p.part = p.colabs + n;
    p.phn = p.part + (size_t)nb * PW;
    sp.bar = (unsigned long long *)(p.phn + (size_t)nb * n);
    static const bool timing = getenv("MFB_NIPALS_TIMING") != nullptr;
    sp.dbg = timing ? sp.bar + 2 : nullptr;
    sp.ppart = (double *)(sp.bar + 2 + 8);
    sp.maxs = maxs;
    p.th = p.th_old = p.spart = nullptr;
    p.chunk = 0;
    p.iters = ibuf;
    p.status = ibuf + ncomp;
    sp.abort_flag = ibuf + ncomp + 1;
    p.m = m; p.n = n; p.ld = ldw;
    p.ncomp = ncomp; p.center = center; p.scale = scale; p.maxiter = maxiter; p.gramschmidt = gramschmidt;
    p.tol = tol;
    p.has_na = 0;
    sp.RB = (int)RB; sp.CS = CS; sp.G = (int)((RB + 63) / 64);
    sp.nst = nst; sp.slot_bytes = (unsigned)slot; sp.ntiles = ntiles;
    if (ld == ldw)
        MFB_CUDA(cudaMemcpyAsync(p.x, xd, (size_t)ld * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    else
        MFB_CUDA(cudaMemcpy2DAsync(p.x, (size_t)ldw * sizeof(double), xd, (size_t)ld * sizeof(double), (size_t)m * sizeof(double),
                                   (size_t)n, cudaMemcpyDeviceToDevice, st));
    CUtensorMap tm1, tmx;
    {
        mfb_encode_tiled_fn enc = mfb_get_encode_tiled();
        if (!enc) { mfb_set_error("cuTensorMapEncodeTiled unavailable"); return MFB_ERR_CUDA; }
        cuuint64_t dims[2] = {(cuuint64_t)m, (cuuint64_t)n};
        cuuint64_t strides[1] = {(cuuint64_t)ldw * sizeof(double)};
        cuuint32_t estr[2] = {1, 1};
        cuuint32_t box1[2] = {NS_CHUNK, NS_CB}, box3[2] = {(cuuint32_t)RB, (cuuint32_t)CS};
        CUresult r = enc(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, p.x, dims, strides, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, p.x, dims, strides, box3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mfb_set_error("cuTensorMapEncodeTiled (nipals) failed with CUresult %d", (int)r); return MFB_ERR_CUDA; }
    }
    void *args[] = {&sp, &tm1, &tmx};
    MFB_CUDA(cudaLaunchCooperativeKernel((void *)nipals_stream_kernel, dim3(nb), dim3(NS_THREADS), args, smem, st));
    if (timing) {
        unsigned long long hd[8];
        MFB_CUDA(cudaMemcpyAsync(hd, sp.dbg, sizeof(hd), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "nipals_stream (block 0, us): loadings %.1f bar1 %.1f gs_p %.1f scores %.1f bar2 %.1f gs_t %.1f | RB %d CS %d nst %d slot %u\n",
                hd[0] * 1e-3, hd[1] * 1e-3, hd[2] * 1e-3, hd[3] * 1e-3, hd[4] * 1e-3, hd[5] * 1e-3, (int)RB, CS, nst, (unsigned)slot);
    }
    return nipals_collect(ctx, p, ncomp, ibuf, T_out, P_out, eig_host, iters_host);
}

int nipals_on_device(mfb_context *ctx, const double *xd, int64_t m, int64_t n, int64_t ld, int has_na, int ncomp,
                     int center, int scale, int maxiter, double tol, int gramschmidt, double *T_out, double *P_out,
                     double *eig_host, int *iters_host) {
    MFB_ARG(ncomp >= 1 && ncomp <= 60, "ncomp must be in 1..60");
    MFB_ARG(ncomp <= m && ncomp <= n, "ncomp exceeds a matrix dimension");
    MFB_ARG(maxiter >= 2, "maxiter must be at least 2");
    static const bool old_kernel = getenv("MFB_NIPALS_OLD") != nullptr;      // development A/B
    if (!has_na && !old_kernel)
        return nipals_stream_on_device(ctx, xd, m, n, ld, ncomp, center, scale, maxiter, tol, gramschmidt, T_out, P_out, eig_host,
                                       iters_host);
    cudaStream_t st = ctx->stream;
    // shared memory: the 256-row pass scratch of the missing-value S3, or the score accumulator of the complete-data
    // S3 (the rows are cut into equal chunks of at most 12288 = 96 KB, so that two blocks still fit an SM)
    const int64_t nchunk = (m + 12287) / 12288;
    const int chunk = (int)(((m + nchunk - 1) / nchunk + NIP_THREADS - 1) / NIP_THREADS * NIP_THREADS);
    size_t smem = 2 * NIP_WARPS * 256 * sizeof(double);
    if (!has_na && (size_t)chunk * sizeof(double) > smem) smem = (size_t)chunk * sizeof(double);
    MFB_CUDA(cudaFuncSetAttribute(nipals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb_max = 0;
    MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_max, nipals_kernel, NIP_THREADS, smem));
    int nb = nb_max * ctx->sm_count;
    if (nb > 2 * ctx->sm_count) nb = 2 * ctx->sm_count;       // two blocks per SM: twice the loads in flight
    const int64_t want = (m + NIP_THREADS - 1) / NIP_THREADS > n ? (m + NIP_THREADS - 1) / NIP_THREADS : n;
    if (want < nb) nb = (int)(want < 1 ? 1 : want);
    const size_t PW = ncomp + 4;
    size_t nd = (size_t)ld * n + (size_t)m * ncomp + (size_t)n * ncomp + ncomp + 2 * (size_t)m + 2 * (size_t)n + (size_t)nb * PW +
                (size_t)nb * n + (has_na ? 0 : (size_t)nb * m);
    DevBuf pool(ctx);
    double *buf = nullptr;
    int *ibuf = nullptr;
    {
        int rc2 = pool.alloc(&buf, nd);
        if (rc2 == MFB_OK) rc2 = pool.alloc(&ibuf, (size_t)ncomp + 2);
        if (rc2 != MFB_OK) return rc2;
    }
    // (the working copy of x, the first ld * n doubles, is overwritten below: zero only what follows it)
    MFB_CUDA(cudaMemsetAsync(buf + (size_t)ld * n, 0, (nd - (size_t)ld * n) * sizeof(double), st));
    MFB_CUDA(cudaMemsetAsync(ibuf, 0, (ncomp + 2) * sizeof(int), st));
    NipalsParams p;
    p.x = buf;
    p.T = p.x + (size_t)ld * n;
    p.P = p.T + (size_t)m * ncomp;
    p.eig = p.P + (size_t)n * ncomp;
    p.th = p.eig + ncomp;
    p.th_old = p.th + m;
    p.ph = p.th_old + m;
    p.colabs = p.ph + n;
    p.part = p.colabs + n;
    p.phn = p.part + (size_t)nb * PW;
    p.spart = p.phn + (size_t)nb * n;
    p.chunk = chunk;
    p.iters = ibuf;
    p.status = ibuf + ncomp;
    p.m = m; p.n = n; p.ld = ld;
    p.ncomp = ncomp; p.center = center; p.scale = scale; p.maxiter = maxiter; p.gramschmidt = gramschmidt;
    p.tol = tol;
    p.has_na = has_na;
    MFB_CUDA(cudaMemcpyAsync(p.x, xd, (size_t)ld * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    void *args[] = {&p};
    MFB_CUDA(cudaLaunchCooperativeKernel((void *)nipals_kernel, dim3(nb), dim3(NIP_THREADS), args, smem, st));
    return nipals_collect(ctx, p, ncomp, ibuf, T_out, P_out, eig_host, iters_host);
}

extern "C" {

int mfb_nipals(mfb_context *ctx, mfb_matrix *x, int ncomp, int center, int scale, int maxiter, double tol,
               int gramschmidt, double *scores, double *loadings, double *eig, int *iters) {
    MFB_ARG(ctx && x && scores && loadings, "NULL argument");
    if (x->elem_bytes != 8) { mfb_set_error("this entry point needs an FP64 resident matrix"); return MFB_ERR_UNSUPPORTED; }
    MFB_ARG(ncomp >= 1 && ncomp <= 60, "ncomp must be in 1..60");
    MFB_CUDA(cudaSetDevice(ctx->device));
    double stats[5];
    int rc = mfb_matrix_stats(x, stats);
    if (rc != MFB_OK) return rc;
    const int64_t m = x->m, n = x->n;
    DevBuf pool(ctx);
    double *dT = nullptr, *dP = nullptr;
    if ((rc = pool.alloc(&dT, (size_t)m * ncomp)) != MFB_OK || (rc = pool.alloc(&dP, (size_t)n * ncomp)) != MFB_OK) return rc;
    const int st_ = nipals_on_device(ctx, x->d, m, n, x->lda, stats[3] > 0, ncomp, center, scale, maxiter, tol, gramschmidt,
                                     dT, dP, eig, iters);
    if (st_ < 0) return st_;
    MFB_CUDA(cudaMemcpyAsync(scores, dT, (size_t)m * ncomp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(loadings, dP, (size_t)n * ncomp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return st_;
}

}  // extern "C"
