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

    // ---- pre-processing: per column mean / sd (about the mean, n-1 denominator), one warp per column ----
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
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t i = lane;
            for (; i + 96 < m; i += 128) { s0 += col[i]; s1 += col[i + 32]; s2 += col[i + 64]; s3 += col[i + 96]; }
            for (; i < m; i += 32) s0 += col[i];
            const double mean = warp_sum((s0 + s1) + (s2 + s3)) / (double)m;
            double sd = 1.0;
            if (p.scale) {
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                i = lane;
                for (; i + 96 < m; i += 128) {
                    const double d0 = col[i] - mean, d1 = col[i + 32] - mean, d2 = col[i + 64] - mean, d3 = col[i + 96] - mean;
                    q0 += d0 * d0; q1 += d1 * d1; q2 += d2 * d2; q3 += d3 * d3;
                }
                for (; i < m; i += 32) { const double d0 = col[i] - mean; q0 += d0 * d0; }
                sd = sqrt(warp_sum((q0 + q1) + (q2 + q3)) / (double)(m - 1));
                if (!(sd > 0.0) && lane == 0) atomicExch(p.status, 2);
            }
            double ab = 0.0;
            for (i = lane; i < m; i += 32) {
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
    grid.sync();

    // rows of the scores owned by this block: a multiple of 32 so that a warp always reads 256 contiguous bytes
    const int64_t rows_per = ((m + nb - 1) / nb + 31) / 32 * 32;
    const int64_t r0 = (b * rows_per < m) ? b * rows_per : m, r1 = (r0 + rows_per < m) ? r0 + rows_per : m;

    for (int h = 0; h < p.ncomp; ++h) {
        // start column: which.max(colSums(abs(x))) -- first maximum; every block finds it redundantly
        int64_t scol = 0;
        {
            double best = -1.0;
            int64_t bj = 0;
            for (int64_t j = tid; j < n; j += NIP_THREADS) {
                const double v = __ldcg(&p.colabs[j]);
                if (v > best) { best = v; bj = j; }
            }
            // block arg-max, ties -> smallest index
            __shared__ double sbv[NIP_THREADS];
            __shared__ long long sbj[NIP_THREADS];
            sbv[tid] = best; sbj[tid] = bj;
            __syncthreads();
            for (int o = NIP_THREADS / 2; o > 0; o >>= 1) {
                if (tid < o) {
                    const double ov = sbv[tid + o];
                    const long long oj = sbj[tid + o];
                    if (ov > sbv[tid] || (ov == sbv[tid] && oj < sbj[tid])) { sbv[tid] = ov; sbj[tid] = oj; }
                }
                __syncthreads();
            }
            scol = sbj[0];
            __syncthreads();
        }
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
        for (int64_t j = gw; j < n; j += tw) {
            double *col = p.x + j * ld;
            const double pj = phn[j];
            double s0 = 0.0, s1 = 0.0;
            int64_t i = lane;
            for (; i + 96 < m; i += 128) {
                const double t0 = __ldcg(&p.th[i]), t1 = __ldcg(&p.th[i + 32]), t2 = __ldcg(&p.th[i + 64]), t3 = __ldcg(&p.th[i + 96]);
                const double v0 = col[i] - t0 * pj, v1 = col[i + 32] - t1 * pj, v2 = col[i + 64] - t2 * pj, v3 = col[i + 96] - t3 * pj;
                col[i] = v0; col[i + 32] = v1; col[i + 64] = v2; col[i + 96] = v3;
                s0 += ((v0 == v0) ? fabs(v0) : 0.0) + ((v2 == v2) ? fabs(v2) : 0.0);
                s1 += ((v1 == v1) ? fabs(v1) : 0.0) + ((v3 == v3) ? fabs(v3) : 0.0);
            }
            for (; i < m; i += 32) {
                const double v0 = col[i] - __ldcg(&p.th[i]) * pj;
                col[i] = v0;
                s0 += (v0 == v0) ? fabs(v0) : 0.0;
            }
            const double ab = warp_sum(s0 + s1);
            if (lane == 0) p.colabs[j] = ab;
        }
        grid.sync();
    }
    // eig <- sqrt(eig); scores <- sweep(scores, 2, eig, "/")  (unit-norm score columns)
    for (int h = 0; h < p.ncomp; ++h) {
        const double e = sqrt(__ldcg(&p.eig[h]));
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) p.T[(size_t)h * m + i] /= e;
        if (b == 0 && tid == 0) p.part[h] = e;     // part[] is free now: reuse for the sqrt(eig) output
    }
}

// NIPALS on a device matrix xd (m x n, leading dimension ld; untouched: a working copy is deflated).  Results stay
// on the device: T (m x ncomp, ld = m, unit columns), P (n x ncomp, ld = n), eig / iters / status through the pinned
// mailbox of the context (eig_host, iters_host: ncomp entries; may be NULL).  Returns MFB_OK, MFB_NOT_CONVERGED or an error.
int nipals_on_device(mfb_context *ctx, const double *xd, int64_t m, int64_t n, int64_t ld, int has_na, int ncomp,
                     int center, int scale, int maxiter, double tol, int gramschmidt, double *T_out, double *P_out,
                     double *eig_host, int *iters_host) {
    MFB_ARG(ncomp >= 1 && ncomp <= 60, "ncomp must be in 1..60");
    MFB_ARG(ncomp <= m && ncomp <= n, "ncomp exceeds a matrix dimension");
    MFB_ARG(maxiter >= 2, "maxiter must be at least 2");
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
        if (rc2 == MFB_OK) rc2 = pool.alloc(&ibuf, (size_t)ncomp + 1);
        if (rc2 != MFB_OK) return rc2;
    }
    // (the working copy of x, the first ld * n doubles, is overwritten below: zero only what follows it)
    MFB_CUDA(cudaMemsetAsync(buf + (size_t)ld * n, 0, (nd - (size_t)ld * n) * sizeof(double), st));
    MFB_CUDA(cudaMemsetAsync(ibuf, 0, (ncomp + 1) * sizeof(int), st));
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
    std::vector<double> hE(ncomp);
    std::vector<int> hI(ncomp + 1);
    MFB_CUDA(cudaMemcpyAsync(T_out, p.T, (size_t)m * ncomp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(P_out, p.P, (size_t)n * ncomp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(hE.data(), p.part, ncomp * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(hI.data(), ibuf, (ncomp + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    for (int h = 0; h < ncomp; ++h) {
        if (eig_host) eig_host[h] = hE[h];
        if (iters_host) iters_host[h] = hI[h];
    }
    if (hI[ncomp] == 2) {
        mfb_set_error("mfb_nipals: a column has zero variance and scale = TRUE");
        return MFB_ERR_ARG;
    }
    return hI[ncomp] == 1 ? MFB_NOT_CONVERGED : MFB_OK;
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
