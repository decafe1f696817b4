// NIPALS PCA on the GPU: replaces nipals::nipals (nipals 0.5; call site R/bicluster.R:226) for complete data.
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
    int ncomp, center, scale, maxiter, gramschmidt;
    double tol;
    double *T;            // m x ncomp unnormalised scores
    double *P;            // n x ncomp loadings
    double *eig;          // ncomp: sum(th^2)
    int *iters;           // ncomp
    double *th, *th_old;  // m
    double *ph;           // n  (raw x'th, then the normalised loading)
    double *colabs;       // n
    double *part;         // [nblocks][ncomp + 4] per-block partials
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

__global__ void __launch_bounds__(NIP_THREADS) nipals_kernel(NipalsParams p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[NIP_THREADS / 32];
    __shared__ double scoef[64];
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int64_t m = p.m, n = p.n, ld = p.ld;
    const int PW = p.ncomp + 4;      // width of a block's partial record

    // ---- pre-processing: per column mean / sd (about the mean, n-1 denominator), one block per column ----
    if (p.center || p.scale) {
        for (int64_t j = b; j < n; j += nb) {
            double *col = p.x + j * ld;
            double s = 0.0;
            for (int64_t i = tid; i < m; i += NIP_THREADS) s += col[i];
            const double mean = block_sum(s, sh) / (double)m;
            double sd = 1.0;
            if (p.scale) {
                double q = 0.0;
                for (int64_t i = tid; i < m; i += NIP_THREADS) { const double d = col[i] - mean; q += d * d; }
                sd = sqrt(block_sum(q, sh) / (double)(m - 1));
                if (!(sd > 0.0) && tid == 0) atomicExch(p.status, 2);
            }
            for (int64_t i = tid; i < m; i += NIP_THREADS) {
                double v = col[i];
                if (p.center) v -= mean;
                if (p.scale) v /= sd;
                col[i] = v;
            }
        }
        grid.sync();
    }
    // column |.| sums of the starting matrix
    for (int64_t j = b; j < n; j += nb) {
        const double *col = p.x + j * ld;
        double s = 0.0;
        for (int64_t i = tid; i < m; i += NIP_THREADS) s += fabs(col[i]);
        s = block_sum(s, sh);
        if (tid == 0) p.colabs[j] = s;
    }
    grid.sync();

    const int64_t rows_per = (m + nb - 1) / nb;
    const int64_t r0 = b * rows_per, r1 = (r0 + rows_per < m) ? r0 + rows_per : m;

    for (int h = 0; h < p.ncomp; ++h) {
        // start column: which.max(colSums(abs(x))) -- first maximum; every block finds it redundantly
        int64_t scol = 0;
        {
            double best = -1.0;
            for (int64_t j = 0; j < n; ++j) {
                const double v = p.colabs[j];
                if (v > best) { best = v; scol = j; }
            }
        }
        // th = x[, scol]; tt = sum(th^2)
        double tts = 0.0;
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) {
            const double v = p.x[scol * ld + i];
            p.th[i] = v;
            tts += v * v;
        }
        tts = block_sum(tts, sh);
        if (tid == 0) p.part[(size_t)b * PW + 0] = tts;
        grid.sync();
        double tt = 0.0;
        for (int bb = 0; bb < nb; ++bb) tt += p.part[(size_t)bb * PW + 0];

        int pciter = 1;
        while (true) {
            // S1: ph_raw[j] = sum_i x[i,j] th[i], one block per column
            for (int64_t j = b; j < n; j += nb) {
                const double *col = p.x + j * ld;
                double s = 0.0;
                for (int64_t i = tid; i < m; i += NIP_THREADS) s += col[i] * p.th[i];
                s = block_sum(s, sh);
                if (tid == 0) p.ph[j] = s;
            }
            grid.sync();
            // S2 (redundant per block): ph = ph_raw / tt; Gram-Schmidt against P[, 0..h); normalise
            double pp_inv;   // 1 / (ph'ph) after normalisation
            {
                // coefficients c_l = sum_j P[j,l] * ph[j] / tt
                const int nh = (p.gramschmidt && h > 0) ? h : 0;
                for (int l = 0; l < nh; ++l) {
                    double s = 0.0;
                    for (int64_t j = tid; j < n; j += NIP_THREADS) s += p.P[(size_t)l * n + j] * (p.ph[j] / tt);
                    s = block_sum(s, sh);
                    if (tid == 0) scoef[l] = s;
                }
                __syncthreads();
                double nrm = 0.0;
                for (int64_t j = tid; j < n; j += NIP_THREADS) {
                    double v = p.ph[j] / tt;
                    double corr = 0.0;
                    for (int l = 0; l < nh; ++l) corr += p.P[(size_t)l * n + j] * scoef[l];
                    v -= corr;
                    nrm += v * v;
                }
                nrm = sqrt(block_sum(nrm, sh));
                // th_new[i] for this block's rows needs the normalised ph: recompute it on the fly per column
                // (ph[j] final value is written by block 0 only, after everybody has finished reading ph_raw)
                double p2 = 0.0;
                for (int64_t j = tid; j < n; j += NIP_THREADS) {
                    double v = p.ph[j] / tt;
                    double corr = 0.0;
                    for (int l = 0; l < nh; ++l) corr += p.P[(size_t)l * n + j] * scoef[l];
                    v = (v - corr) / nrm;
                    p2 += v * v;
                    p.colabs[j] = v;    // every block writes the same value: colabs doubles as the ph scratch
                }
                p2 = block_sum(p2, sh);
                pp_inv = 1.0 / p2;
                __syncthreads();
            }
            grid.sync();   // colabs (= normalised ph) visible everywhere; ph_raw no longer needed
            // S3: th_new = x ph / (ph'ph) for this block's rows; partials of t_l' th_new for Gram-Schmidt
            const int nh = (p.gramschmidt && h > 0) ? h : 0;
            for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int64_t j = 0;
                for (; j + 3 < n; j += 4) {
                    s0 += p.x[j * ld + i] * p.colabs[j];
                    s1 += p.x[(j + 1) * ld + i] * p.colabs[j + 1];
                    s2 += p.x[(j + 2) * ld + i] * p.colabs[j + 2];
                    s3 += p.x[(j + 3) * ld + i] * p.colabs[j + 3];
                }
                for (; j < n; ++j) s0 += p.x[j * ld + i] * p.colabs[j];
                p.th_old[i] = p.th[i];
                p.th[i] = ((s0 + s1) + (s2 + s3)) * pp_inv;
            }
            __syncthreads();
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
                    for (int bb = 0; bb < nb; ++bb) s += p.part[(size_t)bb * PW + 4 + l];
                    scoef[l] = s / p.eig[l];
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
                diff += p.part[(size_t)bb * PW + 1];
                tt += p.part[(size_t)bb * PW + 2];
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
            for (int64_t j = tid; j < n; j += NIP_THREADS) p.P[(size_t)h * n + j] = p.colabs[j];
            if (tid == 0) { p.eig[h] = tt; p.iters[h] = pciter; }
        }
        grid.sync();   // everybody is done reading colabs as ph before it is overwritten below... (P holds ph now)
        for (int64_t j = b; j < n; j += nb) {
            double *col = p.x + j * ld;
            const double pj = p.P[(size_t)h * n + j];
            double s = 0.0;
            for (int64_t i = tid; i < m; i += NIP_THREADS) {
                const double v = col[i] - p.th[i] * pj;
                col[i] = v;
                s += fabs(v);
            }
            s = block_sum(s, sh);
            if (tid == 0) p.colabs[j] = s;
        }
        grid.sync();
    }
    // eig <- sqrt(eig); scores <- sweep(scores, 2, eig, "/")  (unit-norm score columns)
    for (int h = 0; h < p.ncomp; ++h) {
        const double e = sqrt(p.eig[h]);
        for (int64_t i = r0 + tid; i < r1; i += NIP_THREADS) p.T[(size_t)h * m + i] /= e;
        if (b == 0 && tid == 0) p.part[h] = e;     // part[] is free now: reuse for the sqrt(eig) output
    }
}

extern "C" {

int mfb_nipals(mfb_context *ctx, mfb_matrix *x, int ncomp, int center, int scale, int maxiter, double tol,
               int gramschmidt, double *scores, double *loadings, double *eig, int *iters) {
    MFB_ARG(ctx && x && scores && loadings, "NULL argument");
    MFB_ARG(ncomp >= 1 && ncomp <= 60, "ncomp must be in 1..60");
    MFB_ARG(ncomp <= x->m && ncomp <= x->n, "ncomp exceeds a matrix dimension");
    MFB_ARG(maxiter >= 2, "maxiter must be at least 2");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double stats[5];
    int rc = mfb_matrix_stats(x, stats);
    if (rc != MFB_OK) return rc;
    if (stats[3] > 0) {
        mfb_set_error("mfb_nipals: NA values are not supported yet (SURVEY.md 8f)");
        return MFB_ERR_UNSUPPORTED;
    }
    const int64_t m = x->m, n = x->n, ld = x->lda;
    int nb_max = 0;
    MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_max, nipals_kernel, NIP_THREADS, 0));
    int nb = nb_max * ctx->sm_count;
    if (nb > ctx->sm_count) nb = ctx->sm_count;               // one block per SM is plenty; keeps grid.sync cheap
    const int64_t want = (m + NIP_THREADS - 1) / NIP_THREADS > n ? (m + NIP_THREADS - 1) / NIP_THREADS : n;
    if (want < nb) nb = (int)(want < 1 ? 1 : want);
    const size_t PW = ncomp + 4;
    size_t nd = (size_t)ld * n + (size_t)m * ncomp + (size_t)n * ncomp + ncomp + 2 * (size_t)m + 2 * (size_t)n + (size_t)nb * PW;
    double *buf = nullptr;
    int *ibuf = nullptr;
    MFB_CUDA(cudaMalloc(&buf, nd * sizeof(double)));
    MFB_CUDA(cudaMalloc(&ibuf, (ncomp + 1) * sizeof(int)));
    MFB_CUDA(cudaMemsetAsync(buf, 0, nd * sizeof(double), st));
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
    p.iters = ibuf;
    p.status = ibuf + ncomp;
    p.m = m; p.n = n; p.ld = ld;
    p.ncomp = ncomp; p.center = center; p.scale = scale; p.maxiter = maxiter; p.gramschmidt = gramschmidt;
    p.tol = tol;
    MFB_CUDA(cudaMemcpyAsync(p.x, x->d, (size_t)ld * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    void *args[] = {&p};
    MFB_CUDA(cudaLaunchCooperativeKernel((void *)nipals_kernel, dim3(nb), dim3(NIP_THREADS), args, 0, st));
    std::vector<double> hE(ncomp);
    std::vector<int> hI(ncomp + 1);
    MFB_CUDA(cudaMemcpyAsync(scores, p.T, (size_t)m * ncomp * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(loadings, p.P, (size_t)n * ncomp * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(hE.data(), p.part, ncomp * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(hI.data(), ibuf, (ncomp + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    cudaFree(buf);
    cudaFree(ibuf);
    for (int h = 0; h < ncomp; ++h) {
        if (eig) eig[h] = hE[h];
        if (iters) iters[h] = hI[h];
    }
    if (hI[ncomp] == 2) {
        mfb_set_error("mfb_nipals: a column has zero variance and scale = TRUE");
        return MFB_ERR_ARG;
    }
    return hI[ncomp] == 1 ? MFB_NOT_CONVERGED : MFB_OK;
}

}  // extern "C"
