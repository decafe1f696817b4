// svd_pca (R/bicluster.R:528-539 -> stats::prcomp -> La.svd) and the bcv cells
// (R/bcv.R:195-233 bcvGivenKs; prcomp call site :184, MASS::ginv :222) on the GPU.
//
// Both need the leading `kk` singular triplets of a dense FP64 matrix X.  The reference calls LAPACK
// dgesdd on the full matrix (and, inside bcvGivenKs, once more per k through MASS::ginv); here
//   * the Gram matrix of the SHORT side of X is formed by the FP64 tensor-core product (linalg.cu gemm_tn),
//   * its top-kk eigenpairs come from the symmetric solver (Lanczos tridiagonalisation with full
//     re-orthogonalisation run to convergence of the kk leading Ritz pairs, Sturm multisection, inverse iteration;
//     linalg.cu sym_eig_topk),
//   * the other side's vectors are one thin product X v, and the singular values are the NORMS of those products,
//     sigma_j = ||X v_j||, not sqrt(lambda_j): a null direction of X has lambda_j ~ eps lambda_1, whose square root
//     sits right at MASS::ginv's sqrt(eps) cut, while ||X v_j|| ~ eps sigma_1 is dropped as LAPACK's value would be.
// A bcv cell then needs ONE factorisation: with Dc = U S V' the centred hold-in block,
//   ginv(Dc V_k V_k') = V_k' S_k'^-1 U_k''  over the components j <= k with s_j > sqrt(eps) s_1   (MASS::ginv rule)
// so estA_k = sum_{j<=k, kept} B[:,j] C[j,:] with B = Y[r,-s] V S^-1, C = U' Y[-r,s], and the errors of every k
// are the running squared norms of rank-1 down-dates of Y[r,s], accumulated in one pass.
#include <algorithm>

#include "linalg.cuh"

#define MFB_TRY(expr) do { int _r = (expr); if (_r != MFB_OK) return _r; } while (0)

namespace {

// dst (cols x rows, ld = ldd) = src' (src rows x cols, ld = lds): 32 x 32 shared-memory tiles
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ src, int64_t rows, int64_t cols,
                                                        int64_t lds, double *__restrict__ dst, int64_t ldd) {
    __shared__ double tile[32][33];
    const int64_t r0 = blockIdx.x * 32ll, c0 = blockIdx.y * 32ll;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {
        const int64_t r = r0 + tx, c = c0 + q;
        tile[q][tx] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const int64_t c = c0 + tx, r = r0 + q;
        if (r < rows && c < cols) dst[c + r * ldd] = tile[tx][q];
    }
}

// mu[j] = mean_i Y[rows[i], cols[j]]: one block per column, fixed-order two-stage sum
__global__ void __launch_bounds__(256) gathered_colmean_kernel(const double *__restrict__ Y, int64_t ldy,
                                                               const int32_t *__restrict__ rows, int64_t nr,
                                                               const int32_t *__restrict__ cols, int64_t nc,
                                                               double *__restrict__ mu) {
    __shared__ double sh[8];
    for (int64_t j = blockIdx.x; j < nc; j += gridDim.x) {
        const double *col = Y + (int64_t)cols[j] * ldy;
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < nr; i += 256) s += col[rows[i]];
        s = warp_sum(s);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[w];
            mu[j] = t / (double)nr;
        }
    }
}

// dst[i + j ldd] = Y[rows[i], cols[j]] - (mu ? mu[j] : 0)
__global__ void __launch_bounds__(256) gather_kernel(const double *__restrict__ Y, int64_t ldy,
                                                     const int32_t *__restrict__ rows, int64_t nr,
                                                     const int32_t *__restrict__ cols, int64_t nc,
                                                     const double *__restrict__ mu, double *__restrict__ dst,
                                                     int64_t ldd) {
    for (int64_t j = blockIdx.y; j < nc; j += gridDim.y) {
        const double *col = Y + (int64_t)cols[j] * ldy;
        const double mj = mu ? mu[j] : 0.0;
        for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nr; i += (int64_t)gridDim.x * 256)
            dst[i + j * ldd] = col[rows[i]] - mj;
    }
}

// sigma[j] = ||Z[:, j]||  (Z = X v: the singular value computed directly), one block per column, fixed-order sum
__global__ void __launch_bounds__(256) colnorm_kernel(const double *__restrict__ Z, int64_t ldz, int64_t rows,
                                                      double *__restrict__ sigma) {
    __shared__ double sh[8];
    const double *z = Z + (size_t)blockIdx.x * ldz;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) s += z[i] * z[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        sigma[blockIdx.x] = sqrt(t);
    }
}

// colscale[j] = 1 / sigma_j for the components MASS::ginv keeps (sigma_j > max(rel_tol * sigma_1, 0), sigma_1 the
// largest), else 0; with rel_tol < 0 (svd_pca: nothing is ever dropped) every non-zero sigma is kept.
// Then Z[:, j] *= colscale[j]  (unit columns for kept components, zero columns for dropped ones).
__global__ void __launch_bounds__(256) sigma_scale_kernel(double *__restrict__ Z, int64_t ldz, int64_t rows, int kk,
                                                          const double *__restrict__ sigma,
                                                          double *__restrict__ colscale, double rel_tol) {
    const int j = blockIdx.x;
    double s1 = 0.0;
    for (int q = 0; q < kk; ++q) s1 = fmax(s1, sigma[q]);
    const double sj = sigma[j];
    const double cs = (sj > fmax(rel_tol * s1, 0.0)) ? 1.0 / sj : 0.0;
    if (threadIdx.x == 0) colscale[j] = cs;
    double *z = Z + (size_t)j * ldz;
    for (int64_t i = threadIdx.x; i < rows; i += 256) z[i] *= cs;
}

// Running errors of the rank-1 down-dates: for every held-out entry e0 = Y[r_i, s_c],
//   e_j = e_{j-1} - B[i, j] C[j, c],   part[block][j] = sum over the block's entries of e_j^2   (j = 0 .. kk-1)
// and part[block][kk] = sum e0^2.  NaN entries of Y are skipped (sum(..., na.rm = TRUE), R/bcv.R:225-227).
// One launch handles the components j0 .. j0 + KK - 1; with more than KK components the running residual is kept in
// E (|r| x |s|) between the launches.
template <int KK>
__global__ void __launch_bounds__(256) bcv_error_kernel(const double *__restrict__ Y, int64_t ldy,
                                                        const int32_t *__restrict__ rows, int64_t nr,
                                                        const int32_t *__restrict__ cols, int64_t nc,
                                                        const double *__restrict__ B, int64_t ldb,
                                                        const double *__restrict__ C, int64_t ldc, int j0, int kk,
                                                        double *__restrict__ E, int write_e,
                                                        double *__restrict__ part, int pstride) {
    __shared__ double sh[8][KK + 1];
    double acc[KK + 1];
#pragma unroll
    for (int j = 0; j <= KK; ++j) acc[j] = 0.0;
    const int64_t total = nr * nc;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / nr, i = e - c * nr;
        double v = (j0 == 0) ? Y[rows[i] + (int64_t)cols[c] * ldy] : E[e];
        if (v != v) {
            if (write_e) E[e] = v;
            continue;
        }
        acc[KK] += v * v;
#pragma unroll
        for (int j = 0; j < KK; ++j) {
            if (j0 + j < kk) {
                v -= B[i + (int64_t)(j0 + j) * ldb] * C[(j0 + j) + c * ldc];
                acc[j] += v * v;
            }
        }
        if (write_e) E[e] = v;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j <= KK; ++j) {
        const double s = warp_sum(acc[j]);
        if (l == 0) sh[w][j] = s;
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= KK; j += 256) {
        if (j == KK && j0 != 0) continue;          // ||A||^2 comes from the first chunk
        if (j < KK && j0 + j >= kk) continue;
        double s = 0.0;
        for (int ww = 0; ww < 8; ++ww) s += sh[ww][j];
        part[(size_t)blockIdx.x * pstride + (j == KK ? pstride - 1 : j0 + j)] = s;
    }
}

// err[q] = sum_b part[b][ks[q] - 1]  (ks[q] clamped to kk; ks[q] = 0 -> ||A||^2 in the last column)
__global__ void bcv_error_final_kernel(const double *__restrict__ part, int nb, int pstride, int kk,
                                       const int32_t *__restrict__ ks, int nks, double *__restrict__ err) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nks) return;
    int k = ks[q];
    if (k > kk) k = kk;
    const int col = (k <= 0) ? pstride - 1 : k - 1;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += part[(size_t)b * pstride + col];
    err[q] = s;
}

int launch_gather(mfb_context *ctx, const double *Y, int64_t ldy, const int32_t *rows, int64_t nr,
                  const int32_t *cols, int64_t nc, const double *mu, double *dst, int64_t ldd) {
    if (nr == 0 || nc == 0) return MFB_OK;
    int gx = (int)((nr + 255) / 256);
    if (gx > 256) gx = 256;
    const int gy = (int)(nc < 65535 ? nc : 65535);
    gather_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(Y, ldy, rows, nr, cols, nc, mu, dst, ldd);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int launch_transpose(mfb_context *ctx, const double *src, int64_t rows, int64_t cols, int64_t lds, double *dst,
                     int64_t ldd) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    transpose_kernel<<<grid, 256, 0, ctx->stream>>>(src, rows, cols, lds, dst, ldd);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

// Leading kk singular triplets of the device matrix X (rows x cols, ld = ldx; untouched):
//   sigma[kk] (= ||X v_j|| resp. ||X'u_j||, computed directly), colscale[kk] (see sigma_scale_kernel), U (rows x kk,
//   ld = ldu, unit columns for kept components, zero columns for dropped ones) and V (cols x kk, ld = ldv) -- either
//   pointer may be NULL when that side is not wanted.  `Xt` must hold a cols x rows (ld = ldxt) scratch when
//   rows < cols (wide case).
int topk_svd(mfb_context *ctx, const double *X, int64_t rows, int64_t cols, int64_t ldx, double *Xt, int64_t ldxt,
             int kk, double rel_tol, double *sigma, double *colscale, double *U, int64_t ldu, double *V,
             int64_t ldv) {
    DevBuf buf(ctx);
    const bool tall = rows >= cols;
    const int64_t N = tall ? cols : rows, K = tall ? rows : cols;
    const int64_t ldg = round_up(N, 4);
    double *G = nullptr, *evals = nullptr, *evecs = nullptr, *work = nullptr, *Zown = nullptr;
    MFB_TRY(buf.alloc(&G, (size_t)ldg * N));
    MFB_TRY(buf.alloc(&evals, (size_t)kk));
    MFB_TRY(buf.alloc(&evecs, (size_t)N * kk));
    const size_t nwork = gemm_tn_workspace(N, N, K, ctx->sm_count);
    if (nwork) MFB_TRY(buf.alloc(&work, nwork));
    const double *S = X;
    int64_t lds = ldx;
    if (!tall) {
        MFB_TRY(launch_transpose(ctx, X, rows, cols, ldx, Xt, ldxt));
        S = Xt;
        lds = ldxt;
    }
    MFB_TRY(gemm_tn(ctx, S, lds, S, lds, G, ldg, N, N, K, work));     // Gram of the short side (symmetric: half the tiles)
    MFB_TRY(sym_eig_topk(ctx, G, ldg, N, kk, evals, evecs));
    double *Other = tall ? V : U;                      // the short side's vectors are the eigenvectors themselves
    const int64_t ldo = tall ? ldv : ldu;
    if (Other) MFB_CUDA(cudaMemcpy2DAsync(Other, ldo * sizeof(double), evecs, N * sizeof(double), N * sizeof(double),
                                          kk, cudaMemcpyDeviceToDevice, ctx->stream));
    // the long side's vectors Z = S' v (K x kk) -- their norms are the singular values
    double *Z = tall ? U : V;
    if (!Z && rel_tol < 0.0) {                          // svd_pca of a wide matrix: nothing else is wanted
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        return MFB_OK;
    }
    int64_t ldz = tall ? ldu : ldv;
    if (!Z) {
        ldz = round_up(K, 4);
        MFB_TRY(buf.alloc(&Zown, (size_t)ldz * kk));
        Z = Zown;
    }
    MFB_TRY(tall_times_small(ctx, tall ? X : Xt, tall ? ldx : ldxt, evecs, N, Z, ldz, K, N, kk, nullptr));
    colnorm_kernel<<<kk, 256, 0, ctx->stream>>>(Z, ldz, K, sigma);
    sigma_scale_kernel<<<kk, 256, 0, ctx->stream>>>(Z, ldz, K, kk, sigma, colscale, rel_tol);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

}  // namespace

extern "C" {

int mfb_svd_pca(mfb_context *ctx, mfb_matrix *A, int k, double *W, double *H, int *kout) {
    MFB_ARG(ctx && A && W && H, "NULL argument");
    if (A->elem_bytes != 8) { mfb_set_error("this entry point needs an FP64 resident matrix"); return MFB_ERR_UNSUPPORTED; }
    MFB_ARG(k >= 1, "k must be positive");
    MFB_CUDA(cudaSetDevice(ctx->device));
    double stats[5];
    MFB_TRY(mfb_matrix_stats(A, stats));
    if (stats[3] > 0) {
        mfb_set_error("svd_pca: infinite or missing values in 'x'");       // what La.svd reports for NA input
        return MFB_ERR_ARG;
    }
    const int64_t m = A->m, n = A->n;
    int kk = k;                                   // prcomp: k <- min(rank., n, p)
    if (kk > m) kk = (int)m;
    if (kk > n) kk = (int)n;
    if (kout) *kout = kk;
    DevBuf buf(ctx);
    double *sigma, *colscale, *dW, *dH, *At = nullptr, *work = nullptr;
    const int64_t ldw = round_up(m, 4), ldat = round_up(n, 4);
    MFB_TRY(buf.alloc(&sigma, (size_t)kk));
    MFB_TRY(buf.alloc(&colscale, (size_t)kk));
    MFB_TRY(buf.alloc(&dW, (size_t)ldw * kk));
    MFB_TRY(buf.alloc(&dH, (size_t)kk * n));
    if (m < n) MFB_TRY(buf.alloc(&At, (size_t)ldat * m));
    MFB_CUDA(cudaMemsetAsync(dW, 0, (size_t)ldw * kk * sizeof(double), ctx->stream));
    // W = left singular vectors of A (= prcomp(t(A))$rotation); never dropped: rel_tol = -1 keeps all
    MFB_TRY(topk_svd(ctx, A->d, m, n, A->lda, At, ldat, kk, -1.0, sigma, colscale, dW, ldw, nullptr, 0));
    // H = t(t(A) %*% W) = W'A
    const size_t nwork = gemm_tn_workspace(kk, n, m, ctx->sm_count);
    if (nwork) MFB_TRY(buf.alloc(&work, nwork));
    MFB_TRY(gemm_tn(ctx, dW, ldw, A->d, A->lda, dH, kk, kk, n, m, work));
    MFB_CUDA(cudaMemcpy2DAsync(W, m * sizeof(double), dW, ldw * sizeof(double), m * sizeof(double), kk,
                               cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(H, dH, (size_t)kk * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

int mfb_bcv_cells(mfb_context *ctx, mfb_matrix *Y, const int32_t *row_fold, const int32_t *col_fold, int holdouts,
                  const int32_t *ks, int nks, int cell_begin, int cell_end, double *err) {
    MFB_ARG(ctx && Y && row_fold && col_fold && ks && err, "NULL argument");
    if (Y->elem_bytes != 8) { mfb_set_error("this entry point needs an FP64 resident matrix"); return MFB_ERR_UNSUPPORTED; }
    MFB_ARG(holdouts >= 2, "holdouts must be at least 2");
    MFB_ARG(nks >= 1, "ks is empty");
    MFB_ARG(cell_begin >= 0 && cell_end <= holdouts * holdouts && cell_begin <= cell_end, "cell range out of bounds");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = Y->m, p = Y->n, ldy = Y->lda;
    int kmax = 0;
    for (int q = 0; q < nks; ++q) {
        MFB_ARG(ks[q] >= 1, "ks must be positive");
        if (ks[q] > kmax) kmax = ks[q];
    }
    // index lists per fold (host), in increasing order as R's logical subsetting keeps them
    std::vector<std::vector<int32_t>> rin(holdouts), rout(holdouts), cin(holdouts), cout(holdouts);
    for (int64_t i = 0; i < n; ++i) {
        const int f = row_fold[i];
        MFB_ARG(f >= 1 && f <= holdouts, "row_fold entries must be in 1..holdouts");
        for (int x = 0; x < holdouts; ++x) (x == f - 1 ? rout[x] : rin[x]).push_back((int32_t)i);
    }
    for (int64_t j = 0; j < p; ++j) {
        const int f = col_fold[j];
        MFB_ARG(f >= 1 && f <= holdouts, "col_fold entries must be in 1..holdouts");
        for (int y = 0; y < holdouts; ++y) (y == f - 1 ? cout[y] : cin[y]).push_back((int32_t)j);
    }
    int64_t maxr_in = 0, maxr_out = 0, maxc_in = 0, maxc_out = 0;
    for (int x = 0; x < holdouts; ++x) {
        maxr_in = std::max<int64_t>(maxr_in, rin[x].size());
        maxr_out = std::max<int64_t>(maxr_out, rout[x].size());
        maxc_in = std::max<int64_t>(maxc_in, cin[x].size());
        maxc_out = std::max<int64_t>(maxc_out, cout[x].size());
    }
    const int ncell = cell_end - cell_begin;
    if (ncell == 0) return MFB_OK;
    DevBuf buf(ctx);
    int32_t *d_rin, *d_rout, *d_cin, *d_cout, *d_ks;
    double *D, *Dt = nullptr, *R1, *R2, *mu, *sigma, *colscale, *U, *V, *B, *C, *part, *d_err, *work = nullptr;
    double *E = nullptr;                                         // running residual of Y[r, s] when kmax > 64
    const int64_t ldD = round_up(maxr_in, 4), ldDt = round_up(maxc_in, 4), ldR1 = round_up(maxr_out, 4);
    MFB_TRY(buf.alloc(&d_rin, (size_t)maxr_in));
    MFB_TRY(buf.alloc(&d_rout, (size_t)maxr_out));
    MFB_TRY(buf.alloc(&d_cin, (size_t)maxc_in));
    MFB_TRY(buf.alloc(&d_cout, (size_t)maxc_out));
    MFB_TRY(buf.alloc(&d_ks, (size_t)nks));
    MFB_TRY(buf.alloc(&D, (size_t)ldD * maxc_in));
    MFB_TRY(buf.alloc(&Dt, (size_t)ldDt * maxr_in));
    MFB_TRY(buf.alloc(&R1, (size_t)ldR1 * maxc_in));            // Y[r, -s]
    MFB_TRY(buf.alloc(&R2, (size_t)ldD * maxc_out));            // Y[-r, s]
    MFB_TRY(buf.alloc(&mu, (size_t)maxc_in));
    MFB_TRY(buf.alloc(&sigma, (size_t)kmax));
    MFB_TRY(buf.alloc(&colscale, (size_t)kmax));
    MFB_TRY(buf.alloc(&U, (size_t)ldD * kmax));
    MFB_TRY(buf.alloc(&V, (size_t)ldDt * kmax));
    MFB_TRY(buf.alloc(&B, (size_t)ldR1 * kmax));
    MFB_TRY(buf.alloc(&C, (size_t)kmax * maxc_out));
    const int nb_err = 2 * ctx->sm_count;
    const int pstride = kmax + 1;
    MFB_TRY(buf.alloc(&part, (size_t)nb_err * pstride));
    if (kmax > 64) MFB_TRY(buf.alloc(&E, (size_t)maxr_out * maxc_out));
    MFB_TRY(buf.alloc(&d_err, (size_t)ncell * nks));
    {
        size_t nwork = 0;                                        // split-K scratch of the largest C = U'Y[-r, s] product
        for (int cell = cell_begin; cell < cell_end; ++cell) {
            const int x = cell / holdouts, y = cell % holdouts;
            const int64_t nD = (int64_t)rin[x].size(), pD = (int64_t)cin[y].size(), pA = (int64_t)cout[y].size();
            const int64_t kk = std::min<int64_t>(kmax, std::min(nD, pD));
            if (kk > 0 && pA > 0 && nD > 0) nwork = std::max(nwork, gemm_tn_workspace(kk, pA, nD, ctx->sm_count));
        }
        if (nwork) MFB_TRY(buf.alloc(&work, nwork));
    }
    MFB_CUDA(cudaMemcpyAsync(d_ks, ks, nks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const double sqrt_eps = 1.4901161193847656e-08;              // sqrt(.Machine$double.eps): MASS::ginv's tol

    for (int cell = cell_begin; cell < cell_end; ++cell) {
        const int x = cell / holdouts, y = cell % holdouts;
        const int64_t nD = (int64_t)rin[x].size(), nA = (int64_t)rout[x].size();
        const int64_t pD = (int64_t)cin[y].size(), pA = (int64_t)cout[y].size();
        double *e_out = d_err + (size_t)(cell - cell_begin) * nks;
        int kk = kmax;                                           // prcomp(D, rank. = max(ks)): k <- min(rank., dim(D))
        if (kk > nD) kk = (int)nD;
        if (kk > pD) kk = (int)pD;
        if (nD == 0 || pD == 0 || nA == 0 || pA == 0 || kk == 0) {
            MFB_CUDA(cudaMemsetAsync(e_out, 0, nks * sizeof(double), st));
            continue;
        }
        MFB_CUDA(cudaMemcpyAsync(d_rin, rin[x].data(), nD * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_rout, rout[x].data(), nA * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_cin, cin[y].data(), pD * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_cout, cout[y].data(), pA * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        // Dc = scale(Y[-r, -s], center = TRUE, scale = FALSE)       (prcomp(center = TRUE), R/bcv.R:184)
        gathered_colmean_kernel<<<(unsigned)(pD < 4 * ctx->sm_count ? pD : 4 * ctx->sm_count), 256, 0, st>>>(
            Y->d, ldy, d_rin, nD, d_cin, pD, mu);
        MFB_CUDA(cudaGetLastError());
        MFB_CUDA(cudaMemsetAsync(D, 0, (size_t)ldD * pD * sizeof(double), st));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rin, nD, d_cin, pD, mu, D, ldD));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rout, nA, d_cin, pD, nullptr, R1, ldR1));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rin, nD, d_cout, pA, nullptr, R2, ldD));
        MFB_TRY(topk_svd(ctx, D, nD, pD, ldD, Dt, ldDt, kk, sqrt_eps, sigma, colscale, U, ldD, V, ldDt));
        // B = Y[r, -s] V S^-1  (|r| x kk);  C = U' Y[-r, s]  (kk x |s|)
        MFB_TRY(tall_times_small(ctx, R1, ldR1, V, ldDt, B, ldR1, nA, pD, kk, colscale));
        MFB_TRY(gemm_tn(ctx, U, ldD, R2, ldD, C, kk, kk, pA, nD, work));
        if (kk <= 16) bcv_error_kernel<16><<<nb_err, 256, 0, st>>>(Y->d, ldy, d_rout, nA, d_cout, pA, B, ldR1, C, kk, 0, kk, nullptr, 0, part, pstride);
        else if (kk <= 32) bcv_error_kernel<32><<<nb_err, 256, 0, st>>>(Y->d, ldy, d_rout, nA, d_cout, pA, B, ldR1, C, kk, 0, kk, nullptr, 0, part, pstride);
        else
            for (int j0 = 0; j0 < kk; j0 += 64)
                bcv_error_kernel<64><<<nb_err, 256, 0, st>>>(Y->d, ldy, d_rout, nA, d_cout, pA, B, ldR1, C, kk, j0, kk, E,
                                                             (j0 + 64 < kk) ? 1 : 0, part, pstride);
        bcv_error_final_kernel<<<(nks + 63) / 64, 64, 0, st>>>(part, nb_err, pstride, kk, d_ks, nks, e_out);
        MFB_CUDA(cudaGetLastError());
    }
    MFB_CUDA(cudaMemcpyAsync(err, d_err, (size_t)ncell * nks * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

}  // extern "C"
