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
#include <time.h>

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
    MFB_ARG(cell_begin >= 0 && cell_end <= holdouts * holdouts, "cell range out of bounds");
    return mfb_bcv_cells_multi(ctx, Y, 1, row_fold, col_fold, holdouts, ks, nks, cell_begin, cell_end, err);
}

namespace {

// RAII: run the helpers of this file (they launch on ctx->stream) on another stream of the same context
struct StreamScope {
    mfb_context *ctx;
    cudaStream_t saved;
    StreamScope(mfb_context *c, cudaStream_t s) : ctx(c), saved(c->stream) { c->stream = s; }
    ~StreamScope() { ctx->stream = saved; }
};

int ensure_aux(mfb_context *ctx, size_t nstreams, size_t nevents) {
    while (ctx->aux_streams.size() < nstreams) {
        cudaStream_t s;
        MFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        ctx->aux_streams.push_back(s);
    }
    while (ctx->aux_events.size() < nevents) {
        cudaEvent_t e;
        MFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->aux_events.push_back(e);
    }
    return MFB_OK;
}

// device buffers of one cell of a batch
struct CellSlot {
    double *D, *Dt, *R1, *R2, *mu, *G, *evals, *evecs, *sigma, *colscale, *U, *V, *B, *C, *part, *E, *work1, *work2;
};

}  // namespace

int mfb_bcv_cells_multi(mfb_context *ctx, mfb_matrix *Y, int nfolds, const int32_t *row_folds,
                        const int32_t *col_folds, int holdouts, const int32_t *ks, int nks, int cell_begin,
                        int cell_end, double *err) {
    MFB_ARG(ctx && Y && row_folds && col_folds && ks && err, "NULL argument");
    if (Y->elem_bytes != 8) { mfb_set_error("this entry point needs an FP64 resident matrix"); return MFB_ERR_UNSUPPORTED; }
    MFB_ARG(holdouts >= 2, "holdouts must be at least 2");
    MFB_ARG(nks >= 1 && nfolds >= 1, "ks / folds empty");
    const int h2 = holdouts * holdouts;
    MFB_ARG(cell_begin >= 0 && cell_end <= nfolds * h2 && cell_begin <= cell_end, "cell range out of bounds");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = Y->m, p = Y->n, ldy = Y->lda;
    int kmax = 0;
    for (int q = 0; q < nks; ++q) {
        MFB_ARG(ks[q] >= 1, "ks must be positive");
        if (ks[q] > kmax) kmax = ks[q];
    }
    const int ncell = cell_end - cell_begin;
    if (ncell == 0) return MFB_OK;
    // ---- index lists of every fold draw the range touches (host), in increasing order as R's logical subsetting
    //      keeps them; one upload.  Per draw: [rows in | rows out] per row fold, [cols in | cols out] per col fold ----
    const int f_lo = cell_begin / h2, f_hi = (cell_end - 1) / h2;
    const int nf = f_hi - f_lo + 1;
    std::vector<int32_t> idx((size_t)nf * holdouts * (n + p));
    std::vector<int64_t> n_rin((size_t)nf * holdouts), n_cin((size_t)nf * holdouts);
    int64_t maxr_in = 0, maxr_out = 0, maxc_in = 0, maxc_out = 0;
    for (int f = 0; f < nf; ++f) {
        const int32_t *rf = row_folds + (size_t)(f_lo + f) * n, *cf = col_folds + (size_t)(f_lo + f) * p;
        int32_t *base = idx.data() + (size_t)f * holdouts * (n + p);
        for (int x = 0; x < holdouts; ++x) {
            int32_t *rin = base + (size_t)x * n;
            int64_t a = 0, b = n;                    // hold-in from the front, hold-out from the back (reversed below)
            for (int64_t i = 0; i < n; ++i) {
                const int fo = rf[i];
                MFB_ARG(fo >= 1 && fo <= holdouts, "row_fold entries must be in 1..holdouts");
                if (fo - 1 == x) rin[--b] = (int32_t)i; else rin[a++] = (int32_t)i;
            }
            std::reverse(rin + a, rin + n);
            n_rin[(size_t)f * holdouts + x] = a;
            maxr_in = std::max(maxr_in, a);
            maxr_out = std::max(maxr_out, n - a);
        }
        for (int y = 0; y < holdouts; ++y) {
            int32_t *cin = base + (size_t)holdouts * n + (size_t)y * p;
            int64_t a = 0, b = p;
            for (int64_t j = 0; j < p; ++j) {
                const int fo = cf[j];
                MFB_ARG(fo >= 1 && fo <= holdouts, "col_fold entries must be in 1..holdouts");
                if (fo - 1 == y) cin[--b] = (int32_t)j; else cin[a++] = (int32_t)j;
            }
            std::reverse(cin + a, cin + p);
            n_cin[(size_t)f * holdouts + y] = a;
            maxc_in = std::max(maxc_in, a);
            maxc_out = std::max(maxc_out, p - a);
        }
    }
    DevBuf buf(ctx);
    int32_t *d_idx, *d_ks;
    double *d_err;
    MFB_TRY(buf.alloc(&d_idx, idx.size()));
    MFB_TRY(buf.alloc(&d_ks, (size_t)nks));
    MFB_TRY(buf.alloc(&d_err, (size_t)ncell * nks));
    MFB_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(d_ks, ks, nks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // ---- batch size: the cells of a batch go through the eigen-solver together (one cooperative launch, see
    //      linalg.cu) and their thin products run side by side on auxiliary streams; bounded by workspace ----
    const int64_t ldD = round_up(maxr_in, 4), ldDt = round_up(maxc_in, 4), ldR1 = round_up(maxr_out, 4);
    const int64_t Nmax = std::min(maxr_in, maxc_in), Kmax = std::max(maxr_in, maxc_in);
    const int64_t kkmax = std::min<int64_t>(kmax, Nmax);
    const int64_t ldG = round_up(Nmax, 4);
    const int nb_err = 2 * ctx->sm_count;
    const int pstride = kmax + 1;
    const bool any_wide = maxr_in < maxc_in;       // (row counts differ by at most one between folds: all cells alike)
    size_t w1 = 0, w2 = 0;                          // split-K scratch of the Gram and of the C = U'Y[-r, s] products
    for (int cell = cell_begin; cell < cell_end; ++cell) {
        const int f = cell / h2 - f_lo, local = cell % h2, x = local / holdouts, y = local % holdouts;
        const int64_t nD = n_rin[(size_t)f * holdouts + x], pD = n_cin[(size_t)f * holdouts + y], pA = p - pD;
        const int64_t N = std::min(nD, pD), K = std::max(nD, pD), kk = std::min<int64_t>(kmax, N);
        if (N == 0) continue;
        w1 = std::max(w1, gemm_tn_workspace(N, N, K, ctx->sm_count));
        if (kk > 0 && pA > 0) w2 = std::max(w2, gemm_tn_workspace(kk, pA, nD, ctx->sm_count));
    }
    (void)Kmax; (void)kkmax;
    const size_t slot_doubles = (size_t)ldD * maxc_in + ((any_wide || maxr_in == maxc_in) ? (size_t)ldDt * maxr_in : 0) +
                                (size_t)ldR1 * maxc_in + (size_t)ldD * maxc_out + maxc_in + (size_t)ldG * Nmax +
                                (size_t)kmax * 3 + (size_t)Nmax * kmax + (size_t)ldD * kmax + (size_t)ldDt * kmax +
                                (size_t)ldR1 * kmax + (size_t)kmax * maxc_out + (size_t)nb_err * pstride +
                                (kmax > 64 ? (size_t)maxr_out * maxc_out : 0) + w1 + w2 + 64;
    int B = 16;
    {
        const double budget = 6.0 * 1024 * 1024 * 1024;          // bytes of cell workspace per call
        const double per = (double)slot_doubles * sizeof(double);
        if (per * B > budget) B = (int)(budget / per);
        if (B < 1) B = 1;
        if (const char *e = getenv("MFBICLUST_BCV_BATCH")) { const int b = atoi(e); if (b >= 1) B = b; }
        if (B > ncell) B = ncell;
        if (B > 64) B = 64;
    }
    int S = B < 4 ? B : 4;                                      // auxiliary streams
    if (const char *e = getenv("MFBICLUST_BCV_AUX_STREAMS")) { const int v = atoi(e); if (v >= 1) S = std::min(v, B); }
    // two sets of cell buffers: the gather / Gram stage of the next batch runs while the eigen-solver works on this one
    const int nsets = (ncell > B) ? 2 : 1;
    MFB_TRY(ensure_aux(ctx, (size_t)S, (size_t)4 * S + 2));
    std::vector<CellSlot> slot((size_t)nsets * B);
    for (int i = 0; i < nsets * B; ++i) {
        CellSlot &c = slot[i];
        c.Dt = nullptr; c.E = nullptr; c.work1 = nullptr; c.work2 = nullptr;
        MFB_TRY(buf.alloc(&c.D, (size_t)ldD * maxc_in));
        if (any_wide || maxr_in == maxc_in) MFB_TRY(buf.alloc(&c.Dt, (size_t)ldDt * maxr_in));
        MFB_TRY(buf.alloc(&c.R1, (size_t)ldR1 * maxc_in));            // Y[r, -s]
        MFB_TRY(buf.alloc(&c.R2, (size_t)ldD * maxc_out));            // Y[-r, s]
        MFB_TRY(buf.alloc(&c.mu, (size_t)maxc_in));
        MFB_TRY(buf.alloc(&c.G, (size_t)ldG * Nmax));
        MFB_TRY(buf.alloc(&c.evals, (size_t)kmax));
        MFB_TRY(buf.alloc(&c.evecs, (size_t)Nmax * kmax));
        MFB_TRY(buf.alloc(&c.sigma, (size_t)kmax));
        MFB_TRY(buf.alloc(&c.colscale, (size_t)kmax));
        MFB_TRY(buf.alloc(&c.U, (size_t)ldD * kmax));
        MFB_TRY(buf.alloc(&c.V, (size_t)ldDt * kmax));
        MFB_TRY(buf.alloc(&c.B, (size_t)ldR1 * kmax));
        MFB_TRY(buf.alloc(&c.C, (size_t)kmax * maxc_out));
        MFB_TRY(buf.alloc(&c.part, (size_t)nb_err * pstride));
        if (kmax > 64) MFB_TRY(buf.alloc(&c.E, (size_t)maxr_out * maxc_out));
        if (w1) MFB_TRY(buf.alloc(&c.work1, w1));
        if (w2) MFB_TRY(buf.alloc(&c.work2, w2));
    }
    const double sqrt_eps = 1.4901161193847656e-08;              // sqrt(.Machine$double.eps): MASS::ginv's tol
    cudaEvent_t ev_main = ctx->aux_events[4 * S], ev_eig = ctx->aux_events[4 * S + 1];
    // aux_events[parity * S + s]: stage 1 of a batch finished on stream s; aux_events[(2 + parity) S + s]: stage 3 finished

    struct Geo { int64_t nD, nA, pD, pA; const int32_t *rin, *rout, *cin, *cout; int kk; bool tall, empty; };
    auto geometry = [&](int cell) {
        Geo g;
        const int f = cell / h2 - f_lo, local = cell % h2, x = local / holdouts, y = local % holdouts;
        const int32_t *base = d_idx + (size_t)f * holdouts * (n + p);
        g.nD = n_rin[(size_t)f * holdouts + x]; g.nA = n - g.nD;
        g.pD = n_cin[(size_t)f * holdouts + y]; g.pA = p - g.pD;
        g.rin = base + (size_t)x * n; g.rout = g.rin + g.nD;
        g.cin = base + (size_t)holdouts * n + (size_t)y * p; g.cout = g.cin + g.pD;
        int kk = kmax;                                           // prcomp(D, rank. = max(ks)): k <- min(rank., dim(D))
        if (kk > g.nD) kk = (int)g.nD;
        if (kk > g.pD) kk = (int)g.pD;
        g.kk = kk;
        g.tall = g.nD >= g.pD;
        g.empty = g.nD == 0 || g.pD == 0 || g.nA == 0 || g.pA == 0 || kk == 0;
        return g;
    };

    // stage 1 (per cell, auxiliary streams): centre and gather the blocks, Gram matrix of the short side
    auto stage1 = [&](int b0, int nb, int parity, std::vector<EigProblem> &probs) -> int {
        probs.clear();
        for (int i = 0; i < nb; ++i) {
            const Geo g = geometry(b0 + i);
            if (g.empty) continue;
            CellSlot &c = slot[(size_t)parity * B + i];
            cudaStream_t as = ctx->aux_streams[i % S];
            StreamScope scope(ctx, as);
            // Dc = scale(Y[-r, -s], center = TRUE, scale = FALSE)       (prcomp(center = TRUE), R/bcv.R:184)
            gathered_colmean_kernel<<<(unsigned)(g.pD < 4 * ctx->sm_count ? g.pD : 4 * ctx->sm_count), 256, 0, as>>>(
                Y->d, ldy, g.rin, g.nD, g.cin, g.pD, c.mu);
            MFB_CUDA(cudaGetLastError());
            MFB_CUDA(cudaMemsetAsync(c.D, 0, (size_t)ldD * g.pD * sizeof(double), as));
            MFB_TRY(launch_gather(ctx, Y->d, ldy, g.rin, g.nD, g.cin, g.pD, c.mu, c.D, ldD));
            MFB_TRY(launch_gather(ctx, Y->d, ldy, g.rout, g.nA, g.cin, g.pD, nullptr, c.R1, ldR1));
            MFB_TRY(launch_gather(ctx, Y->d, ldy, g.rin, g.nD, g.cout, g.pA, nullptr, c.R2, ldD));
            const double *Sx = c.D;
            int64_t lds = ldD;
            const int64_t N = g.tall ? g.pD : g.nD, K = g.tall ? g.nD : g.pD;
            if (!g.tall) {
                MFB_TRY(launch_transpose(ctx, c.D, g.nD, g.pD, ldD, c.Dt, ldDt));
                Sx = c.Dt;
                lds = ldDt;
            }
            MFB_TRY(gemm_tn(ctx, Sx, lds, Sx, lds, c.G, ldG, N, N, K, c.work1));
            EigProblem pr;
            pr.G = c.G; pr.ldg = ldG; pr.N = N; pr.k = g.kk; pr.evals = c.evals; pr.evecs = c.evecs;
            probs.push_back(pr);
        }
        for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaEventRecord(ctx->aux_events[parity * S + s2], ctx->aux_streams[s2]));
        return MFB_OK;
    };
    const bool timing = getenv("MFB_BCV_TIMING") != nullptr;     // development aid: host clock after a sync per stage
    auto now_ms = [&]() {
        cudaStreamSynchronize(st);
        for (int s2 = 0; s2 < S; ++s2) cudaStreamSynchronize(ctx->aux_streams[s2]);
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    MFB_CUDA(cudaEventRecord(ev_main, st));                      // index lists uploaded
    for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaStreamWaitEvent(ctx->aux_streams[s2], ev_main, 0));
    std::vector<EigProblem> probs_cur, probs_next;
    MFB_TRY(stage1(cell_begin, std::min(B, cell_end - cell_begin), 0, probs_cur));
    int parity = 0;
    for (int b0 = cell_begin; b0 < cell_end; b0 += B, parity ^= 1) {
        const int nb = std::min(B, cell_end - b0);
        double t_b = 0, t_c = 0, t_d = 0;
        for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaStreamWaitEvent(st, ctx->aux_events[parity * S + s2], 0));
        // stage 3 of the batch before the previous one read the eigenvectors this batch is about to overwrite (the
        // previous batch's stage 3 keeps running on the auxiliary streams, beside this batch's eigen-solver)
        if (b0 - 2 * B >= cell_begin)
            for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaStreamWaitEvent(st, ctx->aux_events[(2 + parity) * S + s2], 0));
        // the next batch's gathers and Gram products go to the auxiliary streams now: they overlap the eigen-solver
        if (b0 + B < cell_end && nsets == 2)
            MFB_TRY(stage1(b0 + B, std::min(B, cell_end - b0 - B), parity ^ 1, probs_next));
        if (timing) t_b = now_ms();
        // ---- stage 2 (the whole batch, main stream): leading eigenpairs of every Gram matrix ----
        if (!probs_cur.empty()) MFB_TRY(sym_eig_topk_batch(ctx, (int)probs_cur.size(), probs_cur.data()));
        MFB_CUDA(cudaEventRecord(ev_eig, st));
        if (timing) t_c = now_ms();
        // ---- stage 3 (per cell, auxiliary streams): singular triplets, B = Y[r,-s] V S^-1, C = U'Y[-r,s], errors ----
        for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaStreamWaitEvent(ctx->aux_streams[s2], ev_eig, 0));
        for (int i = 0; i < nb; ++i) {
            const Geo g = geometry(b0 + i);
            double *e_out = d_err + (size_t)(b0 + i - cell_begin) * nks;
            CellSlot &c = slot[(size_t)parity * B + i];
            cudaStream_t as = ctx->aux_streams[i % S];
            StreamScope scope(ctx, as);
            if (g.empty) {
                MFB_CUDA(cudaMemsetAsync(e_out, 0, nks * sizeof(double), as));
                continue;
            }
            const int kk = g.kk;
            // long side: Z = X v (its column norms are the singular values); short side: the eigenvectors
            if (g.tall) {
                MFB_CUDA(cudaMemcpy2DAsync(c.V, ldDt * sizeof(double), c.evecs, g.pD * sizeof(double),
                                           g.pD * sizeof(double), kk, cudaMemcpyDeviceToDevice, as));
                MFB_TRY(tall_times_small(ctx, c.D, ldD, c.evecs, g.pD, c.U, ldD, g.nD, g.pD, kk, nullptr));
                colnorm_kernel<<<kk, 256, 0, as>>>(c.U, ldD, g.nD, c.sigma);
                sigma_scale_kernel<<<kk, 256, 0, as>>>(c.U, ldD, g.nD, kk, c.sigma, c.colscale, sqrt_eps);
            } else {
                MFB_CUDA(cudaMemcpy2DAsync(c.U, ldD * sizeof(double), c.evecs, g.nD * sizeof(double),
                                           g.nD * sizeof(double), kk, cudaMemcpyDeviceToDevice, as));
                MFB_TRY(tall_times_small(ctx, c.Dt, ldDt, c.evecs, g.nD, c.V, ldDt, g.pD, g.nD, kk, nullptr));
                colnorm_kernel<<<kk, 256, 0, as>>>(c.V, ldDt, g.pD, c.sigma);
                sigma_scale_kernel<<<kk, 256, 0, as>>>(c.V, ldDt, g.pD, kk, c.sigma, c.colscale, sqrt_eps);
            }
            MFB_CUDA(cudaGetLastError());
            // B = Y[r, -s] V S^-1  (|r| x kk);  C = U' Y[-r, s]  (kk x |s|)
            MFB_TRY(tall_times_small(ctx, c.R1, ldR1, c.V, ldDt, c.B, ldR1, g.nA, g.pD, kk, c.colscale));
            MFB_TRY(gemm_tn(ctx, c.U, ldD, c.R2, ldD, c.C, kk, kk, g.pA, g.nD, c.work2));
            if (kk <= 16)
                bcv_error_kernel<16><<<nb_err, 256, 0, as>>>(Y->d, ldy, g.rout, g.nA, g.cout, g.pA, c.B, ldR1, c.C, kk, 0, kk,
                                                             nullptr, 0, c.part, pstride);
            else if (kk <= 32)
                bcv_error_kernel<32><<<nb_err, 256, 0, as>>>(Y->d, ldy, g.rout, g.nA, g.cout, g.pA, c.B, ldR1, c.C, kk, 0, kk,
                                                             nullptr, 0, c.part, pstride);
            else
                for (int j0 = 0; j0 < kk; j0 += 64)
                    bcv_error_kernel<64><<<nb_err, 256, 0, as>>>(Y->d, ldy, g.rout, g.nA, g.cout, g.pA, c.B, ldR1, c.C, kk, j0,
                                                                 kk, c.E, (j0 + 64 < kk) ? 1 : 0, c.part, pstride);
            bcv_error_final_kernel<<<(nks + 63) / 64, 64, 0, as>>>(c.part, nb_err, pstride, kk, d_ks, nks, e_out);
            MFB_CUDA(cudaGetLastError());
        }
        for (int s2 = 0; s2 < S; ++s2) MFB_CUDA(cudaEventRecord(ctx->aux_events[(2 + parity) * S + s2], ctx->aux_streams[s2]));
        if (timing) {
            t_d = now_ms();
            fprintf(stderr, "bcv batch of %d cells: eigen %.3f ms, products+errors (+ next gather/Gram) %.3f ms\n", nb,
                    t_c - t_b, t_d - t_c);
        }
        probs_cur.swap(probs_next);
    }
    for (int s2 = 0; s2 < S; ++s2) {                         // every stage 3 has finished
        MFB_CUDA(cudaEventRecord(ev_main, ctx->aux_streams[s2]));
        MFB_CUDA(cudaStreamWaitEvent(st, ev_main, 0));
    }
    MFB_CUDA(cudaMemcpyAsync(err, d_err, (size_t)ncell * nks * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

// ---------------------------------------------------------------------------
// bcvGivenKs for matrices with missing entries (R/bcv.R:177-181: pca := nipals_pca(., center = TRUE, scale = FALSE))
// ---------------------------------------------------------------------------
// Per cell, as the reference does it: NIPALS (missing-value branch) of the hold-in block D gives scores U (unit
// columns) and loadings V -- WITHOUT singular values (nipals_pca returns np$scores, which nipals normalises) -- so
// estD_k = U_k V_k' and ginv(estD_k) = V_k (V_k'V_k)^-1 (U_k'U_k)^-1 U_k' (U_k, V_k have full column rank; NIPALS
// leaves them orthonormal only up to its tolerance, so the small Gram matrices are kept).  Then
//     estA_k = Y[r,-s] ginv(estD_k) Y[-r,s] = Bt[:, :k] M_k Ct[:k, :],   Bt = Y[r,-s] V,  Ct = U'Y[-r,s],
//     M_k = (V_k'V_k)^-1 (U_k'U_k)^-1,
// and sum((A - estA_k)^2, na.rm = TRUE): a missing entry anywhere in a row of Y[r,-s] or a column of Y[-r,s] makes the
// whole row / column of estA NA (R's %*% propagates NA), which na.rm then drops -- NaN arithmetic does the same here.
}  // extern "C"

namespace {

// M_q = (Gv_k)^-1 (Gu_k)^-1 for k = ks[q] (clamped to kk), one block per q; Gauss-Jordan with partial pivoting in
// shared memory.  Gu, Gv: kk x kk (ld = kk).  M: [nks][kk][kk] (row a, column b at a + b * kk).
__global__ void __launch_bounds__(128) na_minv_kernel(const double *__restrict__ Gu, const double *__restrict__ Gv, int kk,
                                                      const int32_t *__restrict__ ks, double *__restrict__ M) {
    extern __shared__ double sm[];               // two k x 2k augmented systems
    int k = ks[blockIdx.x];
    if (k > kk) k = kk;
    double *Au = sm, *Av = sm + (size_t)k * 2 * k;
    const int tid = threadIdx.x;
    for (int e = tid; e < k * 2 * k; e += 128) {
        const int r = e / (2 * k), c = e - r * 2 * k;
        Au[e] = (c < k) ? Gu[r + (size_t)c * kk] : ((c - k == r) ? 1.0 : 0.0);
        Av[e] = (c < k) ? Gv[r + (size_t)c * kk] : ((c - k == r) ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int which = 0; which < 2; ++which) {
        double *A = which ? Av : Au;
        for (int p = 0; p < k; ++p) {
            __shared__ int piv;
            if (tid == 0) {
                int best = p;
                double bv = fabs(A[p * 2 * k + p]);
                for (int r = p + 1; r < k; ++r)
                    if (fabs(A[r * 2 * k + p]) > bv) { bv = fabs(A[r * 2 * k + p]); best = r; }
                piv = best;
            }
            __syncthreads();
            if (piv != p)
                for (int c = tid; c < 2 * k; c += 128) { const double t = A[p * 2 * k + c]; A[p * 2 * k + c] = A[piv * 2 * k + c]; A[piv * 2 * k + c] = t; }
            __syncthreads();
            const double ip = 1.0 / A[p * 2 * k + p];
            __syncthreads();
            for (int c = tid; c < 2 * k; c += 128) A[p * 2 * k + c] *= ip;
            __syncthreads();
            for (int e = tid; e < k * 2 * k; e += 128) {
                const int r = e / (2 * k), c = e - r * 2 * k;
                if (r != p && c != p) A[e] -= A[r * 2 * k + p] * A[p * 2 * k + c];
            }
            __syncthreads();
            for (int r = tid; r < k; r += 128)
                if (r != p) A[r * 2 * k + p] = 0.0;
            __syncthreads();
        }
    }
    // M = Av^-1 Au^-1 (the right halves)
    double *Mq = M + (size_t)blockIdx.x * kk * kk;
    for (int e = tid; e < kk * kk; e += 128) {
        const int a = e % kk, b = e / kk;
        double v = 0.0;
        if (a < k && b < k)
            for (int c = 0; c < k; ++c) v += Av[a * 2 * k + k + c] * Au[c * 2 * k + k + b];
        Mq[e] = v;
    }
}

// Ck[q][a][c] = sum_b M_q[a][b] Ct[b][c]   (a < kk; rows a >= ks[q] are zero because M_q is)
__global__ void __launch_bounds__(256) na_ck_kernel(const double *__restrict__ M, const double *__restrict__ Ct, int kk,
                                                    int64_t pA, double *__restrict__ Ck) {
    const int q = blockIdx.y;
    const double *Mq = M + (size_t)q * kk * kk;
    double *Cq = Ck + (size_t)q * kk * pA;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < (int64_t)kk * pA; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / kk;
        const int a = (int)(e - c * kk);
        double v = 0.0;
        for (int b = 0; b < kk; ++b) v += Mq[a + (size_t)b * kk] * Ct[b + c * kk];
        Cq[a + c * kk] = v;
    }
}

// part[block][q] = sum over this block's held-out entries of (Y[r_i, s_c] - Bt[i, :] Ck[q][:, c])^2, NaN results dropped
__global__ void __launch_bounds__(256) na_error_kernel(const double *__restrict__ Y, int64_t ldy,
                                                       const int32_t *__restrict__ rows, int64_t nr,
                                                       const int32_t *__restrict__ cols, int64_t nc,
                                                       const double *__restrict__ Bt, int64_t ldb,
                                                       const double *__restrict__ Ck, int kk, int nks,
                                                       double *__restrict__ part) {
    __shared__ double sh[8];
    const int q = blockIdx.y;
    const double *Cq = Ck + (size_t)q * kk * nc;
    double acc = 0.0;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < nr * nc; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / nr, i = e - c * nr;
        double v = Y[rows[i] + (int64_t)cols[c] * ldy];
        for (int a = 0; a < kk; ++a) v -= Bt[i + (int64_t)a * ldb] * Cq[a + c * kk];
        if (v == v) acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        part[(size_t)blockIdx.x * nks + q] = t;
    }
}

__global__ void na_error_final_kernel(const double *__restrict__ part, int nb, int nks, double *__restrict__ err) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nks) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += part[(size_t)b * nks + q];
    err[q] = s;
}

// flags[0] != 0 when a row or a column of the gathered block is entirely NA-or-zero (clean() would drop it and the
// reference's products become non-conformable: R/bicluster.R:271, R/helperFunctions.R:94-101)
__global__ void __launch_bounds__(256) na_allbad_kernel(const double *__restrict__ D, int64_t ld, int64_t nr, int64_t nc,
                                                        int *__restrict__ flags) {
    // columns: one warp each; rows: one thread each
    const int64_t gw = (blockIdx.x * 256ll + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t c = gw; c < nc; c += (int64_t)gridDim.x * 8) {
        int good = 0;
        for (int64_t i = lane; i < nr; i += 32) { const double v = D[i + c * ld]; good |= (v == v && v != 0.0); }
        good = __any_sync(0xffffffffu, good);
        if (!good && lane == 0) atomicExch(&flags[0], 1);
    }
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nr; i += (int64_t)gridDim.x * 256) {
        int good = 0, anynan = 0;
        for (int64_t c = 0; c < nc; ++c) { const double v = D[i + c * ld]; good |= (v == v && v != 0.0); anynan |= (v != v); }
        if (!good) atomicExch(&flags[0], 1);
        if (anynan) atomicExch(&flags[1], 1);               // flags[1]: nipals' has.na = any(is.na(x))
    }
}

}  // namespace
extern "C" {

int mfb_bcv_cells_na(mfb_context *ctx, mfb_matrix *Y, int nfolds, const int32_t *row_folds, const int32_t *col_folds,
                     int holdouts, const int32_t *ks, int nks, int cell_begin, int cell_end, double *err) {
    MFB_ARG(ctx && Y && row_folds && col_folds && ks && err, "NULL argument");
    if (Y->elem_bytes != 8) { mfb_set_error("this entry point needs an FP64 resident matrix"); return MFB_ERR_UNSUPPORTED; }
    MFB_ARG(holdouts >= 2 && nks >= 1 && nfolds >= 1, "bad holdouts / ks / folds");
    const int h2 = holdouts * holdouts;
    MFB_ARG(cell_begin >= 0 && cell_end <= nfolds * h2 && cell_begin <= cell_end, "cell range out of bounds");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = Y->m, p = Y->n, ldy = Y->lda;
    int kmax = 0;
    for (int q = 0; q < nks; ++q) {
        MFB_ARG(ks[q] >= 1, "ks must be positive");
        if (ks[q] > kmax) kmax = ks[q];
    }
    MFB_ARG(kmax <= 60, "bcv with missing entries: at most 60 components (nipals)");
    const int ncell = cell_end - cell_begin;
    if (ncell == 0) return MFB_OK;
    DevBuf buf(ctx);
    std::vector<int32_t> rin, rout, cin, cout;
    int32_t *d_rin, *d_rout, *d_cin, *d_cout, *d_ks;
    int *d_flag;
    MFB_TRY(buf.alloc(&d_rin, (size_t)n)); MFB_TRY(buf.alloc(&d_rout, (size_t)n));
    MFB_TRY(buf.alloc(&d_cin, (size_t)p)); MFB_TRY(buf.alloc(&d_cout, (size_t)p));
    MFB_TRY(buf.alloc(&d_ks, (size_t)nks)); MFB_TRY(buf.alloc(&d_flag, (size_t)4));
    MFB_CUDA(cudaMemcpyAsync(d_ks, ks, nks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    double *d_err;
    MFB_TRY(buf.alloc(&d_err, (size_t)ncell * nks));
    const int nb_err = 2 * ctx->sm_count;
    for (int cell = cell_begin; cell < cell_end; ++cell) {
        const int f = cell / h2, local = cell % h2, x = local / holdouts, y = local % holdouts;
        const int32_t *rf = row_folds + (size_t)f * n, *cf = col_folds + (size_t)f * p;
        rin.clear(); rout.clear(); cin.clear(); cout.clear();
        for (int64_t i = 0; i < n; ++i) {
            MFB_ARG(rf[i] >= 1 && rf[i] <= holdouts, "row_fold entries must be in 1..holdouts");
            (rf[i] - 1 == x ? rout : rin).push_back((int32_t)i);
        }
        for (int64_t j = 0; j < p; ++j) {
            MFB_ARG(cf[j] >= 1 && cf[j] <= holdouts, "col_fold entries must be in 1..holdouts");
            (cf[j] - 1 == y ? cout : cin).push_back((int32_t)j);
        }
        const int64_t nD = (int64_t)rin.size(), nA = (int64_t)rout.size(), pD = (int64_t)cin.size(), pA = (int64_t)cout.size();
        double *e_out = d_err + (size_t)(cell - cell_begin) * nks;
        int kk = kmax;
        if (kk > nD) kk = (int)nD;
        if (kk > pD) kk = (int)pD;
        if (nD == 0 || pD == 0 || nA == 0 || pA == 0 || kk == 0) {
            MFB_CUDA(cudaMemsetAsync(e_out, 0, nks * sizeof(double), st));
            continue;
        }
        MFB_CUDA(cudaMemcpyAsync(d_rin, rin.data(), nD * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_rout, rout.data(), nA * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_cin, cin.data(), pD * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_cout, cout.data(), pA * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaStreamSynchronize(st));               // (the host vectors are reused by the next cell)
        DevBuf cb(ctx);
        const int64_t ldD = round_up(nD, 16), ldR1 = round_up(nA, 4);
        double *D, *R1, *R2, *T, *P, *Bt, *Ct, *Gu, *Gv, *M, *Ck, *part, *work = nullptr;
        MFB_TRY(cb.alloc(&D, (size_t)ldD * pD)); MFB_TRY(cb.alloc(&R1, (size_t)ldR1 * pD)); MFB_TRY(cb.alloc(&R2, (size_t)ldD * pA));
        MFB_TRY(cb.alloc(&T, (size_t)nD * kk)); MFB_TRY(cb.alloc(&P, (size_t)pD * kk));
        MFB_TRY(cb.alloc(&Bt, (size_t)ldR1 * kk)); MFB_TRY(cb.alloc(&Ct, (size_t)kk * pA));
        MFB_TRY(cb.alloc(&Gu, (size_t)kk * kk)); MFB_TRY(cb.alloc(&Gv, (size_t)kk * kk));
        MFB_TRY(cb.alloc(&M, (size_t)nks * kk * kk)); MFB_TRY(cb.alloc(&Ck, (size_t)nks * kk * pA));
        MFB_TRY(cb.alloc(&part, (size_t)nb_err * nks));
        size_t nwork = std::max(gemm_tn_workspace(kk, pA, nD, ctx->sm_count),
                                std::max(gemm_tn_workspace(kk, kk, nD, ctx->sm_count), gemm_tn_workspace(kk, kk, pD, ctx->sm_count)));
        if (nwork) MFB_TRY(cb.alloc(&work, nwork));
        MFB_CUDA(cudaMemsetAsync(D, 0, (size_t)ldD * pD * sizeof(double), st));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rin, nD, d_cin, pD, nullptr, D, ldD));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rout, nA, d_cin, pD, nullptr, R1, ldR1));
        MFB_TRY(launch_gather(ctx, Y->d, ldy, d_rin, nD, d_cout, pA, nullptr, R2, ldD));
        MFB_CUDA(cudaMemsetAsync(d_flag, 0, 4 * sizeof(int), st));
        na_allbad_kernel<<<ctx->sm_count, 256, 0, st>>>(D, ldD, nD, pD, d_flag);
        MFB_CUDA(cudaGetLastError());
        int hflags[2] = {0, 0};
        MFB_CUDA(cudaMemcpyAsync(hflags, d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        if (hflags[0]) {
            mfb_set_error("bcv: a hold-in block has a row or column without any non-zero observed entry (clean() would drop it)");
            return MFB_ERR_ARG;
        }
        // nipals(x = D, ncomp = max(ks), center = TRUE, scale = FALSE, tol = 1e-6)   (R/bcv.R:179, R/bicluster.R:226)
        const int nst = nipals_on_device(ctx, D, nD, pD, ldD, hflags[1], kk, 1, 0, 500, 1e-6, 1, T, P, nullptr, nullptr);
        if (nst < 0) return nst;
        MFB_TRY(tall_times_small(ctx, R1, ldR1, P, pD, Bt, ldR1, nA, pD, kk, nullptr));          // Bt = Y[r,-s] V
        MFB_TRY(gemm_tn(ctx, T, nD, R2, ldD, Ct, kk, kk, pA, nD, work));                          // Ct = U'Y[-r,s]
        MFB_TRY(gemm_tn(ctx, T, nD, T, nD, Gu, kk, kk, kk, nD, work));
        MFB_TRY(gemm_tn(ctx, P, pD, P, pD, Gv, kk, kk, kk, pD, work));
        const size_t msm = (size_t)4 * kk * kk * sizeof(double);
        MFB_CUDA(cudaFuncSetAttribute(na_minv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
        na_minv_kernel<<<nks, 128, msm, st>>>(Gu, Gv, kk, d_ks, M);
        {
            int gx = (int)(((int64_t)kk * pA + 255) / 256);
            if (gx > 4 * ctx->sm_count) gx = 4 * ctx->sm_count;
            na_ck_kernel<<<dim3(gx, nks), 256, 0, st>>>(M, Ct, kk, pA, Ck);
        }
        na_error_kernel<<<dim3(nb_err, nks), 256, 0, st>>>(Y->d, ldy, d_rout, nA, d_cout, pA, Bt, ldR1, Ck, kk, nks, part);
        na_error_final_kernel<<<(nks + 63) / 64, 64, 0, st>>>(part, nb_err, nks, e_out);
        MFB_CUDA(cudaGetLastError());
        MFB_CUDA(cudaStreamSynchronize(st));               // per-cell buffers go back to the pool
    }
    MFB_CUDA(cudaMemcpyAsync(err, d_err, (size_t)ncell * nks * sizeof(double), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

}  // extern "C"
