// Dense FP64 building blocks shared by svd_pca and the bcv cells (device pointers, column-major).
#pragma once
#include "common.cuh"

// C[M x N] (ldc) = A' B with A = [K x M] (lda), B = [K x N] (ldb): every entry is a dot product of two
// contiguous columns.  FP64 tensor cores (DMMA.8x8x4), 32 x 32 tile per warp, deterministic split-K.
// `work` must hold at least gemm_tn_workspace(M, N, K) doubles (may be NULL when that is 0).
size_t gemm_tn_workspace(int64_t M, int64_t N, int64_t K, int sm_count);
int gemm_tn(mfb_context *ctx, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
            int64_t M, int64_t N, int64_t K, double *work);

// C[rows x kk] (ldc) = A[rows x inner] (lda) * V[inner x kk] (ldv): 32 rows x 8 column slices per block, 64 columns of
// V per launch.  Optional per-column scale: C[:, j] *= colscale[j].
int tall_times_small(mfb_context *ctx, const double *A, int64_t lda, const double *V, int64_t ldv, double *C,
                     int64_t ldc, int64_t rows, int64_t inner, int kk, const double *colscale);

// Top-`k` eigenpairs (largest first, 1 <= k <= N) of the symmetric N x N matrix G (column-major, ld = ldg, untouched):
// Lanczos tridiagonalisation with full re-orthogonalisation in a persistent cooperative kernel, run until the k
// leading Ritz pairs have converged (or to completion, N steps), Sturm multisection + inverse iteration on the
// tridiagonal matrix, Ritz vectors Q S.  evals: k, evecs: N x k (ld = N), device pointers.
int sym_eig_topk(mfb_context *ctx, const double *G, int64_t ldg, int64_t N, int k, double *evals, double *evecs);

// The same for a batch of independent matrices: ONE cooperative launch advances all of them (blockIdx.y = problem,
// a barrier counter per problem), so that their barrier and latency gaps overlap -- the bcv cells of a rank.
struct EigProblem {
    const double *G;
    int64_t ldg, N;
    int k;
    double *evals, *evecs;
};
int sym_eig_topk_batch(mfb_context *ctx, int np, const EigProblem *probs);

// nipals.cu: NIPALS PCA of a device matrix with the results left on the device (T: m x ncomp unit columns, P: n x ncomp)
int nipals_on_device(mfb_context *ctx, const double *xd, int64_t m, int64_t n, int64_t ld, int has_na, int ncomp,
                     int center, int scale, int maxiter, double tol, int gramschmidt, double *T_out, double *P_out,
                     double *eig_host, int *iters_host);
