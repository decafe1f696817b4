// Dense FP64 building blocks for svd_pca and the bcv cells (see linalg.cuh).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "linalg.cuh"

namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// C = A' B  (dot products of contiguous columns), FP64 tensor cores
// ---------------------------------------------------------------------------
// One warp owns a 32 x 32 tile of C: 4 x 4 m8n8k4 blocks, accumulators in registers; per 16-deep k step a lane
// reads 4 consecutive k of one column per block (A-operand [M = column of A][K], B-operand [K][N = column of B]).
__global__ void __launch_bounds__(128) gemm_tn_kernel(const double *__restrict__ A, int64_t lda,
                                                      const double *__restrict__ B, int64_t ldb,
                                                      double *__restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                      int64_t K, int64_t kz, int vec_ok) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t m0 = ((int64_t)blockIdx.x * 2 + (warp & 1)) * 32;
    const int64_t n0 = ((int64_t)blockIdx.y * 2 + (warp >> 1)) * 32;
    if (m0 >= M || n0 >= N) return;
    const int64_t kbeg = (int64_t)blockIdx.z * kz, kend = (kbeg + kz < K) ? kbeg + kz : K;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double *ap[4], *bp[4];
    bool av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t mc = m0 + 8 * i + g, nc = n0 + 8 * i + g;
        av[i] = mc < M;
        bv[i] = nc < N;
        ap[i] = A + (av[i] ? mc : 0) * lda;
        bp[i] = B + (bv[i] ? nc : 0) * ldb;
    }
    for (int64_t k0 = kbeg; k0 < kend; k0 += 16) {
        double a[4][4], b[4][4];
        const int64_t kk = k0 + 4 * t;
        if (vec_ok && k0 + 16 <= kend) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 a0 = *(const double2 *)(ap[i] + kk), a1 = *(const double2 *)(ap[i] + kk + 2);
                const double2 b0 = *(const double2 *)(bp[i] + kk), b1 = *(const double2 *)(bp[i] + kk + 2);
                a[i][0] = av[i] ? a0.x : 0.0; a[i][1] = av[i] ? a0.y : 0.0; a[i][2] = av[i] ? a1.x : 0.0; a[i][3] = av[i] ? a1.y : 0.0;
                b[i][0] = bv[i] ? b0.x : 0.0; b[i][1] = bv[i] ? b0.y : 0.0; b[i][2] = bv[i] ? b1.x : 0.0; b[i][3] = bv[i] ? b1.y : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const bool in = (kk + s) < kend;
                    a[i][s] = (in && av[i]) ? ap[i][kk + s] : 0.0;
                    b[i][s] = (in && bv[i]) ? bp[i][kk + s] : 0.0;
                }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i][s], b[j][s]);
    }
    double *Cz = C + (size_t)blockIdx.z * (size_t)ldc * (size_t)N;   // split-K partial z (ldc * N doubles apart)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t mr = m0 + 8 * i + g;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t nc = n0 + 8 * j + 2 * t + e;
                if (mr < M && nc < N) Cz[mr + nc * ldc] = acc[i][j][e];
            }
        }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double *__restrict__ part, int Z, int64_t M,
                                                            int64_t N, int64_t ldp, double *__restrict__ C,
                                                            int64_t ldc) {
    const int64_t total = M * N;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t n = e / M, m = e - n * M;
        double s = 0.0;
        for (int z = 0; z < Z; ++z) s += part[(size_t)z * ldp * N + m + n * ldp];
        C[m + n * ldc] = s;
    }
}

static int splitk_factor(int64_t M, int64_t N, int64_t K, int sm_count) {
    const int64_t tiles = ((M + 63) / 64) * ((N + 63) / 64);
    int Z = 1;
    if (tiles < 2 * sm_count) {
        Z = (int)((2 * sm_count + tiles - 1) / tiles);
        const int64_t maxz = (K + 255) / 256;
        if (Z > maxz) Z = (int)maxz;
        if (Z < 1) Z = 1;
        if (Z > 64) Z = 64;
    }
    return Z;
}

size_t gemm_tn_workspace(int64_t M, int64_t N, int64_t K, int sm_count) {
    const int Z = splitk_factor(M, N, K, sm_count);
    return Z > 1 ? (size_t)Z * M * N : 0;
}

int gemm_tn(mfb_context *ctx, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
            int64_t M, int64_t N, int64_t K, double *work) {
    const int Z = splitk_factor(M, N, K, ctx->sm_count);
    int64_t kz = (K + Z - 1) / Z;
    kz = (kz + 15) / 16 * 16;
    const int vec_ok = (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)A) % 32 == 0) && (((uintptr_t)B) % 32 == 0);
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)Z);
    if (Z == 1) {
        gemm_tn_kernel<<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, M, N, K, kz, vec_ok);
    } else {
        MFB_ARG(work != nullptr, "gemm_tn: split-K needs a workspace");
        gemm_tn_kernel<<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, work, M, M, N, K, kz, vec_ok);
        int nb = (int)((M * N + 255) / 256);
        if (nb > 4 * ctx->sm_count) nb = 4 * ctx->sm_count;
        splitk_reduce_kernel<<<nb, 256, 0, ctx->stream>>>(work, Z, M, N, M, C, ldc);
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

// ---------------------------------------------------------------------------
// C = A V  for a tall A and a thin V (kk <= 64)
// ---------------------------------------------------------------------------
// Block = 32 rows x 8 column slices: warp w walks the columns c = w, w + 8, ... of A (a warp reads 256 contiguous
// bytes per column, four columns of loads in flight), V is staged through shared memory 32 columns at a time, and
// the eight slices are summed in shared memory in a fixed order.
template <int KK>
__global__ void __launch_bounds__(256) tall_times_small_kernel(const double *__restrict__ A, int64_t lda,
                                                               const double *__restrict__ V, int64_t ldv,
                                                               double *__restrict__ C, int64_t ldc, int64_t rows,
                                                               int64_t inner, int kk,
                                                               const double *__restrict__ colscale) {
    __shared__ double sv[32 * KK];
    extern __shared__ double red[];            // [8][32][KK + 1]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = blockIdx.x * 32ll + lane;
    const bool rowok = i < rows;
    double acc[KK];
#pragma unroll
    for (int j = 0; j < KK; ++j) acc[j] = 0.0;
    for (int64_t c0 = 0; c0 < inner; c0 += 32) {
        const int cn = (int)((inner - c0 < 32) ? inner - c0 : 32);
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * KK; e += 256) {
            const int c = e / KK, j = e - c * KK;
            sv[e] = (c < cn && j < kk) ? V[(c0 + c) + (int64_t)j * ldv] : 0.0;
        }
        __syncthreads();
        // this warp's four columns of the chunk: c = w, w + 8, w + 16, w + 24
        double a4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = w + 8 * q;
            a4[q] = (rowok && c < cn) ? A[i + (c0 + c) * lda] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = w + 8 * q;
#pragma unroll
            for (int j = 0; j < KK; ++j) acc[j] += a4[q] * sv[c * KK + j];
        }
    }
#pragma unroll
    for (int j = 0; j < KK; ++j) red[(w * 32 + lane) * (KK + 1) + j] = acc[j];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * KK; e += 256) {
        const int r = e % 32, j = e / 32;
        const int64_t ir = blockIdx.x * 32ll + r;
        if (ir < rows && j < kk) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += red[(ww * 32 + r) * (KK + 1) + j];
            C[ir + (int64_t)j * ldc] = colscale ? s * colscale[j] : s;
        }
    }
}

template <int KK>
static int launch_tall_times_small(mfb_context *ctx, const double *A, int64_t lda, const double *V, int64_t ldv,
                                   double *C, int64_t ldc, int64_t rows, int64_t inner, int kk,
                                   const double *colscale) {
    const size_t smem = (size_t)8 * 32 * (KK + 1) * sizeof(double);
    MFB_CUDA(cudaFuncSetAttribute(tall_times_small_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tall_times_small_kernel<KK><<<(unsigned)((rows + 31) / 32), 256, smem, ctx->stream>>>(A, lda, V, ldv, C, ldc, rows,
                                                                                         inner, kk, colscale);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int tall_times_small(mfb_context *ctx, const double *A, int64_t lda, const double *V, int64_t ldv, double *C,
                     int64_t ldc, int64_t rows, int64_t inner, int kk, const double *colscale) {
    MFB_ARG(kk >= 1 && kk <= 64, "tall_times_small: kk must be in 1..64");
    if (kk <= 8) return launch_tall_times_small<8>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    if (kk <= 16) return launch_tall_times_small<16>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    if (kk <= 32) return launch_tall_times_small<32>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    return launch_tall_times_small<64>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
}

// ---------------------------------------------------------------------------
// symmetric eigen-solver, top-k
// ---------------------------------------------------------------------------
#define TRI_THREADS 256

__device__ __forceinline__ double tri_block_sum(double v, double *sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < TRI_THREADS / 32; ++i) s += sh[i];
    return s;
}

// Householder reduction to tridiagonal form (LAPACK dsytd2, lower): A = Q T Q', reflector i stored in A[i+2:N, i]
// (v[i+1] = 1 implicit), tau[i], T = (d, e).  One cooperative launch, ONE pass over the trailing block and ONE
// grid-wide barrier per column: the rank-2 update of step i,  A22 -= v w' + w v',  is fused with the matrix-vector
// product of step i + 1,  p = tau A22' v',  on the freshly updated entries.  That is possible because the next
// reflector only depends on column i + 1 of the updated block, which every block recomputes for itself from the old
// column and the vectors v, w (an O(N) job) before it starts the pass.  Block b owns the columns c with
// c % gridDim.x == b of the (symmetric, fully stored) trailing block; it works on four columns and two rows at a
// time so that every thread has ~16 independent loads in flight.  Arithmetic and summation orders are those of the
// two-pass formulation (update expression a - (v_r w_c + w_r v_c), row-strided sums, warp -> block folds).
// vbuf / pbuf hold two generations each (current / next reflector and product): with one barrier per column a fast
// block writes the next generation while a slow one still reads the current one.
__device__ __forceinline__ void tri_reflector(double alpha, double xnorm, double &beta, double &tq, double &scale) {
    if (xnorm == 0.0) {
        beta = alpha; tq = 0.0; scale = 0.0;
    } else {
        beta = -copysign(hypot(alpha, xnorm), alpha);
        tq = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
    }
}

__global__ void __launch_bounds__(TRI_THREADS) tridiag_kernel(double *A, int64_t N, double *d, double *e,
                                                              double *tau, double *vbuf, double *pbuf,
                                                              double *part) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[TRI_THREADS / 32];
    __shared__ double shq[TRI_THREADS / 32][4];
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    double tq;                   // tau of the current reflector (every block computes it for itself)
    // ---- prologue: reflector 0 from the untouched column 0, p_0 = tau_0 A22 v_0 on the owned columns ----
    {
        const double alpha = __ldcg(&A[1]);
        double q = 0.0;
        for (int64_t r = 2 + tid; r < N; r += TRI_THREADS) { const double v = __ldcg(&A[r]); q += v * v; }
        const double xnorm = sqrt(tri_block_sum(q, sh));
        double beta, scale;
        tri_reflector(alpha, xnorm, beta, tq, scale);
        for (int64_t r = 1 + tid; r < N; r += TRI_THREADS) vbuf[r] = (r == 1) ? 1.0 : __ldcg(&A[r]) * scale;
        if (b == 0 && tid == 0) { d[0] = __ldcg(&A[0]); e[0] = beta; tau[0] = tq; }
        __syncthreads();
        const double *v = vbuf;
        const int64_t cfirst = 1 + (((int64_t)b - 1) % nb + nb) % nb;
        for (int64_t c0 = cfirst; c0 < N; c0 += 4 * (int64_t)nb) {
            const double *col[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int64_t cq = c0 + qq * (int64_t)nb;
                col[qq] = A + ((cq < N) ? cq : c0) * N;
            }
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
            for (int64_t r = 1 + tid; r < N; r += TRI_THREADS) {
                const double vr = v[r];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) s4[qq] += col[qq][r] * vr;
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const double t = warp_sum(s4[qq]);
                if (lane == 0) shq[warp][qq] = t;
            }
            __syncthreads();
            if (tid < 4) {
                const int64_t cq = c0 + tid * (int64_t)nb;
                if (cq < N) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < TRI_THREADS / 32; ++w) t += shq[w][tid];
                    const double pc = tq * t;
                    pbuf[cq] = pc;
                }
            }
            __syncthreads();
        }
        grid.sync();
    }
    for (int64_t i = 0; i + 1 < N; ++i) {
        const int cur = (int)(i & 1);
        const double *v = vbuf + (size_t)cur * N, *pb = pbuf + (size_t)cur * N;
        double *vn = vbuf + (size_t)(cur ^ 1) * N, *pn = pbuf + (size_t)(cur ^ 1) * N;
        // keep reflector i for the back-transformation (nobody reads column i any more)
        if (b == 0)
            for (int64_t r = i + 2 + tid; r < N; r += TRI_THREADS) A[r + i * N] = v[r];
        // w = p - (tau/2)(p'v) v
        // p'v as a fixed-order sum over the columns, formed by every block for itself: the result does not depend on
        // the grid size (the grid is sized by how many cells share the device), and no per-block partial sums have to
        // cross the single barrier of a column
        double pv = 0.0;
#pragma unroll 4
        for (int64_t c = i + 1 + tid; c < N; c += TRI_THREADS) pv += __ldcg(&pb[c]) * v[c];
        pv = tri_block_sum(pv, sh);
        const double alpha2 = -0.5 * tq * pv;
        // column j = i + 1 of the updated block, recomputed by every block: a'[r] = A[r,j] - (v_r w_j + w_r v_j)
        const int64_t j = i + 1;
        const double vj = v[j], wj = __ldcg(&pb[j]) + alpha2 * vj;
        auto upd = [&](int64_t r) {
            const double vr = v[r], wr = __ldcg(&pb[r]) + alpha2 * vr;
            return __ldcg(&A[r + j * N]) - (vr * wj + wr * vj);
        };
        if (j == N - 1) {            // only the last diagonal entry is left
            if (b == 0 && tid == 0) { d[N - 1] = upd(N - 1); e[N - 1] = 0.0; tau[N - 1] = 0.0; }
            break;
        }
        const double alpha = upd(j + 1);
        double q = 0.0;
#pragma unroll 4
        for (int64_t r = j + 2 + tid; r < N; r += TRI_THREADS) { const double a = upd(r); q += a * a; }
        const double xnorm = sqrt(tri_block_sum(q, sh));
        double beta, tqn, scale;
        tri_reflector(alpha, xnorm, beta, tqn, scale);
#pragma unroll 4
        for (int64_t r = j + 1 + tid; r < N; r += TRI_THREADS) vn[r] = (r == j + 1) ? 1.0 : upd(r) * scale;
        if (b == 0 && tid == 0) { d[j] = upd(j); e[j] = beta; tau[j] = tqn; }
        __syncthreads();
        // fused pass over the owned columns c >= j + 1, rows r >= j + 1:  a' = a - (v_r w_c + w_r v_c);  s_c += a' v'_r
        const int64_t cfirst = (j + 1) + (((int64_t)b - (j + 1)) % nb + nb) % nb;
        for (int64_t c0 = cfirst; c0 < N; c0 += 4 * (int64_t)nb) {
            double *col[4];
            double vc[4], wc[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int64_t cq = c0 + qq * (int64_t)nb;
                const bool ok = cq < N;
                col[qq] = ok ? A + cq * N : nullptr;
                vc[qq] = ok ? v[cq] : 0.0;
                wc[qq] = ok ? __ldcg(&pb[cq]) + alpha2 * vc[qq] : 0.0;
            }
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
            // two rows per trip, all loads before the stores (the compiler cannot reorder them itself)
            for (int64_t r = j + 1 + tid; r < N; r += 2 * TRI_THREADS) {
                const int64_t r2 = r + TRI_THREADS;
                const bool ok2 = r2 < N;
                const double vr = v[r], wr = __ldcg(&pb[r]) + alpha2 * vr, nr = vn[r];
                const double vr2 = ok2 ? v[r2] : 0.0, wr2 = ok2 ? __ldcg(&pb[r2]) + alpha2 * vr2 : 0.0;
                const double nr2 = ok2 ? vn[r2] : 0.0;
                double a1[4], a2[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    a1[qq] = col[qq] ? col[qq][r] : 0.0;
                    a2[qq] = (col[qq] && ok2) ? col[qq][r2] : 0.0;
                }
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    if (col[qq]) {
                        const double n1 = a1[qq] - (vr * wc[qq] + wr * vc[qq]);
                        col[qq][r] = n1;
                        s4[qq] += n1 * nr;
                        if (ok2) {
                            const double n2 = a2[qq] - (vr2 * wc[qq] + wr2 * vc[qq]);
                            col[qq][r2] = n2;
                            s4[qq] += n2 * nr2;
                        }
                    }
                }
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const double t = warp_sum(s4[qq]);
                if (lane == 0) shq[warp][qq] = t;
            }
            __syncthreads();
            if (tid < 4) {
                const int64_t cq = c0 + tid * (int64_t)nb;
                if (cq < N) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < TRI_THREADS / 32; ++w) t += shq[w][tid];
                    const double pc = tqn * t;
                    pn[cq] = pc;
                }
            }
            __syncthreads();
        }
        tq = tqn;
        grid.sync();
    }
}

// number of eigenvalues of the symmetric tridiagonal (d, e) that are < x  (Sturm sequence, LAPACK dlaebz style)
__device__ int sturm_count(const double *d, const double *e, int64_t N, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
    for (int64_t i = 1; i < N; ++i) {
        q = d[i] - x - e[i - 1] * e[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

// Top-k eigenpairs of the tridiagonal matrix, one block per eigenpair (cooperative launch, k blocks):
//   * eigenvalue j by 256-way multisection on the Sturm count (7 rounds instead of 55 bisection steps),
//   * inverse iteration: the block's thread 0 runs the pivoted tridiagonal solve in SHARED memory (or in the global
//     workspace when 7 N doubles do not fit), then block 0 re-orthogonalises clustered vectors (dstein's rule:
//     eigenvalues closer than 1e-3 ||T||) by modified Gram-Schmidt in eigenvalue order.
// work: k * 5 * N doubles (only used by the global fall-back).  Z: N x k.
__global__ void __launch_bounds__(TRI_THREADS) tridiag_topk_kernel(const double *d, const double *e, int64_t N, int k,
                                                                   double *evals, double *Z, double *work,
                                                                   int use_smem) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double tsm[];
    __shared__ double sh[TRI_THREADS / 32];
    __shared__ double s_lo, s_hi, s_pivmin, s_norm;
    __shared__ int s_cnt[TRI_THREADS];
    __shared__ double s_lam[64];
    const int tid = threadIdx.x, j = blockIdx.x;
    // tridiagonal in shared memory when it fits
    const double *dd0 = d, *ee0 = e;
    if (use_smem) {
        for (int64_t i = tid; i < N; i += TRI_THREADS) { tsm[i] = d[i]; tsm[N + i] = e[i]; }
        dd0 = tsm; ee0 = tsm + N;
    }
    // Gershgorin bounds
    {
        double lo = INFINITY, hi = -INFINITY, emax = 0.0;
        for (int64_t i = tid; i < N; i += TRI_THREADS) {
            const double el = (i > 0) ? fabs(e[i - 1]) : 0.0, er = (i + 1 < N) ? fabs(e[i]) : 0.0;
            lo = fmin(lo, d[i] - el - er);
            hi = fmax(hi, d[i] + el + er);
            emax = fmax(emax, er * er);
        }
        lo = warp_min(lo); hi = warp_max(hi); emax = warp_max(emax);
        __shared__ double sb[3][TRI_THREADS / 32];
        if ((tid & 31) == 0) { sb[0][tid >> 5] = lo; sb[1][tid >> 5] = hi; sb[2][tid >> 5] = emax; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < TRI_THREADS / 32; ++w) { lo = fmin(lo, sb[0][w]); hi = fmax(hi, sb[1][w]); emax = fmax(emax, sb[2][w]); }
            const double nrm = fmax(fabs(lo), fabs(hi));
            s_lo = lo - 2.0 * nrm * 2.3e-16 * N - 1e-300;
            s_hi = hi + 2.0 * nrm * 2.3e-16 * N + 1e-300;
            s_pivmin = fmax(2.3e-308 * fmax(1.0, emax), 1e-300);
            s_norm = nrm;
        }
        __syncthreads();
    }
    // eigenvalue with index N-1-j (0-based ascending): smallest x with count(x) >= N - j
    {
        const int target = (int)(N - j);
        double lo = s_lo, hi = s_hi;
        for (int round = 0; round < 12; ++round) {
            const double step = (hi - lo) / (TRI_THREADS + 1);
            const double x = lo + step * (tid + 1);
            s_cnt[tid] = (x > lo && x < hi) ? sturm_count(dd0, ee0, N, x, s_pivmin) : -1;
            __syncthreads();
            // first sample whose count reaches the target: every thread scans the (tiny) table in shared memory
            int first = TRI_THREADS;
            for (int t = 0; t < TRI_THREADS; ++t)
                if (s_cnt[t] >= target) { first = t; break; }
            const double nlo = (first == 0) ? lo : lo + step * first;            // sample first-1 (or lo)
            const double nhi = (first == TRI_THREADS) ? hi : lo + step * (first + 1);
            __syncthreads();
            const bool stalled = !(nhi - nlo < hi - lo);
            lo = fmax(lo, nlo); hi = fmin(hi, nhi);
            if (stalled || hi - lo <= 4.0 * 2.3e-16 * fmax(fabs(lo), fabs(hi))) break;
        }
        if (tid == 0) evals[j] = 0.5 * (lo + hi);
    }
    grid.sync();
    // all blocks apply the same separation of coincident eigenvalues (dstein-style perturbation)
    if (tid == 0) {
        const double sep = 10.0 * 2.3e-16 * s_norm;
        for (int q = 0; q < k; ++q) s_lam[q] = __ldcg(&evals[q]);
        for (int q = 1; q < k; ++q)
            if (s_lam[q - 1] - s_lam[q] < sep) s_lam[q] = s_lam[q - 1] - sep;
    }
    __syncthreads();
    double *du, *du2, *dd, *dl, *zl;
    if (use_smem) { du = tsm + 2 * N; du2 = du + N; dd = du2 + N; dl = dd + N; zl = dl + N; }
    else { du = work + (size_t)j * 5 * N; du2 = du + N; dd = du2 + N; dl = dd + N; zl = Z + (size_t)j * N; }
    double *zg = Z + (size_t)j * N;
    for (int iter = 0; iter < 4; ++iter) {
        const double lam = s_lam[j];
        for (int64_t i = tid; i < N; i += TRI_THREADS) {
            dd[i] = dd0[i] - lam;
            du[i] = (i + 1 < N) ? ee0[i] : 0.0;
            dl[i] = (i + 1 < N) ? ee0[i] : 0.0;
            du2[i] = 0.0;
            zl[i] = (iter == 0) ? 1.0 / sqrt((double)N) * (1.0 + 0.01 * (double)((i * 2654435761u + j * 40503u) % 97) / 97.0)
                                : __ldcg(&zg[i]);
        }
        __syncthreads();
        if (tid == 0) {
            // solve (T - lam I) z = b by Gaussian elimination with partial pivoting (tridiagonal)
            double *z = zl;
            const double tiny = 2.3e-16 * s_norm + 1e-300;
            for (int64_t i = 0; i + 1 < N; ++i) {
                if (fabs(dd[i]) >= fabs(dl[i])) {
                    if (fabs(dd[i]) < tiny) dd[i] = copysign(tiny, dd[i] == 0.0 ? 1.0 : dd[i]);
                    const double f = dl[i] / dd[i];
                    dd[i + 1] -= f * du[i];
                    z[i + 1] -= f * z[i];
                    du2[i] = 0.0;
                } else {
                    const double f = dd[i] / dl[i];
                    const double t0 = dd[i + 1];
                    dd[i] = dl[i];
                    dd[i + 1] = du[i] - f * t0;
                    du[i] = t0;
                    if (i + 2 < N) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
                    const double zt = z[i];
                    z[i] = z[i + 1];
                    z[i + 1] = zt - f * z[i + 1];
                }
            }
            if (fabs(dd[N - 1]) < tiny) dd[N - 1] = copysign(tiny, dd[N - 1] == 0.0 ? 1.0 : dd[N - 1]);
            z[N - 1] /= dd[N - 1];
            if (N >= 2) z[N - 2] = (z[N - 2] - du[N - 2] * z[N - 1]) / dd[N - 2];
            for (int64_t i = N - 3; i >= 0; --i) z[i] = (z[i] - du[i] * z[i + 1] - du2[i] * z[i + 2]) / dd[i];
        }
        __syncthreads();
        // crude normalisation against overflow, publish
        double mx = 0.0;
        for (int64_t i = tid; i < N; i += TRI_THREADS) mx = fmax(mx, fabs(zl[i]));
        mx = warp_max(mx);
        __syncthreads();
        if ((tid & 31) == 0) sh[tid >> 5] = mx;
        __syncthreads();
        mx = sh[0];
        for (int w = 1; w < TRI_THREADS / 32; ++w) mx = fmax(mx, sh[w]);
        const double inv = 1.0 / mx;
        for (int64_t i = tid; i < N; i += TRI_THREADS) zg[i] = zl[i] * inv;
        __threadfence();
        grid.sync();
        // block 0: modified Gram-Schmidt in eigenvalue order inside clusters, then normalise every vector
        if (j == 0) {
            for (int q = 0; q < k; ++q) {
                double *zq = Z + (size_t)q * N;
                for (int l = 0; l < q; ++l) {
                    if (fabs(s_lam[l] - s_lam[q]) > 1e-3 * s_norm) continue;
                    const double *zp = Z + (size_t)l * N;       // (other blocks' vectors: read through L2)
                    double s = 0.0;
                    for (int64_t i = tid; i < N; i += TRI_THREADS) s += __ldcg(&zp[i]) * __ldcg(&zq[i]);
                    s = tri_block_sum(s, sh);
                    for (int64_t i = tid; i < N; i += TRI_THREADS) zq[i] = __ldcg(&zq[i]) - s * __ldcg(&zp[i]);
                    __syncthreads();
                }
                double s = 0.0;
                for (int64_t i = tid; i < N; i += TRI_THREADS) { const double t = __ldcg(&zq[i]); s += t * t; }
                s = tri_block_sum(s, sh);
                const double inv2 = 1.0 / sqrt(s);
                for (int64_t i = tid; i < N; i += TRI_THREADS) zq[i] = __ldcg(&zq[i]) * inv2;
                __syncthreads();
            }
            __threadfence();
        }
        grid.sync();
    }
}

// z <- H(0) H(1) ... H(N-2) z, one block per eigenvector column
__global__ void __launch_bounds__(TRI_THREADS) backtransform_kernel(const double *A, int64_t N, const double *tau,
                                                                    double *Z) {
    __shared__ double sh[TRI_THREADS / 32];
    double *z = Z + (size_t)blockIdx.x * N;
    const int tid = threadIdx.x;
    for (int64_t i = N - 2; i >= 0; --i) {
        const double tq = tau[i];
        if (tq == 0.0) continue;
        double s = 0.0;
        for (int64_t r = i + 1 + tid; r < N; r += TRI_THREADS) s += ((r == i + 1) ? 1.0 : A[r + i * N]) * z[r];
        s = tq * tri_block_sum(s, sh);
        for (int64_t r = i + 1 + tid; r < N; r += TRI_THREADS) z[r] -= s * ((r == i + 1) ? 1.0 : A[r + i * N]);
        __syncthreads();
    }
}

int sym_eig_topk(mfb_context *ctx, double *G, int64_t N, int k, double *evals, double *evecs) {
    MFB_ARG(N >= 1 && k >= 1 && k <= 64 && k <= N, "sym_eig_topk: need 1 <= k <= min(N, 64)");
    cudaStream_t st = ctx->stream;
    DevBuf buf(ctx);
    double *wk = nullptr;
    const size_t nwork = (size_t)7 * N + (size_t)4 * ctx->sm_count + (size_t)k * 5 * N + 16;
    {
        int rc = buf.alloc(&wk, nwork);
        if (rc != MFB_OK) return rc;
    }
    double *d = wk, *e = d + N, *tau = e + N, *vbuf = tau + N, *pbuf = vbuf + 2 * N, *part = pbuf + 2 * N;
    double *work = part + 4 * ctx->sm_count;
    if (N == 1) {
        MFB_CUDA(cudaMemcpyAsync(evals, G, sizeof(double), cudaMemcpyDeviceToDevice, st));
        const double one = 1.0;
        MFB_CUDA(cudaMemcpyAsync(evecs, &one, sizeof(double), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        return MFB_OK;
    }
    int nb = (int)((N + 7) / 8);          // one warp per column of the trailing block, 8 warps per block
    {
        // large matrices stream from HBM: two blocks per SM keep twice the loads in flight
        int per_sm = 1;
        if (N >= 4096 && !getenv("MFB_TRI_ONE_BLOCK")) {
            int occ = 0;
            MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tridiag_kernel, TRI_THREADS, 0));
            if (occ >= 2) per_sm = 2;
        }
        if (nb > per_sm * ctx->sm_count) nb = per_sm * ctx->sm_count;
    }
    if (nb > 2 * ctx->sm_count) nb = 2 * ctx->sm_count;
    // small (L2-resident, barrier-bound) matrices come from bcv cells, several of which run concurrently on a device
    // (mfb_context_set_concurrency): with full grids every SM hosts a block of each cell and their grid barriers fight
    // for the same issue slots.  36 cells of D3: 3 streams x 148 blocks 0.38 s, 3 x 98 0.32 s, 6 x 48 0.22 s; a single
    // stream is fastest with the full grid.  (The results do not depend on the grid size.)
    if (N < 4096) {
        const int share = ctx->concurrency < 2 ? 2 : ctx->concurrency;
        int cap = 2 * ctx->sm_count / share;
        if (cap < 32) cap = 32;
        if (nb > cap) nb = cap;
    }
    if (const char *cap = getenv("MFB_TRI_BLOCKS")) {      // development aid: cap the grid of the tridiagonalisation
        const int c = atoi(cap);
        if (c >= 1 && c < nb) nb = c;
    }
    if (nb < 1) nb = 1;
    void *args[] = {&G, &N, &d, &e, &tau, &vbuf, &pbuf, &part};
    MFB_CUDA(cudaLaunchCooperativeKernel((void *)tridiag_kernel, dim3(nb), dim3(TRI_THREADS), args, 0, st));
    {
        size_t tsm = (size_t)7 * N * sizeof(double);
        int use_smem = tsm <= 200 * 1024;
        if (!use_smem) tsm = 0;
        MFB_CUDA(cudaFuncSetAttribute(tridiag_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int kk = k;
        void *targs[] = {&d, &e, &N, &kk, &evals, &evecs, &work, &use_smem};
        MFB_CUDA(cudaLaunchCooperativeKernel((void *)tridiag_topk_kernel, dim3(k), dim3(TRI_THREADS), targs, tsm, st));
    }
    backtransform_kernel<<<k, TRI_THREADS, 0, st>>>(G, N, tau, evecs);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}
