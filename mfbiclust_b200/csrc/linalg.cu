// Dense FP64 building blocks for svd_pca and the bcv cells (see linalg.cuh).
#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>

#include "linalg.cuh"

#define MFB_TRY_RC(expr) do { int _r = (expr); if (_r != MFB_OK) return _r; } while (0)

namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// C = A' B  (dot products of contiguous columns), FP64 tensor cores
// ---------------------------------------------------------------------------
// One warp owns a 32 x 32 tile of C: 4 x 4 m8n8k4 blocks, accumulators in registers; per 16-deep k step a lane
// reads 4 consecutive k of one column per block (A-operand [M = column of A][K], B-operand [K][N = column of B]).
__global__ void __launch_bounds__(128) gemm_tn_kernel(const double *__restrict__ A, int64_t lda,
                                                      const double *__restrict__ B, int64_t ldb,
                                                      double *__restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                      int64_t K, int64_t kz, int vec_ok, int sym) {
    // sym: C = A'A -- only the block tiles on and below the diagonal are computed; dot(a_i, a_j) and dot(a_j, a_i)
    // are the same products in the same order, so the mirrored entry is the same number
    if (sym && blockIdx.y > blockIdx.x) return;
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
                if (mr < M && nc < N) {
                    Cz[mr + nc * ldc] = acc[i][j][e];
                    if (sym == 2 && blockIdx.y < blockIdx.x) Cz[nc + mr * ldc] = acc[i][j][e];   // direct write: mirror
                }
            }
        }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double *__restrict__ part, int Z, int64_t M,
                                                            int64_t N, int64_t ldp, double *__restrict__ C,
                                                            int64_t ldc, int sym) {
    const int64_t total = M * N;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t n = e / M, m = e - n * M;
        // symmetric product: the partials of a block tile above the diagonal live in its mirror image
        const bool flip = sym && (m / 64 < n / 64);
        const int64_t pm = flip ? n : m, pn = flip ? m : n;
        double s = 0.0;
        for (int z = 0; z < Z; ++z) s += part[(size_t)z * ldp * N + pm + pn * ldp];
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
    const int sym = (A == B && lda == ldb && M == N) ? 1 : 0;       // Gram matrix: half the tiles
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)Z);
    if (Z == 1) {
        gemm_tn_kernel<<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, M, N, K, kz, vec_ok, sym ? 2 : 0);
    } else {
        MFB_ARG(work != nullptr, "gemm_tn: split-K needs a workspace");
        gemm_tn_kernel<<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, work, M, M, N, K, kz, vec_ok, sym);
        int nb = (int)((M * N + 255) / 256);
        if (nb > 4 * ctx->sm_count) nb = 4 * ctx->sm_count;
        splitk_reduce_kernel<<<nb, 256, 0, ctx->stream>>>(work, Z, M, N, M, C, ldc, sym);
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

// ---------------------------------------------------------------------------
// C = A V  for a tall A and a thin V (64 columns of V per launch)
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
    MFB_ARG(kk >= 1, "tall_times_small: kk must be positive");
    if (kk > 64) {                                  // 64 output columns per launch
        for (int j0 = 0; j0 < kk; j0 += 64) {
            const int kc = (kk - j0 < 64) ? kk - j0 : 64;
            const int rc = tall_times_small(ctx, A, lda, V + (int64_t)j0 * ldv, ldv, C + (int64_t)j0 * ldc, ldc, rows,
                                            inner, kc, colscale ? colscale + j0 : nullptr);
            if (rc != MFB_OK) return rc;
        }
        return MFB_OK;
    }
    if (kk <= 8) return launch_tall_times_small<8>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    if (kk <= 16) return launch_tall_times_small<16>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    if (kk <= 32) return launch_tall_times_small<32>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
    return launch_tall_times_small<64>(ctx, A, lda, V, ldv, C, ldc, rows, inner, kk, colscale);
}

// ---------------------------------------------------------------------------
// symmetric eigen-solver, top-k: Lanczos tridiagonalisation with full re-orthogonalisation
// ---------------------------------------------------------------------------
// svd_pca and the bcv cells need the kk LARGEST eigenpairs of a Gram matrix G (N x N), kk << N in the configurations
// that cost time (kk = 20 of N = 1333 per D3 cell, 50 of 13 333 per D5 cell).  A Householder reduction touches the
// whole trailing block once per column (N passes, 16 N^3 / 3 bytes); the Lanczos recurrence
//     G Q_J = Q_J T_J + beta_J q_{J+1} e_J',      Q_J orthonormal (classical Gram-Schmidt applied twice per step)
// reads G once per STEP and the leading Ritz pairs of the tridiagonal T_J converge to machine precision after
// J ~ 10 kk steps on these spectra (planted components + the edge of a noise bulk) -- 208 steps instead of 1333
// columns on a D3 cell, ~550 instead of 13 333 on D5, and nothing is written back to G.  Carried to J = N the
// recurrence IS a complete orthogonal tridiagonalisation, so small problems and kk close to N need no second code
// path.  T_J is handed to the bisection / inverse-iteration kernel below; a Ritz pair (theta, s) is accepted when
// its residual  |beta_J s_J|  (= ||G x - theta x||, x = Q_J s) is below 2e-13 theta_1.
//
// One persistent cooperative launch runs a chunk of steps; per step five grid barriers:
//   matvec (a warp per column of G: a dot product of two contiguous vectors) | dots h = Q'w (a warp per basis
//   vector) | update w -= Q h (a 32-row slab per block, the basis vectors split over its warps) | dots | update + norm.
// Every sum has a fixed order that does not depend on the grid size, so results are bit-identical for any number of
// concurrent cells (the grid is sized by mfb_context_set_concurrency).  Exact breakdown (an invariant subspace: G of
// low rank, repeated blocks) continues with a deterministic pseudo-random vector orthogonalised against the basis.
#define LZ_THREADS 512
#define LZ_WARPS (LZ_THREADS / 32)

struct LanczosArgs {
    const double *G;
    int64_t ldg, N;
    double *Q;            // N x (Jcap + 1), ld = ldq
    int64_t ldq;
    double *alpha, *beta; // Jcap each: T = tridiag(beta, alpha, beta); beta[J-1] = residual norm of the J-step recurrence
    double *wa, *wb;      // N each: candidate for the next basis vector / product being orthogonalised
    double *h;            // 2 x hstride: coefficients of the two Gram-Schmidt passes
    int64_t hstride;
    double *nrm;          // per 32-row slab: sum of squares of the candidate
    double *state;        // [0] anorm, [1] which of wa / wb holds the candidate, [2] steps done, [3] error flag
    unsigned *bar;        // grid barrier counter (zeroed before every launch)
    int j_begin, j_end;
    int w_in_smem;
};

__device__ __forceinline__ bool lz_barrier(unsigned *bar, unsigned &target, unsigned nb, int *abort_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nb;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int)(v - target) >= 0) break;
            if ((++spins & 0x3ffu) == 0) {          // bounded wait: a lost block must not hang the device
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) { *abort_flag = 1; break; }
            }
        }
    }
    __syncthreads();
    return *abort_flag != 0;
}

// fixed-order block sum (thread-strided partials -> warp shuffles -> warps in order); every thread gets the result
__device__ __forceinline__ double lz_block_sum(double v, double *sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < LZ_WARPS; ++i) s += sh[i];
    return s;
}

__device__ __forceinline__ double lz_start_entry(int64_t r, int salt) {
    // deterministic and well spread (a constant vector would lie in the null space of a centred row-side Gram matrix)
    unsigned x = (unsigned)(r * 2654435761u) ^ (unsigned)(salt * 40503u + 0x9e3779b9u);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    const double u = (double)(x & 0xffffffu) / 16777216.0;          // [0, 1)
    return 2.0 * u - 1.0;
}

__global__ void __launch_bounds__(LZ_THREADS) lanczos_kernel(LanczosArgs P) {
    extern __shared__ double lsm[];
    __shared__ double sh[LZ_WARPS];
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = gridDim.x, b = blockIdx.x;
    const int64_t N = P.N, ldq = P.ldq, ldg = P.ldg;
    const int64_t nslab = (N + 31) / 32;
    const int WT = nb * LZ_WARPS, gw = b * LZ_WARPS + warp;
    double *red = lsm;                                  // LZ_WARPS x 32
    double *hs = red + LZ_WARPS * 32;                   // j_end + 1 coefficients
    double *ws = hs + (P.j_end + 1);                    // N entries of the current vector (when it fits)
    unsigned target = 0;
    if (tid == 0) s_abort = 0;
    __syncthreads();

    // || cand ||^2 from the slab partials, in slab order
    auto cand_norm2 = [&]() {
        double v = 0.0;
        for (int64_t s = tid; s < nslab; s += LZ_THREADS) v += __ldcg(&P.nrm[s]);
        return lz_block_sum(v, sh);
    };
    // two passes of classical Gram-Schmidt of `buf` against Q[:, 0 .. cnt-1]; leaves the slab norms of the result
    auto reorth = [&](double *buf, int cnt) -> bool {
        for (int pass = 0; pass < 2; ++pass) {
            double *hp = P.h + (size_t)pass * P.hstride;
            if (P.w_in_smem) {
                for (int64_t r = tid; r < N; r += LZ_THREADS) ws[r] = __ldcg(&buf[r]);
                __syncthreads();
            }
            const double *wv = P.w_in_smem ? ws : buf;
            for (int i = gw; i < cnt; i += WT) {
                const double *col = P.Q + (size_t)i * ldq;
                double acc = 0.0;
                int64_t r = lane;
                for (; r + 224 < N; r += 256) {
                    double q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) q[u] = __ldcg(&col[r + 32 * u]);
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc += q[u] * (P.w_in_smem ? wv[r + 32 * u] : __ldcg(&wv[r + 32 * u]));
                }
                for (; r < N; r += 32) acc += __ldcg(&col[r]) * (P.w_in_smem ? wv[r] : __ldcg(&wv[r]));
                acc = warp_sum(acc);
                if (lane == 0) hp[i] = acc;
            }
            if (lz_barrier(P.bar, target, nb, &s_abort)) return true;
            for (int i = tid; i < cnt; i += LZ_THREADS) hs[i] = __ldcg(&hp[i]);
            __syncthreads();
            const int i0 = (int)(((int64_t)cnt * warp) / LZ_WARPS), i1 = (int)(((int64_t)cnt * (warp + 1)) / LZ_WARPS);
            for (int64_t s = b; s < nslab; s += nb) {
                const int64_t r = 32 * s + lane;
                const bool ok = r < N;
                const double *row = P.Q + (ok ? r : 0);
                double acc = 0.0;
                int i = i0;
                for (; i + 8 <= i1; i += 8) {
                    double q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) q[u] = __ldcg(&row[(size_t)(i + u) * ldq]);
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc += q[u] * hs[i + u];
                }
                for (; i < i1; ++i) acc += __ldcg(&row[(size_t)i * ldq]) * hs[i];
                red[warp * 32 + lane] = acc;
                __syncthreads();
                if (warp == 0) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < LZ_WARPS; ++w) t += red[w * 32 + lane];
                    double v = 0.0;
                    if (ok) { v = __ldcg(&buf[r]) - t; buf[r] = v; }
                    const double sq = warp_sum(v * v);
                    if (lane == 0) P.nrm[s] = sq;
                }
                __syncthreads();
            }
            if (lz_barrier(P.bar, target, nb, &s_abort)) return true;
        }
        return false;
    };

    double anorm = __ldcg(&P.state[0]);
    double *cand = (__ldcg(&P.state[1]) == 0.0) ? P.wa : P.wb;
    double *out = (cand == P.wa) ? P.wb : P.wa;
    int j = P.j_begin;
    if (j == 0) {
        // start vector (un-normalised) and its slab norms
        for (int64_t s = b; s < nslab; s += nb) {
            if (warp == 0) {
                const int64_t r = 32 * s + lane;
                double v = 0.0;
                if (r < N) { v = lz_start_entry(r, 0); cand[r] = v; }
                const double sq = warp_sum(v * v);
                if (lane == 0) P.nrm[s] = sq;
            }
        }
        anorm = 0.0;
        if (lz_barrier(P.bar, target, nb, &s_abort)) goto fail;
    }
    {
        bool injected = false;
        int inject_tries = 0;
        for (;;) {
            const double beta_prev = sqrt(cand_norm2());
            if (j > 0 && !injected && b == 0 && tid == 0) P.beta[j - 1] = beta_prev;
            if (j == P.j_end) break;
            if ((j > 0 && !injected && beta_prev <= 1e-12 * anorm) || (injected && !(beta_prev > 1e-6))) {
                // invariant subspace: T decouples (beta = 0); continue with a fresh direction orthogonal to the basis
                if (b == 0 && tid == 0) P.beta[j - 1] = 0.0;
                if (++inject_tries > 4) {              // nothing left outside span(Q): the basis is complete
                    if (b == 0 && tid == 0) P.state[3] = 2.0;
                    break;
                }
                for (int64_t s = b; s < nslab; s += nb) {
                    if (warp == 0) {
                        const int64_t r = 32 * s + lane;
                        if (r < N) cand[r] = lz_start_entry(r, j * 8 + inject_tries);
                    }
                }
                if (lz_barrier(P.bar, target, nb, &s_abort)) goto fail;
                if (reorth(cand, j)) goto fail;
                injected = true;
                continue;
            }
            injected = false;
            inject_tries = 0;
            const double inv = 1.0 / beta_prev;
            // q_j = cand / beta: every block keeps a copy in shared memory, the slab owners store it into Q[:, j]
            if (P.w_in_smem) {
                for (int64_t r = tid; r < N; r += LZ_THREADS) ws[r] = __ldcg(&cand[r]) * inv;
                __syncthreads();
            }
            for (int64_t s = b; s < nslab; s += nb) {
                if (warp == 0) {
                    const int64_t r = 32 * s + lane;
                    if (r < N) P.Q[r + (size_t)j * ldq] = __ldcg(&cand[r]) * inv;
                }
            }
            // out = G q_j: a warp per pair of columns (G symmetric: a column is a row)
            for (int64_t c0 = 2 * (int64_t)gw; c0 < N; c0 += 2 * (int64_t)WT) {
                const double *g0 = P.G + (size_t)c0 * ldg;
                const bool two = c0 + 1 < N;
                const double *g1 = two ? g0 + ldg : g0;
                double a0 = 0.0, a1 = 0.0;
                int64_t r = lane;
                for (; r + 96 < N; r += 128) {
                    double x0[4], x1[4], qv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { x0[u] = __ldcg(&g0[r + 32 * u]); x1[u] = __ldcg(&g1[r + 32 * u]); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) qv[u] = P.w_in_smem ? ws[r + 32 * u] : __ldcg(&cand[r + 32 * u]) * inv;
#pragma unroll
                    for (int u = 0; u < 4; ++u) { a0 += x0[u] * qv[u]; a1 += x1[u] * qv[u]; }
                }
                for (; r < N; r += 32) {
                    const double qv = P.w_in_smem ? ws[r] : __ldcg(&cand[r]) * inv;
                    a0 += __ldcg(&g0[r]) * qv;
                    a1 += __ldcg(&g1[r]) * qv;
                }
                a0 = warp_sum(a0);
                a1 = warp_sum(a1);
                if (lane == 0) { out[c0] = a0; if (two) out[c0 + 1] = a1; }
            }
            if (lz_barrier(P.bar, target, nb, &s_abort)) goto fail;
            if (reorth(out, j + 1)) goto fail;
            const double aj = __ldcg(&P.h[j]) + __ldcg(&P.h[P.hstride + j]);
            if (b == 0 && tid == 0) P.alpha[j] = aj;
            {
                double nb2 = 0.0;
                for (int64_t s = tid; s < nslab; s += LZ_THREADS) nb2 += __ldcg(&P.nrm[s]);
                nb2 = lz_block_sum(nb2, sh);
                anorm = fmax(anorm, fabs(aj) + beta_prev * (j > 0 ? 1.0 : 0.0) + sqrt(nb2));
            }
            double *t = cand; cand = out; out = t;
            ++j;
        }
    }
    if (b == 0 && tid == 0) {
        P.state[0] = anorm;
        P.state[1] = (cand == P.wa) ? 0.0 : 1.0;
        P.state[2] = (double)j;
    }
    return;
fail:
    if (tid == 0) P.state[3] = 1.0;
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

// Top-k eigenpairs of the tridiagonal matrix (cooperative launch; block b owns the eigenpairs b, b + gridDim.x, ...):
//   * eigenvalue j by 256-way multisection on the Sturm count (7 rounds instead of 55 bisection steps),
//   * inverse iteration: the block's thread 0 runs the pivoted tridiagonal solve in SHARED memory (or in the global
//     workspace when 7 N doubles do not fit), then block 0 re-orthogonalises clustered vectors (dstein's rule:
//     eigenvalues closer than 1e-3 ||T||) by modified Gram-Schmidt in eigenvalue order.
// work: gridDim.x * 5 * N doubles (only used by the global fall-back).  Z: N x k (ld = N).  lam: k doubles (scratch).
// resid (optional): resid[j] = |last component of eigenvector j| -- times beta_J it is the Lanczos residual norm.
__global__ void __launch_bounds__(TRI_THREADS) tridiag_topk_kernel(const double *d, const double *e, int64_t N, int k,
                                                                   double *evals, double *Z, double *work,
                                                                   double *lam, int use_smem, double *resid) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double tsm[];
    __shared__ double sh[TRI_THREADS / 32];
    __shared__ double s_lo, s_hi, s_pivmin, s_norm;
    __shared__ int s_cnt[TRI_THREADS];
    const int tid = threadIdx.x, nbk = gridDim.x;
    // tridiagonal in shared memory when it fits
    const double *dd0 = d, *ee0 = e;
    if (use_smem) {
        for (int64_t i = tid; i < N; i += TRI_THREADS) { tsm[i] = d[i]; tsm[N + i] = (i + 1 < N) ? e[i] : 0.0; }
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
    for (int j = blockIdx.x; j < k; j += nbk) {
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
    __threadfence();
    grid.sync();
    // one block applies the separation of coincident eigenvalues (dstein-style perturbation) for everybody
    if (blockIdx.x == 0 && tid == 0) {
        const double sep = 10.0 * 2.3e-16 * s_norm;
        double prev = __ldcg(&evals[0]);
        lam[0] = prev;
        for (int q = 1; q < k; ++q) {
            double v = __ldcg(&evals[q]);
            if (prev - v < sep) v = prev - sep;
            lam[q] = v;
            prev = v;
        }
        __threadfence();
    }
    grid.sync();
    double *du, *du2, *dd, *dl, *zl;
    if (use_smem) { du = tsm + 2 * N; du2 = du + N; dd = du2 + N; dl = dd + N; zl = dl + N; }
    else { du = work + (size_t)blockIdx.x * 5 * N; du2 = du + N; dd = du2 + N; dl = dd + N; zl = nullptr; }
    for (int iter = 0; iter < 4; ++iter) {
        for (int j = blockIdx.x; j < k; j += nbk) {
            double *zg = Z + (size_t)j * N;
            double *z = use_smem ? zl : zg;
            const double lamj = __ldcg(&lam[j]);
            for (int64_t i = tid; i < N; i += TRI_THREADS) {
                dd[i] = dd0[i] - lamj;
                du[i] = (i + 1 < N) ? ee0[i] : 0.0;
                dl[i] = (i + 1 < N) ? ee0[i] : 0.0;
                du2[i] = 0.0;
                const double zi = (iter == 0) ? 1.0 / sqrt((double)N) * (1.0 + 0.01 * (double)((i * 2654435761u + j * 40503u) % 97) / 97.0)
                                              : __ldcg(&zg[i]);
                z[i] = zi;
            }
            __syncthreads();
            if (tid == 0) {
                // solve (T - lam I) z = b by Gaussian elimination with partial pivoting (tridiagonal)
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
            for (int64_t i = tid; i < N; i += TRI_THREADS) mx = fmax(mx, fabs(z[i]));
            mx = warp_max(mx);
            __syncthreads();
            if ((tid & 31) == 0) sh[tid >> 5] = mx;
            __syncthreads();
            mx = sh[0];
            for (int w = 1; w < TRI_THREADS / 32; ++w) mx = fmax(mx, sh[w]);
            const double inv = 1.0 / mx;
            for (int64_t i = tid; i < N; i += TRI_THREADS) zg[i] = z[i] * inv;
            __syncthreads();
        }
        __threadfence();
        grid.sync();
        // block 0: modified Gram-Schmidt in eigenvalue order inside clusters, then normalise every vector
        if (blockIdx.x == 0) {
            for (int q = 0; q < k; ++q) {
                double *zq = Z + (size_t)q * N;
                const double lq = __ldcg(&lam[q]);
                for (int l = 0; l < q; ++l) {
                    if (fabs(__ldcg(&lam[l]) - lq) > 1e-3 * s_norm) continue;
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
                if (resid && iter == 3 && tid == 0) resid[q] = fabs(zq[N - 1]);
            }
            __threadfence();
        }
        grid.sync();
    }
}

// max_j beta * resid[j] against tol * max(|evals[0]|, |evals[k-1]|)  ->  flag[0] = 1 when every pair has converged
__global__ void lanczos_check_kernel(const double *resid, const double *evals, int k, const double *beta_last,
                                     double tol, double *flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double b = fabs(beta_last[0]);
    double mx = 0.0;
    for (int j = 0; j < k; ++j) mx = fmax(mx, b * resid[j]);
    const double scale = fmax(fabs(evals[0]), fabs(evals[k - 1]));
    flag[0] = (mx <= tol * scale) ? 1.0 : 0.0;
    flag[1] = mx;
    flag[2] = scale;
}

static int lanczos_grid(mfb_context *ctx, int64_t N, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lanczos_kernel, LZ_THREADS, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return 0;
    }
    // enough warps for a column pair each, never more blocks than can be resident
    int64_t want = (N / 2 + LZ_WARPS - 1) / LZ_WARPS;
    const int64_t slabs = (N + 31) / 32;
    if (want < 1) want = 1;
    int per_sm = occ > 2 ? 2 : occ;
    int nb = (int)std::min<int64_t>(want, (int64_t)per_sm * ctx->sm_count);
    // small (L2-resident, barrier-bound) problems come from bcv cells, several of which run concurrently on a
    // device (mfb_context_set_concurrency): every cell gets its share of the SMs.  Results do not depend on the grid.
    if (N < 4096) {
        const int share = ctx->concurrency < 1 ? 1 : ctx->concurrency;
        int cap = (ctx->sm_count * (share > 1 ? 2 : 1)) / share;
        if (cap < 16) cap = 16;
        if (nb > cap) nb = cap;
    }
    (void)slabs;
    if (const char *capenv = getenv("MFB_LZ_BLOCKS")) {      // development aid
        const int c = atoi(capenv);
        if (c >= 1 && c < nb) nb = c;
    }
    return nb < 1 ? 1 : nb;
}

int sym_eig_topk(mfb_context *ctx, const double *G, int64_t ldg, int64_t N, int k, double *evals, double *evecs) {
    MFB_ARG(N >= 1 && k >= 1 && k <= N && ldg >= N, "sym_eig_topk: need 1 <= k <= N <= ldg");
    cudaStream_t st = ctx->stream;
    if (N == 1) {
        MFB_CUDA(cudaMemcpyAsync(evals, G, sizeof(double), cudaMemcpyDeviceToDevice, st));
        const double one = 1.0;
        MFB_CUDA(cudaMemcpyAsync(evecs, &one, sizeof(double), cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        return MFB_OK;
    }
    // steps: ~10 k on the spectra of the path (see the header); everything when k is a large part of N
    int64_t J = 10 * (int64_t)k + 32;
    if (const char *e = getenv("MFB_LZ_FIRST")) J = atoll(e) * k + 32;
    if (J > N || 3 * (int64_t)k >= N) J = N;
    int64_t Jcap = std::min<int64_t>(N, J + J / 2 + 64);
    const int64_t ldq = round_up(N, 16);
    DevBuf buf(ctx);
    double *Q = nullptr, *sc = nullptr, *S = nullptr, *twork = nullptr;
    const int64_t nslab = (N + 31) / 32;
    auto alloc_q = [&](int64_t cap, double **q) { return buf.alloc(q, (size_t)ldq * (cap + 1)); };
    { int rc = alloc_q(Jcap, &Q); if (rc != MFB_OK) return rc; }
    // scalars / small vectors: alpha, beta, wa, wb, h(2), nrm, state, flag, resid, lam, barrier
    const size_t n_sc = (size_t)2 * N + 2 * N + 2 * (N + 1) + nslab + 8 + 8 + N + N + 8;
    { int rc = buf.alloc(&sc, n_sc); if (rc != MFB_OK) return rc; }
    double *alpha = sc, *beta = alpha + N, *wa = beta + N, *wb = wa + N, *h = wb + N, *nrm = h + 2 * (N + 1);
    double *state = nrm + nslab, *flag = state + 8, *resid = flag + 8, *lam = resid + N;
    unsigned *bar = (unsigned *)(lam + N);
    MFB_CUDA(cudaMemsetAsync(state, 0, 16 * sizeof(double), st));
    LanczosArgs P;
    P.G = G; P.ldg = ldg; P.N = N; P.Q = Q; P.ldq = ldq; P.alpha = alpha; P.beta = beta; P.wa = wa; P.wb = wb;
    P.h = h; P.hstride = N + 1; P.nrm = nrm; P.state = state; P.bar = bar;
    int64_t done = 0;
    double *hflag = (double *)ctx->pinned;             // (one call at a time per context)
    const double tol = 2e-13;
    int topk_blocks = 0, use_smem = 0;
    size_t tsm = 0;
    for (int round = 0;; ++round) {
        // ---- a chunk of Lanczos steps: done -> J ----
        P.j_begin = (int)done; P.j_end = (int)J;
        size_t smem = (size_t)(LZ_WARPS * 32 + (J + 1) + N) * sizeof(double);
        P.w_in_smem = 1;
        if (smem > 200 * 1024) { P.w_in_smem = 0; smem = (size_t)(LZ_WARPS * 32 + (J + 1)) * sizeof(double); }
        MFB_ARG(smem <= 200 * 1024, "sym_eig_topk: too many Lanczos steps for the shared-memory coefficient buffer");
        MFB_CUDA(cudaFuncSetAttribute(lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int nb = lanczos_grid(ctx, N, smem);
        MFB_ARG(nb >= 1, "sym_eig_topk: the Lanczos kernel does not fit on this device");
        MFB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
        void *args[] = {&P};
        MFB_CUDA(cudaLaunchCooperativeKernel((void *)lanczos_kernel, dim3(nb), dim3(LZ_THREADS), args, smem, st));
        done = J;
        // ---- Ritz pairs of T_J ----
        if (!S) { int rc = buf.alloc(&S, (size_t)N * k); if (rc != MFB_OK) return rc; }   // J x k, J <= N
        tsm = (size_t)7 * J * sizeof(double);
        use_smem = tsm <= 200 * 1024;
        if (!use_smem) tsm = 0;
        MFB_CUDA(cudaFuncSetAttribute(tridiag_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        {
            int occ = 0;
            MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tridiag_topk_kernel, TRI_THREADS, tsm));
            if (occ < 1) occ = 1;
            if (occ > 2) occ = 2;
            topk_blocks = std::min<int64_t>(k, (int64_t)occ * ctx->sm_count);
            if (N < 4096 && ctx->concurrency > 1) topk_blocks = std::min(topk_blocks, std::max(8, 2 * ctx->sm_count / ctx->concurrency));
        }
        if (!use_smem && !twork) { int rc = buf.alloc(&twork, (size_t)2 * ctx->sm_count * 5 * N); if (rc != MFB_OK) return rc; }
        int kk = k;
        int64_t Jn = J;
        void *targs[] = {&alpha, &beta, &Jn, &kk, &evals, &S, &twork, &lam, &use_smem, &resid};
        MFB_CUDA(cudaLaunchCooperativeKernel((void *)tridiag_topk_kernel, dim3(topk_blocks), dim3(TRI_THREADS), targs, tsm, st));
        if (J >= N) break;                           // complete tridiagonalisation: exact
        const double *beta_last = beta + (J - 1);
        lanczos_check_kernel<<<1, 32, 0, st>>>(resid, evals, k, beta_last, tol, flag);
        MFB_CUDA(cudaGetLastError());
        MFB_CUDA(cudaMemcpyAsync(hflag, flag, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaMemcpyAsync(hflag + 4, state, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        if (hflag[4 + 3] == 1.0) { mfb_set_error("sym_eig_topk: grid barrier timed out in the Lanczos kernel"); return MFB_ERR_CUDA; }
        if (hflag[0] == 1.0 || hflag[4 + 3] == 2.0) break;
        // not there yet: 25 % more steps (the basis grows; copy it when the first allocation is too small)
        int64_t Jnext = std::min<int64_t>(N, J + std::max<int64_t>(J / 4, 16));
        if (Jnext > Jcap) {
            double *Q2 = nullptr;
            const int64_t cap2 = std::min<int64_t>(N, std::max<int64_t>(2 * Jcap, Jnext));
            { int rc = alloc_q(cap2, &Q2); if (rc != MFB_OK) return rc; }
            MFB_CUDA(cudaMemcpyAsync(Q2, Q, (size_t)ldq * (Jcap + 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
            Q = Q2; P.Q = Q2; Jcap = cap2;
        }
        J = Jnext;
    }
    {
        MFB_CUDA(cudaMemcpyAsync(hflag + 4, state, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        if (hflag[4 + 3] == 1.0) { mfb_set_error("sym_eig_topk: grid barrier timed out in the Lanczos kernel"); return MFB_ERR_CUDA; }
    }
    // Ritz vectors: evecs (N x k, ld = N) = Q_J S
    MFB_TRY_RC(tall_times_small(ctx, Q, ldq, S, done, evecs, N, N, done, k, nullptr));
    MFB_CUDA(cudaStreamSynchronize(st));
    ctx->last_lanczos_steps = (int)done;
    if (getenv("MFB_LZ_VERBOSE")) fprintf(stderr, "sym_eig_topk: N=%lld k=%d steps=%lld\n", (long long)N, k, (long long)done);
    return MFB_OK;
}
