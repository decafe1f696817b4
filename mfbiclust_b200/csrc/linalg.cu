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

static int splitk_factor(int64_t M, int64_t N, int64_t K, int sm_count, bool sym) {
    // a block = four warps with a 32 x 32 tile each, fed by direct global loads: it takes ~4 blocks per SM to hide
    // the load latency.  A symmetric product only runs the block tiles on and below the diagonal.
    const int64_t tm = (M + 63) / 64, tn = (N + 63) / 64;
    const int64_t tiles = sym ? tm * (tm + 1) / 2 : tm * tn;
    int Z = 1;
    if (tiles < 4 * sm_count) {
        Z = (int)((4 * sm_count + tiles - 1) / tiles);
        const int64_t maxz = (K + 255) / 256;
        if (Z > maxz) Z = (int)maxz;
        if (Z < 1) Z = 1;
        if (Z > 64) Z = 64;
    }
    return Z;
}

size_t gemm_tn_workspace(int64_t M, int64_t N, int64_t K, int sm_count) {
    int Z = splitk_factor(M, N, K, sm_count, false);
    if (M == N) Z = std::max(Z, splitk_factor(M, N, K, sm_count, true));    // (the caller may pass A == B)
    return Z > 1 ? (size_t)Z * M * N : 0;
}

int gemm_tn(mfb_context *ctx, const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
            int64_t M, int64_t N, int64_t K, double *work) {
    const int sym = (A == B && lda == ldb && M == N) ? 1 : 0;       // Gram matrix: half the tiles
    const int Z = splitk_factor(M, N, K, ctx->sm_count, sym != 0);
    int64_t kz = (K + Z - 1) / Z;
    kz = (kz + 15) / 16 * 16;
    const int vec_ok = (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)A) % 32 == 0) && (((uintptr_t)B) % 32 == 0);
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
    unsigned long long *dbg;   // development aid: ns spent by block 0 in [matvec, barrier, dots, barrier, update, barrier, top, -]
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

__device__ __forceinline__ unsigned long long lz_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define LZ_TICK(slot) do { if (P.dbg && b == 0 && tid == 0) { const unsigned long long n_ = lz_now(); P.dbg[slot] += n_ - tmark; tmark = n_; } } while (0)

__global__ void __launch_bounds__(LZ_THREADS) lanczos_kernel(const LanczosArgs *__restrict__ problems) {
    // blockIdx.y = problem (independent matrices advance side by side, each with its own barrier counter: their
    // blocks share the SMs and fill each other's barrier and latency gaps); blockIdx.x = block within the problem
    const LanczosArgs P = problems[blockIdx.y];
    if (P.j_begin >= P.j_end) return;
    extern __shared__ double lsm[];
    __shared__ double sh[LZ_WARPS];
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = gridDim.x, b = blockIdx.x;
    const int64_t N = P.N, ldq = P.ldq, ldg = P.ldg;
    const int64_t nslab = (N + 31) / 32;
    const int WT = nb * LZ_WARPS, gw = b * LZ_WARPS + warp;
    double *red = lsm;                                  // 2 x LZ_WARPS x 32
    double *hs = red + 2 * LZ_WARPS * 32;                   // j_end + 1 coefficients
    double *ws = hs + (P.j_end + 1);                    // N entries of the current vector (when it fits)
    unsigned target = 0;
    unsigned long long tmark = (P.dbg && b == 0 && tid == 0) ? lz_now() : 0ull;
    if (tid == 0) s_abort = 0;
    __syncthreads();

    // || cand ||^2 from the slab partials, in slab order
    auto cand_norm2 = [&]() {
        double v = 0.0;
        for (int64_t s = tid; s < nslab; s += LZ_THREADS) v += __ldcg(&P.nrm[s]);
        return lz_block_sum(v, sh);
    };
    // Orthogonalisation of `buf` against Q[:, 0 .. cnt-1]; leaves the slab norms of the result.
    //  * jloc >= 0 (a Lanczos step, buf = G q_jloc): the two LARGE components, alpha = q_j'w and q_{j-1}'w ~ beta_{j-1},
    //    are removed first, one after the other, by every block on its own shared-memory copy of the vector (no
    //    barrier: the sums have a fixed order, all copies stay identical); `alpha` returns the first.  What is left
    //    has norm ~ beta_j and only rounding-level components along Q, so ONE classical Gram-Schmidt pass restores
    //    orthogonality to the eps level.  (Classical Gram-Schmidt on G q_j itself would multiply the existing loss of
    //    orthogonality by alpha / beta_j > 1 in every step.)
    //  * a second full pass follows when the first one still removed a large part of the vector
    //    (||after|| < 0.5 ||before||: Daniel-Gragg-Kaufman-Stewart), always for jloc < 0 (restart vectors) and when
    //    the vector does not fit in shared memory.
    // `passes` returns how many full passes ran (every block takes the same decision from the same numbers).
    auto reorth = [&](double *buf, int cnt, int jloc, double &alpha, int &passes) -> bool {
        double before2 = 0.0;
        passes = 0;
        alpha = 0.0;
        const bool local = jloc >= 0 && P.w_in_smem;
        for (int pass = 0; pass < 2; ++pass) {
            double *hp = P.h + (size_t)pass * P.hstride;
            if (P.w_in_smem) {
                for (int64_t r = tid; r < N; r += LZ_THREADS) ws[r] = __ldcg(&buf[r]);
                __syncthreads();
            }
            const double *wv = P.w_in_smem ? ws : buf;
            if (pass == 0) {
                if (local) {
                    const double *qj = P.Q + (size_t)jloc * ldq;
                    double v = 0.0;
                    for (int64_t r = tid; r < N; r += LZ_THREADS) v += __ldcg(&qj[r]) * ws[r];
                    alpha = lz_block_sum(v, sh);
                    for (int64_t r = tid; r < N; r += LZ_THREADS) ws[r] -= alpha * __ldcg(&qj[r]);
                    if (jloc > 0) {
                        const double *qp = qj - ldq;
                        v = 0.0;
                        for (int64_t r = tid; r < N; r += LZ_THREADS) v += __ldcg(&qp[r]) * ws[r];
                        const double c = lz_block_sum(v, sh);
                        for (int64_t r = tid; r < N; r += LZ_THREADS) ws[r] -= c * __ldcg(&qp[r]);
                    }
                }
                double v = 0.0;
                for (int64_t r = tid; r < N; r += LZ_THREADS) { const double x = P.w_in_smem ? wv[r] : __ldcg(&wv[r]); v += x * x; }
                before2 = lz_block_sum(v, sh);             // (also orders the updates of ws before the reads below)
            }
            for (int i = gw; i < cnt; i += WT) {
                const double *col = P.Q + (size_t)i * ldq;
                double acc = 0.0;
                int64_t r = lane;
                for (; r + 480 < N; r += 512) {          // 16 independent loads in flight per lane
                    double q[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) q[u] = __ldcg(&col[r + 32 * u]);
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc += q[u] * (P.w_in_smem ? wv[r + 32 * u] : __ldcg(&wv[r + 32 * u]));
                }
                {
                    double q[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) q[u] = (r + 32 * u < N) ? __ldcg(&col[r + 32 * u]) : 0.0;
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (r + 32 * u < N) acc += q[u] * (P.w_in_smem ? wv[r + 32 * u] : __ldcg(&wv[r + 32 * u]));
                }
                acc = warp_sum(acc);
                if (lane == 0) hp[i] = acc;
            }
            LZ_TICK(2);
            if (lz_barrier(P.bar, target, nb, &s_abort)) return true;
            LZ_TICK(3);
            for (int i = tid; i < cnt; i += LZ_THREADS) hs[i] = __ldcg(&hp[i]);
            __syncthreads();
            const int i0 = (int)(((int64_t)cnt * warp) / LZ_WARPS), i1 = (int)(((int64_t)cnt * (warp + 1)) / LZ_WARPS);
            // two 32-row slabs per trip: 32 independent loads in flight per lane
            for (int64_t sA = b; sA < nslab; sA += 2 * (int64_t)nb) {
                const int64_t sB = sA + nb;
                const int64_t rA = 32 * sA + lane, rB = 32 * sB + lane;
                const bool okA = rA < N, okB = sB < nslab && rB < N;
                const double *rowA = P.Q + (okA ? rA : 0), *rowB = P.Q + (okB ? rB : 0);
                double accA = 0.0, accB = 0.0;
                for (int i = i0; i < i1; i += 16) {
                    double qa[16], qb[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const bool in = i + u < i1;
                        qa[u] = (in && okA) ? __ldcg(&rowA[(size_t)(i + u) * ldq]) : 0.0;
                        qb[u] = (in && okB) ? __ldcg(&rowB[(size_t)(i + u) * ldq]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (i + u < i1) {
                            const double hv = hs[i + u];
                            accA += qa[u] * hv;
                            accB += qb[u] * hv;
                        }
                    }
                }
                red[warp * 32 + lane] = accA;
                red[LZ_WARPS * 32 + warp * 32 + lane] = accB;
                __syncthreads();
                if (warp < 2) {
                    const int64_t sx = warp == 0 ? sA : sB, r = warp == 0 ? rA : rB;
                    const bool ok = warp == 0 ? okA : okB;
                    if (sx < nslab) {
                        double t = 0.0;
#pragma unroll
                        for (int w = 0; w < LZ_WARPS; ++w) t += red[warp * LZ_WARPS * 32 + w * 32 + lane];
                        double v = 0.0;
                        // (pass 0 of a Lanczos step starts from the locally orthogonalised copy in shared memory)
                        if (ok) { v = ((pass == 0 && local) ? ws[r] : __ldcg(&buf[r])) - t; buf[r] = v; }
                        const double sq = warp_sum(v * v);
                        if (lane == 0) P.nrm[sx] = sq;
                    }
                }
                __syncthreads();
            }
            LZ_TICK(4);
            if (lz_barrier(P.bar, target, nb, &s_abort)) return true;
            LZ_TICK(5);
            passes = pass + 1;
            if (pass == 0 && local) {
                const double after2 = cand_norm2();
                if (after2 >= 0.25 * before2) break;
            }
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
                int np_ = 0;
                double a_ = 0.0;
                if (reorth(cand, j, -1, a_, np_)) goto fail;
                injected = true;
                continue;
            }
            injected = false;
            inject_tries = 0;
            const double inv = 1.0 / beta_prev;
            LZ_TICK(6);
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
                for (; r + 224 < N; r += 256) {          // 16 independent loads in flight per lane
                    double x0[8], x1[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { x0[u] = __ldcg(&g0[r + 32 * u]); x1[u] = __ldcg(&g1[r + 32 * u]); }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double qv = P.w_in_smem ? ws[r + 32 * u] : __ldcg(&cand[r + 32 * u]) * inv;
                        a0 += x0[u] * qv;
                        a1 += x1[u] * qv;
                    }
                }
                {
                    double x0[8], x1[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool in = r + 32 * u < N;
                        x0[u] = in ? __ldcg(&g0[r + 32 * u]) : 0.0;
                        x1[u] = in ? __ldcg(&g1[r + 32 * u]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (r + 32 * u < N) {
                            const double qv = P.w_in_smem ? ws[r + 32 * u] : __ldcg(&cand[r + 32 * u]) * inv;
                            a0 += x0[u] * qv;
                            a1 += x1[u] * qv;
                        }
                    }
                }
                a0 = warp_sum(a0);
                a1 = warp_sum(a1);
                if (lane == 0) { out[c0] = a0; if (two) out[c0 + 1] = a1; }
            }
            LZ_TICK(0);
            if (lz_barrier(P.bar, target, nb, &s_abort)) goto fail;
            LZ_TICK(1);
            int npass = 0;
            double aloc = 0.0;
            if (reorth(out, j + 1, j, aloc, npass)) goto fail;
            const double aj = aloc + __ldcg(&P.h[j]) + (npass > 1 ? __ldcg(&P.h[P.hstride + j]) : 0.0);
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

// Top-k eigenpairs of the tridiagonal matrices of a batch (cooperative launch; blockIdx.y = problem, block x owns the
// eigenpairs x, x + gridDim.x, ... of its problem):
//   * eigenvalue j by 256-way multisection on the Sturm count (7 rounds instead of 55 bisection steps),
//   * inverse iteration: the block's thread 0 runs the pivoted tridiagonal solve in SHARED memory (or in the global
//     workspace when 7 N doubles do not fit), then block 0 re-orthogonalises clustered vectors (dstein's rule:
//     eigenvalues closer than 1e-3 ||T||) by modified Gram-Schmidt in eigenvalue order.
// work: gridDim.x * 5 * N doubles (only used by the global fall-back).  Z: N x k (ld = N).  lam: k doubles (scratch).
// resid: resid[j] = |last component of eigenvector j| -- times beta_J it is the Lanczos residual norm.
struct TopkArgs {
    const double *d, *e;
    int64_t N;
    int k, use_smem;
    double *evals, *Z, *work, *lam, *resid;
};

__global__ void __launch_bounds__(TRI_THREADS) tridiag_topk_kernel(const TopkArgs *__restrict__ problems) {
    const TopkArgs T = problems[blockIdx.y];
    const double *d = T.d, *e = T.e;
    const int64_t N = T.N;
    const int k = T.k, use_smem = T.use_smem;
    double *evals = T.evals, *Z = T.Z, *work = T.work, *lam = T.lam, *resid = T.resid;
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

// per problem: max_j beta * resid[j] against tol * max(|evals[0]|, |evals[k-1]|)  ->  flag[0] = 1 when every pair has
// converged; flag[3] = the Lanczos kernel's status word
struct CheckArgs {
    const double *resid, *evals, *beta_last, *state;
    int k;
};
__global__ void lanczos_check_kernel(const CheckArgs *__restrict__ problems, int np, double tol, double *flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const CheckArgs C = problems[p];
    const double b = fabs(C.beta_last[0]);
    double mx = 0.0;
    for (int j = 0; j < C.k; ++j) mx = fmax(mx, b * C.resid[j]);
    const double scale = fmax(fabs(C.evals[0]), fabs(C.evals[C.k - 1]));
    flags[4 * p + 0] = (mx <= tol * scale) ? 1.0 : 0.0;
    flags[4 * p + 1] = mx;
    flags[4 * p + 2] = scale;
    flags[4 * p + 3] = C.state[3];
}

namespace {
struct EigState {              // host-side bookkeeping of one problem of a batch
    EigProblem pr;
    int64_t J = 0, Jcap = 0, done = 0, ldq = 0, nslab = 0;
    double *Q = nullptr, *S = nullptr, *sc = nullptr, *twork = nullptr;
    double *alpha, *beta, *wa, *wb, *h, *nrm, *state, *resid, *lam;
    unsigned *bar;
    bool finished = false;
};
}  // namespace

static int sym_eig_group(mfb_context *ctx, int np, const EigProblem *probs, int lz_limit, int topk_limit);

int sym_eig_topk_batch(mfb_context *ctx, int np, const EigProblem *probs) {
    MFB_ARG(np >= 1 && probs, "sym_eig_topk_batch: empty batch");
    for (int p = 0; p < np; ++p)
        MFB_ARG(probs[p].N >= 1 && probs[p].k >= 1 && probs[p].k <= probs[p].N && probs[p].ldg >= probs[p].N,
                "sym_eig_topk: need 1 <= k <= N <= ldg");
    MFB_CUDA(cudaFuncSetAttribute(lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MFB_CUDA(cudaFuncSetAttribute(tridiag_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // co-residency: every block of a cooperative launch must fit on the device at once.  Groups of at most
    // limit / 8 problems (at least 8 blocks each); the pinned mailbox holds the argument blocks of 64 problems.
    int occ = 0;
    MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lanczos_kernel, LZ_THREADS, 48 * 1024));
    if (occ < 1) occ = 1;
    if (occ > 2) occ = 2;
    const int lz_limit = occ * ctx->sm_count;
    int occ2 = 0;
    MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, tridiag_topk_kernel, TRI_THREADS, 64 * 1024));
    if (occ2 < 1) occ2 = 1;
    if (occ2 > 3) occ2 = 3;
    const int topk_limit = occ2 * ctx->sm_count;
    int gmax = std::min(64, std::max(1, lz_limit / 8));
    if (const char *e = getenv("MFB_LZ_GROUP")) { const int g = atoi(e); if (g >= 1 && g < gmax) gmax = g; }
    for (int p0 = 0; p0 < np; p0 += gmax)
        MFB_TRY_RC(sym_eig_group(ctx, std::min(gmax, np - p0), probs + p0, lz_limit, topk_limit));
    return MFB_OK;
}

static int sym_eig_group(mfb_context *ctx, int np, const EigProblem *probs, int lz_limit, int topk_limit) {
    cudaStream_t st = ctx->stream;
    DevBuf buf(ctx);
    std::vector<EigState> E(np);
    LanczosArgs *d_largs = nullptr;
    TopkArgs *d_targs = nullptr;
    CheckArgs *d_cargs = nullptr;
    double *d_flags = nullptr;
    MFB_TRY_RC(buf.alloc(&d_largs, (size_t)np));
    MFB_TRY_RC(buf.alloc(&d_targs, (size_t)np));
    MFB_TRY_RC(buf.alloc(&d_cargs, (size_t)np));
    MFB_TRY_RC(buf.alloc(&d_flags, (size_t)4 * np));
    unsigned long long *d_dbg = nullptr;
    if (getenv("MFB_LZ_PHASES")) {
        MFB_TRY_RC(buf.alloc(&d_dbg, (size_t)8));
        MFB_CUDA(cudaMemsetAsync(d_dbg, 0, 8 * sizeof(unsigned long long), st));
    }
    // pinned mailbox: flags | Lanczos args | top-k args | check args
    double *hflags = (double *)ctx->pinned;
    LanczosArgs *h_largs = (LanczosArgs *)((char *)ctx->pinned + 4096);
    TopkArgs *h_targs = (TopkArgs *)((char *)h_largs + sizeof(LanczosArgs) * 64);
    CheckArgs *h_cargs = (CheckArgs *)((char *)h_targs + sizeof(TopkArgs) * 64);
    const double tol = 2e-13;
    int unfinished = 0;
    for (int p = 0; p < np; ++p) {
        EigState &e = E[p];
        e.pr = probs[p];
        const int64_t N = e.pr.N;
        const int k = e.pr.k;
        if (N == 1) {
            MFB_CUDA(cudaMemcpyAsync(e.pr.evals, e.pr.G, sizeof(double), cudaMemcpyDeviceToDevice, st));
            static const double one = 1.0;
            MFB_CUDA(cudaMemcpyAsync(e.pr.evecs, &one, sizeof(double), cudaMemcpyHostToDevice, st));
            e.finished = true;
            continue;
        }
        // steps: ~10 k on the spectra of the path (see the header); everything when k is a large part of N
        int64_t J = 10 * (int64_t)k + 32;
        if (const char *env = getenv("MFB_LZ_FIRST")) J = atoll(env) * k + 32;
        if (J > N || 3 * (int64_t)k >= N) J = N;
        e.J = J;
        e.Jcap = std::min<int64_t>(N, J + J / 2 + 64);
        e.ldq = round_up(N, 16);
        e.nslab = (N + 31) / 32;
        MFB_TRY_RC(buf.alloc(&e.Q, (size_t)e.ldq * (e.Jcap + 1)));
        MFB_TRY_RC(buf.alloc(&e.S, (size_t)N * k));                 // J x k, J <= N
        // scalars / small vectors: alpha, beta, wa, wb, h (2 passes), nrm, state, resid, lam, barrier counter
        const size_t n_sc = (size_t)4 * N + 2 * (N + 1) + e.nslab + 8 + 2 * N + 8;
        MFB_TRY_RC(buf.alloc(&e.sc, n_sc));
        e.alpha = e.sc; e.beta = e.alpha + N; e.wa = e.beta + N; e.wb = e.wa + N; e.h = e.wb + N;
        e.nrm = e.h + 2 * (N + 1); e.state = e.nrm + e.nslab; e.resid = e.state + 8; e.lam = e.resid + N;
        e.bar = (unsigned *)(e.lam + N);
        MFB_CUDA(cudaMemsetAsync(e.state, 0, 8 * sizeof(double), st));
        ++unfinished;
    }
    while (unfinished > 0) {
        // ---- a chunk of Lanczos steps for every unfinished problem: done -> J ----
        int na = 0;
        size_t smem = 0;
        int64_t want = 1;
        std::vector<int> act;
        for (int p = 0; p < np; ++p) {
            EigState &e = E[p];
            if (e.finished) continue;
            const int64_t N = e.pr.N;
            LanczosArgs &P = h_largs[na];
            P.G = e.pr.G; P.ldg = e.pr.ldg; P.N = N; P.Q = e.Q; P.ldq = e.ldq; P.alpha = e.alpha; P.beta = e.beta;
            P.wa = e.wa; P.wb = e.wb; P.h = e.h; P.hstride = N + 1; P.nrm = e.nrm; P.state = e.state; P.bar = e.bar;
            P.j_begin = (int)e.done; P.j_end = (int)e.J;
            P.dbg = (na == 0) ? d_dbg : nullptr;
            size_t sm = (size_t)(2 * LZ_WARPS * 32 + (e.J + 1) + N) * sizeof(double);
            P.w_in_smem = 1;
            if (sm > 200 * 1024) { P.w_in_smem = 0; sm = (size_t)(2 * LZ_WARPS * 32 + (e.J + 1)) * sizeof(double); }
            MFB_ARG(sm <= 200 * 1024, "sym_eig_topk: too many Lanczos steps for the shared-memory coefficient buffer");
            smem = std::max(smem, sm);
            want = std::max<int64_t>(want, (N / 2 + LZ_WARPS - 1) / LZ_WARPS);   // a column pair per warp
            MFB_CUDA(cudaMemsetAsync(e.bar, 0, sizeof(unsigned), st));
            act.push_back(p);
            ++na;
        }
        int occ = 0;
        MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lanczos_kernel, LZ_THREADS, smem));
        MFB_ARG(occ >= 1, "sym_eig_topk: the Lanczos kernel does not fit on this device");
        if (occ > 2) occ = 2;
        int limit = std::min(lz_limit, occ * ctx->sm_count);
        if (ctx->concurrency > 1) limit = std::max(na, limit / ctx->concurrency);   // other contexts share the device
        int nbc = (int)std::min<int64_t>(want, limit / na);
        if (const char *env = getenv("MFB_LZ_BLOCKS")) { const int c = atoi(env); if (c >= 1) nbc = std::min(c, limit / na); }
        MFB_ARG(nbc >= 1, "sym_eig_topk: batch too large for a cooperative launch");
        MFB_CUDA(cudaMemcpyAsync(d_largs, h_largs, sizeof(LanczosArgs) * na, cudaMemcpyHostToDevice, st));
        {
            const LanczosArgs *argp = d_largs;
            void *args[] = {(void *)&argp};
            MFB_CUDA(cudaLaunchCooperativeKernel((void *)lanczos_kernel, dim3(nbc, na), dim3(LZ_THREADS), args, smem, st));
        }
        // ---- Ritz pairs of every T_J ----
        size_t tsm = 0;
        int kmax = 1;
        bool all_smem = true;
        for (int a = 0; a < na; ++a) {
            EigState &e = E[act[a]];
            e.done = e.J;
            const size_t need = (size_t)7 * e.J * sizeof(double);
            if (need > 200 * 1024) all_smem = false;
            tsm = std::max(tsm, need);
            kmax = std::max(kmax, e.pr.k);
        }
        if (!all_smem) tsm = 0;
        int occ2 = 0;
        MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, tridiag_topk_kernel, TRI_THREADS, tsm));
        if (occ2 < 1) occ2 = 1;
        if (occ2 > 3) occ2 = 3;
        int tlimit = std::min(topk_limit, occ2 * ctx->sm_count);
        if (ctx->concurrency > 1) tlimit = std::max(na, tlimit / ctx->concurrency);
        const int kb = std::max(1, std::min(kmax, tlimit / na));
        for (int a = 0; a < na; ++a) {
            EigState &e = E[act[a]];
            if (!all_smem && !e.twork) MFB_TRY_RC(buf.alloc(&e.twork, (size_t)kb * 5 * e.pr.N));
            TopkArgs &T = h_targs[a];
            T.d = e.alpha; T.e = e.beta; T.N = e.J; T.k = e.pr.k; T.use_smem = all_smem ? 1 : 0;
            T.evals = e.pr.evals; T.Z = e.S; T.work = e.twork; T.lam = e.lam; T.resid = e.resid;
            CheckArgs &C = h_cargs[a];
            C.resid = e.resid; C.evals = e.pr.evals; C.beta_last = e.beta + (e.J - 1); C.state = e.state; C.k = e.pr.k;
        }
        MFB_CUDA(cudaMemcpyAsync(d_targs, h_targs, sizeof(TopkArgs) * na, cudaMemcpyHostToDevice, st));
        MFB_CUDA(cudaMemcpyAsync(d_cargs, h_cargs, sizeof(CheckArgs) * na, cudaMemcpyHostToDevice, st));
        {
            const TopkArgs *argp = d_targs;
            void *targs[] = {(void *)&argp};
            MFB_CUDA(cudaLaunchCooperativeKernel((void *)tridiag_topk_kernel, dim3(kb, na), dim3(TRI_THREADS), targs, tsm, st));
        }
        lanczos_check_kernel<<<(na + 31) / 32, 32, 0, st>>>(d_cargs, na, tol, d_flags);
        MFB_CUDA(cudaGetLastError());
        MFB_CUDA(cudaMemcpyAsync(hflags, d_flags, 4 * na * sizeof(double), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        if (d_dbg) {
            unsigned long long hd[8];
            cudaMemcpy(hd, d_dbg, sizeof(hd), cudaMemcpyDeviceToHost);
            fprintf(stderr, "lanczos phases of problem 0, block 0 (us): matvec %.0f | bar %.0f | dots %.0f | bar %.0f | update %.0f | bar %.0f | top %.0f\n",
                    hd[0] * 1e-3, hd[1] * 1e-3, hd[2] * 1e-3, hd[3] * 1e-3, hd[4] * 1e-3, hd[5] * 1e-3, hd[6] * 1e-3);
        }
        if (getenv("MFB_LZ_VERBOSE"))
            fprintf(stderr, "sym_eig round: %d problems, %d blocks each, top-k %d blocks each, J=%lld, max resid/scale %.2e\n",
                    na, nbc, kb, (long long)E[act[0]].J, hflags[1] / hflags[2]);
        for (int a = 0; a < na; ++a) {
            EigState &e = E[act[a]];
            const double *f = hflags + 4 * a;
            if (f[3] == 1.0) { mfb_set_error("sym_eig_topk: grid barrier timed out in the Lanczos kernel"); return MFB_ERR_CUDA; }
            if (e.J >= e.pr.N || f[0] == 1.0 || f[3] == 2.0) {       // complete tridiagonalisation is exact
                e.finished = true;
                --unfinished;
                continue;
            }
            // not there yet: 25 % more steps (the basis grows; copy it when the first allocation is too small)
            const int64_t Jnext = std::min<int64_t>(e.pr.N, e.J + std::max<int64_t>(e.J / 4, 16));
            if (Jnext > e.Jcap) {
                double *Q2 = nullptr;
                const int64_t cap2 = std::min<int64_t>(e.pr.N, std::max<int64_t>(2 * e.Jcap, Jnext));
                MFB_TRY_RC(buf.alloc(&Q2, (size_t)e.ldq * (cap2 + 1)));
                MFB_CUDA(cudaMemcpyAsync(Q2, e.Q, (size_t)e.ldq * (e.Jcap + 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
                e.Q = Q2; e.Jcap = cap2;
            }
            e.J = Jnext;
        }
    }
    // Ritz vectors: evecs (N x k, ld = N) = Q_J S
    for (int p = 0; p < np; ++p) {
        EigState &e = E[p];
        if (e.pr.N == 1) continue;
        MFB_TRY_RC(tall_times_small(ctx, e.Q, e.ldq, e.S, e.done, e.pr.evecs, e.pr.N, e.pr.N, e.done, e.pr.k, nullptr));
        ctx->last_lanczos_steps = (int)e.done;
        if (getenv("MFB_LZ_VERBOSE"))
            fprintf(stderr, "sym_eig_topk: N=%lld k=%d steps=%lld\n", (long long)e.pr.N, e.pr.k, (long long)e.done);
    }
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

int sym_eig_topk(mfb_context *ctx, const double *G, int64_t ldg, int64_t N, int k, double *evals, double *evecs) {
    EigProblem pr;
    pr.G = G; pr.ldg = ldg; pr.N = N; pr.k = k; pr.evals = evals; pr.evecs = evecs;
    return sym_eig_topk_batch(ctx, 1, &pr);
}
