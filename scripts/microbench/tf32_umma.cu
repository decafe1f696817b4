// Bring-up test for the TF32 x 3 tensor-core contraction of the FP32 ALS path (tcgen05.mma kind::tf32, A operand
// in TMEM, B operand in shared memory, FP32 accumulators in TMEM).
//   D[128 x N] = sum over K of A[128 x K] * B[N x K]'      (A = 128 "units" of a tile, K = reduction, N = rank)
// A is split per element into hi = top 19 bits (TF32) and lo = a - hi by the threads that move it from (here: global;
// in the library: the TMA ring) into TMEM; B comes pre-split as two FP32 arrays.  D += Ahi*Bhi + Ahi*Blo + Alo*Bhi.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_umma tf32_umma.cu ; run: ./tf32_umma
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ int g_abort = 0;
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {      // bounded: never hang the device
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*(volatile int *)&g_abort) return false;
        if (clock64() - t0 > 2000000000ll) { g_abort = 1; return false; }
    }
    return true;
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B (rows of 128 bytes, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset = 8 rows x 128 B, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// instruction descriptor: kind::tf32, FP32 accumulate, A and B K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

constexpr int N_RANK = 16, KSTAGE = 32, TMEM_COLS = 256;

// A: [128][Ktot] floats (unit-major, K contiguous); Bhi / Blo: [N_RANK][Ktot]; D out: [N_RANK][128] doubles
__global__ void __launch_bounds__(192, 1) tf32_umma_kernel(const float *A, const float *Bhi, const float *Blo, int Ktot,
                                                           double *D, int *status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // B tiles: 2 stages x (hi 2 KB + lo 2 KB), each 16 rows x 128 B, swizzle-128B pattern
    unsigned char *sB = sm;
    uint64_t *bars = (uint64_t *)(sm + 2 * 4096);       // [0..1] a_full, [2..3] a_empty, [4] acc_full
    uint32_t *tmem_base_slot = (uint32_t *)(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], 128); mbar_init(&bars[1], 128);
        mbar_init(&bars[2], 1);   mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {          // one warp allocates (and later frees) the tensor memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_base_slot;
    // TMEM map: columns [0,16) accumulator; A stage buffer s at columns 64 + 64 s: hi [0,32), lo [32,64)
    const uint32_t tD = tbase, tA0 = tbase + 64;
    const int nstage = Ktot / KSTAGE;

    if (warp < 4) {
        // ===== converters: thread = TMEM lane = unit of the tile =====
        const uint32_t lanebase = (uint32_t)(32 * warp) << 16;
        for (int s = 0; s < nstage; ++s) {
            const int buf = s & 1;
            if (s >= 2 && !mbar_wait(&bars[2 + buf], ((s >> 1) - 1) & 1)) break;     // MMAs of stage s - 2 have retired
            // the B tiles of this stage (generic-proxy writes into the swizzled layout, then a proxy fence)
            for (int e = tid; e < 2 * N_RANK * 8; e += 128) {       // 16-byte chunks: [hi|lo][row][chunk]
                const int which = e / (N_RANK * 8), r = (e / 8) % N_RANK, c = e % 8;
                const float *src = (which ? Blo : Bhi) + (size_t)r * Ktot + s * KSTAGE + 4 * c;
                float4 v = *(const float4 *)src;
                *(float4 *)(sB + buf * 4096 + which * 2048 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // this thread's 32 values of A, split, into TMEM
            uint32_t hi[32], lo[32];
            const float *arow = A + (size_t)tid * Ktot + s * KSTAGE;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = *(const float4 *)(arow + 4 * q);
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t h = __float_as_uint(x[u]) & 0xffffe000u;
                    hi[4 * q + u] = h;
                    lo[4 * q + u] = __float_as_uint(x[u] - __uint_as_float(h));
                }
            }
            tmem_st32(lanebase | (tA0 + 64 * buf), hi);
            tmem_st32(lanebase | (tA0 + 64 * buf + 32), lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bars[buf]);
        }
        // ===== epilogue: accumulator -> registers -> global =====
        if (mbar_wait(&bars[4], 0)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[16];
            tmem_ld16(lanebase | tD, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < N_RANK; ++j) D[(size_t)j * 128 + tid] = (double)__uint_as_float(v[j]);
        } else if (tid == 0) {
            *status = 2;
        }
    } else if (warp == 5 && lane == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_tf32(128, N_RANK);
        for (int s = 0; s < nstage; ++s) {
            const int buf = s & 1;
            if (!mbar_wait(&bars[buf], (s >> 1) & 1)) { *status = 1; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bhi = smem_u32(sB + buf * 4096), blo = bhi + 2048;
#pragma unroll
            for (int ks = 0; ks < KSTAGE / 8; ++ks) {
                const uint64_t dhi = umma_desc_sw128(bhi + 32 * ks), dlo = umma_desc_sw128(blo + 32 * ks);
                const uint32_t ahi = tA0 + 64 * buf + 8 * ks, alo = ahi + 32;
                umma_tf32_ts(tD, ahi, dhi, idesc, (s | ks) ? 1u : 0u);
                umma_tf32_ts(tD, ahi, dlo, idesc, 1u);
                umma_tf32_ts(tD, alo, dhi, idesc, 1u);
            }
            umma_commit(&bars[2 + buf]);                 // frees the A buffer and the B tiles of this stage
        }
        umma_commit(&bars[4]);                           // accumulator complete
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(TMEM_COLS) : "memory");
    }
}

int main() {
    const int Ktot = 512;
    std::vector<float> A((size_t)128 * Ktot), Bhi((size_t)N_RANK * Ktot), Blo((size_t)N_RANK * Ktot);
    std::vector<double> Bfull((size_t)N_RANK * Ktot);
    srand(7);
    for (auto &x : A) x = (float)rand() / RAND_MAX * 8.0f;                       // like A >= 0 after pseudovalues
    for (size_t i = 0; i < Bfull.size(); ++i) {
        const double w = (double)rand() / RAND_MAX * 0.02;
        Bfull[i] = w;
        float f = (float)w;
        uint32_t u; memcpy(&u, &f, 4); u &= 0xffffe000u; float h; memcpy(&h, &u, 4);
        Bhi[i] = h;
        Blo[i] = (float)(w - (double)h);
    }
    float *dA, *dBhi, *dBlo; double *dD; int *dS;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBhi, Bhi.size() * 4)); CK(cudaMalloc(&dBlo, Blo.size() * 4));
    CK(cudaMalloc(&dD, (size_t)N_RANK * 128 * 8)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBhi, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBlo, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)N_RANK * 128 * 8)); CK(cudaMemset(dS, 0, 4));
    const size_t smem = 2 * 4096 + 256 + 1024;
    CK(cudaFuncSetAttribute(tf32_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tf32_umma_kernel<<<1, 192, smem>>>(dA, dBhi, dBlo, Ktot, dD, dS);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<double> D((size_t)N_RANK * 128);
    int st = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    double maxrel = 0.0, maxrel_plain = 0.0;
    for (int j = 0; j < N_RANK; ++j)
        for (int c = 0; c < 128; ++c) {
            double ref = 0.0, plain = 0.0;
            for (int k = 0; k < Ktot; ++k) {
                ref += (double)A[(size_t)c * Ktot + k] * Bfull[(size_t)j * Ktot + k];
                float a = A[(size_t)c * Ktot + k]; uint32_t u; memcpy(&u, &a, 4); u &= 0xffffe000u; float ah; memcpy(&ah, &u, 4);
                plain += (double)ah * (double)Bhi[(size_t)j * Ktot + k];             // what a single TF32 product would give
            }
            maxrel = fmax(maxrel, fabs(D[(size_t)j * 128 + c] - ref) / fabs(ref));
            maxrel_plain = fmax(maxrel_plain, fabs(plain - ref) / fabs(ref));
        }
    printf("status %d; TF32x3 max rel err vs FP64 %.3e (single TF32 product would be %.3e); D[0][0] = %.9g\n", st, maxrel,
           maxrel_plain, D[0]);
    printf(maxrel < 2e-6 ? "TF32_UMMA_OK\n" : "TF32_UMMA_MISMATCH\n");
    return 0;
}
