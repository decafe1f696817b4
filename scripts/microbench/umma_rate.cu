// How long does one tcgen05.mma kind::tf32 (A operand in TMEM, M = 128, K = 8) take as a function of N, and do
// consecutive MMAs overlap when they accumulate into the same / into different TMEM columns?
// One CTA, one issuing thread, REPS MMAs back to back, one commit at the end; clock64 around issue .. mbarrier wait.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int reps, long long *out, int nw = 1) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(sm + 65536 + 16384);
    uint32_t *slot = (uint32_t *)(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 16384) / 4; i += 128) ((float *)sm)[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = *slot;
    if (tid == 0 && mode != 4) {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sm)), adesc = umma_desc_sw128(smem_u32(sm + 65536));
        // mode 0: TS, all MMAs accumulate into the same D; 1: TS, D alternates between two column blocks;
        // 2: TS, fresh D (accumulate = 0) every time, same columns; 3: SS (A from shared memory), same D
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tb + ((mode == 1 && (r & 1)) ? 256u : 0u);
            if (mode == 3) umma_ss(d, adesc, bdesc, idesc, r ? 1u : 0u);
            else umma_ts(d, tb + 480, bdesc, idesc, (mode == 2) ? 0u : (r > 1 ? 1u : 0u));
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    if (mode == 4 && (tid & 31) == 0 && warp < nw) {
        __shared__ uint64_t wb[4];
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wb[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t idesc = idesc_tf32(128, N);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sm));
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) umma_ts(tb + 64 * warp, tb + 480, bdesc, idesc, r ? 1u : 0u);
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&wb[warp])) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&wb[warp])) : "memory");
        const long long t2 = clock64();
        out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    long long *d; CK(cudaMalloc(&d, 64));
    const size_t smem = 65536 + 16384 + 1024 + 64;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char *names[4] = {"TS same D", "TS two D blocks", "TS fresh D", "SS same D"};
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {16, 256}) {
            if (mode == 1 && N > 128) continue;
            for (int reps : {1, 16, 64}) {
                rate_kernel<<<1, 128, smem>>>(N, mode, reps, d);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                printf("%-16s N=%3d reps=%2d: issue %6lld clk, issue..complete %6lld clk  (%.1f clk per MMA)\n", names[mode], N, reps, h[0], h[1], (double)h[1] / reps);
            }
        }
    for (int nw : {1, 2, 4}) {
        rate_kernel<<<1, 128, smem>>>(16, 4, 64, d, nw);
        CK(cudaDeviceSynchronize());
        long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
        printf("%d issuing warps, 64 MMAs each (N=16): per-warp issue..complete", nw);
        for (int w = 0; w < nw; ++w) printf(" %lld", h[2 * w + 1]);
        printf(" clk\n");
    }
    return 0;
}
