// Shared host/device helpers for libmfbiclust_cuda.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mfbiclust.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void mfb_set_error(const char *fmt, ...);

#define MFB_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            mfb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                          __FILE__, __LINE__);                                          \
            return MFB_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define MFB_ARG(cond, msg)                                                              \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            mfb_set_error("bad argument: %s (%s:%d)", msg, __FILE__, __LINE__);         \
            return MFB_ERR_ARG;                                                         \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------------------
// objects behind the opaque handles
// ---------------------------------------------------------------------------
struct mfb_context {
    int device = 0;
    int sm_count = 0;
    int concurrency = 1;          // contexts working on this device at the same time (mfb_context_set_concurrency)
    int last_lanczos_steps = 0;   // steps the last sym_eig_topk call needed (diagnostics)
    cudaStream_t stream = nullptr;
    // auxiliary streams / events: independent small jobs of one call (the per-cell stages of a bcv batch) run side by
    // side and join the main stream again through events; created on first use
    std::vector<cudaStream_t> aux_streams;
    std::vector<cudaEvent_t> aux_events;
    // small pinned mailbox for control blocks / scalars
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // device workspace pool: blocks are recycled by rounded size instead of going back to cudaFree (which
    // synchronises the device and unmaps); safe because every kernel and copy of a context runs on its one stream
    std::multimap<size_t, void *> pool_free;
    std::unordered_map<void *, size_t> pool_live;
    size_t pool_cached_bytes = 0;
};

bool mfb_context_alive(mfb_context *ctx);      // false once mfb_context_destroy has run

// pooled device allocations (context.cu).  mfb_pool_alloc returns MFB_OK / MFB_ERR_ALLOC.
int mfb_pool_alloc(mfb_context *ctx, void **p, size_t bytes);
void mfb_pool_free(mfb_context *ctx, void *p);
void mfb_pool_release(mfb_context *ctx);      // cudaFree every cached block

// RAII bundle of pooled buffers
struct DevBuf {
    mfb_context *ctx;
    std::vector<void *> ptrs;
    explicit DevBuf(mfb_context *c) : ctx(c) {}
    ~DevBuf() { for (void *q : ptrs) mfb_pool_free(ctx, q); }
    template <typename T> int alloc(T **p, size_t count) {
        int rc = mfb_pool_alloc(ctx, (void **)p, (count ? count : 1) * sizeof(T));
        if (rc == MFB_OK) ptrs.push_back((void *)*p);
        return rc;
    }
};

struct mfb_matrix {
    mfb_context *ctx = nullptr;
    double *d = nullptr;      // column-major, leading dimension lda (elements); float data when elem_bytes == 4
    int elem_bytes = 8;
    int64_t m = 0, n = 0, lda = 0;
    bool owned = false;
    bool stats_valid = false;
    double stats[5] = {0, 0, 0, 0, 0};   // min, max, sumsq, nan count, sum
};

static inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

// driver entry point for tensor-map encoding (resolved at run time: the library has no
// link-time dependency on libcuda, so it loads on a machine without a driver and then
// fails loudly at mfb_context_create()).
typedef CUresult (*mfb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                        const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
mfb_encode_tiled_fn mfb_get_encode_tiled();

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// D(8x8) += A(8x4) * B(4x8), FP64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
#endif
