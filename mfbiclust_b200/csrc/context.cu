// Context, error string and the HBM-resident matrix handle of libmfbiclust_cuda.so.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mfb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

mfb_encode_tiled_fn mfb_get_encode_tiled() {
    static mfb_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (mfb_encode_tiled_fn)p;
    return fn;
}

// contexts that have not been destroyed: handles that outlive their context (e.g. finalisers run in arbitrary
// order by a garbage collector) must not touch it -- mfb_context_destroy has already released their device memory
static std::vector<mfb_context *> g_live_contexts;
bool mfb_context_alive(mfb_context *ctx) {
    for (mfb_context *c : g_live_contexts)
        if (c == ctx) return true;
    return false;
}

static size_t pool_round(size_t bytes) {
    if (bytes <= 512) return 512;
    if (bytes <= ((size_t)1 << 20)) {
        size_t r = 512;
        while (r < bytes) r <<= 1;
        return r;
    }
    const size_t q = (size_t)2 << 20;
    return (bytes + q - 1) / q * q;
}

int mfb_pool_alloc(mfb_context *ctx, void **p, size_t bytes) {
    const size_t r = pool_round(bytes);
    auto it = ctx->pool_free.find(r);
    if (it != ctx->pool_free.end()) {
        *p = it->second;
        ctx->pool_free.erase(it);
        ctx->pool_cached_bytes -= r;
        ctx->pool_live[*p] = r;
        return MFB_OK;
    }
    cudaError_t e = cudaMalloc(p, r);
    if (e != cudaSuccess) {                    // give the cached blocks back to the driver and try once more
        cudaGetLastError();
        mfb_pool_release(ctx);
        e = cudaMalloc(p, r);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        mfb_set_error("cudaMalloc(%zu) failed: %s", r, cudaGetErrorString(e));
        *p = nullptr;
        return MFB_ERR_ALLOC;
    }
    ctx->pool_live[*p] = r;
    return MFB_OK;
}

void mfb_pool_free(mfb_context *ctx, void *p) {
    if (!p) return;
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) { cudaFree(p); return; }
    ctx->pool_free.emplace(it->second, p);
    ctx->pool_cached_bytes += it->second;
    ctx->pool_live.erase(it);
}

void mfb_pool_release(mfb_context *ctx) {
    if (ctx->pool_free.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pool_free) cudaFree(kv.second);
    ctx->pool_free.clear();
    ctx->pool_cached_bytes = 0;
}

extern "C" {

int mfb_version(void) { return 100; }
const char *mfb_last_error(void) { return g_err; }

int mfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mfb_context_create(int device, mfb_context **out) {
    MFB_ARG(out != nullptr, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        mfb_set_error("no CUDA device available (%s): libmfbiclust_cuda has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MFB_ERR_CUDA;
    }
    MFB_ARG(device >= 0 && device < n, "device index out of range");
    MFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        mfb_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        return MFB_ERR_UNSUPPORTED;
    }
    if (!mfb_get_encode_tiled()) {
        mfb_set_error("cuTensorMapEncodeTiled not available from the driver");
        return MFB_ERR_CUDA;
    }
    mfb_context *c = new mfb_context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    MFB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->pinned_bytes = 1 << 20;
    MFB_CUDA(cudaMallocHost(&c->pinned, c->pinned_bytes));
    g_live_contexts.push_back(c);
    *out = c;
    return MFB_OK;
}

void mfb_context_destroy(mfb_context *ctx) {
    if (!ctx || !mfb_context_alive(ctx)) return;
    for (size_t i = 0; i < g_live_contexts.size(); ++i)
        if (g_live_contexts[i] == ctx) { g_live_contexts.erase(g_live_contexts.begin() + i); break; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    mfb_pool_release(ctx);
    for (auto &kv : ctx->pool_live) cudaFree(kv.first);     // handles the caller forgot to release
    ctx->pool_live.clear();
    for (cudaStream_t a : ctx->aux_streams) cudaStreamDestroy(a);
    for (cudaEvent_t e : ctx->aux_events) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    delete ctx;
}

int mfb_context_sync(mfb_context *ctx) {
    MFB_ARG(ctx, "ctx is NULL");
    MFB_CUDA(cudaSetDevice(ctx->device));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

void *mfb_context_stream(mfb_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int mfb_context_set_concurrency(mfb_context *ctx, int n) {
    MFB_ARG(ctx && n >= 1, "bad argument");
    ctx->concurrency = n;
    return MFB_OK;
}

int mfb_context_trim(mfb_context *ctx) {
    MFB_ARG(ctx, "ctx is NULL");
    MFB_CUDA(cudaSetDevice(ctx->device));
    mfb_pool_release(ctx);
    return MFB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// matrix statistics: min, max, sum of squares, NaN count, sum  (two-stage, deterministic)
// ---------------------------------------------------------------------------
struct Stat5 {
    double mn, mx, ss, nan, sum;
};

template <typename T>
__global__ void __launch_bounds__(256) stats_partial_kernel(const T *__restrict__ A, int64_t m, int64_t n,
                                                            int64_t lda, Stat5 *__restrict__ part) {
    double mn = INFINITY, mx = -INFINITY, ss = 0.0, nn = 0.0, sm = 0.0;
    // one column at a time per block; threads stride over rows (coalesced)
    for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
        const T *col = A + j * lda;
        for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
            double v = (double)col[i];
            if (v != v) {
                nn += 1.0;
            } else {
                mn = fmin(mn, v);
                mx = fmax(mx, v);
                ss += v * v;
                sm += v;
            }
        }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    ss = warp_sum(ss);
    nn = warp_sum(nn);
    sm = warp_sum(sm);
    __shared__ Stat5 sh[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = Stat5{mn, mx, ss, nn, sm};
    __syncthreads();
    if (threadIdx.x == 0) {
        Stat5 r = sh[0];
        for (int i = 1; i < 8; ++i) {
            r.mn = fmin(r.mn, sh[i].mn);
            r.mx = fmax(r.mx, sh[i].mx);
            r.ss += sh[i].ss;
            r.nan += sh[i].nan;
            r.sum += sh[i].sum;
        }
        part[blockIdx.x] = r;
    }
}

__global__ void stats_final_kernel(const Stat5 *__restrict__ part, int nb, double *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        Stat5 r = part[0];
        for (int i = 1; i < nb; ++i) {
            r.mn = fmin(r.mn, part[i].mn);
            r.mx = fmax(r.mx, part[i].mx);
            r.ss += part[i].ss;
            r.nan += part[i].nan;
            r.sum += part[i].sum;
        }
        out[0] = r.mn;
        out[1] = r.mx;
        out[2] = r.ss;
        out[3] = r.nan;
        out[4] = r.sum;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) shift_kernel(T *__restrict__ A, int64_t m, int64_t n, int64_t lda,
                                                    double shift) {
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
        T *col = A + j * lda;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
            col[i] = (T)((double)col[i] + shift);
    }
}

// dst (float, ld = ldd, zero padded rows) <- src (double, ld = lds)
__global__ void __launch_bounds__(256) to_f32_kernel(const double *__restrict__ src, int64_t m, int64_t n, int64_t lds,
                                                     float *__restrict__ dst, int64_t ldd) {
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
        const double *col = src + j * lds;
        float *out = dst + j * ldd;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ldd; i += (int64_t)gridDim.x * blockDim.x)
            out[i] = (i < m) ? (float)col[i] : 0.0f;
    }
}

extern "C" {

int mfb_matrix_upload(mfb_context *ctx, const double *host, int64_t m, int64_t n, mfb_matrix **out) {
    MFB_ARG(ctx && host && out, "NULL argument");
    MFB_ARG(m > 0 && n > 0, "matrix dimensions must be positive");
    MFB_CUDA(cudaSetDevice(ctx->device));
    mfb_matrix *A = new mfb_matrix();
    A->ctx = ctx;
    A->m = m;
    A->n = n;
    A->lda = round_up(m, 16);   // 128-byte aligned columns: TMA strides, 32-byte vector loads
    A->owned = true;
    size_t bytes = (size_t)A->lda * (size_t)n * sizeof(double) + 256;
    if (mfb_pool_alloc(ctx, (void **)&A->d, bytes) != MFB_OK) {
        delete A;
        return MFB_ERR_ALLOC;
    }
    if (A->lda != m) MFB_CUDA(cudaMemsetAsync(A->d, 0, bytes, ctx->stream));
    MFB_CUDA(cudaMemcpy2DAsync(A->d, A->lda * sizeof(double), host, m * sizeof(double), m * sizeof(double), n,
                               cudaMemcpyHostToDevice, ctx->stream));
    *out = A;
    return MFB_OK;
}

int mfb_matrix_upload_f32(mfb_context *ctx, const double *host, int64_t m, int64_t n, mfb_matrix **out) {
    MFB_ARG(ctx && host && out, "NULL argument");
    MFB_ARG(m > 0 && n > 0, "matrix dimensions must be positive");
    MFB_CUDA(cudaSetDevice(ctx->device));
    DevBuf tmp(ctx);
    double *stage = nullptr;
    {
        int rc = tmp.alloc(&stage, (size_t)m * n);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(stage, host, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    mfb_matrix *A = new mfb_matrix();
    A->ctx = ctx;
    A->m = m;
    A->n = n;
    A->elem_bytes = 4;
    A->lda = round_up(m, 32);   // 128-byte aligned columns of floats
    A->owned = true;
    if (mfb_pool_alloc(ctx, (void **)&A->d, (size_t)A->lda * n * sizeof(float) + 256) != MFB_OK) {
        delete A;
        return MFB_ERR_ALLOC;
    }
    int gy = (int)(n < 65535 ? n : 65535), gx = (int)((A->lda + 255) / 256);
    if (gx > 64) gx = 64;
    to_f32_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(stage, m, n, m, (float *)A->d, A->lda);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));     // the staging buffer goes back to the pool
    *out = A;
    return MFB_OK;
}

int mfb_matrix_wrap_dev(mfb_context *ctx, double *dev, int64_t m, int64_t n, mfb_matrix **out) {
    MFB_ARG(ctx && dev && out, "NULL argument");
    MFB_ARG(m > 0 && n > 0, "matrix dimensions must be positive");
    MFB_ARG(m % 16 == 0 && ((uintptr_t)dev % 128) == 0,
            "wrapped device matrices need a row count that is a multiple of 16 and a 128-byte aligned base (TMA boxes)");
    mfb_matrix *A = new mfb_matrix();
    A->ctx = ctx;
    A->d = dev;
    A->m = m;
    A->n = n;
    A->lda = m;
    A->owned = false;
    *out = A;
    return MFB_OK;
}

int mfb_matrix_view(mfb_context *ctx, mfb_matrix *A, mfb_matrix **out) {
    MFB_ARG(ctx && A && out, "NULL argument");
    MFB_ARG(ctx->device == A->ctx->device, "a view must live on the device of the matrix");
    mfb_matrix *V = new mfb_matrix();
    *V = *A;
    V->ctx = ctx;
    V->owned = false;
    *out = V;
    return MFB_OK;
}

void mfb_matrix_free(mfb_matrix *A) {
    if (!A) return;
    if (A->owned && A->d && mfb_context_alive(A->ctx)) {
        cudaSetDevice(A->ctx->device);
        mfb_pool_free(A->ctx, A->d);
    }
    delete A;
}

void *mfb_matrix_dev_ptr(mfb_matrix *A) { return A ? (void *)A->d : nullptr; }

int mfb_matrix_stats(mfb_matrix *A, double stats[5]) {
    MFB_ARG(A && stats, "NULL argument");
    mfb_context *ctx = A->ctx;
    MFB_CUDA(cudaSetDevice(ctx->device));
    if (!A->stats_valid) {
        int nb = (int)(A->n < 4 * ctx->sm_count ? A->n : 4 * ctx->sm_count);
        DevBuf buf(ctx);
        Stat5 *part = nullptr;
        double *dout = nullptr;
        {
            unsigned char *raw = nullptr;
            int rc = buf.alloc(&raw, sizeof(Stat5) * nb + 5 * sizeof(double));
            if (rc != MFB_OK) return rc;
            part = (Stat5 *)raw;
        }
        dout = (double *)(part + nb);
        if (A->elem_bytes == 4)
            stats_partial_kernel<float><<<nb, 256, 0, ctx->stream>>>((const float *)A->d, A->m, A->n, A->lda, part);
        else
            stats_partial_kernel<double><<<nb, 256, 0, ctx->stream>>>(A->d, A->m, A->n, A->lda, part);
        stats_final_kernel<<<1, 32, 0, ctx->stream>>>(part, nb, dout);
        MFB_CUDA(cudaGetLastError());
        MFB_CUDA(cudaMemcpyAsync(A->stats, dout, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        A->stats_valid = true;
    }
    memcpy(stats, A->stats, 5 * sizeof(double));
    return MFB_OK;
}

int mfb_matrix_shift(mfb_matrix *A, double shift) {
    MFB_ARG(A, "NULL argument");
    mfb_context *ctx = A->ctx;
    MFB_CUDA(cudaSetDevice(ctx->device));
    int gy = (int)(A->n < 65535 ? A->n : 65535);
    int gx = (int)((A->m + 255) / 256);
    if (gx > 64) gx = 64;
    if (A->elem_bytes == 4)
        shift_kernel<float><<<dim3(gx, gy), 256, 0, ctx->stream>>>((float *)A->d, A->m, A->n, A->lda, shift);
    else
        shift_kernel<double><<<dim3(gx, gy), 256, 0, ctx->stream>>>(A->d, A->m, A->n, A->lda, shift);
    MFB_CUDA(cudaGetLastError());
    A->stats_valid = false;
    return MFB_OK;
}

}  // extern "C"
