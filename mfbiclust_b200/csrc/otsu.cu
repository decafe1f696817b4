// Otsu thresholds and membership masks on the GPU: replaces otsuHelper / threshold
// (R/BiclusterStrategy.R:442-456, 508-528) and EBImage::otsu (call site :448), which bins with
// graphics::hist.default.  Bit-exact with the R arithmetic: every histogram moment is a dyadic
// rational, and the few rounded operations (two divisions, one subtraction, two products, the
// un-rescale multiply and add) are issued with explicit round-to-nearest intrinsics so that no
// FMA contraction can change a bit.
#include "common.cuh"

#define OTSU_LEVELS 256

__device__ __forceinline__ double fuzzy_break(int i) {
    // hist.default: breaks = i/256; diddle = 1e-7 * median(diff(breaks)); right = TRUE, include.lowest = TRUE
    const double diddle = 1e-7 * (1.0 / OTSU_LEVELS);
    return (i == 0) ? (0.0 - diddle) : __dadd_rn((double)i / OTSU_LEVELS, diddle);
}

// one CTA per column of M (rows x cols, column-major)
__global__ void __launch_bounds__(512) otsu_kernel(const double *__restrict__ M, int64_t rows, int64_t ld,
                                                   double *__restrict__ th) {
    __shared__ unsigned int hist[OTSU_LEVELS];
    __shared__ double s_mn[16], s_mx[16];
    const double *x = M + (size_t)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = tid; i < rows; i += blockDim.x) {
        const double v = x[i];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    for (int i = tid; i < OTSU_LEVELS; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    mn = s_mn[0];
    mx = s_mx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, s_mn[w]); mx = fmax(mx, s_mx[w]); }
    if (mx == mn) {                     // constant column: the threshold is that value (:451-453)
        if (tid == 0) th[blockIdx.x] = x[0];
        return;
    }
    const double range = __dsub_rn(mx, mn);
    for (int64_t i = tid; i < rows; i += blockDim.x) {
        const double r = __ddiv_rn(__dsub_rn(x[i], mn), range);     // (x - min) / (max - min)
        int b = (int)floor(r * OTSU_LEVELS);
        b = b < 0 ? 0 : (b > OTSU_LEVELS - 1 ? OTSU_LEVELS - 1 : b);
        while (b > 0 && r <= fuzzy_break(b)) --b;                   // fuzzy[b] < r <= fuzzy[b+1]
        while (b < OTSU_LEVELS - 1 && r > fuzzy_break(b + 1)) ++b;
        atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        // EBImage::otsu: inclusive class weights, var = w1 * w2 * (m2/w2 - m1/w1)^2 = (w1*w2) * (d*d) in R's
        // evaluation order (`^` binds tighter than `*`, x^2 is x*x), threshold = mean of the
        // first and last maximising mid-points.
        double wtot = 0.0, mtot = 0.0;
        for (int i = 0; i < OTSU_LEVELS; ++i) {
            const double c = (double)hist[i];
            wtot += c;
            mtot += c * ((2.0 * i + 1.0) / (2.0 * OTSU_LEVELS));
        }
        double w1 = 0.0, m1 = 0.0, best = -INFINITY;
        int first = -1, last = -1;
        for (int i = 0; i < OTSU_LEVELS; ++i) {
            const double c = (double)hist[i];
            const double cm = c * ((2.0 * i + 1.0) / (2.0 * OTSU_LEVELS));
            w1 += c;
            m1 += cm;
            const double w2 = wtot + c - w1;
            const double m2 = mtot + cm - m1;
            if (w1 == 0.0 || w2 == 0.0) continue;                    // 0/0 -> NaN in R, dropped by na.rm
            const double d = __dsub_rn(__ddiv_rn(m2, w2), __ddiv_rn(m1, w1));
            const double var = __dmul_rn(__dmul_rn(w1, w2), __dmul_rn(d, d));
            if (var > best) { best = var; first = i; last = i; }
            else if (var == best) last = i;
        }
        const double t = __dmul_rn(__dadd_rn((2.0 * first + 1.0) / (2.0 * OTSU_LEVELS),
                                             (2.0 * last + 1.0) / (2.0 * OTSU_LEVELS)), 0.5);
        th[blockIdx.x] = __dadd_rn(__dmul_rn(t, range), mn);        // t * (max - min) + min
    }
}

// margin 2: th per column; margin 1: th per row.  `<` when th < 0, else `>` (strict).
__global__ void __launch_bounds__(256) threshold_kernel(const double *__restrict__ M, int64_t rows, int64_t cols,
                                                        const double *__restrict__ th, int margin,
                                                        int32_t *__restrict__ out) {
    const int64_t total = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / rows, i = e - j * rows;
        const double t = th[margin == 1 ? i : j];
        const double v = M[e];
        out[e] = (t < 0.0) ? (v < t) : (v > t);
    }
}

extern "C" {

int mfb_otsu(mfb_context *ctx, const double *M, int64_t rows, int64_t cols, double *th) {
    MFB_ARG(ctx && M && th, "NULL argument");
    MFB_ARG(rows >= 1 && cols >= 1, "empty matrix");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double *dM = nullptr, *dth = nullptr;
    DevBuf buf(ctx);
    {
        int rc = buf.alloc(&dM, (size_t)rows * cols);
        if (rc == MFB_OK) rc = buf.alloc(&dth, (size_t)cols);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(dM, M, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, st));
    otsu_kernel<<<(unsigned)cols, 512, 0, st>>>(dM, rows, rows, dth);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaMemcpyAsync(th, dth, sizeof(double) * cols, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

int mfb_threshold(mfb_context *ctx, const double *M, int64_t rows, int64_t cols, const double *th, int margin,
                  int32_t *out) {
    MFB_ARG(ctx && M && th && out, "NULL argument");
    MFB_ARG(margin == 1 || margin == 2, "MARGIN must be 1 or 2");
    MFB_ARG(rows >= 1 && cols >= 1, "empty matrix");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nth = (margin == 1) ? rows : cols;
    double *dM = nullptr, *dth = nullptr;
    int32_t *dout = nullptr;
    DevBuf buf(ctx);
    {
        int rc = buf.alloc(&dM, (size_t)rows * cols);
        if (rc == MFB_OK) rc = buf.alloc(&dth, (size_t)nth);
        if (rc == MFB_OK) rc = buf.alloc(&dout, (size_t)rows * cols);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(dM, M, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(dth, th, sizeof(double) * nth, cudaMemcpyHostToDevice, st));
    int64_t nb = (rows * cols + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    threshold_kernel<<<(unsigned)nb, 256, 0, st>>>(dM, rows, cols, dth, margin, dout);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaMemcpyAsync(out, dout, sizeof(int32_t) * rows * cols, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

}  // extern "C"
