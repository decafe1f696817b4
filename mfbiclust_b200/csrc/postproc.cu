// Bicluster post-processing on the GPU: replaces sizes / overlap / union.biclust, the arithmetic inside filter.biclust
// (R/helperFunctions.R:150-206 filter.biclust, :226-247 overlap, :303-308 sizes, :336-345 union.biclust) on the
// membership matrices threshold() produces (RowxBicluster m x k and BiclusterxCol k x n, R logicals = int32).
//
// Integer / bit work, HBM-bound: every membership column is packed into a bit vector (one ballot per 32 entries), the
// pairwise intersections |rows_i & rows_j| and |cols_i & cols_j| are popcounts of ANDed words, and the union of the
// chosen biclusters is one coalesced pass that writes the m x n 0/1 matrix: element (r, c) is 1 iff the k-bit masks of
// row r and column c share a chosen bicluster.  The greedy selection of filter.biclust (k steps on k numbers) stays on
// the host (backends.filter_biclust, r/R/backends.R).
#include "common.cuh"

#define PP_MAXW 64          // 64-bit words per row / column mask in the union kernel: k <= 4096

__device__ __forceinline__ bool r_true(int32_t v) { return v != 0 && v != INT32_MIN; }      // NA_LOGICAL is not TRUE

// bits[c * W + w] bit b = M(32 w + b, c) for the len x k logical matrix M whose element (e, c) sits at e * se + c * sc
__global__ void __launch_bounds__(256) pp_pack_kernel(const int32_t *__restrict__ M, int64_t len, int k, int64_t se, int64_t sc,
                                                      int64_t W, uint32_t *__restrict__ bits) {
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t q = gw; q < W * k; q += nw) {
        const int64_t c = q / W, w = q - c * W, e = 32 * w + lane;
        const bool on = e < len && r_true(M[e * se + c * sc]);
        const uint32_t word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) bits[c * W + w] = word;
    }
}

// inter[i * k + j] = |set_i & set_j| (a warp per pair; the diagonal is the set size)
__global__ void __launch_bounds__(256) pp_pair_kernel(const uint32_t *__restrict__ bits, int k, int64_t W, long long *__restrict__ inter) {
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t q = gw; q < (int64_t)k * k; q += nw) {
        const int i = (int)(q / k), j = (int)(q - (int64_t)i * k);
        long long s = 0;
        for (int64_t w = lane; w < W; w += 32) s += __popc(bits[i * W + w] & bits[j * W + w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) inter[q] = s;
    }
}

// masks[e * KW + w] bit b = M(e, 64 w + b): the biclusters that contain row (or column) e
__global__ void __launch_bounds__(256) pp_mask_kernel(const int32_t *__restrict__ M, int64_t len, int k, int64_t se, int64_t sc, int KW,
                                                      const int32_t *__restrict__ choose, unsigned long long *__restrict__ masks) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x)
        for (int w = 0; w < KW; ++w) {
            unsigned long long v = 0ull;
            for (int b = 0; b < 64 && 64 * w + b < k; ++b) {
                const int c = 64 * w + b;
                if ((!choose || r_true(choose[c])) && r_true(M[e * se + c * sc])) v |= 1ull << b;
            }
            masks[e * KW + w] = v;
        }
}

// out(r, c) = 1 iff row r and column c share a (chosen) bicluster; column-major m x n doubles
__global__ void __launch_bounds__(256) pp_union_kernel(const unsigned long long *__restrict__ rmask, const unsigned long long *__restrict__ cmask,
                                                       int64_t m, int64_t n, int KW, double *__restrict__ out) {
    for (int64_t c = blockIdx.y; c < n; c += gridDim.y) {
        unsigned long long cm[PP_MAXW];
        bool any = false;
        for (int w = 0; w < KW; ++w) { cm[w] = cmask[c * KW + w]; any |= cm[w] != 0ull; }
        double *col = out + c * m;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
            bool hit = false;
            if (any)
                for (int w = 0; w < KW; ++w) hit |= (rmask[r * KW + w] & cm[w]) != 0ull;
            col[r] = hit ? 1.0 : 0.0;
        }
    }
}

static int pp_grid(mfb_context *ctx, int64_t items, int per_block) {
    int64_t nb = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (nb > cap) nb = cap;
    return (int)(nb < 1 ? 1 : nb);
}

extern "C" {

int mfb_bicluster_overlap(mfb_context *ctx, const int32_t *rowxBicluster, int64_t m, const int32_t *biclusterxCol, int64_t n, int k,
                          double *sizes, double *overlaps) {
    MFB_ARG(ctx && rowxBicluster && biclusterxCol && sizes && overlaps, "NULL argument");
    MFB_ARG(m >= 1 && n >= 1 && k >= 1, "empty membership matrix");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t Wr = (m + 31) / 32, Wc = (n + 31) / 32;
    DevBuf buf(ctx);
    int32_t *dR = nullptr, *dC = nullptr;
    uint32_t *bR = nullptr, *bC = nullptr;
    long long *iR = nullptr, *iC = nullptr;
    {
        int rc = buf.alloc(&dR, (size_t)m * k);
        if (rc == MFB_OK) rc = buf.alloc(&dC, (size_t)n * k);
        if (rc == MFB_OK) rc = buf.alloc(&bR, (size_t)Wr * k);
        if (rc == MFB_OK) rc = buf.alloc(&bC, (size_t)Wc * k);
        if (rc == MFB_OK) rc = buf.alloc(&iR, (size_t)k * k);
        if (rc == MFB_OK) rc = buf.alloc(&iC, (size_t)k * k);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(dR, rowxBicluster, sizeof(int32_t) * m * k, cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(dC, biclusterxCol, sizeof(int32_t) * n * k, cudaMemcpyHostToDevice, st));
    // RowxBicluster(e, c) at e + c * m; BiclusterxCol(c, e) at c + e * k
    pp_pack_kernel<<<pp_grid(ctx, Wr * k, 8), 256, 0, st>>>(dR, m, k, 1, m, Wr, bR);
    pp_pack_kernel<<<pp_grid(ctx, Wc * k, 8), 256, 0, st>>>(dC, n, k, k, 1, Wc, bC);
    pp_pair_kernel<<<pp_grid(ctx, (int64_t)k * k, 8), 256, 0, st>>>(bR, k, Wr, iR);
    pp_pair_kernel<<<pp_grid(ctx, (int64_t)k * k, 8), 256, 0, st>>>(bC, k, Wc, iC);
    MFB_CUDA(cudaGetLastError());
    std::vector<long long> hR((size_t)k * k), hC((size_t)k * k);
    MFB_CUDA(cudaMemcpyAsync(hR.data(), iR, sizeof(long long) * k * k, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(hC.data(), iC, sizeof(long long) * k * k, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    // sizes (:303-308) and overlap(i, j) = |rows_i & rows_j| * |cols_i & cols_j| / size_i (:233-243; 0 / 0 = NaN as in R);
    // overlaps is the k x k matrix of overlap(matrix = TRUE), column-major: entry (i, j) at i + k * j
    for (int i = 0; i < k; ++i) sizes[i] = (double)hR[(size_t)i * k + i] * (double)hC[(size_t)i * k + i];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            overlaps[(size_t)i + (size_t)k * j] = ((double)hR[(size_t)i * k + j] * (double)hC[(size_t)i * k + j]) / sizes[i];
    return MFB_OK;
}

int mfb_bicluster_union(mfb_context *ctx, const int32_t *rowxBicluster, int64_t m, const int32_t *biclusterxCol, int64_t n, int k,
                        const int32_t *choose, double *biclustered) {
    MFB_ARG(ctx && rowxBicluster && biclusterxCol && biclustered, "NULL argument");
    MFB_ARG(m >= 1 && n >= 1 && k >= 0, "empty membership matrix");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (k == 0) { memset(biclustered, 0, sizeof(double) * (size_t)m * n); return MFB_OK; }
    const int KW = (k + 63) / 64;
    MFB_ARG(KW <= PP_MAXW, "more than 4096 biclusters");
    DevBuf buf(ctx);
    int32_t *dR = nullptr, *dC = nullptr, *dch = nullptr;
    unsigned long long *mR = nullptr, *mC = nullptr;
    double *dout = nullptr;
    {
        int rc = buf.alloc(&dR, (size_t)m * k);
        if (rc == MFB_OK) rc = buf.alloc(&dC, (size_t)n * k);
        if (rc == MFB_OK) rc = buf.alloc(&dch, (size_t)k);
        if (rc == MFB_OK) rc = buf.alloc(&mR, (size_t)m * KW);
        if (rc == MFB_OK) rc = buf.alloc(&mC, (size_t)n * KW);
        if (rc == MFB_OK) rc = buf.alloc(&dout, (size_t)m * n);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(dR, rowxBicluster, sizeof(int32_t) * m * k, cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(dC, biclusterxCol, sizeof(int32_t) * n * k, cudaMemcpyHostToDevice, st));
    if (choose) MFB_CUDA(cudaMemcpyAsync(dch, choose, sizeof(int32_t) * k, cudaMemcpyHostToDevice, st));
    // the choice is folded into the row masks only: a bicluster that is not chosen never matches
    pp_mask_kernel<<<pp_grid(ctx, m, 256), 256, 0, st>>>(dR, m, k, 1, m, KW, choose ? dch : nullptr, mR);
    pp_mask_kernel<<<pp_grid(ctx, n, 256), 256, 0, st>>>(dC, n, k, k, 1, KW, nullptr, mC);
    dim3 grid((unsigned)pp_grid(ctx, m, 256), (unsigned)(n < 65535 ? n : 65535));
    if ((int64_t)grid.x * grid.y > (int64_t)ctx->sm_count * 16) {
        grid.y = (unsigned)(((int64_t)ctx->sm_count * 16 + grid.x - 1) / grid.x);
        if (grid.y < 1) grid.y = 1;
    }
    pp_union_kernel<<<grid, 256, 0, st>>>(mR, mC, m, n, KW, dout);
    MFB_CUDA(cudaGetLastError());
    MFB_CUDA(cudaMemcpyAsync(biclustered, dout, sizeof(double) * m * n, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    return MFB_OK;
}

}  // extern "C"
