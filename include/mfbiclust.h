/*
 * mfbiclust.h -- C ABI of libmfbiclust_cuda.so (B200 / sm_100a).
 *
 * The reference (jonalim/mfBiclust) is pure R with no native boundary: its
 * back-ends are reached by *function name* through
 *   do.call(get(method), c(list(A = obj, k = k), list(...)))   R/BiclusterStrategy.R:82-87
 * and its arithmetic lives in NMF::.fcnnls, nipals::nipals, stats::prcomp,
 * MASS::ginv, EBImage::otsu.  This header declares the entry points an R
 * `.Call` shim binds so that als_nmf / svd_pca /
 * nipals_pca / bcv / auto_bcv / otsuHelper / threshold keep their R
 * signatures while the arithmetic runs on the GPU.  Each entry point cites the
 * reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes; all matrices COLUMN-MAJOR (R layout), FP64;
 *   - caller-owned HOST buffers unless the name says `_dev`; inputs are never
 *     written;
 *   - return value: 0 = MFB_OK; > 0 = algorithmic condition the R wrapper
 *     reacts to (restart, fallback); < 0 = bad argument / CUDA error, with a
 *     message available from mfb_last_error();
 *   - no global mutable state besides the explicit mfb_context; one host
 *     thread per context;
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     MFB_ERR_CUDA.
 */
#ifndef MFBICLUST_H
#define MFBICLUST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
#define MFB_OK               0
#define MFB_SINGULAR         1  /* Gram matrix not invertible: R's solve() error inside .fcnnls (caught by try(), R/bicluster.R:123,150) */
#define MFB_ZERO_ROW         2  /* any(rowSums(H) == 0), R/bicluster.R:137 */
#define MFB_NOT_CONVERGED    3  /* informational (nipals maxiter, auto_bcv maxIter) */
#define MFB_ERR_ARG         -1
#define MFB_ERR_CUDA        -2
#define MFB_ERR_ALLOC       -3
#define MFB_ERR_NEGATIVE    -4  /* "Negative values are not allowed", R/bicluster.R:76 */
#define MFB_ERR_UNSUPPORTED -5

typedef struct mfb_context mfb_context;
typedef struct mfb_matrix mfb_matrix;   /* an m x n FP64 matrix resident in HBM */
typedef struct mfb_als mfb_als;         /* one als_nmf replicate in flight */

/* ---- context ----------------------------------------------------------- */
int         mfb_version(void);
const char *mfb_last_error(void);
int  mfb_device_count(void);
int  mfb_context_create(int device, mfb_context **out);
void mfb_context_destroy(mfb_context *ctx);
int  mfb_context_sync(mfb_context *ctx);
/* CUDA stream (cudaStream_t) every kernel of this context is launched on; for callers that time with CUDA events. */
void *mfb_context_stream(mfb_context *ctx);
/* Device buffers are recycled through a per-context pool; this returns the cached ones to the driver. */
int  mfb_context_trim(mfb_context *ctx);
/* Hint: `n` contexts (host threads, each with its own context / stream) work on this device at the same time, e.g.
 * on the independent cells of a bcv call.  Barrier-bound cooperative kernels then size their grids to 1/n of the
 * device instead of fighting for every SM.  Results never depend on it.  Default 1. */
int  mfb_context_set_concurrency(mfb_context *ctx, int n);

/* ---- resident matrix ---------------------------------------------------- */
/* Upload a column-major host matrix (the `A` / `Y` argument of every back-end). */
int  mfb_matrix_upload(mfb_context *ctx, const double *host, int64_t m, int64_t n, mfb_matrix **out);
/* Same, but the resident copy is kept in SINGLE precision (half the HBM traffic of the ALS stream kernels; the
 * factors, the tensor-core accumulation and the NNLS stay FP64).  Accepted by the mfb_als_* entry points only. */
int  mfb_matrix_upload_f32(mfb_context *ctx, const double *host, int64_t m, int64_t n, mfb_matrix **out);
/* Wrap an existing device buffer (column-major, ld = m); not owned. */
int  mfb_matrix_wrap_dev(mfb_context *ctx, double *dev, int64_t m, int64_t n, mfb_matrix **out);
/* Non-owning alias of a resident matrix for ANOTHER context on the same device (one host thread per context: this is
 * how several threads, each with its own context and stream, work on one resident matrix, e.g. the independent cells
 * of a bcv call).  The owner must outlive the view and must have finished writing (mfb_context_sync). */
int  mfb_matrix_view(mfb_context *ctx, mfb_matrix *A, mfb_matrix **out);
void mfb_matrix_free(mfb_matrix *A);
/* stats[0..4] = min, max, sum of squares, #NaN/NA, sum  (max(A): R/bicluster.R:81; any(A<0): :76; pseudovalues: R/helperFunctions.R:270-278) */
int  mfb_matrix_stats(mfb_matrix *A, double stats[5]);
/* A <- A + shift  (pseudovalues(), R/helperFunctions.R:275) */
int  mfb_matrix_shift(mfb_matrix *A, double shift);
void *mfb_matrix_dev_ptr(mfb_matrix *A);

/* ---- als_nmf (R/bicluster.R:70-216; NMF::.fcnnls call sites :123, :150) --- */
/*
 * One replicate of the `while (i < maxIter)` loop from a GIVEN start:
 *   W0     m x k, column-normalised   (R/bicluster.R:95-101)
 *   Hold0  k x n                      (R/bicluster.R:104)
 *   resid_old0 = max(A)               (R/bicluster.R:105)
 * The session keeps W, Wold, H, Hold, residNormOld on the device.
 */
int  mfb_als_begin(mfb_context *ctx, mfb_matrix *A, int k,
                   const double *W0, const double *Hold0, double resid_old0,
                   mfb_als **out);
/*
 * Run up to `max_new_iters` further iterations (R/bicluster.R:119-203).
 * fixed_iters != 0 disables the convergence break (:179).
 * Returns MFB_OK (converged flag in *converged, or iteration budget used),
 * MFB_SINGULAR / MFB_ZERO_ROW when the iteration that was running failed the
 * way .fcnnls / the zero-row test fail in R: the iteration counter has been
 * advanced (R's `i <- i + 1` precedes the failure), W/Wold/Hold/residNormOld
 * are those of the last completed iteration, and the caller is expected to
 * call mfb_als_restart() with a fresh W (R's restart(), :107-116) and iterate
 * again.
 *   *iters_total  iterations counted so far (R's `i`)
 *   *converged    1 if the break at :179 fired
 */
int  mfb_als_iterate(mfb_als *s, int max_new_iters, double eps_conv, int fixed_iters,
                     int *iters_total, int *converged);
/*
 * The same, and *maxcol_draws = the number of uniforms NMF::.fcnnls would have consumed from R's RNG during these
 * iterations through max.col(ties.method = "random") (call sites R/bicluster.R:123,150): one
 * max.col row per Lawson-Hanson step back and one per variable entering a passive set, each costing one draw per
 * (near-)tie met while scanning it -- including the structural ties among leading -Inf / zero entries.  The count
 * depends on the values only, so the R wrapper can advance .Random.seed by it (runif(draws)) before it draws the
 * next replicate's start or a restart W: set.seed(); als_nmf() then follows the reference's stream without explicit
 * starts.  -1 when unknown (ranks above 32 use block pivoting; row-sharded sessions).
 */
int  mfb_als_iterate_draws(mfb_als *s, int max_new_iters, double eps_conv, int fixed_iters,
                           int *iters_total, int *converged, long long *maxcol_draws);
int  mfb_als_restart(mfb_als *s, const double *Wnew);
/*
 * Row-sharded replicate: one very large A split by rows over several devices / processes (north star: "very
 * large single matrices are row-sharded, with the k x k and k x n partial products all-reduced by NCCL").  Each
 * rank opens a session on ITS row block A_rows (m_local x n) with its rows of W0 and the full (replicated) Hold0;
 * every rank then calls mfb_als_iterate / mfb_als_restart (with its rows of the new W) / mfb_als_result (its rows of
 * W, the full H) in lockstep.  Per iteration the library asks the host for two sum-all-reduces, IN PLACE on a device
 * buffer and ordered on the context's stream: [W_r'A_r | W_r'W_r] (k n + k^2 doubles, before the H-step NNLS, which
 * every rank then solves identically) and the three iteration scalars (||A - WH||^2 terms, ||dW||^2).  The library
 * has no link-time dependency on NCCL / MPI: the host supplies the collective (ncclAllReduce on `stream`, or
 * torch.distributed in the Python harness).  The callback returns 0 on success.  All status codes, the restart
 * protocol, traces and convergence decisions are identical on every rank (they derive from all-reduced values).
 * resid_old0 = max(A) over the WHOLE matrix.  mfb_als_begin_sharded itself performs one 4-double all-reduce
 * (||A||^2; negative / NA verdict), so all ranks must call it together.
 */
typedef int (*mfb_allreduce_fn)(void *user, double *dev_buf, int64_t count, void *stream);
int  mfb_als_begin_sharded(mfb_context *ctx, mfb_matrix *A_rows, int64_t m_global, int k,
                           const double *W0_rows, const double *Hold0, double resid_old0,
                           mfb_allreduce_fn allreduce, void *user, mfb_als **out);
/*
 * Optional, when all ranks are on ONE node: exchange over peer memory instead of the callback.  Every rank exports
 * the cudaIpc handle (64 bytes) of its exchange block, the host gathers the handles of all ranks (rank order) and
 * every rank connects BEFORE the first iteration.  From then on the per-iteration exchanges are one-shot all-reduces
 * fused into the kernels that consume the sums: each rank raises a system-scope release flag in every peer's block
 * once its W_r'W_r, its folded W_r'A_r (stored straight into every peer's block for up to 8 ranks, read remotely
 * beyond) and its scalars are in place, and the consumers add the ranks' buffers in rank order (bit-identical on
 * every rank; no collective launch, no host involvement).  Only the W'A exchange is a rendezvous on the critical
 * path; W'W and the scalars are consumed a kernel later, when every flag has long arrived.  At most 16 ranks.  mfb_als_shard_connect returns MFB_ERR_UNSUPPORTED when a peer block
 * cannot be mapped (the session then keeps using the callback; the ranks must agree on the outcome).  No rank may
 * call mfb_als_end before every rank has its result (the blocks are read remotely until then).  A peer that never
 * arrives makes the waiting kernels time out (20 s) and the call fail with MFB_ERR_CUDA.
 */
int  mfb_als_shard_export(mfb_als *s, void *handle64);
int  mfb_als_shard_connect(mfb_als *s, int rank, int world, const void *handles /* world x 64 bytes */);
/* Undo a successful connect before the first iteration (another rank could not connect): back to the callback. */
int  mfb_als_shard_disconnect(mfb_als *s);
/*
 * Measurement aid for bench.py: runs `iters` further fixed iterations with a CUDA event recorded (on the
 * context's stream) after every kernel, and returns the mean duration in ms of the four launches of an
 * iteration: ms[0] H-step stream, ms[1] H-step solve, ms[2] W-step stream, ms[3] W-step solve.  A row-sharded
 * session has seven launches with the callback exchange (H stream, fold, invert, H solve, W stream, W solve, finish)
 * and five over peer memory (H stream, fold + wait for the peers + invert, H solve, W stream, W solve + wait for
 * the peers' scalars + finish).  ms must hold 8 doubles.
 */
int  mfb_als_profile(mfb_als *s, int iters, double *ms /* [8] */);
/* trace: iters_total x 3 row-major (residNorm, rms dW, rms dH) per completed iteration (NaN rows for failed ones). */
int  mfb_als_result(mfb_als *s, double *W, double *H, double *resid, double *trace);
void mfb_als_end(mfb_als *s);

/* Convenience: begin + iterate + result + end on a resident matrix. */
int  mfb_als_run(mfb_context *ctx, mfb_matrix *A, int k,
                 const double *W0, const double *Hold0, double resid_old0,
                 int max_iter, double eps_conv, int fixed_iters,
                 double *W, double *H, double *resid, int *iters_done, int *converged,
                 double *trace);
/* Same from a HOST matrix (upload inside): the end-to-end call bench.py times. */
int  mfb_als_run_host(mfb_context *ctx, const double *A_host, int64_t m, int64_t n, int k,
                      const double *W0, const double *Hold0,
                      int max_iter, double eps_conv, int fixed_iters,
                      double *W, double *H, double *resid, int *iters_done, int *converged,
                      double *trace);

/* Exact NNLS  min ||C K - B||_F, K >= 0  for every column of B: the unit-testable
 * equivalent of NMF::.fcnnls(x = C, y = B)[[1]].  C is r x k, B is r x p, K is k x p. */
int  mfb_nnls(mfb_context *ctx, const double *C, const double *B, int64_t r, int k, int64_t p, double *K);
/* ... and the number of max.col draws of that .fcnnls call (see mfb_als_iterate_draws); -1 for k > 32 */
int  mfb_nnls_draws(mfb_context *ctx, const double *C, const double *B, int64_t r, int k, int64_t pcols,
                    double *K, long long *maxcol_draws);

/* ---- snmf (R/bicluster.R:341-441; the NMF::nmf(method = "snmf/r") call at :363) ---------------------------------
 * Kim & Park's sparse NMF by alternating NNLS as shipped in NMF 0.21.0 (.nmf_snmf, version "R"):
 *   H <- argmin_{H >= 0} ||[W; sqrt(beta) 1'] H - [A; 0]||     i.e. (W'W + beta 11') H = W'A
 *   W <- argmin_{W >= 0} ||[H'; eta I] W' - [A'; 0]||          i.e. (HH' + eta^2 I) W' = H A'
 * on the ALS kernels (same two passes over A per iteration; the regularisers are added where the Gram matrices are
 * formed).  An snmf session is an mfb_als session: mfb_als_result / mfb_als_end apply (resid / trace are not R's).
 *   W0    m x k start (NMF's seed; its columns are scaled to unit length here, as .nmf_snmf does)
 *   eta < 0 means max(A) (NMF's default); mfBiclust passes beta = eta = mean(A) (R/bicluster.R:351-352)
 */
int  mfb_snmf_begin(mfb_context *ctx, mfb_matrix *A, int k, const double *W0, double beta, double eta, mfb_als **out);
/*
 * Up to `max_new_iters` further iterations including NMF's convergence test -- on the first iteration after a
 * (re)start and on every fifth: arg-max memberships of the rows of W and columns of H (max.col, ties "first")
 * changed in at most `wminchange` rows / no column for `iconv` consecutive tests (bi_conv = c(0, 10)) AND the mean
 * absolute KKT residual <= eps_conv (1e-4) times its first value.  The residual needs W'A of the NEW W: the stream
 * kernel of the next H-step runs ahead for it and is not repeated when the iteration goes on.
 * Returns MFB_OK (*converged = 1, or budget used), MFB_ZERO_ROW (zero row in H: .nmf_snmf redraws W -- call
 * mfb_snmf_restart with matrix(runif(m * k), m, k) and iterate again; it gives up at the 10th), MFB_SINGULAR (solve()
 * inside .fcnnls failed: R/bicluster.R:399-410 halves eta and starts over).  The iteration counter is advanced for a
 * failed iteration as in R.  *objective = .snmf.objective = 1/2 (||A - WH||^2 + eta sum(W^2) + beta sum(colSums(H)^2)).
 */
int  mfb_snmf_iterate(mfb_als *s, int max_new_iters, int wminchange, int iconv, double eps_conv,
                      int *iters_total, int *converged, double *objective);
int  mfb_snmf_restart(mfb_als *s, const double *Wnew);

/* ---- otsuHelper / threshold (R/BiclusterStrategy.R:442-456, 508-528; EBImage::otsu call site :448) */
/* One Otsu threshold per COLUMN of the rows x cols matrix M. */
int  mfb_otsu(mfb_context *ctx, const double *M, int64_t rows, int64_t cols, double *th);
/* margin = 2: th per column; margin = 1: th per row.  out: rows x cols int32 (R logical), column-major. */
int  mfb_threshold(mfb_context *ctx, const double *M, int64_t rows, int64_t cols,
                   const double *th, int margin, int32_t *out);

/* ---- bicluster post-processing (R/helperFunctions.R:150-206 filter.biclust, :226-247 overlap, :303-308 sizes,
 *      :336-345 union.biclust) on the membership matrices threshold() produces */
/* rowxBicluster: m x k, biclusterxCol: k x n, R logicals (int32, column-major; NA is not TRUE).
 * sizes[i] = |rows_i| * |cols_i|;  overlaps: k x k column-major, entry (i, j) at i + k * j =
 * |rows_i & rows_j| * |cols_i & cols_j| / sizes[i]  -- overlap(BiclusterRows, BiclusterCols, matrix = TRUE). */
int  mfb_bicluster_overlap(mfb_context *ctx, const int32_t *rowxBicluster, int64_t m, const int32_t *biclusterxCol,
                           int64_t n, int k, double *sizes, double *overlaps);
/* biclustered: m x n column-major, 1 where any bicluster b with choose[b] TRUE (choose == NULL: all) contains the
 * element, else 0 -- union.biclust(RowxBiclust, BiclustxCol, choose). */
int  mfb_bicluster_union(mfb_context *ctx, const int32_t *rowxBicluster, int64_t m, const int32_t *biclusterxCol,
                         int64_t n, int k, const int32_t *choose, double *biclustered);

/* ---- svd_pca (R/bicluster.R:528-539 -> stats::prcomp(t(A), rank. = k, center = FALSE)) */
/* W: m x k (left singular vectors of A), H: k x n (= t(W) %*% A), sdev-free.  kout = min(k, m, n). */
int  mfb_svd_pca(mfb_context *ctx, mfb_matrix *A, int k, double *W, double *H, int *kout);

/* ---- nipals_pca (R/bicluster.R:218-242 -> nipals::nipals, call site :226) -- */
/* scores: m x ncomp (unit-norm columns), loadings: n x ncomp, eig: ncomp, iters: ncomp.
 * NaN entries are treated as NA (nipals' has.na branch: sums over observed cells only).  x is not modified. */
int  mfb_nipals(mfb_context *ctx, mfb_matrix *x, int ncomp, int center, int scale,
                int maxiter, double tol, int gramschmidt,
                double *scores, double *loadings, double *eig, int *iters);

/* ---- bcv (R/bcv.R:164-239 bcvGivenKs; prcomp call site :184, MASS::ginv :222) */
/*
 * Errors of the cells [cell_begin, cell_end) of the h x h grid (cell = (x-1)*h + (y-1),
 * x = row fold, y = column fold, both 1-based as drawn by sample() at :170-174) for
 * every k in ks[0..nks).  err: (cell_end - cell_begin) x nks, row-major.  The caller
 * (R wrapper / other ranks) sums over cells (:231-233).
 */
int  mfb_bcv_cells(mfb_context *ctx, mfb_matrix *Y,
                   const int32_t *row_fold, const int32_t *col_fold, int holdouts,
                   const int32_t *ks, int nks, int cell_begin, int cell_end, double *err);
/*
 * The same for `nfolds` fold draws at once -- auto_bcv's outer iterations (R/bcv.R:56-58): their fold draws do not
 * depend on the data, so the cells of several bcv() calls form ONE work list.  row_folds: nfolds x n, col_folds:
 * nfolds x p (draw after draw); cells are numbered draw * h^2 + (x-1)*h + (y-1); err: (cell_end - cell_begin) x nks.
 * The cells of a batch share one launch of the eigen-solver; values do not depend on the batching.
 */
int  mfb_bcv_cells_multi(mfb_context *ctx, mfb_matrix *Y, int nfolds,
                         const int32_t *row_folds, const int32_t *col_folds, int holdouts,
                         const int32_t *ks, int nks, int cell_begin, int cell_end, double *err);

/*
 * The same work list for a matrix WITH missing entries (NaN = NA): the reference then factors every hold-in block
 * with nipals_pca(center = TRUE, scale = FALSE) instead of prcomp (R/bcv.R:177-181) -- scores and loadings without
 * singular values -- and NA entries of the products are dropped by sum(..., na.rm = TRUE) (:225-227).  At most 60
 * components.  MFB_ERR_ARG when a hold-in block has a row or column without a non-zero observed entry (clean()
 * would drop it and the reference's products become non-conformable).
 */
int  mfb_bcv_cells_na(mfb_context *ctx, mfb_matrix *Y, int nfolds,
                      const int32_t *row_folds, const int32_t *col_folds, int holdouts,
                      const int32_t *ks, int nks, int cell_begin, int cell_end, double *err);

#ifdef __cplusplus
}
#endif
#endif /* MFBICLUST_H */
