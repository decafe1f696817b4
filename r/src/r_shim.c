/* .Call shim between the mfBiclust R package and libmfbiclust_cuda.so (include/mfbiclust.h).
 *
 * Pointer marshalling only: REAL()/INTEGER()/LOGICAL() extraction, allocation of result SEXPs, Rf_error on a
 * negative status.  No arithmetic and no CPU fallback live here.  The seam it fills is the reference's
 * do.call(get(method), ...) at R/BiclusterStrategy.R:82-87 and the direct calls at R/addStrat.R:72, R/bcv.R:155,
 * R/BiclusterStrategy.R:485-494; the R side is r/R/backends.R.
 *
 * R is not installed in the build image of this repository: the file is compiled there against the minimal
 * declarations in r/stub/ (tests/test_r_shim.py: gcc -fsyntax-only -Wall -Werror), which checks every call against
 * the prototypes of include/mfbiclust.h.  With a real R: R CMD SHLIB with src/Makevars.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "mfbiclust.h"

static mfb_context *ctx_of(SEXP x) {
    mfb_context *c = (mfb_context *) R_ExternalPtrAddr(x);
    if (!c) Rf_error("mfBiclust: GPU context has been released");
    return c;
}
static void ctx_finalizer(SEXP x) { mfb_context_destroy((mfb_context *) R_ExternalPtrAddr(x)); R_ClearExternalPtr(x); }
static void mat_finalizer(SEXP x) { mfb_matrix_free((mfb_matrix *) R_ExternalPtrAddr(x)); R_ClearExternalPtr(x); }
static void als_finalizer(SEXP x) { mfb_als_end((mfb_als *) R_ExternalPtrAddr(x)); R_ClearExternalPtr(x); }
#define CHECK(call) do { int st_ = (call); if (st_ < 0) Rf_error("%s", mfb_last_error()); } while (0)

SEXP C_mfb_context(SEXP device) {                       /* options(mfBiclust.gpu = 0L) */
    mfb_context *c = NULL;
    CHECK(mfb_context_create(Rf_asInteger(device), &c));  /* fails loudly without a B200 */
    SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP C_mfb_upload(SEXP ctx, SEXP A) {                   /* A: numeric matrix, column-major as is */
    SEXP dim = Rf_getAttrib(A, R_DimSymbol);
    mfb_matrix *m = NULL;
    CHECK(mfb_matrix_upload(ctx_of(ctx), REAL(A), INTEGER(dim)[0], INTEGER(dim)[1], &m));
    SEXP p = PROTECT(R_MakeExternalPtr(m, R_NilValue, ctx));   /* keeps the context alive */
    R_RegisterCFinalizerEx(p, mat_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP C_mfb_stats(SEXP A) {                              /* c(min, max, sumsq, nNA, sum) */
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 5));
    CHECK(mfb_matrix_stats((mfb_matrix *) R_ExternalPtrAddr(A), REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP C_mfb_als_begin(SEXP ctx, SEXP A, SEXP k, SEXP W0, SEXP Hold0, SEXP residOld) {
    mfb_als *s = NULL;
    CHECK(mfb_als_begin(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(A), Rf_asInteger(k), REAL(W0), REAL(Hold0),
                        Rf_asReal(residOld), &s));
    SEXP p = PROTECT(R_MakeExternalPtr(s, R_NilValue, A));
    R_RegisterCFinalizerEx(p, als_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP C_mfb_als_iterate(SEXP s, SEXP maxNew, SEXP eps, SEXP fixed) {   /* c(status, i, converged, maxcol_draws) */
    int it = 0, conv = 0;
    long long draws = 0;                                   /* uniforms NMF::.fcnnls would have consumed (-1: unknown) */
    int st = mfb_als_iterate_draws((mfb_als *) R_ExternalPtrAddr(s), Rf_asInteger(maxNew), Rf_asReal(eps),
                                   Rf_asLogical(fixed), &it, &conv, &draws);
    if (st < 0) Rf_error("%s", mfb_last_error());
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 4));
    INTEGER(out)[0] = st; INTEGER(out)[1] = it; INTEGER(out)[2] = conv; INTEGER(out)[3] = (int) draws;
    UNPROTECT(1);
    return out;
}

SEXP C_mfb_als_restart(SEXP s, SEXP Wnew) { CHECK(mfb_als_restart((mfb_als *) R_ExternalPtrAddr(s), REAL(Wnew))); return R_NilValue; }

SEXP C_mfb_als_result(SEXP s, SEXP m, SEXP n, SEXP k, SEXP iters) {   /* list(W, H, resid, trace) */
    int M = Rf_asInteger(m), N = Rf_asInteger(n), K = Rf_asInteger(k), I = Rf_asInteger(iters);
    SEXP W = PROTECT(Rf_allocMatrix(REALSXP, M, K)), H = PROTECT(Rf_allocMatrix(REALSXP, K, N));
    SEXP tr = PROTECT(Rf_allocMatrix(REALSXP, 3, I > 0 ? I : 1)), r = PROTECT(Rf_allocVector(REALSXP, 1));
    CHECK(mfb_als_result((mfb_als *) R_ExternalPtrAddr(s), REAL(W), REAL(H), REAL(r), REAL(tr)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, W); SET_VECTOR_ELT(out, 1, H); SET_VECTOR_ELT(out, 2, r); SET_VECTOR_ELT(out, 3, tr);
    UNPROTECT(5);
    return out;
}

SEXP C_mfb_svd_pca(SEXP ctx, SEXP A, SEXP k, SEXP m, SEXP n) {        /* list(W, H) */
    int K = Rf_asInteger(k), M = Rf_asInteger(m), N = Rf_asInteger(n), kk = K < M ? (K < N ? K : N) : (M < N ? M : N);
    SEXP W = PROTECT(Rf_allocMatrix(REALSXP, M, kk)), H = PROTECT(Rf_allocMatrix(REALSXP, kk, N));
    int kout = 0;
    CHECK(mfb_svd_pca(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(A), K, REAL(W), REAL(H), &kout));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, W); SET_VECTOR_ELT(out, 1, H);
    UNPROTECT(3);
    return out;
}

SEXP C_mfb_nipals(SEXP ctx, SEXP x, SEXP m, SEXP n, SEXP ncomp, SEXP center, SEXP scale, SEXP maxiter, SEXP tol,
                  SEXP gramschmidt) {                                  /* list(scores, loadings, eig, iter, status) */
    int M = Rf_asInteger(m), N = Rf_asInteger(n), K = Rf_asInteger(ncomp);
    SEXP sc = PROTECT(Rf_allocMatrix(REALSXP, M, K)), ld = PROTECT(Rf_allocMatrix(REALSXP, N, K));
    SEXP eig = PROTECT(Rf_allocVector(REALSXP, K)), it = PROTECT(Rf_allocVector(INTSXP, K));
    int st = mfb_nipals(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(x), K, Rf_asLogical(center), Rf_asLogical(scale),
                        Rf_asInteger(maxiter), Rf_asReal(tol), Rf_asLogical(gramschmidt), REAL(sc), REAL(ld), REAL(eig),
                        INTEGER(it));
    if (st < 0) Rf_error("%s", mfb_last_error());
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, sc); SET_VECTOR_ELT(out, 1, ld); SET_VECTOR_ELT(out, 2, eig); SET_VECTOR_ELT(out, 3, it);
    SET_VECTOR_ELT(out, 4, Rf_ScalarInteger(st));
    UNPROTECT(5);
    return out;
}

SEXP C_mfb_bcv_cells(SEXP ctx, SEXP Y, SEXP rowFold, SEXP colFold, SEXP holdouts, SEXP ks, SEXP cellBegin, SEXP cellEnd) {
    int nks = LENGTH(ks), c0 = Rf_asInteger(cellBegin), c1 = Rf_asInteger(cellEnd);
    SEXP err = PROTECT(Rf_allocMatrix(REALSXP, nks, c1 - c0));         /* row-major cells x nks == column-major nks x cells */
    CHECK(mfb_bcv_cells(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(Y), INTEGER(rowFold), INTEGER(colFold),
                        Rf_asInteger(holdouts), INTEGER(ks), nks, c0, c1, REAL(err)));
    UNPROTECT(1);
    return err;                                                        /* caller: rowSums(err) */
}

SEXP C_mfb_otsu(SEXP ctx, SEXP M) {
    SEXP dim = Rf_getAttrib(M, R_DimSymbol);
    SEXP th = PROTECT(Rf_allocVector(REALSXP, INTEGER(dim)[1]));
    CHECK(mfb_otsu(ctx_of(ctx), REAL(M), INTEGER(dim)[0], INTEGER(dim)[1], REAL(th)));
    UNPROTECT(1);
    return th;
}

SEXP C_mfb_threshold(SEXP ctx, SEXP M, SEXP th, SEXP margin) {
    SEXP dim = Rf_getAttrib(M, R_DimSymbol);
    SEXP out = PROTECT(Rf_allocMatrix(LGLSXP, INTEGER(dim)[0], INTEGER(dim)[1]));
    CHECK(mfb_threshold(ctx_of(ctx), REAL(M), INTEGER(dim)[0], INTEGER(dim)[1], REAL(th), Rf_asInteger(margin),
                        (int32_t *) LOGICAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP C_mfb_upload_f32(SEXP ctx, SEXP A) {               /* options(mfBiclust.precision = "f32"): A resident in FP32 */
    SEXP dim = Rf_getAttrib(A, R_DimSymbol);
    mfb_matrix *m = NULL;
    CHECK(mfb_matrix_upload_f32(ctx_of(ctx), REAL(A), INTEGER(dim)[0], INTEGER(dim)[1], &m));
    SEXP p = PROTECT(R_MakeExternalPtr(m, R_NilValue, ctx));
    R_RegisterCFinalizerEx(p, mat_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP C_mfb_shift(SEXP A, SEXP s) { CHECK(mfb_matrix_shift((mfb_matrix *) R_ExternalPtrAddr(A), Rf_asReal(s))); return R_NilValue; }

/* auto_bcv: the cells of several fold draws in one call; rowFolds: n x nfolds, colFolds: p x nfolds integer matrices
 * (one draw per column == draw after draw in memory) */
SEXP C_mfb_bcv_cells_multi(SEXP ctx, SEXP Y, SEXP rowFolds, SEXP colFolds, SEXP holdouts, SEXP ks, SEXP cellBegin,
                           SEXP cellEnd) {
    int nks = LENGTH(ks), c0 = Rf_asInteger(cellBegin), c1 = Rf_asInteger(cellEnd);
    int nfolds = INTEGER(Rf_getAttrib(rowFolds, R_DimSymbol))[1];
    SEXP err = PROTECT(Rf_allocMatrix(REALSXP, nks, c1 - c0));
    CHECK(mfb_bcv_cells_multi(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(Y), nfolds, INTEGER(rowFolds),
                              INTEGER(colFolds), Rf_asInteger(holdouts), INTEGER(ks), nks, c0, c1, REAL(err)));
    UNPROTECT(1);
    return err;
}

/* the same work list for a matrix with NA entries: NIPALS-based pca per hold-in block (R/bcv.R:177-181) */
SEXP C_mfb_bcv_cells_na(SEXP ctx, SEXP Y, SEXP rowFolds, SEXP colFolds, SEXP holdouts, SEXP ks, SEXP cellBegin,
                        SEXP cellEnd) {
    int nks = LENGTH(ks), c0 = Rf_asInteger(cellBegin), c1 = Rf_asInteger(cellEnd);
    int nfolds = INTEGER(Rf_getAttrib(rowFolds, R_DimSymbol))[1];
    SEXP err = PROTECT(Rf_allocMatrix(REALSXP, nks, c1 - c0));
    CHECK(mfb_bcv_cells_na(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(Y), nfolds, INTEGER(rowFolds),
                           INTEGER(colFolds), Rf_asInteger(holdouts), INTEGER(ks), nks, c0, c1, REAL(err)));
    UNPROTECT(1);
    return err;
}

SEXP C_mfb_trim(SEXP ctx) { CHECK(mfb_context_trim(ctx_of(ctx))); return R_NilValue; }

/* sizes() + overlap(matrix = TRUE) of all biclusters: list(sizes, overlaps)  (R/helperFunctions.R:226-247, 303-308) */
SEXP C_mfb_bicluster_overlap(SEXP ctx, SEXP rowx, SEXP bicx) {
    SEXP dr = Rf_getAttrib(rowx, R_DimSymbol), dc = Rf_getAttrib(bicx, R_DimSymbol);
    const int k = INTEGER(dr)[1];
    SEXP sz = PROTECT(Rf_allocVector(REALSXP, k));
    SEXP ov = PROTECT(Rf_allocMatrix(REALSXP, k, k));
    CHECK(mfb_bicluster_overlap(ctx_of(ctx), (const int32_t *) LOGICAL(rowx), INTEGER(dr)[0], (const int32_t *) LOGICAL(bicx),
                                INTEGER(dc)[1], k, REAL(sz), REAL(ov)));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, sz);
    SET_VECTOR_ELT(out, 1, ov);
    UNPROTECT(3);
    return out;
}

/* union.biclust(RowxBiclust, BiclustxCol, choose)  (R/helperFunctions.R:336-345) */
SEXP C_mfb_bicluster_union(SEXP ctx, SEXP rowx, SEXP bicx, SEXP choose) {
    SEXP dr = Rf_getAttrib(rowx, R_DimSymbol), dc = Rf_getAttrib(bicx, R_DimSymbol);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, INTEGER(dr)[0], INTEGER(dc)[1]));
    CHECK(mfb_bicluster_union(ctx_of(ctx), (const int32_t *) LOGICAL(rowx), INTEGER(dr)[0], (const int32_t *) LOGICAL(bicx),
                              INTEGER(dc)[1], INTEGER(dr)[1], (const int32_t *) LOGICAL(choose), REAL(out)));
    UNPROTECT(1);
    return out;
}

/* ---- snmf: the NMF::nmf(method = "snmf/r") run behind R/bicluster.R:363 ---- */
SEXP C_mfb_snmf_begin(SEXP ctx, SEXP A, SEXP k, SEXP W0, SEXP beta, SEXP eta) {
    mfb_als *s = NULL;
    CHECK(mfb_snmf_begin(ctx_of(ctx), (mfb_matrix *) R_ExternalPtrAddr(A), Rf_asInteger(k), REAL(W0), Rf_asReal(beta),
                         Rf_asReal(eta), &s));
    SEXP p = PROTECT(R_MakeExternalPtr(s, R_NilValue, A));
    R_RegisterCFinalizerEx(p, als_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP C_mfb_snmf_iterate(SEXP s, SEXP maxNew, SEXP biConv, SEXP epsConv) {   /* c(status, i, converged, objective) */
    int it = 0, conv = 0;
    double obj = NA_REAL;
    int st = mfb_snmf_iterate((mfb_als *) R_ExternalPtrAddr(s), Rf_asInteger(maxNew), (int) REAL(biConv)[0],
                              (int) REAL(biConv)[1], Rf_asReal(epsConv), &it, &conv, &obj);
    if (st < 0) Rf_error("%s", mfb_last_error());
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
    REAL(out)[0] = st; REAL(out)[1] = it; REAL(out)[2] = conv; REAL(out)[3] = obj;
    UNPROTECT(1);
    return out;
}

SEXP C_mfb_snmf_restart(SEXP s, SEXP Wnew) { CHECK(mfb_snmf_restart((mfb_als *) R_ExternalPtrAddr(s), REAL(Wnew))); return R_NilValue; }

static const R_CallMethodDef calls[] = {
    {"C_mfb_context", (DL_FUNC) &C_mfb_context, 1}, {"C_mfb_upload", (DL_FUNC) &C_mfb_upload, 2},
    {"C_mfb_stats", (DL_FUNC) &C_mfb_stats, 1}, {"C_mfb_als_begin", (DL_FUNC) &C_mfb_als_begin, 6},
    {"C_mfb_als_iterate", (DL_FUNC) &C_mfb_als_iterate, 4}, {"C_mfb_als_restart", (DL_FUNC) &C_mfb_als_restart, 2},
    {"C_mfb_als_result", (DL_FUNC) &C_mfb_als_result, 5}, {"C_mfb_svd_pca", (DL_FUNC) &C_mfb_svd_pca, 5},
    {"C_mfb_nipals", (DL_FUNC) &C_mfb_nipals, 10}, {"C_mfb_bcv_cells", (DL_FUNC) &C_mfb_bcv_cells, 8},
    {"C_mfb_otsu", (DL_FUNC) &C_mfb_otsu, 2}, {"C_mfb_threshold", (DL_FUNC) &C_mfb_threshold, 4},
    {"C_mfb_upload_f32", (DL_FUNC) &C_mfb_upload_f32, 2}, {"C_mfb_shift", (DL_FUNC) &C_mfb_shift, 2},
    {"C_mfb_bcv_cells_multi", (DL_FUNC) &C_mfb_bcv_cells_multi, 8}, {"C_mfb_trim", (DL_FUNC) &C_mfb_trim, 1},
    {"C_mfb_bcv_cells_na", (DL_FUNC) &C_mfb_bcv_cells_na, 8},
    {"C_mfb_bicluster_overlap", (DL_FUNC) &C_mfb_bicluster_overlap, 3}, {"C_mfb_bicluster_union", (DL_FUNC) &C_mfb_bicluster_union, 4},
    {"C_mfb_snmf_begin", (DL_FUNC) &C_mfb_snmf_begin, 6}, {"C_mfb_snmf_iterate", (DL_FUNC) &C_mfb_snmf_iterate, 4},
    {"C_mfb_snmf_restart", (DL_FUNC) &C_mfb_snmf_restart, 2},
    {NULL, NULL, 0}};
void R_init_mfBiclust(DllInfo *dll) { R_registerRoutines(dll, NULL, calls, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }
