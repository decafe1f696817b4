// Entry points declared in include/mfbiclust.h whose kernels are not written yet.
#include "common.cuh"
extern "C" {
int mfb_svd_pca(mfb_context *, mfb_matrix *, int, double *, double *, int *) {
    mfb_set_error("mfb_svd_pca: not implemented yet");
    return MFB_ERR_UNSUPPORTED;
}
int mfb_nipals(mfb_context *, mfb_matrix *, int, int, int, int, double, int, double *, double *, double *, int *) {
    mfb_set_error("mfb_nipals: not implemented yet");
    return MFB_ERR_UNSUPPORTED;
}
int mfb_bcv_cells(mfb_context *, mfb_matrix *, const int32_t *, const int32_t *, int, const int32_t *, int, int, int,
                  double *) {
    mfb_set_error("mfb_bcv_cells: not implemented yet");
    return MFB_ERR_UNSUPPORTED;
}
}
