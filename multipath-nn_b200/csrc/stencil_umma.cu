// placeholder until the tcgen05 kernels land
#include "common.cuh"
#include "../../include/mpnn.h"
extern "C" int mpnn_has_umma(void) { return 0; }
int mpnn_stencil_gemm_umma(const void*, int, const void*, int, const void*, int, const float*, void*, int, int,
                           void*, int, int, Geom, float*, int, int*, int, cudaStream_t) {
    mpnn_set_error("tcgen05 path not built"); return MPNN_ERR_UNSUPPORTED;
}
int mpnn_stencil_wgrad_umma(const void*, int, int, float*, const void*, int, int, float*, const void*, int,
                            int, float*, int, Geom, cudaStream_t) {
    mpnn_set_error("tcgen05 path not built"); return MPNN_ERR_UNSUPPORTED;
}
extern "C" int mpnn_umma_selftest(const void*, const void*, float*, int, int, int, int, int, int, int, int, void*) {
    mpnn_set_error("tcgen05 path not built"); return MPNN_ERR_UNSUPPORTED;
}
