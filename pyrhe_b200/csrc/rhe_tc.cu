// Tensor-core (tcgen05 / TMEM / TMA) path -- placeholder until the int8 kernels land.
#include "rhe_common.cuh"

int rhe_tc_create(rhe_ctx*) { rhe_set_error("RHE_PATH_TCGEN05 is not built yet"); return RHE_ERR_UNSUPPORTED; }
void rhe_tc_destroy(rhe_ctx*) {}
int rhe_tc_set_rhs(rhe_ctx*, cudaStream_t) { return RHE_ERR_UNSUPPORTED; }
int rhe_tc_pass_a(rhe_ctx*, const uint8_t*, int, cudaStream_t) { return RHE_ERR_UNSUPPORTED; }
int rhe_tc_pass_b(rhe_ctx*, const uint8_t*, int, const int32_t*, const int32_t*, float*, float*, cudaStream_t) { return RHE_ERR_UNSUPPORTED; }
