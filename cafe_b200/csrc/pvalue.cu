// pvalue.cu — K5 (placeholder until the p-value kernels land).
#include "common.cuh"
int run_pvalues(cafe_gpu_ctx* ctx, const double*, int, int, double*) {
    CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "pvalues: not built yet");
}
