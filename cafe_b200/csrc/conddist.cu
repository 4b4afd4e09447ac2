// conddist.cu — K4 (placeholder until the simulate-and-prune kernels land).
#include "common.cuh"
int run_conditional_distribution(cafe_gpu_ctx* ctx, int, const double*, uint64_t, double*) {
    CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "conditional_distribution: not built yet");
}
