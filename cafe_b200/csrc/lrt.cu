// lrt.cu — branch-stretch likelihood-ratio test for every family at once.
//
// Replaces the per-family loop of __cafe_likelihood_ratio_test_thread_func (cafe/cafe_main.c:342-396): for every non-root branch b
// the reference lengthens b by rint(0.15 * length) for as long as the family's maximum root likelihood grows, pruning the whole
// tree again after every step.  The sequence of lengths does not depend on the family, so here one step is ONE batched
// evaluation: K1 builds the single matrix of the lengthened branch (key (int t', lambda_b, mu_b) appended behind the tree's keys,
// the other matrices stay), K2 prunes every family, and k_lrt_step advances the per-family state (previous best, still growing?)
// on the device; the host reads back one counter per step and stops when no family is still growing.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

__global__ void k_lrt_begin_branch(const double* __restrict__ base, int F, double* __restrict__ prev, int* __restrict__ steps) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    prev[f] = base[f];
    steps[f] = 0;
}

// one `while (prevlh < nextlh)` turn of cafe_main.c:377-386 for every family that is still growing
__global__ void k_lrt_step(const double* __restrict__ maxlik, int F, double* __restrict__ prev, int* __restrict__ steps,
                           unsigned char* __restrict__ active, int* __restrict__ n_growing) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F || !active[f]) return;
    const double next = maxlik[f];
    if (prev[f] < next) {
        prev[f] = next;
        steps[f]++;
        atomicAdd(n_growing, 1);
    } else {
        active[f] = 0;
    }
}

struct LrtBuffers {
    double *d_base = nullptr, *d_prev = nullptr;
    int *d_steps = nullptr, *d_count = nullptr;
    unsigned char* d_active = nullptr;
    ~LrtBuffers() { cudaFree(d_base); cudaFree(d_prev); cudaFree(d_steps); cudaFree(d_count); cudaFree(d_active); }
};

}  // namespace


int run_lrt_branch_stretch(cafe_gpu_ctx* ctx, const uint8_t* tested, const double* lengthened_mu, double* base_out, double* best_out,
                           int32_t* steps_out) {
    const int F = ctx->F, n = ctx->n_nodes;
    const int D = (int)ctx->keys.size();
    if ((size_t)D + 1 > ctx->mat_cap) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "lrt: no room for one more matrix");
    LrtBuffers B;
    CAFE_CK(ctx, cudaMalloc(&B.d_base, F * sizeof(double)));
    CAFE_CK(ctx, cudaMalloc(&B.d_prev, F * sizeof(double)));
    CAFE_CK(ctx, cudaMalloc(&B.d_steps, F * sizeof(int)));
    CAFE_CK(ctx, cudaMalloc(&B.d_count, sizeof(int)));
    CAFE_CK(ctx, cudaMalloc(&B.d_active, F));
    const int threads = 256, blocks = (F + threads - 1) / threads;

    // the unlengthened tree: maxlh of cafe_main.c:365
    int rc = launch_prune(ctx, nullptr);
    if (rc) return rc;
    CAFE_CK(ctx, cudaMemcpyAsync(B.d_base, ctx->d_maxlik, F * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (base_out) CAFE_CK(ctx, cudaMemcpyAsync(base_out, B.d_base, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));

    // the thread's tree copy keeps the parsed (double) lengths only until the first tested family has gone over a branch: the
    // length is restored through `int old_bl` (cafe_main.c:350,375,390).  So the first tested family starts from the parsed
    // length, every later one from the truncated length.
    int first_tested = -1;
    for (int f = 0; f < F && first_tested < 0; ++f)
        if (!tested || tested[f]) first_tested = f;
    std::vector<unsigned char> mask_rest(F), mask_first(F, 0);
    for (int f = 0; f < F; ++f) mask_rest[f] = (!tested || tested[f]) ? 1 : 0;
    if (first_tested >= 0) mask_first[first_tested] = 1;

    // the tree's own keys come back on every way out
    struct KeyGuard {
        cafe_gpu_ctx* ctx;
        const std::vector<BdKey> keys0;
        const std::vector<int> node_key0;
        ~KeyGuard() {
            ctx->keys = keys0;
            ctx->node_key = node_key0;
            build_schedule(ctx);
            ctx->results_valid = false;  // d_maxlik holds the last lengthened tree, not the tree's own
        }
    } guard{ctx, ctx->keys, ctx->node_key};
    const std::vector<BdKey>& keys0 = guard.keys0;
    const std::vector<int>& node_key0 = guard.node_key0;

    for (int b = 0; b < n; ++b) {
        double* best_row = best_out + (size_t)b * F;
        int32_t* steps_row = steps_out ? steps_out + (size_t)b * F : nullptr;
        if (b == ctx->root) {
            std::fill(best_row, best_row + F, -1.0);
            if (steps_row) std::fill(steps_row, steps_row + F, 0);
            continue;
        }
        k_lrt_begin_branch<<<blocks, threads, 0, ctx->stream>>>(B.d_base, F, B.d_prev, B.d_steps);
        ctx->launches++;
        const double parsed = ctx->branchlength[b], truncated = (double)(int)parsed;
        const bool two_starts = parsed != truncated && first_tested >= 0;
        for (int variant = 0; variant < (two_starts ? 2 : 1); ++variant) {
            // variant 0: everyone from the truncated length (minus the first tested family when the parsed length is fractional)
            // variant 1: the first tested family from the parsed length
            std::vector<unsigned char> mask = (variant == 0) ? mask_rest : mask_first;
            if (variant == 0 && two_starts) mask[first_tested] = 0;
            if (std::none_of(mask.begin(), mask.end(), [](unsigned char m) { return m != 0; })) continue;
            CAFE_CK(ctx, cudaMemcpyAsync(B.d_active, mask.data(), F, cudaMemcpyHostToDevice, ctx->stream));
            CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));  // `mask` is pageable and dies with this scope
            double bl = (variant == 0) ? truncated : parsed;
            for (int iter = 0; iter < 100000; ++iter) {
                bl += rint(bl * 0.15);                                             // cafe_main.c:380
                ctx->keys = keys0;
                // birthdeath_cache_get_matrix, birthdeath.c:363-370
                ctx->keys.push_back(BdKey{(int)bl, ctx->lambda[b], lengthened_mu ? lengthened_mu[b] : ctx->mu[b]});
                ctx->node_key = node_key0;
                ctx->node_key[b] = D;
                rc = build_schedule(ctx);
                if (!rc) rc = build_one_matrix(ctx, D);
                if (!rc) rc = launch_prune(ctx, nullptr);
                if (rc) return rc;
                CAFE_CK(ctx, cudaMemsetAsync(B.d_count, 0, sizeof(int), ctx->stream));
                k_lrt_step<<<blocks, threads, 0, ctx->stream>>>(ctx->d_maxlik, F, B.d_prev, B.d_steps, B.d_active, B.d_count);
                ctx->launches++;
                int growing = 0;
                CAFE_CK(ctx, cudaMemcpyAsync(&growing, B.d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
                if (growing == 0) break;
            }
        }
        CAFE_CK(ctx, cudaMemcpyAsync(best_row, B.d_prev, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (steps_row) CAFE_CK(ctx, cudaMemcpyAsync(steps_row, B.d_steps, F * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return CAFE_GPU_OK;
}
