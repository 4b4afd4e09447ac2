// lrt.cu — branch-stretch likelihood-ratio test for every family at once.
//
// Replaces the per-family loop of __cafe_likelihood_ratio_test_thread_func (cafe/cafe_main.c:342-396): for every non-root branch b
// the reference lengthens b by rint(0.15 * length) for as long as the family's maximum root likelihood grows, pruning the whole
// tree again after every step.  The sequence of lengths does not depend on the family, so here one step is ONE batched
// evaluation: K1 builds the single matrix of the lengthened branch (key (int t', lambda_b, mu_b) appended behind the tree's keys,
// the other matrices stay), K2 prunes the families that are still growing (packed into a table of their own - most families stop
// after one or two steps), and k_lrt_step advances the per-family state (previous best, step count) on the device; the host reads
// back one flag per surviving family and stops when none is left.
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

// the families that are still growing, packed: counts_c[leaf][i] = counts[leaf][idx[i]] (leaf-major, same leaf stride)
__global__ void k_lrt_gather_counts(const int* __restrict__ counts, size_t leaf_stride, const int* __restrict__ idx, int n_active,
                                    int* __restrict__ counts_c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const size_t row = (size_t)blockIdx.y * leaf_stride;
    counts_c[row + i] = counts[row + idx[i]];
}

// one `while (prevlh < nextlh)` turn of cafe_main.c:377-386 for the packed families: maxlik[i] belongs to family idx[i]
__global__ void k_lrt_step(const double* __restrict__ maxlik, const int* __restrict__ idx, int n_active, double* __restrict__ prev,
                           int* __restrict__ steps, unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const int f = idx[i];
    const double next = maxlik[i];
    const bool grew = prev[f] < next;
    if (grew) {
        prev[f] = next;
        steps[f]++;
    }
    keep[i] = grew ? 1 : 0;
}

struct LrtBuffers {
    double *d_base = nullptr, *d_prev = nullptr;
    int *d_steps = nullptr, *d_idx = nullptr, *d_counts_c = nullptr;
    unsigned char* d_keep = nullptr;
    ~LrtBuffers() { cudaFree(d_base); cudaFree(d_prev); cudaFree(d_steps); cudaFree(d_idx); cudaFree(d_counts_c); cudaFree(d_keep); }
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
    CAFE_CK(ctx, cudaMalloc(&B.d_idx, F * sizeof(int)));
    CAFE_CK(ctx, cudaMalloc(&B.d_keep, F));
    const size_t counts_ints = (size_t)ctx->n_leaves * ctx->F_pad;
    CAFE_CK(ctx, cudaMalloc(&B.d_counts_c, counts_ints * sizeof(int)));
    CAFE_CK(ctx, cudaMemsetAsync(B.d_counts_c, 0, counts_ints * sizeof(int), ctx->stream));
    const int threads = 256, blocks = (F + threads - 1) / threads;

    // the unlengthened tree: maxlh of cafe_main.c:365
    int rc = launch_prune(ctx, nullptr);
    if (rc) return rc;
    CAFE_CK(ctx, cudaMemcpyAsync(B.d_base, ctx->d_maxlik, F * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (base_out) CAFE_CK(ctx, cudaMemcpyAsync(base_out, B.d_base, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));

    // the thread's tree copy keeps the parsed (double) lengths only until the first tested family has gone over a branch: the
    // length is restored through `int old_bl` (cafe_main.c:350,375,390).  So the first tested family starts from the parsed
    // length, every later one from the truncated length.
    // tested[f] == 2 marks a tested family that is NOT the table's first tested one even if it is the first of this context
    // (a later shard of a table split over ranks): it starts from the truncated lengths like every later family.
    int first_tested = -1;
    for (int f = 0; f < F; ++f)
        if (!tested || tested[f]) {
            if (!tested || tested[f] == 1) first_tested = f;
            break;
        }
    std::vector<int> all_tested;
    for (int f = 0; f < F; ++f)
        if (!tested || tested[f]) all_tested.push_back(f);

    // the tree's own keys and the full family table come back on every way out
    struct KeyGuard {
        cafe_gpu_ctx* ctx;
        const std::vector<BdKey> keys0;
        const std::vector<int> node_key0;
        int* const d_counts0;
        const int F0;
        ~KeyGuard() {
            ctx->keys = keys0;
            ctx->node_key = node_key0;
            ctx->d_counts = d_counts0;
            ctx->F = F0;
            build_schedule(ctx);
            ctx->results_valid = false;  // d_maxlik holds the last lengthened tree, not the tree's own
        }
    } guard{ctx, ctx->keys, ctx->node_key, ctx->d_counts, ctx->F};
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
            // Only the families that are still growing are pruned again: they are packed into a table of their own (a family's
            // result does not depend on its position, tests/test_gpu_parity.py), so a step costs K2 over the survivors only.
            std::vector<int> idx;
            if (variant == 1) idx.push_back(first_tested);
            else for (int f : all_tested) if (!(two_starts && f == first_tested)) idx.push_back(f);
            std::vector<unsigned char> keep;
            double bl = (variant == 0) ? truncated : parsed;
            for (int iter = 0; iter < 100000 && !idx.empty(); ++iter) {
                const int n_active = (int)idx.size();
                bl += rint(bl * 0.15);                                             // cafe_main.c:380
                ctx->keys = keys0;
                // birthdeath_cache_get_matrix, birthdeath.c:363-370
                ctx->keys.push_back(BdKey{(int)bl, ctx->lambda[b], lengthened_mu ? lengthened_mu[b] : ctx->mu[b]});
                ctx->node_key = node_key0;
                ctx->node_key[b] = D;
                rc = build_schedule(ctx);
                if (!rc) rc = build_one_matrix(ctx, D);
                if (rc) return rc;
                CAFE_CK(ctx, cudaMemcpyAsync(B.d_idx, idx.data(), n_active * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                const dim3 ggrid((n_active + threads - 1) / threads, ctx->n_leaves);
                k_lrt_gather_counts<<<ggrid, threads, 0, ctx->stream>>>(guard.d_counts0, ctx->F_pad, B.d_idx, n_active, B.d_counts_c);
                ctx->launches++;
                ctx->d_counts = B.d_counts_c;
                ctx->F = n_active;
                rc = launch_prune(ctx, nullptr);
                ctx->d_counts = guard.d_counts0;
                ctx->F = F;
                if (rc) return rc;
                k_lrt_step<<<(n_active + threads - 1) / threads, threads, 0, ctx->stream>>>(ctx->d_maxlik, B.d_idx, n_active, B.d_prev,
                                                                                              B.d_steps, B.d_keep);
                ctx->launches++;
                keep.resize(n_active);
                CAFE_CK(ctx, cudaMemcpyAsync(keep.data(), B.d_keep, n_active, cudaMemcpyDeviceToHost, ctx->stream));
                CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
                int kept = 0;
                for (int i = 0; i < n_active; ++i)
                    if (keep[i]) idx[kept++] = idx[i];
                idx.resize(kept);
            }
        }
        CAFE_CK(ctx, cudaMemcpyAsync(best_row, B.d_prev, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (steps_row) CAFE_CK(ctx, cudaMemcpyAsync(steps_row, B.d_steps, F * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return CAFE_GPU_OK;
}
