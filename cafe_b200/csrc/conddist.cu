// conddist.cu — K4: Monte-Carlo conditional distribution.
//
// Replaces cafe_conditional_distribution / get_random_probabilities
// (cafe/conditional_distribution.cpp:10-120) and cafe_tree_random_familysize (cafe/cafe_tree.c:533-569):
// for every root size s in [root_min, root_max], n_samples simulated families are drawn down the tree
// (child size = first c with cumsum_{c'<=c} M[parent][c'] >= u, c < maxFamilySize-1), each is pruned with
// root range {s} and the shrinking column window of conditional_distribution.cpp:29, and the n_samples
// root likelihoods of every s are sorted ascending.
//
// Parallel restatement of the reference's serial loops:
//   * the reference's running `cumul +=` sums are reproduced as per-row prefix sums computed serially in
//     the same order (k_row_cdf), so a binary search finds exactly the size the serial scan would;
//   * trials are independent: the simulation uses the maxFamilySize captured before the loop (:20,:26);
//     the range.max ratchet is a prefix-min over the trials of one root size (k_ratchet);
//   * "replay" mode consumes the caller's uniform stream in the reference's order (s, trial, prefix-order
//     non-root node) and is draw-for-draw comparable with the single-threaded reference; otherwise a
//     counter-based generator keyed by (seed, s, trial, node) is used.
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

// prefix sums of every matrix row, in the reference's summation order (cafe_tree.c:549-553)
__global__ void k_row_cdf(const double* __restrict__ M, double* __restrict__ CDF, int S, int Sp, int D) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (row >= S || d >= D) return;
    const double* m = M + ((size_t)d * Sp + row) * Sp;
    double* c = CDF + ((size_t)d * Sp + row) * Sp;
    double cumul = 0.0;
    for (int j = 0; j < S; ++j) { cumul += m[j]; c[j] = cumul; }
}

__device__ __forceinline__ double counter_uniform(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    // splitmix64-style mixing of (seed, root size, trial, node) -> [0,1) with 53 random bits
    uint64_t x = seed ^ (a * 0x9E3779B97F4A7C15ull) ^ (b * 0xBF58476D1CE4E5B9ull) ^ (c * 0x94D049BB133111EBull);
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// one thread per (root size, trial): walk the tree in prefix order (libtree/tree.c:101-124)
__global__ void k_simulate(const double* __restrict__ CDF, int Sp, const int* __restrict__ prefix_nodes,
                           const int* __restrict__ parent, const int* __restrict__ node_key, int n_nonroot, int root,
                           int s_lo, int n_samples, int Fc, int Fc_pad, int range_max,
                           const double* __restrict__ uniforms /* nullable, chunk-local */, uint64_t seed,
                           int* __restrict__ sizes /* [n_nodes][Fc_pad] */, int* __restrict__ trial_max, int* __restrict__ root_size) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Fc) return;
    const int s = s_lo + f / n_samples, trial = f % n_samples;
    root_size[f] = s;
    const int max_family_size = max(s, range_max);  // conditional_distribution.cpp:20
    sizes[(size_t)root * Fc_pad + f] = s;
    int mx = 0;
    for (int k = 0; k < n_nonroot; ++k) {
        const int v = prefix_nodes[k];
        const int ps = sizes[(size_t)parent[v] * Fc_pad + f];
        const double rnd = uniforms ? uniforms[(size_t)f * n_nonroot + k] : counter_uniform(seed, (uint64_t)s, (uint64_t)trial, (uint64_t)k);
        const double* cdf = CDF + ((size_t)node_key[v] * Sp + ps) * Sp;
        // first c in [0, max_family_size-2] with cdf[c] >= rnd, else max_family_size-1
        int lo = 0, hi = max_family_size - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] >= rnd) hi = mid; else lo = mid + 1;
        }
        sizes[(size_t)v * Fc_pad + f] = lo;
        mx = max(mx, lo);
    }
    trial_max[f] = mx;
}

// range.max = MIN(max + MAX(50, max/5), range.max), carried from trial to trial within one root size
__global__ void k_ratchet(const int* __restrict__ trial_max, int n_root_sizes, int n_samples, int range_max, int* __restrict__ colmax) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_root_sizes) return;
    int cap = range_max;
    for (int t = 0; t < n_samples; ++t) {
        const int m = trial_max[(size_t)r * n_samples + t];
        cap = min(m + max(50, m / 5), cap);
        colmax[(size_t)r * n_samples + t] = cap;
    }
}

struct RootChild {
    int is_leaf;
    int key;
    int leaf;           // leaf ordinal when is_leaf
    int slot;           // vector slot when internal
    const int* err_rowptr; const int* err_col; const double* err_val;
};

// root with the single row s: L0 = prod over the two children of (row s of M_child) . (child vector)
// one warp per simulated family
__global__ void __launch_bounds__(256)
k_root_single_row(RootChild c0, RootChild c1, const double* __restrict__ M, const double* __restrict__ vec, int Sp, int Vp,
                  size_t slot_stride, const int* __restrict__ sizes, int Fc, int Fc_pad, int s_lo, int n_samples,
                  const int* __restrict__ colmax, double* __restrict__ L0) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= Fc) return;
    const int s = s_lo + f / n_samples;
    const int cm = colmax[f];
    double prod = 1.0;
    for (int w = 0; w < 2; ++w) {
        const RootChild c = w ? c1 : c0;
        const double* row = M + ((size_t)c.key * Sp + s) * Sp;
        double fac;
        if (c.is_leaf) {
            const int cnt = sizes[(size_t)(2 * c.leaf) * Fc_pad + f];
            if (c.err_rowptr == nullptr) {
                fac = (cnt <= cm) ? row[cnt] : 0.0;
            } else {
                fac = 0.0;
                for (int k = c.err_rowptr[cnt]; k < c.err_rowptr[cnt + 1]; ++k) {
                    const int j = c.err_col[k];
                    if (j <= cm) fac = __dadd_rn(fac, __dmul_rn(row[j], c.err_val[k]));
                }
            }
        } else {
            const double* L = vec + (size_t)c.slot * slot_stride + (size_t)f * Vp;
            double acc = 0.0;
            for (int j = lane; j <= cm; j += 32) acc += row[j] * L[j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            fac = acc;
        }
        prod *= fac;
    }
    if (lane == 0) L0[f] = prod;
}

// The simulated families of a chunk in the fused kernel's own order (see the dealing in run_conditional_distribution):
// position p holds family order[p]; only the leaves' sizes travel (the kernel reads nothing else of the size table).
__global__ void k_deal_chunk(const int* __restrict__ sizes, int Fc_pad, const int* __restrict__ colmax, const int* __restrict__ root_size,
                             const int* __restrict__ order, int Fc, int* __restrict__ leaf_sizes_p, int* __restrict__ colmax_p,
                             int* __restrict__ root_size_p) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (p >= Fc) return;
    const int f = order[p];
    leaf_sizes_p[(size_t)k * Fc_pad + p] = sizes[(size_t)(2 * k) * Fc_pad + f];
    if (k == 0) { colmax_p[p] = colmax[f]; root_size_p[p] = root_size[f]; }
}
__global__ void k_undeal_L0(const double* __restrict__ L0p, const int* __restrict__ order, int Fc, double* __restrict__ L0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < Fc) L0[order[p]] = L0p[p];
}

}  // namespace

// rows [row_lo, row_hi) of the distribution (root sizes root_min + row): cd_out is [(row_hi - row_lo)][n_samples]; the replay
// stream `uniforms` always starts at row 0.  Rows are independent (the device RNG is keyed by (seed, root size, trial, node)), so
// ranks can split them (sharding.conditional_distribution_sharded).
int run_conditional_distribution(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed, int row_lo, int row_hi, double* cd_out) {
    const int R = ctx->R, n = ctx->n_nodes, n_nonroot = n - 1, D = (int)ctx->keys.size();
    const size_t mat_bytes = (size_t)D * ctx->Sp * ctx->Sp * sizeof(double);
    double* d_cdf = nullptr;
    int *d_prefix = nullptr, *d_parent = nullptr, *d_node_key = nullptr;
    int *d_sizes = nullptr, *d_trial_max = nullptr, *d_colmax = nullptr, *d_root_size = nullptr;
    int *d_order = nullptr, *d_leaf_p = nullptr, *d_colmax_p = nullptr, *d_root_p = nullptr;
    double* d_L0p = nullptr;
    int order_Fc = -1;  // the chunk size d_order was built for
    // the fused kernel (prune_fused2.cu, windowed mode) prunes the simulated families whenever it can; the per-node kernels remain
    // for error-model leaves and as the A/B switch CAFE_GPU_NO_FUSED
    const bool fused = fused2_windowed_supported(ctx) && std::getenv("CAFE_GPU_NO_FUSED") == nullptr;
    // CAFE_GPU_STAGE_TIMES=1: wall-clock of the stages of this call on stderr (each stage ends with a stream synchronisation)
    const bool stage_times = std::getenv("CAFE_GPU_STAGE_TIMES") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char* what) {
        if (!stage_times) return;
        cudaStreamSynchronize(ctx->stream);
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "cond_dist stage %-26s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    double *d_uniforms = nullptr, *d_L0 = nullptr, *d_sorted = nullptr;
    void* d_tmp = nullptr;
    int* d_offsets = nullptr;
    int rc = CAFE_GPU_OK;
    auto cleanup = [&]() {
        for (const void* p : {(const void*)d_cdf, (const void*)d_prefix, (const void*)d_parent, (const void*)d_node_key, (const void*)d_sizes,
                              (const void*)d_trial_max, (const void*)d_colmax, (const void*)d_root_size, (const void*)d_uniforms, (const void*)d_L0,
                              (const void*)d_sorted, (const void*)d_tmp, (const void*)d_offsets, (const void*)d_order, (const void*)d_leaf_p,
                              (const void*)d_colmax_p, (const void*)d_root_p, (const void*)d_L0p})
            work_free(ctx, p);
    };
#define CD_CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); cleanup(); return CAFE_GPU_ERR_CUDA; } } while (0)

    // root children (the root node itself is evaluated by k_root_single_row)
    RootChild rc_child[2];
    {
        const int kids[2] = {ctx->left[ctx->root], ctx->right[ctx->root]};
        for (int w = 0; w < 2; ++w) {
            RootChild c{};
            const int v = kids[w];
            c.is_leaf = ctx->left[v] < 0;
            c.key = ctx->node_key[v];
            c.leaf = v / 2;
            c.slot = -1;
            if (c.is_leaf) {
                int e = ctx->leaf_err.empty() ? -1 : ctx->leaf_err[c.leaf];
                if (e >= 0) { c.err_rowptr = ctx->errs[e].d_rowptr; c.err_col = ctx->errs[e].d_col; c.err_val = ctx->errs[e].d_val; }
            } else {
                for (const PruneOp& op : ctx->ops)
                    if (op.is_root && op.gemm_child == v) c.slot = op.in_slot;
                if (c.slot < 0) { ctx->err = "conditional_distribution: schedule has no slot for a root child"; return CAFE_GPU_ERR_STATE; }
            }
            rc_child[w] = c;
        }
    }

    CD_CK(work_malloc(ctx, &d_cdf, mat_bytes));
    {
        dim3 grid((ctx->S + 127) / 128, D);
        k_row_cdf<<<grid, 128, 0, ctx->stream>>>(ctx->d_M, d_cdf, ctx->S, ctx->Sp, D);
        ctx->launches++;
    }
    CD_CK(work_malloc(ctx, &d_prefix, n_nonroot * sizeof(int)));
    CD_CK(work_malloc(ctx, &d_parent, n * sizeof(int)));
    CD_CK(work_malloc(ctx, &d_node_key, n * sizeof(int)));
    CD_CK(cudaMemcpyAsync(d_prefix, ctx->prefix_nonroot.data(), n_nonroot * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(cudaMemcpyAsync(d_parent, ctx->parent.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CD_CK(cudaMemcpyAsync(d_node_key, ctx->node_key.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

    // root sizes are processed in chunks that bound the vector-slot memory
    const int rows_per_chunk = std::max(1, std::min(R, (1 << 16) / std::max(1, n_samples)));
    const int Fc_max = rows_per_chunk * n_samples, Fc_pad = round_up(Fc_max, 128);
    CD_CK(work_malloc(ctx, &d_sizes, (size_t)n * Fc_pad * sizeof(int)));
    CD_CK(cudaMemsetAsync(d_sizes, 0, (size_t)n * Fc_pad * sizeof(int), ctx->stream));
    CD_CK(work_malloc(ctx, &d_trial_max, (size_t)Fc_pad * sizeof(int)));
    CD_CK(work_malloc(ctx, &d_colmax, (size_t)Fc_pad * sizeof(int)));
    CD_CK(cudaMemsetAsync(d_colmax, 0, (size_t)Fc_pad * sizeof(int), ctx->stream));
    CD_CK(work_malloc(ctx, &d_root_size, (size_t)Fc_pad * sizeof(int)));
    CD_CK(cudaMemsetAsync(d_root_size, 0, (size_t)Fc_pad * sizeof(int), ctx->stream));
    CD_CK(work_malloc(ctx, &d_L0, (size_t)R * n_samples * sizeof(double)));
    CD_CK(work_malloc(ctx, &d_sorted, (size_t)R * n_samples * sizeof(double)));
    if (uniforms) CD_CK(work_malloc(ctx, &d_uniforms, (size_t)Fc_max * n_nonroot * sizeof(double)));
    if (!fused) {
        rc = ensure_vec_buffers(ctx, Fc_pad);
        if (rc) { cleanup(); return rc; }
    } else {
        CD_CK(work_malloc(ctx, &d_order, (size_t)Fc_pad * sizeof(int)));
        CD_CK(work_malloc(ctx, &d_leaf_p, (size_t)ctx->n_leaves * Fc_pad * sizeof(int)));
        CD_CK(cudaMemsetAsync(d_leaf_p, 0, (size_t)ctx->n_leaves * Fc_pad * sizeof(int), ctx->stream));
        CD_CK(work_malloc(ctx, &d_colmax_p, (size_t)Fc_pad * sizeof(int)));
        CD_CK(cudaMemsetAsync(d_colmax_p, 0, (size_t)Fc_pad * sizeof(int), ctx->stream));
        CD_CK(work_malloc(ctx, &d_root_p, (size_t)Fc_pad * sizeof(int)));
        CD_CK(cudaMemsetAsync(d_root_p, 0, (size_t)Fc_pad * sizeof(int), ctx->stream));
        CD_CK(work_malloc(ctx, &d_L0p, (size_t)Fc_pad * sizeof(double)));
    }
    const size_t slot_stride = (size_t)Fc_pad * ctx->Vp;
    stage("allocations + row CDFs");

    for (int r_lo = row_lo; r_lo < row_hi; r_lo += rows_per_chunk) {
        const int rows = std::min(rows_per_chunk, row_hi - r_lo), Fc = rows * n_samples, s_lo = ctx->root_min + r_lo;
        if (uniforms)
            CD_CK(cudaMemcpyAsync(d_uniforms, uniforms + (size_t)r_lo * n_samples * n_nonroot, (size_t)Fc * n_nonroot * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
        k_simulate<<<(Fc + 127) / 128, 128, 0, ctx->stream>>>(d_cdf, ctx->Sp, d_prefix, d_parent, d_node_key, n_nonroot, ctx->root, s_lo,
                                                              n_samples, Fc, Fc_pad, ctx->rmax, d_uniforms, seed, d_sizes, d_trial_max, d_root_size);
        k_ratchet<<<(rows + 127) / 128, 128, 0, ctx->stream>>>(d_trial_max, rows, n_samples, ctx->rmax, d_colmax);
        ctx->launches += 2;
        if (fused) {
            // one persistent launch: every node with the per-trial column window, the root over its whole range, and of the
            // root's likelihoods the one of the family's own root size (the reference prunes with the root range {s})
            //
            // The kernel stops a tile's K loops and output passes at the tile's largest window, and it splits the families over
            // its CTAs statically, so it gets them in an order of its own: largest root size first (the window grows with the root
            // size and hardly changes along the draws of one root size), dealt to the launch's tiles round by round - first tile
            // of every CTA, then the second of every CTA, ...  A tile of 96 then holds draws of one or two root sizes, and every CTA
            // the same mix of wide and narrow tiles; in row order the CTAs at the end of a chunk would hold all the wide ones.
            if (order_Fc != Fc) {
                std::vector<std::array<int, 3>> slots;
                fused2_tile_slots(ctx, Fc, slots);
                std::stable_sort(slots.begin(), slots.end(), [](const std::array<int, 3>& a, const std::array<int, 3>& b) { return a[0] < b[0]; });
                std::vector<int> order(Fc_pad, 0);
                int next = Fc - 1;
                for (const auto& sl : slots)
                    for (int i = 0; i < sl[2]; ++i) order[sl[1] + i] = next--;
                if (next != -1) { ctx->err = "conditional_distribution: the tile slots do not cover the chunk"; cleanup(); return CAFE_GPU_ERR_STATE; }
                CD_CK(cudaMemcpyAsync(d_order, order.data(), (size_t)Fc_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                CD_CK(cudaStreamSynchronize(ctx->stream));  // `order` is a local
                order_Fc = Fc;
            }
            k_deal_chunk<<<dim3((Fc + 255) / 256, ctx->n_leaves), 256, 0, ctx->stream>>>(d_sizes, Fc_pad, d_colmax, d_root_size, d_order, Fc,
                                                                                         d_leaf_p, d_colmax_p, d_root_p);
            Fused2Job job;
            job.counts = d_leaf_p; job.leaf_stride = (size_t)Fc_pad; job.F = Fc; job.F_pad = Fc_pad;
            job.d_colmax = d_colmax_p;
            job.root_r0 = ctx->root_min; job.root_rows = R;
            job.d_root_pick = d_root_p; job.d_L0_out = d_L0p;
            rc = launch_prune_fused2_job(ctx, job);
            if (rc) { cleanup(); return rc; }
            k_undeal_L0<<<(Fc + 255) / 256, 256, 0, ctx->stream>>>(d_L0p, d_order, Fc, d_L0 + (size_t)r_lo * n_samples);
            ctx->launches += 2;
        } else {
            // all non-root nodes with the per-trial column window; leaf k is node 2k of the size table
            rc = launch_prune_ops(ctx, d_sizes, (size_t)2 * Fc_pad, Fc, Fc_pad, d_colmax, 0, 0, /*skip_root=*/true, nullptr);
            if (rc) { cleanup(); return rc; }
            k_root_single_row<<<(Fc + 7) / 8, 256, 0, ctx->stream>>>(rc_child[0], rc_child[1], ctx->d_M, ctx->d_vec, ctx->Sp, ctx->Vp, slot_stride,
                                                                     d_sizes, Fc, Fc_pad, s_lo, n_samples, d_colmax,
                                                                     d_L0 + (size_t)r_lo * n_samples);
            ctx->launches++;
        }
        CD_CK(cudaGetLastError());
        stage("chunk (simulate + prune)");
    }

    // ascending sort of every row (std::sort, conditional_distribution.cpp:41)
    {
        std::vector<int> off(R + 1);
        for (int r = 0; r <= R; ++r) off[r] = (row_lo + std::min(r, row_hi - row_lo)) * n_samples;  // segments of the requested rows only
        CD_CK(work_malloc(ctx, &d_offsets, (R + 1) * sizeof(int)));
        CD_CK(cudaMemcpyAsync(d_offsets, off.data(), (R + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        size_t tmp_bytes = 0;
        CD_CK(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, d_L0, d_sorted, R * n_samples, row_hi - row_lo, d_offsets, d_offsets + 1, ctx->stream));
        CD_CK(work_malloc(ctx, &d_tmp, std::max<size_t>(tmp_bytes, 16)));
        CD_CK(cub::DeviceSegmentedSort::SortKeys(d_tmp, tmp_bytes, d_L0, d_sorted, R * n_samples, row_hi - row_lo, d_offsets, d_offsets + 1, ctx->stream));
        ctx->launches++;
    }
    CD_CK(cudaMemcpyAsync(cd_out, d_sorted + (size_t)row_lo * n_samples, (size_t)(row_hi - row_lo) * n_samples * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CD_CK(cudaStreamSynchronize(ctx->stream));
    stage("sort + download");
    cleanup();
    stage("frees");
    ctx->results_valid = false;  // the vector slots were reused
#undef CD_CK
    return CAFE_GPU_OK;
}
