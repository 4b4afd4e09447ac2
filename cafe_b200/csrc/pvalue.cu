// pvalue.cu — K5: family-wide p-values.
//
// Replaces, for every family at once, viterbi_section's p-value part (cafe/viterbi.cpp:88-97):
//   cafe_family_set_size_with_family_forced (cafe/cafe_family.c:236-255): per-family range
//       root 1..rint(1.25*max_f), columns 0..max_f + max(50, max_f/5)
//   cafe_tree_p_values (cafe/pvalue.cpp:143-154): prune, p[s] = pvalue(L[s], cd[s], n_samples)
//   pvalue (libcommon/mathfunc.c:663-689): rank of L[s] in the ascending row, ties at their midpoint
//   viterbi_set_max_pvalue (cafe/viterbi.cpp:32-39): max over s, 0 for an empty root range
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace {

// Ties: the reference compares with == (mathfunc.c:674-681) and an observed family that coincides with a simulated
// one gives a bit-identical likelihood there, so the tie midpoint matters.  Here the two values come from
// differently ordered sums (GEMM vs row dot product) and agree only to ~1e-15 relative; values within
// TIE_RTOL of each other are therefore treated as the tie the reference would see.
constexpr double TIE_RTOL = 1e-11;

__device__ __forceinline__ double pvalue_dev(double v, const double* __restrict__ cd, int size) {
    const double lo_v = v * (1.0 - TIE_RTOL), hi_v = v * (1.0 + TIE_RTOL);
    int from = 0, to = size - 1;
    while (from < to) {
        const int mi = from + (to - from) / 2;
        const double c = cd[mi];
        if (c > hi_v) to = mi - 1;
        else if (c < lo_v) from = mi + 1;
        else {
            from = mi; while (from > 0 && cd[from - 1] >= lo_v) --from;
            to = mi; while (to + 1 < size && cd[to + 1] <= hi_v) ++to;
            break;
        }
    }
    if (from > to) to = from;
    return ((double)from + (cd[from] <= hi_v ? 1.0 : 0.0) + (double)(to - from) / 2.0) / (double)size;
}

// one warp per family: lanes take root sizes s, warp-max of the p-values
// order (nullable): row f of Lroot / rfsize belongs to family order[f] of the table (run_pvalues sorts the families by window)
__global__ void __launch_bounds__(256)
k_family_pvalue(const double* __restrict__ Lroot, size_t Vp, int F, const int* __restrict__ rfsize, const double* __restrict__ cd,
                int cd_rows, int n_samples, const int* __restrict__ order, double* __restrict__ out) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= F) return;
    const int rf = min(rfsize[f], cd_rows);
    double best = -1.0;
    for (int s = lane; s < rf; s += 32) best = fmax(best, pvalue_dev(Lroot[(size_t)f * Vp + s], cd + (size_t)s * n_samples, n_samples));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) out[order ? order[f] : f] = (rf > 0) ? best : 0.0;
}

// counts of the families in another order: dst[k][i] = src[k][order[i]] for every leaf k (leaf-major tables)
__global__ void k_permute_counts(const int* __restrict__ src, int* __restrict__ dst, const int* __restrict__ order, int F, int F_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i < F) dst[(size_t)k * F_pad + i] = src[(size_t)k * F_pad + order[i]];
}

// ---------------------------------------------------------------- branch cutting (cafe/branch_cutting.cpp:20-44, :101-150)
// One side of the cut is a single leaf: cutPvalue = max_s pvalue(L[s], cd[s]) over the whole root range (cafe_tree_p_values on
// the other tree, :125-131).  One warp per family.
__global__ void __launch_bounds__(256)
k_cut_pvalue_one(const double* __restrict__ L, int F, int rf, const double* __restrict__ cd, int cdlen, double* __restrict__ out) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= F) return;
    double best = -INFINITY;
    for (int s = lane; s < rf; s += 32) best = fmax(best, pvalue_dev(L[(size_t)f * rf + s], cd + (size_t)s * cdlen, cdlen));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) out[f] = best;
}

// Both sides are trees (p_values_of_two_trees, :20-44): p2[s1][s2] = (1/cdlen) sum_t pvalue(L1[s1] L2[s2] / cd2[s2][t], cd1[s1]),
// cutPvalue = max over (s1, s2), never below 0 (:136-147).  One thread per (s1, s2) pair, the sum over t in the reference's order
// (one rounding per product, quotient and addition), block maximum, one atomic maximum per block: p >= 0, so the bit patterns
// order like the values.  grid = (pair chunks, families).
__global__ void __launch_bounds__(256)
k_cut_pvalue_two(const double* __restrict__ L1, const double* __restrict__ L2, int rf, const double* __restrict__ cd1,
                 const double* __restrict__ cd2, int cdlen, unsigned long long* __restrict__ out_bits) {
    const int f = blockIdx.y;
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    double p = 0.0;
    if (pair < rf * rf) {
        const int s2 = pair / rf, s1 = pair - s2 * rf;
        const double prod = __dmul_rn(L1[(size_t)f * rf + s1], L2[(size_t)f * rf + s2]);
        const double* row1 = cd1 + (size_t)s1 * cdlen;
        const double* row2 = cd2 + (size_t)s2 * cdlen;
        for (int t = 0; t < cdlen; ++t) p = __dadd_rn(p, pvalue_dev(__ddiv_rn(prod, row2[t]), row1, cdlen));
        p = __ddiv_rn(p, (double)cdlen);
    }
    if (!(p > 0.0)) p = 0.0;  // the running maximum starts at 0 (:136); a NaN never replaces it (:142)
    __shared__ double red[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = fmax(p, __shfl_xor_sync(0xffffffffu, p, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) p = fmax(p, red[w]);
        atomicMax(out_bits + f, (unsigned long long)__double_as_longlong(p));
    }
}

}  // namespace

int run_pvalues(cafe_gpu_ctx* ctx, const double* cd, int cd_rows, int n_samples, double* out) {
    const int F = ctx->F, nl = ctx->n_leaves;
    // CAFE_GPU_STAGE_TIMES=1: wall-clock of the stages of this call on stderr (each stage ends with a stream synchronisation)
    const bool stage_times = std::getenv("CAFE_GPU_STAGE_TIMES") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_last = now();
    auto stage = [&](const char* what) {
        if (!stage_times) return;
        cudaStreamSynchronize(ctx->stream);
        const auto t = now();
        std::fprintf(stderr, "pvalues stage %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    // per-family forced range (host: n_leaves ints per family)
    std::vector<int> colmax(ctx->F_pad, 0), rfsize(ctx->F_pad, 0);
    int rf_max = 0;
    for (int f = 0; f < F; ++f) {
        const int mx = ctx->h_fam_max[f];
        const int root_max = (int)std::rint(mx * 1.25);
        colmax[f] = std::min(mx + std::max(50, mx / 5), ctx->W - 1);
        rfsize[f] = root_max;  // root_min is 1 in the forced range
        rf_max = std::max(rf_max, root_max);
    }
    stage("forced ranges (host)");
    if (rf_max > ctx->Vp || 1 + rf_max > ctx->S)
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "pvalues: a family's root range rint(1.25*max) exceeds the matrices (set_ranges from the table's max first)");
    int *d_colmax = nullptr, *d_rf = nullptr, *d_order = nullptr, *d_counts_sorted = nullptr;
    double *d_cd = nullptr, *d_out = nullptr;
    auto cleanup = [&]() {  // work_free: the buffers stay with the context (common.cuh)
        for (const void* q : {(const void*)d_colmax, (const void*)d_rf, (const void*)d_cd, (const void*)d_out, (const void*)d_order, (const void*)d_counts_sorted}) work_free(ctx, q);
    };
    // root rows 1..rf_max for everybody (rows beyond a family's own range are ignored by k_family_pvalue)
    const int root_rows = std::min(rf_max, ctx->S - 1);
    const bool fused = root_rows >= 1 && fused2_windowed_supported(ctx) && std::getenv("CAFE_GPU_NO_FUSED") == nullptr;
    // The fused kernel ends a tile's K loops and output sizes at the tile's largest window (prune_fused2.cu, tile_wmax), so it
    // gets the families in an order of its own: sorted by window (a counting sort, ties in table order), widest first, and dealt
    // to the launch's tiles round by round - first tile of every CTA, then the second of every CTA, ... - so that a tile of 96
    // holds families of nearly the same range AND every CTA gets the same mix of wide and narrow tiles (sorted order alone would
    // hand all the wide families to the last few CTAs of the static split).  A family's result does not depend on its
    // position; the p-values go back to the table's order.
    std::vector<int> order;
    if (fused && F > 1) {
        std::vector<int> start(ctx->W + 1, 0), sorted(F);
        for (int f = 0; f < F; ++f) start[colmax[f] + 1]++;
        for (int w = 0; w < ctx->W; ++w) start[w + 1] += start[w];
        for (int f = 0; f < F; ++f) sorted[start[colmax[f]]++] = f;          // ascending window
        std::vector<std::array<int, 3>> slots;
        fused2_tile_slots(ctx, F, slots);
        std::stable_sort(slots.begin(), slots.end(), [](const std::array<int, 3>& a, const std::array<int, 3>& b) { return a[0] < b[0]; });
        order.assign(ctx->F_pad, 0);
        int next = F - 1;                                                       // widest first
        for (const auto& sl : slots)
            for (int i = 0; i < sl[2]; ++i) order[sl[1] + i] = sorted[next--];
        if (next != -1) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "pvalues: the tile slots do not cover the families");
        std::vector<int> cm(colmax), rf(rfsize);
        for (int i = 0; i < F; ++i) { colmax[i] = cm[order[i]]; rfsize[i] = rf[order[i]]; }
    }
    stage("family order (host)");
#define PV_CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); cleanup(); return CAFE_GPU_ERR_CUDA; } } while (0)
    PV_CK(work_malloc(ctx, &d_colmax, ctx->F_pad * sizeof(int)));
    PV_CK(work_malloc(ctx, &d_rf, ctx->F_pad * sizeof(int)));
    PV_CK(work_malloc(ctx, &d_cd, (size_t)cd_rows * n_samples * sizeof(double)));
    PV_CK(work_malloc(ctx, &d_out, ctx->F_pad * sizeof(double)));
    PV_CK(cudaMemcpyAsync(d_colmax, colmax.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PV_CK(cudaMemcpyAsync(d_rf, rfsize.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PV_CK(cudaMemcpyAsync(d_cd, cd, (size_t)cd_rows * n_samples * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    stage("allocations + uploads");
    const double* Lroot = nullptr;
    size_t Lstride = 0;
    int rc = CAFE_GPU_OK;
    if (fused) {
        // the fused kernel in windowed mode: every family with its own forced range, all root rows copied out
        const size_t need = (size_t)F * root_rows;
        if (need > ctx->Lroot_cache_cap) {  // kept across calls: a malloc / free pair of this size costs more than the pass itself
            cudaFree(ctx->d_Lroot_cache); ctx->d_Lroot_cache = nullptr; ctx->Lroot_cache_cap = 0;
            PV_CK(cudaMalloc(&ctx->d_Lroot_cache, need * sizeof(double)));
            ctx->Lroot_cache_cap = need;
        }
        double* d_Lroot = ctx->d_Lroot_cache;
        const int* d_counts_job = ctx->d_counts;
        if (!order.empty()) {
            PV_CK(work_malloc(ctx, &d_order, ctx->F_pad * sizeof(int)));
            PV_CK(work_malloc(ctx, &d_counts_sorted, (size_t)nl * ctx->F_pad * sizeof(int)));
            PV_CK(cudaMemsetAsync(d_counts_sorted, 0, (size_t)nl * ctx->F_pad * sizeof(int), ctx->stream));
            PV_CK(cudaMemcpyAsync(d_order, order.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            k_permute_counts<<<dim3((F + 255) / 256, nl), 256, 0, ctx->stream>>>(ctx->d_counts, d_counts_sorted, d_order, F, ctx->F_pad);
            ctx->launches++;
            d_counts_job = d_counts_sorted;
        }
        Fused2Job job;
        job.counts = d_counts_job; job.leaf_stride = (size_t)ctx->F_pad; job.F = F; job.F_pad = ctx->F_pad;
        job.d_colmax = d_colmax;
        job.root_r0 = 1; job.root_rows = root_rows;
        job.d_Lroot_out = d_Lroot;
        job.d_root_need = d_rf;  // k_family_pvalue reads the rows s < rfsize[f] only
        rc = launch_prune_fused2_job(ctx, job);
        if (rc) { cleanup(); return rc; }
        Lroot = d_Lroot; Lstride = (size_t)root_rows;
    } else {
        rc = ensure_vec_buffers(ctx, ctx->F_pad);
        if (rc) { cleanup(); return rc; }
        int root_slot = -1;
        rc = launch_prune_ops(ctx, ctx->d_counts, ctx->F_pad, F, ctx->F_pad, d_colmax, 1, root_rows, false, &root_slot);
        if (rc) { cleanup(); return rc; }
        Lroot = ctx->d_vec + (size_t)root_slot * ctx->F_pad * ctx->Vp; Lstride = (size_t)ctx->Vp;
    }
    stage("pruning");
    k_family_pvalue<<<(F + 7) / 8, 256, 0, ctx->stream>>>(Lroot, Lstride, F, d_rf, d_cd, cd_rows, n_samples, d_order, d_out);
    ctx->launches++;
    PV_CK(cudaGetLastError());
    PV_CK(cudaMemcpyAsync(out, d_out, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CK(cudaStreamSynchronize(ctx->stream));
    stage("lookup + download");
    cleanup();
    stage("frees");
    ctx->results_valid = false;
#undef PV_CK
    return CAFE_GPU_OK;
}

// cafe_gpu_cut_pvalues: see include/cafe_gpu.h.  Host arrays in, host array out; the arithmetic of one family does not depend on
// the others.
int run_cut_pvalues(cafe_gpu_ctx* ctx, const double* L1, const double* L2, int F, int rf, const double* cd1, const double* cd2,
                    int cdlen, double* out) {
    if (F <= 0) return CAFE_GPU_OK;
    double *dL1 = nullptr, *dL2 = nullptr, *dcd1 = nullptr, *dcd2 = nullptr, *dout = nullptr;
    auto cleanup = [&]() { cudaFree(dL1); cudaFree(dL2); cudaFree(dcd1); cudaFree(dcd2); cudaFree(dout); };
#define CUT_CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); cleanup(); return CAFE_GPU_ERR_CUDA; } } while (0)
    const size_t lb = (size_t)F * rf * sizeof(double), cb = (size_t)rf * cdlen * sizeof(double);
    CUT_CK(cudaMalloc(&dL1, lb)); CUT_CK(cudaMalloc(&dcd1, cb)); CUT_CK(cudaMalloc(&dout, F * sizeof(double)));
    CUT_CK(cudaMemcpyAsync(dL1, L1, lb, cudaMemcpyHostToDevice, ctx->stream));
    CUT_CK(cudaMemcpyAsync(dcd1, cd1, cb, cudaMemcpyHostToDevice, ctx->stream));
    if (L2) {
        CUT_CK(cudaMalloc(&dL2, lb)); CUT_CK(cudaMalloc(&dcd2, cb));
        CUT_CK(cudaMemcpyAsync(dL2, L2, lb, cudaMemcpyHostToDevice, ctx->stream));
        CUT_CK(cudaMemcpyAsync(dcd2, cd2, cb, cudaMemcpyHostToDevice, ctx->stream));
        CUT_CK(cudaMemsetAsync(dout, 0, F * sizeof(double), ctx->stream));  // +0.0: the maximum starts at 0
        const int pairs = rf * rf;
        for (int f0 = 0; f0 < F; f0 += 32768) {  // gridDim.y limit
            const int nf = std::min(32768, F - f0);
            k_cut_pvalue_two<<<dim3((pairs + 255) / 256, nf), 256, 0, ctx->stream>>>(dL1 + (size_t)f0 * rf, dL2 + (size_t)f0 * rf, rf, dcd1, dcd2, cdlen,
                                                                                    reinterpret_cast<unsigned long long*>(dout) + f0);
            ctx->launches++;
        }
    } else {
        k_cut_pvalue_one<<<(F + 7) / 8, 256, 0, ctx->stream>>>(dL1, F, rf, dcd1, cdlen, dout);
        ctx->launches++;
    }
    CUT_CK(cudaGetLastError());
    CUT_CK(cudaMemcpyAsync(out, dout, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUT_CK(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef CUT_CK
    return CAFE_GPU_OK;
}
