// viterbi.cu — Viterbi ancestral reconstruction for every family.
//
// Replaces cafe_tree_viterbi (cafe/viterbi.cpp:494-521) as driven per family by the report / viterbi commands:
// the max-product pruning of __cafe_tree_node_compute_viterbi (:209-321) in post-order, then the back-track of
// __cafe_tree_node_backtrack_viterbi (:323-351) in prefix order.  Same data flow as K2 with (max, argmax) in place of the sum:
//     factor_c[i] = max_j M_c[r0+i][j] * L_c[j]      (strict ">" from 0: the first maximum wins, an all-zero row keeps pointer 0)
//     L_v[i]      = factor_left[i] * factor_right[i]
// A (max, x) semiring product has no tensor-core form; the kernel keeps the K2 layout instead (transposed matrices, so that the
// threads of a block - consecutive output sizes - read consecutive addresses, 8 families per block to reuse every matrix element
// from registers) and is bound by the fp64 pipe (one DMUL + one compare per matrix element and family).
// A leaf with count -1 carries no data (the reference's "familysize < 0" branch, viterbi.cpp:236-250): its vector is all ones over
// its PARENT's range (the root range below the root) - zeros beyond, as in a freshly allocated tree - and its size is
// reconstructed like an ancestor's.  (In the reference the entries beyond keep whatever an earlier family left there.)
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int VT_THREADS = 128;  // output sizes per block
constexpr int VT_FB = 8;         // families per block

struct VitChild {
    int is_leaf;
    const double* MT;        // transposed matrix of the child's branch
    const int* counts;       // leaf: observed sizes [F_pad] of this leaf (+ family offset)
    const int* err_rowptr;   // leaf with error model: sparse rows, else nullptr
    const int* err_col;
    const double* err_val;
    const double* L;         // internal child: its vector [FC][Vp]
    short* vit;              // back-pointers of the child [FC][Vp]
};

// one internal node: both children, FB families x VT_THREADS output sizes per block
__global__ void __launch_bounds__(VT_THREADS)
k_viterbi_node(VitChild A, VitChild B, int Sp, int Vp, int W, int r0, int nrows, int n_fam, double* __restrict__ Lout,
               const int* __restrict__ colmax, const int* __restrict__ rfsize, int is_root) {
    extern __shared__ double sL[];  // [VT_FB][W] child vector of the families of this block
    const int i = blockIdx.x * VT_THREADS + threadIdx.x;  // output size index
    const int f0 = blockIdx.y * VT_FB;
    const int nf = min(VT_FB, n_fam - f0);
    double prod[VT_FB];
    int cm[VT_FB], nr[VT_FB];  // per family: last column of the window, number of output rows (forced ranges of the report)
#pragma unroll
    for (int u = 0; u < VT_FB; ++u) {
        prod[u] = 1.0;
        cm[u] = (colmax && u < nf) ? colmax[f0 + u] : W - 1;
        nr[u] = !colmax ? nrows : (u < nf ? (is_root ? rfsize[f0 + u] : cm[u] + 1) : 0);
    }
    // forced ranges: nothing of this block's families lies above the widest window / beyond the longest range among them - the
    // column loop ends there, and an output size no family has is a zero without looking (same values as the full loops)
    int cm_blk = 0, nr_blk = 0;
#pragma unroll
    for (int u = 0; u < VT_FB; ++u) { if (u < nf) cm_blk = max(cm_blk, cm[u]); nr_blk = max(nr_blk, nr[u]); }
    const int Wb = min(W, cm_blk + 1);
    const bool live = i < nrows && i < nr_blk;

    for (int side = 0; side < 2; ++side) {
        const VitChild& C = side ? B : A;
        double best[VT_FB]; int arg[VT_FB];
#pragma unroll
        for (int u = 0; u < VT_FB; ++u) { best[u] = 0.0; arg[u] = 0; }
        if (C.is_leaf) {
            if (live) {
                for (int u = 0; u < nf; ++u) {
                    const int cnt = C.counts[f0 + u];
                    if (cnt < 0) {
                        // no data: L[j] = 1 for the sizes of this node's range (:236-250), so the factor is the row maximum
                        const int last = is_root ? min(cm[u], nr[u] - 1) : cm[u];
                        for (int j = 0; j <= last; ++j) {
                            const double v = C.MT[(size_t)j * Sp + r0 + i];
                            if (v > best[u]) { best[u] = v; arg[u] = j; }
                        }
                    } else if (C.err_rowptr == nullptr) {
                        // one-hot leaf (:262-266): the only non-zero product is M[s][count]
                        const double v = (cnt <= cm[u]) ? C.MT[(size_t)cnt * Sp + r0 + i] : 0.0;
                        if (v > 0.0) { best[u] = v; arg[u] = cnt; }
                    } else {
                        // error-model leaf (:252-260): L[j] = errormatrix[count][j], ascending j
                        for (int k = C.err_rowptr[cnt]; k < C.err_rowptr[cnt + 1]; ++k) {
                            const int j = C.err_col[k];
                            if (j <= cm[u]) {
                                const double v = __dmul_rn(C.MT[(size_t)j * Sp + r0 + i], C.err_val[k]);
                                if (v > best[u]) { best[u] = v; arg[u] = j; }
                            }
                        }
                    }
                }
            }
        } else {
            __syncthreads();
            for (int x = threadIdx.x; x < VT_FB * Wb; x += VT_THREADS) {
                const int u = x / Wb, j = x - u * Wb;
                sL[u * W + j] = (u < nf) ? C.L[(size_t)(f0 + u) * Vp + j] : 0.0;
            }
            __syncthreads();
            if (live) {
                const double* __restrict__ mcol = C.MT + r0 + i;
                for (int j = 0; j < Wb; ++j) {
                    const double m = mcol[(size_t)j * Sp];
#pragma unroll
                    for (int u = 0; u < VT_FB; ++u) {
                        const double v = __dmul_rn(m, sL[u * W + j]);
                        if (v > best[u] && j <= cm[u]) { best[u] = v; arg[u] = j; }
                    }
                }
            }
        }
        if (i < nrows) {
#pragma unroll
            for (int u = 0; u < VT_FB; ++u) {
                if (u < nf && i < nr[u]) C.vit[(size_t)(f0 + u) * Vp + i] = (short)arg[u];
                prod[u] = __dmul_rn(prod[u], best[u]);
            }
        }
    }
    if (i < Vp) {
#pragma unroll
        for (int u = 0; u < VT_FB; ++u)
            if (u < nf) Lout[(size_t)(f0 + u) * Vp + i] = (i < nr[u]) ? prod[u] : 0.0;
    }
}

// back-track: one thread per family, prefix order (parent before child)
__global__ void __launch_bounds__(128)
k_viterbi_backtrack(const int* __restrict__ prefix, int n_prefix, const int* __restrict__ parent, const int* __restrict__ is_leaf,
                    const int* __restrict__ leaf_ord, int root, const double* __restrict__ Lroot, const short* __restrict__ vit,
                    size_t node_stride, int Vp, int R, int root_min, int range_min, const int* __restrict__ counts, int F_pad,
                    int fam0, int n_fam, int n_nodes, int* __restrict__ sizes_out, double* __restrict__ maxlik_out,
                    const int* __restrict__ rfsize) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fam) return;
    if (rfsize) R = rfsize[f];
    int* sz = sizes_out + (size_t)(fam0 + f) * n_nodes;
    for (int p = 0; p < n_prefix; ++p) {
        const int v = prefix[p];
        if (is_leaf[v]) {
            const int c = counts[(size_t)leaf_ord[v] * F_pad + fam0 + f];
            if (c >= 0) { sz[v] = c; continue; }  // observed leaves keep their size (:327); one without data is reconstructed
        }
        if (v == root) {
            const double* L = Lroot + (size_t)f * Vp;
            double ml = (R > 0) ? L[0] : 0.0; int am = 0;
            for (int i = 1; i < R; ++i) if (L[i] > ml) { ml = L[i]; am = i; }  // __maxidx: first maximum
            sz[v] = root_min + am;
            if (maxlik_out) maxlik_out[fam0 + f] = ml;
        } else {
            const int par = parent[v];
            const int base = (par == root) ? root_min : range_min;
            // A family with an empty root range (all counts 0: root 1..rint(1.25*0), viterbi.cpp / cafe_family.c:236-255) never
            // writes the back-pointers of the root's children (:289-303 loops over no root size); the reference then reads
            // whatever the previous family left there, 0 in a freshly allocated tree.  We return that 0.
            if (par == root && R <= 0) { sz[v] = range_min; continue; }
            sz[v] = vit[(size_t)v * node_stride + (size_t)f * Vp + (sz[par] - base)] + range_min;
        }
    }
}

// viterbi_sum_probabilities (cafe/viterbi.cpp:42-70): one warp per (family, non-root node): the row of the branch's matrix at the
// parent's reconstructed size; entries equal to the realised transition count half, smaller ones fully.
__global__ void __launch_bounds__(256)
k_viterbi_branch_pvalues(const double* __restrict__ M, const int* __restrict__ node_key, const int* __restrict__ parent, int Sp,
                         int n_nodes, int n_fam_total, const int* __restrict__ sizes, const int* __restrict__ colmax, int W,
                         double* __restrict__ out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)n_fam_total * n_nodes) return;
    const int f = (int)(w / n_nodes), c = (int)(w - (long long)f * n_nodes);
    const int par = parent[c];
    if (par < 0) { if (lane == 0) out[w] = -1.0; return; }
    const int* sz = sizes + (size_t)f * n_nodes;
    const double* __restrict__ row = M + (size_t)node_key[c] * Sp * Sp + (size_t)sz[par] * Sp;
    const double p = row[sz[c]];
    const int cmax = colmax ? colmax[f] : W - 1;
    double acc = 0.0;
    for (int m = lane; m <= cmax; m += 32) {
        const double x = row[m];
        if (x == p) acc += x / 2.0; else if (x < p) acc += x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[w] = acc;
}

// the families in another order (run_viterbi sorts them by forced range): leaf-major count tables in, per-family rows out
__global__ void k_vit_gather_counts(const int* __restrict__ src, int* __restrict__ dst, const int* __restrict__ order, int F, int F_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i < F) dst[(size_t)k * F_pad + i] = src[(size_t)k * F_pad + order[i]];
}
template <typename T>
__global__ void k_vit_scatter_rows(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ order, int F, int n) {
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (long long)F * n) return;
    const int i = (int)(x / n), v = (int)(x - (long long)i * n);
    dst[(size_t)order[i] * n + v] = src[x];
}

}  // namespace

// forced: per-family ranges as viterbi_section uses them (cafe_family_set_size_with_family_forced, cafe/cafe_family.c:236-255):
// root 1..rint(1.25*max_f), columns 0..max_f + max(50, max_f/5).  branch_pv_out (nullable): [F][n_nodes], see k_viterbi_branch_pvalues.
int run_viterbi(cafe_gpu_ctx* ctx, int32_t* sizes_out, double* maxlik_out, bool forced, double* branch_pv_out) {
    const int n = ctx->n_nodes, F = ctx->F, Vp = ctx->Vp, W = ctx->W, Sp = ctx->Sp;
    std::vector<int> h_colmax, h_rf;
    if (forced) {
        const int nl = ctx->n_leaves;
        h_colmax.assign(ctx->F_pad, 0); h_rf.assign(ctx->F_pad, 0);
        for (int f = 0; f < F; ++f) {
            const int mx = ctx->h_fam_max[f];
            h_colmax[f] = std::min(mx + std::max(50, mx / 5), W - 1);
            h_rf[f] = (int)std::rint(mx * 1.25);
            if (h_rf[f] > Vp || 1 + h_rf[f] > ctx->S)
                CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "viterbi: a family's root range rint(1.25*max) exceeds the matrices (set_ranges from the table's max first)");
        }
    }
    // Forced ranges: k_viterbi_node ends its loops at the widest range among the VT_FB families of a block, so the families are
    // processed sorted by range (a counting sort, ties in table order) and the results scattered back to the table's order at
    // the end; a family's reconstruction does not depend on its position.
    std::vector<int> order;
    if (forced && F > VT_FB) {
        std::vector<int> start(W + 1, 0);
        for (int f = 0; f < F; ++f) start[h_colmax[f] + 1]++;
        for (int w = 0; w < W; ++w) start[w + 1] += start[w];
        order.assign(ctx->F_pad, 0);
        for (int f = 0; f < F; ++f) order[start[h_colmax[f]]++] = f;
        std::vector<int> cm(h_colmax), rf(h_rf);
        for (int i = 0; i < F; ++i) { h_colmax[i] = cm[order[i]]; h_rf[i] = rf[order[i]]; }
    }
    const bool sorted = !order.empty();
    if (W > 32767) CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "viterbi: vector longer than the 16-bit back-pointers");
    const size_t mat = (size_t)Sp * Sp;
    // families per chunk: vectors (8 B) of the internal nodes + back-pointers (2 B) of all nodes, <= ~1.5 GB
    const int n_internal = n / 2;
    size_t per_family = (size_t)Vp * (8 * (size_t)n_internal + 2 * (size_t)n);
    int FC = (int)std::max<size_t>(VT_FB, std::min<size_t>((size_t)F, (size_t)(1500u << 20) / per_family));
    FC = (FC + VT_FB - 1) / VT_FB * VT_FB;

    // prefix order with the root first; parents, leaf flags
    std::vector<int> prefix, parent(n, -1), is_leaf(n, 0), leaf_ord(n, 0), slot_of(n, -1);
    {
        std::vector<int> st{ctx->root};
        while (!st.empty()) {
            int v = st.back(); st.pop_back();
            prefix.push_back(v);
            if (ctx->left[v] >= 0) { st.push_back(ctx->right[v]); st.push_back(ctx->left[v]); }
        }
        int s = 0;
        for (int v = 0; v < n; ++v) {
            if (ctx->left[v] >= 0) { parent[ctx->left[v]] = v; parent[ctx->right[v]] = v; slot_of[v] = s++; }
            else { is_leaf[v] = 1; leaf_ord[v] = v / 2; }
        }
    }
    int *d_prefix = nullptr, *d_parent = nullptr, *d_is_leaf = nullptr, *d_leaf_ord = nullptr, *d_sizes = nullptr;
    int *d_colmax = nullptr, *d_rf = nullptr, *d_node_key = nullptr;
    double *d_L = nullptr, *d_ml = nullptr, *d_bpv = nullptr;
    short* d_vit = nullptr;
    int *d_order = nullptr, *d_counts_sorted = nullptr, *d_sizes_tab = nullptr;
    double *d_ml_tab = nullptr, *d_bpv_tab = nullptr;
    auto cleanup = [&]() {  // work_free: the buffers stay with the context (common.cuh)
        for (const void* q : {(const void*)d_prefix, (const void*)d_parent, (const void*)d_is_leaf, (const void*)d_leaf_ord, (const void*)d_sizes, (const void*)d_L,
                              (const void*)d_ml, (const void*)d_vit, (const void*)d_colmax, (const void*)d_rf, (const void*)d_node_key, (const void*)d_bpv,
                              (const void*)d_order, (const void*)d_counts_sorted, (const void*)d_sizes_tab, (const void*)d_ml_tab, (const void*)d_bpv_tab})
            work_free(ctx, q);
    };
#define VT_CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return CAFE_GPU_ERR_CUDA; } } while (0)
    VT_CK(work_malloc(ctx, &d_prefix, n * sizeof(int)));
    VT_CK(work_malloc(ctx, &d_parent, n * sizeof(int)));
    VT_CK(work_malloc(ctx, &d_is_leaf, n * sizeof(int)));
    VT_CK(work_malloc(ctx, &d_leaf_ord, n * sizeof(int)));
    VT_CK(work_malloc(ctx, &d_sizes, (size_t)F * n * sizeof(int)));
    VT_CK(work_malloc(ctx, &d_ml, (size_t)F * sizeof(double)));
    VT_CK(work_malloc(ctx, &d_L, (size_t)n_internal * FC * Vp * sizeof(double)));
    VT_CK(work_malloc(ctx, &d_vit, (size_t)n * FC * Vp * sizeof(short)));
    if (forced) {
        VT_CK(work_malloc(ctx, &d_colmax, ctx->F_pad * sizeof(int)));
        VT_CK(work_malloc(ctx, &d_rf, ctx->F_pad * sizeof(int)));
        VT_CK(cudaMemcpyAsync(d_colmax, h_colmax.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        VT_CK(cudaMemcpyAsync(d_rf, h_rf.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    const int* d_counts_use = ctx->d_counts;
    if (sorted) {
        VT_CK(work_malloc(ctx, &d_order, ctx->F_pad * sizeof(int)));
        VT_CK(work_malloc(ctx, &d_counts_sorted, (size_t)ctx->n_leaves * ctx->F_pad * sizeof(int)));
        VT_CK(cudaMemsetAsync(d_counts_sorted, 0, (size_t)ctx->n_leaves * ctx->F_pad * sizeof(int), ctx->stream));
        VT_CK(cudaMemcpyAsync(d_order, order.data(), ctx->F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        k_vit_gather_counts<<<dim3((F + 255) / 256, ctx->n_leaves), 256, 0, ctx->stream>>>(ctx->d_counts, d_counts_sorted, d_order, F, ctx->F_pad);
        ctx->launches++;
        d_counts_use = d_counts_sorted;
    }
    VT_CK(cudaMemcpyAsync(d_prefix, prefix.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    VT_CK(cudaMemcpyAsync(d_parent, parent.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    VT_CK(cudaMemcpyAsync(d_is_leaf, is_leaf.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    VT_CK(cudaMemcpyAsync(d_leaf_ord, leaf_ord.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const size_t node_stride = (size_t)FC * Vp;

    // post-order list of the internal nodes
    std::vector<int> post;
    {
        std::vector<std::pair<int, int>> st{{ctx->root, 0}};
        while (!st.empty()) {
            auto [v, state] = st.back(); st.pop_back();
            if (ctx->left[v] < 0) continue;
            if (state == 0) { st.push_back({v, 1}); st.push_back({ctx->right[v], 0}); st.push_back({ctx->left[v], 0}); }
            else post.push_back(v);
        }
    }
    const size_t smem = (size_t)VT_FB * W * sizeof(double);
    VT_CK(cudaFuncSetAttribute(k_viterbi_node, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));

    for (int fam0 = 0; fam0 < F; fam0 += FC) {
        const int nf = std::min(FC, F - fam0);
        for (int v : post) {
            const bool is_root = (v == ctx->root);
            const int r0 = is_root ? (forced ? 1 : ctx->root_min) : ctx->rmin;
            const int nrows = is_root ? (forced ? std::min(Vp, ctx->S - 1) : ctx->R) : W;
            VitChild ch[2];
            const int kids[2] = {ctx->left[v], ctx->right[v]};
            for (int s = 0; s < 2; ++s) {
                const int c = kids[s];
                VitChild& C = ch[s];
                C.is_leaf = ctx->left[c] < 0;
                C.MT = ctx->d_MT + (size_t)ctx->node_key[c] * mat;
                C.counts = nullptr; C.err_rowptr = nullptr; C.err_col = nullptr; C.err_val = nullptr; C.L = nullptr;
                C.vit = d_vit + (size_t)c * node_stride;
                if (C.is_leaf) {
                    const int k = c / 2;
                    C.counts = d_counts_use + (size_t)k * ctx->F_pad + fam0;
                    const int e = ctx->leaf_err.empty() ? -1 : ctx->leaf_err[k];
                    if (e >= 0) { C.err_rowptr = ctx->errs[e].d_rowptr; C.err_col = ctx->errs[e].d_col; C.err_val = ctx->errs[e].d_val; }
                } else {
                    C.L = d_L + (size_t)slot_of[c] * node_stride;
                }
            }
            dim3 grid((Vp + VT_THREADS - 1) / VT_THREADS, (nf + VT_FB - 1) / VT_FB);
            k_viterbi_node<<<grid, VT_THREADS, smem, ctx->stream>>>(ch[0], ch[1], Sp, Vp, W, r0, nrows, nf,
                                                                    d_L + (size_t)slot_of[v] * node_stride,
                                                                    forced ? d_colmax + fam0 : nullptr, forced ? d_rf + fam0 : nullptr, is_root ? 1 : 0);
            ctx->launches++;
        }
        k_viterbi_backtrack<<<(nf + 127) / 128, 128, 0, ctx->stream>>>(
            d_prefix, n, d_parent, d_is_leaf, d_leaf_ord, ctx->root, d_L + (size_t)slot_of[ctx->root] * node_stride, d_vit, node_stride, Vp,
            ctx->R, forced ? 1 : ctx->root_min, ctx->rmin, d_counts_use, ctx->F_pad, fam0, nf, n, d_sizes, d_ml, forced ? d_rf + fam0 : nullptr);
        ctx->launches++;
        VT_CK(cudaGetLastError());
    }
    if (branch_pv_out) {
        VT_CK(work_malloc(ctx, &d_node_key, n * sizeof(int)));
        VT_CK(work_malloc(ctx, &d_bpv, (size_t)F * n * sizeof(double)));
        VT_CK(cudaMemcpyAsync(d_node_key, ctx->node_key.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        const long long warps = (long long)F * n;
        k_viterbi_branch_pvalues<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_M, d_node_key, d_parent, Sp, n, F, d_sizes, forced ? d_colmax : nullptr, W, d_bpv);
        ctx->launches++;
        VT_CK(cudaGetLastError());
        const double* src = d_bpv;
        if (sorted) {
            VT_CK(work_malloc(ctx, &d_bpv_tab, (size_t)F * n * sizeof(double)));
            k_vit_scatter_rows<double><<<(unsigned)((warps + 255) / 256), 256, 0, ctx->stream>>>(d_bpv, d_bpv_tab, d_order, F, n);
            ctx->launches++;
            src = d_bpv_tab;
        }
        VT_CK(cudaMemcpyAsync(branch_pv_out, src, (size_t)F * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (sizes_out) {
        const int* src = d_sizes;
        if (sorted) {
            VT_CK(work_malloc(ctx, &d_sizes_tab, (size_t)F * n * sizeof(int)));
            k_vit_scatter_rows<int><<<(unsigned)(((long long)F * n + 255) / 256), 256, 0, ctx->stream>>>(d_sizes, d_sizes_tab, d_order, F, n);
            ctx->launches++;
            src = d_sizes_tab;
        }
        VT_CK(cudaMemcpyAsync(sizes_out, src, (size_t)F * n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (maxlik_out) {
        const double* src = d_ml;
        if (sorted) {
            VT_CK(work_malloc(ctx, &d_ml_tab, (size_t)F * sizeof(double)));
            k_vit_scatter_rows<double><<<(F + 255) / 256, 256, 0, ctx->stream>>>(d_ml, d_ml_tab, d_order, F, 1);
            ctx->launches++;
            src = d_ml_tab;
        }
        VT_CK(cudaMemcpyAsync(maxlik_out, src, (size_t)F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    VT_CK(cudaStreamSynchronize(ctx->stream));
#undef VT_CK
    cleanup();
    return CAFE_GPU_OK;
}
