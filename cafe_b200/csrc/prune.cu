// prune.cu — K2: batched Felsenstein pruning over all gene families.
//
// Replaces initialize_leaf_likelihoods + compute_internal_node_likelihood + square_matrix_multiply
// (cafe/cafe_tree.c:191-323, libtree/birthdeath.c:163-182) and the root reduction of
// compute_posterior (cafe/lambda.cpp:657-689), for every family at once.
//
// Data layout in HBM
//   d_M  [D][Sp][Sp]   transition matrices M[s][c], zero padded, Sp % 16 == 0
//   d_MT [D][Sp][Sp]   transposed copies (leaf edges are column gathers -> contiguous row reads)
//   d_vec[slot][F_pad][Vp]  node likelihood vectors, family-major, zero beyond W, Vp % 16 == 0
//   d_counts[leaf][F_pad]   leaf-major observed sizes
//
// Per internal tree edge the work is a batched matvec = GEMM  Out[f][i] = sum_j M[r0+i][j] * L[f][j]
// on the fp64 tensor pipe (DMMA.8x8x4); leaf edges are gathers fused into the epilogue together with
// the child product.
#include <algorithm>
#include <cstdlib>
#include <functional>

#include "common.cuh"

// =============================================================================================
// host: schedule (post-order, slot allocation)
// =============================================================================================
int build_schedule(cafe_gpu_ctx* ctx) {
    ctx->ops.clear();
    const int n = ctx->n_nodes;
    if (n < 3) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "tree needs at least two leaves");
    // Sethi–Ullman style need: number of simultaneously live vector slots to evaluate a subtree
    std::vector<int> need(n, 0);
    std::function<int(int)> calc = [&](int v) -> int {
        if (ctx->left[v] < 0) return need[v] = 0;
        int a = calc(ctx->left[v]), b = calc(ctx->right[v]);
        int hi = std::max(a, b), lo = std::min(a, b);
        // evaluate the needier child first, hold its result (1 slot) while evaluating the other,
        // then one more slot for the output
        int k = std::max(hi, lo + (hi > 0 ? 1 : 0));
        int live_children = (a > 0) + (b > 0);
        return need[v] = std::max(k, live_children + 1);
    };
    calc(ctx->root);

    std::vector<int> free_slots;
    int n_slots = 0;
    auto alloc = [&]() {
        if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
        return n_slots++;
    };
    auto leaf_ord = [&](int node) { return node / 2; };
    std::function<int(int)> eval = [&](int v) -> int {
        int a = ctx->left[v], b = ctx->right[v];
        bool ia = ctx->left[a] >= 0, ib = ctx->left[b] >= 0;
        PruneOp op{};
        op.node = v;
        op.is_root = (v == ctx->root);
        op.gemm_child = -1; op.gemm_key = -1; op.in_slot = -1; op.other_kind = 0;
        op.leaf_a = op.leaf_b = -1; op.leaf_a_key = op.leaf_b_key = -1;
        if (!ia && !ib) {
            op.out_slot = alloc();
            op.leaf_a = leaf_ord(a); op.leaf_a_key = ctx->node_key[a];
            op.leaf_b = leaf_ord(b); op.leaf_b_key = ctx->node_key[b];
            ctx->ops.push_back(op);
            return op.out_slot;
        }
        if (ia != ib) {
            int g = ia ? a : b, l = ia ? b : a;
            int sg = eval(g);
            op.out_slot = alloc();
            op.gemm_child = g; op.gemm_key = ctx->node_key[g]; op.in_slot = sg;
            op.other_kind = 1; op.leaf_a = leaf_ord(l); op.leaf_a_key = ctx->node_key[l];
            ctx->ops.push_back(op);
            free_slots.push_back(sg);
            return op.out_slot;
        }
        int first = need[a] >= need[b] ? a : b, second = (first == a) ? b : a;
        int s1 = eval(first);
        int s2 = eval(second);
        op.out_slot = alloc();
        op.gemm_child = first; op.gemm_key = ctx->node_key[first]; op.in_slot = s1; op.other_kind = 0;
        ctx->ops.push_back(op);
        free_slots.push_back(s1);
        PruneOp op2 = op;
        op2.gemm_child = second; op2.gemm_key = ctx->node_key[second]; op2.in_slot = s2; op2.other_kind = 2;
        ctx->ops.push_back(op2);
        free_slots.push_back(s2);
        return op.out_slot;
    };
    eval(ctx->root);
    ctx->n_slots = n_slots;
    return CAFE_GPU_OK;
}

// =============================================================================================
// device
// =============================================================================================
namespace {

struct LeafSrc {
    const double* MT;  // transposed matrix of the leaf's branch (key base)
    const int* counts; // [F_pad] observed sizes of this leaf
    const int* err_rowptr;  // nullable: sparse error rows
    const int* err_col;
    const double* err_val;
};

// factor[r] = sum_j M[r][j] * leafvec[j]  for a leaf: a column gather (cafe_tree.c:204-210 one-hot)
// or the sparse error-row combination (cafe_tree.c:196-203), restricted to columns <= colmax.
__device__ __forceinline__ double leaf_factor(const LeafSrc& L, int Sp, int count, int colmax, int r) {
    if (L.err_rowptr == nullptr) {
        return (count <= colmax) ? L.MT[(size_t)count * Sp + r] : 0.0;
    }
    double s = 0.0;
    for (int k = L.err_rowptr[count]; k < L.err_rowptr[count + 1]; ++k) {
        int j = L.err_col[k];
        if (j <= colmax) s = __dadd_rn(s, __dmul_rn(L.MT[(size_t)j * Sp + r], L.err_val[k]));
    }
    return s;
}

// ---- leaf pair: both children are leaves (cherries) -------------------------------------------
// one warp per family; i contiguous across lanes (coalesced reads of two MT rows, coalesced write)
__global__ void __launch_bounds__(256)
k_leaf_pair(LeafSrc A, LeafSrc B, int Sp, int F, int r0, int nrows, int Vp, const int* __restrict__ colmax,
            int default_colmax, double* __restrict__ out, int mask_rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= F) return;
    const int f = warp;
    const int ca = A.counts[f], cb = B.counts[f];
    const int cm = colmax ? colmax[f] : default_colmax;
    double* o = out + (size_t)f * Vp;
    for (int i = lane; i < Vp; i += 32) {
        double v = 0.0;
        if (i < nrows && (!mask_rows || i <= cm)) {
            v = leaf_factor(A, Sp, ca, cm, r0 + i) * leaf_factor(B, Sp, cb, cm, r0 + i);
        }
        o[i] = v;
    }
}

// ---- GEMM node: Out[f][i] = (sum_j M[r0+i][j] * In[f][j]) (*) other ---------------------------
constexpr int TM = 128;          // families per CTA tile
constexpr int TN = 128;          // output rows (i) per CTA tile
constexpr int BK = 16;           // k per stage: 16 doubles = 128 B rows -> 128B swizzle
constexpr int GEMM_THREADS = 256;  // 8 warps, all along N: warp tile 128 x 16
constexpr int MB = TM / 8;       // 16 m-blocks
constexpr int NB = 2;            // n-blocks per warp

// swizzled byte offset of (row, 16-byte chunk) inside a [rows][128 B] tile (== TMA SWIZZLE_128B)
__device__ __forceinline__ int swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_node_gemm(const double* __restrict__ Min, int Sp, int r0, int nrows, int K,  // matrix of the gemm child
            const double* __restrict__ in, double* __restrict__ out, int Vp, int F,
            int other_kind, LeafSrc leaf, const int* __restrict__ colmax, int default_colmax, int mask_rows) {
    __shared__ __align__(1024) unsigned char sA[TM * 128];
    __shared__ __align__(1024) unsigned char sB[TN * 128];

    const int f0 = blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int pg = mma_row_perm(g);

    double acc[MB][NB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

    const int rows_valid = min(TM, F - f0);
    const int mb_valid = (rows_valid + 7) >> 3;
    const int n_valid = nrows - n0;  // output rows of this tile that exist (may exceed TN)
    const bool warp_active = (warp * 16) < n_valid;

    // fragment byte offsets inside a tile for the 4 k4-steps of a stage
    int frag_off[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) frag_off[kk] = pg * 128 + (((2 * kk + (q >> 1)) ^ pg) << 4) + ((q & 1) << 3);

    // Stage pipeline: the next K block travels global -> registers while the DMMAs of the current one run, then registers ->
    // shared memory between two barriers (the fused kernels use TMA with the same swizzle instead).
    constexpr int PIECES = TM * 8 / GEMM_THREADS;  // 16-byte pieces per thread and operand
    double2 pa[PIECES], pb[PIECES];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
            const int ch = threadIdx.x + i * GEMM_THREADS, row = ch >> 3, c = ch & 7;
            pa[i] = *reinterpret_cast<const double2*>(in + (size_t)(f0 + row) * Vp + k0 + 2 * c);
            const int mr = r0 + n0 + row;
            pb[i] = make_double2(0.0, 0.0);
            if (mr < Sp) pb[i] = *reinterpret_cast<const double2*>(Min + (size_t)mr * Sp + k0 + 2 * c);
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
            const int ch = threadIdx.x + i * GEMM_THREADS, row = ch >> 3, c = ch & 7;
            *reinterpret_cast<double2*>(sA + swz(row, c)) = pa[i];
            *reinterpret_cast<double2*>(sB + swz(row, c)) = pb[i];
        }
    };
    gload(0);
    sstore();
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += BK) {
        const bool more = k0 + BK < K;
        if (more) gload(k0 + BK);
        if (warp_active) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (k0 + 4 * kk < K) {
                    double b[NB];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
                        b[nb] = *reinterpret_cast<const double*>(sB + (warp * 16 + nb * 8) * 128 + frag_off[kk]);
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
                        if (mb < mb_valid) {
                            double a = *reinterpret_cast<const double*>(sA + mb * 1024 + frag_off[kk]);
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a, b[nb]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (more) { sstore(); __syncthreads(); }
    }

    // ---- epilogue: child product, masks, store --------------------------------------------------
    // lane holds C[row g][cols 2q, 2q+1] of each 8x8 block -> tile row pi(g), tile cols pi(2q), pi(2q+1)
    const int pc0 = mma_row_perm(2 * q), pc1 = mma_row_perm(2 * q + 1);
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int f = f0 + mb * 8 + pg;
        if (mb < mb_valid && f < F) {
            const int cm = colmax ? colmax[f] : default_colmax;
            const int cnt = (other_kind == 1) ? leaf.counts[f] : 0;
            double* o = out + (size_t)f * Vp;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = n0 + warp * 16 + nb * 8 + (h ? pc1 : pc0);
                    if (i < Vp) {
                        double v = 0.0;
                        if (i < nrows && (!mask_rows || i <= cm)) {
                            v = acc[mb][nb][h];
                            if (other_kind == 1) v *= leaf_factor(leaf, Sp, cnt, cm, r0 + i);
                            else if (other_kind == 2) v *= o[i];
                        }
                        o[i] = v;
                    }
                }
            }
        }
    }
}

// ---- root: per-family posterior reduction (cafe/lambda.cpp:657-689) ------------------------------
// one warp per family: max_j L[j] (first maximum, mathfunc.c:9-40), max_j exp(log L[j] + log prior[j])
__global__ void __launch_bounds__(256)
k_root_posterior(const double* __restrict__ Lroot, int Vp, int F, int R, const double* __restrict__ logprior,
                 double* __restrict__ logpost, double* __restrict__ maxlik, int* __restrict__ argmax) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= F) return;
    const double* L = Lroot + (size_t)warp * Vp;
    double ml = -1.0, mp = -1.0;
    int am = 0x7fffffff;
    for (int j = lane; j < R; j += 32) {
        double l = L[j];
        if (l > ml) { ml = l; am = j; }  // strict >: keeps the first maximum within the lane (j ascending)
        double p = exp(log(l) + logprior[j]);
        if (p > mp) mp = p;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double oml = __shfl_xor_sync(0xffffffffu, ml, off);
        int oam = __shfl_xor_sync(0xffffffffu, am, off);
        double omp = __shfl_xor_sync(0xffffffffu, mp, off);
        if (oml > ml || (oml == ml && oam < am)) { ml = oml; am = oam; }
        if (omp > mp) mp = omp;
    }
    if (lane == 0) {
        logpost[warp] = log(mp);
        maxlik[warp] = ml;
        argmax[warp] = am;
    }
}

}  // namespace

// =============================================================================================
// host: launch the schedule
// =============================================================================================
static LeafSrc make_leaf_src(cafe_gpu_ctx* ctx, int leaf, int key) {
    LeafSrc L{};
    L.MT = ctx->d_MT + (size_t)key * ctx->Sp * ctx->Sp;
    L.counts = ctx->d_counts + (size_t)leaf * ctx->F_pad;
    int e = ctx->leaf_err.empty() ? -1 : ctx->leaf_err[leaf];
    if (e >= 0) {
        L.err_rowptr = ctx->errs[e].d_rowptr;
        L.err_col = ctx->errs[e].d_col;
        L.err_val = ctx->errs[e].d_val;
    }
    return L;
}


int launch_prune(cafe_gpu_ctx* ctx, double* d_Lroot_out) {
    if (ctx->timing) CAFE_CK(ctx, cudaEventRecord(ctx->evt(ctx->ring_k2, EV_K2_BEGIN), ctx->stream));
    const bool no_fused = std::getenv("CAFE_GPU_NO_FUSED") != nullptr;  // A/B switches for tests and profiling (read per call)
    const bool fused_v1 = std::getenv("CAFE_GPU_FUSED_V1") != nullptr;  // first-generation fused kernel
    if (!no_fused && !fused_v1 && fused2_supported(ctx)) {
        // fused persistent kernel, one CTA per SM: whole tree + root reduction in one launch
        int rc = launch_prune_fused2(ctx, d_Lroot_out);
        if (rc) return rc;
    } else if (!no_fused && fused_supported(ctx)) {
        // first-generation fused kernel: error-model leaves, leaf counts outside the vector, two-leaf trees
        int rc = launch_prune_fused(ctx, d_Lroot_out);
        if (rc) return rc;
    } else {
        // per-node kernels (also the path for per-family column windows, see pvalue.cu / conddist.cu)
        int root_slot = -1;
        int rc = launch_prune_ops(ctx, ctx->d_counts, ctx->F_pad, ctx->F, ctx->F_pad, nullptr, ctx->root_min, ctx->R, false, &root_slot);
        if (rc) return rc;
        const double* Lroot = ctx->d_vec + (size_t)root_slot * ctx->F_pad * ctx->Vp;
        const int warps_per_block = 8;
        k_root_posterior<<<(ctx->F + warps_per_block - 1) / warps_per_block, 256, 0, ctx->stream>>>(
            Lroot, ctx->Vp, ctx->F, ctx->R, ctx->d_logprior, ctx->d_logpost, ctx->d_maxlik, ctx->d_argmax);
        ctx->launches++;
        if (d_Lroot_out) {
            CAFE_CK(ctx, cudaMemcpy2DAsync(d_Lroot_out, (size_t)ctx->R * sizeof(double), Lroot,
                                           (size_t)ctx->Vp * sizeof(double), (size_t)ctx->R * sizeof(double), ctx->F,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (ctx->timing) { CAFE_CK(ctx, cudaEventRecord(ctx->evt(ctx->ring_k2, EV_K2_END), ctx->stream)); ctx->ring_k2++; }
    CAFE_CK(ctx, cudaGetLastError());
    return CAFE_GPU_OK;
}

// Runs the schedule with the per-node kernels over F families.  counts_base[leaf * leaf_stride + f] is the
// observed (or simulated) size of family f at leaf `leaf`; d_colmax (nullable) gives a per-family column
// window (cafe_family.c:250-254, conditional_distribution.cpp:29); the root rows are root_r0..root_r0+root_rows-1.
// With skip_root the root node is left to the caller (single-row roots of the conditional distribution).
// The caller guarantees ctx->d_vec holds n_slots*F_pad*Vp doubles.
int launch_prune_ops(cafe_gpu_ctx* ctx, const int* counts_base, size_t leaf_stride, int F, int F_pad, const int* d_colmax,
                     int root_r0, int root_rows, bool skip_root, int* root_slot_out) {
    const size_t slot_stride = (size_t)F_pad * ctx->Vp;
    const size_t mat_stride = (size_t)ctx->Sp * ctx->Sp;
    const int K = ctx->W;  // columns min..max of the matvec (cafe_tree.c:223)
    for (const PruneOp& op : ctx->ops) {
        if (op.is_root && skip_root) continue;
        const int r0 = op.is_root ? root_r0 : 0;
        const int nrows = op.is_root ? root_rows : ctx->W;
        const int mask_rows = op.is_root ? 0 : 1;
        double* out = ctx->d_vec + (size_t)op.out_slot * slot_stride;
        auto leaf_src = [&](int leaf, int key) {
            LeafSrc L = make_leaf_src(ctx, leaf, key);
            L.counts = counts_base + (size_t)leaf * leaf_stride;
            return L;
        };
        if (op.gemm_child < 0) {
            LeafSrc A = leaf_src(op.leaf_a, op.leaf_a_key), B = leaf_src(op.leaf_b, op.leaf_b_key);
            const int warps_per_block = 8;
            k_leaf_pair<<<(F + warps_per_block - 1) / warps_per_block, 256, 0, ctx->stream>>>(
                A, B, ctx->Sp, F, r0, nrows, ctx->Vp, d_colmax, ctx->W - 1, out, mask_rows);
        } else {
            LeafSrc L{};
            if (op.other_kind == 1) L = leaf_src(op.leaf_a, op.leaf_a_key);
            dim3 grid((F + TM - 1) / TM, (ctx->Vp + TN - 1) / TN);
            k_node_gemm<<<grid, GEMM_THREADS, 0, ctx->stream>>>(
                ctx->d_M + (size_t)op.gemm_key * mat_stride, ctx->Sp, r0, nrows, K,
                ctx->d_vec + (size_t)op.in_slot * slot_stride, out, ctx->Vp, F, op.other_kind, L, d_colmax,
                ctx->W - 1, mask_rows);
        }
        ctx->launches++;
        if (op.is_root && root_slot_out) *root_slot_out = op.out_slot;
    }
    CAFE_CK(ctx, cudaGetLastError());
    return CAFE_GPU_OK;
}
