// common.cuh — shared declarations of the sm_100a likelihood path (context, launch helpers, PTX wrappers).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <array>
#include <string>
#include <vector>

#include "../../include/cafe_gpu.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: every CUDA call goes through CK(); failures are recorded in ctx->err and surface
// as CAFE_GPU_ERR_CUDA at the C-ABI.  There is no CPU fallback anywhere in this library.
// ---------------------------------------------------------------------------------------------
#define CAFE_CK(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ +   \
                         ":" + std::to_string(__LINE__) + ")";                                     \
            return CAFE_GPU_ERR_CUDA;                                                              \
        }                                                                                          \
    } while (0)

#define CAFE_FAIL(ctx, code, msg)                                                                  \
    do {                                                                                           \
        (ctx)->err = (msg);                                                                        \
        return (code);                                                                             \
    } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// One distinct transition-matrix key, cafe/cafe_tree.c:374-391 (int branch length!).
struct BdKey {
    int t;
    double lambda, mu;
};

// Host-computed scalars of one key (libtree/birthdeath.c:246-262), uploaded for K1.
struct BdKeyParams {
    double log_alpha, log_beta, log_coeff, coeff;
    double q;  // coeff / (alpha * beta): the part of term(j+1) / term(j) that does not depend on (s, c, j)  (bd_matrix.cu)
    int mode;  // 0: rows>=1 zero (coeff<=0); 1: identity (coeff==1); 2: mu<0 sum; 3: mu>=0 sum
    int rec;   // 1: the anchored term recurrence of bd_matrix.cu applies (all scalars finite, q in range)
};

// Sparse rows of an error matrix: for observed count `o`, the non-zero (true size j, value) pairs in
// ascending j — the dense sum of cafe/cafe_tree.c:196-203 + libtree/birthdeath.c:172-180 adds exact
// zeros for every other j, so the result is identical.
struct ErrModelDev {
    int dim = 0;
    int max_up = 0;           // largest (true size - observed size) with a non-zero entry: how far a row reaches above its diagonal
    int* d_rowptr = nullptr;  // [dim+1]
    int* d_col = nullptr;
    double* d_val = nullptr;
};

// One step of the pruning schedule: a node of the tree in post-order.
struct PruneOp {
    int node;        // parent node v
    int is_root;
    int gemm_child;  // internal child whose vector is multiplied by its matrix (-1: none, leaf pair)
    int gemm_key;    // key index of gemm_child's branch
    int in_slot;     // slot of gemm_child's vector
    int out_slot;    // slot of v's vector
    int other_kind;  // 0: none (first of two internal children), 1: leaf sibling, 2: multiply into out_slot (second internal child)
    int leaf_a, leaf_a_key;  // leaf ordinal (leaf order) / key for leaf factors; leaf_a used by other_kind==1 and leaf pairs
    int leaf_b, leaf_b_key;  // second leaf of a leaf pair
};

struct cafe_gpu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    int sm_count = 148;

    // tree (nlist order)
    int n_nodes = 0, n_leaves = 0, root = -1;
    std::vector<int> left, right, parent, t_int;
    std::vector<double> branchlength;
    std::vector<int> prefix_nonroot;  // non-root nodes in prefix order (RNG consumption order of K4)

    // ranges
    bool have_ranges = false;
    int rmin = 0, rmax = 0, root_min = 0, root_max = 0;
    int W = 0, R = 0, S = 0;  // vector length, root rows, matrix dimension
    int Sp = 0;               // padded matrix leading dimension
    int Vp = 0;               // padded vector leading dimension (>= max(W,R))

    // lnC
    double* d_lnc = nullptr;   // [lnc_rows][lnc_cols]   lnC(n,x)
    double* d_lncT = nullptr;  // [lnc_cols][lnc_rows]   transposed copy
    int lnc_rows = 0, lnc_cols = 0;

    // families
    int F = 0, F_pad = 0;
    int* d_counts = nullptr;  // [n_leaves][F_pad]  (leaf-major: coalesced over families)
    size_t counts_cap = 0;    // ints allocated for d_counts
    int* d_mult = nullptr;    // [F_pad]
    int* d_first = nullptr;   // [F_pad]
    std::vector<int> h_counts;  // [F][n_leaves] as given
    std::vector<int> h_fam_max; // [F] largest count of a family: its forced range (cafe_family.c:236-255) follows from it
    int max_count = 0;
    bool has_missing = false;   // some leaf count is -1 (a species without data): Viterbi only

    // prior
    double* d_logprior = nullptr;  // [R] log(prior[i])
    double* d_prior_mant = nullptr;  // [R] prior[i] = mant * 2^exp with mant in [1,2) (exact compare of L*prior, prune_fused2.cu)
    int* d_prior_exp = nullptr;      // [R]
    std::vector<double> h_prior;

    // error models, per leaf (leaf order)
    std::vector<ErrModelDev> errs;        // owned models
    std::vector<int> leaf_err;            // per leaf: index into errs or -1
    int* d_leaf_err_rowptr_base = nullptr;

    // rates / keys
    std::vector<double> lambda, mu;  // per node
    std::vector<BdKey> keys;
    std::vector<int> node_key;  // per node: key index (-1 root)
    BdKeyParams* d_keyparams = nullptr;
    int keys_cap = 0;
    bool matrices_valid = false;
    // K1 sharded over ranks (cafe_gpu_set_key_shard): this context builds keys [key_lo, key_hi) only, the caller all-gathers
    int shard_rank = 0, shard_world = 1;
    int key_lo = 0, key_hi = 0, keys_per_rank = 0;
    bool matrices_need_exchange = false;
    double* d_M = nullptr;   // [D][Sp][Sp]  M[s][c]
    double* d_MT = nullptr;  // [D][Sp][Sp]  MT[c][s]
    size_t mat_cap = 0;      // allocated matrices

    // pruning schedule + vector slots
    std::vector<PruneOp> ops;
    int n_slots = 0;
    double* d_vec = nullptr;  // [n_slots][F_pad][Vp]
    size_t vec_cap = 0;

    // per-family outputs
    double* d_logpost = nullptr;  // [F_pad]
    double* d_maxlik = nullptr;   // [F_pad]
    int* d_argmax = nullptr;      // [F_pad]
    double* d_score = nullptr;    // [2] partial score, min first index of a zero family (as double)
    double* h_score = nullptr;    // pinned [2]
    bool results_valid = false;

    double* d_Lroot_cache = nullptr;  // root likelihood rows of all families, grown on demand (p-values)
    size_t Lroot_cache_cap = 0;
    void* fused_state = nullptr;   // prune_fused.cu private state (device schedule, scratch)
    void* fused2_state = nullptr;  // prune_fused2.cu private state
    void* k1_state = nullptr;      // bd_matrix.cu private state (ratio tables of the term recurrence)
    // work buffers of the multi-launch passes (conditional distribution): handed out by work_malloc, taken back by work_free
    // WITHOUT a cudaFree - a call allocates and frees a dozen buffers of up to 100 MB, and on a busy host the driver calls took
    // anything from 30 to 600 ms per pass against 200 ms of kernels
    struct WorkBlock { void* p; size_t bytes; bool used; };
    std::vector<WorkBlock> work_blocks;

    // multi-GPU (comm.cu): a context is one rank of an NCCL communicator.  Either one process per GPU (cafe_gpu_comm_init:
    // world ranks in world processes) or one process with several devices (cafe_gpu_create_multi: the leader context owns
    // the contexts of the other devices in `peers`, every public call on the leader fans out).
    void* nccl_comm = nullptr;            // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    std::vector<cafe_gpu_ctx*> peers;     // leader only: the other local contexts, in rank order (ranks 1..n-1)
    cafe_gpu_ctx* leader = nullptr;       // peers only
    double* d_score_all = nullptr;        // [comm_world][2] gathered partial scores
    double* d_score_final = nullptr;      // [2]
    int fam_lo = 0;                       // in-process multi: first unique pattern of the leader's list held by this context

    // bookkeeping
    int64_t launches = 0;
    bool timing = false;
    // ring of CUDA-event octets, one per objective evaluation (see the EV_* indices below)
    static constexpr int kRing = 256;
    static constexpr int kEv = 8;
    std::vector<cudaEvent_t> ring;   // kEv * kRing events, created on enable_timing
    int ring_k1 = 0, ring_k2 = 0;    // evaluations recorded since the last collect
    int ring_x = 0, ring_r = 0;      // exchanges / reductions recorded (multi-GPU only)
    cudaEvent_t evt(int i, int which) { return ring[kEv * (i % kRing) + which]; }
};
enum { EV_K1_BEGIN = 0, EV_K1_END, EV_K2_BEGIN, EV_K2_END, EV_XCHG_BEGIN, EV_XCHG_END, EV_RED_BEGIN, EV_RED_END };

template <class T>
inline cudaError_t work_malloc(cafe_gpu_ctx* ctx, T** out, size_t bytes) {
    bytes = std::max<size_t>(bytes, 16);
    int best = -1;
    for (int i = 0; i < (int)ctx->work_blocks.size(); ++i) {
        const auto& b = ctx->work_blocks[i];
        if (!b.used && b.bytes >= bytes && (best < 0 || b.bytes < ctx->work_blocks[best].bytes)) best = i;
    }
    if (best >= 0 && ctx->work_blocks[best].bytes <= 2 * bytes + (1u << 20)) {
        ctx->work_blocks[best].used = true;
        *out = (T*)ctx->work_blocks[best].p;
        return cudaSuccess;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // give the cached blocks back and try once more
        for (auto& b : ctx->work_blocks) if (!b.used) { cudaFree(b.p); b.p = nullptr; }
        ctx->work_blocks.erase(std::remove_if(ctx->work_blocks.begin(), ctx->work_blocks.end(), [](const cafe_gpu_ctx::WorkBlock& b) { return b.p == nullptr; }),
                               ctx->work_blocks.end());
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
    }
    ctx->work_blocks.push_back({p, bytes, true});
    *out = (T*)p;
    return cudaSuccess;
}
inline void work_free(cafe_gpu_ctx* ctx, const void* p) {
    if (!p) return;
    for (auto& b : ctx->work_blocks) if (b.p == p) { b.used = false; return; }
}
inline void work_release_all(cafe_gpu_ctx* ctx) {
    for (auto& b : ctx->work_blocks) cudaFree(b.p);
    ctx->work_blocks.clear();
}

// kernels (defined in the .cu files of this directory); all launch on ctx->stream
int launch_bd_matrices(cafe_gpu_ctx* ctx);                              // bd_matrix.cu   (K1: M and MT of the keys [key_lo, key_hi))
int launch_transpose_keys(cafe_gpu_ctx* ctx, int lo, int hi, int lo2, int hi2);  // bd_matrix.cu (MT of the keys [lo,hi) and [lo2,hi2))
int launch_prune(cafe_gpu_ctx* ctx, double* d_Lroot_out /*nullable*/);  // prune.cu       (K2)
int launch_score_reduce(cafe_gpu_ctx* ctx, double* d_out2);             // reduce.cu      (K3)
int build_schedule(cafe_gpu_ctx* ctx);                                  // prune.cu (host)
int launch_prune_ops(cafe_gpu_ctx* ctx, const int* counts_base, size_t leaf_stride, int F, int F_pad, const int* d_colmax,
                     int root_r0, int root_rows, bool skip_root, int* root_slot_out);  // prune.cu (per-node kernels)
int ensure_vec_buffers(cafe_gpu_ctx* ctx, size_t F_pad);                // api.cu
bool fused_supported(const cafe_gpu_ctx* ctx);                          // prune_fused.cu
int launch_prune_fused(cafe_gpu_ctx* ctx, double* d_Lroot_out);         // prune_fused.cu (K2, fused persistent kernel)
void fused_release(cafe_gpu_ctx* ctx);
void k1_release(cafe_gpu_ctx* ctx);                                      // bd_matrix.cu
bool fused2_supported(const cafe_gpu_ctx* ctx);                         // prune_fused2.cu
int launch_prune_fused2(cafe_gpu_ctx* ctx, double* d_Lroot_out);        // prune_fused2.cu (K2, one CTA per SM, default)
// The same kernel on any table of sizes: the observed families (score), or simulated / windowed ones (conditional distribution,
// p-values).  counts[k * leaf_stride + f] = size of family f at leaf k; d_colmax (nullable): per-family column window;
// root rows root_r0 .. root_r0 + root_rows - 1; d_root_pick / d_L0_out (nullable): root SIZE per family whose likelihood is
// written to d_L0_out[f]; d_Lroot_out (nullable): all root rows, [F][root_rows]; posterior: the root reduction of the score.
struct Fused2Job {
    const int* counts = nullptr; size_t leaf_stride = 0; int F = 0, F_pad = 0;
    const int* d_colmax = nullptr;
    int root_r0 = 0, root_rows = 0;
    const int* d_root_pick = nullptr; double* d_L0_out = nullptr;
    const int* d_root_need = nullptr;  // windowed jobs: of the root rows copied to d_Lroot_out only the first d_root_need[f] of family f are read
    double* d_Lroot_out = nullptr;
    bool posterior = false;
};
bool fused2_windowed_supported(const cafe_gpu_ctx* ctx);
void fused2_tile_slots(const cafe_gpu_ctx* ctx, int F, std::vector<std::array<int, 3>>& slots);  // host mirror of the kernel's tile split
int launch_prune_fused2_job(cafe_gpu_ctx* ctx, const Fused2Job& job);
void fused2_release(cafe_gpu_ctx* ctx);
int run_conditional_distribution(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed, int row_lo, int row_hi,
                                 double* cd_out);                       // conddist.cu    (K4)
int run_pvalues(cafe_gpu_ctx* ctx, const double* cd, int cd_rows, int n_samples, double* out);  // pvalue.cu (K5)
int run_cut_pvalues(cafe_gpu_ctx* ctx, const double* L1, const double* L2, int F, int rf, const double* cd1, const double* cd2,
                    int cdlen, double* out);  // pvalue.cu (branch cutting)
int run_lrt_branch_stretch(cafe_gpu_ctx* ctx, const uint8_t* tested, const double* lengthened_mu, double* base_out, double* best_out,
                           int32_t* steps_out);                     // lrt.cu
int build_one_matrix(cafe_gpu_ctx* ctx, int key);                       // api.cu (K1 for one key)
int run_viterbi(cafe_gpu_ctx* ctx, int32_t* sizes_out, double* maxlik_out, bool forced, double* branch_pv_out);  // viterbi.cu
// comm.cu: NCCL plumbing (libnccl is loaded lazily with dlopen, a context without a communicator never touches it)
int comm_exchange_matrices(std::vector<cafe_gpu_ctx*>& locals);   // in-place all-gather of d_M over the ranks + local transposes
int comm_reduce_scores(std::vector<cafe_gpu_ctx*>& locals);      // all-gather of {partial score, first zero} + ordered final sum
void comm_release(cafe_gpu_ctx* ctx);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// fp64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col).  SASS: DMMA.8x8x4 (the only fp64 MMA shape
// sm_100a has; tcgen05.mma has no .kind::f64 — SURVEY.md §7).
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Row permutation inside an 8-row MMA block.  With 128-byte rows and the 128B swizzle (16-byte chunk
// index ^= row & 7) the natural rows 0..3 of a half-warp collide pairwise; mapping MMA row g to
// tile row pi(g) = 2*(g&3) + (g>>2) makes every fragment load conflict-free.
__device__ __forceinline__ int mma_row_perm(int g) { return 2 * (g & 3) + (g >> 2); }

#endif  // __CUDACC__
