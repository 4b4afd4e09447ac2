// api.cu — the C-ABI (include/cafe_gpu.h) over the CUDA kernels of this directory.
//
// Host-side bookkeeping only: topology, key de-duplication (gather_keys/add_key,
// cafe/cafe_tree.c:374-446), key scalars (libtree/birthdeath.c:246-262), buffer management.
// All arithmetic on families happens in the kernels; there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"

static thread_local std::string g_create_error;

extern "C" {

int cafe_gpu_abi_version(void) { return CAFE_GPU_ABI_VERSION; }

int cafe_gpu_create(cafe_gpu_ctx** out, int device) {
    if (!out) return CAFE_GPU_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                         " — this library has no CPU fallback";
        cudaGetLastError();
        return CAFE_GPU_ERR_NO_DEVICE;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n) { g_create_error = "device index out of range"; return CAFE_GPU_ERR_ARG; }
    cafe_gpu_ctx* ctx = new cafe_gpu_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice/cudaStreamCreate failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return CAFE_GPU_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaMalloc(&ctx->d_score, 2 * sizeof(double)) != cudaSuccess || cudaMalloc(&ctx->d_score_final, 2 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_score, 2 * sizeof(double)) != cudaSuccess) {
        g_create_error = std::string("cafe_gpu_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        cudaFree(ctx->d_score); cudaFree(ctx->d_score_final);
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return CAFE_GPU_ERR_CUDA;
    }
    *out = ctx;
    return CAFE_GPU_OK;
}

static void free_err_models(cafe_gpu_ctx* ctx) {
    for (auto& e : ctx->errs) { cudaFree(e.d_rowptr); cudaFree(e.d_col); cudaFree(e.d_val); }
    ctx->errs.clear();
}
// drop the error models no leaf refers to any more (a session re-uploads its models whenever they change)
static void gc_err_models(cafe_gpu_ctx* ctx) {
    std::vector<int> remap(ctx->errs.size(), -1);
    std::vector<ErrModelDev> kept;
    for (int& e : ctx->leaf_err) {
        if (e < 0) continue;
        if (remap[e] < 0) { remap[e] = (int)kept.size(); kept.push_back(ctx->errs[e]); }
        e = remap[e];
    }
    for (size_t i = 0; i < ctx->errs.size(); ++i)
        if (remap[i] < 0) { cudaFree(ctx->errs[i].d_rowptr); cudaFree(ctx->errs[i].d_col); cudaFree(ctx->errs[i].d_val); }
    ctx->errs.swap(kept);
}

static void destroy_one(cafe_gpu_ctx* ctx);
void cafe_gpu_destroy(cafe_gpu_ctx* ctx) {
    if (!ctx) return;
    for (cafe_gpu_ctx* p : ctx->peers) destroy_one(p);
    ctx->peers.clear();
    destroy_one(ctx);
}
static void destroy_one(cafe_gpu_ctx* ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    cudaFree(ctx->d_score_all); cudaFree(ctx->d_score_final); cudaFree(ctx->d_Lroot_cache);
    cudaFree(ctx->d_lnc); cudaFree(ctx->d_lncT); cudaFree(ctx->d_counts); cudaFree(ctx->d_mult); cudaFree(ctx->d_first);
    cudaFree(ctx->d_logprior); cudaFree(ctx->d_prior_mant); cudaFree(ctx->d_prior_exp); cudaFree(ctx->d_keyparams); cudaFree(ctx->d_M); cudaFree(ctx->d_MT); cudaFree(ctx->d_vec);
    cudaFree(ctx->d_logpost); cudaFree(ctx->d_maxlik); cudaFree(ctx->d_argmax); cudaFree(ctx->d_score);
    cudaFreeHost(ctx->h_score);
    free_err_models(ctx);
    fused_release(ctx);
    fused2_release(ctx);
    k1_release(ctx);
    work_release_all(ctx);
    for (cudaEvent_t e : ctx->ring) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* cafe_gpu_last_error(const cafe_gpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cafe_gpu_set_stream(cafe_gpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return CAFE_GPU_OK;
}

static int one_synchronize(cafe_gpu_ctx* ctx) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CAFE_GPU_OK;
}

// ------------------------------------------------------------------------------------------- setup
static int one_set_tree(cafe_gpu_ctx* ctx, int n_nodes, const int32_t* left, const int32_t* right, const double* branchlength) {
    if (!ctx || n_nodes < 3 || (n_nodes & 1) == 0 || !left || !right || !branchlength)
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: need an odd number (>=3) of nodes and non-null arrays");
    work_release_all(ctx);  // cached work buffers were sized for the old problem
    ctx->n_nodes = n_nodes;
    ctx->n_leaves = (n_nodes + 1) / 2;
    ctx->left.assign(left, left + n_nodes);
    ctx->right.assign(right, right + n_nodes);
    ctx->branchlength.assign(branchlength, branchlength + n_nodes);
    ctx->parent.assign(n_nodes, -1);
    for (int i = 0; i < n_nodes; ++i) {
        bool leaf = left[i] < 0;
        if (leaf != (right[i] < 0)) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: tree must be binary");
        if (leaf != ((i & 1) == 0)) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: nodes must be in nlist (infix) order: leaves even, internal odd");
        if (!leaf) {
            if (left[i] >= n_nodes || right[i] >= n_nodes || left[i] == right[i]) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: bad child index");
            ctx->parent[left[i]] = i;
            ctx->parent[right[i]] = i;
        }
    }
    ctx->root = -1;
    for (int i = 0; i < n_nodes; ++i)
        if (ctx->parent[i] < 0) {
            if (ctx->root >= 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: more than one root");
            ctx->root = i;
        }
    ctx->t_int.resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        if (i != ctx->root && !(branchlength[i] > 0))  // cafe_cmd_tree rejects these, cafe_commands.cpp:1113-1116
            CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_tree: non-root branch length must be > 0");
        ctx->t_int[i] = (int)branchlength[i];  // add_key, cafe_tree.c:376
    }
    // prefix order (node, head subtree, tail subtree) — libtree/tree.c:101-124
    ctx->prefix_nonroot.clear();
    std::vector<int> st{ctx->root};
    while (!st.empty()) {
        int v = st.back(); st.pop_back();
        if (ctx->left[v] >= 0) { st.push_back(ctx->right[v]); st.push_back(ctx->left[v]); }
        if (v != ctx->root) ctx->prefix_nonroot.push_back(v);
    }
    ctx->keys.clear(); ctx->node_key.clear(); ctx->ops.clear();
    ctx->matrices_valid = false; ctx->results_valid = false;
    ctx->leaf_err.assign(ctx->n_leaves, -1);
    free_err_models(ctx);
    // every buffer whose size depends on the number of leaves or nodes goes: the family table (n_leaves x F_pad), the
    // matrices (capacity >= n_nodes - 1 keys, plus n_leaves error-leaf matrices behind the transposed copies)
    cudaFree(ctx->d_counts); ctx->d_counts = nullptr; ctx->counts_cap = 0;
    ctx->F = 0; ctx->h_counts.clear();
    cudaFree(ctx->d_M); cudaFree(ctx->d_MT); ctx->d_M = ctx->d_MT = nullptr; ctx->mat_cap = 0;
    cudaFree(ctx->d_keyparams); ctx->d_keyparams = nullptr; ctx->keys_cap = 0;
    return CAFE_GPU_OK;
}

static int one_set_ranges(cafe_gpu_ctx* ctx, int range_min, int range_max, int root_min, int root_max) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (range_min != 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "set_ranges: range_min must be 0 (init_family_size, cafe_family.c:357-364)");
    if (range_max < 1 || root_min < 0 || root_max < root_min) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_ranges: bad range");
    if (range_max != ctx->rmax || std::max(range_max, root_max) + 1 != ctx->S) work_release_all(ctx);  // sized by the old vectors / matrices
    ctx->rmin = range_min; ctx->rmax = range_max; ctx->root_min = root_min; ctx->root_max = root_max;
    ctx->W = range_max - range_min + 1;
    ctx->R = root_max - root_min + 1;
    ctx->S = std::max(range_max, root_max) + 1;  // cafe_main.c:325, birthdeath.c:241
    ctx->Sp = round_up(ctx->S, 16);
    ctx->Vp = round_up(std::max(ctx->W, ctx->R), 16);
    ctx->have_ranges = true;
    ctx->matrices_valid = false; ctx->results_valid = false;
    // geometry changed: drop size-dependent buffers
    cudaFree(ctx->d_M); cudaFree(ctx->d_MT); ctx->d_M = ctx->d_MT = nullptr; ctx->mat_cap = 0;
    cudaFree(ctx->d_vec); ctx->d_vec = nullptr; ctx->vec_cap = 0;
    // the prior arrays are sized by the old root range: a new cafe_gpu_set_prior is required (check_ready)
    cudaFree(ctx->d_logprior); cudaFree(ctx->d_prior_mant); cudaFree(ctx->d_prior_exp);
    ctx->d_logprior = ctx->d_prior_mant = nullptr; ctx->d_prior_exp = nullptr;
    ctx->h_prior.clear();
    return CAFE_GPU_OK;
}

static int one_set_lnc_table(cafe_gpu_ctx* ctx, const double* lnc, int rows, int cols) {
    if (!ctx || !lnc || rows < 2 || cols < 2) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_lnc_table: bad arguments");
    cudaFree(ctx->d_lnc); cudaFree(ctx->d_lncT); ctx->d_lnc = ctx->d_lncT = nullptr;
    k1_release(ctx);  // K1's ratio tables are derived from this table
    size_t bytes = (size_t)rows * cols * sizeof(double);
    CAFE_CK(ctx, cudaMalloc(&ctx->d_lnc, bytes));
    CAFE_CK(ctx, cudaMalloc(&ctx->d_lncT, bytes));
    std::vector<double> T((size_t)rows * cols);
    for (int n = 0; n < rows; ++n)
        for (int x = 0; x < cols; ++x) T[(size_t)x * rows + n] = lnc[(size_t)n * cols + x];
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_lnc, lnc, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_lncT, T.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->lnc_rows = rows; ctx->lnc_cols = cols;
    ctx->matrices_valid = false;
    return CAFE_GPU_OK;
}

static int one_set_families(cafe_gpu_ctx* ctx, int n_families, int n_leaves, const int32_t* counts,
                          const int32_t* multiplicity, const int32_t* first_index) {
    if (!ctx || n_families < 1 || !counts) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_families: bad arguments");
    if (ctx->n_nodes == 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "set_families: set_tree first");
    if (n_leaves != ctx->n_leaves) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_families: n_leaves does not match the tree");
    const int F = n_families, F_pad = round_up(F, 128);
    if (F_pad != ctx->F_pad) work_release_all(ctx);  // cached work buffers of the p-value / Viterbi passes follow the table's size
    std::vector<int> T((size_t)n_leaves * F_pad, 0), mult(F_pad, 0), first(F_pad, 0);
    int mx = 0;
    bool missing = false;
    std::vector<int> fam_max(F, 0);
    for (int f = 0; f < F; ++f) {
        for (int k = 0; k < n_leaves; ++k) {
            int c = counts[(size_t)f * n_leaves + k];
            // -1: the species has no column in the family table (familysize < 0, cafe_family.c:214-216).  Only the Viterbi
            // reconstruction defines a meaning for it (viterbi.cpp:236-250); the likelihood paths refuse such a table (check_ready).
            if (c < -1) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_families: negative count");
            if (c < 0) missing = true;
            mx = std::max(mx, c);
            fam_max[f] = std::max(fam_max[f], c);
            T[(size_t)k * F_pad + f] = c;
        }
        mult[f] = multiplicity ? multiplicity[f] : 1;
        first[f] = first_index ? first_index[f] : f;
    }
    if (T.size() > ctx->counts_cap) {
        cudaFree(ctx->d_counts); ctx->d_counts = nullptr; ctx->counts_cap = 0;
        CAFE_CK(ctx, cudaMalloc(&ctx->d_counts, T.size() * sizeof(int)));
        ctx->counts_cap = T.size();
    }
    if (F_pad != ctx->F_pad) {
        cudaFree(ctx->d_mult); cudaFree(ctx->d_first);
        cudaFree(ctx->d_logpost); cudaFree(ctx->d_maxlik); cudaFree(ctx->d_argmax);
        ctx->d_mult = ctx->d_first = ctx->d_argmax = nullptr; ctx->d_logpost = ctx->d_maxlik = nullptr;
        ctx->F_pad = 0;
        CAFE_CK(ctx, cudaMalloc(&ctx->d_mult, F_pad * sizeof(int)));
        CAFE_CK(ctx, cudaMalloc(&ctx->d_first, F_pad * sizeof(int)));
        CAFE_CK(ctx, cudaMalloc(&ctx->d_logpost, F_pad * sizeof(double)));
        CAFE_CK(ctx, cudaMalloc(&ctx->d_maxlik, F_pad * sizeof(double)));
        CAFE_CK(ctx, cudaMalloc(&ctx->d_argmax, F_pad * sizeof(int)));
        cudaFree(ctx->d_vec); ctx->d_vec = nullptr; ctx->vec_cap = 0;
    }
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_counts, T.data(), T.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_mult, mult.data(), F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_first, first.data(), F_pad * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->F = F; ctx->F_pad = F_pad; ctx->max_count = mx; ctx->has_missing = missing;
    ctx->h_counts.assign(counts, counts + (size_t)F * n_leaves);
    ctx->h_fam_max.swap(fam_max);
    ctx->results_valid = false;
    return CAFE_GPU_OK;
}

static int one_set_prior(cafe_gpu_ctx* ctx, const double* prior, int len) {
    if (!ctx || !prior) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_prior: bad arguments");
    if (!ctx->have_ranges) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "set_prior: set_ranges first");
    if (len < ctx->R) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_prior: need at least root_max-root_min+1 values");
    ctx->h_prior.assign(prior, prior + ctx->R);
    std::vector<double> lp(ctx->R);
    for (int i = 0; i < ctx->R; ++i) lp[i] = std::log(prior[i]);  // lambda.cpp:682
    // prior = mant * 2^exp for the exact product compare of the fused root reduction (prune_fused2.cu)
    std::vector<double> pm(ctx->R);
    std::vector<int> pe(ctx->R);
    for (int i = 0; i < ctx->R; ++i) {
        if (prior[i] > 0 && std::isfinite(prior[i])) { int e; pm[i] = 2.0 * std::frexp(prior[i], &e); pe[i] = e - 1; }
        else { pm[i] = 1.0; pe[i] = -(1 << 30); }
    }
    cudaFree(ctx->d_logprior); ctx->d_logprior = nullptr;
    cudaFree(ctx->d_prior_mant); ctx->d_prior_mant = nullptr;
    cudaFree(ctx->d_prior_exp); ctx->d_prior_exp = nullptr;
    CAFE_CK(ctx, cudaMalloc(&ctx->d_logprior, ctx->R * sizeof(double)));
    CAFE_CK(ctx, cudaMalloc(&ctx->d_prior_mant, ctx->R * sizeof(double)));
    CAFE_CK(ctx, cudaMalloc(&ctx->d_prior_exp, ctx->R * sizeof(int)));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_logprior, lp.data(), ctx->R * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_prior_mant, pm.data(), ctx->R * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_prior_exp, pe.data(), ctx->R * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->results_valid = false;
    return CAFE_GPU_OK;
}

static int one_set_error_model(cafe_gpu_ctx* ctx, int leaf, const double* errormatrix, int dim) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (ctx->n_nodes == 0 || !ctx->have_ranges) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "set_error_model: set_tree and set_ranges first");
    if (leaf >= ctx->n_leaves) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_error_model: leaf out of range");
    ctx->results_valid = false;
    if (!errormatrix) {
        if (leaf < 0) ctx->leaf_err.assign(ctx->n_leaves, -1); else ctx->leaf_err[leaf] = -1;
        gc_err_models(ctx);
        return CAFE_GPU_OK;
    }
    if (dim < ctx->rmax + 1) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_error_model: dim must be >= range_max+1");
    std::vector<int> rowptr(dim + 1, 0), col;
    std::vector<double> val;
    int max_up = 0;
    for (int o = 0; o < dim; ++o) {
        for (int j = 0; j < dim; ++j) {
            double e = errormatrix[(size_t)o * dim + j];
            if (e != 0.0) { col.push_back(j); val.push_back(e); max_up = std::max(max_up, j - o); }
        }
        rowptr[o + 1] = (int)col.size();
    }
    ErrModelDev E;
    E.dim = dim;
    E.max_up = max_up;
    CAFE_CK(ctx, cudaMalloc(&E.d_rowptr, rowptr.size() * sizeof(int)));
    CAFE_CK(ctx, cudaMalloc(&E.d_col, std::max<size_t>(1, col.size()) * sizeof(int)));
    CAFE_CK(ctx, cudaMalloc(&E.d_val, std::max<size_t>(1, val.size()) * sizeof(double)));
    CAFE_CK(ctx, cudaMemcpy(E.d_rowptr, rowptr.data(), rowptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!col.empty()) {
        CAFE_CK(ctx, cudaMemcpy(E.d_col, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice));
        CAFE_CK(ctx, cudaMemcpy(E.d_val, val.data(), val.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->errs.push_back(E);
    int idx = (int)ctx->errs.size() - 1;
    if (leaf < 0) ctx->leaf_err.assign(ctx->n_leaves, idx); else ctx->leaf_err[leaf] = idx;
    gc_err_models(ctx);
    return CAFE_GPU_OK;
}

// ------------------------------------------------------------------------------------------- rates / K1
static BdKeyParams key_params(const BdKey& k) {
    // libtree/birthdeath.c:246-262 — evaluated on the host with the same libm as the reference
    BdKeyParams P{};
    double alpha, beta, coeff;
    const double t = (double)k.t;
    if (k.mu < 0 || k.lambda == k.mu) {
        alpha = k.lambda * t / (1 + k.lambda * t);
        beta = alpha;
        coeff = 1 - 2 * alpha;
    } else {
        double e_diff = std::exp((k.lambda - k.mu) * t);
        double numerator = e_diff - 1;
        double denominator = k.lambda * e_diff - k.mu;
        alpha = (k.mu * numerator) / denominator;
        beta = (k.lambda * numerator) / denominator;
        coeff = 1 - alpha - beta;
    }
    P.coeff = coeff;
    if (!(coeff > 0)) { P.mode = 0; return P; }      // init_zero_matrix (also catches NaN like the reference's else-branch would not fill)
    if (coeff == 1) { P.mode = 1; return P; }
    P.log_alpha = std::log(alpha);
    P.log_beta = std::log(beta);
    P.log_coeff = std::log(coeff);
    P.mode = (k.mu < 0) ? 2 : 3;                     // birthdeath.c:272-275
    // term(j+1) / term(j) = q (s-j)(c-j) / ((j+1)(s+c-1-j)); K1's recurrence kernel needs q finite and small enough that a term
    // cannot climb from below 2^-2000 to the normal range within one 16-term segment (bd_matrix.cu) - anything else (mu = 0,
    // lambda t below ~1e-6, ...) takes the term-by-term kernel
    // (from the rounded logs, not from alpha and beta: the reference's terms are exp() of sums of THOSE)
    P.q = (P.mode == 2) ? coeff * std::exp(-2 * P.log_alpha) : std::exp(P.log_coeff - P.log_alpha - P.log_beta);
    P.rec = std::isfinite(P.q) && P.q > 0 && P.q < 1e12 && std::isfinite(P.log_alpha) && std::isfinite(P.log_beta) &&
            std::isfinite(P.log_coeff);
    return P;
}

static int one_set_rates(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node) {
    if (!ctx || !lambda_per_node || !mu_per_node) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_rates: bad arguments");
    if (ctx->n_nodes == 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "set_rates: set_tree first");
    const int n = ctx->n_nodes;
    ctx->lambda.assign(lambda_per_node, lambda_per_node + n);
    ctx->mu.assign(mu_per_node, mu_per_node + n);
    ctx->keys.clear();
    ctx->node_key.assign(n, -1);
    // gather_keys over the prefix traversal, add_key's linear exact-compare de-duplication
    for (int v : ctx->prefix_nonroot) {
        BdKey k{ctx->t_int[v], ctx->lambda[v], ctx->mu[v]};
        int found = -1;
        for (size_t i = 0; i < ctx->keys.size(); ++i)
            if (ctx->keys[i].t == k.t && ctx->keys[i].lambda == k.lambda && ctx->keys[i].mu == k.mu) { found = (int)i; break; }
        if (found < 0) { ctx->keys.push_back(k); found = (int)ctx->keys.size() - 1; }
        ctx->node_key[v] = found;
    }
    ctx->matrices_valid = false; ctx->results_valid = false;
    return build_schedule(ctx);
}

static int ensure_matrix_buffers(cafe_gpu_ctx* ctx) {
    const size_t D = ctx->keys.size();
    // key range of this context: everything, or one of shard_world equal chunks (the all-gather needs equal chunks, so the
    // buffers hold keys_per_rank * shard_world matrices)
    ctx->keys_per_rank = (int)((D + ctx->shard_world - 1) / ctx->shard_world);
    ctx->key_lo = std::min<int>((int)D, ctx->shard_rank * ctx->keys_per_rank);
    ctx->key_hi = std::min<int>((int)D, ctx->key_lo + ctx->keys_per_rank);
    const size_t need = (size_t)ctx->keys_per_rank * ctx->shard_world;
    if (need > ctx->mat_cap) {
        cudaFree(ctx->d_M); cudaFree(ctx->d_MT); ctx->d_M = ctx->d_MT = nullptr;
        size_t cap = std::max<size_t>(need, (size_t)ctx->n_nodes - 1 + ctx->shard_world);  // never more keys than branches
        size_t bytes = cap * ctx->Sp * ctx->Sp * sizeof(double);
        // the transposed copy carries n_leaves extra matrices behind the keys: the error-model leaf matrices of prune_fused2.cu
        size_t bytes_t = (cap + ctx->n_leaves) * ctx->Sp * ctx->Sp * sizeof(double);
        CAFE_CK(ctx, cudaMalloc(&ctx->d_M, bytes));
        CAFE_CK(ctx, cudaMalloc(&ctx->d_MT, bytes_t));
        CAFE_CK(ctx, cudaMemsetAsync(ctx->d_M, 0, bytes, ctx->stream));   // padding stays zero forever
        CAFE_CK(ctx, cudaMemsetAsync(ctx->d_MT, 0, bytes_t, ctx->stream));
        ctx->mat_cap = cap;
    }
    if ((int)D > ctx->keys_cap) {
        cudaFree(ctx->d_keyparams); ctx->d_keyparams = nullptr;
        int cap = std::max<int>((int)D, ctx->n_nodes - 1);
        CAFE_CK(ctx, cudaMalloc(&ctx->d_keyparams, cap * sizeof(BdKeyParams)));
        ctx->keys_cap = cap;
    }
    return CAFE_GPU_OK;
}

static int one_build_matrices(cafe_gpu_ctx* ctx) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (!ctx->have_ranges || ctx->keys.empty()) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "build_matrices: set_ranges and set_rates first");
    if (!ctx->d_lnc) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "build_matrices: set_lnc_table first");
    const int maxfs = ctx->S - 1;
    if (ctx->lnc_rows < 2 * maxfs || ctx->lnc_cols < maxfs + 1)
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "build_matrices: lnC table too small (need rows >= 2*(S-1), cols >= S)");
    int rc = ensure_matrix_buffers(ctx);
    if (rc) return rc;
    std::vector<BdKeyParams> kp(ctx->keys.size());
    for (size_t i = 0; i < kp.size(); ++i) kp[i] = key_params(ctx->keys[i]);
    // small synchronous-semantics copy from pageable memory: staged by the runtime before return
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_keyparams, kp.data(), kp.size() * sizeof(BdKeyParams), cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_bd_matrices(ctx);
    if (rc) return rc;
    ctx->matrices_need_exchange = ctx->shard_world > 1;
    ctx->matrices_valid = !ctx->matrices_need_exchange;
    ctx->results_valid = false;
    return CAFE_GPU_OK;
}

int cafe_gpu_set_key_shard(cafe_gpu_ctx* ctx, int rank, int world) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (world < 1 || rank < 0 || rank >= world) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_key_shard: need 0 <= rank < world");
    ctx->shard_rank = rank; ctx->shard_world = world;
    ctx->matrices_valid = false; ctx->results_valid = false;
    return CAFE_GPU_OK;
}

int cafe_gpu_matrix_storage(cafe_gpu_ctx* ctx, void** d_M, void** d_MT, int64_t* doubles_per_key, int32_t* keys_per_rank) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (!ctx->d_M) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "matrix_storage: build_matrices first");
    if (d_M) *d_M = ctx->d_M;
    if (d_MT) *d_MT = ctx->d_MT;
    if (doubles_per_key) *doubles_per_key = (int64_t)ctx->Sp * ctx->Sp;
    if (keys_per_rank) *keys_per_rank = ctx->keys_per_rank;
    return CAFE_GPU_OK;
}

int cafe_gpu_matrices_exchanged(cafe_gpu_ctx* ctx) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (!ctx->matrices_need_exchange) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "matrices_exchanged: no sharded build_matrices pending");
    // the transposed copies of the keys other ranks built (this rank's own were transposed right after K1)
    cudaSetDevice(ctx->device);
    int rc = launch_transpose_keys(ctx, 0, ctx->key_lo, ctx->key_hi, (int)ctx->keys.size());
    if (rc) return rc;
    ctx->matrices_need_exchange = false;
    ctx->matrices_valid = true;
    return CAFE_GPU_OK;
}

int cafe_gpu_num_keys(const cafe_gpu_ctx* ctx) { return ctx ? (int)ctx->keys.size() : 0; }

static int one_get_matrix(cafe_gpu_ctx* ctx, int node, double* out, int out_dim) {
    if (!ctx || !out) return CAFE_GPU_ERR_ARG;
    if (!ctx->matrices_valid) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "get_matrix: build_matrices first");
    if (node < 0 || node >= ctx->n_nodes || ctx->node_key[node] < 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "get_matrix: node has no branch");
    if (out_dim != ctx->S) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "get_matrix: out_dim must equal S = max(range_max, root_max)+1");
    const double* src = ctx->d_M + (size_t)ctx->node_key[node] * ctx->Sp * ctx->Sp;
    CAFE_CK(ctx, cudaMemcpy2DAsync(out, (size_t)ctx->S * sizeof(double), src, (size_t)ctx->Sp * sizeof(double),
                                   (size_t)ctx->S * sizeof(double), ctx->S, cudaMemcpyDeviceToHost, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CAFE_GPU_OK;
}

// ------------------------------------------------------------------------------------------- K2 + K3
}  // extern "C"

// K1 for the single key `key` of ctx->keys (the other matrices stay as they are): the lengthened branch of lrt.cu
int build_one_matrix(cafe_gpu_ctx* ctx, int key) {
    const int D = (int)ctx->keys.size();
    if (key < 0 || key >= D || (size_t)D > ctx->mat_cap) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "build_one_matrix: bad key");
    if (D > ctx->keys_cap) {  // grow, keeping the parameters of the keys already built
        BdKeyParams* grown = nullptr;
        const int cap = std::max(D, 2 * ctx->keys_cap);
        CAFE_CK(ctx, cudaMalloc(&grown, cap * sizeof(BdKeyParams)));
        if (ctx->d_keyparams)
            CAFE_CK(ctx, cudaMemcpyAsync(grown, ctx->d_keyparams, ctx->keys_cap * sizeof(BdKeyParams), cudaMemcpyDeviceToDevice, ctx->stream));
        CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_keyparams);
        ctx->d_keyparams = grown;
        ctx->keys_cap = cap;
    }
    const BdKeyParams kp = key_params(ctx->keys[key]);
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->d_keyparams + key, &kp, sizeof(BdKeyParams), cudaMemcpyHostToDevice, ctx->stream));
    const int lo = ctx->key_lo, hi = ctx->key_hi;
    ctx->key_lo = key; ctx->key_hi = key + 1;
    const int rc = launch_bd_matrices(ctx);
    ctx->key_lo = lo; ctx->key_hi = hi;
    return rc;
}

int ensure_vec_buffers(cafe_gpu_ctx* ctx, size_t F_pad) {
    size_t need = (size_t)ctx->n_slots * F_pad * ctx->Vp;
    if (need > ctx->vec_cap) {
        cudaFree(ctx->d_vec); ctx->d_vec = nullptr;
        CAFE_CK(ctx, cudaMalloc(&ctx->d_vec, need * sizeof(double)));
        CAFE_CK(ctx, cudaMemsetAsync(ctx->d_vec, 0, need * sizeof(double), ctx->stream));
        ctx->vec_cap = need;
    }
    return CAFE_GPU_OK;
}

extern "C" {

static int check_ready(cafe_gpu_ctx* ctx, const char* who, bool missing_ok = false) {
    if (ctx->has_missing && !missing_ok)
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, std::string(who) + ": a leaf without data (count -1) - only cafe_gpu_viterbi reconstructs such families; the reference's pruning asserts (cafe_tree.c:207)");
    if (!ctx->matrices_valid) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, std::string(who) + (ctx->matrices_need_exchange ? ": all-gather the matrices and call cafe_gpu_matrices_exchanged first" : ": build_matrices first"));
    if (ctx->F == 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, std::string(who) + ": set_families first");
    if (!ctx->d_logprior) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, std::string(who) + ": set_prior first");
    if (ctx->leaf_err.empty() || std::all_of(ctx->leaf_err.begin(), ctx->leaf_err.end(), [](int e) { return e < 0; })) {
        // one-hot leaves must fit the vector (initialize_leaf_likelihoods asserts, cafe_tree.c:207)
        if (ctx->max_count >= std::max(ctx->W, ctx->R))
            CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, std::string(who) + ": a family count exceeds the likelihood vector (size_of_factor)");
    }
    for (int e : ctx->leaf_err)
        if (e >= 0 && ctx->max_count >= ctx->errs[e].dim) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, std::string(who) + ": a family count exceeds the error matrix");
    return CAFE_GPU_OK;
}

static int score_device(cafe_gpu_ctx* ctx, double* d_out2) {
    int rc = check_ready(ctx, "score");
    if (rc) return rc;
    rc = ensure_vec_buffers(ctx, ctx->F_pad);
    if (rc) return rc;
    rc = launch_prune(ctx, nullptr);
    if (rc) return rc;
    rc = launch_score_reduce(ctx, d_out2);
    if (rc) return rc;
    ctx->results_valid = true;
    return CAFE_GPU_OK;
}





static int one_family_results(cafe_gpu_ctx* ctx, double* log_max_posterior, double* max_likelihood, int32_t* argmax_likelihood) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (!ctx->results_valid) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "family_results: score first");
    if (log_max_posterior) CAFE_CK(ctx, cudaMemcpyAsync(log_max_posterior, ctx->d_logpost, ctx->F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (max_likelihood) CAFE_CK(ctx, cudaMemcpyAsync(max_likelihood, ctx->d_maxlik, ctx->F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (argmax_likelihood) CAFE_CK(ctx, cudaMemcpyAsync(argmax_likelihood, ctx->d_argmax, ctx->F * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CAFE_GPU_OK;
}

static int one_viterbi(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* max_likelihood_out) {
    if (!ctx || !node_sizes_out) return CAFE_GPU_ERR_ARG;
    int rc = check_ready(ctx, "viterbi", true);
    if (rc) return rc;
    return run_viterbi(ctx, node_sizes_out, max_likelihood_out, false, nullptr);
}

static int one_viterbi_report(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* branch_pvalues_out) {
    if (!ctx || !node_sizes_out) return CAFE_GPU_ERR_ARG;
    int rc = check_ready(ctx, "viterbi_report");
    if (rc) return rc;
    return run_viterbi(ctx, node_sizes_out, nullptr, true, branch_pvalues_out);
}

static int one_likelihood_ratio_test(cafe_gpu_ctx* ctx, const uint8_t* tested, const double* lengthened_mu_per_node,
                                   double* base_max_likelihood_out, double* best_max_likelihood_out, int32_t* steps_out) {
    if (!ctx || !best_max_likelihood_out) return CAFE_GPU_ERR_ARG;
    int rc = check_ready(ctx, "likelihood_ratio_test");
    if (rc) return rc;
    rc = ensure_vec_buffers(ctx, ctx->F_pad);
    if (rc) return rc;
    return run_lrt_branch_stretch(ctx, tested, lengthened_mu_per_node, base_max_likelihood_out, best_max_likelihood_out, steps_out);
}

static int one_family_likelihoods(cafe_gpu_ctx* ctx, double* L_out) {
    if (!ctx || !L_out) return CAFE_GPU_ERR_ARG;
    int rc = check_ready(ctx, "family_likelihoods");
    if (rc) return rc;
    rc = ensure_vec_buffers(ctx, ctx->F_pad);
    if (rc) return rc;
    double* d_L = nullptr;
    size_t bytes = (size_t)ctx->F * ctx->R * sizeof(double);
    CAFE_CK(ctx, cudaMalloc(&d_L, bytes));
    rc = launch_prune(ctx, d_L);
    if (rc) { cudaFree(d_L); return rc; }
    cudaError_t e = cudaMemcpyAsync(L_out, d_L, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_L);
    CAFE_CK(ctx, e);
    return CAFE_GPU_OK;
}

// ------------------------------------------------------------------------------------------- K4 / K5

static int one_conditional_distribution_rows(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed, int row_lo, int row_hi,
                                           double* cd_out) {
    if (!ctx || !cd_out || n_samples < 1) return CAFE_GPU_ERR_ARG;
    if (!ctx->matrices_valid) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "conditional_distribution_rows: build_matrices first");
    if (row_lo < 0 || row_hi > ctx->R || row_lo > row_hi) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "conditional_distribution_rows: need 0 <= row_lo <= row_hi <= root_max-root_min+1");
    if (row_lo == row_hi) return CAFE_GPU_OK;
    return run_conditional_distribution(ctx, n_samples, uniforms, seed, row_lo, row_hi, cd_out);
}

static int one_pvalues(cafe_gpu_ctx* ctx, const double* cd, int cd_rows, int n_samples, double* max_pvalue_out) {
    if (!ctx || !cd || !max_pvalue_out || cd_rows < 1 || n_samples < 1) return CAFE_GPU_ERR_ARG;
    if (!ctx->matrices_valid) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "pvalues: build_matrices first");
    if (ctx->F == 0) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "pvalues: set_families first");
    if (ctx->has_missing) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "pvalues: a leaf without data (count -1): the reference's pruning asserts (cafe_tree.c:207)");
    return run_pvalues(ctx, cd, cd_rows, n_samples, max_pvalue_out);
}

// ------------------------------------------------------------------------------------------- bookkeeping
int64_t cafe_gpu_launch_count(const cafe_gpu_ctx* ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    for (const cafe_gpu_ctx* p : ctx->peers) n += p->launches;
    return n;
}
void cafe_gpu_reset_launch_count(cafe_gpu_ctx* ctx) {
    if (!ctx) return;
    ctx->launches = 0;
    for (cafe_gpu_ctx* p : ctx->peers) p->launches = 0;
}
int cafe_gpu_enable_timing(cafe_gpu_ctx* ctx, int on) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    if (on && ctx->ring.empty()) {
        ctx->ring.resize((size_t)cafe_gpu_ctx::kEv * cafe_gpu_ctx::kRing);
        for (auto& e : ctx->ring) CAFE_CK(ctx, cudaEventCreate(&e));
    }
    ctx->timing = on != 0;
    ctx->ring_k1 = ctx->ring_k2 = ctx->ring_x = ctx->ring_r = 0;
    return CAFE_GPU_OK;
}
int cafe_gpu_timing_collect4(cafe_gpu_ctx* ctx, float* k1_ms, float* exchange_ms, float* k2_ms, float* reduce_ms, int cap) {
    if (!ctx || cap < 0) return CAFE_GPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    const int done = std::min(ctx->ring_k1, ctx->ring_k2);
    const bool multi = ctx->ring_x == ctx->ring_k1 && ctx->ring_r == ctx->ring_k2 && ctx->comm_world > 1;
    int n = std::min(done, std::min(cap, (int)cafe_gpu_ctx::kRing));
    int first = done - n;
    for (int i = 0; i < n; ++i) {
        const int e = first + i;
        if (k1_ms) CAFE_CK(ctx, cudaEventElapsedTime(&k1_ms[i], ctx->evt(e, EV_K1_BEGIN), ctx->evt(e, EV_K1_END)));
        if (k2_ms) CAFE_CK(ctx, cudaEventElapsedTime(&k2_ms[i], ctx->evt(e, EV_K2_BEGIN), ctx->evt(e, EV_K2_END)));
        if (exchange_ms) { exchange_ms[i] = 0.f; if (multi) CAFE_CK(ctx, cudaEventElapsedTime(&exchange_ms[i], ctx->evt(e, EV_XCHG_BEGIN), ctx->evt(e, EV_XCHG_END))); }
        if (reduce_ms) { reduce_ms[i] = 0.f; if (multi) CAFE_CK(ctx, cudaEventElapsedTime(&reduce_ms[i], ctx->evt(e, EV_RED_BEGIN), ctx->evt(e, EV_RED_END))); }
    }
    ctx->ring_k1 = ctx->ring_k2 = ctx->ring_x = ctx->ring_r = 0;
    return n;
}
int cafe_gpu_timing_collect(cafe_gpu_ctx* ctx, float* k1_ms, float* k2_ms, int cap) {
    return cafe_gpu_timing_collect4(ctx, k1_ms, nullptr, k2_ms, nullptr, cap);
}

double cafe_gpu_score_flops(const cafe_gpu_ctx* ctx) {
    if (!ctx || ctx->n_nodes == 0 || !ctx->have_ranges) return 0.0;
    double peers = 0.0;
    for (const cafe_gpu_ctx* p : ctx->peers) peers += cafe_gpu_score_flops(p);
    // internal edges only; 2*W*W per non-root parent, 2*R*W under the root
    double per_family = 0.0;
    for (int v = 0; v < ctx->n_nodes; ++v) {
        if (v == ctx->root || ctx->left[v] < 0) continue;  // v is an internal child => its edge is a real matvec
        per_family += (ctx->parent[v] == ctx->root) ? 2.0 * ctx->R * ctx->W : 2.0 * ctx->W * ctx->W;
    }
    return per_family * ctx->F + peers;
}

}  // extern "C"

// =================================================================================================
// The public entry points: every call on a context fans out over its local contexts (itself, plus the peers of a leader
// created by cafe_gpu_create_multi) with the right device current, and the evaluation calls run the NCCL exchange steps of
// comm.cu when the context is a rank of a communicator.
// =================================================================================================
#include <thread>

static std::vector<cafe_gpu_ctx*> locals_of(cafe_gpu_ctx* ctx) {
    std::vector<cafe_gpu_ctx*> v{ctx};
    v.insert(v.end(), ctx->peers.begin(), ctx->peers.end());
    return v;
}

#define CAFE_NEED_LEADER(ctx)                                                                                      \
    do {                                                                                                           \
        if (!(ctx)) return CAFE_GPU_ERR_ARG;                                                                       \
        if ((ctx)->leader) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "call this on the leader context of cafe_gpu_create_multi"); \
    } while (0)

// run fn(c) for every local context, in rank order, with c's device current; the first failure is reported on the leader
template <typename Fn>
static int each_local(cafe_gpu_ctx* ctx, Fn fn) {
    for (cafe_gpu_ctx* c : locals_of(ctx)) {
        cudaSetDevice(c->device);
        const int rc = fn(c);
        if (rc < 0) {
            if (c != ctx) ctx->err = "device " + std::to_string(c->device) + ": " + c->err;
            cudaSetDevice(ctx->device);
            return rc;
        }
    }
    cudaSetDevice(ctx->device);
    return CAFE_GPU_OK;
}
// the same with one host thread per local context: for the calls that synchronise inside (whole passes over the families)
template <typename Fn>
static int each_local_parallel(cafe_gpu_ctx* ctx, Fn fn) {
    std::vector<cafe_gpu_ctx*> L = locals_of(ctx);
    if (L.size() == 1) { cudaSetDevice(ctx->device); return fn(ctx); }
    std::vector<int> rc(L.size(), 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < L.size(); ++i)
        th.emplace_back([&, i] { cudaSetDevice(L[i]->device); rc[i] = fn(L[i]); });
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < L.size(); ++i)
        if (rc[i] < 0) {
            if (L[i] != ctx) ctx->err = "device " + std::to_string(L[i]->device) + ": " + L[i]->err;
            return rc[i];
        }
    return CAFE_GPU_OK;
}

// balanced contiguous slice of n items for rank r of w (sizes differ by at most one)
static void shard_bounds(int n, int w, int r, int& lo, int& hi) {
    const int base = n / w, extra = n % w;
    lo = r * base + std::min(r, extra);
    hi = lo + base + (r < extra ? 1 : 0);
}

extern "C" {

int cafe_gpu_synchronize(cafe_gpu_ctx* ctx) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [](cafe_gpu_ctx* c) { return one_synchronize(c); });
}
int cafe_gpu_set_tree(cafe_gpu_ctx* ctx, int n_nodes, const int32_t* left, const int32_t* right, const double* branchlength) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_tree(c, n_nodes, left, right, branchlength); });
}
int cafe_gpu_set_ranges(cafe_gpu_ctx* ctx, int range_min, int range_max, int root_min, int root_max) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_ranges(c, range_min, range_max, root_min, root_max); });
}
int cafe_gpu_set_lnc_table(cafe_gpu_ctx* ctx, const double* lnc, int rows, int cols) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_lnc_table(c, lnc, rows, cols); });
}
int cafe_gpu_set_prior(cafe_gpu_ctx* ctx, const double* prior, int len) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_prior(c, prior, len); });
}
int cafe_gpu_set_error_model(cafe_gpu_ctx* ctx, int leaf, const double* errormatrix, int dim) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_error_model(c, leaf, errormatrix, dim); });
}
int cafe_gpu_set_rates(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) { return one_set_rates(c, lambda_per_node, mu_per_node); });
}

// The families of a multi-device context are split into contiguous, balanced slices in rank order (no cross-family state
// except the sum and the first zero family, cafe/lambda.cpp:698-722).  With one process per GPU the caller passes its own slice.
int cafe_gpu_set_families(cafe_gpu_ctx* ctx, int n_families, int n_leaves, const int32_t* counts, const int32_t* multiplicity,
                          const int32_t* first_index) {
    CAFE_NEED_LEADER(ctx);
    if (ctx->peers.empty()) { cudaSetDevice(ctx->device); return one_set_families(ctx, n_families, n_leaves, counts, multiplicity, first_index); }
    const int w = (int)ctx->peers.size() + 1;
    if (n_families < w) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "set_families: fewer families than devices");
    std::vector<int32_t> first_all;
    if (!first_index) { first_all.resize(n_families); for (int f = 0; f < n_families; ++f) first_all[f] = f; first_index = first_all.data(); }
    int r = 0;
    return each_local(ctx, [&](cafe_gpu_ctx* c) {
        int lo, hi;
        shard_bounds(n_families, w, r++, lo, hi);
        c->fam_lo = lo;
        return one_set_families(c, hi - lo, n_leaves, counts + (size_t)lo * n_leaves, multiplicity ? multiplicity + lo : nullptr, first_index + lo);
    });
}

int cafe_gpu_build_matrices(cafe_gpu_ctx* ctx) {
    CAFE_NEED_LEADER(ctx);
    int rc = each_local(ctx, [](cafe_gpu_ctx* c) {
        if (c->comm_world > 1) { c->shard_rank = c->comm_rank; c->shard_world = c->comm_world; }  // K1 sharded over the ranks
        return one_build_matrices(c);
    });
    if (rc || ctx->comm_world == 1) return rc;
    std::vector<cafe_gpu_ctx*> L = locals_of(ctx);
    rc = comm_exchange_matrices(L);
    cudaSetDevice(ctx->device);
    if (rc && L[0] != ctx) ctx->err = L[0]->err;
    return rc;
}

int cafe_gpu_get_matrix(cafe_gpu_ctx* ctx, int node, double* out, int out_dim) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    return one_get_matrix(ctx, node, out, out_dim);
}

// K2 + K3 on every local context, then (with a communicator) the all-gather of the partial scores and the ordered final sum;
// the result lands in every local context's d_score_final (comm) or d_score (single).
static int score_all_device(cafe_gpu_ctx* ctx) {
    int rc = each_local(ctx, [](cafe_gpu_ctx* c) { return score_device(c, c->d_score); });
    if (rc || ctx->comm_world == 1) return rc;
    std::vector<cafe_gpu_ctx*> L = locals_of(ctx);
    rc = comm_reduce_scores(L);
    cudaSetDevice(ctx->device);
    return rc;
}
static int read_score(cafe_gpu_ctx* ctx, double* score_out, int32_t* first_zero_family) {
    const double* src = ctx->comm_world > 1 ? ctx->d_score_final : ctx->d_score;
    CAFE_CK(ctx, cudaMemcpyAsync(ctx->h_score, src, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (std::isinf(ctx->h_score[1])) {
        *score_out = ctx->h_score[0];
        if (first_zero_family) *first_zero_family = -1;
        return CAFE_GPU_OK;
    }
    *score_out = -std::numeric_limits<double>::infinity();  // log(0), lambda.cpp:753-760
    if (first_zero_family) *first_zero_family = (int32_t)ctx->h_score[1];
    return CAFE_GPU_ZERO_LIKELIHOOD;
}

int cafe_gpu_score(cafe_gpu_ctx* ctx, double* score_out, int32_t* first_zero_family) {
    CAFE_NEED_LEADER(ctx);
    if (!score_out) return CAFE_GPU_ERR_ARG;
    int rc = score_all_device(ctx);
    if (rc) return rc;
    return read_score(ctx, score_out, first_zero_family);
}
int cafe_gpu_objective(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node, double* score_out,
                       int32_t* first_zero_family) {
    int rc = cafe_gpu_set_rates(ctx, lambda_per_node, mu_per_node);
    if (rc) return rc;
    rc = cafe_gpu_build_matrices(ctx);
    if (rc) return rc;
    return cafe_gpu_score(ctx, score_out, first_zero_family);
}
int cafe_gpu_score_device(cafe_gpu_ctx* ctx, double* out_device) {
    CAFE_NEED_LEADER(ctx);
    if (!out_device) return CAFE_GPU_ERR_ARG;
    int rc = score_all_device(ctx);
    if (rc) return rc;
    CAFE_CK(ctx, cudaMemcpyAsync(out_device, ctx->comm_world > 1 ? ctx->d_score_final : ctx->d_score, 2 * sizeof(double),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    return CAFE_GPU_OK;
}
int cafe_gpu_objective_device(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node, double* out_device) {
    int rc = cafe_gpu_set_rates(ctx, lambda_per_node, mu_per_node);
    if (rc) return rc;
    rc = cafe_gpu_build_matrices(ctx);
    if (rc) return rc;
    return cafe_gpu_score_device(ctx, out_device);
}

// per-family outputs: every local context fills its slice of the leader's family order
int cafe_gpu_family_results(cafe_gpu_ctx* ctx, double* log_max_posterior, double* max_likelihood, int32_t* argmax_likelihood) {
    CAFE_NEED_LEADER(ctx);
    return each_local(ctx, [&](cafe_gpu_ctx* c) {
        const size_t o = ctx->peers.empty() ? 0 : (size_t)c->fam_lo;
        return one_family_results(c, log_max_posterior ? log_max_posterior + o : nullptr, max_likelihood ? max_likelihood + o : nullptr,
                                  argmax_likelihood ? argmax_likelihood + o : nullptr);
    });
}
int cafe_gpu_family_likelihoods(cafe_gpu_ctx* ctx, double* L_out) {
    CAFE_NEED_LEADER(ctx);
    if (!L_out) return CAFE_GPU_ERR_ARG;
    return each_local_parallel(ctx, [&](cafe_gpu_ctx* c) {
        const size_t o = ctx->peers.empty() ? 0 : (size_t)c->fam_lo;
        return one_family_likelihoods(c, L_out + o * c->R);
    });
}
int cafe_gpu_viterbi(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* max_likelihood_out) {
    CAFE_NEED_LEADER(ctx);
    if (!node_sizes_out) return CAFE_GPU_ERR_ARG;
    return each_local_parallel(ctx, [&](cafe_gpu_ctx* c) {
        const size_t o = ctx->peers.empty() ? 0 : (size_t)c->fam_lo;
        return one_viterbi(c, node_sizes_out + o * c->n_nodes, max_likelihood_out ? max_likelihood_out + o : nullptr);
    });
}
int cafe_gpu_viterbi_report(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* branch_pvalues_out) {
    CAFE_NEED_LEADER(ctx);
    if (!node_sizes_out) return CAFE_GPU_ERR_ARG;
    return each_local_parallel(ctx, [&](cafe_gpu_ctx* c) {
        const size_t o = ctx->peers.empty() ? 0 : (size_t)c->fam_lo;
        return one_viterbi_report(c, node_sizes_out + o * c->n_nodes, branch_pvalues_out ? branch_pvalues_out + o * c->n_nodes : nullptr);
    });
}
int cafe_gpu_pvalues(cafe_gpu_ctx* ctx, const double* cd, int cd_rows, int n_samples, double* max_pvalue_out) {
    CAFE_NEED_LEADER(ctx);
    if (!max_pvalue_out) return CAFE_GPU_ERR_ARG;
    return each_local_parallel(ctx, [&](cafe_gpu_ctx* c) {
        const size_t o = ctx->peers.empty() ? 0 : (size_t)c->fam_lo;
        return one_pvalues(c, cd, cd_rows, n_samples, max_pvalue_out + o);
    });
}

int cafe_gpu_cut_pvalues(cafe_gpu_ctx* ctx, const double* L_rest, const double* L_sub, int n_families, int rfsize,
                         const double* cd_rest, const double* cd_sub, int cdlen, double* cut_pvalue_out) {
    CAFE_NEED_LEADER(ctx);
    if (!L_rest || !cd_rest || !cut_pvalue_out || n_families < 0 || rfsize < 1 || cdlen < 1 || ((L_sub == nullptr) != (cd_sub == nullptr)))
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "cut_pvalues: bad arguments");
    cudaSetDevice(ctx->device);
    return run_cut_pvalues(ctx, L_rest, L_sub, n_families, rfsize, cd_rest, cd_sub, cdlen, cut_pvalue_out);
}

// The branch-stretch test shards by families like the score.  Only the table's first tested family starts from the parsed
// branch lengths (cafe/cafe_main.c:350,390), so the slices after the one that holds it mark their families 2.
int cafe_gpu_likelihood_ratio_test(cafe_gpu_ctx* ctx, const uint8_t* tested, const double* lengthened_mu_per_node,
                                   double* base_max_likelihood_out, double* best_max_likelihood_out, int32_t* steps_out) {
    CAFE_NEED_LEADER(ctx);
    if (!best_max_likelihood_out) return CAFE_GPU_ERR_ARG;
    if (ctx->peers.empty()) {
        cudaSetDevice(ctx->device);
        return one_likelihood_ratio_test(ctx, tested, lengthened_mu_per_node, base_max_likelihood_out, best_max_likelihood_out, steps_out);
    }
    std::vector<cafe_gpu_ctx*> L = locals_of(ctx);
    int F_all = 0;
    for (cafe_gpu_ctx* c : L) F_all += c->F;
    int first_tested = -1;
    for (int f = 0; f < F_all && first_tested < 0; ++f)
        if (!tested || tested[f] == 1) first_tested = f;
    const int n_nodes = ctx->n_nodes;
    std::vector<std::vector<uint8_t>> t_loc(L.size());
    std::vector<std::vector<double>> best_loc(L.size());
    std::vector<std::vector<int32_t>> steps_loc(L.size());
    for (size_t i = 0; i < L.size(); ++i) {
        cafe_gpu_ctx* c = L[i];
        t_loc[i].resize(c->F);
        for (int f = 0; f < c->F; ++f) {
            uint8_t t = tested ? tested[c->fam_lo + f] : 1;
            if (t == 1 && first_tested >= 0 && c->fam_lo > first_tested) t = 2;
            t_loc[i][f] = t;
        }
        best_loc[i].resize((size_t)n_nodes * c->F);
        steps_loc[i].resize((size_t)n_nodes * c->F);
    }
    int rc = CAFE_GPU_OK;
    {
        std::vector<int> rcs(L.size(), 0);
        std::vector<std::thread> th;
        for (size_t i = 0; i < L.size(); ++i)
            th.emplace_back([&, i] {
                cudaSetDevice(L[i]->device);
                rcs[i] = one_likelihood_ratio_test(L[i], t_loc[i].data(), lengthened_mu_per_node,
                                                   base_max_likelihood_out ? base_max_likelihood_out + L[i]->fam_lo : nullptr,
                                                   best_loc[i].data(), steps_loc[i].data());
            });
        for (auto& t : th) t.join();
        cudaSetDevice(ctx->device);
        for (size_t i = 0; i < L.size(); ++i)
            if (rcs[i] < 0 && rc == 0) { rc = rcs[i]; if (L[i] != ctx) ctx->err = L[i]->err; }
    }
    if (rc) return rc;
    for (size_t i = 0; i < L.size(); ++i)
        for (int b = 0; b < n_nodes; ++b)
            for (int f = 0; f < L[i]->F; ++f) {
                best_max_likelihood_out[(size_t)b * F_all + L[i]->fam_lo + f] = best_loc[i][(size_t)b * L[i]->F + f];
                if (steps_out) steps_out[(size_t)b * F_all + L[i]->fam_lo + f] = steps_loc[i][(size_t)b * L[i]->F + f];
            }
    return CAFE_GPU_OK;
}

// The rows (root sizes) of the conditional distribution are independent: a multi-device context splits them over its
// devices — the distributed form of the reference's pthreads over root sizes (cafe/conditional_distribution.cpp:86-120).
int cafe_gpu_conditional_distribution_rows(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed, int row_lo, int row_hi,
                                           double* cd_out) {
    CAFE_NEED_LEADER(ctx);
    if (!cd_out || n_samples < 1) return CAFE_GPU_ERR_ARG;
    if (ctx->peers.empty()) {
        cudaSetDevice(ctx->device);
        return one_conditional_distribution_rows(ctx, n_samples, uniforms, seed, row_lo, row_hi, cd_out);
    }
    const int w = (int)ctx->peers.size() + 1;
    std::vector<std::pair<int, int>> span(w);
    for (int i = 0; i < w; ++i) { int lo, hi; shard_bounds(row_hi - row_lo, w, i, lo, hi); span[i] = {row_lo + lo, row_lo + hi}; }
    std::vector<cafe_gpu_ctx*> L = locals_of(ctx);
    std::vector<int> rcs(w, 0);
    std::vector<std::thread> th;
    for (int i = 0; i < w; ++i)
        th.emplace_back([&, i] {
            cudaSetDevice(L[i]->device);
            rcs[i] = one_conditional_distribution_rows(L[i], n_samples, uniforms, seed, span[i].first, span[i].second,
                                                       cd_out + (size_t)(span[i].first - row_lo) * n_samples);
        });
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->device);
    for (int i = 0; i < w; ++i)
        if (rcs[i] < 0) { if (L[i] != ctx) ctx->err = L[i]->err; return rcs[i]; }
    return CAFE_GPU_OK;
}
int cafe_gpu_conditional_distribution(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed, double* cd_out) {
    if (!ctx) return CAFE_GPU_ERR_ARG;
    return cafe_gpu_conditional_distribution_rows(ctx, n_samples, uniforms, seed, 0, ctx->R, cd_out);
}

}  // extern "C"
