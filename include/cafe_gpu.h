/*
 * cafe_gpu.h — C-ABI of the B200-native birth–death pruning likelihood path.
 *
 * This is the drop-in boundary for CAFE's per-family likelihood hot path.  The
 * reference has no FFI layer; the seams this ABI replaces are, outermost first (paths relative to
 * the reference root):
 *
 *   B1  double __cafe_best_lambda_search(double* x, void* args)        cafe/lambda.cpp:726
 *       double cafe_best_lambda_mu_search(double* x, void* args)       cafe/lambdamu.cpp:323
 *         -> after cafe_shell_set_lambdas (cafe/cafe_shell.c:31) has put (lambda, mu) on every node,
 *            the body  reset_birthdeath_cache + get_posterior + cafe_free_birthdeath_cache  becomes
 *            cafe_gpu_set_rates + cafe_gpu_build_matrices + cafe_gpu_score  (= cafe_gpu_objective).
 *   B2  double get_posterior(pCafeFamily, pCafeTree, std::vector<double>& prior)   cafe/lambda.cpp:691
 *         -> cafe_gpu_score  (+ cafe_gpu_family_results for maxlh / per-family values)
 *   B3  void compute_tree_likelihoods(pCafeTree); double* get_likelihoods(pCafeTree)  cafe/cafe_tree.c:320-329
 *         -> cafe_gpu_family_likelihoods (all families at once)
 *       void reset_birthdeath_cache(pCafeTree, int, family_size_range*)  cafe/cafe_main.c:319
 *         -> cafe_gpu_build_matrices  (+ cafe_gpu_get_matrix for inspection)
 *   B4  matrix cafe_conditional_distribution(pCafeTree, family_size_range*, int, int)
 *                                                        cafe/conditional_distribution.cpp:86
 *         -> cafe_gpu_conditional_distribution
 *   B5  void cafe_tree_p_values(...) cafe/pvalue.cpp:143 + viterbi_set_max_pvalue cafe/viterbi.cpp:32
 *       as driven per family by viterbi_section cafe/viterbi.cpp:88-97
 *         -> cafe_gpu_pvalues
 *
 * Conventions: plain C, plain pointers and sizes, host pointers unless a parameter says "device".
 * Every function returns 0 on success, a negative cafe_gpu_status on error (text via
 * cafe_gpu_last_error), and cafe_gpu_score/objective return CAFE_GPU_ZERO_LIKELIHOOD (1) when some
 * family has likelihood 0 at every root size — the reference's "posterior = 0" exception
 * (cafe/lambda.cpp:715-720) — in which case *score_out = -inf and *first_zero_family names it.
 * One host thread per context; a context owns one device, one stream and all device memory.
 * There is NO CPU fallback: without a CUDA device cafe_gpu_create fails with CAFE_GPU_ERR_NO_DEVICE.
 *
 * Node numbering everywhere is the reference's nlist (infix) order — leaves at even indices,
 * internal nodes at odd indices (cafe/cafe_commands.cpp:1985-2051).  "Leaf order" means leaf k is
 * node 2k.  Only binary trees (cafe/cafe_tree.c:248-249).
 */
#ifndef CAFE_GPU_H
#define CAFE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cafe_gpu_ctx cafe_gpu_ctx;

enum cafe_gpu_status {
    CAFE_GPU_OK = 0,
    CAFE_GPU_ZERO_LIKELIHOOD = 1,
    CAFE_GPU_ERR_NO_DEVICE = -1,
    CAFE_GPU_ERR_CUDA = -2,
    CAFE_GPU_ERR_ARG = -3,
    CAFE_GPU_ERR_STATE = -4,
    CAFE_GPU_ERR_UNSUPPORTED = -5
};

/* ABI version of this header (bumped on incompatible change). */
#define CAFE_GPU_ABI_VERSION 2
int cafe_gpu_abi_version(void);

/* device < 0 selects the current CUDA device.  Fails (no CPU fallback) when there is none. */
int cafe_gpu_create(cafe_gpu_ctx** out, int device);
void cafe_gpu_destroy(cafe_gpu_ctx* ctx);
const char* cafe_gpu_last_error(const cafe_gpu_ctx* ctx);

/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*), e.g. the
 * caller's framework stream; NULL restores the context's own stream. */
int cafe_gpu_set_stream(cafe_gpu_ctx* ctx, void* cuda_stream);
int cafe_gpu_synchronize(cafe_gpu_ctx* ctx);

/* Tree topology.  left/right = children->head / children->tail (cafe/cafe_tree.c:248-249), -1 at
 * leaves.  branchlength as parsed (double); the (int) truncation of the matrix key
 * (cafe/cafe_tree.c:376, libtree/birthdeath.h:26-31) is applied inside. */
int cafe_gpu_set_tree(cafe_gpu_ctx* ctx, int n_nodes, const int32_t* left, const int32_t* right,
                      const double* branchlength);

/* family_size_range (libtree/family.h:10-15) as copied to the tree by copy_range_to_tree
 * (cafe/cafe_main.c:52-60).  range_min must be 0 (init_family_size, cafe/cafe_family.c:357-364).
 * Matrix dimension S = max(range_max, root_max) + 1 (cafe/cafe_main.c:325). */
int cafe_gpu_set_ranges(cafe_gpu_ctx* ctx, int range_min, int range_max, int root_min, int root_max);

/* lnC table built ON THE HOST with the reference's Lanczos gammaln (libcommon/mathfunc.c:87-119,
 * 224-229; libtree/chooseln_cache.h:27-41): row-major [rows][cols], rows >= 2*(S-1), cols >= S,
 * table[n*cols + x] = chooseln(n, x). */
int cafe_gpu_set_lnc_table(cafe_gpu_ctx* ctx, const double* lnc, int rows, int cols);

/* Unique family count patterns: counts[f*n_leaves + k] = size of family f in leaf k (leaf order).
 * multiplicity (NULL => 1): how many families of the original list share the pattern (the `ref`
 * short-cut of cafe/lambda.cpp:702-714, cafe/cafe_family.c:9-34).  first_index (NULL => f): index
 * of the first such family in the original list, reported by cafe_gpu_score on zero likelihood.
 * A count of -1 marks a species of the tree without data (familysize < 0 after cafe_family_set_size, cafe/cafe_family.c:214-216):
 * cafe_gpu_viterbi reconstructs it; every likelihood entry point refuses such a table (the reference's pruning asserts,
 * cafe/cafe_tree.c:207). */
int cafe_gpu_set_families(cafe_gpu_ctx* ctx, int n_families, int n_leaves, const int32_t* counts,
                          const int32_t* multiplicity, const int32_t* first_index);

/* Root prior, prior[i] for root size root_min+i (cafe/lambda.cpp:841-852); len >= root_max-root_min+1. */
int cafe_gpu_set_prior(cafe_gpu_ctx* ctx, const double* prior, int len);

/* Dense error matrix errormatrix[observed][true], row-major dim x dim (libtree/family.h:31-38,
 * cafe/error_model.cpp:145-259), attached to leaf `leaf` (leaf order) or to all leaves when
 * leaf < 0.  errormatrix == NULL detaches.  dim must be >= range_max+1. */
int cafe_gpu_set_error_model(cafe_gpu_ctx* ctx, int leaf, const double* errormatrix, int dim);

/* Per-node (lambda, mu) as cafe_shell_set_lambdas leaves them in
 * pcnode->birth_death_probabilities (cafe/cafe_shell.c:148-244); mu < 0 selects the lambda-only
 * formulas (libtree/birthdeath.c:250-254,272-273).  Root entries are ignored.  Distinct
 * (int t, lambda, mu) keys are found as gather_keys/add_key do (cafe/cafe_tree.c:374-446). */
int cafe_gpu_set_rates(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node);

/* K1: build one S x S transition matrix per distinct key (reset_birthdeath_cache,
 * cafe/cafe_main.c:319; compute_birthdeath_rates, libtree/birthdeath.c:238-286). */
int cafe_gpu_build_matrices(cafe_gpu_ctx* ctx);
int cafe_gpu_num_keys(const cafe_gpu_ctx* ctx);
/* Copy back the matrix used by the branch above `node`, row-major S x S. */
int cafe_gpu_get_matrix(cafe_gpu_ctx* ctx, int node, double* out, int out_dim);

/* K2+K3: pruning over all families, root posterior, score = sum_f mult_f * log(max_j L_f[j]*prior[j])
 * (cafe/lambda.cpp:657-724).  first_zero_family may be NULL. */
int cafe_gpu_score(cafe_gpu_ctx* ctx, double* score_out, int32_t* first_zero_family);
/* cafe_gpu_set_rates + cafe_gpu_build_matrices + cafe_gpu_score: one objective evaluation. */
int cafe_gpu_objective(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node,
                       double* score_out, int32_t* first_zero_family);
/* Same, asynchronous and device-resident: out_device[0] = score (of this context's families; of all ranks' families when a
 * communicator is attached), out_device[1] = (double) smallest first_index with zero likelihood, or +inf.  No host
 * synchronisation. */
int cafe_gpu_objective_device(cafe_gpu_ctx* ctx, const double* lambda_per_node, const double* mu_per_node,
                              double* out_device /* device pointer, 2 doubles */);

/* Device-resident score of the current matrices (K2+K3 only): out_device as for cafe_gpu_objective_device. */
int cafe_gpu_score_device(cafe_gpu_ctx* ctx, double* out_device /* device pointer, 2 doubles */);

/* ---- Multi-GPU.  Families shard naturally: get_posterior has no cross-family state except the running sum and
 * the first zero family (cafe/lambda.cpp:698-722).  With a communicator attached, one objective evaluation is: K1 on this
 * rank's ceil(D / world) distinct keys -> ncclAllGather of the matrices in place (+ a local transpose of the received ones) ->
 * K2 + K3 on this rank's families -> ncclAllGather of {partial score, first zero family} and a sum in rank order, all on the
 * context's stream, so cafe_gpu_score / cafe_gpu_objective(_device) return the same all-reduced result on every rank.  This is
 * the drop-in for the body of __cafe_best_lambda_search (cafe/lambda.cpp:726-769) on 1..8 GPUs.
 *
 * (a) one process per GPU (torchrun, mpirun): rank 0 calls cafe_gpu_comm_unique_id, the launcher hands the id_bytes (>= 128)
 *     bytes to every rank, every rank calls cafe_gpu_comm_init on its context and passes ITS slice of the families to
 *     cafe_gpu_set_families (first_index = indices into the whole table).
 * (b) one process, several devices: cafe_gpu_create_multi returns a leader context that owns one context per device
 *     (ncclCommInitAll).  Every call in this header made on the leader fans out: setters go to all devices,
 *     cafe_gpu_set_families splits the table into contiguous balanced slices in device order, per-family outputs come back
 *     in table order, the rows of the conditional distribution are split over the devices.  devices == NULL means 0..n-1.
 * libnccl.so.2 is loaded on first use (dlopen); contexts without a communicator never touch it. */
int cafe_gpu_comm_unique_id(void* id_out, int id_bytes);
int cafe_gpu_comm_init(cafe_gpu_ctx* ctx, const void* id, int id_bytes, int rank, int world);
int cafe_gpu_comm_size(const cafe_gpu_ctx* ctx);
int cafe_gpu_comm_rank(const cafe_gpu_ctx* ctx);
int cafe_gpu_create_multi(cafe_gpu_ctx** out, const int* devices, int n_devices);
int cafe_gpu_num_devices(const cafe_gpu_ctx* ctx);

/* K1 sharded across ranks with a collective of the CALLER's (when the library's own NCCL path above is not wanted).  After
 * cafe_gpu_set_key_shard(rank, world), cafe_gpu_build_matrices builds only this rank's contiguous chunk of keys_per_rank =
 * ceil(D / world) distinct keys; the caller all-gathers the chunks IN PLACE over the device buffer d_M returned by
 * cafe_gpu_matrix_storage (chunk r = doubles [r * keys_per_rank * doubles_per_key, +keys_per_rank * doubles_per_key)) ON THE
 * CONTEXT'S STREAM (cafe_gpu_set_stream) — or orders its own stream after the build and before the next call with events —
 * and calls cafe_gpu_matrices_exchanged, which transposes the received matrices into d_MT locally; scoring before that fails
 * with CAFE_GPU_ERR_STATE.  The reference builds all matrices in one `omp for` (libtree/birthdeath.c:331-343) — this is its
 * distributed equivalent.  world == 1 restores the unsharded behaviour. */
int cafe_gpu_set_key_shard(cafe_gpu_ctx* ctx, int rank, int world);
int cafe_gpu_matrix_storage(cafe_gpu_ctx* ctx, void** d_M, void** d_MT, int64_t* doubles_per_key, int32_t* keys_per_rank);
int cafe_gpu_matrices_exchanged(cafe_gpu_ctx* ctx);

/* Per-family results of the last score: log(max posterior), max likelihood, argmax_j L[j]
 * (the `maxlh` side effect, cafe/lambda.cpp:672-676).  Any pointer may be NULL. */
int cafe_gpu_family_results(cafe_gpu_ctx* ctx, double* log_max_posterior, double* max_likelihood,
                            int32_t* argmax_likelihood);
/* Root likelihood vectors of all families, row-major [n_families][root_max-root_min+1]
 * (get_likelihoods, cafe/cafe_tree.c:325-329).  Recomputes with the current matrices. */
int cafe_gpu_family_likelihoods(cafe_gpu_ctx* ctx, double* L_out);

/* Viterbi ancestral reconstruction (cafe_tree_viterbi, cafe/viterbi.cpp:494-521; max-product pruning :209-321 and
 * back-track :323-351) for every family with the current matrices: node_sizes_out[f * n_nodes + v] = reconstructed size of
 * node v (nlist order; leaves keep their observed size), max_likelihood_out[f] (nullable) = max_i of the root's max-product
 * vector.  A leaf with count -1 carries no data (the familysize < 0 branch, :236-250): its vector is all ones over its parent's
 * range - zeros beyond, as in a freshly allocated tree; the reference keeps there what an earlier family left - and its size
 * is reconstructed like an ancestor's. */
int cafe_gpu_viterbi(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* max_likelihood_out);
/* The same as the `report` command runs it (viterbi_section, cafe/viterbi.cpp:88-119): every family with its own forced range
 * (cafe_family_set_size_with_family_forced, cafe/cafe_family.c:236-255: root 1..rint(1.25*max_f), sizes 0..max_f+max(50,max_f/5)),
 * then viterbi_sum_probabilities (:42-70): branch_pvalues_out[f * n_nodes + c] (nullable) for the branch above node c =
 * sum over sizes m of the branch's transition row at the parent's reconstructed size, entries equal to the realised
 * transition probability counted half and smaller ones fully; -1 at the root.  (The reference stores the same numbers under
 * node id 2*j+k for child k of internal node 2j+1, and overwrites them with -1 when the family's max p-value exceeds the
 * cut-off - that filter is the caller's.) */
int cafe_gpu_viterbi_report(cafe_gpu_ctx* ctx, int32_t* node_sizes_out, double* branch_pvalues_out);

/* Branch-stretch likelihood-ratio test (cafe_likelihood_ratio_test / __cafe_likelihood_ratio_test_thread_func,
 * cafe/cafe_main.c:342-431, run by `report ... likelihood`, cafe/reports.cpp:684-685) for every family with the current rates and
 * matrices.  For each non-root branch b (nlist order) the branch is lengthened by rint(0.15 * length) for as long as the family's
 * maximum root likelihood grows (:377-386; the lengthened branch's matrix is keyed (int length, lambda_b, mu_b),
 * libtree/birthdeath.c:363-370).  Outputs, row-major [n_nodes][n_families]:
 *   best_max_likelihood_out[b][f] = the reference's `prevlh` when the loop stops (-1 in the root's row),
 *   steps_out[b][f] (nullable)    = number of lengthenings that raised the likelihood,
 *   base_max_likelihood_out[f] (nullable) = `maxlh` of the unlengthened tree.
 * The caller forms likelihoodRatios[b][f] = best == base ? 1 : 1 - chi2cdf(2 * (log best - log base), 1)   (:388).
 * tested (nullable = all): 1 tests a family, 2 tests it as a family that is not the table's first tested one (see below; for the
 * later shards of a table split over several contexts), 0 skips a family — the reference's `maximumPvalues[i] > param->pvalue` filter (:358-362), whose rows
 * the caller sets to -1; skipped families report best = base.  Like the reference, the first tested family starts from the parsed
 * branch lengths and all later ones from the (int)-truncated lengths (the length is restored through an int, :350,:390).
 * lengthened_mu_per_node (nullable = the nodes' own mu): the mu the LENGTHENED branch is keyed with.  The reference runs this test
 * on cafe_tree_copy(param->pcafe), and cafe_tree_node_copy (cafe/cafe_tree.c:485-494) copies lambda but not mu: the copy's nodes
 * carry the tree-level pcafe->mu (cafe_tree.c:39), which cafe_tree_new leaves at 0.  Passing zeros reproduces the stock binary's
 * numbers (keys (t, lambda, 0): log(alpha) = -inf, the NaN entries clamp to 1); NULL runs the algorithm as written.
 * Afterwards the context is back at the tree's own keys; cafe_gpu_family_results needs a new cafe_gpu_score. */
int cafe_gpu_likelihood_ratio_test(cafe_gpu_ctx* ctx, const uint8_t* tested, const double* lengthened_mu_per_node,
                                   double* base_max_likelihood_out, double* best_max_likelihood_out, int32_t* steps_out);

/* K4: conditional distribution (cafe/conditional_distribution.cpp:10-120): for every root size
 * s = root_min..root_max, n_samples simulated families (cafe/cafe_tree.c:533-569), each pruned with
 * root range {s} and the range.max ratchet of conditional_distribution.cpp:29; rows sorted ascending.
 * cd_out is row-major [root_max-root_min+1][n_samples].
 *   uniforms != NULL : "replay" — the stream of unifrnd() draws the single-threaded reference would
 *                      consume, [(root_max-root_min+1) * n_samples * (n_nodes-1)] doubles, in order
 *                      (s, trial, prefix-order non-root node); results match the reference draw for draw.
 *   uniforms == NULL : counter-based device RNG keyed by (seed, s, trial, node). */
int cafe_gpu_conditional_distribution(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms,
                                      uint64_t seed, double* cd_out);

/* Rows [row_lo, row_hi) of the same distribution (root sizes root_min + row), cd_out row-major [(row_hi-row_lo)][n_samples]:
 * the rows are independent, so ranks can split them and all-gather the result — the distributed form of the reference's
 * pthreads over root sizes (cafe/conditional_distribution.cpp:86-120).  `uniforms` still starts at row 0. */
int cafe_gpu_conditional_distribution_rows(cafe_gpu_ctx* ctx, int n_samples, const double* uniforms, uint64_t seed,
                                           int row_lo, int row_hi, double* cd_out);

/* K5: family-wide p-values (cafe/viterbi.cpp:88-97,32-39; cafe/pvalue.cpp:143-154;
 * libcommon/mathfunc.c:663-689): per family the forced range of cafe/cafe_family.c:236-255, prune,
 * p[s] = pvalue(L[s], cd[s]), result = max_s (0 when the family's root range is empty).
 * cd is [cd_rows][n_samples] ascending rows, row r = root size 1+r. */
int cafe_gpu_pvalues(cafe_gpu_ctx* ctx, const double* cd, int cd_rows, int n_samples,
                     double* max_pvalue_out);

/* Branch cutting, the p-value part (cafe/branch_cutting.cpp:20-44 p_values_of_two_trees, :101-150 compute_cutpvalues).  The caller
 * cuts a branch (phylogeny_split_tree, libtree/phylogeny.c:571-614), gives each side its own context (tree, rates, the families'
 * counts at that side's leaves), and passes the root likelihood rows of the two sides (cafe_gpu_family_likelihoods, row-major
 * [n_families][rfsize]) with their conditional distributions ([rfsize][cdlen] ascending rows, cafe_gpu_conditional_distribution):
 *   L_sub == cd_sub == NULL (one side is a single leaf):  out[f] = max_s pvalue(L_rest[f][s], cd_rest[s])
 *   otherwise:  out[f] = max(0, max_{s1,s2} (1/cdlen) sum_t pvalue(L_rest[f][s1] * L_sub[f][s2] / cd_sub[s2][t], cd_rest[s1]))
 * with the sum over t in ascending order, one rounding per operation as on the CPU.  `ctx` only names the device and stream. */
int cafe_gpu_cut_pvalues(cafe_gpu_ctx* ctx, const double* L_rest, const double* L_sub, int n_families, int rfsize,
                         const double* cd_rest, const double* cd_sub, int cdlen, double* cut_pvalue_out);

/* Bookkeeping for measurement: kernels launched by this context since creation / last reset; and,
 * when timing is enabled, one CUDA-event quad per objective evaluation recorded on the context's
 * stream around K1 (matrix build) and K2 (pruning + root reduction).  cafe_gpu_timing_collect
 * synchronises, writes the per-evaluation device times (ms) of the evaluations recorded since the
 * last collect (at most `cap`, at most the last 256) and returns how many it wrote. */
int64_t cafe_gpu_launch_count(const cafe_gpu_ctx* ctx);
void cafe_gpu_reset_launch_count(cafe_gpu_ctx* ctx);
int cafe_gpu_enable_timing(cafe_gpu_ctx* ctx, int on);
int cafe_gpu_timing_collect(cafe_gpu_ctx* ctx, float* k1_ms, float* k2_ms, int cap);
/* The same with the two exchange steps of a multi-GPU evaluation (zeros without a communicator): K1 (incl. the local
 * transposes), all-gather of the matrices, K2, reduction of the score.  Any pointer may be NULL. */
int cafe_gpu_timing_collect4(cafe_gpu_ctx* ctx, float* k1_ms, float* exchange_ms, float* k2_ms, float* reduce_ms, int cap);
/* Algorithmic fp64 flops of one cafe_gpu_score over the current families:
 * sum over internal edges of 2*W*W (2*R*W at the root), leaf edges excluded. */
double cafe_gpu_score_flops(const cafe_gpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CAFE_GPU_H */
