/*
 * cafe_oracle.h — CPU restatement of CAFE's birth–death pruning likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by the
 * product path (cafe_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / baseline.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (a) the known-answer vectors of the reference's own unit tests, and
 *   (b) the unmodified reference compiled by oracle/Makefile into oracle/_ref/libcafe_ref.so
 *       (bit-for-bit on matrices, per-family likelihoods, scores, CD rows and p-values).
 *
 * Every function cites the reference file:line it restates (paths relative to the reference root).
 * Trees are passed flat: nodes are numbered in the reference's nlist (infix) order — leaves at even
 * indices, internal nodes at odd indices (cafe/cafe_commands.cpp:2028-2051).
 */
#ifndef CAFE_ORACLE_H
#define CAFE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* libcommon/mathfunc.c:87-89,112-119 — 6-term Lanczos log-gamma (NOT libm lgamma). */
double orc_gammaln(double a);
/* libcommon/mathfunc.c:224-229 */
double orc_chooseln(double n, double r);
/* libcommon/mathfunc.c:352-355 */
double orc_poisspdf(int x, double lambda);
/* libcommon/mathfunc.c:91-94 — glibc rand()/(RAND_MAX+1.0) */
double orc_unifrnd(void);

/* libtree/chooseln_cache.h:27-41 — dense table of lnC(n,x), n < 2*size, x <= size,
 * out is row-major [2*size][size+1].  (The reference fills lazily; values are identical.) */
void orc_lnc_table(int size, double *out);

/* libtree/birthdeath.c:238-286 (+ :34-73, :211-225).  M is row-major (maxfs+1)^2,
 * M[s*(maxfs+1)+c] = P(child = c | parent = s).  `branchlength` is used as given: the caller
 * applies the (int) truncation of cafe/cafe_tree.c:376 (see orc_key_branchlength). */
void orc_bd_matrix(double branchlength, double lambda, double mu, int maxfs, double *M);
/* cafe/cafe_tree.c:374-391 — the cache key truncates the branch length to int. */
int orc_key_branchlength(double branchlength);

/* libtree/birthdeath.c:163-182 (non-BLAS branch) */
void orc_matvec(const double *M, int S, const double *vec, int row_start, int row_end,
                int col_start, int col_end, double *result);

/* cafe/cafe_tree.c:191-323 — post-order pruning for ONE family.
 *   n_nodes, left[], right[] : flat binary tree (children head/tail), -1 for leaves; `root` index.
 *   node_matrix[i]           : S x S transition matrix of the branch above node i (ignored for root)
 *   leaf_count[i]            : observed size at leaf i (ignored for internal nodes)
 *   leaf_err[i]              : NULL, or row-major E x E errormatrix[observed][true] for leaf i
 *                              (cafe/cafe_tree.c:196-203)
 *   range                    : min,max (non-root rows/cols), root_min,root_max (root rows)
 *   L_root                   : out, root_max-root_min+1 values (cafe/cafe_tree.c:325-329)
 * Returns 0, or -1 if a leaf count is outside [0, size_of_factor) (the reference asserts). */
int orc_prune(int n_nodes, const int *left, const int *right, int root,
              const double *const *node_matrix, int S,
              const int *leaf_count, const double *const *leaf_err, int E,
              int range_min, int range_max, int root_min, int root_max,
              double *L_root);

/* cafe/lambda.cpp:657-689 */
void orc_posterior(const double *L_root, const double *prior, int rfsize,
                   double *max_likelihood, double *max_posterior, int *argmax_likelihood);

/* cafe/lambda.cpp:691-724 — score over F families with the `ref` (duplicate) short-cut.
 *   counts      : [F][n_leaves] in leaf order = even nlist indices 0,2,4,…
 *   ref         : NULL, or per family the index of the first identical family (cafe_family.c:9-34)
 *   per_family_logpost / per_family_maxlik / per_family_argmax : optional outputs [F]
 *   *first_zero : index of the first family with max_likelihood == 0 (then score = -inf), else -1
 * Returns the score Σ_f log(max_posterior_f). */
double orc_score(int n_nodes, const int *left, const int *right, int root,
                 const double *const *node_matrix, int S,
                 const double *const *leaf_err, int E,
                 int range_min, int range_max, int root_min, int root_max,
                 const double *prior,
                 int F, const int *counts, const int *ref,
                 double *per_family_logpost, double *per_family_maxlik, int *per_family_argmax,
                 double *L_all /* optional [F][rfsize] */, int *first_zero);

/* cafe/cafe_tree.c:533-569 — simulate sizes down the tree (prefix order, one unifrnd() per
 * non-root node).  sizes[] (per node) is filled; returns the max simulated non-root size.
 * `uniforms` may be NULL (draw from glibc rand()) or a pre-drawn stream consumed in order;
 * *n_used is advanced. */
int orc_random_familysize(int n_nodes, const int *left, const int *right, int root,
                          const double *const *node_matrix, int S,
                          int root_size, int max_family_size,
                          const double *uniforms, long *n_used, int *sizes);

/* cafe/conditional_distribution.cpp:10-44 — one row of the conditional distribution:
 * `trials` simulated families at root size s, pruned with root range {s} and the range.max
 * ratchet of :29, sorted ascending. */
void orc_random_probabilities(int n_nodes, const int *left, const int *right, int root,
                              const double *const *node_matrix, int S,
                              int range_min, int range_max, int root_size, int trials,
                              const double *uniforms, long *n_used, double *probs_sorted,
                              double *probs_unsorted /* optional */, int *leaf_sizes /* optional [trials][n_nodes] */,
                              int *caps /* optional [trials] */,
                              const double *const *leaf_err, int E);

/* libcommon/mathfunc.c:663-689 */
double orc_pvalue(double v, const double *conddist, int size);

/* cafe/cafe_family.c:357-364 */
void orc_init_family_size(int max, int *root_min, int *root_max, int *min, int *max_out);

/* cafe/cafe_family.c:236-255 + cafe/pvalue.cpp:143-154 + cafe/viterbi.cpp:32-39 — family p-value:
 * per-family range (root 1..rint(1.25*max_f), cols 0..max_f+max(50,max_f/5)), prune,
 * p[s] = pvalue(L[s], cd[s], n_samples), return max_s (0 when the root range is empty).
 * cd is row-major [cd_rows][n_samples], row r = root size 1+r. */
double orc_family_pvalue(int n_nodes, const int *left, const int *right, int root,
                         const double *const *node_matrix, int S,
                         const double *const *leaf_err, int E,
                         const int *leaf_count_by_node,
                         const double *cd, int cd_rows, int n_samples,
                         double *pvalues_out /* optional [rfsize_f] */, int *rfsize_out);

/* cafe_tree_viterbi, cafe/viterbi.cpp:209-351,494-521 (every leaf observed).  Returns non-zero when a leaf count does not fit. */
int orc_viterbi(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
                int S, const int *leaf_count, const double *const *leaf_err, int E, int range_min,
                int range_max, int root_min, int root_max, int *sizes_out, double *root_max_lik);

/* viterbi_sum_probabilities, cafe/viterbi.cpp:42-70; out[c] for the branch above node c, -1 at the root */
void orc_viterbi_branch_pvalues(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
                                int S, const int *sizes, int range_max, double *out);

/* libcommon/mathfunc.c:128-151,260-263,284-287 — chi2cdf through the series-only incomplete gamma */
double orc_chi2cdf(double x, int df);

/* __cafe_likelihood_ratio_test_thread_func, cafe/cafe_main.c:342-396, for ONE family that passed the p-value filter: per non-root
 * branch, lengthen by rint(0.15*length) while the maximum root likelihood grows; ratios_out[b] = 1 or
 * 1 - chi2cdf(2*(log best - log base), 1), -1 at the root.  branchlength[] is in/out (the reference restores lengths through an int). */
void orc_lrt_cache_clear(void); /* frees the memoised matrices of lengthened branches (birthdeath.c:363-382) */
int orc_lrt_family(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix, int S,
                   const double *lambda, const double *mu, double *branchlength, const int *leaf_count,
                   const double *const *leaf_err, int E, int range_min, int range_max, int root_min, int root_max,
                   double *ratios_out, double *best_out /* optional */, int *steps_out /* optional */);

#ifdef __cplusplus
}
#endif
#endif
