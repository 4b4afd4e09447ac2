/*
 * cafe_oracle.c — CPU restatement of CAFE's birth–death pruning likelihood path (see cafe_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 * Parity status: PINNED (tests/test_oracle.py: reference KATs + bitwise comparison with the compiled
 * reference in oracle/_ref/).
 *
 * This is a restatement, not a copy: the reference works on pointer-linked trees, lazily filled
 * caches and per-node heap buffers; here everything is flat arrays.  The ARITHMETIC (operation
 * order, formulas, clamps, special cases) follows the cited reference lines exactly, because the
 * point of the oracle is bit-level agreement with the reference on x86-64/glibc.
 */
#include "cafe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------ special functions */

/* Lanczos coefficients, libcommon/mathfunc.c:87-89 */
static const double k_lanczos[7] = {1.000000000190015,  76.18009172947146,     -86.50532032941677,
                                    24.01409824083091,  -1.231739572450155,    1.208650973866179e-3,
                                    -5.395239384953e-6};
#define ORC_SQRT_2PI 2.5066282746310002416123552393401042 /* mathfunc.c:103 */

double orc_gammaln(double a) /* mathfunc.c:112-119 */
{
    double p = k_lanczos[0];
    double a55 = a + 5.5;
    for (int n = 1; n <= 6; n++) p += k_lanczos[n] / (a + n);
    return (a + 0.5) * log(a55) - a55 + log(ORC_SQRT_2PI * p / a);
}

double orc_chooseln(double n, double r) /* mathfunc.c:224-229 */
{
    if (r == 0 || (n == 0 && r == 0)) return 0;
    if (n <= 0 || r <= 0) return log(0);
    return orc_gammaln(n + 1) - orc_gammaln(r + 1) - orc_gammaln(n - r + 1);
}

double orc_poisspdf(int x, double lambda) /* mathfunc.c:352-355 */
{
    return exp(x * log(lambda) - orc_gammaln(x + 1) - lambda);
}

double orc_unifrnd(void) /* mathfunc.c:91-94 */
{
    return rand() / (RAND_MAX + 1.0);
}

void orc_lnc_table(int size, double *out) /* chooseln_cache.h:27-41: values[n][x] = chooseln(n,x) */
{
    int cols = size + 1;
    for (int n = 0; n < 2 * size; n++)
        for (int x = 0; x <= size; x++) out[(size_t)n * cols + x] = orc_chooseln(n, x);
}

/* ------------------------------------------------------------------ transition matrix */

int orc_key_branchlength(double branchlength) /* cafe_tree.c:376 */
{
    return (int)branchlength;
}

/* birthdeath.c:52-73 — mu < 0 path */
static double bd_rate_log_alpha(int s, int c, double log_alpha, double coeff)
{
    int m = ORC_MIN(c, s);
    double lastterm = 1;
    double p = 0.0;
    for (int j = 0; j <= m; j++) {
        double t = orc_chooseln(s, j) + orc_chooseln(s + c - 1 - j, s - 1) + (s + c - 2 * j) * log_alpha;
        p += exp(t) * lastterm;
        lastterm *= coeff;
    }
    return ORC_MAX(ORC_MIN(p, 1), 0);
}

/* birthdeath.c:34-50 — mu >= 0 path */
static double bd_rate_log_alpha_beta(int s, int c, double log_alpha, double log_beta, double log_coeff)
{
    int m = ORC_MIN(c, s);
    double p = 0;
    for (int j = 0; j <= m; j++) {
        double t = orc_chooseln(s, j) + orc_chooseln(s + c - 1 - j, s - 1) + (s - j) * log_alpha +
                   (c - j) * log_beta + j * log_coeff;
        p += exp(t);
    }
    return ORC_MAX(ORC_MIN(p, 1), 0);
}

void orc_bd_matrix(double branchlength, double lambda, double mu, int maxfs, double *M) /* birthdeath.c:238-286 */
{
    int S = maxfs + 1;
    memset(M, 0, sizeof(double) * (size_t)S * S); /* memory_new is calloc (memalloc.c) */
    M[0] = 1; /* :244 */

    double alpha, beta, coeff;
    if (mu < 0 || lambda == mu) { /* :250-254 */
        alpha = lambda * branchlength / (1 + lambda * branchlength);
        beta = alpha;
        coeff = 1 - 2 * alpha;
    } else { /* :255-262 */
        double e_diff = exp((lambda - mu) * branchlength);
        double numerator = e_diff - 1;
        double denominator = lambda * e_diff - mu;
        alpha = (mu * numerator) / denominator;
        beta = (lambda * numerator) / denominator;
        coeff = 1 - alpha - beta;
    }
    /* init_matrix :211-225 — row 0 = [1,0,...]; coeff<=0: rows>=1 zero; coeff==1: identity */
    if (coeff <= 0) return;
    if (coeff == 1) {
        for (int s = 1; s < S; s++) M[(size_t)s * S + s] = 1;
        return;
    }
    double la = log(alpha), lb = log(beta), lc = log(coeff);
    for (int s = 1; s <= maxfs; s++)
        for (int c = 0; c <= maxfs; c++)
            M[(size_t)s * S + c] = (mu < 0) ? bd_rate_log_alpha(s, c, la, coeff) /* :272-273 */
                                            : bd_rate_log_alpha_beta(s, c, la, lb, lc); /* :274-275 */
}

/* ------------------------------------------------------------------ pruning */

void orc_matvec(const double *M, int S, const double *vec, int row_start, int row_end, int col_start,
                int col_end, double *result) /* birthdeath.c:172-180 */
{
    for (int s = row_start, i = 0; s <= row_end; s++, i++) {
        result[i] = 0;
        for (int c = col_start, j = 0; c <= col_end; c++, j++) result[i] += M[(size_t)s * S + c] * vec[j];
    }
}

typedef struct {
    int n_nodes;
    const int *left, *right;
    int root;
    const double *const *node_matrix;
    int S;
    const int *leaf_count;
    const double *const *leaf_err;
    int E;
    int rmin, rmax, root_min, root_max;
    int size_of_factor;
    double *lik; /* [n_nodes][size_of_factor] */
    double *f1, *f2;
    int err;
} prune_ctx;

static void prune_node(prune_ctx *cx, int v) /* cafe_tree.c:301-318 recursion, :191-271 per node */
{
    double *Lv = cx->lik + (size_t)v * cx->size_of_factor;
    if (cx->left[v] < 0) { /* initialize_leaf_likelihoods :191-211 */
        memset(Lv, 0, sizeof(double) * cx->size_of_factor);
        int fs = cx->leaf_count[v];
        if (cx->leaf_err && cx->leaf_err[v]) {
            const double *row = cx->leaf_err[v] + (size_t)fs * cx->E;
            for (int j = 0; j < cx->size_of_factor; j++) Lv[j] = (j < cx->E) ? row[j] : 0.0;
            /* NB: the reference reads errormatrix[fs][j] for all j < size_of_factor and would run
             * past the row when size_of_factor > E; the oracle fences that. */
        } else {
            if (fs < 0 || fs >= cx->size_of_factor) { cx->err = -1; return; }
            Lv[fs] = 1;
        }
        return;
    }
    int a = cx->left[v], b = cx->right[v];
    prune_node(cx, a);
    prune_node(cx, b);
    /* compute_internal_node_likelihood :226-271 */
    int r0, r1;
    if (v == cx->root) { r0 = cx->root_min; r1 = cx->root_max; }
    else               { r0 = cx->rmin;     r1 = cx->rmax; }
    orc_matvec(cx->node_matrix[a], cx->S, cx->lik + (size_t)a * cx->size_of_factor, r0, r1, cx->rmin, cx->rmax, cx->f1);
    orc_matvec(cx->node_matrix[b], cx->S, cx->lik + (size_t)b * cx->size_of_factor, r0, r1, cx->rmin, cx->rmax, cx->f2);
    int size = r1 - r0 + 1;
    for (int i = 0; i < size; i++) Lv[i] = cx->f1[i] * cx->f2[i];
}

int orc_prune(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
              int S, const int *leaf_count, const double *const *leaf_err, int E, int range_min,
              int range_max, int root_min, int root_max, double *L_root)
{
    prune_ctx cx;
    cx.n_nodes = n_nodes; cx.left = left; cx.right = right; cx.root = root;
    cx.node_matrix = node_matrix; cx.S = S; cx.leaf_count = leaf_count; cx.leaf_err = leaf_err; cx.E = E;
    cx.rmin = range_min; cx.rmax = range_max; cx.root_min = root_min; cx.root_max = root_max;
    int rsize = root_max - root_min + 1, fsize = range_max - range_min + 1;
    cx.size_of_factor = ORC_MAX(rsize, fsize); /* cafe_tree.c:56-58 */
    /* the reference sizes node buffers once from the *global* range; leaf counts up to S-1 must fit */
    if (cx.size_of_factor < S) cx.size_of_factor = S;
    cx.lik = (double *)calloc((size_t)n_nodes * cx.size_of_factor, sizeof(double));
    cx.f1 = (double *)calloc(cx.size_of_factor, sizeof(double));
    cx.f2 = (double *)calloc(cx.size_of_factor, sizeof(double));
    cx.err = 0;
    prune_node(&cx, root);
    if (!cx.err) memcpy(L_root, cx.lik + (size_t)root * cx.size_of_factor, sizeof(double) * rsize);
    free(cx.lik); free(cx.f1); free(cx.f2);
    return cx.err;
}

void orc_posterior(const double *L, const double *prior, int rfsize, double *max_likelihood,
                   double *max_posterior, int *argmax_likelihood) /* lambda.cpp:657-689 */
{
    double ml = L[0]; int am = 0;               /* __max/__maxidx: first maximum, mathfunc.c */
    for (int j = 1; j < rfsize; j++) if (L[j] > ml) { ml = L[j]; am = j; }
    double mp = exp(log(L[0]) + log(prior[0])); /* :678-686 (std::max_element: first maximum) */
    for (int j = 1; j < rfsize; j++) {
        double p = exp(log(L[j]) + log(prior[j]));
        if (p > mp) mp = p;
    }
    *max_likelihood = ml; *max_posterior = mp; *argmax_likelihood = am;
}

double orc_score(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
                 int S, const double *const *leaf_err, int E, int range_min, int range_max, int root_min,
                 int root_max, const double *prior, int F, const int *counts, const int *ref,
                 double *per_family_logpost, double *per_family_maxlik, int *per_family_argmax,
                 double *L_all, int *first_zero) /* lambda.cpp:691-724 */
{
    int rfsize = root_max - root_min + 1;
    int n_leaves = (n_nodes + 1) / 2;
    double *ml = (double *)malloc(sizeof(double) * F), *mp = (double *)malloc(sizeof(double) * F);
    int *am = (int *)malloc(sizeof(int) * F);
    int *leaf_count = (int *)malloc(sizeof(int) * n_nodes);
    double *L = (double *)malloc(sizeof(double) * rfsize);
    double score = 0;
    if (first_zero) *first_zero = -1;
    int zero_seen = 0;
    for (int i = 0; i < F; i++) {
        if (!ref || ref[i] < 0 || ref[i] == i) {
            for (int k = 0; k < n_nodes; k++) leaf_count[k] = -1;
            for (int k = 0; k < n_leaves; k++) leaf_count[2 * k] = counts[(size_t)i * n_leaves + k];
            orc_prune(n_nodes, left, right, root, node_matrix, S, leaf_count, leaf_err, E, range_min, range_max,
                      root_min, root_max, L);
            orc_posterior(L, prior, rfsize, &ml[i], &mp[i], &am[i]);
            if (L_all) memcpy(L_all + (size_t)i * rfsize, L, sizeof(double) * rfsize);
        } else {
            ml[i] = ml[ref[i]]; mp[i] = mp[ref[i]]; am[i] = am[ref[i]];
            if (L_all) memcpy(L_all + (size_t)i * rfsize, L_all + (size_t)ref[i] * rfsize, sizeof(double) * rfsize);
        }
        if (ml[i] == 0 && !zero_seen) { /* :715-720: throw => caller scores log(0) */
            zero_seen = 1;
            if (first_zero) *first_zero = i;
        }
        if (per_family_logpost) per_family_logpost[i] = log(mp[i]);
        if (per_family_maxlik) per_family_maxlik[i] = ml[i];
        if (per_family_argmax) per_family_argmax[i] = am[i];
        score += log(mp[i]);
    }
    free(ml); free(mp); free(am); free(leaf_count); free(L);
    if (zero_seen) return log(0); /* lambda.cpp:753-760 */
    return score;
}

/* ------------------------------------------------------------------ conditional distribution */

int orc_random_familysize(int n_nodes, const int *left, const int *right, int root,
                          const double *const *node_matrix, int S, int root_size, int max_family_size,
                          const double *uniforms, long *n_used, int *sizes) /* cafe_tree.c:533-569 */
{
    int max = 0;
    int *stack = (int *)malloc(sizeof(int) * (n_nodes + 2));
    int *parent = (int *)malloc(sizeof(int) * n_nodes);
    for (int i = 0; i < n_nodes; i++) parent[i] = -1;
    for (int i = 0; i < n_nodes; i++) if (left[i] >= 0) { parent[left[i]] = i; parent[right[i]] = i; }
    sizes[root] = root_size;
    int sp = 0;
    stack[sp++] = root;
    while (sp > 0) { /* tree_traveral_prefix, tree.c:101-124: node, then head subtree, then tail subtree */
        int v = stack[--sp];
        if (left[v] >= 0) { stack[sp++] = right[v]; stack[sp++] = left[v]; }
        if (v == root) continue;
        double rnd = uniforms ? uniforms[(*n_used)++] : (++(*n_used), orc_unifrnd());
        double cumul = 0;
        int ps = sizes[parent[v]];
        int c = 0;
        for (; c < max_family_size - 1; c++) {
            cumul += node_matrix[v][(size_t)ps * S + c];
            if (cumul >= rnd) break;
        }
        sizes[v] = c;
        if (max < c) max = c;
    }
    free(stack); free(parent);
    return max;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

void orc_random_probabilities(int n_nodes, const int *left, const int *right, int root,
                              const double *const *node_matrix, int S, int range_min, int range_max,
                              int root_size, int trials, const double *uniforms, long *n_used,
                              double *probs_sorted, double *probs_unsorted, int *leaf_sizes, int *caps,
                              const double *const *leaf_err, int E)
/* conditional_distribution.cpp:10-44.  leaf_err (nullable): the error models attached to the tree's leaves - compute_tree_likelihoods
 * applies them to the simulated leaf sizes exactly as to observed ones (cafe_tree.c:196-203). */
{
    int rmax = range_max;
    int maxFamilySize = ORC_MAX(root_size, range_max); /* :20 (root_max == root_size here) */
    int *sizes = (int *)malloc(sizeof(int) * n_nodes);
    for (int i = 0; i < trials; i++) {
        int max = orc_random_familysize(n_nodes, left, right, root, node_matrix, S, root_size, maxFamilySize,
                                        uniforms, n_used, sizes);
        rmax = ORC_MIN(max + ORC_MAX(50, max / 5), rmax); /* :29 — the range.max ratchet */
        if (leaf_sizes) memcpy(leaf_sizes + (size_t)i * n_nodes, sizes, sizeof(int) * n_nodes);
        if (caps) caps[i] = rmax;
        double L0 = 0;
        /* a simulated leaf size above the (ratcheted) column window contributes 0 (cafe_tree.c:223);
         * in the reference the one-hot lands outside cols min..max of the matvec. */
        orc_prune(n_nodes, left, right, root, node_matrix, S, sizes, leaf_err, E, range_min, rmax, root_size, root_size, &L0);
        probs_sorted[i] = L0;
    }
    if (probs_unsorted) memcpy(probs_unsorted, probs_sorted, sizeof(double) * trials);
    qsort(probs_sorted, trials, sizeof(double), cmp_double); /* :41 std::sort ascending */
    free(sizes);
}

double orc_pvalue(double v, const double *cd, int size) /* mathfunc.c:663-689 */
{
    int from = 0, to = size - 1;
    while (from < to) {
        int mi = from + (to - from) / 2;
        if (cd[mi] > v) to = mi - 1;
        else if (cd[mi] < v) from = mi + 1;
        else {
            for (from = mi - 1; from >= 0 && cd[from] == v; from--) ;
            for (to = mi + 1; to < size && cd[to] == v; to++) ;
            from++; to--;
            break;
        }
    }
    if (from > to) to = from;
    return (double)(from + (cd[from] <= v ? 1 : 0) + (to - from) / 2.0) / (double)size;
}

void orc_init_family_size(int max, int *root_min, int *root_max, int *min, int *max_out) /* cafe_family.c:357-364 */
{
    *root_min = 1;
    *root_max = (int)ORC_MAX(30, rint(max * 1.25));
    *max_out = max + ORC_MAX(50, max / 5);
    *min = 0;
}

double orc_family_pvalue(int n_nodes, const int *left, const int *right, int root,
                         const double *const *node_matrix, int S, const double *const *leaf_err, int E,
                         const int *leaf_count_by_node, const double *cd, int cd_rows, int n_samples,
                         double *pvalues_out, int *rfsize_out)
{
    int max = 0; /* cafe_family.c:236-255 */
    for (int i = 0; i < n_nodes; i += 2) if (max < leaf_count_by_node[i]) max = leaf_count_by_node[i];
    int root_min = 1, root_max = (int)rint(max * 1.25), rmax = max + ORC_MAX(50, max / 5);
    int rfsize = root_max - root_min + 1;
    if (rfsize_out) *rfsize_out = rfsize;
    if (rfsize <= 0) return 0; /* viterbi.cpp:32-39 — empty => 0 */
    double *L = (double *)malloc(sizeof(double) * rfsize);
    orc_prune(n_nodes, left, right, root, node_matrix, S, leaf_count_by_node, leaf_err, E, 0, rmax, root_min, root_max, L);
    double best = 0; int first = 1;
    for (int s = 0; s < rfsize && s < cd_rows; s++) { /* pvalue.cpp:149-153 */
        double p = orc_pvalue(L[s], cd + (size_t)s * n_samples, n_samples);
        if (pvalues_out) pvalues_out[s] = p;
        if (first || p > best) { best = p; first = 0; }
    }
    free(L);
    return best;
}

/* ------------------------------------------------------------------ Viterbi ancestral reconstruction */

/* cafe_tree_viterbi (cafe/viterbi.cpp:494-521): max-product pruning in post-order
 * (__cafe_tree_node_compute_viterbi :209-321) followed by the back-track in prefix order
 * (__cafe_tree_node_backtrack_viterbi :323-351).  All leaves carry an observed size (the reference's
 * a leaf with leaf_count < 0 has no data: the "familysize < 0" branch, restated for a fresh tree, see vit_node).  sizes_out[v] = reconstructed size of node v
 * (leaves keep their observed size); *root_max_lik = max_i L_root[i].
 * Per child: factor[i] = max_j M_child[s][c] * L_child[j] with a strict ">" from 0 (first maximum wins; a row
 * whose products are all 0 keeps back-pointer 0, the calloc'ed initial value of pcnode->viterbi). */
typedef struct {
    int n_nodes; const int *left, *right; int root;
    const double *const *node_matrix; int S;
    const int *leaf_count; const double *const *leaf_err; int E;
    int rmin, rmax, root_min, root_max, size_of_factor;
    double *lik; int *vit; double *f[2];
} vit_ctx;

static void vit_node(vit_ctx *cx, int v, int parent_is_root)
{
    double *Lv = cx->lik + (size_t)v * cx->size_of_factor;
    if (cx->left[v] < 0) { /* leaf, :251-267 */
        memset(Lv, 0, sizeof(double) * cx->size_of_factor);
        int fs = cx->leaf_count[v];
        if (fs < 0) {
            /* no data for this species (:236-250): likelihoods[i] = 1 for the sizes of the PARENT's range - the root range when
             * the parent is the root (:221-230) - and whatever the array held before elsewhere: zeros in a fresh tree, which is
             * what is restated here.  (The leaf's own factors / viterbi pointers of :240-249 are overwritten by the parent.) */
            int n1 = parent_is_root ? (cx->root_max - cx->root_min + 1) : (cx->rmax - cx->rmin + 1);
            for (int i = 0; i < n1 && i < cx->size_of_factor; i++) Lv[i] = 1;
            return;
        }
        if (cx->leaf_err && cx->leaf_err[v]) {
            for (int j = 0; j < cx->size_of_factor; j++) Lv[j] = (fs < cx->E && j < cx->E) ? cx->leaf_err[v][(size_t)fs * cx->E + j] : 0.0;
        } else {
            Lv[fs] = 1;
        }
        return;
    }
    int child[2] = { cx->left[v], cx->right[v] };
    vit_node(cx, child[0], v == cx->root);
    vit_node(cx, child[1], v == cx->root);
    int r0 = (v == cx->root) ? cx->root_min : cx->rmin, r1 = (v == cx->root) ? cx->root_max : cx->rmax; /* :273-286 */
    for (int idx = 0; idx < 2; idx++) { /* :291-311 */
        const double *M = cx->node_matrix[child[idx]];
        const double *Lc = cx->lik + (size_t)child[idx] * cx->size_of_factor;
        int *vc = cx->vit + (size_t)child[idx] * cx->size_of_factor;
        memset(cx->f[idx], 0, sizeof(double) * cx->size_of_factor);
        for (int s = r0, i = 0; s <= r1; s++, i++) {
            for (int c = cx->rmin, j = 0; c <= cx->rmax; c++, j++) {
                double tmp = M[(size_t)s * cx->S + c] * Lc[j];
                if (tmp > cx->f[idx][i]) { cx->f[idx][i] = tmp; vc[i] = j; }
            }
        }
    }
    int size = r1 - r0 + 1;
    for (int i = 0; i < size; i++) Lv[i] = cx->f[0][i] * cx->f[1][i]; /* :313-317 */
}

int orc_viterbi(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
                int S, const int *leaf_count, const double *const *leaf_err, int E, int range_min,
                int range_max, int root_min, int root_max, int *sizes_out, double *root_max_lik)
{
    vit_ctx cx;
    cx.n_nodes = n_nodes; cx.left = left; cx.right = right; cx.root = root;
    cx.node_matrix = node_matrix; cx.S = S; cx.leaf_count = leaf_count; cx.leaf_err = leaf_err; cx.E = E;
    cx.rmin = range_min; cx.rmax = range_max; cx.root_min = root_min; cx.root_max = root_max;
    int rsize = root_max - root_min + 1, fsize = range_max - range_min + 1;
    cx.size_of_factor = ORC_MAX(rsize, fsize);
    if (cx.size_of_factor < S) cx.size_of_factor = S;
    for (int v = 0; v < n_nodes; v += 2)
        if (leaf_count[v] >= cx.size_of_factor) return 1; /* leaf_count < 0: a species without data (missing) */
    cx.lik = (double *)calloc((size_t)n_nodes * cx.size_of_factor, sizeof(double));
    cx.vit = (int *)calloc((size_t)n_nodes * cx.size_of_factor, sizeof(int));
    cx.f[0] = (double *)calloc(cx.size_of_factor, sizeof(double));
    cx.f[1] = (double *)calloc(cx.size_of_factor, sizeof(double));
    vit_node(&cx, root, 0);
    /* back-track in prefix order: node, head subtree, tail subtree (tree.c:101-124) */
    int *parent = (int *)malloc(sizeof(int) * n_nodes), *stack = (int *)malloc(sizeof(int) * (n_nodes + 2));
    for (int i = 0; i < n_nodes; i++) parent[i] = -1;
    for (int i = 0; i < n_nodes; i++) if (left[i] >= 0) { parent[left[i]] = i; parent[right[i]] = i; }
    int sp = 0;
    stack[sp++] = root;
    while (sp > 0) {
        int v = stack[--sp];
        if (left[v] >= 0) { stack[sp++] = right[v]; stack[sp++] = left[v]; }
        if (left[v] < 0 && leaf_count[v] >= 0) { sizes_out[v] = leaf_count[v]; continue; } /* :327: observed leaves keep their size */
        if (v == root) { /* :329-339, __maxidx: first maximum */
            const double *L = cx.lik + (size_t)root * cx.size_of_factor;
            int am = 0; double ml = L[0];
            for (int i = 1; i < rsize; i++) if (L[i] > ml) { ml = L[i]; am = i; }
            sizes_out[v] = root_min + am;
            if (root_max_lik) *root_max_lik = ml;
        } else { /* :341-350 */
            int p = parent[v];
            int base = (p == root) ? root_min : range_min;
            sizes_out[v] = cx.vit[(size_t)v * cx.size_of_factor + (sizes_out[p] - base)] + range_min;
        }
    }
    free(parent); free(stack); free(cx.lik); free(cx.vit); free(cx.f[0]); free(cx.f[1]);
    return 0;
}

/* viterbi_sum_probabilities (cafe/viterbi.cpp:42-70): for the branch above every non-root node c with parent size ps and
 * child size cs (both from the Viterbi reconstruction), p = M_c[ps][cs] and the branch p-value is
 * sum_{m=0..range_max} ( M_c[ps][m] == p ? M_c[ps][m] / 2 : M_c[ps][m] < p ? M_c[ps][m] : 0 ).  out[c], -1 at the root. */
void orc_viterbi_branch_pvalues(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix,
                                int S, const int *sizes, int range_max, double *out)
{
    for (int v = 0; v < n_nodes; v++) out[v] = -1;
    for (int v = 0; v < n_nodes; v++) {
        if (left[v] < 0) continue;
        int child[2] = { left[v], right[v] };
        for (int k = 0; k < 2; k++) {
            const double *M = node_matrix[child[k]];
            double p = M[(size_t)sizes[v] * S + sizes[child[k]]];
            double acc = 0;
            for (int m = 0; m <= range_max; m++) {
                double x = M[(size_t)sizes[v] * S + m];
                if (x == p) acc += x / 2.0;
                else if (x < p) acc += x;
            }
            out[child[k]] = acc;
        }
    }
    (void)root;
}

/* ------------------------------------------------------------------ likelihood-ratio test (branch stretch) */

#define ORC_EPS 1e-8 /* libcommon/mathfunc.h:12 */

/* incgammaln_lower + gammaincln + gammainc + gamcdf + chi2cdf, libcommon/mathfunc.c:128-151,260-263,284-287:
 * series only, at most 999 terms, stop at the first term below 1e-8; a series that never stops gives exactly 1. */
double orc_chi2cdf(double x, int df)
{
    const double a = df / 2.0;
    const double xs = x / 2;
    double p = 1 / a, t = 1 / a;
    int i;
    for (i = 1; i < 1000; i++) {
        t *= xs / (a + i);
        if (t < ORC_EPS) break;
        p += t;
    }
    double lower = (i == 1000) ? orc_gammaln(a) : log(p) + a * log(xs) - xs;
    return exp(lower - orc_gammaln(a));
}

/* birthdeath_cache_get_matrix, libtree/birthdeath.c:363-382: matrices of lengthened branches are memoised by (int t, lambda, mu)
 * (the sequence of lengths does not depend on the family, so every family after the first hits the cache). */
typedef struct { int t, maxfs; double lambda, mu; double *M; } orc_cached_matrix;
static orc_cached_matrix *g_lrt_cache = NULL;
static int g_lrt_cache_n = 0, g_lrt_cache_cap = 0;

void orc_lrt_cache_clear(void)
{
    for (int i = 0; i < g_lrt_cache_n; i++) free(g_lrt_cache[i].M);
    free(g_lrt_cache);
    g_lrt_cache = NULL;
    g_lrt_cache_n = g_lrt_cache_cap = 0;
}

static const double *lrt_cached_matrix(int t, double lambda, double mu, int maxfs)
{
    for (int i = 0; i < g_lrt_cache_n; i++) {
        const orc_cached_matrix *c = &g_lrt_cache[i];
        if (c->t == t && c->maxfs == maxfs && c->lambda == lambda && c->mu == mu) return c->M;
    }
    if (g_lrt_cache_n == g_lrt_cache_cap) {
        g_lrt_cache_cap = g_lrt_cache_cap ? 2 * g_lrt_cache_cap : 64;
        g_lrt_cache = (orc_cached_matrix *)realloc(g_lrt_cache, sizeof(orc_cached_matrix) * (size_t)g_lrt_cache_cap);
    }
    orc_cached_matrix *c = &g_lrt_cache[g_lrt_cache_n++];
    c->t = t; c->maxfs = maxfs; c->lambda = lambda; c->mu = mu;
    c->M = (double *)malloc(sizeof(double) * (size_t)(maxfs + 1) * (maxfs + 1));
    orc_bd_matrix((double)t, lambda, mu, maxfs, c->M);
    return c->M;
}

/* __cafe_likelihood_ratio_test_thread_func, cafe/cafe_main.c:342-396, body for ONE family that passed the
 * maximumPvalues filter (:358-362).  For every non-root node b in nlist order: start from the family's maximum root likelihood,
 * lengthen the branch above b by rint(0.15 * length) (:380) for as long as the maximum root likelihood grows (:377-386; the matrix
 * of the lengthened branch is keyed by the (int) length, libtree/birthdeath.c:363-370), then
 *   ratio[b] = prev == maxlh ? 1 : 1 - chi2cdf(2 * (log(prev) - log(maxlh)), 1)     (:388), -1 at the root (:369-373).
 * branchlength[] is IN/OUT like the thread's tree copy: the reference restores the length through an `int old_bl` (:350,:375,:390),
 * so after the first family every non-root length is truncated.  best_out / steps_out (optional): prevlh at the stop and the
 * number of lengthenings that raised the likelihood. */
int orc_lrt_family(int n_nodes, const int *left, const int *right, int root, const double *const *node_matrix, int S,
                   const double *lambda, const double *mu, double *branchlength, const int *leaf_count,
                   const double *const *leaf_err, int E, int range_min, int range_max, int root_min, int root_max,
                   double *ratios_out, double *best_out, int *steps_out)
{
    /* mu[] is the mu the LENGTHENED branch is keyed with.  The stock reference runs on cafe_tree_copy(param->pcafe), whose nodes
     * carry the tree-level pcafe->mu (cafe_tree.c:39; cafe_tree_node_copy :485-494 copies lambda only) = 0 after cafe_tree_new, not
     * the node's own mu: pass zeros to restate the stock binary, the nodes' mu to restate the algorithm as written. */
    const int rf = root_max - root_min + 1;
    double *L = (double *)malloc(sizeof(double) * (size_t)rf);
    const double **mats = (const double **)malloc(sizeof(double *) * (size_t)n_nodes);
    memcpy(mats, node_matrix, sizeof(double *) * (size_t)n_nodes);
    int rc = orc_prune(n_nodes, left, right, root, mats, S, leaf_count, leaf_err, E, range_min, range_max, root_min, root_max, L);
    if (rc) { free(L); free(mats); return rc; }
    double maxlh = L[0];
    for (int i = 1; i < rf; i++) if (L[i] > maxlh) maxlh = L[i]; /* __max, libcommon/mathfunc.c */
    for (int b = 0; b < n_nodes; b++) {
        if (b == root) {
            ratios_out[b] = -1;
            if (best_out) best_out[b] = -1;
            if (steps_out) steps_out[b] = 0;
            continue;
        }
        const int old_bl = (int)branchlength[b];
        double prevlh = -1, nextlh = maxlh;
        int steps = -1;
        while (prevlh < nextlh) {
            prevlh = nextlh;
            steps++;
            branchlength[b] += rint(branchlength[b] * 0.15);
            mats[b] = lrt_cached_matrix(orc_key_branchlength(branchlength[b]), lambda[b], mu[b], S - 1);
            orc_prune(n_nodes, left, right, root, mats, S, leaf_count, leaf_err, E, range_min, range_max, root_min, root_max, L);
            nextlh = L[0];
            for (int i = 1; i < rf; i++) if (L[i] > nextlh) nextlh = L[i];
        }
        ratios_out[b] = (prevlh == maxlh) ? 1 : 1 - orc_chi2cdf(2 * (log(prevlh) - log(maxlh)), 1);
        if (best_out) best_out[b] = prevlh;
        if (steps_out) steps_out[b] = steps;
        branchlength[b] = old_bl;
        mats[b] = node_matrix[b];
    }
    free(L); free(mats);
    return 0;
}
