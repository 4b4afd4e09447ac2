// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" doors into the UNMODIFIED reference objects (compiled by oracle/Makefile from the
// sources under /root/reference into oracle/_ref/).  This file is OUR code: it only calls the
// reference's public functions so that tests/ and tests/golden/make_golden.py can drive the reference
// through ctypes and pin the oracle (oracle/cafe_oracle.c) and the CUDA path against it.
// It is linked into oracle/_ref/libcafe_ref.so, never into the product.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
#include "cafe.h"
#include "cafe_shell.h"
#include <family.h>
#include <mathfunc.h>
#include <chooseln_cache.h>
extern pBirthDeathCacheArray probability_cache;
extern struct chooseln_cache cache;
int __check_error_model_columnsums(pErrorStruct errormodel);
void __cafe_famliy_check_the_pattern(pCafeFamily pcf);
}
#include "gene_family.h"
#include "lambda.h"
#include "conditional_distribution.h"
#include "pvalue.h"
#include "viterbi.h"
#include "error_model.h"
#include "branch_cutting.h"

namespace {
struct Session {
    pCafeTree tree = nullptr;
    family_size_range range;
    pCafeFamily family = nullptr;
    std::vector<pErrorStruct> errs;
};
family_size_range make_range(int rmin, int rmax, int root_min, int root_max) {
    family_size_range r; r.min = rmin; r.max = rmax; r.root_min = root_min; r.root_max = root_max; return r;
}
// The reference's chooseln_cache_resize2 keeps rows allocated at the OLD width (chooseln_cache.c:36-56 vs
// chooseln_cache.h:33-37), so growing the cache reads past them.  The shim never grows: it frees and
// re-initialises, which is what a fresh reference process sees.
void ensure_lnc(int size) {
    if (chooseln_is_init2(&cache) && get_chooseln_cache_size2(&cache) >= size) return;
    if (chooseln_is_init2(&cache)) chooseln_cache_free2(&cache);
    chooseln_cache_init2(&cache, size);
}
pCafeNode node_at(Session* s, int i) { return (pCafeNode)s->tree->super.nlist->array[i]; }
}  // namespace

extern "C" {

double refshim_gammaln(double a) { return gammaln(a); }
double refshim_chooseln(double n, double r) { return chooseln(n, r); }
double refshim_poisspdf(int x, double l) { return poisspdf(x, l); }
double refshim_pvalue(double v, const double* cd, int n) { return pvalue(v, cd, n); }
void refshim_srand(unsigned seed) { srand(seed); }
double refshim_unifrnd() { return unifrnd(); }

void refshim_init_family_size(int max, int* out4) {
    family_size_range r; init_family_size(&r, max);
    out4[0] = r.root_min; out4[1] = r.root_max; out4[2] = r.min; out4[3] = r.max;
}

// compute_birthdeath_rates (libtree/birthdeath.c:238) with the global lnC cache sized as the
// reference does in birthdeath_cache_init (birthdeath.c:331-343).
void refshim_bd_matrix(double t, double lambda, double mu, int maxfs, double* out) {
    ensure_lnc(maxfs);
    struct square_matrix* m = compute_birthdeath_rates(t, lambda, mu, maxfs);
    memcpy(out, m->values, sizeof(double) * (size_t)m->size * m->size);
    free(m->values); free(m);
}

double refshim_bd_rate_log_alpha(int s, int c, double log_alpha, double coeff, int cache_size) {
    ensure_lnc(cache_size);
    return birthdeath_rate_with_log_alpha(s, c, log_alpha, coeff, &cache);
}
double refshim_bd_likelihood_s_c(int s, int c, double t, double lambda, double mu, int cache_size) {
    ensure_lnc(cache_size);
    return birthdeath_likelihood_with_s_c(s, c, t, lambda, mu, &cache);
}

void refshim_matvec(const double* M, int S, const double* vec, int r0, int r1, int c0, int c1, double* out) {
    struct square_matrix m; m.values = const_cast<double*>(M); m.size = S;
    square_matrix_multiply(&m, const_cast<double*>(vec), r0, r1, c0, c1, out);
}

// ---------------------------------------------------------------- tree sessions
void* refshim_session_new(const char* newick, int rmin, int rmax, int root_min, int root_max) {
    Session* s = new Session();
    s->range = make_range(rmin, rmax, root_min, root_max);
    s->tree = cafe_tree_new(newick, &s->range, 0, 0);
    if (!s->tree) { delete s; return nullptr; }
    return s;
}
int refshim_n_nodes(void* h) { return ((Session*)h)->tree->super.nlist->size; }

// nlist-order description: children (head/tail), parent, branch length, '\n'-joined names
void refshim_describe(void* h, int* left, int* right, int* parent, double* bl, char* names, int names_len) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    std::string all;
    for (int i = 0; i < n; i++) {
        pTreeNode tn = (pTreeNode)s->tree->super.nlist->array[i];
        left[i] = right[i] = parent[i] = -1;
        if (tn->children && tn->children->head) {
            left[i] = ((pTreeNode)tn->children->head->data)->id;
            right[i] = ((pTreeNode)tn->children->tail->data)->id;
        }
        if (tn->parent) parent[i] = tn->parent->id;
        bl[i] = ((pPhylogenyNode)tn)->branchlength;
        const char* nm = ((pPhylogenyNode)tn)->name;
        all += (nm ? nm : "");
        all += "\n";
    }
    strncpy(names, all.c_str(), names_len - 1);
    names[names_len - 1] = 0;
}

void refshim_set_range(void* h, int rmin, int rmax, int root_min, int root_max) {
    Session* s = (Session*)h;
    s->range = make_range(rmin, rmax, root_min, root_max);
    cafe_tree_set_parameters(s->tree, &s->range, 0);
}

void refshim_set_rates(void* h, const double* lambda, const double* mu) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    for (int i = 0; i < n; i++) {
        node_at(s, i)->birth_death_probabilities.lambda = lambda[i];
        node_at(s, i)->birth_death_probabilities.mu = mu[i];
    }
}
// reset_birthdeath_cache (cafe/cafe_main.c:319): rebuilds every (int t, lambda, mu) matrix
void refshim_reset_cache(void* h) {
    Session* s = (Session*)h;
    const int need = s->range.max > s->range.root_max ? s->range.max : s->range.root_max;
    ensure_lnc(need);  // never let the reference GROW its lnC cache (see ensure_lnc): a fresh process would start at this size
    reset_birthdeath_cache(s->tree, 0, &s->range);
}
int refshim_get_matrix(void* h, int node, double* out) {
    Session* s = (Session*)h;
    struct square_matrix* m = node_at(s, node)->birthdeath_matrix;
    if (!m) return 0;
    if (out) memcpy(out, m->values, sizeof(double) * (size_t)m->size * m->size);
    return m->size;
}

// dense errormatrix[observed][true] (dim x dim) attached to leaf `node` (nlist index); dim==0 clears
void refshim_set_errormodel(void* h, int node, const double* matrix, int dim, int fromdiff, int todiff) {
    Session* s = (Session*)h;
    if (dim == 0) { node_at(s, node)->errormodel = nullptr; return; }
    pErrorStruct e = (pErrorStruct)calloc(1, sizeof(ErrorStruct));
    e->maxfamilysize = dim - 1; e->fromdiff = fromdiff; e->todiff = todiff;
    e->errormatrix = (double**)calloc(dim, sizeof(double*));
    for (int i = 0; i < dim; i++) {
        e->errormatrix[i] = (double*)calloc(dim + 2048, sizeof(double));  // slack: the reference may read past dim
        memcpy(e->errormatrix[i], matrix + (size_t)i * dim, sizeof(double) * dim);
    }
    s->errs.push_back(e);
    node_at(s, node)->errormodel = e;
}

// error-model file -> dense matrix via the reference's reader + column-sum fix (error_model.cpp:145-204,
// cafe_shell.c:585-622).  Returns dim; out may be NULL to query.
int refshim_read_errormodel(const char* path, int range_max, double* out, int* fromdiff, int* todiff) {
    ErrorStruct e; memset(&e, 0, sizeof(e));
    e.maxfamilysize = range_max;
    std::ifstream ifs(path);
    if (!ifs) return -1;
    ifs >> e;
    __check_error_model_columnsums(&e);
    int dim = e.maxfamilysize + 1;
    if (out) for (int i = 0; i < dim; i++) memcpy(out + (size_t)i * dim, e.errormatrix[i], sizeof(double) * dim);
    if (fromdiff) *fromdiff = e.fromdiff;
    if (todiff) *todiff = e.todiff;
    return dim;
}

// compute_tree_likelihoods for one family; counts in leaf order (even nlist indices)
int refshim_likelihoods(void* h, const int* counts, double* L_out) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    for (int i = 0; i < n; i++) node_at(s, i)->familysize = -1;
    for (int i = 0, k = 0; i < n; i += 2, k++) node_at(s, i)->familysize = counts[k];
    compute_tree_likelihoods(s->tree);
    memcpy(L_out, get_likelihoods(s->tree), sizeof(double) * s->tree->rfsize);
    return s->tree->rfsize;
}

// cafe_tree_viterbi for one family (cafe/viterbi.cpp:494): sizes of all nodes in nlist order; returns max_i L_root[i]
double refshim_viterbi(void* h, const int* counts, int* sizes_out) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    for (int i = 0; i < n; i++) {
        pCafeNode nd = node_at(s, i);
        nd->familysize = -1;
        if (nd->viterbi) memset(nd->viterbi, 0, sizeof(int) * s->tree->size_of_factor);  // as freshly calloc'ed (cafe_tree.c)
    }
    for (int i = 0, k = 0; i < n; i += 2, k++) node_at(s, i)->familysize = counts[k];
    cafe_tree_viterbi(s->tree);
    for (int i = 0; i < n; i++) sizes_out[i] = node_at(s, i)->familysize;
    double* L = ((pCafeNode)s->tree->super.root)->likelihoods;
    double ml = L[0];
    for (int i = 1; i < s->tree->rfsize; i++) if (L[i] > ml) ml = L[i];
    return ml;
}

// families attached to the session; species order = leaf order
void refshim_set_families(void* h, int F, const int* counts, int dedup) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    int nl = (n + 1) / 2;
    std::vector<std::string> species;
    for (int i = 0; i < n; i += 2) species.push_back(((pPhylogenyNode)s->tree->super.nlist->array[i])->name);
    if (s->family) cafe_family_free(s->family);
    s->family = cafe_family_init(species);
    for (int f = 0; f < F; f++) {
        std::ostringstream id; id << "F" << f;
        std::vector<int> v(counts + (size_t)f * nl, counts + (size_t)(f + 1) * nl);
        cafe_family_add_item(s->family, gene_family(id.str(), "d", v));
    }
    cafe_family_set_species_index(s->family, s->tree);
    if (dedup) __cafe_famliy_check_the_pattern(s->family);
}
void refshim_get_refs(void* h, int* ref_out) {
    Session* s = (Session*)h;
    for (int i = 0; i < s->family->flist->size; i++) ref_out[i] = ((pCafeFamilyItem)s->family->flist->array[i])->ref;
}

// get_posterior (cafe/lambda.cpp:691); *threw = 1 when the zero-likelihood exception fired
double refshim_get_posterior(void* h, const double* prior1000, int* threw, char* msg, int msg_len) {
    Session* s = (Session*)h;
    std::vector<double> pr(prior1000, prior1000 + FAMILYSIZEMAX);
    *threw = 0;
    try { return get_posterior(s->family, s->tree, pr); }
    catch (std::runtime_error& e) {
        *threw = 1;
        if (msg) { strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
        return log(0);
    }
}
void refshim_get_maxlh(void* h, int* out) {
    Session* s = (Session*)h;
    for (int i = 0; i < s->family->flist->size; i++) out[i] = ((pCafeFamilyItem)s->family->flist->array[i])->maxlh;
}

void refshim_prior_poisson(int shift, double lambda, double* out1000) {
    std::vector<double> pr;
    cafe_set_prior_rfsize_poisson_lambda(pr, shift, &lambda);
    memcpy(out1000, pr.data(), sizeof(double) * FAMILYSIZEMAX);
}
// find_poisson_lambda (lambda.cpp:808) — consumes one rand()
double refshim_find_poisson_lambda(void* h, int* iters, double* score) {
    Session* s = (Session*)h;
    poisson_lambda pl = find_poisson_lambda(s->family);
    double v = pl.parameters[0];
    if (iters) *iters = pl.num_iterations;
    if (score) *score = pl.score;
    free(pl.parameters);
    return v;
}

// cafe_conditional_distribution (conditional_distribution.cpp:86); out [rfsize][n_samples]
int refshim_cond_dist(void* h, int nthreads, int n_samples, double* out) {
    Session* s = (Session*)h;
    matrix cd = cafe_conditional_distribution(s->tree, &s->range, nthreads, n_samples);
    for (size_t r = 0; r < cd.size(); r++) memcpy(out + r * n_samples, cd[r].data(), sizeof(double) * n_samples);
    return (int)cd.size();
}
// one row, unsorted copy too is not available from the reference; sorted only
void refshim_random_probabilities(void* h, int root_size, int trials, double* out_sorted) {
    Session* s = (Session*)h;
    std::vector<double> p = get_random_probabilities(s->tree, root_size, trials);
    memcpy(out_sorted, p.data(), sizeof(double) * trials);
}
int refshim_random_familysize(void* h, int root_size, int max_family_size, int* sizes_out) {
    Session* s = (Session*)h;
    int m = cafe_tree_random_familysize(s->tree, root_size, max_family_size);
    int n = refshim_n_nodes(h);
    for (int i = 0; i < n; i++) sizes_out[i] = node_at(s, i)->familysize;
    return m;
}

// family p-value exactly as viterbi_section does it (viterbi.cpp:88-97): forced per-family range,
// cafe_tree_p_values, max.  cd is [cd_rows][n_samples].
double refshim_family_pvalue(void* h, int idx, const double* cd, int cd_rows, int n_samples, double* pvals_out, int* rfsize_out) {
    Session* s = (Session*)h;
    std::vector<std::vector<double> > cdv(cd_rows);
    for (int r = 0; r < cd_rows; r++) cdv[r].assign(cd + (size_t)r * n_samples, cd + (size_t)(r + 1) * n_samples);
    cafe_family_set_size_with_family_forced(s->family, idx, s->tree);
    int rf = s->tree->rfsize;
    if (rfsize_out) *rfsize_out = rf;
    double best = 0;
    if (rf > 0) {
        std::vector<double> p1(rf);
        cafe_tree_p_values(s->tree, p1, cdv, n_samples);
        best = p1[0];
        for (int i = 0; i < rf; i++) { if (pvals_out) pvals_out[i] = p1[i]; if (p1[i] > best) best = p1[i]; }
    }
    copy_range_to_tree(s->tree, &s->range);
    return best;
}

// Viterbi as viterbi_section does it for family `idx` (viterbi.cpp:88-119): forced per-family range, cafe_tree_viterbi,
// viterbi_sum_probabilities.  sizes_out[n_nodes]; branch_pv_out[n_nodes] indexed by the CHILD node of each branch (root: -1).
void refshim_viterbi_forced(void* h, int idx, int* sizes_out, double* branch_pv_out) {
    Session* s = (Session*)h;
    int n = refshim_n_nodes(h);
    for (int i = 0; i < n; i++) {
        pCafeNode nd = node_at(s, i);
        if (nd->viterbi) memset(nd->viterbi, 0, sizeof(int) * s->tree->size_of_factor);
    }
    cafe_family_set_size_with_family_forced(s->family, idx, s->tree);
    cafe_tree_viterbi(s->tree);
    for (int i = 0; i < n; i++) { sizes_out[i] = node_at(s, i)->familysize; branch_pv_out[i] = -1; }
    viterbi_parameters vp;
    pCafeFamilyItem pitem = (pCafeFamilyItem)s->family->flist->array[idx];
    viterbi_sum_probabilities(&vp, s->tree, pitem);
    int nnodes = (n - 1) / 2;
    for (int j = 0; j < nnodes; j++) {
        pTreeNode pn = (pTreeNode)s->tree->super.nlist->array[2 * j + 1];
        pTreeNode child[2] = {(pTreeNode)pn->children->head->data, (pTreeNode)pn->children->tail->data};
        for (int k = 0; k < 2; k++) {
            int ci = -1;
            for (int i = 0; i < n; i++) if ((pTreeNode)s->tree->super.nlist->array[i] == child[k]) ci = i;
            branch_pv_out[ci] = vp.viterbiPvalues[viterbi_parameters::NodeFamilyKey(2 * j + k, pitem)];
        }
    }
    copy_range_to_tree(s->tree, &s->range);
}

// reference family-table reader + dedup (gene_family.cpp:186-225, cafe_family.c:9-34)
int refshim_load_families(const char* path, int max_size, int* n_species, int* F_out, int* counts_out, int cap, int* ref_out, int* max_size_out) {
    std::ifstream ifs(path);
    if (!ifs) return -1;
    std::string p(path);
    char sep = (p.size() >= 3 && p.substr(p.size() - 3) == "csv") ? ',' : '\t';
    pCafeFamily f = load_gene_families(ifs, sep, max_size);
    if (!f) return -2;
    *n_species = f->num_species; *F_out = f->flist->size;
    if (max_size_out) *max_size_out = f->max_size;
    if (counts_out) {
        if ((long)f->flist->size * f->num_species > cap) return -3;
        for (int i = 0; i < f->flist->size; i++) {
            pCafeFamilyItem it = (pCafeFamilyItem)f->flist->array[i];
            memcpy(counts_out + (size_t)i * f->num_species, it->count, sizeof(int) * f->num_species);
            if (ref_out) ref_out[i] = it->ref;
        }
    }
    cafe_family_free(f);
    return 0;
}

// the reference's Nelder-Mead (libcommon/fminsearch.cpp) on a caller-supplied function
int refshim_fminsearch(math_func f, void* args, int n, const double* x0, double tolx, double tolf, double* x_out, double* f_out, int* iters_out) {
    pFMinSearch pfm = fminsearch_new_with_eq(f, n, args);
    pfm->tolx = tolx; pfm->tolf = tolf;
    std::vector<double> start(x0, x0 + n);
    fminsearch_min(pfm, &start[0]);
    memcpy(x_out, fminsearch_get_minX(pfm), n * sizeof(double));
    *f_out = fminsearch_get_minF(pfm);
    *iters_out = pfm->iters;
    fminsearch_free(pfm);
    return 0;
}

double refshim_chi2cdf(double x, int df) { return chi2cdf(x, df); }

// cafe_likelihood_ratio_test (cafe/cafe_main.c:398-431) with num_threads = 1 on the session's tree, families and current
// matrices (refshim_reset_cache first); out is row-major [n_nodes][F].
// Reference defect, fenced: the test works on cafe_tree_copy(param->pcafe), and cafe_tree_node_copy (cafe_tree.c:485-494)
// copies lambda but not mu - the copy's nodes get the TREE-level pcafe->mu (cafe_tree.c:39), which cafe_tree_new leaves at 0,
// so the lengthened branches would be keyed (t, lambda, 0) whatever the model (and mu = 0 makes log(alpha) = -inf, NaN entries).
// The shim sets the tree-level mu to the nodes' common mu first, so the copy carries the rates the caller set; callers
// therefore use one mu for all nodes here.
// keep_node_mu = 0 runs the stock behaviour (tree-level mu as cafe_tree_new left it: 0).
void refshim_likelihood_ratio_test(void* h, const double* max_pvalues, double pvalue_cutoff, int keep_node_mu, double* out) {
    Session* s = (Session*)h;
    s->tree->mu = keep_node_mu ? node_at(s, 0)->birth_death_probabilities.mu : 0;
    CafeParam param;
    memset(&param, 0, sizeof(param));
    param.pcafe = s->tree;
    param.pfamily = s->family;
    param.num_threads = 1;
    param.pvalue = pvalue_cutoff;
    param.quiet = 1;
    param.flog = stdout;
    const int F = s->family->flist->size, n = refshim_n_nodes(h);
    std::vector<double> mp(max_pvalues, max_pvalues + F);
    cafe_likelihood_ratio_test(&param, mp.data());
    for (int b = 0; b < n; b++) {
        memcpy(out + (size_t)b * F, param.likelihoodRatios[b], sizeof(double) * F);
        free(param.likelihoodRatios[b]);
    }
    free(param.likelihoodRatios);
}

void refshim_session_free(void* h) {
    Session* s = (Session*)h;
    if (s->family) cafe_family_free(s->family);
    // tree and error structs intentionally leaked (test process lifetime)
    delete s;
}


// Branch cutting for ONE branch b (nlist index) and every family of the session: cut_branch (cafe/branch_cutting.cpp:185-219:
// copy + split of the tree, conditional distributions of the remaining tree and of the cut-off subtree from the rand() stream,
// n_samples draws, or n_samples / 10 each when neither side is a single leaf) followed by compute_cutpvalues (:101-150).
// The stock driver cafe_branch_cutting (:221-272) cannot run - it never fills pfamily / viterbi / num_random_samples / pvalue of
// its thread parameters - so the two functions are called the way its thread function would have called them.
// keep_node_mu: as in refshim_likelihood_ratio_test - cafe_tree_copy gives the copies' nodes the TREE-level mu (cafe_tree.c:39,
// :485-494), which cafe_tree_new leaves at 0; 1 sets it to the nodes' common mu first, 0 is the stock behaviour.
// out[F]: cutPvalues[b][f] (-1 where max_pvalues[f] > cutoff; 0 for a root b: compute_cutpvalues returns before writing).
// dims = {rows of the first distribution, its draws, rows of the second, its draws}; cd1 / cd2 (nullable) receive them row-major.
int refshim_branch_cut(void* h, int b, int n_samples, int keep_node_mu, const double* max_pvalues, double cutoff, double* out,
                       double* cd1, double* cd2, int* dims, char* log, int log_len) {
    Session* s = (Session*)h;
    const int n = refshim_n_nodes(h), F = s->family->flist->size;
    s->tree->mu = keep_node_mu ? node_at(s, 0)->birth_death_probabilities.mu : 0;
    CutBranch cb(n);
    std::ostringstream ost;
    cut_branch(cb, (pTree)s->tree, s->tree, s->range, 1, n_samples, b, ost);
    if (log) { strncpy(log, ost.str().c_str(), log_len - 1); log[log_len - 1] = 0; }
    matrix& m1 = cb.pCDSs[b].first; matrix& m2 = cb.pCDSs[b].second;
    dims[0] = (int)m1.size(); dims[1] = m1.empty() ? 0 : (int)m1[0].size();
    dims[2] = (int)m2.size(); dims[3] = m2.empty() ? 0 : (int)m2[0].size();
    if (cd1) for (size_t i = 0; i < m1.size(); i++) memcpy(cd1 + i * m1[i].size(), m1[i].data(), sizeof(double) * m1[i].size());
    if (cd2) for (size_t i = 0; i < m2.size(); i++) memcpy(cd2 + i * m2[i].size(), m2[i].data(), sizeof(double) * m2[i].size());
    viterbi_parameters v;
    viterbi_parameters_init(&v, n, F);
    std::vector<double> mp(max_pvalues, max_pvalues + F);
    v.maximumPvalues = mp.data();
    v.cutPvalues = (double**)memory_new_2dim(n, F, sizeof(double));
    std::vector<double> p1(s->tree->rfsize);
    double** p2 = (double**)memory_new_2dim(s->tree->rfsize, s->tree->rfsize, sizeof(double));
    compute_cutpvalues(s->tree, s->family, n_samples, b, 0, F, v, cutoff, p1, p2, cb);
    for (int f = 0; f < F; f++) out[f] = v.cutPvalues[b][f];
    memory_free_2dim((void**)p2, s->tree->rfsize, s->tree->rfsize, NULL);
    memory_free_2dim((void**)v.cutPvalues, n, F, NULL);
    v.maximumPvalues = nullptr; v.cutPvalues = nullptr;
    return F;
}
}  // extern "C"
