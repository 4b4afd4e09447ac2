// capi.cpp — extern "C" doors into the host library for the Python tests / bench (ctypes).
// No torch types, plain pointers and sizes.  Errors are returned as negative codes; the text of the
// last one is available from cafe_host_last_error().
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "cafe_commands.h"
#include "cafe_math.h"

static thread_local std::string g_host_err;

#define HOST_TRY try {
#define HOST_CATCH(rv)                                         \
    }                                                          \
    catch (std::exception & e) { g_host_err = e.what(); return rv; }

extern "C" {

const char* cafe_host_last_error() { return g_host_err.c_str(); }

// ---- math (parity anchors) ----
double cafe_host_gammaln(double a) { return cafe::gammaln(a); }
double cafe_host_chooseln(double n, double r) { return cafe::chooseln(n, r); }
double cafe_host_poisspdf(int x, double l) { return cafe::poisspdf(x, l); }
double cafe_host_chi2cdf(double x, int df) { return cafe::chi2cdf(x, df); }
double cafe_host_pvalue(double v, const double* cd, int n) { return cafe::pvalue(v, cd, n); }
void cafe_host_lnc_table(int size, double* out) {
    std::vector<double> T = cafe::lnc_table(size);
    std::memcpy(out, T.data(), T.size() * sizeof(double));
}
void cafe_host_init_family_size(int max, int* out4) {
    family_size_range r;
    init_family_size(&r, max);
    out4[0] = r.root_min; out4[1] = r.root_max; out4[2] = r.min; out4[3] = r.max;
}

// ---- Nelder–Mead on a caller-supplied function (parity with libcommon/fminsearch.cpp) ----
int cafe_host_fminsearch(math_func f, void* args, int n, const double* x0, double tolx, double tolf, double* x_out,
                         double* f_out, int* iters_out) {
    HOST_TRY
    pFMinSearch pfm = fminsearch_new_with_eq(f, n, args);
    pfm->tolx = tolx; pfm->tolf = tolf;
    std::vector<double> start(x0, x0 + n);
    fminsearch_min(pfm, start.data());
    std::memcpy(x_out, fminsearch_get_minX(pfm), n * sizeof(double));
    *f_out = fminsearch_get_minF(pfm);
    *iters_out = pfm->iters;
    fminsearch_free(pfm);
    return 0;
    HOST_CATCH(-1)
}

// ---- tree parsing without a session ----
// returns n_nodes (or -1); arrays must hold max_nodes entries; names are '\n'-joined
int cafe_host_parse_tree(const char* newick, int max_nodes, int* left, int* right, int* parent, double* bl, char* names, int names_len) {
    HOST_TRY
    family_size_range rg{0, 1, 0, 1};
    pCafeTree t = cafe_tree_new(newick, &rg, 0, 0);
    int n = t->num_nodes();
    if (n > max_nodes) { cafe_tree_free(t); g_host_err = "max_nodes too small"; return -1; }
    std::string all;
    for (int i = 0; i < n; ++i) {
        left[i] = t->nlist[i].left; right[i] = t->nlist[i].right; parent[i] = t->nlist[i].parent; bl[i] = t->nlist[i].branchlength;
        all += t->nlist[i].name; all += "\n";
    }
    std::strncpy(names, all.c_str(), names_len - 1);
    names[names_len - 1] = 0;
    cafe_tree_free(t);
    return n;
    HOST_CATCH(-1)
}

int cafe_host_parse_lambda_tree(const char* tree_newick, const char* lambda_newick, int* taxaid_out, int max_nodes) {
    HOST_TRY
    family_size_range rg{0, 1, 0, 1};
    pCafeTree t = cafe_tree_new(tree_newick, &rg, 0, 0);
    std::vector<int> ids;
    int m = parse_lambda_tree(lambda_newick, *t, ids);
    if ((int)ids.size() > max_nodes) { cafe_tree_free(t); return -1; }
    std::copy(ids.begin(), ids.end(), taxaid_out);
    cafe_tree_free(t);
    return m;
    HOST_CATCH(-1)
}

// error-model file -> dense matrix (reader + column-sum fix); returns dim or -1
int cafe_host_read_errormodel(const char* path, int range_max, double* out, int out_cap, int* fromdiff, int* todiff) {
    HOST_TRY
    ErrorStruct em;
    em.maxfamilysize = range_max;
    std::ifstream ifs(path);
    if (!ifs) { g_host_err = "cannot open file"; return -1; }
    ifs >> em;
    __check_error_model_columnsums(&em);
    int dim = em.maxfamilysize + 1;
    if (out) {
        if ((long)dim * dim > out_cap) { g_host_err = "out_cap too small"; return -1; }
        std::memcpy(out, em.errormatrix.data(), sizeof(double) * dim * dim);
    }
    if (fromdiff) *fromdiff = em.fromdiff;
    if (todiff) *todiff = em.todiff;
    return dim;
    HOST_CATCH(-1)
}

// family table reader; returns 0 or -1.  counts_out may be NULL to query sizes first.
int cafe_host_load_families(const char* path, int max_size, int* n_species, int* n_families, int* counts_out, long cap,
                            int* ref_out, int* max_size_out) {
    HOST_TRY
    std::ifstream ifs(path);
    if (!ifs) { g_host_err = "cannot open file"; return -1; }
    std::string p(path);
    char sep = (p.size() >= 3 && p.compare(p.size() - 3, 3, "csv") == 0) ? ',' : '\t';
    pCafeFamily f = load_gene_families(ifs, sep, max_size);
    *n_species = f->num_species;
    *n_families = (int)f->flist.size();
    if (max_size_out) *max_size_out = f->max_size;
    if (counts_out) {
        if ((long)f->flist.size() * f->num_species > cap) { cafe_family_free(f); g_host_err = "cap too small"; return -1; }
        for (size_t i = 0; i < f->flist.size(); ++i) {
            std::copy(f->flist[i].count.begin(), f->flist[i].count.end(), counts_out + i * f->num_species);
            if (ref_out) ref_out[i] = f->flist[i].ref;
        }
    }
    cafe_family_free(f);
    return 0;
    HOST_CATCH(-1)
}

// ---- sessions: the reference's Globals + command dispatcher ----
void* cafe_host_new(int quiet) {
    Globals* g = new Globals();
    g->param.quiet = quiet;
    return g;
}
void cafe_host_free(void* h) { delete static_cast<Globals*>(h); }
void cafe_host_release_gpu() { cafe_gpu_engine_release(); }

int cafe_host_command(void* h, const char* line) { return cafe_shell_dispatch_command(*static_cast<Globals*>(h), line); }

void cafe_host_srand(unsigned seed) { std::srand(seed); }
// find_poisson_lambda on the loaded table (consumes one rand()), cafe/lambda.cpp:808-838
int cafe_host_find_poisson_lambda(void* h, double* lambda, int* iters, double* score) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.pfamily) { g_host_err = "no family table loaded"; return -1; }
    poisson_lambda r = find_poisson_lambda(p.pfamily);
    *lambda = r.parameters[0]; *iters = r.num_iterations; *score = r.score;
    return 0;
    HOST_CATCH(-1)
}
int cafe_host_get_family_table(void* h, int* counts_out, long cap, int* ref_out, int* index_out) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.pfamily) return -1;
    const int ns = p.pfamily->num_species;
    if ((long)p.pfamily->flist.size() * ns > cap) return -1;
    for (size_t i = 0; i < p.pfamily->flist.size(); ++i) {
        std::copy(p.pfamily->flist[i].count.begin(), p.pfamily->flist[i].count.end(), counts_out + i * ns);
        if (ref_out) ref_out[i] = p.pfamily->flist[i].ref;
    }
    if (index_out) std::copy(p.pfamily->index.begin(), p.pfamily->index.end(), index_out);
    return ns;
}

int cafe_host_num_params(void* h) { return static_cast<Globals*>(h)->param.num_params; }
int cafe_host_get_parameters(void* h, double* out, int cap) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.input.parameters || cap < p.num_params) return -1;
    std::copy(p.input.parameters, p.input.parameters + p.num_params, out);
    return p.num_params;
}
int cafe_host_objective_calls(void* h) { return static_cast<Globals*>(h)->param.objective_calls; }
int cafe_host_get_ranges(void* h, int* out4) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    out4[0] = p.family_size.min; out4[1] = p.family_size.max; out4[2] = p.family_size.root_min; out4[3] = p.family_size.root_max;
    return 0;
}
int cafe_host_get_prior(void* h, double* out, int n) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.prior_rfsize || n > FAMILYSIZEMAX) return -1;
    std::copy(p.prior_rfsize, p.prior_rfsize + n, out);
    return n;
}
int cafe_host_num_families(void* h) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    return p.pfamily ? (int)p.pfamily->flist.size() : 0;
}
// one objective evaluation through the reference-named callback (seam B1); returns -score in *out
int cafe_host_objective(void* h, const double* x, int n, double* out) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (n != p.num_params) { g_host_err = "wrong number of parameters"; return -1; }
    std::vector<double> xx(x, x + n);
    *out = (p.optimizer_init_type == LAMBDA_MU) ? cafe_best_lambda_mu_search(xx.data(), &p) : __cafe_best_lambda_search(xx.data(), &p);
    return 0;
    HOST_CATCH(-1)
}
// per-family root likelihood vectors at the current rates: out [F][rfsize]
int cafe_host_family_likelihoods(void* h, double* out, long cap) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    std::vector<double> L = compute_tree_likelihoods_all(p.pfamily, p.pcafe);
    if ((long)L.size() > cap) { g_host_err = "cap too small"; return -1; }
    std::copy(L.begin(), L.end(), out);
    return p.pcafe->rfsize;
    HOST_CATCH(-1)
}
int cafe_host_get_cond_dist(void* h, double* out, long cap, int* rows, int* cols) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (p.cond_dist.empty()) return -1;
    *rows = (int)p.cond_dist.size(); *cols = (int)p.cond_dist[0].size();
    if ((long)*rows * *cols > cap) return -1;
    for (int r = 0; r < *rows; ++r) std::copy(p.cond_dist[r].begin(), p.cond_dist[r].end(), out + (size_t)r * *cols);
    return 0;
}
int cafe_host_get_max_pvalues(void* h, double* out, int cap) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    if ((int)p.max_pvalues.size() > cap) return -1;
    std::copy(p.max_pvalues.begin(), p.max_pvalues.end(), out);
    return (int)p.max_pvalues.size();
}
// The text report from given per-family results (no device work): sizes / branch_pv row-major [families][nodes], max_pv [families],
// lr nullable [nodes][families].  Lets the CPU tests pin the writer against the stock binary's report with the oracle's numbers.
int cafe_host_report_text_from(void* h, const double* lambda, int num_lambdas, const int* sizes, const double* branch_pv,
                               const double* max_pv, const double* lr, const char* path) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.pcafe || !p.pfamily) { g_host_err = "load and tree first"; return -1; }
    const size_t F = p.pfamily->flist.size(), n = p.pcafe->num_nodes();
    p.max_pvalues.assign(max_pv, max_pv + F);
    viterbi_parameters v;
    cafe_viterbi_from(&p, std::vector<int>(sizes, sizes + F * n), std::vector<double>(branch_pv, branch_pv + F * n), v);
    std::vector<double> lam(lambda, lambda + num_lambdas);
    double* const saved_lambda = p.lambda;
    const int saved_n = p.num_lambdas;
    const auto saved_lr = p.likelihoodRatios;
    p.lambda = lam.data();
    p.num_lambdas = num_lambdas;
    p.likelihoodRatios.clear();
    if (lr)
        for (size_t b = 0; b < n; ++b) p.likelihoodRatios.emplace_back(lr + b * F, lr + (b + 1) * F);
    std::ofstream out(path);
    if (out) cafe_report_text(out, &p, v);
    p.lambda = saved_lambda;
    p.num_lambdas = saved_n;
    p.likelihoodRatios = saved_lr;
    if (!out) { g_host_err = "cannot open report file"; return -1; }
    return 0;
    HOST_CATCH(-1)
}
int cafe_host_set_max_pvalues(void* h, const double* in, int n) {
    CafeParam& p = static_cast<Globals*>(h)->param;
    p.max_pvalues.assign(in, in + n);
    return n;
}
// cafe_likelihood_ratio_test with the family p-values of the last report (or all families when none were computed);
// out row-major [nodes][families]
int cafe_host_likelihood_ratio_test(void* h, int tree_level_mu, double* out, long cap, int* nodes, int* families) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    std::vector<double> mp = p.max_pvalues;
    if (mp.size() != p.pfamily->flist.size()) mp.assign(p.pfamily->flist.size(), 0.0);
    const int saved = p.lrt_tree_level_mu;
    p.lrt_tree_level_mu = tree_level_mu;
    try { cafe_likelihood_ratio_test(&p, mp.data()); } catch (...) { p.lrt_tree_level_mu = saved; throw; }
    p.lrt_tree_level_mu = saved;
    *nodes = (int)p.likelihoodRatios.size();
    *families = (int)p.pfamily->flist.size();
    if ((long)*nodes * *families > cap) { g_host_err = "cap too small"; return -1; }
    for (int b = 0; b < *nodes; ++b) std::copy(p.likelihoodRatios[b].begin(), p.likelihoodRatios[b].end(), out + (size_t)b * *families);
    return 0;
    HOST_CATCH(-1)
}

// cafe_branch_cutting with the family p-values set by cafe_host_set_max_pvalues / the last report; out row-major [nodes][families]
int cafe_host_branch_cutting(void* h, int num_random_samples, int tree_level_mu, double* out, long cap, int* nodes, int* families) {
    HOST_TRY
    CafeParam& p = static_cast<Globals*>(h)->param;
    if (!p.pcafe || !p.pfamily) { g_host_err = "load and tree first"; return -1; }
    if (p.max_pvalues.size() != p.pfamily->flist.size()) p.max_pvalues.assign(p.pfamily->flist.size(), 0.0);
    const int saved = p.lrt_tree_level_mu;
    p.lrt_tree_level_mu = tree_level_mu;
    try { cafe_branch_cutting(&p, num_random_samples); } catch (...) { p.lrt_tree_level_mu = saved; throw; }
    p.lrt_tree_level_mu = saved;
    *nodes = (int)p.cutPvalues.size();
    *families = (int)p.pfamily->flist.size();
    if ((long)*nodes * *families > cap) { g_host_err = "cap too small"; return -1; }
    for (int b = 0; b < *nodes; ++b) std::copy(p.cutPvalues[b].begin(), p.cutPvalues[b].end(), out + (size_t)b * *families);
    return 0;
    HOST_CATCH(-1)
}

}  // extern "C"
