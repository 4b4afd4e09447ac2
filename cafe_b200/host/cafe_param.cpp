#include "cafe_param.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>

#include "cafe_math.h"

void cafe_log(pCafeParam param, const char* msg, ...) {
    va_list ap;
    va_start(ap, msg);
    if (param->flog && param->flog != stderr && param->flog != stdout) {
        va_list ap2;
        va_copy(ap2, ap);
        vfprintf(param->flog, msg, ap2);
        va_end(ap2);
    }
    if (!param->quiet) {
        vfprintf(stdout, msg, ap);
        fflush(stdout);
    }
    if (param->flog) fflush(param->flog);
    va_end(ap);
}

// ================================================================================================
// GPU engine: one context behind the reference's global-state entry points
// ================================================================================================
namespace {

struct EngineState {
    cafe_gpu_ctx* ctx = nullptr;
    // what the device currently holds
    std::vector<int> left, right;
    std::vector<double> branchlength;
    family_size_range range{-1, -1, -1, -1};
    int lnc_size = 0;
    const CafeFamily* family = nullptr;
    size_t family_F = 0;
    std::vector<int> family_index;
    std::vector<int> unique_first;  // unique pattern u -> index of its first family in flist
    std::vector<int> family_unique; // family i -> unique pattern index
    std::string err_signature;
    std::vector<double> prior;
    bool matrices_valid = false;
};
EngineState g_eng;

void gpu_check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string("cafe_gpu ") + what + ": " + cafe_gpu_last_error(g_eng.ctx));
}

void sync_tree(pCafeTree t) {
    const int n = t->num_nodes();
    std::vector<int> l(n), r(n);
    std::vector<double> bl(n);
    for (int i = 0; i < n; ++i) { l[i] = t->nlist[i].left; r[i] = t->nlist[i].right; bl[i] = t->nlist[i].branchlength; }
    if (l != g_eng.left || r != g_eng.right || bl != g_eng.branchlength) {
        gpu_check(cafe_gpu_set_tree(g_eng.ctx, n, l.data(), r.data(), bl.data()), "set_tree");
        g_eng.left = l; g_eng.right = r; g_eng.branchlength = bl;
        g_eng.family = nullptr;       // leaf mapping may have changed
        g_eng.err_signature = "?";
        g_eng.matrices_valid = false;
    }
}

void sync_ranges(const family_size_range& rg) {
    if (std::memcmp(&rg, &g_eng.range, sizeof(rg)) == 0) return;
    gpu_check(cafe_gpu_set_ranges(g_eng.ctx, rg.min, rg.max, rg.root_min, rg.root_max), "set_ranges");
    g_eng.range = rg;
    g_eng.prior.clear();
    g_eng.err_signature = "?";
    g_eng.matrices_valid = false;
    const int maxfs = std::max(rg.max, rg.root_max);
    if (g_eng.lnc_size < maxfs) {  // birthdeath_cache_init, libtree/birthdeath.c:331-343
        std::vector<double> T = cafe::lnc_table(maxfs);
        gpu_check(cafe_gpu_set_lnc_table(g_eng.ctx, T.data(), 2 * maxfs, maxfs + 1), "set_lnc_table");
        g_eng.lnc_size = maxfs;
    }
}

void sync_family(pCafeFamily f, pCafeTree t) {
    if (g_eng.family == f && g_eng.family_F == f->flist.size() && g_eng.family_index == f->index) return;
    const int nl = t->num_leaves();
    std::vector<int> leaf_species(nl, -1);  // leaf k (node 2k) <- species column
    for (int i = 0; i < f->num_species; ++i) {
        int idx = f->index[i];
        if (idx >= 0 && idx < t->num_nodes() && (idx & 1) == 0) leaf_species[idx / 2] = i;
    }
    for (int k = 0; k < nl; ++k)
        if (leaf_species[k] < 0) throw std::runtime_error("Warning: Tree and family indices not synchronized");
    const size_t F = f->flist.size();
    g_eng.unique_first.clear();
    g_eng.family_unique.assign(F, -1);
    std::vector<int> mult;
    for (size_t i = 0; i < F; ++i) {
        int ref = f->flist[i].ref;
        if (ref < 0 || ref == (int)i) {
            g_eng.family_unique[i] = (int)g_eng.unique_first.size();
            g_eng.unique_first.push_back((int)i);
            mult.push_back(1);
        } else {
            g_eng.family_unique[i] = g_eng.family_unique[ref];
            mult[g_eng.family_unique[ref]]++;
        }
    }
    const size_t U = g_eng.unique_first.size();
    std::vector<int32_t> counts(U * nl);
    for (size_t u = 0; u < U; ++u) {
        const CafeFamilyItem& it = f->flist[g_eng.unique_first[u]];
        for (int k = 0; k < nl; ++k) counts[u * nl + k] = it.count[leaf_species[k]];
    }
    gpu_check(cafe_gpu_set_families(g_eng.ctx, (int)U, nl, counts.data(), mult.data(), g_eng.unique_first.data()), "set_families");
    g_eng.family = f;
    g_eng.family_F = F;
    g_eng.family_index = f->index;
}

void sync_error_models(pCafeFamily f, pCafeTree t) {
    std::ostringstream sig;
    for (int k = 0; k < t->num_leaves(); ++k) {
        int e = t->nlist[2 * k].errormodel;
        sig << e;
        if (e >= 0 && f) {
            // file name, size and an FNV-1a hash of the matrix bits: a model edited in place is uploaded again
            const ErrorStruct& em = f->errors[e];
            uint64_t h = 1469598103934665603ULL;
            const unsigned char* b = reinterpret_cast<const unsigned char*>(em.errormatrix.data());
            for (size_t i = 0; i < em.errormatrix.size() * sizeof(double); ++i) { h ^= b[i]; h *= 1099511628211ULL; }
            sig << ":" << em.errorfilename << ":" << em.maxfamilysize << ":" << h;
        }
        sig << ";";
    }
    if (sig.str() == g_eng.err_signature) return;
    gpu_check(cafe_gpu_set_error_model(g_eng.ctx, -1, nullptr, 0), "set_error_model(clear)");
    for (int k = 0; k < t->num_leaves(); ++k) {
        int e = t->nlist[2 * k].errormodel;
        if (e < 0 || !f) continue;
        const ErrorStruct& em = f->errors[e];
        gpu_check(cafe_gpu_set_error_model(g_eng.ctx, k, em.errormatrix.data(), em.maxfamilysize + 1), "set_error_model");
    }
    g_eng.err_signature = sig.str();
}

// The reference runs the conditional distribution, the family p-values / Viterbi pass and the likelihood-ratio test on
// cafe_tree_copy(...) of the tree (conditional_distribution.cpp:77, viterbi.cpp:125, cafe_main.c:347), and cafe_tree_node_copy
// (cafe_tree.c:485-494) copies lambda, family size and matrix pointer only: those passes see NO error model, whatever the
// `errormodel` command attached (the stock binary's report is the same file with and without it).  So do ours.
void sync_without_error_models() {
    if (g_eng.err_signature == "-") return;
    gpu_check(cafe_gpu_set_error_model(g_eng.ctx, -1, nullptr, 0), "set_error_model(clear)");
    g_eng.err_signature = "-";
}

void sync_prior(const double* prior, int R) {
    if ((int)g_eng.prior.size() == R && std::equal(prior, prior + R, g_eng.prior.begin())) return;
    gpu_check(cafe_gpu_set_prior(g_eng.ctx, prior, R), "set_prior");
    g_eng.prior.assign(prior, prior + R);
}

void push_rates(pCafeTree t) {
    const int n = t->num_nodes();
    std::vector<double> lam(n), mu(n);
    for (int i = 0; i < n; ++i) {
        lam[i] = t->nlist[i].birth_death_probabilities.lambda;
        mu[i] = t->nlist[i].birth_death_probabilities.mu;
    }
    gpu_check(cafe_gpu_set_rates(g_eng.ctx, lam.data(), mu.data()), "set_rates");
}

}  // namespace

// Devices of the engine: the environment variable CAFE_GPUS is a count ("8" = devices 0..7) or a comma-separated list of
// CUDA device indices ("0,2,5"); unset = the current device.  More than one device makes the engine a multi-device context
// (cafe_gpu_create_multi): the family table is split over the devices and every objective evaluation of the lambda search
// runs K1 / all-gather / K2 / score reduction on all of them (include/cafe_gpu.h, "Multi-GPU").
static std::vector<int> engine_devices() {
    std::vector<int> dev;
    const char* e = std::getenv("CAFE_GPUS");
    if (!e || !*e) return dev;
    const std::string s(e);
    if (s.find(',') == std::string::npos) {
        const int n = std::atoi(s.c_str());
        for (int i = 0; i < n; ++i) dev.push_back(i);
        return dev;
    }
    size_t pos = 0;
    while (pos < s.size()) {
        size_t c = s.find(',', pos);
        if (c == std::string::npos) c = s.size();
        if (c > pos) dev.push_back(std::atoi(s.substr(pos, c - pos).c_str()));
        pos = c + 1;
    }
    return dev;
}

std::vector<int> cafe_gpu_engine_devices() {
    std::vector<int> dev = engine_devices();
    if (dev.empty()) dev.push_back(-1);  // the current device
    return dev;
}

cafe_gpu_ctx* cafe_gpu_engine() {
    if (!g_eng.ctx) {
        const std::vector<int> dev = engine_devices();
        int rc = dev.size() > 1 ? cafe_gpu_create_multi(&g_eng.ctx, dev.data(), (int)dev.size())
                                : cafe_gpu_create(&g_eng.ctx, dev.empty() ? -1 : dev[0]);
        if (rc != 0) throw std::runtime_error(std::string("cafe_gpu_create failed: ") + cafe_gpu_last_error(nullptr));
    }
    return g_eng.ctx;
}

void cafe_gpu_engine_release() {
    if (g_eng.ctx) cafe_gpu_destroy(g_eng.ctx);
    g_eng = EngineState();
}

void cafe_gpu_sync_state(pCafeFamily pfamily, pCafeTree pcafe, const double* prior_rfsize, bool with_error_models) {
    cafe_gpu_engine();
    sync_tree(pcafe);
    sync_ranges(pcafe->range);
    if (pfamily) sync_family(pfamily, pcafe);
    if (with_error_models) sync_error_models(pfamily, pcafe);
    else sync_without_error_models();
    if (prior_rfsize) sync_prior(prior_rfsize, pcafe->rfsize);
}

void reset_birthdeath_cache(pCafeTree tree, int k_value, family_size_range* range) {
    if (k_value > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    (void)range;  // matrix entries do not depend on the matrix size; the tree's own range fixes S
    cafe_gpu_engine();
    sync_tree(tree);
    sync_ranges(tree->range);
    push_rates(tree);
    gpu_check(cafe_gpu_build_matrices(g_eng.ctx), "build_matrices");
    g_eng.matrices_valid = true;
}

void cafe_free_birthdeath_cache(pCafeTree) { g_eng.matrices_valid = false; }

double get_posterior(pCafeFamily pfamily, pCafeTree pcafe, std::vector<double>& prior_rfsize) {
    cafe_gpu_sync_state(pfamily, pcafe, prior_rfsize.data());
    if (!g_eng.matrices_valid) {  // compute_child_factor sets missing matrices lazily, cafe_tree.c:219-220
        family_size_range rg = pcafe->range;
        reset_birthdeath_cache(pcafe, 0, &rg);
    }
    double score = 0;
    int32_t first_zero = -1;
    int rc = cafe_gpu_score(g_eng.ctx, &score, &first_zero);
    gpu_check(rc, "score");
    // maxlh side effect of compute_posterior (lambda.cpp:673-676): only families still at -1
    bool pending = false;
    for (const auto& it : pfamily->flist)
        if (it.maxlh < 0) { pending = true; break; }
    if (pending) {
        std::vector<int32_t> argmax(g_eng.unique_first.size());
        gpu_check(cafe_gpu_family_results(g_eng.ctx, nullptr, nullptr, argmax.data()), "family_results");
        const size_t upto = (rc == CAFE_GPU_ZERO_LIKELIHOOD) ? (size_t)first_zero + 1 : pfamily->flist.size();
        for (size_t i = 0; i < upto; ++i) {
            CafeFamilyItem& it = pfamily->flist[i];
            if ((it.ref < 0 || it.ref == (int)i) && it.maxlh < 0) it.maxlh = argmax[g_eng.family_unique[i]];
        }
    }
    if (rc == CAFE_GPU_ZERO_LIKELIHOOD) {
        std::ostringstream ost;
        ost << "WARNING: Calculated posterior probability for family " << pfamily->flist[first_zero].id << " = 0" << std::endl;
        throw std::runtime_error(ost.str());
    }
    return score;
}

std::vector<double> compute_tree_likelihoods_all(pCafeFamily pfamily, pCafeTree pcafe) {
    std::vector<double> unit_prior(pcafe->rfsize, 1.0);
    cafe_gpu_sync_state(pfamily, pcafe, g_eng.prior.size() == (size_t)pcafe->rfsize ? g_eng.prior.data() : unit_prior.data());
    if (!g_eng.matrices_valid) {
        family_size_range rg = pcafe->range;
        reset_birthdeath_cache(pcafe, 0, &rg);
    }
    const size_t U = g_eng.unique_first.size(), R = pcafe->rfsize;
    std::vector<double> LU(U * R);
    gpu_check(cafe_gpu_family_likelihoods(g_eng.ctx, LU.data()), "family_likelihoods");
    std::vector<double> L(pfamily->flist.size() * R);
    for (size_t i = 0; i < pfamily->flist.size(); ++i)
        std::copy(LU.begin() + g_eng.family_unique[i] * R, LU.begin() + (g_eng.family_unique[i] + 1) * R, L.begin() + i * R);
    return L;
}

// ================================================================================================
// parameter -> node mapping (cafe/cafe_shell.c:31-287)
// ================================================================================================
void cafe_shell_set_lambdas(pCafeParam param, double* parameters) {
    if (param->optimizer_init_type == LAMBDA_ONLY) cafe_shell_set_lambda(param, parameters);
    if (param->optimizer_init_type == LAMBDA_MU) cafe_shell_set_lambda_mu(param, parameters);
}

void cafe_shell_set_lambda(pCafeParam param, double* parameters) {
    if (param->input.parameters[0] != parameters[0])
        std::memcpy(param->input.parameters, parameters, param->num_params * sizeof(double));
    param->lambda = param->input.parameters;
    if (param->parameterized_k_value > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    param->pcafe->k = 0;
    // initialize_k_bd -> set_birth_death_probabilities4: lambda = parameters[taxaid], mu = -1
    for (int i = 0; i < param->pcafe->num_nodes(); ++i) {
        int taxa_id = 0;
        if (!param->lambda_tree.empty()) taxa_id = param->lambda_tree[i];
        if (taxa_id < 0) taxa_id = 0;
        param->pcafe->nlist[i].birth_death_probabilities.lambda = parameters[taxa_id];
        param->pcafe->nlist[i].birth_death_probabilities.mu = -1;
    }
}

void cafe_shell_set_lambda_mu(pCafeParam param, double* parameters) {
    if (param->input.parameters[0] != parameters[0])
        std::memcpy(param->input.parameters, parameters, param->num_params * sizeof(double));
    param->lambda = param->input.parameters;
    if (param->parameterized_k_value > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    param->mu = &param->input.parameters[param->num_lambdas];
    param->pcafe->k = 0;
    const int first_mu = param->num_lambdas;
    // initialize_k_bd2 -> set_birth_death_probabilities / set_birth_death_probabilities2
    for (int i = 0; i < param->pcafe->num_nodes(); ++i) {
        probabilities& p = param->pcafe->nlist[i].birth_death_probabilities;
        if (!param->lambda_tree.empty()) {
            int taxa_id = param->lambda_tree[i];
            if (taxa_id < 0) taxa_id = 0;
            p.lambda = parameters[taxa_id];
            if (param->eqbg) p.mu = (taxa_id == 0) ? p.lambda : parameters[first_mu + (taxa_id - param->eqbg)];
            else p.mu = parameters[first_mu + taxa_id];
        } else {
            p.lambda = parameters[0];
            p.mu = parameters[first_mu];
        }
    }
}

int __cafe_cmd_lambda_tree(pCafeParam param, const char* arg1, const char* arg2) {
    int idx = 1;
    const char* plambdastr = arg1;
    if (arg2 != nullptr) { std::sscanf(arg1, "%d", &idx); plambdastr = arg2; }
    std::vector<int> taxaid;
    int m = parse_lambda_tree(plambdastr, *param->pcafe, taxaid);
    if (idx == 2) return 1;  // the second lambda tree is only used by lhtest (outside the path)
    param->lambda_tree = taxaid;
    param->lambda_tree_string = plambdastr;
    param->num_lambdas = m;
    if (!param->quiet) std::printf("The number of lambdas is %d\n", m);
    return 0;
}

// ================================================================================================
// root prior (cafe/lambda.cpp:771-870)
// ================================================================================================
std::vector<int> collect_leaf_sizes(pCafeFamily pfamily) {
    std::vector<int> sizes;
    for (const auto& it : pfamily->flist)
        for (int i = 0; i < pfamily->num_species; ++i) {
            if (pfamily->index[i] < 0) continue;
            if (it.count[i] > 0) sizes.push_back(it.count[i] - 1);  // root size is conditioned to be >= 1
        }
    return sizes;
}

static double lnL_poisson(double* plambda, void* data) {  // __lnLPoisson
    const std::vector<int>& sizes = *static_cast<std::vector<int>*>(data);
    double score = 0;
    for (int x : sizes) {
        double ll = cafe::poisspdf(x, plambda[0]);
        if (std::isnan(ll)) ll = 0;
        score += std::log(ll);
    }
    return -score;
}

poisson_lambda find_poisson_lambda(pCafeFamily pfamily) {
    std::vector<int> sizes = collect_leaf_sizes(pfamily);
    pFMinSearch pfm = fminsearch_new_with_eq(lnL_poisson, 1, &sizes);
    pfm->tolx = 1e-6;
    pfm->tolf = 1e-6;
    double start[1] = {cafe::unifrnd()};  // consumes one rand()
    fminsearch_min(pfm, start);
    poisson_lambda r;
    r.parameters.assign(1, fminsearch_get_minX(pfm)[0]);
    r.num_params = 1;
    r.num_iterations = pfm->iters;
    r.score = *pfm->fv;
    fminsearch_free(pfm);
    return r;
}

void cafe_set_prior_rfsize_poisson_lambda(std::vector<double>& prior_rfsize, int shift, double* lambda) {
    prior_rfsize.resize(FAMILYSIZEMAX);
    for (int i = 0; i < FAMILYSIZEMAX; ++i) prior_rfsize[i] = cafe::poisspdf(shift - 1 + i, lambda[0]);  // shifted Poisson
}

double cafe_set_prior_rfsize_empirical(pCafeParam param, std::vector<double>& prior_rfsize) {
    poisson_lambda result = find_poisson_lambda(param->pfamily);
    cafe_log(param, "Empirical Prior Estimation Result: (%d iterations)\n", result.num_iterations);
    cafe_log(param, "Poisson lambda: %f & Score: %f\n", result.parameters[0], result.score);
    cafe_set_prior_rfsize_poisson_lambda(prior_rfsize, param->pcafe->range.root_min, result.parameters.data());
    return 0;
}

void input_values_randomize(input_values* vals, int lambda_len, int mu_len, int k, int kfix, double max_branch_length,
                            double* k_weights) {
    (void)kfix; (void)k_weights;
    if (mu_len < 0) mu_len = 0;
    if (k > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    for (int i = 0; i < lambda_len; ++i) vals->parameters[i] = 1.0 / max_branch_length * cafe::unifrnd();
    for (int i = 0; i < mu_len; ++i) vals->parameters[lambda_len + i] = 1.0 / max_branch_length * cafe::unifrnd();
}

// ================================================================================================
// objective callbacks and search drivers — seam B1
// ================================================================================================
static std::string join_doubles(int n, const double* v) {  // string_pchar_join_double: "%15.14lf" joined by ","
    std::string out;
    char buf[64];
    for (int i = 0; i < n; ++i) {
        std::snprintf(buf, sizeof buf, "%15.14lf", v[i]);
        out += buf;
        if (i < n - 1) out += ",";
    }
    return out;
}

static double objective_body(pCafeParam param, double* x, int n_checked) {
    // shared by the lambda and lambda/mu callbacks: negative guard, set rates, K1, K2+K3
    for (int i = 0; i < n_checked; ++i)
        if (x[i] < 0) return -std::numeric_limits<double>::infinity();  // log(0)
    cafe_shell_set_lambdas(param, x);
    reset_birthdeath_cache(param->pcafe, param->parameterized_k_value, &param->family_size);
    double score;
    try {
        std::vector<double> pr(param->prior_rfsize, param->prior_rfsize + FAMILYSIZEMAX);
        score = get_posterior(param->pfamily, param->pcafe, pr);
    } catch (std::runtime_error& e) {
        if (!param->quiet || param->optimizer_init_type == LAMBDA_MU) std::cerr << e.what();
        score = -std::numeric_limits<double>::infinity();
    }
    cafe_free_birthdeath_cache(param->pcafe);
    param->objective_calls++;
    return score;
}

double __cafe_best_lambda_search(double* plambda, void* args) {
    pCafeParam param = static_cast<pCafeParam>(args);
    double score = objective_body(param, plambda, param->num_lambdas);
    cafe_log(param, "Lambda : %s & Score: %f\n", join_doubles(param->num_lambdas, plambda).c_str(), score);
    cafe_log(param, ".");
    return -score;
}

double cafe_best_lambda_mu_search(double* parameters, void* args) {
    pCafeParam param = static_cast<pCafeParam>(args);
    double score = objective_body(param, parameters, param->num_params);
    cafe_log(param, "Lambda : %s ", join_doubles(param->num_lambdas, parameters).c_str());
    cafe_log(param, "Mu : %s & Score: %f\n",
             join_doubles(param->num_mus - param->eqbg, parameters + param->num_lambdas).c_str(), score);
    cafe_log(param, ".");
    return -score;
}

double* cafe_best_lambda_by_fminsearch(pCafeParam param, int lambda_len, int k) {
    if (k > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    const int max_runs = 10;
    std::vector<double> scores;
    bool converged = false;
    int runs = 0;
    do {
        if (param->num_params > 0)
            input_values_randomize(&param->input, param->num_lambdas, param->num_mus, param->parameterized_k_value, 0,
                                   max_branch_length(param->pcafe), nullptr);
        copy_range_to_tree(param->pcafe, &param->family_size);
        pFMinSearch pfm = fminsearch_new_with_eq(__cafe_best_lambda_search, lambda_len, param);
        pfm->tolx = 1e-6;
        pfm->tolf = 1e-6;
        // a copy of the start point goes to the search (lambda.cpp:560-563)
        std::vector<double> start(param->input.parameters, param->input.parameters + param->num_params);
        fminsearch_min(pfm, start.data());
        double* re = fminsearch_get_minX(pfm);
        for (int i = 0; i < param->num_params; ++i) param->input.parameters[i] = re[i];
        cafe_log(param, "\n");
        cafe_log(param, "Lambda Search Result: %d\n", pfm->iters);
        cafe_log(param, "Lambda : %s & Score: %f\n", join_doubles(param->num_lambdas, param->input.parameters).c_str(), *pfm->fv);
        if (runs > 0) {
            double minscore = *std::min_element(scores.begin(), scores.end());
            if (std::fabs(minscore - *pfm->fv) < 10 * pfm->tolf) converged = true;
        }
        scores.push_back(*pfm->fv);
        fminsearch_free(pfm);
        copy_range_to_tree(param->pcafe, &param->family_size);
        ++runs;
    } while (param->checkconv && !converged && runs < max_runs);
    if (param->checkconv) {
        if (converged) cafe_log(param, "score converged in %d runs.\n", runs);
        else cafe_log(param, "score failed to converge in %d runs.\n", max_runs);
    }
    return param->input.parameters;
}

void best_lambda_mu_by_fminsearch(pCafeParam param, int lambda_len, int mu_len, int k, std::ostream& log) {
    (void)lambda_len; (void)mu_len;
    if (k > 0) throw std::runtime_error("the clustered (-k) model is outside the GPU hot path");
    const int max_runs = 10;
    std::vector<double> scores(max_runs, 0.0);  // zero-filled like the reference's vector<double>(max_runs)
    bool converged = false;
    int runs = 0;
    do {
        if (param->num_params > 0)
            input_values_randomize(&param->input, param->num_lambdas, param->num_mus, param->parameterized_k_value, 0,
                                   max_branch_length(param->pcafe), nullptr);
        copy_range_to_tree(param->pcafe, &param->family_size);
        pFMinSearch pfm = fminsearch_new_with_eq(cafe_best_lambda_mu_search, param->num_params, param);
        pfm->tolx = 1e-6;
        pfm->tolf = 1e-6;
        // NB: the live parameter array is the start point (lambdamu.cpp:400): the objective copies each
        // trial into it (cafe_shell_set_lambda_mu), which shifts the later initial vertices. Kept.
        fminsearch_min(pfm, param->input.parameters);
        double* re = fminsearch_get_minX(pfm);
        for (int i = 0; i < param->num_params; ++i) param->input.parameters[i] = re[i];
        log << "\n";
        log << "Lambda Search Result: " << pfm->iters << "\n";
        log << "Lambda : " << join_doubles(param->num_lambdas, param->input.parameters) << " & Score: " << *pfm->fv;
        log << "Mu : " << join_doubles(param->num_mus - param->eqbg, param->input.parameters + param->num_lambdas)
            << " & Score: " << *pfm->fv << "\n";
        if (runs > 0) {
            double minscore = *std::min_element(scores.begin(), scores.end());
            if (std::fabs(minscore - *pfm->fv) < 10 * pfm->tolf) converged = true;
        }
        scores[runs] = *pfm->fv;
        fminsearch_free(pfm);
        copy_range_to_tree(param->pcafe, &param->family_size);
        ++runs;
    } while (param->checkconv && !converged && runs < max_runs);
    if (param->checkconv) {
        if (converged) log << "score converged in " << runs << " runs.\n";
        else log << "score failed to converge in " << max_runs << " runs.\n";
    }
}

// ================================================================================================
// conditional distribution and family p-values — seams B4/B5
// ================================================================================================
matrix cafe_conditional_distribution(pCafeTree pTree, family_size_range* range, int numthreads, int num_random_samples) {
    cafe_gpu_engine();
    sync_tree(pTree);
    family_size_range rg = pTree->range;
    rg.root_min = range->root_min; rg.root_max = range->root_max;
    sync_ranges(rg);
    sync_without_error_models();  // the thread's tree copy carries none (conditional_distribution.cpp:77)
    if (!g_eng.matrices_valid) { push_rates(pTree); gpu_check(cafe_gpu_build_matrices(g_eng.ctx), "build_matrices"); g_eng.matrices_valid = true; }
    const int R = rg.root_max - rg.root_min + 1;
    std::vector<double> flat((size_t)R * num_random_samples);
    const char* mode = std::getenv("CAFE_GPU_CD_RNG");
    bool replay = mode ? std::strcmp(mode, "replay") == 0 : numthreads <= 1;
    if (replay) {
        // the draws the single-threaded reference would make, in its order (cafe_tree.c:533-569)
        const size_t n = (size_t)R * num_random_samples * (pTree->num_nodes() - 1);
        std::vector<double> u(n);
        for (size_t i = 0; i < n; ++i) u[i] = cafe::unifrnd();
        gpu_check(cafe_gpu_conditional_distribution(g_eng.ctx, num_random_samples, u.data(), 0, flat.data()), "conditional_distribution");
    } else {
        uint64_t seed = ((uint64_t)std::rand() << 32) ^ (uint64_t)std::rand();
        gpu_check(cafe_gpu_conditional_distribution(g_eng.ctx, num_random_samples, nullptr, seed, flat.data()), "conditional_distribution");
    }
    matrix cd(R);
    for (int r = 0; r < R; ++r) cd[r].assign(flat.begin() + (size_t)r * num_random_samples, flat.begin() + (size_t)(r + 1) * num_random_samples);
    return cd;
}

void cafe_family_pvalues(pCafeParam param, std::vector<double>& max_pvalues) {
    if (param->cond_dist.empty()) throw std::runtime_error("conditional distribution not computed");
    cafe_gpu_sync_state(param->pfamily, param->pcafe, param->prior_rfsize, false);  // viterbi.cpp:125: a copy without error models
    if (!g_eng.matrices_valid) {
        reset_birthdeath_cache(param->pcafe, 0, &param->family_size);
    }
    const int rows = (int)param->cond_dist.size(), n = (int)param->cond_dist[0].size();
    std::vector<double> flat((size_t)rows * n);
    for (int r = 0; r < rows; ++r) std::copy(param->cond_dist[r].begin(), param->cond_dist[r].end(), flat.begin() + (size_t)r * n);
    std::vector<double> pu(g_eng.unique_first.size());
    gpu_check(cafe_gpu_pvalues(g_eng.ctx, flat.data(), rows, n, pu.data()), "pvalues");
    max_pvalues.resize(param->pfamily->flist.size());
    for (size_t i = 0; i < max_pvalues.size(); ++i) max_pvalues[i] = pu[g_eng.family_unique[i]];
}

void cafe_viterbi_all(pCafeParam param, std::vector<int>& node_sizes, std::vector<double>& branch_pvalues) {
    std::vector<double> unit_prior(param->pcafe->rfsize, 1.0);
    cafe_gpu_sync_state(param->pfamily, param->pcafe, param->prior_rfsize ? param->prior_rfsize : unit_prior.data(), false);  // viterbi.cpp:125
    if (!g_eng.matrices_valid) reset_birthdeath_cache(param->pcafe, 0, &param->family_size);
    const int nnodes = param->pcafe->num_nodes();
    const size_t nrows = param->pfamily->flist.size(), U = g_eng.unique_first.size();
    std::vector<int32_t> su(U * nnodes);
    std::vector<double> pu(U * nnodes);
    gpu_check(cafe_gpu_viterbi_report(g_eng.ctx, su.data(), pu.data()), "viterbi_report");
    node_sizes.resize(nrows * nnodes);
    branch_pvalues.resize(nrows * nnodes);
    for (size_t i = 0; i < nrows; ++i) {
        const size_t u = g_eng.family_unique[i];
        std::copy(su.begin() + u * nnodes, su.begin() + (u + 1) * nnodes, node_sizes.begin() + i * nnodes);
        std::copy(pu.begin() + u * nnodes, pu.begin() + (u + 1) * nnodes, branch_pvalues.begin() + i * nnodes);
    }
}

void cafe_likelihood_ratio_test(pCafeParam param, double* maximumPvalues) {
    cafe_log(param, "Running Likelihood Ratio Test....\n");
    std::vector<double> unit_prior(param->pcafe->rfsize, 1.0);
    cafe_gpu_sync_state(param->pfamily, param->pcafe, param->prior_rfsize ? param->prior_rfsize : unit_prior.data(), false);  // cafe_main.c:347
    if (!g_eng.matrices_valid) reset_birthdeath_cache(param->pcafe, 0, &param->family_size);
    const int nnodes = param->pcafe->num_nodes();
    const size_t nrows = param->pfamily->flist.size(), U = g_eng.unique_first.size();
    std::vector<uint8_t> tested(U);
    for (size_t u = 0; u < U; ++u) tested[u] = !(maximumPvalues[g_eng.unique_first[u]] > param->pvalue);  // cafe_main.c:358
    std::vector<double> base(U), best((size_t)nnodes * U);
    const std::vector<double> tree_mu(nnodes, 0.0);  // pcafe->mu as cafe_tree_new(..., 0, 0) leaves it (cafe_commands.cpp:1171)
    gpu_check(cafe_gpu_likelihood_ratio_test(g_eng.ctx, tested.data(), param->lrt_tree_level_mu ? tree_mu.data() : nullptr, base.data(),
                                             best.data(), nullptr), "likelihood_ratio_test");
    param->likelihoodRatios.assign(nnodes, std::vector<double>(nrows, -1.0));
    const int root = param->pcafe->root;
    for (int b = 0; b < nnodes; ++b) {
        if (b == root) continue;  // cafe_main.c:369-373
        std::vector<double> ratio(U, -1.0);
        for (size_t u = 0; u < U; ++u) {
            if (!tested[u]) continue;
            const double prevlh = best[(size_t)b * U + u], maxlh = base[u];
            ratio[u] = (prevlh == maxlh) ? 1 : 1 - cafe::chi2cdf(2 * (std::log(prevlh) - std::log(maxlh)), 1);  // :388
        }
        for (size_t i = 0; i < nrows; ++i) param->likelihoodRatios[b][i] = ratio[g_eng.family_unique[i]];  // :419-427
    }
    cafe_log(param, "Done : Likelihood Ratio test\n");
}

void write_pvalues(std::ostream& ost, const matrix& cd, int count) {
    ost << std::setw(10) << std::setprecision(9);
    for (const auto& row : cd) {
        ost << row[0];
        for (int j = 1; j < count; ++j) ost << "\t" << row[j];
        ost << "\n";
    }
}

matrix read_pvalues(std::istream& ist, int count) {
    matrix m;
    std::string line;
    while (std::getline(ist, line)) {
        std::vector<double> data(count);
        std::istringstream iss(line);
        for (int i = 0; i < count; ++i) iss >> data[i];
        m.push_back(data);
    }
    return m;
}
