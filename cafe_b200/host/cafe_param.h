// cafe_param.h — host mirror of the reference's global state (CafeParam, libtree/family.h:115-172) and
// of the entry points on the likelihood path.  The bodies that the reference runs on the CPU
// (reset_birthdeath_cache, get_posterior, cafe_conditional_distribution, the p-value loop) are
// forwarded to the CUDA library through the C-ABI of include/cafe_gpu.h — there is no CPU path.
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
#include <vector>

#include "../../include/cafe_gpu.h"
#include "cafe_family.h"
#include "cafe_tree.h"
#include "fminsearch.h"

#define FAMILYSIZEMAX 1000  // libtree/family.h:8

enum OPTIMIZER_INIT_TYPE { UNKNOWN, DO_NOTHING, LAMBDA_ONLY, LAMBDA_MU };

struct input_values {  // libtree/input_values.h
    std::vector<double> store;
    double* parameters = nullptr;
    void construct(int n) { store.assign(n > 0 ? n : 1, 0.0); parameters = store.data(); }
};

struct CafeParam {  // libtree/family.h:115-172 (fields used by the lambda / lambdamu / pvalue paths)
    FILE* fout = stdout;
    FILE* flog = stdout;
    std::string str_fdata;
    pCafeTree pcafe = nullptr;
    pCafeFamily pfamily = nullptr;
    int eqbg = 0;
    int posterior = 0;
    std::vector<double> prior_rfsize_store;
    double* prior_rfsize = nullptr;  // FAMILYSIZEMAX values once set
    input_values input;
    int num_params = 0;
    OPTIMIZER_INIT_TYPE optimizer_init_type = LAMBDA_ONLY;
    double* lambda = nullptr;
    std::vector<int> lambda_tree;  // per node taxaid (label-1) of the -t tree; empty = none (param->lambda_tree == NULL)
    std::string lambda_tree_string;
    int num_lambdas = 0;
    double* mu = nullptr;
    int num_mus = 0;
    int parameterized_k_value = 0;  // the clustered (-k) model is out of scope: must stay 0
    int fixcluster0 = 0;
    int checkconv = 0;
    int num_branches = 0;
    family_size_range family_size{0, 1, 0, 1};
    double pvalue = 0.01;
    int num_threads = 1;
    int num_random_samples = 1000;  // Globals::num_random_samples
    int quiet = 0;
    std::vector<std::vector<double>> cond_dist;  // ConditionalDistribution::matrix
    std::vector<double> max_pvalues;             // viterbi.maximumPvalues
    std::vector<std::vector<double>> likelihoodRatios;  // [node][family], libtree/family.h:169
    std::vector<std::vector<double>> cutPvalues;        // [node][family], viterbi_parameters::cutPvalues (cafe/viterbi.h:52)
    // 1: key the lengthened branches of the likelihood-ratio test with the tree-level mu (0 after cafe_tree_new), as the stock
    // reference does because cafe_tree_node_copy drops the nodes' mu (cafe/cafe_tree.c:485-494); 0: the nodes' own mu
    int lrt_tree_level_mu = 0;
    int objective_calls = 0;
};
typedef CafeParam* pCafeParam;

// cafe/cafe_main.c:26-44
void cafe_log(pCafeParam param, const char* msg, ...);

// --- the GPU engine behind the reference's entry points (replaces the global probability_cache) ---
cafe_gpu_ctx* cafe_gpu_engine();   // lazily created; throws std::runtime_error when no device
void cafe_gpu_engine_release();
std::vector<int> cafe_gpu_engine_devices();  // CAFE_GPUS as a device list; {-1} (the current device) when unset
// push tree / ranges / families / error models / prior to the device when they changed
// with_error_models = false: the passes the reference runs on a copy of the tree, which drops the leaves' error models (cafe_param.cpp)
void cafe_gpu_sync_state(pCafeFamily pfamily, pCafeTree pcafe, const double* prior_rfsize, bool with_error_models = true);

// cafe/cafe_main.c:319-326 — rebuild every (int t, lambda, mu) matrix for the tree's per-node rates   [K1]
void reset_birthdeath_cache(pCafeTree tree, int k_value, family_size_range* range);
void cafe_free_birthdeath_cache(pCafeTree pcafe);
// cafe/lambda.cpp:691-724 — throws std::runtime_error("WARNING: Calculated posterior probability for
// family <id> = 0") for the first zero family; sets pitem->maxlh on first use                       [K2+K3]
double get_posterior(pCafeFamily pfamily, pCafeTree pcafe, std::vector<double>& prior_rfsize);
// cafe/cafe_tree.c:320-329 for every family at once: row-major [flist.size()][rfsize]
std::vector<double> compute_tree_likelihoods_all(pCafeFamily pfamily, pCafeTree pcafe);

// cafe/cafe_shell.c:31-38,194-287
void cafe_shell_set_lambdas(pCafeParam param, double* parameters);
void cafe_shell_set_lambda(pCafeParam param, double* parameters);
void cafe_shell_set_lambda_mu(pCafeParam param, double* parameters);
// cafe/cafe_shell.c:334-393 — returns 0, sets param->lambda_tree / num_lambdas; throws on mismatch
int __cafe_cmd_lambda_tree(pCafeParam param, const char* arg1, const char* arg2);

// cafe/lambda.cpp:771-870
struct poisson_lambda { std::vector<double> parameters; int num_params; int num_iterations; double score; };
std::vector<int> collect_leaf_sizes(pCafeFamily pfamily);
poisson_lambda find_poisson_lambda(pCafeFamily pfamily);
void cafe_set_prior_rfsize_poisson_lambda(std::vector<double>& prior_rfsize, int shift, double* lambda);
double cafe_set_prior_rfsize_empirical(pCafeParam param, std::vector<double>& prior_rfsize);

// cafe/cafe_main.c:124-163
void input_values_randomize(input_values* vals, int lambda_len, int mu_len, int k, int kfix, double max_branch_length,
                            double* k_weights);

// objective callbacks (math_func) and search drivers — the drop-in seam B1
double __cafe_best_lambda_search(double* plambda, void* args);                    // cafe/lambda.cpp:726-769
double* cafe_best_lambda_by_fminsearch(pCafeParam param, int lambda_len, int k);  // cafe/lambda.cpp:525-647
double cafe_best_lambda_mu_search(double* parameters, void* args);                // cafe/lambdamu.cpp:323-367
void best_lambda_mu_by_fminsearch(pCafeParam param, int lambda_len, int mu_len, int k, std::ostream& log);  // lambdamu.cpp:370-473

// conditional distribution + family p-values (cafe/conditional_distribution.cpp:86-120, cafe/pvalue.cpp:143-154,
// cafe/viterbi.cpp:32-39,88-97)                                                                      [K4, K5]
typedef std::vector<std::vector<double>> matrix;
matrix cafe_conditional_distribution(pCafeTree pTree, family_size_range* range, int numthreads, int num_random_samples);
void cafe_family_pvalues(pCafeParam param, std::vector<double>& max_pvalues);
// cafe/cafe_main.c:398-431 (+ the per-family body :342-396) — branch-stretch likelihood-ratio test of `report ... likelihood`:
// fills param->likelihoodRatios[b][i]; -1 for the root's row and for families whose maximumPvalues[i] > param->pvalue;
// duplicates copy their first occurrence.  All families at once through cafe_gpu_likelihood_ratio_test.
void cafe_likelihood_ratio_test(pCafeParam param, double* maximumPvalues);
// cafe/branch_cutting.cpp:101-272 — `report ... branchcutting`: param->cutPvalues[b][i] = p-value of cutting the branch above node b
// for family i; -1 for the root's row and for families whose family-wide p-value exceeds param->pvalue; needs param->max_pvalues.
// param->lrt_tree_level_mu selects the stock keying of the copies' nodes, as for the likelihood-ratio test (branch_cutting.cpp).
void cafe_branch_cutting(pCafeParam param, int num_random_samples);
// Viterbi pass + text report of `report` (cafe/viterbi.h, cafe/viterbi.cpp:88-173,570-595, cafe/reports.cpp:157-194,230-500)
struct change { int expand = 0, remain = 0, decrease = 0; };  // cafe/viterbi.h
struct viterbi_parameters {
    int num_nodes = 0, num_rows = 0;
    std::vector<std::vector<int>> node_sizes;         // [family][node]: viterbiNodeFamilysizes + the observed leaves
    std::vector<std::vector<double>> viterbiPvalues;  // [family][2j+k] for child k of internal node 2j+1; -1 when filtered
    std::vector<double> maximumPvalues;
    std::vector<double> averageExpansion;             // [2j+k]
    std::vector<change> expandRemainDecrease;         // [2j+k]
};
// cafe_gpu_viterbi_report for every family: node sizes and per-branch p-values, row-major [family][node]
void cafe_viterbi_all(pCafeParam param, std::vector<int>& node_sizes, std::vector<double>& branch_pvalues);
void cafe_viterbi(pCafeParam param, viterbi_parameters& viterbi);  // needs param->max_pvalues (cafe_family_pvalues)
void cafe_viterbi_from(pCafeParam param, const std::vector<int>& node_sizes, const std::vector<double>& branch_pvalues,
                       viterbi_parameters& viterbi);  // the host part of cafe_viterbi (no device work)
void cafe_report_text(std::ostream& ost, pCafeParam param, const viterbi_parameters& viterbi);
// cafe/pvalue.cpp:63-93, cafe/cafe_commands.cpp:1373-1396 — text format of `pvalue -o / -i`
void write_pvalues(std::ostream& ost, const matrix& cd, int count);
matrix read_pvalues(std::istream& ist, int count);
