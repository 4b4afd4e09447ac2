// branch_cutting.cpp — host mirror of `report ... branchcutting` (cafe/branch_cutting.cpp:101-272).
//
// For every non-root branch b the reference cuts the tree above node b (cafe_tree_split -> phylogeny_split_tree,
// libtree/phylogeny.c:571-614: the cut-off subtree, and the remaining tree in which b's sibling replaces b's parent and
// inherits its branch length), simulates a conditional distribution for each side that is more than a single leaf
// (cut_branch, :185-219: num_random_samples draws, or a tenth of that for each side when both are trees), and gives every family
// whose family-wide p-value passed the cutoff the p-value of the cut (compute_cutpvalues, :101-150).
// The stock driver cafe_branch_cutting cannot run (it never fills pfamily / viterbi / num_random_samples / pvalue of its thread
// parameters, :247-257); what is mirrored here is what its thread function would do with them filled — the same call sequence
// oracle/ref_shim.cpp::refshim_branch_cut uses to pin the oracle against the reference, bit for bit.
//
// Device side: each side of a cut is a tree of its own, so it gets a context of its own (matrices K1, conditional distribution
// K4, root likelihood rows of the tested families K2), and cafe_gpu_cut_pvalues does the p-value double loop.  Error models are
// not attached to the copies (cafe_tree_node_copy, cafe/cafe_tree.c:485-494, copies lambda, family size and matrix only), and
// with param->lrt_tree_level_mu the copies' nodes carry the tree-level mu as in the stock binary (same defect as the
// likelihood-ratio test, see cafe_param.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "cafe_math.h"
#include "cafe_param.h"

namespace {

struct SideTree {  // one side of a cut in its own infix numbering (tree_build_node_list)
    std::vector<int> left, right, orig;  // orig: node id in the uncut tree
    std::vector<double> bl;
    int n() const { return (int)left.size(); }
};

void ck(cafe_gpu_ctx* ctx, int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string("cafe_gpu ") + what + ": " + cafe_gpu_last_error(ctx));
}

SideTree flatten(const std::vector<int>& left, const std::vector<int>& right, const std::vector<double>& bl, int root) {
    std::vector<int> order;
    std::function<void(int)> infix = [&](int v) {
        if (left[v] >= 0) { infix(left[v]); order.push_back(v); infix(right[v]); }
        else order.push_back(v);
    };
    infix(root);
    std::vector<int> nw(left.size(), -1);
    for (size_t i = 0; i < order.size(); ++i) nw[order[i]] = (int)i;
    SideTree s;
    for (int v : order) {
        s.left.push_back(left[v] >= 0 ? nw[left[v]] : -1);
        s.right.push_back(right[v] >= 0 ? nw[right[v]] : -1);
        s.bl.push_back(v == root ? -1.0 : bl[v]);  // phylogeny.c:611-612
        s.orig.push_back(v);
    }
    return s;
}

std::string side_string(const CafeTree& t, const SideTree& s) {  // name:branchlength, as cafe_tree_string without family sizes
    std::ostringstream o;
    std::function<void(int)> rec = [&](int v) {
        if (s.left[v] >= 0) { o << "("; rec(s.left[v]); o << ","; rec(s.right[v]); o << ")"; }
        o << t.nlist[s.orig[v]].name;
        if (s.bl[v] >= 0) o << ":" << s.bl[v];
    };
    int root = 0;
    std::vector<char> is_child(s.n(), 0);
    for (int v = 0; v < s.n(); ++v) if (s.left[v] >= 0) { is_child[s.left[v]] = 1; is_child[s.right[v]] = 1; }
    for (int v = 0; v < s.n(); ++v) if (!is_child[v]) root = v;
    rec(root);
    return o.str();
}

// one side of a cut on the device
struct Side {
    cafe_gpu_ctx* ctx = nullptr;
    ~Side() { if (ctx) cafe_gpu_destroy(ctx); }
    void setup(const CafeTree& t, const SideTree& s, const family_size_range& rg, bool tree_level_mu, int device) {
        ck(nullptr, cafe_gpu_create(&ctx, device), "create");
        ck(ctx, cafe_gpu_set_tree(ctx, s.n(), s.left.data(), s.right.data(), s.bl.data()), "set_tree");
        ck(ctx, cafe_gpu_set_ranges(ctx, rg.min, rg.max, rg.root_min, rg.root_max), "set_ranges");
        const int maxfs = std::max(rg.max, rg.root_max);
        std::vector<double> T = cafe::lnc_table(maxfs);
        ck(ctx, cafe_gpu_set_lnc_table(ctx, T.data(), 2 * maxfs, maxfs + 1), "set_lnc_table");
        std::vector<double> lam(s.n()), mu(s.n());
        for (int v = 0; v < s.n(); ++v) {
            lam[v] = t.nlist[s.orig[v]].birth_death_probabilities.lambda;
            mu[v] = tree_level_mu ? 0.0 : t.nlist[s.orig[v]].birth_death_probabilities.mu;
        }
        ck(ctx, cafe_gpu_set_rates(ctx, lam.data(), mu.data()), "set_rates");
        ck(ctx, cafe_gpu_build_matrices(ctx), "build_matrices");
        std::vector<double> unit(rg.root_max - rg.root_min + 1, 1.0);
        ck(ctx, cafe_gpu_set_prior(ctx, unit.data(), (int)unit.size()), "set_prior");
    }
    // cafe_conditional_distribution on this side (cafe/conditional_distribution.cpp:86-120), rows flattened
    std::vector<double> distribution(const SideTree& s, const family_size_range& rg, int n_samples, bool replay, uint64_t seed) {
        const int R = rg.root_max - rg.root_min + 1;
        std::vector<double> flat((size_t)R * n_samples);
        if (replay) {  // the draws the single-threaded reference would make, in its order
            std::vector<double> u((size_t)R * n_samples * (s.n() - 1));
            for (double& x : u) x = cafe::unifrnd();
            ck(ctx, cafe_gpu_conditional_distribution(ctx, n_samples, u.data(), 0, flat.data()), "conditional_distribution");
        } else {
            ck(ctx, cafe_gpu_conditional_distribution(ctx, n_samples, nullptr, seed, flat.data()), "conditional_distribution");
        }
        return flat;
    }
    // root likelihood rows of the tested families (set_size_for_split + compute_tree_likelihoods, :125-135)
    std::vector<double> likelihood_rows(const CafeTree& t, const SideTree& s, pCafeFamily fam, const std::vector<int>& tested,
                                        const std::vector<int>& node_species, int R) {
        const int nl = (s.n() + 1) / 2;
        std::vector<int32_t> counts(tested.size() * nl);
        for (size_t j = 0; j < tested.size(); ++j)
            for (int k = 0; k < nl; ++k) counts[j * nl + k] = fam->flist[tested[j]].count[node_species[s.orig[2 * k]]];
        ck(ctx, cafe_gpu_set_families(ctx, (int)tested.size(), nl, counts.data(), nullptr, nullptr), "set_families");
        std::vector<double> L(tested.size() * (size_t)R);
        ck(ctx, cafe_gpu_family_likelihoods(ctx, L.data()), "family_likelihoods");
        (void)t;
        return L;
    }
};

}  // namespace

// phylogeny_split_tree on the flat tree: (rest, sub) for a non-root node b
static void split_tree(const CafeTree& t, int b, SideTree& rest, SideTree& sub) {
    const int n = t.num_nodes();
    std::vector<int> left(n), right(n), parent(n);
    std::vector<double> bl(n);
    for (int v = 0; v < n; ++v) { left[v] = t.nlist[v].left; right[v] = t.nlist[v].right; parent[v] = t.nlist[v].parent; bl[v] = t.nlist[v].branchlength; }
    const int par = parent[b];
    const int sib = (left[par] == b) ? right[par] : left[par];
    int rest_root = t.root;
    if (par == t.root) {
        rest_root = sib;                       // :586-591
    } else {
        bl[sib] += bl[par];                    // :594
        const int grand = parent[par];
        if (left[grand] == par) left[grand] = sib; else right[grand] = sib;
    }
    rest = flatten(left, right, bl, rest_root);
    sub = flatten(left, right, bl, b);
}

void cafe_branch_cutting(pCafeParam param, int num_random_samples) {
    cafe_log(param, "Running Branch Cutting....\n");
    if (param->max_pvalues.size() != param->pfamily->flist.size()) throw std::runtime_error("branch cutting: family-wide p-values not computed");
    const CafeTree& t = *param->pcafe;
    pCafeFamily fam = param->pfamily;
    const int nnodes = t.num_nodes();
    const size_t nrows = fam->flist.size();
    const family_size_range rg = param->family_size;
    const int R = rg.root_max - rg.root_min + 1;
    const char* mode = std::getenv("CAFE_GPU_CD_RNG");
    const bool replay = mode ? std::strcmp(mode, "replay") == 0 : param->num_threads <= 1;

    std::vector<int> node_species(nnodes, -1);  // leaf node id -> column of the family table (set_size_for_split, :61-87)
    for (int i = 0; i < fam->num_species; ++i)
        if (fam->index[i] >= 0 && fam->index[i] < nnodes) node_species[fam->index[i]] = i;
    for (int v = 0; v < nnodes; v += 2)
        if (node_species[v] < 0) throw std::runtime_error("Warning: Tree and family indices not synchronized");

    // families to compute: first occurrences (:113-114) that passed the cutoff (:115-119)
    std::vector<int> tested;
    for (size_t i = 0; i < nrows; ++i) {
        const int ref = fam->flist[i].ref;
        if (ref >= 0 && ref != (int)i) continue;
        if (!(param->max_pvalues[i] > param->pvalue)) tested.push_back((int)i);
    }

    param->cutPvalues.assign(nnodes, std::vector<double>(nrows, 0.0));
    std::vector<std::string> logs(nnodes);
    const bool tree_level_mu = param->lrt_tree_level_mu != 0;
    // device-generator seeds of the two distributions of every branch, drawn here in branch order: the result does not depend
    // on how the branches are spread over devices (unused with the rand() replay)
    std::vector<uint64_t> seeds(2 * (size_t)nnodes, 0);
    if (!replay)
        for (uint64_t& x : seeds) x = ((uint64_t)std::rand() << 32) ^ (uint64_t)std::rand();
    // one branch: both sides of the cut on `device`, the row of cutPvalues and the log text of cut_branch
    auto do_branch = [&](int b, int device) {
        if (b == t.root) { std::fill(param->cutPvalues[b].begin(), param->cutPvalues[b].end(), -1.0); return; }  // :236-240
        SideTree rest, sub;
        split_tree(t, b, rest, sub);
        std::ostringstream ost;
        ost << ">> " << b << "  --------------------\n" << side_string(t, rest) << "\n" << side_string(t, sub) << "\n";
        logs[b] = ost.str();
        std::vector<double> cut(tested.size(), 0.0);
        const bool one = sub.n() == 1 || rest.n() == 1;
        if (one) {
            const SideTree& s = (sub.n() == 1) ? rest : sub;  // :192-201
            Side side;
            side.setup(t, s, rg, tree_level_mu, device);
            std::vector<double> cd = side.distribution(s, rg, num_random_samples, replay, seeds[2 * b]);
            if (!tested.empty()) {
                std::vector<double> L = side.likelihood_rows(t, s, fam, tested, node_species, R);
                ck(side.ctx, cafe_gpu_cut_pvalues(side.ctx, L.data(), nullptr, (int)tested.size(), R, cd.data(), nullptr, num_random_samples, cut.data()), "cut_pvalues");
            }
        } else {
            const int n10 = num_random_samples / 10;  // :204
            if (n10 < 1) throw std::runtime_error("branch cutting: fewer than 10 random samples");
            Side a, c;
            a.setup(t, rest, rg, tree_level_mu, device);
            c.setup(t, sub, rg, tree_level_mu, device);
            std::vector<double> cd1 = a.distribution(rest, rg, n10, replay, seeds[2 * b]);
            std::vector<double> cd2 = c.distribution(sub, rg, n10, replay, seeds[2 * b + 1]);
            if (!tested.empty()) {
                std::vector<double> L1 = a.likelihood_rows(t, rest, fam, tested, node_species, R);
                std::vector<double> L2 = c.likelihood_rows(t, sub, fam, tested, node_species, R);
                ck(a.ctx, cafe_gpu_cut_pvalues(a.ctx, L1.data(), L2.data(), (int)tested.size(), R, cd1.data(), cd2.data(), n10, cut.data()), "cut_pvalues");
            }
        }
        std::vector<double>& row = param->cutPvalues[b];
        for (size_t i = 0; i < nrows; ++i) {
            const int ref = fam->flist[i].ref;
            if (ref >= 0 && ref != (int)i) continue;
            if (param->max_pvalues[i] > param->pvalue) row[i] = -1.0;
        }
        for (size_t j = 0; j < tested.size(); ++j) row[tested[j]] = cut[j];
    };
    // The branches are independent.  With the rand() replay they must run in order (the stream is shared); with the device
    // generator and several devices (CAFE_GPUS) every device takes every n-th branch on a host thread of its own - the reference
    // spreads the FAMILIES of a branch over its pthreads instead (:242-258), the result is the same.
    const std::vector<int> devices = cafe_gpu_engine_devices();
    if (replay || devices.size() < 2) {
        for (int b = 0; b < nnodes; ++b) do_branch(b, devices[0]);
    } else {
        std::vector<std::string> errors(devices.size());
        std::vector<std::thread> workers;
        for (size_t d = 0; d < devices.size(); ++d)
            workers.emplace_back([&, d]() {
                try {
                    for (int b = (int)d; b < nnodes; b += (int)devices.size()) do_branch(b, devices[d]);
                } catch (std::exception& e) { errors[d] = e.what(); }
            });
        for (std::thread& w : workers) w.join();
        for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error(e);
    }
    for (int b = 0; b < nnodes; ++b) if (!logs[b].empty()) cafe_log(param, "%s", logs[b].c_str());
    for (size_t i = 0; i < nrows; ++i) {  // duplicates take their first occurrence's values, :259-267
        const int ref = fam->flist[i].ref;
        if (ref < 0 || ref == (int)i) continue;
        for (int b = 0; b < nnodes; ++b) param->cutPvalues[b][i] = param->cutPvalues[b][ref];
    }
    cafe_log(param, "Done : Branch Cutting\n");
}
