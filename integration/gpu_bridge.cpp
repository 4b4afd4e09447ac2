// gpu_bridge.cpp — the reference-side binding of include/cafe_gpu.h: compiled INTO the unmodified reference (plus two edited
// statements, see oracle/Makefile target `bridge` and INTEGRATION.md §2) it flattens pCafeParam into the C-ABI and runs every
// objective evaluation of `lambda -s` / `lambdamu -s` on the GPU.  This file is the cgo/JNI stub of a project whose host
// language is C++: plain C++11 against the reference's own headers, nothing of cafe_b200/host.
extern "C" {
#include "family.h"     // pCafeParam, pCafeTree, pCafeFamily, pErrorStruct (libtree/family.h)
#include "cafe.h"
#include <mathfunc.h>   // chooseln (libcommon/mathfunc.c:224-229)
}
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "cafe_gpu.h"
#include "gpu_bridge.h"

namespace {

cafe_gpu_ctx* g_gpu = NULL;

void gpu_ck(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string("gpu_bridge: ") + what + ": " + cafe_gpu_last_error(g_gpu) + "\n");
}

// what the device currently holds: re-bound when the session state behind pCafeParam changes (load / tree / errormodel)
struct Bound {
    void* tree; void* family; int n_families; int n_nodes;
    family_size_range range;
    std::vector<void*> err;
    std::vector<double> branchlength;
    Bound() : tree(NULL), family(NULL), n_families(-1), n_nodes(-1) { range.min = range.max = range.root_min = range.root_max = -1; }
};
Bound g_bound;

bool same_binding(pCafeParam param) {
    pArrayList nl = param->pcafe->super.nlist;
    if (g_bound.tree != (void*)param->pcafe || g_bound.family != (void*)param->pfamily) return false;
    if (g_bound.n_families != param->pfamily->flist->size || g_bound.n_nodes != nl->size) return false;
    const family_size_range& r = param->family_size;
    if (r.min != g_bound.range.min || r.max != g_bound.range.max || r.root_min != g_bound.range.root_min || r.root_max != g_bound.range.root_max)
        return false;
    for (int k = 0; k < nl->size; ++k) {
        if ((void*)((pCafeNode)nl->array[k])->errormodel != g_bound.err[k]) return false;
        if (((pPhylogenyNode)nl->array[k])->branchlength != g_bound.branchlength[k]) return false;
    }
    return true;
}

// once per (tree, table, ranges, error models): topology, ranges, lnC table, unique count patterns, error matrices
void bind(pCafeParam param) {
    if (!g_gpu) {
        int rc = cafe_gpu_create(&g_gpu, -1);
        if (rc) throw std::runtime_error(std::string("gpu_bridge: cafe_gpu_create: ") + cafe_gpu_last_error(NULL) + "\n");
    }
    pCafeTree t = param->pcafe;
    pArrayList nl = t->super.nlist;                       // infix order: leaves even, internal odd
    const int n = nl->size;
    std::vector<int32_t> left(n, -1), right(n, -1);
    std::vector<double> bl(n);
    for (int i = 0; i < n; ++i) {
        pTreeNode v = (pTreeNode)nl->array[i];
        if (v->children && v->children->head) {           // cafe/cafe_tree.c:248-249
            left[i] = ((pTreeNode)v->children->head->data)->id;
            right[i] = ((pTreeNode)v->children->tail->data)->id;
        }
        bl[i] = ((pPhylogenyNode)v)->branchlength;        // the (int) truncation happens behind the ABI
        if (!(bl[i] > 0)) bl[i] = 1;                      // the root carries -1; the ABI ignores the root's entry
    }
    gpu_ck(cafe_gpu_set_tree(g_gpu, n, left.data(), right.data(), bl.data()), "set_tree");
    family_size_range* r = &param->family_size;
    gpu_ck(cafe_gpu_set_ranges(g_gpu, r->min, r->max, r->root_min, r->root_max), "set_ranges");
    const int size = r->max > r->root_max ? r->max : r->root_max;       // birthdeath_cache_init, birthdeath.c:331
    std::vector<double> lnc((size_t)2 * size * (size + 1));
    for (int a = 0; a < 2 * size; ++a)
        for (int x = 0; x <= size; ++x) lnc[(size_t)a * (size + 1) + x] = chooseln(a, x);   // Lanczos gammaln, not lgamma
    gpu_ck(cafe_gpu_set_lnc_table(g_gpu, lnc.data(), 2 * size, size + 1), "set_lnc_table");

    pCafeFamily f = param->pfamily;                       // unique patterns via pitem->ref (cafe_family.c:9-34)
    const int n_leaves = (n + 1) / 2;
    std::vector<int> col_of_leaf(n_leaves, 0);
    for (int i = 0; i < f->num_species; ++i)
        if (f->index[i] >= 0) col_of_leaf[f->index[i] / 2] = i;
    std::vector<int32_t> counts, mult, first, uniq_of(f->flist->size);
    for (int i = 0; i < f->flist->size; ++i) {
        pCafeFamilyItem it = (pCafeFamilyItem)f->flist->array[i];
        if (it->ref < 0 || it->ref == i) {
            uniq_of[i] = (int)first.size(); first.push_back(i); mult.push_back(1);
            for (int k = 0; k < n_leaves; ++k) counts.push_back(it->count[col_of_leaf[k]]);
        } else {
            uniq_of[i] = uniq_of[it->ref]; mult[uniq_of[i]]++;
        }
    }
    gpu_ck(cafe_gpu_set_families(g_gpu, (int)first.size(), n_leaves, counts.data(), mult.data(), first.data()), "set_families");
    gpu_ck(cafe_gpu_set_error_model(g_gpu, -1, NULL, 0), "set_error_model");
    for (int k = 0; k < n_leaves; ++k) {                 // dense errormatrix[observed][true], family.h:31-38
        pErrorStruct e = ((pCafeNode)nl->array[2 * k])->errormodel;
        if (!e) continue;
        const int dim = e->maxfamilysize + 1;
        std::vector<double> E((size_t)dim * dim);
        for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j) E[(size_t)i * dim + j] = e->errormatrix[i][j];
        gpu_ck(cafe_gpu_set_error_model(g_gpu, k, E.data(), dim), "set_error_model");
    }
    g_bound.tree = param->pcafe; g_bound.family = param->pfamily; g_bound.n_families = f->flist->size; g_bound.n_nodes = n;
    g_bound.range = *r;
    g_bound.err.resize(n); g_bound.branchlength.resize(n);
    for (int k = 0; k < n; ++k) {
        g_bound.err[k] = (void*)((pCafeNode)nl->array[k])->errormodel;
        g_bound.branchlength[k] = ((pPhylogenyNode)nl->array[k])->branchlength;
    }
}

}  // namespace

double gpu_bridge_score(pCafeParam param) {
    if (!same_binding(param)) bind(param);
    pArrayList nl = param->pcafe->super.nlist;
    std::vector<double> lam(nl->size), mu(nl->size);
    for (int i = 0; i < nl->size; ++i) {                  // left there by cafe_shell_set_lambdas (cafe_shell.c:31)
        lam[i] = ((pCafeNode)nl->array[i])->birth_death_probabilities.lambda;
        mu[i] = ((pCafeNode)nl->array[i])->birth_death_probabilities.mu;
    }
    gpu_ck(cafe_gpu_set_prior(g_gpu, param->prior_rfsize, FAMILYSIZEMAX), "set_prior");
    double score = 0;
    int32_t first_zero = -1;
    const int rc = cafe_gpu_objective(g_gpu, lam.data(), mu.data(), &score, &first_zero);
    gpu_ck(rc, "objective");
    if (rc == CAFE_GPU_ZERO_LIKELIHOOD) {                 // the exception of lambda.cpp:715-720
        pCafeFamilyItem it = (pCafeFamilyItem)param->pfamily->flist->array[first_zero];
        throw std::runtime_error(std::string("WARNING: Calculated posterior probability for family ") + it->id + " = 0\n");
    }
    return score;
}
