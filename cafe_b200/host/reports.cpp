// reports.cpp — host mirror of the Viterbi pass and the text report of `report` (cafe/viterbi.cpp:88-173,570-595,
// cafe/reports.cpp:157-194,230-357,400-500).  The numbers come from the CUDA library (cafe_gpu_viterbi_report, family p-values,
// likelihood ratios); this file only arranges them the way the reference's Report does, quirks included:
//   * the 'cut P-value' / 'Likelihood Ratio' column HEADERS are written when the columns are ABSENT (reports.cpp:445-446 test
//     `.empty()` where `!empty()` was meant);
//   * the branch p-values of a family are stored under id 2j+k for child k of internal node 2j+1 (viterbi.cpp:42-70), so the
//     pairs of a family line are children pairs in internal-node order, as the "# Output format" line says.
#include <cmath>
#include <functional>
#include <ostream>
#include <sstream>
#include <stdexcept>

#include "cafe_param.h"

namespace {
// phylogeny_string (libtree/phylogeny.c:408-470): "(left,right)" + label + ":%g" branch length (>= 0 only)
std::string newick(const CafeTree& t, const std::function<void(std::ostream&, const CafeNode&)>& label, bool with_bl) {
    std::function<void(std::ostream&, int)> rec = [&](std::ostream& o, int v) {
        const CafeNode& n = t.nlist[v];
        if (!n.is_leaf()) {
            o << "(";
            rec(o, n.left);
            o << ",";
            rec(o, n.right);
            o << ")";
        }
        label(o, n);
        if (with_bl && n.branchlength >= 0) o << ":" << n.branchlength;  // ostream's default format is %g
    };
    std::ostringstream o;
    rec(o, t.root);
    return o.str();
}
void write_doubles(std::ostream& ost, const std::vector<double>& items) {  // reports.cpp:240-260
    ost << "(";
    for (size_t i = 0; i < items.size(); ++i) {
        if (items[i] == -1) ost << "-";
        else ost << items[i];
        if (i + 1 < items.size()) ost << ",";
    }
    ost << ")";
}
}  // namespace

// cafe_viterbi (viterbi.cpp:144-173) = viterbi_section for every family (:88-117): forced range, family p-values (already in
// maximumPvalues), Viterbi reconstruction, size deltas (:570-595), branch p-values or -1 when the family is filtered.
// cafe_viterbi_from is the host part: it arranges reconstructed sizes and branch p-values ([family][node]) the way
// viterbi_parameters holds them.
void cafe_viterbi_from(pCafeParam param, const std::vector<int>& sizes, const std::vector<double>& branch_pv, viterbi_parameters& viterbi) {
    const int nnodes = param->pcafe->num_nodes();
    const size_t nrows = param->pfamily->flist.size();
    if (param->max_pvalues.size() != nrows) throw std::runtime_error("cafe_viterbi: family p-values not computed");
    if (sizes.size() != nrows * nnodes || branch_pv.size() != nrows * nnodes) throw std::runtime_error("cafe_viterbi: wrong array sizes");
    viterbi.num_nodes = nnodes;
    viterbi.num_rows = (int)nrows;
    viterbi.maximumPvalues = param->max_pvalues;
    viterbi.node_sizes.assign(nrows, std::vector<int>(nnodes));
    viterbi.viterbiPvalues.assign(nrows, std::vector<double>(nnodes > 1 ? nnodes - 1 : 0, -1.0));
    viterbi.averageExpansion.assign(nnodes - 1, 0.0);
    viterbi.expandRemainDecrease.assign(nnodes - 1, change());
    const CafeTree& t = *param->pcafe;
    for (size_t i = 0; i < nrows; ++i) {
        for (int v = 0; v < nnodes; ++v) viterbi.node_sizes[i][v] = sizes[i * nnodes + v];
        const bool filtered = viterbi.maximumPvalues[i] > param->pvalue;  // viterbi.cpp:105-114
        for (int j = 0; j < (nnodes - 1) / 2; ++j) {
            const CafeNode& pn = t.nlist[2 * j + 1];
            const int child[2] = {pn.left, pn.right};
            for (int k = 0; k < 2; ++k) {
                const int m = 2 * j + k;
                const int d = viterbi.node_sizes[i][child[k]] - viterbi.node_sizes[i][pn.id];
                if (d > 0) viterbi.expandRemainDecrease[m].expand++;
                else if (d == 0) viterbi.expandRemainDecrease[m].remain++;
                else viterbi.expandRemainDecrease[m].decrease++;
                viterbi.averageExpansion[m] += d;
                viterbi.viterbiPvalues[i][m] = filtered ? -1.0 : branch_pv[i * nnodes + child[k]];
            }
        }
    }
    for (double& a : viterbi.averageExpansion) a /= (double)nrows;  // viterbi.cpp:166-169
}

void cafe_viterbi(pCafeParam param, viterbi_parameters& viterbi) {
    cafe_log(param, "Running Viterbi algorithm....\n");
    std::vector<int> sizes;
    std::vector<double> branch_pv;
    cafe_viterbi_all(param, sizes, branch_pv);  // [family][node], cafe_gpu_viterbi_report
    cafe_viterbi_from(param, sizes, branch_pv, viterbi);
}

// operator<<(ostream&, const Report&), text format (reports.cpp:447-500) with the family lines of :339-357
void cafe_report_text(std::ostream& ost, pCafeParam param, const viterbi_parameters& viterbi) {
    const CafeTree& t = *param->pcafe;
    const int nnodes = t.num_nodes();
    ost << "Tree:" << newick(t, [](std::ostream& o, const CafeNode& n) { o << n.name; }, true) << "\n";
    ost << "Lambda:";
    for (int i = 0; i < param->num_lambdas; ++i) ost << "\t" << param->lambda[i];
    ost << "\n";
    if (!param->lambda_tree.empty()) {
        CafeTree lt = t;  // same topology; lambda_tree_string (reports.cpp:131-137) prints taxaid + 1 where it is set
        ost << "Lambda tree:\t"
            << newick(lt, [&](std::ostream& o, const CafeNode& n) { if (param->lambda_tree[n.id] != -1) o << param->lambda_tree[n.id] + 1; }, false)
            << "\n";
    }
    ost << "# IDs of nodes:" << newick(t, [](std::ostream& o, const CafeNode& n) { o << n.name << "<" << n.id << ">"; }, false) << "\n";
    ost << "# Output format for: ' Average Expansion', 'Expansions', 'No Change', 'Contractions', and 'Branch-specific P-values' = (node ID, node ID): ";
    for (int b = 1; b < nnodes; b += 2) ost << "(" << t.nlist[b].left << "," << t.nlist[b].right << ") ";
    ost << "\n";
    ost << "# Output format for 'Branch cutting P-values' and 'Likelihood Ratio Test': (0";
    for (int i = 1; i < nnodes; ++i) ost << ", " << i;
    ost << ")\n";
    // write_viterbi, reports.cpp:157-194
    const size_t npairs = viterbi.averageExpansion.size() / 2;
    ost << "Average Expansion:";
    for (size_t b = 0; b < npairs; ++b) ost << "\t(" << viterbi.averageExpansion[2 * b] << "," << viterbi.averageExpansion[2 * b + 1] << ")";
    ost << "\nExpansion :";
    for (size_t b = 0; b < npairs; ++b) ost << "\t(" << viterbi.expandRemainDecrease[2 * b].expand << "," << viterbi.expandRemainDecrease[2 * b + 1].expand << ")";
    ost << "\nnRemain :";
    for (size_t b = 0; b < npairs; ++b) ost << "\t(" << viterbi.expandRemainDecrease[2 * b].remain << "," << viterbi.expandRemainDecrease[2 * b + 1].remain << ")";
    ost << "\nnDecrease :";
    for (size_t b = 0; b < npairs; ++b) ost << "\t(" << viterbi.expandRemainDecrease[2 * b].decrease << "," << viterbi.expandRemainDecrease[2 * b + 1].decrease << ")";
    ost << "\n";
    const bool have_lr = !param->likelihoodRatios.empty(), have_cut = !param->cutPvalues.empty();
    // write_families_header with the inverted flags of reports.cpp:456-457: a column's title appears when the column is ABSENT
    ost << "'ID'\t'Newick'\t'Family-wide P-value'\t'Viterbi P-values'";
    if (!have_cut) ost << "\t'cut P-value'";
    if (!have_lr) ost << "\t'Likelihood Ratio'";
    ost << "\n";
    CafeTree ft = t;
    for (size_t i = 0; i < param->pfamily->flist.size(); ++i) {
        for (int v = 0; v < nnodes; ++v) ft.nlist[v].familysize = viterbi.node_sizes[i][v];  // cafe_report_set_viterbi, :139-144
        ost << param->pfamily->flist[i].id << "\t";
        ost << newick(ft, [](std::ostream& o, const CafeNode& n) { o << n.name; if (n.familysize >= 0) o << "_" << n.familysize; }, true) << "\t";
        ost << viterbi.maximumPvalues[i] << "\t(";
        for (size_t b = 0; b < npairs; ++b) {
            const double p1 = viterbi.viterbiPvalues[i][2 * b], p2 = viterbi.viterbiPvalues[i][2 * b + 1];
            if (p1 < 0) ost << "(-,-)";
            else ost << "(" << p1 << "," << p2 << ")";
            if (b + 1 < npairs) ost << ",";
        }
        ost << ")\t";
        if (have_cut) {  // reports.cpp:340-344
            std::vector<double> cp(nnodes);
            for (int b = 0; b < nnodes; ++b) cp[b] = param->cutPvalues[b][i];
            write_doubles(ost, cp);
            ost << "\t";
        }
        if (have_lr) {
            std::vector<double> lr(nnodes);
            for (int b = 0; b < nnodes; ++b) lr[b] = param->likelihoodRatios[b][i];
            write_doubles(ost, lr);
        }
        ost << "\n";
    }
}
