#include "cafe_tree.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace {

struct RawNode {
    std::string name;
    double branchlength = -1;
    std::vector<std::unique_ptr<RawNode>> kids;
};

// Recursive-descent Newick reader: names, ':length', '[...]' comments (NHX tags are skipped — none of
// them reaches the likelihood path), optional trailing ';', blanks ignored.
class NewickReader {
   public:
    explicit NewickReader(const std::string& s) {
        for (char c : s)
            if (!std::isspace((unsigned char)c)) text_.push_back(c);
        while (!text_.empty() && text_.back() == ';') text_.pop_back();
        int depth = 0;
        for (char c : text_) {
            if (c == '(') ++depth;
            if (c == ')') --depth;
            if (depth < 0) break;
        }
        if (depth != 0) throw std::runtime_error("Tree error (Unbalanced parentheses): " + s);
    }
    std::unique_ptr<RawNode> parse() {
        if (text_.empty()) throw std::runtime_error("Failed to load tree from provided string");
        auto root = node();
        if (pos_ != text_.size()) throw std::runtime_error("Failed to load tree from provided string");
        return root;
    }

   private:
    std::unique_ptr<RawNode> node() {
        auto n = std::make_unique<RawNode>();
        if (peek() == '(') {
            ++pos_;
            for (;;) {
                n->kids.push_back(node());
                if (peek() == ',') { ++pos_; continue; }
                if (peek() == ')') { ++pos_; break; }
                throw std::runtime_error("Failed to load tree from provided string");
            }
        }
        label(*n);
        return n;
    }
    void label(RawNode& n) {
        size_t start = pos_;
        while (pos_ < text_.size() && !std::strchr("(),:[", text_[pos_])) ++pos_;
        n.name = text_.substr(start, pos_ - start);
        skip_comment();
        if (peek() == ':') {
            ++pos_;
            start = pos_;
            while (pos_ < text_.size() && !std::strchr("(),[", text_[pos_])) ++pos_;
            std::string num = text_.substr(start, pos_ - start);
            if (!num.empty()) n.branchlength = std::atof(num.c_str());
            skip_comment();
        }
    }
    void skip_comment() {
        while (peek() == '[') {
            size_t close = text_.find(']', pos_);
            if (close == std::string::npos) throw std::runtime_error("keep format : [&&NHX ... ]");
            pos_ = close + 1;
        }
    }
    char peek() const { return pos_ < text_.size() ? text_[pos_] : '\0'; }
    std::string text_;
    size_t pos_ = 0;
};

// infix numbering: head subtree, node, tail subtree (tree_traveral_infix)
void number_infix(RawNode* n, std::vector<RawNode*>& order) {
    if (n->kids.empty()) { order.push_back(n); return; }
    if (n->kids.size() != 2) throw std::runtime_error("Tree must be binary");
    number_infix(n->kids[0].get(), order);
    order.push_back(n);
    number_infix(n->kids[1].get(), order);
}

void build_orders(CafeTree& t) {
    t.prefix.clear();
    t.postfix.clear();
    std::function<void(int)> walk = [&](int v) {
        t.prefix.push_back(v);
        if (!t.nlist[v].is_leaf()) { walk(t.nlist[v].left); walk(t.nlist[v].right); }
        t.postfix.push_back(v);
    };
    walk(t.root);
}

void flatten(RawNode* rootn, CafeTree& t) {
    std::vector<RawNode*> order;
    number_infix(rootn, order);
    t.nlist.assign(order.size(), CafeNode());
    auto id_of = [&](RawNode* p) { return (int)(std::find(order.begin(), order.end(), p) - order.begin()); };
    for (size_t i = 0; i < order.size(); ++i) {
        CafeNode& n = t.nlist[i];
        n.id = (int)i;
        n.name = order[i]->name;
        n.branchlength = order[i]->branchlength;
        if (!order[i]->kids.empty()) {
            n.left = id_of(order[i]->kids[0].get());
            n.right = id_of(order[i]->kids[1].get());
        }
    }
    for (auto& n : t.nlist)
        if (!n.is_leaf()) { t.nlist[n.left].parent = n.id; t.nlist[n.right].parent = n.id; }
    t.root = id_of(rootn);
    build_orders(t);
}

}  // namespace

pCafeTree cafe_tree_new(const char* sztree, family_size_range* range, double lambda, double mu) {
    NewickReader rd(sztree ? sztree : "");
    auto raw = rd.parse();
    auto t = std::make_unique<CafeTree>();
    flatten(raw.get(), *t);
    if (t->nlist.size() < 3) throw std::runtime_error("Failed to load tree from provided string");
    // "name_size" leaf labels carry a family size (cafe_tree_parse_node, cafe_commands.cpp:2062-2073)
    for (auto& n : t->nlist) {
        size_t us = n.name.find('_');
        if (us != std::string::npos) {
            n.familysize = std::atoi(n.name.c_str() + us + 1);
            n.name.resize(us);
        }
    }
    copy_range_to_tree(t.get(), range);
    int rsize = range->root_max - range->root_min + 1, fsize = range->max - range->min + 1;
    t->size_of_factor = std::max(rsize, fsize);
    t->lambda = lambda;
    t->mu = mu;
    for (auto& n : t->nlist) { n.birth_death_probabilities.lambda = lambda; n.birth_death_probabilities.mu = mu; }
    return t.release();
}

void cafe_tree_free(pCafeTree pcafe) { delete pcafe; }
pCafeTree cafe_tree_copy(pCafeTree psrc) { return new CafeTree(*psrc); }

void copy_range_to_tree(pCafeTree tree, family_size_range* range) {
    tree->range = *range;
    tree->rfsize = range->root_max - range->root_min + 1;
}

void cafe_tree_set_parameters(pCafeTree pcafe, family_size_range* range, double lambda) {
    copy_range_to_tree(pcafe, range);
    pcafe->lambda = lambda;
    int fsize = range->max - range->min + 1;
    pcafe->size_of_factor = std::max(pcafe->size_of_factor, std::max(pcafe->rfsize, fsize));
}

int max_branch_length(pCafeTree ptree) {
    int longest = 0;
    for (const auto& n : ptree->nlist) {
        if (n.branchlength > 0) {
            if (longest < n.branchlength) longest = (int)n.branchlength;  // int, as in the reference
        } else if (n.id != ptree->root) {
            throw std::runtime_error("Failed to load tree from provided string (branch length missing)");
        }
    }
    return longest;
}

bool is_ultrametric(pCafeTree ptree) {
    std::vector<double> depth;
    for (const auto& n : ptree->nlist) {
        if (!n.is_leaf()) continue;
        double d = 0;
        for (int v = n.id; v != ptree->root; v = ptree->nlist[v].parent) d += ptree->nlist[v].branchlength;
        depth.push_back(d);
    }
    double mx = *std::max_element(depth.begin(), depth.end());
    double tol = mx * 0.0001;
    for (double d : depth)
        if (std::fabs(d - mx) > tol) return false;
    return true;
}

int parse_lambda_tree(const char* sztree, const CafeTree& like, std::vector<int>& taxaid_per_node) {
    NewickReader rd(sztree ? sztree : "");
    auto raw = rd.parse();
    CafeTree lt;
    flatten(raw.get(), lt);
    if (lt.nlist.size() != like.nlist.size()) throw std::runtime_error("Lambda has a different topology from the tree");
    taxaid_per_node.assign(lt.nlist.size(), -1);
    int labelled = 0;
    std::vector<int> seen;
    for (size_t i = 0; i < lt.nlist.size(); ++i) {
        int id = -1;  // phylogeny_clear_node leaves taxaid = -1 ... then `taxaid--` (cafe_shell.c:324-332)
        if (!lt.nlist[i].name.empty()) {
            id = std::atoi(lt.nlist[i].name.c_str());
            ++labelled;
        }
        taxaid_per_node[i] = id - 1;
        if (taxaid_per_node[i] >= 0 && std::find(seen.begin(), seen.end(), taxaid_per_node[i]) == seen.end())
            seen.push_back(taxaid_per_node[i]);
    }
    if (labelled != (int)lt.nlist.size() - 1) {
        std::ostringstream o;
        o << "ERROR(lambda -t): Branch lambda classes not totally specified.\n" << sztree
          << "\nYou have to specify lambda classes for all branches including the internal branches of the tree.\n"
          << "There are total " << lt.nlist.size() - 1 << " branches in the tree.\n";
        throw std::runtime_error(o.str());
    }
    return (int)seen.size();
}

std::string cafe_tree_string(const CafeTree& t) {
    std::function<std::string(int)> rec = [&](int v) -> std::string {
        const CafeNode& n = t.nlist[v];
        std::ostringstream o;
        if (!n.is_leaf()) o << "(" << rec(n.left) << "," << rec(n.right) << ")";
        o << n.name;
        if (n.familysize >= 0) o << "_" << n.familysize;
        if (v != t.root && n.branchlength >= 0) o << ":" << n.branchlength;
        return o.str();
    };
    return rec(t.root);
}
