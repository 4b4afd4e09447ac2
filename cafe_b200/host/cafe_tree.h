// cafe_tree.h — host mirror of the reference's tree structures for the likelihood path.
//
// Same names and meaning as libtree/family.h:10-85 (family_size_range, CafeTree, CafeNode) and
// libtree/tree.h (PhylogenyNode fields), but stored flat: nodes live in one vector in the reference's
// nlist (infix) order — leaves at even indices, internal nodes at odd indices
// (cafe/cafe_commands.cpp:1985-2051) — which is exactly what the C-ABI (include/cafe_gpu.h) takes.
#pragma once
#include <string>
#include <vector>

struct family_size_range {  // libtree/family.h:10-15
    int min, max, root_min, root_max;
};

struct probabilities {  // libtree/family.h:40-46 (the scalar pair; the clustered -k arrays are out of scope)
    double lambda = 0, mu = 0;
};

struct CafeNode {  // libtree/family.h:53-78 + PhylogenyNode libtree/tree.h
    int id = -1;
    int parent = -1, left = -1, right = -1;  // children->head / children->tail
    std::string name;
    double branchlength = -1;
    int taxaid = -1;
    int familysize = -1;
    probabilities birth_death_probabilities;
    int errormodel = -1;  // index into CafeFamily::errors, -1 = none
    bool is_leaf() const { return left < 0; }
};

struct CafeTree {  // libtree/family.h:17-28
    std::vector<CafeNode> nlist;  // infix order
    int root = -1;
    family_size_range range{0, 1, 0, 1};
    double lambda = 0, mu = 0;
    int k = 0;
    int size_of_factor = 0;
    int rfsize = 0;
    std::vector<int> prefix;   // node ids in prefix order (libtree/tree.c:101-124)
    std::vector<int> postfix;  // node ids in postfix order

    int num_nodes() const { return (int)nlist.size(); }
    int num_leaves() const { return ((int)nlist.size() + 1) / 2; }
};
typedef CafeTree* pCafeTree;

// cafe/cafe_commands.cpp:2076-2107.  Throws std::runtime_error on malformed or non-binary input
// ("Tree must be binary", cafe_commands.cpp:1994-1996).
pCafeTree cafe_tree_new(const char* sztree, family_size_range* range, double lambda, double mu);
void cafe_tree_free(pCafeTree pcafe);
pCafeTree cafe_tree_copy(pCafeTree psrc);
// cafe/cafe_main.c:52-60
void copy_range_to_tree(pCafeTree tree, family_size_range* range);
// cafe/cafe_tree.c:46-67
void cafe_tree_set_parameters(pCafeTree pcafe, family_size_range* range, double lambda);
// cafe/cafe_commands.cpp:1099-1120 — an int, and it throws when a non-root branch length is missing
int max_branch_length(pCafeTree ptree);
// cafe/cafe_commands.cpp:1075-1089 (0.01 % tolerance)
bool is_ultrametric(pCafeTree ptree);
// Newick with the topology of `like` and integer labels -> per-node taxaid (label-1), cafe/cafe_shell.c:324-393.
// Returns the number of distinct labels (param->num_lambdas); throws on topology mismatch / missing labels.
int parse_lambda_tree(const char* sztree, const CafeTree& like, std::vector<int>& taxaid_per_node);
std::string cafe_tree_string(const CafeTree& t);
