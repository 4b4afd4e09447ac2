// cafe_family.h — host mirror of the reference's family table for the likelihood path
// (libtree/family.h:31-38,87-104, cafe/cafe.h:11-27, cafe/gene_family.cpp, cafe/cafe_family.c,
// cafe/error_model.cpp:145-259).  The table ends up on the GPU as a packed int32 matrix of the
// UNIQUE count patterns plus multiplicities (cafe_gpu_set_families).
#pragma once
#include <iosfwd>
#include <string>
#include <vector>

#include "cafe_tree.h"

struct ErrorStruct {  // libtree/family.h:31-38; errormatrix dense row-major, [observed][true]
    std::string errorfilename;
    int fromdiff = 0, todiff = 0, maxfamilysize = 0;
    std::vector<double> errormatrix;  // (maxfamilysize+1)^2
    double at(int observed, int truth) const { return errormatrix[(size_t)observed * (maxfamilysize + 1) + truth]; }
    double& at(int observed, int truth) { return errormatrix[(size_t)observed * (maxfamilysize + 1) + truth]; }
};
typedef ErrorStruct* pErrorStruct;

struct CafeFamilyItem {  // cafe/cafe.h:11-27
    std::string id, desc;
    std::vector<int> count;  // in species (table column) order
    int maxlh = -1;
    int ref = -1;  // index of the first family with identical counts (cafe_family.c:9-34)
    int holder = 1;
};
typedef CafeFamilyItem* pCafeFamilyItem;

struct CafeFamily {  // libtree/family.h:87-104
    std::vector<std::string> species;
    int num_species = 0;
    std::vector<int> index;      // species -> node id in the tree (nlist), cafe_family_set_species_index
    std::vector<int> error_ptr;  // species -> index into errors, -1 = none
    int max_size = 0;
    std::vector<CafeFamilyItem> flist;
    std::vector<ErrorStruct> errors;
};
typedef CafeFamily* pCafeFamily;

struct gene_family {  // cafe/gene_family.h:28-44
    std::string id, desc;
    std::vector<int> values;
};

pCafeFamily cafe_family_init(const std::vector<std::string>& species_list);
void cafe_family_free(pCafeFamily pcf);
void cafe_family_add_item(pCafeFamily pcf, const gene_family& gf);
// cafe/gene_family.cpp:186-225 (+ duplicate detection).  Throws std::runtime_error like the reference.
pCafeFamily load_gene_families(std::istream& ist, char separator, int max_size);
// cafe/cafe_family.c:9-34 — same `ref` semantics (lowest index wins), hashed instead of O(F^2)
void __cafe_famliy_check_the_pattern(pCafeFamily pcf);
// cafe/gene_family.cpp:413-445
void cafe_family_set_species_index(pCafeFamily pcf, pCafeTree pcafe);
// cafe/cafe_family.c:357-364
void init_family_size(family_size_range* fs, int max);
// cafe/cafe_family.c:211-234
void cafe_family_set_size(pCafeFamily pcf, pCafeFamilyItem pitem, pCafeTree pcafe);
// cafe/cafe_family.c:236-255
void cafe_family_set_size_with_family_forced(pCafeFamily pcf, int idx, pCafeTree pcafe);
void cafe_family_reset_maxlh(pCafeFamily pcf);

// error model: reader (cafe/error_model.cpp:145-204), column-sum fix (cafe/cafe_shell.c:585-622),
// attachment to species / tree leaves (cafe/error_model.cpp:206-259)
std::istream& operator>>(std::istream& ifst, ErrorStruct& errormodel);
int __check_error_model_columnsums(pErrorStruct errormodel);
int set_error_matrix_from_file(pCafeFamily family, pCafeTree pTree, family_size_range& range, std::string filename,
                               std::string speciesname);
int remove_error_model(pCafeFamily family, pCafeTree pcafe, std::string species_name);
