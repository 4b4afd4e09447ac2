#include "cafe_family.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace {

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream is(s);
    while (std::getline(is, cur, sep)) out.push_back(cur);
    return out;
}

std::string lower(const std::string& s) {
    std::string r(s);
    for (auto& c : r) c = (char)std::tolower((unsigned char)c);
    return r;
}

void chomp(std::string& s) {
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
}

struct CountsHash {
    size_t operator()(const std::vector<int>& v) const {
        size_t h = 1469598103934665603ull;  // FNV-1a over the ints
        for (int x : v) { h ^= (size_t)(unsigned)x; h *= 1099511628211ull; }
        return h;
    }
};

}  // namespace

pCafeFamily cafe_family_init(const std::vector<std::string>& species_list) {
    pCafeFamily pcf = new CafeFamily();
    pcf->species = species_list;
    pcf->num_species = (int)species_list.size();
    pcf->index.assign(pcf->num_species, -1);
    pcf->error_ptr.assign(pcf->num_species, -1);
    return pcf;
}

void cafe_family_free(pCafeFamily pcf) { delete pcf; }

void cafe_family_add_item(pCafeFamily pcf, const gene_family& gf) {
    CafeFamilyItem item;
    item.id = gf.id;
    item.desc = gf.desc;
    if (gf.values.size() != (size_t)pcf->num_species)
        std::cerr << "Inconsistency in column count: expected " << pcf->num_species + 2 << ", but found " << gf.values.size() + 2;
    item.count.assign(pcf->num_species, 0);
    std::copy(gf.values.begin(), gf.values.begin() + std::min(gf.values.size(), (size_t)pcf->num_species), item.count.begin());
    pcf->max_size = std::max(pcf->max_size, *std::max_element(item.count.begin(), item.count.end()));
    pcf->flist.push_back(std::move(item));
}

pCafeFamily load_gene_families(std::istream& ist, char separator, int max_size) {
    if (!ist) return nullptr;
    std::string header;
    std::getline(ist, header);
    chomp(header);
    std::vector<std::string> species = split(header, separator);
    if (species.size() < 2) throw std::runtime_error("Failed to identify species for gene families");
    species.erase(species.begin(), species.begin() + 2);  // description and ID columns
    pCafeFamily pcf = cafe_family_init(species);
    std::string line;
    while (std::getline(ist, line)) {
        chomp(line);
        std::vector<std::string> cols = split(line, separator);
        gene_family gf;
        if (cols.size() < 2) throw std::runtime_error("Error reading family '" + gf.id + "'");
        gf.desc = cols[0];
        gf.id = cols[1];
        try {
            for (size_t i = 2; i < cols.size(); ++i) gf.values.push_back(std::stoi(cols[i]));
        } catch (const std::invalid_argument&) {
            delete pcf;
            throw std::runtime_error("Error reading family '" + gf.id + "'");
        }
        if (gf.values.empty()) continue;
        if (max_size < 0 || *std::max_element(gf.values.begin(), gf.values.end()) <= max_size) cafe_family_add_item(pcf, gf);
    }
    __cafe_famliy_check_the_pattern(pcf);
    return pcf;
}

void __cafe_famliy_check_the_pattern(pCafeFamily pcf) {
    std::unordered_map<std::vector<int>, int, CountsHash> first_seen;
    first_seen.reserve(pcf->flist.size() * 2);
    for (size_t i = 0; i < pcf->flist.size(); ++i) {
        CafeFamilyItem& it = pcf->flist[i];
        if (it.ref != -1) continue;  // already classified by an earlier call, like the reference
        auto ins = first_seen.emplace(it.count, (int)i);
        it.ref = ins.first->second;
        it.holder = ins.second ? 1 : 0;
    }
}

void cafe_family_set_species_index(pCafeFamily pcf, pCafeTree pcafe) {
    std::map<std::string, int> leaf_names;
    for (int j = 0; j < pcafe->num_nodes(); j += 2) leaf_names[lower(pcafe->nlist[j].name)] = j;
    std::set<std::string> all_species;
    for (int i = 0; i < pcf->num_species; ++i) {
        std::string sp = lower(pcf->species[i]);
        all_species.insert(sp);
        if (!pcf->species[i].empty() && pcf->species[i][0] == '-') {
            pcf->index[i] = std::atoi(pcf->species[i].c_str() + 1);
        } else {
            auto it = leaf_names.find(sp);
            if (it == leaf_names.end()) throw std::runtime_error("No species '" + sp + "' was found in the tree");
            pcf->index[i] = it->second;
        }
    }
    for (auto& leaf : leaf_names)
        if (!all_species.count(leaf.first)) throw std::runtime_error("No species '" + leaf.first + "' was found in the family list");
}

void init_family_size(family_size_range* fs, int max) {
    fs->root_min = 1;  // must be 1, not 0
    fs->root_max = (int)std::max(30.0, std::rint(max * 1.25));
    fs->max = max + std::max(50, max / 5);
    fs->min = 0;
}

void cafe_family_set_size(pCafeFamily pcf, pCafeFamilyItem pitem, pCafeTree pcafe) {
    for (auto& n : pcafe->nlist) n.familysize = -1;
    for (int i = 0; i < pcf->num_species; ++i) {
        int idx = pcf->index[i];
        if (idx < 0 || idx >= pcafe->num_nodes()) {
            std::cerr << "Inconsistency in tree size";
            throw std::runtime_error("Inconsistency in tree size");  // the reference exit(-1)s here
        }
        pcafe->nlist[idx].familysize = pitem->count[i];
    }
}

void cafe_family_set_size_with_family_forced(pCafeFamily pcf, int idx, pCafeTree pcafe) {
    pCafeFamilyItem pitem = &pcf->flist[idx];
    cafe_family_set_size(pcf, pitem, pcafe);
    int max = 0;
    for (int i = 0; i < pcf->num_species; ++i) {
        if (pcf->index[i] < 0) continue;
        max = std::max(max, pitem->count[i]);
    }
    pcafe->range.root_min = 1;
    pcafe->range.root_max = (int)std::rint(max * 1.25);
    pcafe->range.max = max + std::max(50, max / 5);
    pcafe->rfsize = pcafe->range.root_max - pcafe->range.root_min + 1;
}

void cafe_family_reset_maxlh(pCafeFamily pcf) {
    for (auto& it : pcf->flist) it.maxlh = -1;
}

// ------------------------------------------------------------------------------------------ error model

std::istream& operator>>(std::istream& ifst, ErrorStruct& em) {
    std::string line;
    if (!std::getline(ifst, line)) throw std::runtime_error("Empty file");
    chomp(line);
    {   // "maxcnt:K"
        std::vector<std::string> first = split(split(line, ' ').at(0), ':');
        int file_rows = first.size() > 1 ? std::atoi(first[1].c_str()) : 0;
        em.maxfamilysize = std::max(em.maxfamilysize, file_rows);
    }
    if (std::getline(ifst, line)) {  // "cntdiff a ... b"
        chomp(line);
        std::vector<std::string> d = split(line, ' ');
        em.fromdiff = std::atoi(d.at(1).c_str());
        em.todiff = std::atoi(d.at(d.size() - 1).c_str());
    }
    const int N = em.maxfamilysize;
    em.errormatrix.assign((size_t)(N + 1) * (N + 1), 0.0);
    const int width = em.todiff - em.fromdiff + 1;
    auto copy_down = [&](int j) {  // a missing true-size row repeats the previous one, shifted by one
        for (int i = em.fromdiff; i <= em.todiff; ++i)
            if (i + j >= 0 && i + j <= N && i + j - 1 >= 0 && j - 1 >= 0) em.at(i + j, j) = em.at(i + j - 1, j - 1);
    };
    int j = 0;
    while (std::getline(ifst, line)) {
        chomp(line);
        std::vector<std::string> d = split(line, ' ');
        if ((int)d.size() != width + 1) continue;
        int col1 = std::atoi(d[0].c_str());
        while (j && j < col1) { copy_down(j); ++j; }
        for (int i = em.fromdiff, k = 1; i <= em.todiff; ++i, ++k)
            if (i + j >= 0 && i + j <= N) em.at(i + j, j) = std::atof(d[k].c_str());
        ++j;
    }
    while (j && j <= N) { copy_down(j); ++j; }
    return ifst;
}

int __check_error_model_columnsums(pErrorStruct em) {
    const int N = em->maxfamilysize, diff = em->todiff;
    auto colsum = [&](int j) { double s = 0; for (int i = 0; i <= N; ++i) s += em->at(i, j); return s; };
    for (int j = 0; j < diff; ++j) em->at(0, j) = em->at(0, j) + (1 - colsum(j));
    for (int j = diff; j <= N - diff; ++j) {
        double s = colsum(j);
        // the reference calls the INTEGER abs() here (cafe_shell.c:603, a C file): |1-s| is truncated to
        // int first, so the renormalisation only fires when the column is off by a whole unit.
        if (std::abs((int)(1 - s)) > 0.00000000000001)
            for (int i = 0; i <= N; ++i) em->at(i, j) = em->at(i, j) / s;
    }
    for (int j = N - diff + 1; j <= N; ++j) {
        if (j < 0) continue;
        em->at(N, j) = em->at(N, j) + (1 - colsum(j));
    }
    return 0;
}

static bool iequal(const std::string& a, const std::string& b) { return lower(a) == lower(b); }

int set_error_matrix_from_file(pCafeFamily family, pCafeTree pTree, family_size_range& range, std::string filename,
                               std::string speciesname) {
    int which = -1;
    for (size_t i = 0; i < family->errors.size(); ++i)
        if (iequal(family->errors[i].errorfilename, filename)) { which = (int)i; break; }
    if (which < 0) {
        ErrorStruct em;
        em.errorfilename = filename;
        em.maxfamilysize = range.max;
        std::ifstream ifst(filename.c_str());
        if (!ifst) throw std::runtime_error("ERROR(errormodel): Cannot open " + filename + " in read mode.\n");
        ifst >> em;
        __check_error_model_columnsums(&em);
        family->errors.push_back(std::move(em));
        which = (int)family->errors.size() - 1;
    }
    // init_error_ptr, error_model.cpp:206-229
    for (int i = 0; i < family->num_species; ++i) {
        if (!speciesname.empty() && !iequal(family->species[i], speciesname)) continue;
        family->error_ptr[i] = which;
        if (family->index[i] >= 0) pTree->nlist[family->index[i]].errormodel = which;
        if (!speciesname.empty()) break;
    }
    return 0;
}

int remove_error_model(pCafeFamily family, pCafeTree pcafe, std::string species_name) {
    for (int i = 0; i < family->num_species; ++i) {
        if (!species_name.empty() && !iequal(family->species[i], species_name)) continue;
        family->error_ptr[i] = -1;
        if (family->index[i] >= 0) pcafe->nlist[family->index[i]].errormodel = -1;
    }
    return 0;
}
