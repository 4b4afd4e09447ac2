#include "fminsearch.h"

#include <cmath>
#include <numeric>

pFMinSearch fminsearch_new() { return new FMinSearch(); }

pFMinSearch fminsearch_new_with_eq(math_func eq, int Xsize, void* args) {
    pFMinSearch pfm = fminsearch_new();
    fminsearch_set_equation(pfm, eq, Xsize, args);
    return pfm;
}

void fminsearch_set_equation(pFMinSearch pfm, math_func eq, int Xsize, void* args) {
    pfm->N = Xsize;
    pfm->N1 = Xsize + 1;
    pfm->v.assign(pfm->N1, std::vector<double>(Xsize, 0.0));
    pfm->fv_store.assign(pfm->N1, 0.0);
    pfm->fv = pfm->fv_store.data();
    pfm->x_mean.assign(Xsize, 0.0);
    pfm->x_r.assign(Xsize, 0.0);
    pfm->x_tmp.assign(Xsize, 0.0);
    pfm->eq = eq;
    pfm->args = args;
}

void fminsearch_free(pFMinSearch pfm) { delete pfm; }

namespace {

// The reference orders vertices with its own in-place quicksort that drags an index array along
// (fminsearch.cpp:73-118).  Ties (several vertices at +inf are routine on the lambda*t = 1 wall) come
// out in an order that depends on that exact partition scheme, and that order decides which vertex
// is "worst" next — so the same scheme is restated here: first element is the pivot, the hole
// alternates between the two ends, `<=`/`>=` let equal keys stay where they are.
void sort_keys_with_index(double* key, int* idx, int lo, int hi) {
    const double pivot = key[lo];
    const int pivot_idx = idx[lo];
    int i = lo, j = hi;
    while (i < j) {
        while (pivot <= key[j] && i < j) --j;
        if (i != j) { key[i] = key[j]; idx[i] = idx[j]; ++i; }
        while (pivot >= key[i] && i < j) ++i;
        if (i != j) { key[j] = key[i]; idx[j] = idx[i]; --j; }
    }
    key[i] = pivot;
    idx[i] = pivot_idx;
    if (lo < i) sort_keys_with_index(key, idx, lo, i - 1);
    if (hi > i) sort_keys_with_index(key, idx, i + 1, hi);
}

void sort_simplex(pFMinSearch pfm) {
    std::vector<int> idx(pfm->N1);
    std::iota(idx.begin(), idx.end(), 0);
    sort_keys_with_index(pfm->fv, idx.data(), 0, pfm->N);
    std::vector<std::vector<double>> sorted(pfm->N1);
    for (int i = 0; i < pfm->N1; ++i) sorted[i] = pfm->v[idx[i]];
    pfm->v.swap(sorted);
}

bool vertices_converged(const FMinSearch& s) {  // adjacent vertices, fminsearch.cpp:121-136
    double worst = -1.7976931348623157e+308;
    for (int i = 0; i < s.N; ++i)
        for (int j = 0; j < s.N; ++j) worst = std::fmax(worst, std::fabs(s.v[i + 1][j] - s.v[i][j]));
    return worst <= s.tolx;
}

bool values_converged(const FMinSearch& s) {  // fminsearch.cpp:138-149
    double worst = -1.7976931348623157e+308;
    for (int i = 1; i < s.N1; ++i) {
        double t = std::fabs(s.fv[i] - s.fv[0]);
        if (t > worst) worst = t;  // NaN (inf-inf) never replaces, as in the reference's `t > max`
    }
    return worst <= s.tolf;
}

void replace_worst(pFMinSearch pfm, const std::vector<double>& x, double f) {
    pfm->v[pfm->N] = x;
    pfm->fv[pfm->N] = f;
    sort_simplex(pfm);
}

void shrink(pFMinSearch pfm) {
    for (int i = 1; i < pfm->N1; ++i) {
        for (int j = 0; j < pfm->N; ++j) pfm->v[i][j] = pfm->v[0][j] + pfm->sigma * (pfm->v[i][j] - pfm->v[0][j]);
        pfm->fv[i] = pfm->eq(pfm->v[i].data(), pfm->args);
    }
    sort_simplex(pfm);
}

void init_simplex(pFMinSearch pfm, const double* X0) {  // fminsearch.cpp:151-182
    for (int i = 0; i < pfm->N1; ++i) {
        // quirk: after an infinite previous vertex the step is 100x larger (only from the 3rd vertex on)
        const bool big = i > 1 && std::isinf(pfm->fv[i - 1]);
        const double grow = big ? 1 + pfm->delta * 100 : 1 + pfm->delta;
        for (int j = 0; j < pfm->N; ++j) {
            if (i - 1 == j) pfm->v[i][j] = X0[j] ? grow * X0[j] : pfm->zero_delta;
            else pfm->v[i][j] = X0[j];
        }
        pfm->fv[i] = pfm->eq(pfm->v[i].data(), pfm->args);
    }
    sort_simplex(pfm);
}

}  // namespace

int fminsearch_min(pFMinSearch pfm, double* X0) {
    init_simplex(pfm, X0);
    const int N = pfm->N;
    int it = 0;
    for (; it < pfm->maxiters; ++it) {
        if (vertices_converged(*pfm) && values_converged(*pfm)) break;
        for (int c = 0; c < N; ++c) {  // centroid of the N best
            double s = 0;
            for (int r = 0; r < N; ++r) s += pfm->v[r][c];
            pfm->x_mean[c] = s / N;
        }
        for (int c = 0; c < N; ++c) pfm->x_r[c] = pfm->x_mean[c] + pfm->rho * (pfm->x_mean[c] - pfm->v[N][c]);
        const double f_r = pfm->eq(pfm->x_r.data(), pfm->args);
        if (f_r < pfm->fv[0]) {
            for (int c = 0; c < N; ++c) pfm->x_tmp[c] = pfm->x_mean[c] + pfm->chi * (pfm->x_r[c] - pfm->x_mean[c]);
            const double f_e = pfm->eq(pfm->x_tmp.data(), pfm->args);
            if (f_e < f_r) replace_worst(pfm, pfm->x_tmp, f_e);
            else replace_worst(pfm, pfm->x_r, f_r);
        } else if (f_r >= pfm->fv[N]) {  // compared with the WORST vertex, not the second worst
            if (f_r > pfm->fv[N]) {
                for (int c = 0; c < N; ++c) pfm->x_tmp[c] = pfm->x_mean[c] + pfm->psi * (pfm->x_mean[c] - pfm->v[N][c]);
                const double f_cc = pfm->eq(pfm->x_tmp.data(), pfm->args);
                if (f_cc < pfm->fv[N]) replace_worst(pfm, pfm->x_tmp, f_cc);
                else shrink(pfm);
            } else {
                for (int c = 0; c < N; ++c) pfm->x_tmp[c] = pfm->x_mean[c] + pfm->psi * (pfm->x_r[c] - pfm->x_mean[c]);
                const double f_c = pfm->eq(pfm->x_tmp.data(), pfm->args);
                if (f_c <= f_r) replace_worst(pfm, pfm->x_tmp, f_c);
                else shrink(pfm);
            }
        } else {
            replace_worst(pfm, pfm->x_r, f_r);
        }
    }
    pfm->bymax = it == pfm->maxiters;
    pfm->iters = it;
    return pfm->bymax;
}

double* fminsearch_get_minX(pFMinSearch pfm) { return pfm->v[0].data(); }
double fminsearch_get_minF(pFMinSearch pfm) { return pfm->fv[0]; }
