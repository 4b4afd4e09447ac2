// fminsearch.h — the reference's Nelder–Mead, behaviour for behaviour (libcommon/fminsearch.cpp,
// libcommon/mathfunc.h:26-50).  The search driver is the CALLER of the GPU objective; lambda-hat
// parity to 1e-6 relative is only reachable when the host walks the same simplex path, including
// the reference's non-textbook accept rules and its tie order among equal (e.g. +inf) values
// (SURVEY.md App. C).
#pragma once
#include <vector>

typedef double (*math_func)(double* x, void* args);

struct FMinSearch {
    int maxiters = 10000;
    int bymax = 0;
    double rho = 1, chi = 2, psi = 0.5, sigma = 0.5;  // reflection, expansion, contraction, shrink
    double tolx = 1e-6, tolf = 1e-6;
    double delta = 0.05, zero_delta = 0.00025;
    int N = 0, N1 = 0;
    int iters = 0;
    std::vector<std::vector<double>> v;  // N1 vertices
    std::vector<double> fv_store;
    double* fv = nullptr;                // -> fv_store (the reference exposes *pfm->fv as the best value)
    std::vector<double> x_mean, x_r, x_tmp;
    void* args = nullptr;
    math_func eq = nullptr;
};
typedef FMinSearch* pFMinSearch;

pFMinSearch fminsearch_new();
pFMinSearch fminsearch_new_with_eq(math_func eq, int Xsize, void* args);
void fminsearch_set_equation(pFMinSearch pfm, math_func eq, int Xsize, void* args);
void fminsearch_free(pFMinSearch pfm);
int fminsearch_min(pFMinSearch pfm, double* X0);
double* fminsearch_get_minX(pFMinSearch pfm);
double fminsearch_get_minF(pFMinSearch pfm);
