// cafe_math.h — host-side scalar math of the likelihood path, kept bit-compatible with the reference
// (libcommon/mathfunc.c) because its outputs are INPUTS of the GPU path: the lnC table
// (cafe_gpu_set_lnc_table) and the root prior (cafe_gpu_set_prior).
#pragma once
#include <vector>

namespace cafe {

// 6-term Lanczos log-gamma — libcommon/mathfunc.c:87-89,112-119 (not libm lgamma: SURVEY.md fact 4)
double gammaln(double a);
// libcommon/mathfunc.c:224-229
double chooseln(double n, double r);
// libcommon/mathfunc.c:352-355
double poisspdf(int x, double lambda);
// libcommon/mathfunc.c:91-94 — glibc rand()/(RAND_MAX+1.0); the search start and the Poisson prior fit
// consume this stream, so the host keeps glibc's generator.
double unifrnd();
// libcommon/mathfunc.c:663-689
double pvalue(double v, const double* conddist, int size);

// chi2cdf -> gamcdf -> gammainc -> incgammaln_lower, libcommon/mathfunc.c:128-151,260-263,284-287: the series-only lower incomplete
// gamma (at most 999 terms, first term below 1e-8 stops it; a series that never stops yields exactly 1)
double chi2cdf(double x, int df);

// Dense lnC table for the GPU matrix builder: row-major [2*size][size+1], T[n][x] = chooseln(n, x)
// — the values libtree/chooseln_cache.h:27-41 memoises lazily.
std::vector<double> lnc_table(int size);

}  // namespace cafe
