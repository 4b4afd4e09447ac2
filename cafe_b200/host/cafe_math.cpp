#include "cafe_math.h"

#include <cmath>
#include <cstdlib>

namespace cafe {

namespace {
// Lanczos series coefficients (g = 5, n = 6), libcommon/mathfunc.c:87-89
constexpr double kQ[7] = {1.000000000190015,  76.18009172947146,  -86.50532032941677, 24.01409824083091,
                          -1.231739572450155, 1.208650973866179e-3, -5.395239384953e-6};
constexpr double kSqrt2Pi = 2.5066282746310002416123552393401042;
}  // namespace

double gammaln(double a) {
    const double shifted = a + 5.5;
    double series = kQ[0];
    for (int n = 1; n <= 6; ++n) series += kQ[n] / (a + n);
    return (a + 0.5) * std::log(shifted) - shifted + std::log(kSqrt2Pi * series / a);
}

double chooseln(double n, double r) {
    if (r == 0) return 0;                       // also covers n == 0 && r == 0
    if (n <= 0 || r <= 0) return std::log(0.0);  // -inf
    return gammaln(n + 1) - gammaln(r + 1) - gammaln(n - r + 1);
}

double poisspdf(int x, double lambda) { return std::exp(x * std::log(lambda) - gammaln(x + 1) - lambda); }

double unifrnd() { return std::rand() / (RAND_MAX + 1.0); }

double pvalue(double v, const double* cd, int size) {
    int lo = 0, hi = size - 1;
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (cd[mid] > v) {
            hi = mid - 1;
        } else if (cd[mid] < v) {
            lo = mid + 1;
        } else {  // widen to the whole run of ties
            lo = mid;
            while (lo > 0 && cd[lo - 1] == v) --lo;
            hi = mid;
            while (hi + 1 < size && cd[hi + 1] == v) ++hi;
            break;
        }
    }
    if (lo > hi) hi = lo;
    return (lo + (cd[lo] <= v ? 1 : 0) + (hi - lo) / 2.0) / size;
}

double chi2cdf(double x, int df) {
    const double shape = df / 2.0, scaled = x / 2;
    double sum = 1 / shape, term = 1 / shape;
    int i = 1;
    for (; i < 1000; ++i) {
        term *= scaled / (shape + i);
        if (term < 1e-8) break;  // __EPS__, libcommon/mathfunc.h:12
        sum += term;
    }
    const double log_lower = (i == 1000) ? gammaln(shape) : std::log(sum) + shape * std::log(scaled) - scaled;
    return std::exp(log_lower - gammaln(shape));
}

std::vector<double> lnc_table(int size) {
    const int rows = 2 * size, cols = size + 1;
    std::vector<double> T((size_t)rows * cols);
    for (int n = 0; n < rows; ++n)
        for (int x = 0; x < cols; ++x) T[(size_t)n * cols + x] = chooseln(n, x);
    return T;
}

}  // namespace cafe
