// Negative-binomial pmf for a large size parameter (K4, the E-step and estimate_threshold of the mixture fit).
//
// scipy.stats.nbinom.pmf(k, n, p) -- what lib/python/analysis/peaks.py:57,75 calls with n = 1/alpha, p = 1/(1+alpha*mu) --
// is accurate to a few ulp.  The textbook form exp(lgamma(n+k) - lgamma(k+1) - lgamma(n) + n log p + k log q) is not once
// n is large: with the dispersion at its floor (alpha = 1e-6, em_fitting.py:32,74) n = 1e6, lgamma(n+k) ~ 1.3e7 and its
// rounding alone is 1e-9 relative in the pmf, which the slowly contracting EM map carries into the 8th digit of the fitted
// parameters.  For n above NB_SADDLE_MIN the pmf is evaluated by the saddle-point expansion of C. Loader, "Fast and
// accurate computation of binomial probabilities" (2000): no large terms ever cancel.
//
//   NB(k; n, p) = n/(n+k) * Binom(n; n+k, p)
//   Binom(x; N, p) = exp( d(N) - d(x) - d(N-x) - D(x, Np) - D(N-x, Nq) ) / sqrt(2 pi x (N-x)/N)
//   d(z) = log z! - log( sqrt(2 pi z) (z/e)^z )          (Stirling error)
//   D(x, m) = x log(x/m) + m - x                          (deviance term, by series when x ~ m)
//
// Host + device so that tests/test_library.py can compare it with scipy on the host.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

constexpr double NB_SADDLE_MIN = 1000.0;   // 1/alpha above which the saddle-point form is used

// Stirling error d(z), z > 0
NB_HD double nb_stirling_error(double z) {
    if (z > 15.0) {
        // asymptotic series 1/(12 z) - 1/(360 z^3) + 1/(1260 z^5) - 1/(1680 z^7) + 1/(1188 z^9)
        const double r = 1.0 / z, r2 = r * r;
        return r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0 - r2 * (1.0 / 1680.0 - r2 * (1.0 / 1188.0)))));
    }
    return lgamma(z + 1.0) - (z + 0.5) * log(z) + z - 0.91893853320467274178;   // - log sqrt(2 pi)
}

// D(x, m) = x log(x/m) + m - x for x, m > 0
NB_HD double nb_deviance(double x, double m) {
    const double d = x - m, sum = x + m;
    if (fabs(d) < 0.1 * sum) {
        // log(x/m) = 2 atanh(t), t = d/(x+m)  =>  D = d t + 2 x t * sum_{j>=1} t^(2j)/(2j+1);  t^2 < 0.01: 9 terms
        const double t = d / sum, u = t * t;
        double h = 1.0 / 19.0;
        h = fma(h, u, 1.0 / 17.0); h = fma(h, u, 1.0 / 15.0); h = fma(h, u, 1.0 / 13.0); h = fma(h, u, 1.0 / 11.0);
        h = fma(h, u, 1.0 / 9.0); h = fma(h, u, 1.0 / 7.0); h = fma(h, u, 1.0 / 5.0); h = fma(h, u, 1.0 / 3.0);
        return fma(2.0 * x * t, h * u, d * t);
    }
    return x * log(x / m) + m - x;
}

// pmf of the negative binomial with real size n > 0 and success probability p at the integer k >= 0; q = 1 - p is
// passed separately so that the caller can form it without cancellation (alpha mu / (1 + alpha mu))
NB_HD double nb_pmf_saddle(double k, double n, double p, double q) {
    if (k <= 0.0) return exp(n * log(p));
    const double N = n + k;
    const double lc = nb_stirling_error(N) - nb_stirling_error(n) - nb_stirling_error(k) - nb_deviance(n, N * p) - nb_deviance(k, N * q);
    const double lf = log(6.283185307179586477 * n * (k / N));
    return (n / N) * exp(lc - 0.5 * lf);
}
