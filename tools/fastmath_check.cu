// Host-side accuracy check of the FP64-lean sqrt / exp forms of csrc/fast_math.cuh against libm (ulp statistics).
// nvcc -O2 -std=c++17 -o tools/fastmath_check tools/fastmath_check.cu && ./tools/fastmath_check
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../hdbo_benchmark_b200/csrc/fast_math.cuh"
int main() {
    const double tab[64] = {
#define X(v) v,
#include "exp2_table_values.inc"
    };
    std::mt19937_64 g(1);
    double worst_s = 0, worst_e = 0, worst_k = 0;
    for (int i = 0; i < 4000000; i++) {
        const double u = std::uniform_real_distribution<double>(-69.0, 69.0)(g);
        const double x = std::pow(10.0, u / 2.3) ;                     // 1e-30 .. 1e30 (the float stand-in seed of the host build)
        const double s = hdbo::sqrt_pos_cubic(x), sr = std::sqrt(x);
        worst_s = std::fmax(worst_s, std::fabs(s - sr) / (std::nextafter(sr, INFINITY) - sr));
        const double t = -std::pow(10.0, std::uniform_real_distribution<double>(-20.0, 2.87)(g));   // -1e-20 .. -741
        const double e = (t > -707.0) ? hdbo::exp_neg_tab_inrange(t, tab) : hdbo::exp_neg_tab(t, tab), er = std::exp(std::fmax(t, -740.0));
        if (t > -707.0 && e != hdbo::exp_neg_tab(t, tab)) { printf("mismatch at %g\n", t); return 1; }
        if (er > 1e-300) worst_e = std::fmax(worst_e, std::fabs(e - er) / (std::nextafter(er, INFINITY) - er));
        // Matern-5/2 as the cross kernel evaluates it: x = 5 r^2
        const double x5 = std::pow(10.0, std::uniform_real_distribution<double>(-29.0, 4.0)(g));
        const double a = hdbo::sqrt_pos_cubic(x5), k = fma(x5, 3.33333333333333314830e-01, 1.0 + a) * hdbo::exp_neg_tab(-a, tab);
        const long double al = sqrtl((long double)x5), kl = (1.0L + al + (long double)x5 / 3.0L) * expl(-al);
        worst_k = std::fmax(worst_k, (double)fabsl((long double)k - kl));   // absolute (k <= 1; the digits are absolute 2^-48)
    }
    printf("{\"sqrt_max_ulp\": %.3f, \"exp_max_ulp\": %.3f, \"matern_max_abs_err\": %.3e, \"two_pow_minus_53\": %.3e}\n", worst_s, worst_e, worst_k, ldexp(1.0, -53));
    const double edge[] = {-0.0, -1e-300, -1e-17, -0.0054, -0.00542, -708.0, -709.9, -739.9, -740.0, -745.0, -1e6};
    for (double t : edge) printf("  exp(%g) = %.17g  (libm %.17g)\n", t, hdbo::exp_neg_tab(t, tab), std::exp(t));
    return 0;
}
