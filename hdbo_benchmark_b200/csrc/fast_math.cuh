// Branch-free FP64 sqrt and exp for the distance -> kernel epilogues.
//
// The CUDA library sqrt()/exp() carry special-case branches; in the K* epilogue (64 kernel values per thread, 2 warps per
// scheduler) those branches cut the unrolled code into small blocks, the scheduler cannot interleave the independent
// elements, and every dependent FP64 op costs ~40 clk: measured, the epilogue of k_cross took as long as its 256-deep
// DMMA mainloop.  These versions are straight-line FMA sequences (no branches, no slow paths), accurate to ~1 ulp on
// the domain the kernels need, so 8 elements per fragment row pipeline through the FP64 unit.
//   sqrt_pos(x): x in [1e-30, 1e30]  (squared scaled distances, clamped below by the Matern 1e-30 floor)
//   exp_nonpos(t): t <= 0            (-sqrt5 r and -r^2/2); results below 2^-1021 flush to 0
#pragma once
#include <cuda_runtime.h>

namespace hdbo {

__device__ __forceinline__ double sqrt_pos(double x) {
    double y = (double)rsqrtf((float)x);          // MUFU.RSQ seed, relative error ~2^-22
    const double e = fma(-x * y, y, 1.0);         // one Newton step for 1/sqrt(x): y <- y + y e / 2   (error ~2^-43)
    y = fma(0.5 * y, e, y);
    const double r = x * y;                       // sqrt(x) estimate, then one Heron-style correction: error ~2^-86 before rounding
    return fma(fma(-r, r, x), 0.5 * y, r);
}

__device__ __forceinline__ double exp_nonpos(double t) {
    t = fmax(t, -740.0);
    const double nf = rint(t * 1.4426950408889634074);                  // n = round(t / ln 2)
    double f = fma(nf, -6.93147180369123816490e-01, t);                 // f = t - n ln2  (hi/lo split), |f| <= 0.3466
    f = fma(nf, -1.90821492927058770002e-10, f);
    double p = 1.6059043836821614599e-10;                               // Taylor of exp(f), degree 13: remainder < 5e-18
    p = fma(p, f, 2.0876756987868098979e-09);
    p = fma(p, f, 2.5052108385441718775e-08);
    p = fma(p, f, 2.7557319223985890653e-07);
    p = fma(p, f, 2.7557319223985892511e-06);
    p = fma(p, f, 2.4801587301587301566e-05);
    p = fma(p, f, 1.9841269841269841253e-04);
    p = fma(p, f, 1.3888888888888889419e-03);
    p = fma(p, f, 8.3333333333333332177e-03);
    p = fma(p, f, 4.1666666666666664354e-02);
    p = fma(p, f, 1.6666666666666665741e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const int n = (int)nf;                                              // in [-1068, 0]
    const long long bits = __double_as_longlong(p) + ((long long)n << 52);
    return (n < -1021) ? 0.0 : __longlong_as_double(bits);
}

}  // namespace hdbo
