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
#include <cstring>
#include <cmath>
#include "exp2_table.cuh"

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

// ---- the FP64-lean forms used by the int8 cross kernel -----------------------------------------------------------------------
// On B200 the scalar FP64 pipe does not run while tcgen05.mma is streaming on the same SM (tools/fp64_vs_tc.cu: a DFMA stream takes
// 4-14x longer next to a saturated int8 MMA stream, FFMA / integer work is unaffected), so in a kernel that feeds the tensor cores
// every DFMA is paid for in tensor time: these forms trade FP64 instructions for integer / table work.
__host__ __device__ __forceinline__ int fm_hi(double v) {
#ifdef __CUDA_ARCH__
    return __double2hiint(v);
#else
    long long b; memcpy(&b, &v, 8); return (int)(b >> 32);
#endif
}
__host__ __device__ __forceinline__ int fm_lo(double v) {
#ifdef __CUDA_ARCH__
    return __double2loint(v);
#else
    long long b; memcpy(&b, &v, 8); return (int)b;
#endif
}
__host__ __device__ __forceinline__ double fm_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    long long b = ((long long)hi << 32) | (unsigned)lo; double v; memcpy(&v, &b, 8); return v;
#endif
}

// sqrt(x), x in [1e-30, 1e300]: 22-bit reciprocal-square-root seed, then ONE cubically convergent correction
//   h = 1 - x y^2,  sqrt(x) = x y (1 - h)^(-1/2) = x y (1 + h/2 + 3 h^2/8 + O(h^3)),  |h| <~ 2^-21  ->  dropped terms < 2^-64
// (5 FP64 instructions; sqrt_pos above needs 8).
__host__ __device__ __forceinline__ double sqrt_pos_cubic(double x) {
    double y;
#ifdef __CUDA_ARCH__
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = (double)(1.0f / sqrtf((float)fm_make(fm_hi(x), 0)));     // host stand-in of the same accuracy class (tests)
#endif
    const double r0 = x * y;
    const double h = fma(-r0, y, 1.0);
    const double c = fma(h, 0.375, 0.5);
    return fma(r0, h * c, r0);
}

// exp(t), t < 0 (sign bit set; results below 2^-1021 flush to 0): n = round(64 t / ln 2) from the mantissa of 64 t log2(e) + 1.5 2^52,
// exp(t) = 2^(n >> 6) * 2^((n & 63) / 64) * exp(f), |f| <= ln2 / 128: table entry times a degree-5 polynomial (remainder < 4e-17).
// 10 FP64 instructions (exp_nonpos: 19 + two conversions).  `tab` = EXP2_TAB64 (shared-memory copy in kernels).
__host__ __device__ __forceinline__ double exp_neg_tab(double t, const double* tab) {
    if ((unsigned)fm_hi(t) > 0xC0872000u) t = -740.0;                   // integer compare: more negative than -740
    const double nm = fma(t, 9.23324826168936567683e+01, 6755399441055744.0);
    const int n = fm_lo(nm);                                            // in [-68352, 0]
    const double nf = nm - 6755399441055744.0;
    double f = fma(nf, -1.08304246932675596327e-02, t);                // ln2/64 split hi (32 significant bits: n hi exact) / lo
    f = fma(nf, -2.98158582698529328128e-12, f);
    double p = fma(f, 8.33333333333333321769e-03, 4.16666666666666643537e-02);
    p = fma(p, f, 1.66666666666666657415e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = tab[n & 63] * p;                                   // in [1, 2.02)
    const int e = n >> 6;
    return (e < -1021) ? 0.0 : fm_make(fm_hi(r) + e * 1048576, fm_lo(r));
}

// the same for t in [-707, -0] (callers clamp the argument): no range checks
__host__ __device__ __forceinline__ double exp_neg_tab_inrange(double t, const double* tab) {
    const double nm = fma(t, 9.23324826168936567683e+01, 6755399441055744.0);
    const int n = fm_lo(nm);
    const double nf = nm - 6755399441055744.0;
    double f = fma(nf, -1.08304246932675596327e-02, t);
    f = fma(nf, -2.98158582698529328128e-12, f);
    double p = fma(f, 8.33333333333333321769e-03, 4.16666666666666643537e-02);
    p = fma(p, f, 1.66666666666666657415e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = tab[n & 63] * p;
    return fm_make(fm_hi(r) + (n & ~63) * 16384, fm_lo(r));             // 2^(n >> 6): (n >> 6) << 20 added to the exponent field
}

}  // namespace hdbo
