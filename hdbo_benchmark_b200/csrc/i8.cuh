// Digit slicing shared by the producers and consumers of the exact int8 (tcgen05) path -- kernels_i8.cu.
//
// A value is held as a 48-bit fixed-point integer and split into S = 6 base-256 digits, one int8 plane per digit:
//   signed   (L^-1 rows, scaled inputs; |x| < 1 after a power-of-two row scale):  X = rint(x 2^47) = sum_t b_t 256^(5-t),
//            balanced digits b_t in [-128, 127]  ->  x = sum_t b_t 2^-(8t+7)
//   unsigned (K* / outputscale in [0, 1]):                      X = min(rint(x 2^48), 2^48 - 1) = sum_s a_s 256^(5-s),
//            plain bytes a_s in [0, 255]        ->  x = sum_s a_s 2^-(8s+8)
// Products of two digit planes with s + t = g share one weight; only pairs with g <= S - 1 are formed (21 of 36): the dropped
// ones weigh <= 2^-48 of the row scale, the level of the fixed-point rounding itself.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdbo {

constexpr int I8_S = 6;
constexpr int I8_SIGNED_FRAC_BITS = 47;
constexpr int I8_A_IMAGE = I8_S * 128 * 32;   // bytes of one 128-row tile image (128 rows x 32 k) -- layout in kernels_i8.cu
constexpr int I8_B_IMAGE = I8_S * 64 * 32;    // bytes of one 64-row tile image
constexpr int I8_MAX_N_PAD = 8192;            // unsigned x signed group sums stay below 2^31: 6 pairs x n x 255 x 128
constexpr int I8_MAX_D = 16384;               // signed x signed (cross GEMM): 6 pairs x d x 128 x 128 < 2^31

// Balanced base-256 digits without a carry chain: adding 128 to every digit turns them into the plain bytes of X + bias, and
// b_t = byte_t - 128 = byte_t ^ 0x80 as int8.  Representable range: |X| <= 0x7F7F7F7F7F7F (0.996 2^47) -- see i8_row_exponent.
constexpr long long I8_SIGNED_BIAS = 0x808080808080LL;

// exponent e with |x| 2^-e <= 0.99 for every |x| <= mx, so that rint(x 2^(47-e)) has 6 balanced digits
__device__ __forceinline__ int i8_row_exponent(double mx) {
    int ex = 0;
    if (mx > 0.0) { const double f = frexp(mx, &ex); if (f > 0.99) ex++; }   // mx = f 2^ex, f in [0.5, 1)
    return ex;
}

// the 6 balanced digits of X (|X| within range), most significant first, as raw bytes (two's-complement int8)
__device__ __forceinline__ void i8_signed_digits(long long X, unsigned (&d)[I8_S]) {
    const unsigned long long y = (unsigned long long)(X + I8_SIGNED_BIAS) ^ 0x808080808080ULL;
    const unsigned lo = (unsigned)y, hi = (unsigned)(y >> 32);
    d[5] = lo & 255u; d[4] = (lo >> 8) & 255u; d[3] = (lo >> 16) & 255u; d[2] = lo >> 24; d[1] = hi & 255u; d[0] = (hi >> 8) & 255u;
}

__device__ __forceinline__ double i8_pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// batch = number of GPs that share inputs but differ in hyper-parameters (SAAS samples); per-GP data is contiguous per GP
cudaError_t launch_slice_linv(const double* Linv, int ld, int n_pad, int8_t* img, double* scale, int batch, cudaStream_t st);
cudaError_t launch_slice_rows(int rows_per_tile, const double* X, long long ldx, int rows_valid, int rows_pad, int d, const double* center,
                              const double* inv_ls, int8_t* img, double* scale, double* norms, long long bs_rows, int batch,
                              cudaStream_t st);
cudaError_t launch_cross_i8(int kind, const int8_t* Zimg, const int8_t* Ximg, const double* zscale, const double* xscale, const double* zn,
                            const double* xn, const double* alpha, const double* hyp, int8_t* Kimg, double* meanpart, int m, int n,
                            int rows_pad, int rows_alloc, int n_pad, int d, int batch, int* flag, cudaStream_t st);
cudaError_t launch_var_i8(const int8_t* Aimg, int rows_pad, int rows_alloc, const int8_t* Bimg, const double* colscale, const double* hyp,
                          int n_pad, double* varpart, int batch, int* flag, cudaStream_t st);

}  // namespace hdbo
