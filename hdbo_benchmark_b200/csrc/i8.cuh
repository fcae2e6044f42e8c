// Digit slicing shared by the producers (k_cross epilogue, L^-1 slicing) and the consumer (k_var_i8) of the exact int8 path.
// A value x in [-1, 1] is held as a 49-bit fixed-point integer X = rint(x 2^48) and split into S = 7 balanced base-128
// digits d_t in [-64, 64]:  x = sum_t d_t 2^-(7t+6).  Products of two digit planes with s + t = g share the weight 2^-(7g+12).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdbo {

constexpr int I8_S = 7;
constexpr int I8_FRAC_BITS = 48;
constexpr int I8_A_IMAGE = I8_S * 128 * 32;   // bytes of one K* tile image (128 candidates x 32 k) -- layout in kernels_i8.cu
constexpr int I8_B_IMAGE = I8_S * 64 * 32;    // bytes of one L^-1 tile image (64 rows x 32 k)

// sum_{t=1..6} 64 128^(6-t): adding it to a NON-NEGATIVE fixed-point value makes its 6 low balanced digits plain 7-bit fields + 64
constexpr long long I8_LOW_BIAS = 0x20408102040LL;

__device__ __forceinline__ void i8_digits(long long x, signed char (&d)[I8_S]) {
#pragma unroll
    for (int t = I8_S - 1; t >= 1; t--) {
        const int r = (int)((x + 64) & 127) - 64;   // balanced remainder in [-64, 63]
        d[t] = (signed char)r;
        x = (x - r) >> 7;
    }
    d[0] = (signed char)x;                          // |x| <= 64 for |value| <= 1
}

__device__ __forceinline__ double i8_group_weight(int g) {   // 2^-(7g+12)
    return __longlong_as_double((long long)(1023 - (7 * g + 12)) << 52);
}

cudaError_t launch_slice_linv(const double* Linv, int ld, int n_pad, int8_t* planes, double* scale, cudaStream_t st);
cudaError_t launch_var_i8(const int8_t* Aimg, int rows_pad, const int8_t* Bimg, const double* colscale, const double* hyp,
                          int n_pad, double* varpart, int* flag, cudaStream_t st);

cudaError_t launch_slice_rows(int rows_per_tile, const double* X, long long ldx, int rows_valid, int rows_pad, int d, const double* center,
                              const double* inv_ls, int8_t* img, double* scale, double* norms, cudaStream_t st);
cudaError_t launch_cross_i8(int kind, const int8_t* Zimg, const int8_t* Ximg, const double* zscale, const double* xscale, const double* zn,
                            const double* xn, const double* alpha, const double* hyp, int8_t* Kimg, double* meanpart, int m, int n,
                            int rows_pad, int n_pad, int d, int* flag, cudaStream_t st);
cudaError_t launch_i8_peak(int iters, int sms, int* flag, double* ops_out, cudaStream_t st);

}  // namespace hdbo
