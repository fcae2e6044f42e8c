/* libhdbo_b200_dev.so -- developer / measurement hooks, deliberately NOT part of the product library libhdbo_b200.so.
 *
 * Loaded only by tests (device math routines against libm) and by bench.py (live tensor-pipe ceilings = roofline denominators).
 * Same conventions as hdbo_b200.h: device pointers, caller's stream, 0 / >0 cudaError_t / <0 argument index.
 */
#ifndef HDBO_B200_DEV_H
#define HDBO_B200_DEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sqrt_out[i] = sqrt(clamp(x[i], 1e-30, 1e30)), exp_out[i] = exp(-|x[i]|) computed with the branch-free device routines the
 * kernel epilogues use (csrc/fast_math.cuh); either output may be NULL. */
int hdbo_test_math(const double* x, int64_t n, double* sqrt_out, double* exp_out, void* stream);
/* Same for the FP64-lean forms of the int8 cross kernel (cubic sqrt correction; table-based exp, |x| clamped to 700). */
int hdbo_test_math_lean(const double* x, int64_t n, double* sqrt_out, double* exp_out, void* stream);
/* Launches a register-only DMMA.8x8x4 loop on `sms` CTAs x 512 threads (scratch: >= 8 bytes of device memory) and
 * writes the flops it performs to *flops_out (HOST pointer).  Timed by the caller: measured FP64 tensor peak. */
int hdbo_fp64_peak_probe(double* scratch, int iters, int sms, double* flops_out, void* stream);

/* Same for the int8 tensor pipe: `sms` CTAs issue resident-operand tcgen05.mma kind::i8 (M128 N256 K32) only; *ops_out (HOST) =
 * integer operations performed (2 per multiply-accumulate).  flag: device int32, set if the completion wait timed out. */
int hdbo_i8_peak_probe(int32_t* flag, int iters, int sms, double* ops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HDBO_B200_DEV_H */
