// extern "C" entry points of libhdbo_b200_dev.so (include/hdbo_b200_dev.h): developer / measurement hooks kept OUT of the product
// library -- the device math routines on arrays (tests compare them with libm) and the tensor-pipe issue-ceiling probes that
// bench.py uses as roofline denominators.  Self-contained: links against cudart only.
#include "gemm_core.cuh"
#include "fast_math.cuh"
#include "tc05.cuh"
#include "../../include/hdbo_b200_dev.h"

#define TRY_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return (int)_e; } while (0)

namespace hdbo {

// test hook: the branch-free sqrt / exp of fast_math.cuh evaluated on arrays (tests compare with the library functions)
__global__ void k_test_math(const double* __restrict__ x, long long n, double* __restrict__ s, double* __restrict__ e) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (s) s[i] = sqrt_pos(fmin(fmax(x[i], 1e-30), 1e30));
    if (e) e[i] = exp_nonpos(-fabs(x[i]));
}
// the FP64-lean forms of the int8 cross kernel: sqrt (clamped to [1e-30, 1e300]) and exp(-|x|) with |x| clamped to 700
__global__ void k_test_math_lean(const double* __restrict__ x, long long n, double* __restrict__ s, double* __restrict__ e) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = EXP2_TAB64[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (s) s[i] = sqrt_pos_cubic(fmin(fmax(x[i], 1e-30), 1e300));
    if (e) e[i] = exp_neg_tab_inrange(-fmin(fabs(x[i]), 700.0) - 0.0, tab);
}
static cudaError_t launch_test_math_lean(const double* x, long long n, double* s, double* e, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_test_math_lean<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, s, e);
    return cudaGetLastError();
}
static cudaError_t launch_test_math(const double* x, long long n, double* s, double* e, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_test_math<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, s, e);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------
// int8 tensor-pipe ceiling, measured live (bench.py roofline denominator): every CTA issues `iters` x 8 resident-operand MMAs
// (M = 128, N = 256, K = 32, pseudo-random digits in shared memory, no global traffic) and waits for the last commit.
__global__ void __launch_bounds__(128, 1) k_i8_peak(int iters, int* flag) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(base + 16384);
    uint32_t* tmem_slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4096; i += 128) ((uint32_t*)base)[i] = (uint32_t)(i * 2654435761u) & 0x3f3f3f3fu;   // digits in [0, 63]
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes above -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        const uint64_t ad = make_desc_sw32(smem_u32(base)), bd = make_desc_sw32(smem_u32(base + 4096));
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) umma_i8(tmem_base + (j & 1) * 256, ad, bd, idesc_i8(256, true), (uint32_t)(it != 0));
        }
        umma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0, flag);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

static cudaError_t launch_i8_peak(int iters, int sms, int* flag, double* ops_out, cudaStream_t st) {
    constexpr int SM = 16384 + 64 + 1024;
    k_i8_peak<<<sms, 128, SM, st>>>(iters, flag);
    if (ops_out) *ops_out = 2.0 * 128 * 256 * 32 * 8.0 * iters * sms;
    return cudaGetLastError();
}



// ---- FP64 tensor (DMMA.8x8x4) issue-ceiling probe: the denominator of every FP64 roofline fraction ----
__global__ void __launch_bounds__(512) k_dmma_peak(double* out, int iters) {
    double a = threadIdx.x * 1e-9, b = threadIdx.x * 2e-9;
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) { c[j][0] = 0; c[j][1] = 0; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace hdbo

using namespace hdbo;

extern "C" int hdbo_test_math(const double* x, int64_t n, double* sqrt_out, double* exp_out, void* stream_) {
    if (!x) return -1;
    if (n < 0) return -2;
    TRY_CUDA(launch_test_math(x, n, sqrt_out, exp_out, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_test_math_lean(const double* x, int64_t n, double* sqrt_out, double* exp_out, void* stream_) {
    if (!x) return -1;
    if (n < 0) return -2;
    TRY_CUDA(launch_test_math_lean(x, n, sqrt_out, exp_out, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_fp64_peak_probe(double* scratch, int iters, int sms, double* flops_out, void* stream_) {
    if (!scratch) return -1;
    if (iters <= 0 || sms <= 0) return -2;
    k_dmma_peak<<<sms, 512, 0, (cudaStream_t)stream_>>>(scratch, iters);
    TRY_CUDA(cudaGetLastError());
    if (flops_out) *flops_out = 2.0 * 256 * 8 * (double)iters * 16 * sms;   /* host pointer */
    return 0;
}

extern "C" int hdbo_i8_peak_probe(int32_t* flag, int iters, int sms, double* ops_out, void* stream_) {
    if (!flag || iters <= 0 || sms <= 0) return -1;
    TRY_CUDA(launch_i8_peak(iters, sms, flag, ops_out, (cudaStream_t)stream_));
    return 0;
}
