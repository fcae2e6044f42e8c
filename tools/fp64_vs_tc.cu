// Does the scalar FP64 pipe run at full rate while tcgen05.mma kind::i8 is streaming on the same SM?
// One CTA per SM: thread 0 issues M128 N256 K32 int8 MMAs from resident shared-memory operands (as hdbo_i8_peak_probe does),
// warps 2..17 run independent DFMA (or FFMA / IMAD) chains.  Modes: 1 = MMA only, 2 = ALU only, 3 = both.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp64_vs_tc tools/fp64_vs_tc.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;              // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(256 >> 4) << 32;     // SBO: 8 rows x 32 B
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)6 << 61;              // SWIZZLE_32B
    return d;
}
__device__ __forceinline__ constexpr uint32_t idesc_i8(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_c), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// what: 0 DFMA, 1 FFMA, 2 IMAD
__global__ void __launch_bounds__(576, 1) k_mix(int mode, int what, int mma_iters, int alu_iters, int alu_warps, long long* t_mma, long long* t_alu, double* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(base + 16384);
    uint32_t* tmem_slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4096; i += blockDim.x) ((uint32_t*)base)[i] = (uint32_t)(i * 2654435761u) & 0x3f3f3f3fu;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0 && (mode & 1)) {
        const long long t0 = clock64();
        const uint64_t ad = make_desc_sw32(smem_u32(base)), bd = make_desc_sw32(smem_u32(base + 4096));
        for (int it = 0; it < mma_iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) umma_i8(tmem_base + (j & 1) * 256, ad, bd, idesc_i8(256), (uint32_t)(it != 0));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        if (blockIdx.x == 0) t_mma[0] = clock64() - t0;
    }
    if (warp >= 2 && warp < 2 + alu_warps && (mode & 2)) {
        const long long t0 = clock64();
        if (what == 0) {
            double a[8]; for (int j = 0; j < 8; j++) a[j] = 1.0 + tid * 1e-9 + j;
            const double m = 1.0000001, c = 1e-9;
            for (int it = 0; it < alu_iters; it++) {
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = fma(a[j], m, c);
            }
            double s = 0; for (int j = 0; j < 8; j++) s += a[j];
            if (s == 12345.678) sink[0] = s;
        } else if (what == 1) {
            float a[8]; for (int j = 0; j < 8; j++) a[j] = 1.0f + tid * 1e-6f + j;
            const float m = 1.00001f, c = 1e-6f;
            for (int it = 0; it < alu_iters; it++) {
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = fmaf(a[j], m, c);
            }
            float s = 0; for (int j = 0; j < 8; j++) s += a[j];
            if (s == 12345.678f) sink[0] = s;
        } else {
            unsigned a[8]; for (int j = 0; j < 8; j++) a[j] = tid + j;
            for (int it = 0; it < alu_iters; it++) {
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = a[j] * 1664525u + 1013904223u;
            }
            unsigned s = 0; for (int j = 0; j < 8; j++) s ^= a[j];
            if (s == 0x12345678u) sink[0] = s;
        }
        if (blockIdx.x == 0 && (tid & 31) == 0) t_alu[warp - 2] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
    constexpr int SM = 16384 + 64 + 1024;
    long long *t_mma, *t_alu; double* sink;
    CK(cudaMallocManaged(&t_mma, 8)); CK(cudaMallocManaged(&t_alu, 16 * 8)); CK(cudaMalloc(&sink, 8));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int mma_iters = 4000;                      // 32000 MMAs x 2.1 MOP
    const char* names[3] = {"dfma", "ffma", "imad"};
    for (int what = 0; what < 3; what++) {
        for (int warps : {4, 16}) {
            const int alu_iters = what == 0 ? 20000 : 80000;
            for (int mode = 1; mode <= 3; mode++) {
                if (mode == 1 && (what != 0 || warps != 4)) continue;
                float best = 1e30f; long long bm = 0, ba = 0;
                for (int rep = 0; rep < 3; rep++) {
                    t_mma[0] = 0; for (int i = 0; i < 16; i++) t_alu[i] = 0;
                    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    CK(cudaEventRecord(e0));
                    k_mix<<<sms, 576, SM>>>(mode, what, mma_iters, alu_iters, warps, t_mma, t_alu, sink);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) { best = ms; bm = t_mma[0]; ba = 0; for (int i = 0; i < warps; i++) if (t_alu[i] > ba) ba = t_alu[i]; }
                }
                const double mma_clk_per = bm ? (double)bm / (mma_iters * 8.0) : 0.0;            // clocks per M128 N256 K32 MMA
                const double alu_clk_per = ba ? (double)ba / (alu_iters * 8.0 * (warps / 4.0)) : 0.0;   // clocks per warp-instruction per scheduler
                printf("{\"alu\": \"%s\", \"alu_warps\": %d, \"mode\": %d, \"ms\": %.3f, \"mma_clk_per_mma\": %.1f, \"alu_clk_per_warp_instr_per_smsp\": %.2f}\n",
                       names[what], warps, mode, best, mma_clk_per, alu_clk_per);
            }
        }
    }
    return 0;
}
