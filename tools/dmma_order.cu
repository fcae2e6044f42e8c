// Register-only DMMA.8x8x4 microbenchmark with a GEMM-like operand pattern (8 A x 4 B fragments, 32 accumulator
// tiles per warp, 2 warps per scheduler) and several instruction orderings.  Answers: is the ~87 % DMMA-pipe occupancy of
// k_var a property of the operand/accumulator register pattern rather than of memory staging?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_order tools/dmma_order.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds64(const double* p) {
    double v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds128(const double* p) {
    double2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

// ORDER 0: mt-major (a reused 4x in a row)   1: nt-major (b reused 8x)   2: snake over nt within mt
// ORDER 3: 2x2 blocks (each a and b reused twice in a row)   MTv x NTv accumulator tiles
template <int MTv, int NTv, int ORDER>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, double seed) {
    double a[MTv], b[NTv], c[MTv][NTv][2];
#pragma unroll
    for (int i = 0; i < MTv; i++) a[i] = seed * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < NTv; j++) b[j] = seed * (threadIdx.x * 3 + j);
#pragma unroll
    for (int i = 0; i < MTv; i++)
#pragma unroll
        for (int j = 0; j < NTv; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
    for (int it = 0; it < iters; it++) {
        if (ORDER == 0) {
#pragma unroll
            for (int i = 0; i < MTv; i++)
#pragma unroll
                for (int j = 0; j < NTv; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        } else if (ORDER == 1) {
#pragma unroll
            for (int j = 0; j < NTv; j++)
#pragma unroll
                for (int i = 0; i < MTv; i++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        } else if (ORDER == 2) {
#pragma unroll
            for (int i = 0; i < MTv; i++)
#pragma unroll
                for (int jj = 0; jj < NTv; jj++) { const int j = (i & 1) ? NTv - 1 - jj : jj; dmma(c[i][j][0], c[i][j][1], a[i], b[j]); }
        } else {
#pragma unroll
            for (int i = 0; i < MTv; i += 2)
#pragma unroll
                for (int j = 0; j < NTv; j += 2) {
                    dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
                    dmma(c[i][j + 1][0], c[i][j + 1][1], a[i], b[j + 1]);
                    dmma(c[i + 1][j + 1][0], c[i + 1][j + 1][1], a[i + 1], b[j + 1]);
                    dmma(c[i + 1][j][0], c[i + 1][j][1], a[i + 1], b[j]);
                }
        }
        // perturb the operands a little so nothing is hoisted (cheap integer-free update every iteration)
        a[it & (MTv - 1)] += 1e-30;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MTv; i++)
#pragma unroll
        for (int j = 0; j < NTv; j++) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// DMMA fed from shared memory like the GEMM mainloop (no global traffic, no barriers):
// MODE 0: 12 LDS.64 per 32 DMMA (one k4 step per iteration)   MODE 1: 12 LDS.128 per 64 DMMA (two k4 steps per iteration,
// lane t holds k = 2t, 2t+1 so each step still covers 4 distinct k)   MODE 2: like 0 but operands loaded once (no LDS in loop)
template <int MODE>
__global__ void __launch_bounds__(256, 1) kl(double* out, int iters) {
    __shared__ __align__(16) double sm[2 * 128 * 24];
    for (int i = threadIdx.x; i < 2 * 128 * 24; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
    double c[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
    const int SLDv = (MODE == 1) ? 24 : 20;
    const double* as = sm + (wm * 64 + (lane >> 2)) * SLDv + (lane & 3) * (MODE == 1 ? 2 : 1);
    const double* bs = sm + 128 * 24 + (wn * 32 + (lane >> 2)) * SLDv + (lane & 3) * (MODE == 1 ? 2 : 1);
    double a0[8], b0[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a0[i] = as[i * 8 * SLDv];
#pragma unroll
    for (int j = 0; j < 4; j++) b0[j] = bs[j * 8 * SLDv];
    for (int it = 0; it < iters; it++) {
        const int off = (it & 1) * 8;   // alternate between two k positions so the loads are real
        if (MODE == 0) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = lds64(as + i * 8 * SLDv + off);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = lds64(bs + j * 8 * SLDv + off);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        } else if (MODE == 1) {
            double2 a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = lds128(as + i * 8 * SLDv + off);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = lds128(bs + j * 8 * SLDv + off);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(c[i][j][0], c[i][j][1], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(c[i][j][0], c[i][j][1], a[i].y, b[j].y);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(c[i][j][0], c[i][j][1], a0[i], b0[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// smem-fed DMMA loop (MODE 0 above) plus the GEMM's cp.async traffic: every 4 iterations (= one 16-wide k-block,
// 128 DMMA per warp) each thread issues NCP 16-byte cp.async into a 3-slot ring from an L2-resident buffer
// (NCP = 8 is the k_var ratio), commit + wait_group, optionally a block barrier.
template <int NCP, int BAR>
__global__ void __launch_bounds__(256, 1) kc(double* out, const double* __restrict__ src, int iters) {
    extern __shared__ __align__(16) double smd[];
    double* ring = smd;                       // 3 x 2 x 128 x 20 doubles
    double* sm = smd + 3 * 2 * 128 * 20;      // operand area actually read by the DMMA loop
    for (int i = threadIdx.x; i < 2 * 128 * 24; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
    double c[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
    const double* as = sm + (wm * 64 + (lane >> 2)) * 20 + (lane & 3);
    const double* bs = sm + 128 * 24 + (wn * 32 + (lane >> 2)) * 20 + (lane & 3);
    const double* g = src + (size_t)blockIdx.x * 256 * 4096;   // 256 rows x 4096 doubles per CTA
    int slot = 0;
    for (int it = 0; it < iters; it++) {
        if ((it & 3) == 0) {
            if (NCP > 0) {
                asm volatile("cp.async.wait_group 1;\n" ::);
                if (BAR) __syncthreads();
#pragma unroll
                for (int q = 0; q < NCP; q++) {
                    const int cidx = threadIdx.x + q * 256;
                    unsigned d = (unsigned)__cvta_generic_to_shared(ring + slot * 5120 + (cidx >> 3) * 20 + (cidx & 7) * 2);
                    const double* sp = g + (size_t)((it >> 2) & 255) * 16 + (size_t)(cidx >> 3) * 4096 + (cidx & 7) * 2;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(sp));
                }
                asm volatile("cp.async.commit_group;\n" ::);
                slot = slot == 2 ? 0 : slot + 1;
            } else if (BAR) __syncthreads();
        }
        const int off = (it & 1) * 8;
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = lds64(as + i * 8 * 20 + off);
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = lds64(bs + j * 8 * 20 + off);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

template <int NCP, int BAR>
void runc(const char* name, double* out, const double* src, int sms) {
    const int iters = 8192, smem = (3 * 2 * 128 * 20 + 2 * 128 * 24) * 8;
    CK(cudaFuncSetAttribute(kc<NCP, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        kc<NCP, BAR><<<sms, 256, smem>>>(out, src, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
    }
    double flops = 2.0 * 256 * 32 * (double)iters * 8 * sms;
    printf("{\"pattern\": \"%s\", \"tflops\": %.2f}\n", name, flops / best / 1e9);
}

template <int MODE>
void runl(const char* name, double* out, int sms) {
    const int iters = 8192;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        kl<MODE><<<sms, 256>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
    }
    double flops = 2.0 * 256 * 32 * (MODE == 1 ? 2 : 1) * (double)iters * 8 * sms;
    printf("{\"pattern\": \"%s\", \"tflops\": %.2f}\n", name, flops / best / 1e9);
}

template <int MTv, int NTv, int ORDER>
void run(const char* name, double* out, int sms, int threads) {
    const int iters = 4096;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        k<MTv, NTv, ORDER><<<sms, threads>>>(out, iters, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
    }
    double flops = 2.0 * 256 * MTv * NTv * (double)iters * (threads / 32) * sms;
    printf("{\"pattern\": \"%s\", \"acc_tiles\": %d, \"threads\": %d, \"tflops\": %.2f}\n", name, MTv * NTv, threads, flops / best / 1e9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; CK(cudaMalloc(&out, 8));
    const int sms = p.multiProcessorCount;
    run<8, 4, 0>("8x4 mt-major (k_var order)", out, sms, 256);
    run<8, 4, 1>("8x4 nt-major", out, sms, 256);
    run<8, 4, 2>("8x4 snake", out, sms, 256);
    run<8, 4, 3>("8x4 2x2 blocks", out, sms, 256);
    run<4, 4, 0>("4x4 mt-major", out, sms, 256);
    run<4, 8, 0>("4x8 mt-major", out, sms, 256);
    runl<2>("smem-fed: operands loaded once", out, sms);
    runl<0>("smem-fed: 12 LDS.64 per 32 DMMA (k_var ratio)", out, sms);
    runl<1>("smem-fed: 12 LDS.128 per 64 DMMA", out, sms);
    double* src; CK(cudaMalloc(&src, (size_t)sms * 256 * 4096 * 8)); CK(cudaMemset(src, 0, (size_t)sms * 256 * 4096 * 8));
    runc<0, 1>("smem-fed + barrier per k-block, no cp.async", out, src, sms);
    runc<8, 0>("smem-fed + 8 cp.async/thread per k-block (k_var ratio), no barrier", out, src, sms);
    runc<8, 1>("smem-fed + 8 cp.async/thread per k-block + barrier (= k_var mainloop shape)", out, src, sms);
    runc<4, 1>("smem-fed + 4 cp.async/thread per k-block + barrier", out, src, sms);
    runc<16, 1>("smem-fed + 16 cp.async/thread per k-block + barrier", out, src, sms);
    return 0;
}
