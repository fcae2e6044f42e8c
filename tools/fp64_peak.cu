// FP64 peak microbenchmark for B200 (sm_100a): measures the DMMA.8x8x4 and DFMA issue
// ceilings that every roofline fraction in this repo is quoted against ("of measured").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters) {
    double a = threadIdx.x * 1e-9, b = threadIdx.x * 2e-9;
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; j++) { c[j][0] = 0; c[j][1] = 0; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < NACC; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += c[j][0] + c[j][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-12, b = threadIdx.x * 2e-9;
    double c[NACC];
#pragma unroll
    for (int j = 0; j < NACC; j++) c[j] = j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) c[j] = fma(c[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += c[j];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 4096;
    int threads_list[] = {128, 256, 512, 1024};
    bool first = true;
    for (int t : threads_list) {
        {
            float ms = time_it([&] { dmma_kernel<8><<<sms, t>>>(out, iters); }, 5);
            double flops = 2.0 * 256 * 8 * (double)iters * (t / 32) * sms;
            printf("%s{\"op\": \"dmma884\", \"threads_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}", first ? "" : ",\n", t, ms, flops / ms / 1e9);
            first = false;
        }
        {
            float ms = time_it([&] { dfma_kernel<8><<<sms, t>>>(out, iters); }, 5);
            double flops = 2.0 * 32 * 8 * (double)iters * (t / 32) * sms;
            printf(",\n{\"op\": \"dfma\", \"threads_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}", t, ms, flops / ms / 1e9);
        }
    }
    // sustained (about 2 s) DMMA loop at 512 threads/SM, to see the power-capped rate
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int n = 0; CK(cudaEventRecord(e0));
        for (; n < 400; n++) dmma_kernel<8><<<sms, 512>>>(out, iters * 4);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 256 * 8 * (double)iters * 4 * 16 * sms * n;
        printf(",\n{\"op\": \"dmma884_sustained\", \"threads_per_sm\": 512, \"ms\": %.2f, \"tflops\": %.3f}", ms, flops / ms / 1e9);
    }
    printf("\n]}\n");
    return 0;
}
