// Scalar FP64 pipe throughput probe: independent DFMA chains per thread; prints lane-FMA per clock per SM for several
// warps-per-SM and ILP settings.  (Developer tool: sets the floor of the FP64 epilogues.)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* clk) {
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP> void run(int threads, double* out, long long* clk) {
    const int iters = 4096;
    k<ILP><<<148, threads>>>(out, iters, clk); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<ILP><<<148, threads>>>(out, iters, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; i++) cyc += h[i]; cyc /= 148;
    printf("threads/SM %4d ILP %2d: %.1f lane-FMA/clk/SM (clock64), %.2f TFLOP/s (events)\n", threads, ILP,
           (double)threads * ILP * iters / cyc, 2.0 * 148 * threads * ILP * iters / (ms * 1e-3) / 1e12);
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
    for (int th : {128, 256, 512, 1024}) { run<1>(th, out, clk); run<4>(th, out, clk); run<16>(th, out, clk); }
    return 0;
}
