// A/B harness for the variance contraction V = K* L^-T (k_var schedule: triangular k-range, heavy column tiles first,
// row sum-of-squares epilogue) over mainloop configurations of csrc/gemm_core.cuh.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/var_bench tools/var_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../hdbo_benchmark_b200/csrc/gemm_core.cuh"
using namespace hdbo;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <class G>
__global__ void __launch_bounds__(G::THREADS, 1) k_var_t(const double* __restrict__ Kstar, const double* __restrict__ Linv,
                                                         double* __restrict__ varpart, int ldk, int ldl, int mtiles, int ntiles) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.x, bn = ntiles - 1 - blockIdx.y;
    typename G::Frag acc; acc.zero();
    G::mainloop(Kstar + (size_t)bm * 128 * ldk, ldk, Linv + (size_t)bn * 128 * ldl, ldl, 0, (bn + 1) * (128 / G::BK), smem, acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp / G::WN, wn = warp % G::WN;
#pragma unroll
    for (int mt = 0; mt < G::MT; mt++) {
        double s = 0.0;
#pragma unroll
        for (int nt = 0; nt < G::NT; nt++) { s = fma(acc.v[mt][nt][0], acc.v[mt][nt][0], s); s = fma(acc.v[mt][nt][1], acc.v[mt][nt][1], s); }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((lane & 3) == 0) smem[wn * 128 + wm * (128 / G::WM) + mt * 8 + (lane >> 2)] = s;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        double r = 0.0;
        for (int w = 0; w < G::WN; w++) r += smem[w * 128 + threadIdx.x];
        varpart[(size_t)bn * (mtiles * 128) + bm * 128 + threadIdx.x] = r;
    }
}

template <class G>
void run(const char* name, const double* K, const double* L, double* vp, int rows, int n) {
    const int mt = rows / 128, nt = n / 128;
    CK(cudaFuncSetAttribute(k_var_t<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; r++) {
        CK(cudaEventRecord(e0));
        k_var_t<G><<<dim3(mt, nt), G::THREADS, G::SMEM_BYTES>>>(K, L, vp, n, n, mt, nt);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0 && ms < best) best = ms;
    }
    std::vector<double> h(16);
    CK(cudaMemcpy(h.data(), vp + (size_t)(nt - 1) * rows, 16 * 8, cudaMemcpyDeviceToHost));
    const double flops = (double)rows * n * n;   // algorithmic: n^2 per candidate
    printf("{\"variant\": \"%s\", \"threads\": %d, \"bk\": %d, \"stages\": %d, \"smem\": %d, \"ms\": %.3f, \"tflops_alg\": %.2f, \"check\": %.12e}\n", name,
           G::THREADS, G::BK, G::STAGES, G::SMEM_BYTES, best, flops / best / 1e9, h[3]);
}

__global__ void fill(double* p, size_t n, unsigned seed, double scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15; p[i] = scale * ((double)(x & 0xffffff) / 16777216.0 - 0.5); }
}

int main(int argc, char** argv) {
    const int n = 4096, rows = 148 * 128;
    double *K, *L, *vp;
    CK(cudaMalloc(&K, (size_t)rows * n * 8)); CK(cudaMalloc(&L, (size_t)n * n * 8)); CK(cudaMalloc(&vp, (size_t)(n / 128) * rows * 8));
    fill<<<(unsigned)(((size_t)rows * n + 255) / 256), 256>>>(K, (size_t)rows * n, 1u, 1.0);
    fill<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(L, (size_t)n * n, 7u, 0.1);
    CK(cudaDeviceSynchronize());
    run<GemmT<128, 128, 2, 4, 16, 3>>("base_256t_bk16_s3", K, L, vp, rows, n);
    run<GemmT<128, 128, 2, 4, 16, 4>>("256t_bk16_s4", K, L, vp, rows, n);
    run<GemmT<128, 128, 2, 4, 32, 3>>("256t_bk32_s3", K, L, vp, rows, n);
    run<GemmT<128, 128, 2, 4, 32, 2>>("256t_bk32_s2", K, L, vp, rows, n);
    run<GemmT<128, 128, 4, 4, 16, 3>>("512t_bk16_s3", K, L, vp, rows, n);
    run<GemmT<128, 128, 4, 4, 32, 3>>("512t_bk32_s3", K, L, vp, rows, n);
    run<GemmT<128, 128, 4, 2, 16, 3>>("256t_4x2_bk16_s3", K, L, vp, rows, n);
    run<GemmT<128, 128, 4, 4, 16, 4>>("512t_bk16_s4", K, L, vp, rows, n);
    return 0;
}
