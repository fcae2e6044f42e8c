// Standalone timing harness for the 128x128 leaf (per-phase clock64 stamps).  nvcc -DLF_TIMING ...
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ long long lf_stamps[64];
#include "../hdbo_benchmark_b200/csrc/leaf.cuh"
using namespace hdbo;
int main() {
    const int n = 128;
    std::vector<double> A(n * n);
    const char* mode = getenv("LEAF_MATRIX");
    double w = (mode && mode[0] == 's') ? 1e-5 : 0.05;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i * n + j] = exp(-w * (i - j) * (i - j)) + (i == j ? 0.1 : 0.0);
    double *dA, *dL, *dLi, *dLt; int* info;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dLi, n * n * 8); cudaMalloc(&dLt, n * n * 8); cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    cudaFuncSetAttribute(k_leaf_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM_BYTES);
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_leaf_blocked<<<1, LF_THREADS, LF_SMEM_BYTES>>>(dA, dL, dLi, dLt, n, 0, 0, info);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long st[64]; cudaMemcpyFromSymbol(st, lf_stamps, sizeof(st));
        printf("rep %d: %.1f us total; err=%s\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        const char* names[] = {"load", "chol0", "trsm+inv0", "upd0", "chol1", "trsm+inv1", "upd1", "chol2", "trsm+inv2", "upd2", "chol3", "trsm+inv3", "upd3", "Lout+diag", "assemble", "write"};
        for (int i = 0; i < 16; i++) printf("  %-10s %8lld clk\n", names[i], st[i + 1] - st[i]);
    }
    std::vector<double> L(n * n), Li(n * n);
    cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Li.data(), dLi, n * n * 8, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) { double s = 0; for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k]; e1 = fmax(e1, fabs(s - A[i * n + j])); }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += L[i * n + k] * Li[k * n + j]; e2 = fmax(e2, fabs(s - (i == j))); }
    printf("|LL^T-A| %.2e  |L Linv - I| %.2e\n", e1, e2);
    // CPU inverse by forward substitution, per 32x32 block max abs error of the GPU inverse
    std::vector<double> Xi(n * n, 0.0);
    for (int c = 0; c < n; c++) for (int i = c; i < n; i++) { double s = (i == c); for (int k = c; k < i; k++) s -= L[i * n + k] * Xi[k * n + c]; Xi[i * n + c] = s / L[i * n + i]; }
    for (int ib = 0; ib < 4; ib++) { for (int jb = 0; jb < 4; jb++) { double e = 0, m = 0; for (int i = 0; i < 32; i++) for (int j = 0; j < 32; j++) { e = fmax(e, fabs(Li[(ib*32+i)*n + jb*32+j] - Xi[(ib*32+i)*n + jb*32+j])); m = fmax(m, fabs(Xi[(ib*32+i)*n + jb*32+j])); } printf("  blk(%d,%d) err %.1e max %.1e |", ib, jb, e, m); } printf("\n"); }
    return 0;
}
