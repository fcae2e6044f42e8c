// Probe: int8 x int8 -> int32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8, operands by TMA with 128B swizzle,
// accumulators in TMEM), the building block of an exact (Ozaki-sliced) replacement for the FP64 DMMA variance GEMM.
//   C[M x N] (int32) = A[M x K] (int8, K contiguous) * B[N x K]^T (int8, K contiguous)
// Every mbarrier wait is bounded (error flag instead of a hang).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/i8gemm_probe tools/i8gemm_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TM = 128, TN = 128, TKB = 128;       // tile rows / cols, K bytes per stage (= one 128B swizzle atom)
constexpr int STAGES = 4;
constexpr int SLAB = TM * TKB;                      // 16 KB per operand per stage
constexpr int SMEM_BYTES = 2 * STAGES * SLAB + 256 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    for (long i = 0; i < 40000000L; i++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major operand, SWIZZLE_128B: 8-row groups 1024 B apart (SBO = 64 x 16 B), LBO = 1, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_i8gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                   int32_t* __restrict__ C, int ldc, int kblocks, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* As = base;
    uint8_t* Bs = base + STAGES * SLAB;
    uint64_t* bars = (uint64_t*)(base + 2 * STAGES * SLAB);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum = smem_u32(bars + 2 * STAGES);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bm = blockIdx.x, bn = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // M = 128, N = 128, A/B signed 8-bit K-major, C = S32
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    if (warp == 0 && lane == 0) {            // ---- TMA producer ----
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            if (!mbar_wait(empty0 + 8 * s, ph ^ 1, err)) break;
            mbar_expect_tx(full0 + 8 * s, 2 * SLAB);
            tma_load_2d(smem_u32(As + s * SLAB), &tmA, full0 + 8 * s, kb * TKB, bm * TM);
            tma_load_2d(smem_u32(Bs + s * SLAB), &tmB, full0 + 8 * s, kb * TKB, bn * TN);
        }
    } else if (warp == 1 && lane == 0) {     // ---- MMA issuer ----
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            if (!mbar_wait(full0 + 8 * s, ph, err)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = make_desc(smem_u32(As + s * SLAB)), bd = make_desc(smem_u32(Bs + s * SLAB));
#pragma unroll
            for (int k = 0; k < TKB / 32; k++) umma_i8(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
            umma_commit(empty0 + 8 * s);
        }
        umma_commit(accum);
    }
    __syncwarp();
    // ---- epilogue: TMEM -> registers -> global (thread = one row of the tile) ----
    mbar_wait(accum, 0, err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int32_t* crow = C + (size_t)(bm * TM + warp * 32 + lane) * ldc + bn * TN;
#pragma unroll
    for (int ch = 0; ch < TN / 32; ch++) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                     "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<int4*>(crow + ch * 32 + j) = make_int4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* ptr, uint64_t rows, uint64_t kbytes) {
    CUtensorMap m;
    cuuint64_t dims[2] = {kbytes, rows};
    cuuint64_t strides[1] = {kbytes};
    cuuint32_t box[2] = {TKB, TM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

static bool run(EncodeFn enc, int M, int N, int K, bool check) {
    std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
    for (size_t i = 0; i < A.size(); i++) A[i] = (int8_t)((i * 1103515245u + 12345u) >> 16) % 65 - 32;
    for (size_t i = 0; i < B.size(); i++) B[i] = (int8_t)((i * 214013u + 2531011u) >> 15) % 61 - 30;
    int8_t *dA, *dB; int32_t* dC; int* derr;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(derr, 0, 4)); CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    CUtensorMap ta = make_map(enc, dA, M, K), tb = make_map(enc, dB, N, K);
    CK(cudaFuncSetAttribute(k_i8gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dim3 grid(M / TM, N / TN);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < (check ? 1 : 5); r++) {
        CK(cudaEventRecord(e0));
        k_i8gemm<<<grid, 128, SMEM_BYTES>>>(ta, tb, dC, N, K / TKB, derr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    bool ok = herr == 0;
    if (check) {
        std::vector<int32_t> C((size_t)M * N);
        CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0; long first = -1;
        for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
            int32_t s = 0; for (int k = 0; k < K; k++) s += (int32_t)A[(size_t)i * K + k] * (int32_t)B[(size_t)j * K + k];
            if (s != C[(size_t)i * N + j]) { if (first < 0) first = (long)i * N + j; bad++; }
        }
        printf("{\"check\": \"M=%d N=%d K=%d\", \"timeout_flag\": %d, \"mismatches\": %ld, \"first_bad\": %ld, \"c00\": %d}\n", M, N, K, herr, bad, first, C[0]);
        ok = ok && bad == 0;
    } else {
        printf("{\"bench\": \"M=%d N=%d K=%d\", \"timeout_flag\": %d, \"ms\": %.3f, \"tops\": %.1f}\n", M, N, K, herr, best, 2.0 * M * N * (double)K / best / 1e9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(derr);
    return ok;
}

int main() {
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    if (!run(enc, 128, 128, 128, true)) return 2;
    if (!run(enc, 256, 384, 1024, true)) return 3;
    run(enc, 8192, 8192, 8192, false);
    run(enc, 18944, 4096, 4096, false);
    return 0;
}
