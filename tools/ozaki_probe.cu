// Probe of the digit-sliced (Ozaki) int8 variance contraction on tcgen05: S digit planes per operand, all S+... weight groups
// (g = s + t) accumulated in TMEM at once, operands of a 64-byte k-block loaded once per stage by two 3-D TMA copies.
//   acc[g][M x N] (int32) = sum_{s+t=g} A_s[M x K] * B_t[N x K]^T ,   0-based s,t in [0,S), pairs kept while s + t <= S - 1
// Checked EXACTLY against the CPU.  Every mbarrier wait is bounded.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ozaki_probe tools/ozaki_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int S = 7;                     // digit planes per operand
constexpr int G = S;                     // weight groups g = s + t in [0, S-1]
constexpr int TM = 128, TN = 64, TKB = 64;
constexpr int STAGES = 2;
constexpr int A_STAGE = S * TM * TKB;    // 57344
constexpr int B_STAGE = S * TN * TKB;    // 28672
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) + 256 + 1024;
constexpr int TMEM_COLS = 512;

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major operand, SWIZZLE_64B (64-byte rows): 8-row groups 512 B apart (SBO = 32 x 16 B), LBO = 1, version 1
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)32 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// grid (M/128, N/64); 192 threads: warp 0 TMA, warp 1 MMA, warps 2..5 epilogue (warp w reads TMEM lanes 32*(w%4)..)
__global__ void __launch_bounds__(192, 1) k_ozaki(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  int32_t* __restrict__ C /*[G][M][N]*/, int M, int N, int kblocks, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* As = base;
    uint8_t* Bs = base + STAGES * A_STAGE;
    uint64_t* bars = (uint64_t*)(base + STAGES * (A_STAGE + B_STAGE));
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    if (warp == 0 && lane == 0) {            // ---- TMA producer: all S planes of both operands for one 64-byte k-block ----
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            if (!mbar_wait(empty0 + 8 * s, ph ^ 1, err)) break;
            mbar_expect_tx(full0 + 8 * s, A_STAGE + B_STAGE);
            tma_load_3d(smem_u32(As + s * A_STAGE), &tmA, full0 + 8 * s, kb * TKB, bm * TM, 0);
            tma_load_3d(smem_u32(Bs + s * B_STAGE), &tmB, full0 + 8 * s, kb * TKB, bn * TN, 0);
        }
    } else if (warp == 1 && lane == 0) {     // ---- MMA issuer: 28 digit pairs x 2 k-steps per k-block, grouped by weight ----
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            if (!mbar_wait(full0 + 8 * s, ph, err)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(As + s * A_STAGE), b0 = smem_u32(Bs + s * B_STAGE);
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int sa = 0; sa <= g; sa++) {
                    const int tb = g - sa;
                    const uint64_t ad = make_desc_sw64(a0 + sa * TM * TKB), bd = make_desc_sw64(b0 + tb * TN * TKB);
#pragma unroll
                    for (int k = 0; k < TKB / 32; k++)
                        umma_i8(tmem_base + g * TN, ad + 2 * k, bd + 2 * k, idesc, (kb | sa | k) != 0);
                }
            }
            umma_commit(empty0 + 8 * s);
        }
        umma_commit(accum);
    }
    __syncwarp();
    if (warp >= 2) {                         // ---- epilogue: raw int32 group accumulators to global ----
        mbar_wait(accum, 0, err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int row = bm * TM + q * 32 + lane;
        for (int g = 0; g < G; g++) {
#pragma unroll
            for (int ch = 0; ch < TN / 32; ch++) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + g * TN + ch * 32, r);
                int32_t* dst = C + ((size_t)g * M + row) * N + bn * TN + ch * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<int4*>(dst + j) = make_int4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// planes [S][rows][kbytes] -> 3-D map, box {64 bytes of K, box_rows, S planes}, 64B swizzle
static CUtensorMap make_map3(EncodeFn enc, void* ptr, uint64_t rows, uint64_t kbytes, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[3] = {kbytes, rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {kbytes, kbytes * rows};
    cuuint32_t box[3] = {TKB, box_rows, (cuuint32_t)S};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

static bool run(EncodeFn enc, int M, int N, int K, bool check) {
    std::vector<int8_t> A((size_t)S * M * K), B((size_t)S * N * K);
    for (size_t i = 0; i < A.size(); i++) A[i] = (int8_t)(((i * 1103515245u + 12345u) >> 16) % 129) - 64;
    for (size_t i = 0; i < B.size(); i++) B[i] = (int8_t)(((i * 214013u + 2531011u) >> 15) % 129) - 64;
    int8_t *dA, *dB; int32_t* dC; int* derr;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, (size_t)G * M * N * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(derr, 0, 4)); CK(cudaMemset(dC, 0xff, (size_t)G * M * N * 4));
    CUtensorMap ta = make_map3(enc, dA, M, K, TM), tb = make_map3(enc, dB, N, K, TN);
    CK(cudaFuncSetAttribute(k_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dim3 grid(M / TM, N / TN);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < (check ? 1 : 5); r++) {
        CK(cudaEventRecord(e0));
        k_ozaki<<<grid, 192, SMEM_BYTES>>>(ta, tb, dC, M, N, K / TKB, derr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    bool ok = herr == 0;
    const double pairs = S * (S + 1) / 2.0;
    if (check) {
        std::vector<int32_t> C((size_t)G * M * N);
        CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0, first = -1;
        for (int g = 0; g < G; g++) for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
            int64_t s = 0;
            for (int sa = 0; sa <= g; sa++) {
                const int8_t* a = &A[((size_t)sa * M + i) * K]; const int8_t* b = &B[((size_t)(g - sa) * N + j) * K];
                for (int k = 0; k < K; k++) s += (int32_t)a[k] * (int32_t)b[k];
            }
            if ((int32_t)s != C[((size_t)g * M + i) * N + j]) { if (first < 0) first = ((long)g * M + i) * N + j; bad++; }
        }
        printf("{\"check\": \"M=%d N=%d K=%d S=%d\", \"timeout_flag\": %d, \"mismatches\": %ld, \"first_bad\": %ld}\n", M, N, K, S, herr, bad, first);
        ok = ok && bad == 0;
    } else {
        printf("{\"bench\": \"M=%d N=%d K=%d S=%d pairs=%.0f\", \"timeout_flag\": %d, \"ms\": %.3f, \"int8_tops\": %.1f, \"fp64_equiv_tflops\": %.1f}\n", M, N, K, S,
               pairs, herr, best, 2.0 * M * N * (double)K * pairs / best / 1e9, 2.0 * M * N * (double)K / best / 1e9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(derr);
    return ok;
}

int main() {
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    if (!run(enc, 128, 64, 64, true)) return 2;
    if (!run(enc, 256, 192, 512, true)) return 3;
    run(enc, 18944, 4096, 4096, false);
    run(enc, 18944, 4096, 2048, false);
    return 0;
}
