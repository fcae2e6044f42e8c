// FP64 "NT" GEMM tile mainloop for sm_100a:  acc[TM x TN] += A[TM x K] * B[TN x K]^T
//
// Both operands are row-major with the contraction index contiguous, which is how every
// contraction on the GP hot path is laid out in HBM (scaled inputs [rows x d], K* [m x n],
// L^-1 [n x n], L^-T [n x n]); see DESIGN.md "Data layout".
//
// Hardware mapping (measured, profiles/r01_fp64_peak.json): tcgen05 has no f64 kind, the FP64
// tensor instruction on sm_100a is DMMA.8x8x4 (every mma.sync f64 shape lowers to it), peak
// 37.0 TFLOP/s = 64 FMA/clk/SM.  The default configuration G128 is one CTA = 256 threads = 8 warps in a
// 2 (M) x 4 (N) grid, each warp owning a 64x32 sub-tile = 8x4 DMMA tiles = 64 accumulator doubles per
// thread.  Operands are staged with cp.async (LDGSTS, 16 B) through a 3-stage ring of [rows x 16] slabs; the
// smem row stride is 20 doubles so that the 8 rows x 4 k fragment read of a warp hits 16 distinct 8-byte
// banks per half-warp (conflict-free).  Per k4 step a G128 warp issues 12 LDS.64 and 32 DMMA, so the DMMA
// pipe is the only saturated resource (smem traffic is 24 B/clk/SM of 128).
// G64 (64x64, 4 warps) and G32 (32x32, 4 warps) exist for the latency-bound nodes of the Cholesky /
// inverse recursion, where a launch has far fewer than 148 128-tiles and smaller tiles buy SM parallelism.
//
// All operand matrices handed to the mainloop are padded: row counts to multiples of 128, the
// contraction length to a multiple of 16, leading dimensions to multiples of 16 doubles (128 B), so
// the mainloop has no bounds checks and every cp.async is a full aligned 16-byte copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdbo {

constexpr int BK = 16;                       // k-block of the default configurations (and the unit of every kblocks argument)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int TM_, int TN_, int WM_, int WN_, int BK_ = 16, int STAGES_ = 3>
struct GemmT {
    static constexpr int TM = TM_, TN = TN_, WM = WM_, WN = WN_;
    static constexpr int BK = BK_, STAGES = STAGES_;
    static constexpr int SLD = BK + 4;           // smem row stride (doubles), = 4 mod 16: conflict-free fragment reads
    static constexpr int CH = BK / 2;            // 16-byte chunks per slab row
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int MT = TM / (8 * WM);     // DMMA tiles per warp along M
    static constexpr int NT = TN / (8 * WN);     // DMMA tiles per warp along N
    static constexpr int A_ELEMS = TM * SLD, B_ELEMS = TN * SLD;
    static constexpr int SMEM_BYTES = STAGES * (A_ELEMS + B_ELEMS) * 8;
    static constexpr int KB_PER_TM = TM / BK, KB_PER_TN = TN / BK;

    // accumulator fragment of one thread: v[mt][nt][e] = C[frag_row(mt)][frag_col(nt) + e]
    struct Frag {
        double v[MT][NT][2];
        __device__ __forceinline__ void zero() {
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) { v[i][j][0] = 0.0; v[i][j][1] = 0.0; }
        }
    };
    // thread index inside the GEMM's own group of THREADS threads: a CTA may hold several independent groups (warp groups of the
    // multi-pipeline side kernel, kernels_gemm.cu); THREADS is a power of two, so this is the identity for single-group CTAs
    static_assert((THREADS & (THREADS - 1)) == 0, "THREADS must be a power of two");
    static __device__ __forceinline__ int group_tid() { return threadIdx.x & (THREADS - 1); }
    static __device__ __forceinline__ int frag_row(int mt) {
        return ((group_tid() >> 5) / WN) * (TM / WM) + mt * 8 + ((threadIdx.x & 31) >> 2);
    }
    static __device__ __forceinline__ int frag_col(int nt) {
        return ((group_tid() >> 5) % WN) * (TN / WN) + nt * 8 + ((threadIdx.x & 3) << 1);
    }
    // barrier over the group: BAR = 0 -> the whole CTA (__syncthreads), BAR > 0 -> named barrier BAR over THREADS threads
    template <int BAR>
    static __device__ __forceinline__ void group_sync() {
        if (BAR == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(THREADS) : "memory");
    }

    template <int ROWS>
    static __device__ __forceinline__ void load_slab(double* s, const double* __restrict__ g, int ld, int kb, int tid) {
#pragma unroll
        for (int q = 0; q < ROWS * CH / THREADS; q++) {
            const int c = tid + q * THREADS, row = c / CH, ch = c % CH;
            cp_async16(s + row * SLD + ch * 2, g + (size_t)row * ld + (size_t)kb * BK + ch * 2);
        }
    }

    // gA / gB point at the first row of the operand slabs (column 0).  k-blocks [kb_begin, kb_end).
    template <int BAR = 0>
    static __device__ __forceinline__ void mainloop(const double* __restrict__ gA, int lda, const double* __restrict__ gB,
                                                    int ldb, int kb_begin, int kb_end, double* smem, Frag& acc) {
        const int tid = group_tid(), lane = tid & 31, warp = tid >> 5;
        const int wm = warp / WN, wn = warp % WN;
        double* sA = smem;
        double* sB = smem + STAGES * A_ELEMS;
        const int nkb = kb_end - kb_begin;
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if (s < nkb) {
                load_slab<TM>(sA + s * A_ELEMS, gA, lda, kb_begin + s, tid);
                load_slab<TN>(sB + s * B_ELEMS, gB, ldb, kb_begin + s, tid);
            }
            cp_async_commit();
        }
        const int frag_off_a = (wm * (TM / WM) + (lane >> 2)) * SLD + (lane & 3);
        const int frag_off_b = (wn * (TN / WN) + (lane >> 2)) * SLD + (lane & 3);
        int slot = 0;
        constexpr int KS = BK / 4;                       // k4 steps per k-block
        constexpr int CPT_A = TM * CH / THREADS;         // 16-byte chunks per thread per k-block, A and B
        constexpr int CPT_B = TN * CH / THREADS;
        static_assert(TM * CH % THREADS == 0 && TN * CH % THREADS == 0, "slab chunks must divide evenly over the threads");
        for (int it = 0; it < nkb; it++) {
            cp_async_wait<STAGES - 2>();
            group_sync<BAR>();
            // The cp.async for stage it+STAGES-1 are NOT issued here in one burst: measured (tools/dmma_order.cu), a burst of
            // 8 LDGSTS per thread right after the barrier costs 10 % of DMMA throughput (32.4 vs 36.0 TFLOP/s) because all
            // warps queue on the LSU while the tensor pipe drains.  They are spread over the k4 steps, behind the LDS.
            const int nx = it + STAGES - 1;
            int nslot = slot + STAGES - 1; if (nslot >= STAGES) nslot -= STAGES;
            const bool do_load = nx < nkb;
            double* nA = sA + nslot * A_ELEMS;
            double* nB = sB + nslot * B_ELEMS;
            const double* a_s = sA + slot * A_ELEMS + frag_off_a;
            const double* b_s = sB + slot * B_ELEMS + frag_off_b;
#pragma unroll
            for (int k4 = 0; k4 < KS; k4++) {
                double a[MT], b[NT];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) a[mt] = a_s[mt * 8 * SLD + k4 * 4];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) b[nt] = b_s[nt * 8 * SLD + k4 * 4];
                if (do_load) {
#pragma unroll
                    for (int q = 0; q < CPT_A; q++)
                        if (q * KS / CPT_A == k4) {
                            const int c = tid + q * THREADS, row = c / CH, ch = c % CH;
                            cp_async16(nA + row * SLD + ch * 2, gA + (size_t)row * lda + (size_t)(kb_begin + nx) * BK + ch * 2);
                        }
#pragma unroll
                    for (int q = 0; q < CPT_B; q++)
                        if (q * KS / CPT_B == k4) {
                            const int c = tid + q * THREADS, row = c / CH, ch = c % CH;
                            cp_async16(nB + row * SLD + ch * 2, gB + (size_t)row * ldb + (size_t)(kb_begin + nx) * BK + ch * 2);
                        }
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) dmma884(acc.v[mt][nt][0], acc.v[mt][nt][1], a[mt], b[nt]);
            }
            cp_async_commit();
            slot = (slot + 1 == STAGES) ? 0 : slot + 1;
        }
        cp_async_wait<0>();
        group_sync<BAR>();
    }
};

using G128 = GemmT<128, 128, 2, 4>;
// same warp grid / fragment layout as G128, 32-wide k-blocks (3 x 72 KB ring = 221 KB): half the barriers; measured best of the
// variants in tools/var_bench.cu (35.0 vs 34.5 TFLOP/s for BK = 16 inside k_var)
using G128W = GemmT<128, 128, 2, 4, 32, 3>;
using G64 = GemmT<64, 64, 2, 2>;
using G32 = GemmT<32, 32, 2, 2>;

// ---- the default 128x128 configuration under its historical names (used by the fused kernels) ----
constexpr int BM = 128;
constexpr int BN = 128;
constexpr int NTHREADS = G128::THREADS;
constexpr int GEMM_SMEM_BYTES = G128::SMEM_BYTES;   // 122880
constexpr int KB_PER_TILE = BM / BK;                // k-blocks per 128-wide tile (8)
using Frag = G128::Frag;

__device__ __forceinline__ int frag_row(int mt) { return G128::frag_row(mt); }
__device__ __forceinline__ int frag_col(int nt) { return G128::frag_col(nt); }
__device__ __forceinline__ void gemm_nt_mainloop(const double* __restrict__ gA, int lda, const double* __restrict__ gB,
                                                 int ldb, int kb_begin, int kb_end, double* smem, Frag& acc) {
    G128::mainloop(gA, lda, gB, ldb, kb_begin, kb_end, smem, acc);
}

// Deterministic per-row reduction of a per-thread value rowval[mt] (already summed over the thread's 8 columns):
// lanes sharing a row (lane&3) via shuffles, then the 4 N-warps through smem in fixed order.
// Result: out[r] for r in [0,128) valid in thread r (threads 0..127) after the call.
__device__ __forceinline__ double tile_row_reduce(double (&rowval)[8], double* red /* >= 4*128 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        double x = rowval[mt];
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        if ((lane & 3) == 0) red[wn * 128 + wm * 64 + mt * 8 + (lane >> 2)] = x;
    }
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 128) r = ((red[threadIdx.x] + red[128 + threadIdx.x]) + red[256 + threadIdx.x]) + red[384 + threadIdx.x];
    __syncthreads();
    return r;
}

// Same along columns: colval[nt][e] summed over the thread's 8 rows; lanes sharing a column (lane>>2) via shuffles,
// then the 2 M-warps through smem.  Result for column c valid in thread c (threads 0..127).
__device__ __forceinline__ double tile_col_reduce(double (&colval)[4][2], double* red /* >= 2*128 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            double x = colval[nt][e];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if ((lane >> 2) == 0) red[wm * 128 + wn * 32 + nt * 8 + (lane & 3) * 2 + e] = x;
        }
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 128) r = red[threadIdx.x] + red[128 + threadIdx.x];
    __syncthreads();
    return r;
}

}  // namespace hdbo
