// 128x128 diagonal-tile Cholesky + triangular inverse in ONE CTA (the latency-critical leaf of chol_inverse).
//
// Blocked 4x4 over 32x32 sub-blocks, everything resident in shared memory (row stride 132 doubles = 4 mod 16 so the
// 8-row x 4-k DMMA fragment reads are bank-conflict free):
//   * diagonal 32x32 block: ONE warp, the block lives in registers (lane i = row i, 32 doubles), pivots and columns are
//     exchanged with warp shuffles -- no shared-memory round trips, no block barriers on the 32-pivot critical path;
//   * its inverse: another warp, also in registers (lane c = column c of L^-1), overlapped with the panel solve;
//   * panel rows below: one thread per row, right-looking substitution in registers (L_kk read as smem broadcasts);
//   * trailing 32x32 block updates and the assembly of the tile inverse  X_ij = -(sum_k X_ik L_kj) X_jj : DMMA.8x8x4 on
//     fragments read straight from shared memory, one 16x32 half-block per warp.
// A non-positive pivot records info = 1-based global index (first one wins) and continues with a unit pivot.
#pragma once
#include "gemm_core.cuh"

#ifdef LF_TIMING   /* tools/leaf_bench.cu defines `__device__ long long lf_stamps[64]` before including this file */
#define LF_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) lf_stamps[i] = clock64(); } while (0)
#else
#define LF_STAMP(i) do { } while (0)
#endif

namespace hdbo {

constexpr int LF_THREADS = 256;
constexpr int LF_LD = 132;
constexpr int LF_SMEM_DOUBLES = 128 * LF_LD + 4 * 32 * 36 + 128;
constexpr int LF_SMEM_BYTES = LF_SMEM_DOUBLES * 8;
constexpr unsigned LF_FULL = 0xffffffffu;

// 1/sqrt(a) and sqrt(a) for a pivot on the 128-long critical path: MUFU.RSQ seed (2^-22) + two FP64 Newton steps (-> 2^-88
// before rounding) + one Heron correction for the square root; straight-line, ~12 dependent FP64 ops.  (With the pivot chain
// software-pipelined this beats the library rsqrt()/sqrt() pair; an earlier attempt inside the un-pipelined loop did not.)
__device__ __forceinline__ void lf_rsqrt_sqrt(double a, double& inv, double& d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));   // MUFU.RSQ64H on the high word: relative error ~2^-22, no F2F round trip
    double e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    const double r = a * y;
    inv = y;
    d = fma(fma(-r, r, a), 0.5 * y, r);
}

// lane i holds row i (lower part) of a 32x32 SPD block; on exit row i of its Cholesky factor.  invd[j] = 1/L_jj.
// Software-pipelined: the dependent chain of a pivot is  shuffle(pivot) -> rsqrt/sqrt -> mul -> shuffle(L_{j+1,j}) -> fma, and the
// NEXT pivot only needs column j+1 of the update.  So column j+1 is updated first, the next pivot is broadcast and its rsqrt/sqrt
// are issued immediately, and the remaining 30-j shuffle+fma pairs of the rank-1 update fill that latency (in program order the
// whole update loop sat in front of the next pivot's chain: 440 clk per pivot instead of ~250).
__device__ __forceinline__ int lf_warp_chol32(double (&a)[32], double* invd) {
    const int lane = threadIdx.x & 31;
    int bad = 0;
    double ajj = __shfl_sync(LF_FULL, a[0], 0);
    if (!(ajj > 0.0)) { bad = 1; ajj = 1.0; }
    double inv, d;
    lf_rsqrt_sqrt(ajj, inv, d);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const double l = (lane == j) ? d : a[j] * inv;
        a[j] = l;
        if (lane == 0) invd[j] = inv;
        if (j + 1 < 32) {
            // lane j+1 already owns L_{j+1,j} (its own l): the next pivot a_{j+1,j+1} - l^2 needs ONE shuffle, not two
            double nxt = __shfl_sync(LF_FULL, fma(-l, l, a[j + 1]), j + 1);
            if (!(nxt > 0.0)) { if (!bad) bad = j + 2; nxt = 1.0; }
            lf_rsqrt_sqrt(nxt, inv, d);
            const double l1 = __shfl_sync(LF_FULL, l, j + 1);
            a[j + 1] = fma(-l, l1, a[j + 1]);
        }
#pragma unroll
        for (int k = j + 2; k < 32; k++) {
            const double lk = __shfl_sync(LF_FULL, l, k);
            a[k] = fma(-l, lk, a[k]);
        }
    }
    return bad;
}

// lane r holds row r of L (32x32 lower); returns in x[] column `lane` of L^-1 (x[i] = Linv[i][lane]).
__device__ __forceinline__ void lf_warp_inv32(const double (&a)[32], const double* invd, double (&x)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const double xi = x[i] * invd[i];
        x[i] = xi;
#pragma unroll
        for (int r = i + 1; r < 32; r++) {
            const double lri = __shfl_sync(LF_FULL, a[i], r);
            x[r] = fma(-lri, xi, x[r]);
        }
    }
}

// acc (16 rows x 32 cols, rows r0.., as 2x4 DMMA tiles) = sum over kk in [0,K) of A(r, kk) * B(kk, c)
// fa(row, k) / fb(k, col) return shared-memory element addresses.
template <class FA, class FB>
__device__ __forceinline__ void lf_block_mma(double (&acc)[2][4][2], int K, FA fa, FB fb) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
#pragma unroll 4
    for (int k4 = 0; k4 < K; k4 += 4) {
        double a[2], b[4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) a[mt] = *fa(mt * 8 + (lane >> 2), k4 + (lane & 3));
#pragma unroll
        for (int nt = 0; nt < 4; nt++) b[nt] = *fb(k4 + (lane & 3), nt * 8 + (lane >> 2));
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
}

__global__ void __launch_bounds__(LF_THREADS, 1)
k_leaf_blocked(double* __restrict__ Awork, double* __restrict__ L, double* __restrict__ Linv, double* __restrict__ Linvt,
               int ld, long long bs, int tile, int32_t* __restrict__ info) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                          // [128][132]
    double* Dv = sm + 128 * LF_LD;           // [4][32][36] inverses of the diagonal blocks
    double* invd = Dv + 4 * 32 * 36;         // [128] 1 / L_jj
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)b * bs + (size_t)tile * 128 * ld + (size_t)tile * 128;
    const double* Ag = Awork + base;
    LF_STAMP(0);
    // whole tile with 16-byte async copies, 32 in flight per thread (the strict upper part is never consumed:
    // every store below masks it)
#pragma unroll 8
    for (int c = tid; c < 128 * 64; c += LF_THREADS) {
        const int i = c >> 6, ch = c & 63;
        cp_async16(S + i * LF_LD + ch * 2, Ag + (size_t)i * ld + ch * 2);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    LF_STAMP(1);

    // (Measured and rejected again with this structure: look-ahead -- warp 0 updating and factorising the next diagonal block while
    //  warps 1-6 finish the trailing update -- and a progressive write-out of L by warp 7: 55.6 us vs 51.2 us; the diagonal warp's
    //  dependent chain slows down when the other warps issue DMMA/LDS next to it.)
    // Warp 7 computes the inverses of the four diagonal blocks OFF the critical path: block kb's inverse is only needed by the
    // final assembly of the tile inverse, and it depends only on the finished diagonal factor (never touched again) and invd, so it
    // overlaps the panel solves, the trailing updates and the next diagonal factorisations of warps 0-6.  Warps 0-6 synchronise
    // among themselves on named barrier 1 (224 threads); warp 0 hands each finished diagonal block to warp 7 through barrier 2+kb.
    if (warp == 7) {
        for (int kb = 0; kb < 4; kb++) {
            const int o = kb * 32;
            asm volatile("bar.sync %0, 64;" ::"r"(2 + kb) : "memory");        // diagonal block kb is in shared memory
            double a[32], x[32];
#pragma unroll
            for (int k = 0; k < 32; k++) a[k] = S[(o + lane) * LF_LD + o + k];
            lf_warp_inv32(a, invd + o, x);
#pragma unroll
            for (int i = 0; i < 32; i++) Dv[(kb * 32 + i) * 36 + lane] = (i >= lane) ? x[i] : 0.0;
        }
    } else {
        for (int kb = 0; kb < 4; kb++) {
            const int o = kb * 32;
            // ---- diagonal block: warp 0, in registers ----
            if (warp == 0) {
                double a[32];
#pragma unroll
                for (int k = 0; k < 32; k++) a[k] = S[(o + lane) * LF_LD + o + k];
                const int bad = lf_warp_chol32(a, invd + o);
                if (bad && lane == 0) atomicCAS(info + b, 0, tile * 128 + o + bad);
#pragma unroll
                for (int k = 0; k < 32; k++)
                    if (k <= lane) S[(o + lane) * LF_LD + o + k] = a[k];
                __threadfence_block();
                asm volatile("bar.arrive %0, 64;" ::"r"(2 + kb) : "memory");  // -> warp 7
            }
            asm volatile("bar.sync 1, 224;" ::: "memory");
            LF_STAMP(2 + kb * 3);
            // ---- panel rows below the diagonal block: TWO threads per row (adjacent lanes; even / odd columns), right-looking
            //      substitution in registers; the pivot-column value hops to the partner lane with one shuffle per step.
            //      (One thread per row -- 528 dependent-ish FMAs + 528 shared-memory reads -- took as long as a 32x32 diagonal
            //      factorisation, 12 k clk per step of the 4-step leaf.)
            if (tid < 64 * (3 - kb)) {           // whole warps: 64 threads per 32 rows
                const int r = o + 32 + (tid >> 1), h = tid & 1;
                double acc[16];                  // columns h, h+2, ..., h+30
#pragma unroll
                for (int c = 0; c < 16; c++) acc[c] = S[r * LF_LD + o + 2 * c + h];
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    double xk = acc[k >> 1] * invd[o + k];                      // meaningful on the owner lane (h == k & 1)
                    xk = __shfl_sync(LF_FULL, xk, (lane & ~1) | (k & 1));       // owner of column k -> both lanes of the pair
                    if (h == (k & 1)) acc[k >> 1] = xk;
#pragma unroll
                    for (int c = (k >> 1); c < 16; c++) {
                        const int col = 2 * c + h;
                        if (col > k) acc[c] = fma(-xk, S[(o + col) * LF_LD + o + k], acc[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; c++) S[r * LF_LD + o + 2 * c + h] = acc[c];
            }
            asm volatile("bar.sync 1, 224;" ::: "memory");
            LF_STAMP(3 + kb * 3);
            // ---- trailing updates C_ib,jb -= L_ib,kb L_jb,kb^T, one 16x32 half-block per warp item (warps 0-6) ----
            {
                int item = 0;
                for (int ib = kb + 1; ib < 4; ib++)
                    for (int jb = kb + 1; jb <= ib; jb++)
                        for (int h = 0; h < 2; h++, item++) {
                            if ((item % 7) != warp) continue;
                            const int r0 = ib * 32 + h * 16, c0 = jb * 32;
                            double acc[2][4][2];
                            lf_block_mma(acc, 32,
                                         [&](int r, int k) { return S + (r0 + r) * LF_LD + o + k; },
                                         [&](int k, int c) { return S + (c0 + c) * LF_LD + o + k; });
#pragma unroll
                            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                                for (int nt = 0; nt < 4; nt++) {
                                    double* p = S + (r0 + mt * 8 + (lane >> 2)) * LF_LD + c0 + nt * 8 + (lane & 3) * 2;
                                    p[0] -= acc[mt][nt][0];
                                    p[1] -= acc[mt][nt][1];
                                }
                        }
            }
            asm volatile("bar.sync 1, 224;" ::: "memory");
            LF_STAMP(4 + kb * 3);
        }
    }
    __syncthreads();
    // ---- L tile out; then the diagonal blocks of S become their inverses ----
    double* Lg = L + base;
    for (int e = tid; e < 128 * 128; e += LF_THREADS) {
        const int i = e >> 7, k = e & 127;
        Lg[(size_t)i * ld + k] = (k <= i) ? S[i * LF_LD + k] : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < 4 * 32 * 32; e += LF_THREADS) {
        const int kb = e >> 10, i = (e >> 5) & 31, c = e & 31;
        S[(kb * 32 + i) * LF_LD + kb * 32 + c] = Dv[(kb * 32 + i) * 36 + c];
    }
    __syncthreads();
    LF_STAMP(14);
    // ---- tile inverse, block columns right to left: X_ib,jb = -(sum_{k=jb+1..ib} X_ib,k L_k,jb) X_jb,jb ----
    for (int jb = 2; jb >= 0; jb--) {
        const int nitems = (3 - jb) * 2;
        const int ib = jb + 1 + (warp >> 1), h = warp & 1;
        const bool active = warp < nitems;
        const int r0 = ib * 32 + h * 16, c0 = jb * 32;
        double acc[2][4][2];
        if (active)   // sum over the 32*(ib-jb) rows k of column block jb, starting at block row jb+1
            lf_block_mma(acc, 32 * (ib - jb),
                         [&](int r, int k) { return S + (r0 + r) * LF_LD + c0 + 32 + k; },
                         [&](int k, int c) { return S + (c0 + 32 + k) * LF_LD + c0 + c; });
        __syncthreads();
        if (active) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    double* p = S + (r0 + mt * 8 + (lane >> 2)) * LF_LD + c0 + nt * 8 + (lane & 3) * 2;
                    p[0] = acc[mt][nt][0];
                    p[1] = acc[mt][nt][1];
                }
        }
        __syncthreads();
        if (active)
            lf_block_mma(acc, 32,
                         [&](int r, int k) { return S + (r0 + r) * LF_LD + c0 + k; },
                         [&](int k, int c) { return S + (c0 + k) * LF_LD + c0 + c; });
        __syncthreads();
        if (active) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    double* p = S + (r0 + mt * 8 + (lane >> 2)) * LF_LD + c0 + nt * 8 + (lane & 3) * 2;
                    p[0] = -acc[mt][nt][0];
                    p[1] = -acc[mt][nt][1];
                }
        }
        __syncthreads();
    }
    LF_STAMP(15);
    double* Lig = Linv + base;
    double* Ltg = Linvt + base;
    // (a 32-byte-per-thread vectorised write-out with lanes along the row index was measured slower: the transposed tile
    //  then costs one 32-byte sector per thread on the store side)
    for (int e = tid; e < 128 * 128; e += LF_THREADS) {
        const int i = e >> 7, k = e & 127;
        Lig[(size_t)i * ld + k] = (k <= i) ? S[i * LF_LD + k] : 0.0;
        Ltg[(size_t)i * ld + k] = (k >= i) ? S[k * LF_LD + i] : 0.0;
    }
    LF_STAMP(16);
}

}  // namespace hdbo
