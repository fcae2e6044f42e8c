// Internal declarations shared by the .cu files of libhdbo_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hdbo_b200.h"
#include "fast_math.cuh"

#include <atomic>

namespace hdbo {

// The > 48 KB dynamic-shared-memory opt-in (cudaFuncSetAttribute) is a PER-DEVICE function attribute: a process that uses a second
// GPU after the first must set it again there.  One bit per device ordinal, checked against the calling thread's current device.
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool pending(int& dev) {
        dev = 0;
        cudaGetDevice(&dev);
        return dev >= 64 || !((done.load(std::memory_order_acquire) >> dev) & 1ull);
    }
    void mark(int dev) { if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release); }
};

constexpr double SQRT5 = 2.23606797749978969640917366873128;
constexpr double INV_SQRT_2 = 0.70710678118654752440084436210485;
constexpr double INV_SQRT_2PI = 0.39894228040143267793994605993438;
constexpr double LOG_2PI = 1.83787706640934548356065947281123;
constexpr double LOG_SQRT_PI_DIV_2 = 0.22579135264472743236309761494744;

// k-block range selectors for triangular operands (tile indices are relative to the sub-matrix handed in, which is
// always aligned on the diagonal of the triangular operand).
enum KRange : int {
    KR_FULL = 0,   // [0, K)
    KR_LE_N = 1,   // k-tile <= n-tile   (B lower triangular, row-major)
    KR_LE_M = 2,   // k-tile <= m-tile   (A lower triangular)
    KR_GE_N = 3,   // k-tile >= n-tile   (B upper triangular)
    KR_GE_M = 4,   // k-tile >= m-tile   (A upper triangular)
    KR_GE_MN = 5,  // k-tile >= max(m,n) (both upper)
    // the two triangular ranges with the long half of the columns cut in two at K/2 (C2 / ldc2 receive the second parts; beta = 0,
    // batch 1, even ntiles): 1.5 ntiles virtual columns of at most K/2 each, so that a launch with few row tiles is not as long as
    // its single longest k-range (the consumer adds C2 onto C)
    KR_LE_N_SPLIT = 6, KR_GE_N_SPLIT = 7,
};

struct GemmParams {
    const double* A; const double* B; double* C; double* Ct;  // C = alpha*A*B^T + beta*C ; Ct (optional) gets the transpose
    int lda, ldb, ldc, ldct;
    long long bsA, bsB, bsC, bsCt;                            // batch strides (elements)
    int mtiles, ntiles, kblocks;                              // 128-tiles in M and N; 16-blocks in K
    int krange;                                               // KRange
    int lower_only;                                           // only tiles with m-tile >= n-tile (grid.x enumerates them)
    double alpha, beta;
    int tile_hint;                                            // 0: choose the CTA tile by the launch's tile count; 32 / 64 / 128: force it
    double* C2; int ldc2, c2_col0;                            // KR_*_SPLIT: second-part output: element (i, j) of C lives at C2[i * ldc2 + j - c2_col0]
};

// kernel value k(r^2) and derivative dk/dr^2 for unit outputscale.  kind: 0 Matern-5/2, 1 RBF.
__device__ __forceinline__ void kernel_eval(int kind, double r2, double& k, double& dk) {
    // branch-free sqrt / exp (fast_math.cuh) so that the unrolled epilogues interleave their independent elements; the FP64-lean
    // forms (15 instead of 27 FP64 instructions, no conversion-pipe work): an epilogue runs after its CTA's DMMA mainloop, so its
    // instruction count is tile time (measured: MLL 5.60 -> 5.54 ms, FP64-mode pool +0.7 %)
    if (kind == 0) {
        const double r = sqrt_pos_cubic(fmin(fmax(r2, 1e-30), 1e30));
        const double e = exp_neg_tab(-SQRT5 * r, EXP2_TAB64);      // (64-entry table through L1)
        const double t = 1.0 + SQRT5 * r;
        k = (t + (5.0 / 3.0) * r * r) * e;
        dk = -(5.0 / 6.0) * t * e;
    } else {
        const double e = exp_neg_tab(-0.5 * r2 - 0.0, EXP2_TAB64);
        k = e;
        dk = -0.5 * e;
    }
}

// ---- launch wrappers (each enqueues on `stream` and returns cudaGetLastError()) ----
cudaError_t launch_gemm_nt(const GemmParams& p, int batch, cudaStream_t stream);

cudaError_t launch_scale_inputs(const double* X, int ldx, long long bsX, int nrows, int d, const double* center,
                                const double* inv_ls, long long bs_ls, double* Xs, int d_pad, int rows_pad,
                                long long bsXs, double* norms, long long bsn, int batch, cudaStream_t stream);

struct CrossParams {
    const double* Zs; const double* Xs; const double* zn; const double* xn; const double* alpha; const double* hyp;
    double* Kstar; double* DK; double* meanpart;
    int ldz, ldx, ldk;
    long long bsZ, bsX, bszn, bsxn, bsalpha, bsK, bsmp;
    int m, n, mtiles, ntiles, kblocks, kind;
};
cudaError_t launch_cross(const CrossParams& p, int batch, cudaStream_t stream);

struct SymBuildParams {
    const double* Xs; const double* xn; const double* hyp;
    double* Khat; double* R2;
    int ldx, ldk;
    long long bsX, bsxn, bsK;
    int n, tiles, kblocks, kind;
    int colmajor;                // enumerate the lower tiles column by column (column 0 first) instead of row by row
    int t0, t1;                  // tile range [t0, t1) of that enumeration (t1 = 0: all)
    int max_ctas;                // > 0: at most this many CTAs, each walking a strided list of tiles (128x128 tiles only)
};
cudaError_t launch_sym_build(const SymBuildParams& p, int batch, cudaStream_t stream);

struct VarParams {
    const double* Kstar; const double* Linv; double* V; double* varpart;
    int ldk, ldl, ldv;
    long long bsK, bsL, bsV, bsvp;
    int mtiles, ntiles;
};
cudaError_t launch_var(const VarParams& p, int batch, cudaStream_t stream);

// the rank-256 updates one pair of columns of the pipelined factorisation makes available, as a single persistent launch
// (k_side_updates, kernels_gemm.cu)
struct SideParams {
    double *Aw, *L, *Lt, *Kinv;
    int ld, T;
    int k0, kblocks;             // first column tile of the pair, contraction length in 16-blocks (16: a pair, 8: a single column)
    int n0, nn;                  // first tile column / row of the NEXT pair and its width (0, 1 or 2)
    int f;                       // first "far" tile column / row (= n0 + 2)
    int jmax;                    // last tile column of this pair (= n0 - 1): rows j <= jmax of L^-T exist
    double* Li;                  // L^-1 (region s writes its rows)
    int e_k0, e_jmax, e_kblocks; // the PREVIOUS pair's K^-1 region (deferred by one pair: filler without consumers before the end)
    int e2_k0, e2_kblocks;       // source columns of region e2 (the last pair but one merges its own and the previous pair's: K = 512)
    int w;                       // columns in this pair (1 or 2)
    // job counts of the regions, in list order.  Launch A (one round of the grid) holds only jobs whose operands exist as soon as the
    // pair's panel does -- a, s, c, e_prev, starting with the "head" (a, s and the first mb + 1 jobs of c: the next pair's columns,
    // this pair's rows of L^-1, the diagonal block of the pair after next); launch B = the rest, which reads those rows of L^-1:
    //   a  bulk, next pair's columns, far rows      A(i, j)    i >= f, n0 <= j < n0 + nn
    //   s  rows k0 .. k0+w-1 of L^-1                Linv(k0+r, j) = Linv(k0+r, k0..k0+r) W(k0..k0+r, j), j < k0  (+ L^-T(j, k0+r))
    //   c  bulk, lower triangle from f (column-major: (f, f), (f+1, f), ... so the next-but-one pair's diagonal block comes early)
    //   e  K^-1 of the previous pair                Kinv(a, b) b <= a <= e_jmax
    //   b  inverse, next pair's rows                W^T(j, i)  j <= jmax, n0 <= i < n0 + nn
    //   d  inverse, far rows                        W^T(j, i)  j <= jmax, i >= f
    //   e2 K^-1 of this pair (last two pairs only; the last but one folds the previous pair's columns in: the zero lower tiles of
    //      L^-T make that exact)                    Kinv(a, b) b <= a <= jmax
    int n_a, n_s, n_c, n_e, n_b, n_d, n_e2;
    int job0, job1;              // this launch walks jobs [job0, job1) of the pair's list
};
cudaError_t launch_side_updates(const SideParams& p, int max_ctas, cudaStream_t stream);

// caller-owned side streams / events of the pipelined factorisation (hdbo_gp.aux_streams / aux_events), see kernels_chol_pipe.cu
struct ChainStreams {
    cudaStream_t chain, bulk, inverse, panel;
    cudaEvent_t ev[12];
};
constexpr int PIPE_MIN_TILES = 8;   // below n_pad = 1024 the recursion's handful of launches wins
cudaError_t chol_inverse_pipelined(double* Aw, double* L, double* Li, double* Lt, double* Kinv, int ld, int T, int32_t* info,
                                   cudaStream_t origin, const ChainStreams& cs, const SymBuildParams* build_rest);
int pipe_build_head_tiles(int T);   // lower tiles (column-major) the caller must build before chol_inverse_pipelined: columns 0..3

// Cholesky factor + inverse of the padded SPD matrix (recursive, GEMM-rich); see kernels_chol.cu
cudaError_t chol_inverse(double* Awork, double* L, double* Linv, double* Linvt, int ld, long long bs, int tiles,
                         int n_true, int32_t* info, int batch, cudaStream_t stream);

}  // namespace hdbo
