// FP64 Cholesky factor AND its inverse for the padded SPD matrix Khat = K + noise I  (SURVEY.md §8a row A2).
//
// Recursive block algorithm over 128-tiles [a, b), split at h:
//     chol_inv([a,h))                                  -> L11, L11^-1 (and its transpose)
//     L21    = A21 * L11^-T                            GEMM, B lower-triangular  (k-tile <= n-tile)
//     A22   -= L21 * L21^T                             SYRK, lower tiles only
//     chol_inv([h,b))                                  -> L22, L22^-1
//     T^T    = L11^-T(as U11) * L21^T                  GEMM, A upper-triangular  (k-tile >= m-tile), scratch = upper-right of Awork
//     L21inv = -L22^-1 * T                             GEMM, A lower-triangular  (k-tile <= m-tile), also stored transposed
// so every flop above the 128x128 leaves runs in the DMMA mainloop with large K (n^3/3 for the factor, n^3/3 for
// the inverse).  Having L^-1 explicitly turns every later triangular solve on the hot path (posterior variance,
// alpha, K^-1 for the MLL gradient) into a GEMM / matvec -- the same trade BoTorch's fast_pred_var cache makes.
//
// The leaf (leaf.cuh) factors AND inverts one 128x128 diagonal tile in a single CTA: 32x32 diagonal blocks in
// registers with warp shuffles, panel rows by per-thread substitution, block updates with DMMA from shared memory.
// A non-positive pivot records info = 1-based global index (first one wins) and continues with a unit pivot so the
// stream never hangs; the host applies the jitter schedule.
#include <cstdlib>

#include "gemm_core.cuh"
#include "kernels.cuh"
#include "leaf.cuh"

namespace hdbo {

cudaError_t launch_leaf(double* Awork, double* L, double* Linv, double* Linvt, int ld, long long bs, int tile,
                        int32_t* info, int batch, cudaStream_t stream) {
    // A warp-specialised, look-ahead ("pipelined") variant of the leaf was built and measured slower (82-104 us vs 72 us):
    // with one CTA per SM and each role running its own long unrolled code path the warps thrash the instruction cache.
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_leaf_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    k_leaf_blocked<<<batch, LF_THREADS, LF_SMEM_BYTES, stream>>>(Awork, L, Linv, Linvt, ld, bs, tile, info);
    return cudaGetLastError();
}

#define HDBO_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

static cudaError_t rec(double* Aw, double* L, double* Li, double* Lt, int ld, long long bs, int a, int b, int32_t* info,
                       int batch, cudaStream_t stream) {
    if (b - a == 1) return launch_leaf(Aw, L, Li, Lt, ld, bs, a, info, batch, stream);
    const int h = a + (b - a + 1) / 2;
    HDBO_TRY(rec(Aw, L, Li, Lt, ld, bs, a, h, info, batch, stream));
    auto at = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    GemmParams g{};
    g.lda = g.ldb = g.ldc = g.ldct = ld;
    g.bsA = g.bsB = g.bsC = g.bsCt = bs;
    // L21 = A21 * Linv11^T
    g.A = at(Aw, h, a); g.B = at(Li, a, a); g.C = at(L, h, a); g.Ct = nullptr;
    g.mtiles = b - h; g.ntiles = h - a; g.kblocks = (h - a) * KB_PER_TILE; g.krange = KR_LE_N; g.lower_only = 0;
    g.alpha = 1.0; g.beta = 0.0;
    HDBO_TRY(launch_gemm_nt(g, batch, stream));
    // A22 -= L21 L21^T (lower tiles)
    g.A = at(L, h, a); g.B = at(L, h, a); g.C = at(Aw, h, h);
    g.mtiles = g.ntiles = b - h; g.krange = KR_FULL; g.lower_only = 1; g.alpha = -1.0; g.beta = 1.0;
    HDBO_TRY(launch_gemm_nt(g, batch, stream));
    HDBO_TRY(rec(Aw, L, Li, Lt, ld, bs, h, b, info, batch, stream));
    // T^T[j,i] = sum_{k>=j} U11[j,k] L21[i,k]   -> scratch in the (unused) upper-right block of Awork
    g.A = at(Lt, a, a); g.B = at(L, h, a); g.C = at(Aw, a, h);
    g.mtiles = h - a; g.ntiles = b - h; g.kblocks = (h - a) * KB_PER_TILE; g.krange = KR_GE_M; g.lower_only = 0;
    g.alpha = 1.0; g.beta = 0.0;
    HDBO_TRY(launch_gemm_nt(g, batch, stream));
    // Linv21[i,j] = -sum_{k<=i} Linv22[i,k] T^T[j,k]   (+ transposed copy into Linv^T)
    g.A = at(Li, h, h); g.B = at(Aw, a, h); g.C = at(Li, h, a); g.Ct = at(Lt, a, h);
    g.mtiles = b - h; g.ntiles = h - a; g.kblocks = (b - h) * KB_PER_TILE; g.krange = KR_LE_M;
    g.alpha = -1.0; g.beta = 0.0;
    HDBO_TRY(launch_gemm_nt(g, batch, stream));
    return cudaSuccess;
}

cudaError_t chol_inverse(double* Awork, double* L, double* Linv, double* Linvt, int ld, long long bs, int tiles,
                         int n_true, int32_t* info, int batch, cudaStream_t stream) {
    (void)n_true;
    return rec(Awork, L, Linv, Linvt, ld, bs, 0, tiles, info, batch, stream);
}

}  // namespace hdbo
