// Pipelined ("look-ahead") Cholesky factor + inverse (+ K^-1) of ONE padded SPD matrix over caller-owned side streams.
//
// Why: the recursive factor-then-invert of kernels_chol.cu is a dependency chain -- at n = 4096 its 32 leaves (one SM each) and the
// small GEMMs between them take 2.7 ms during which 147 SMs idle, next to 2.4 ms of large GEMMs (profiles/r01_launches_fit_*.csv).
// Here the chain is isolated on a high-priority stream and everything that is not on it runs beside it:
//
//   chain stream (per 128-column step k):   leaf(k) -> panel L(k+1.., k) = A(k+1.., k) L_kk^-T -> update of column k+1 only
//   bulk stream:                            trailing update of columns >= k+2 with panel k      (right-looking, K = 128)
//   inverse stream:                         row k of L^-1 as soon as row k of L and L_kk^-1 exist:
//                                               P^T = L^-T(0..k-1, 0..k-1) L(k, 0..k-1)^T  ->  L^-1(k, 0..k-1) = -L_kk^-1 P
//                                           and, when the MLL gradient is wanted, the rank-128 update of K^-1 = L^-T L^-1 with that row
//
// so potrf's trailing updates (n^3/3), trtri (n^3/3) and lauum (n^3/3) overlap the latency-bound chain instead of following it.
// Dependencies cross streams through caller-owned events (stream-ordered record -> wait; an event is re-recorded only after the
// wait on its previous record has been enqueued).  Everything is enqueued asynchronously; the side streams are forked from and joined
// back to the caller's stream, so the whole sequence is capturable into a CUDA graph.  Same kernels, same arithmetic per tile as the
// recursive path: k_leaf_blocked + k_gemm_nt.
#include "gemm_core.cuh"
#include "kernels.cuh"

namespace hdbo {

cudaError_t launch_leaf(double* Awork, double* L, double* Linv, double* Linvt, int ld, long long bs, int tile, int32_t* info,
                        int batch, cudaStream_t stream);

#define PIPE_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t chol_inverse_pipelined(double* Aw, double* L, double* Li, double* Lt, double* Kinv, int ld, int T, int32_t* info,
                                   cudaStream_t origin, const ChainStreams& cs) {
    cudaStream_t sA = cs.chain, sB = cs.bulk, sC = cs.inverse;
    // events: 0 fork, 1-2 leaf done (k % 2), 3-4 panel done, 5-6 bulk done, 7 join
    cudaEvent_t eFork = cs.ev[0], eJoin = cs.ev[7];
    cudaEvent_t eLeaf[2] = {cs.ev[1], cs.ev[2]}, ePanel[2] = {cs.ev[3], cs.ev[4]}, eBulk[2] = {cs.ev[5], cs.ev[6]};
    auto at = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    const long long bs = (long long)ld * ld;
    // fork: the side streams start after everything already enqueued on the caller's stream (the kernel build)
    PIPE_TRY(cudaEventRecord(eFork, origin));
    PIPE_TRY(cudaStreamWaitEvent(sA, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sB, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sC, eFork, 0));
    if (Kinv) PIPE_TRY(cudaMemsetAsync(Kinv, 0, sizeof(double) * (size_t)ld * ld, sC));
    GemmParams g{};
    g.lda = g.ldb = g.ldc = g.ldct = ld;
    g.bsA = g.bsB = g.bsC = g.bsCt = bs;
    for (int k = 0; k < T; k++) {
        // ---- chain: diagonal tile k -> L_kk, L_kk^-1 (and its transpose)
        PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, k, info, 1, sA));
        PIPE_TRY(cudaEventRecord(eLeaf[k & 1], sA));
        // ---- inverse stream: row k of L^-1 (needs L(k, 0..k-1): panel k-1; L_kk^-1: leaf k; rows < k of L^-1: stream order)
        if (k >= 1) {
            PIPE_TRY(cudaStreamWaitEvent(sC, ePanel[(k - 1) & 1], 0));
            // P^T[j, k] = sum_{kk >= j} L^-T[j, kk] L[k, kk]^T   -> scratch: upper tiles (j, k) of Awork (never touched otherwise)
            g.A = at(Lt, 0, 0); g.B = at(L, k, 0); g.C = at(Aw, 0, k); g.Ct = nullptr;
            g.mtiles = k; g.ntiles = 1; g.kblocks = k * KB_PER_TILE; g.krange = KR_GE_M; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
            PIPE_TRY(launch_gemm_nt(g, 1, sC));
            PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[k & 1], 0));
            // L^-1[k, j] = -sum_kk L_kk^-1[., kk] P^T[j, kk]   (+ transposed copy into L^-T[j, k])
            g.A = at(Li, k, k); g.B = at(Aw, 0, k); g.C = at(Li, k, 0); g.Ct = at(Lt, 0, k);
            g.mtiles = 1; g.ntiles = k; g.kblocks = KB_PER_TILE; g.krange = KR_LE_M; g.alpha = -1.0; g.beta = 0.0;
            PIPE_TRY(launch_gemm_nt(g, 1, sC));
        } else if (Kinv) {
            PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[0], 0));
        }
        if (Kinv) {
            // K^-1[a, b] += sum_{r in row block k} L^-1[r, a] L^-1[r, b]  for a >= b, a, b <= k   (lower tiles)
            g.A = at(Lt, 0, k); g.B = at(Lt, 0, k); g.C = Kinv; g.Ct = nullptr;
            g.mtiles = g.ntiles = k + 1; g.kblocks = KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 1; g.alpha = 1.0; g.beta = 1.0;
            PIPE_TRY(launch_gemm_nt(g, 1, sC));
            g.lower_only = 0;
        }
        if (k == T - 1) break;
        // ---- chain: panel below the diagonal tile, L[i, k] = A[i, k] L_kk^-T
        g.A = at(Aw, k + 1, k); g.B = at(Li, k, k); g.C = at(L, k + 1, k); g.Ct = nullptr;
        g.mtiles = T - k - 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_LE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        PIPE_TRY(launch_gemm_nt(g, 1, sA));
        PIPE_TRY(cudaEventRecord(ePanel[k & 1], sA));
        // ---- bulk: A[i, j] -= L[i, k] L[j, k]^T for k+2 <= j <= i
        if (k + 2 < T) {
            PIPE_TRY(cudaStreamWaitEvent(sB, ePanel[k & 1], 0));
            g.A = at(L, k + 2, k); g.B = at(L, k + 2, k); g.C = at(Aw, k + 2, k + 2); g.Ct = nullptr;
            g.mtiles = g.ntiles = T - k - 2; g.kblocks = KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 1; g.alpha = -1.0; g.beta = 1.0;
            PIPE_TRY(launch_gemm_nt(g, 1, sB));
            PIPE_TRY(cudaEventRecord(eBulk[k & 1], sB));
            g.lower_only = 0;
        }
        // ---- chain: column k+1 only, A[i, k+1] -= L[i, k] L[k+1, k]^T  (its earlier updates came from bulk(k-1): wait for it)
        if (k >= 1) PIPE_TRY(cudaStreamWaitEvent(sA, eBulk[(k - 1) & 1], 0));
        g.A = at(L, k + 1, k); g.B = at(L, k + 1, k); g.C = at(Aw, k + 1, k + 1); g.Ct = nullptr;
        g.mtiles = T - k - 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_FULL; g.alpha = -1.0; g.beta = 1.0;
        PIPE_TRY(launch_gemm_nt(g, 1, sA));
    }
    // join: the caller's stream continues when all three side streams are done
    PIPE_TRY(cudaEventRecord(eJoin, sA));
    PIPE_TRY(cudaStreamWaitEvent(origin, eJoin, 0));
    PIPE_TRY(cudaEventRecord(eFork, sB));
    PIPE_TRY(cudaStreamWaitEvent(origin, eFork, 0));
    PIPE_TRY(cudaEventRecord(eLeaf[0], sC));
    PIPE_TRY(cudaStreamWaitEvent(origin, eLeaf[0], 0));
    return cudaSuccess;
}

}  // namespace hdbo
