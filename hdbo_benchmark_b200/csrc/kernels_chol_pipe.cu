// Pipelined ("look-ahead") Cholesky factor + inverse (+ K^-1) of ONE padded SPD matrix over caller-owned side streams.
//
// Why: the recursive factor-then-invert of kernels_chol.cu is a dependency chain -- at n = 4096 its 32 leaves (one SM each) and the
// small GEMMs between them take 2.7 ms during which 147 SMs idle, next to 2.4 ms of large GEMMs (profiles/r01_launches_fit_*.csv).
// Here the chain is isolated on a high-priority stream and everything that is not on it runs beside it, ALL of it right-looking, i.e.
// as rank-128 updates that become available the moment panel k exists:
//
//   chain stream (per 128-column step k):   leaf(k) -> panel L(k+1.., k) = A(k+1.., k) L_kk^-T -> update of column k+1 only
//   bulk stream:                            A(i, j) -= L(i, k) L(j, k)^T for k+2 <= j <= i           ((T-k)^2 / 2 tiles at step k)
//   inverse stream:                         forward substitution of the identity, right-looking: row k of L^-1 is finished by one
//                                           scaling with L_kk^-1, then every row below is updated with it,
//                                               W^T(j, i) -= L^-T(j, k) L(i, k)^T   for j <= k < i    ((k+1)(T-k-1) tiles)
//                                           (W^T = the running, unscaled rows of L^-1, kept transposed in the unused upper tiles of
//                                           Awork so that both steps are K-contiguous "NT" products)
//   K^-1 stream (MLL gradient only):        K^-1(a, b) += L^-T(a, k) L^-T(b, k)^T  for b <= a <= k     ((k+1)^2 / 2 tiles)
//
// The three side loads add up to T^2 / 2 tile updates at EVERY step (the row-by-row inverse of the first version put 58 % of trtri +
// lauum behind the last quarter of the chain and left a 0.7 ms tail), so potrf's trailing updates, trtri and lauum (n^3 / 3 flops
// each) overlap the latency-bound chain evenly instead of following it.
// Dependencies cross streams through caller-owned events (stream-ordered record -> wait; an event is re-recorded only after the
// wait on its previous record has been enqueued).  Everything is enqueued asynchronously; the side streams are forked from and joined
// back to the caller's stream, so the whole sequence is capturable into a CUDA graph.  Same kernels as the recursive path:
// k_leaf_blocked + k_gemm_nt.
#include <cstdlib>

#include "gemm_core.cuh"
#include "kernels.cuh"

namespace hdbo {

cudaError_t launch_leaf(double* Awork, double* L, double* Linv, double* Linvt, int ld, long long bs, int tile, int32_t* info,
                        int batch, cudaStream_t stream);

// developer knobs (read once per process).  HDBO_PIPE_SKIP times the streams in isolation -- results are WRONG when set: bit 0 skip
// bulk, 1 skip inverse rows, 2 skip K^-1.
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static int pipe_skip_mask() { static const int m = env_int("HDBO_PIPE_SKIP", 0); return m; }

// CTA tile for a launch of `tiles128` 128x128 output tiles: enough CTAs to cover the 148 SMs, else the largest tile
static int pipe_tile(long long tiles128) {
    static const int t128 = env_int("HDBO_PIPE_T128", 100), t64 = env_int("HDBO_PIPE_T64", 24);
    return tiles128 >= t128 ? 128 : (tiles128 >= t64 ? 64 : 32);
}

#define PIPE_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t chol_inverse_pipelined(double* Aw, double* L, double* Li, double* Lt, double* Kinv, int ld, int T, int32_t* info,
                                   cudaStream_t origin, const ChainStreams& cs) {
    cudaStream_t sA = cs.chain, sB = cs.bulk, sC = cs.inverse, sD = cs.kinv;
    // events: 0 fork, 1-2 leaf done (k % 2), 3-4 panel done, 5-6 bulk done, 7-8 row k of L^-1 done, 9-11 joins
    cudaEvent_t eFork = cs.ev[0];
    cudaEvent_t eLeaf[2] = {cs.ev[1], cs.ev[2]}, ePanel[2] = {cs.ev[3], cs.ev[4]}, eBulk[2] = {cs.ev[5], cs.ev[6]};
    cudaEvent_t eRow[2] = {cs.ev[7], cs.ev[8]};
    auto at = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    const long long bs = (long long)ld * ld;
    const int skip = pipe_skip_mask();
    // fork: the side streams start after everything already enqueued on the caller's stream (the kernel build; the upper tiles of
    // Awork -- W^T -- were zeroed before it)
    PIPE_TRY(cudaEventRecord(eFork, origin));
    PIPE_TRY(cudaStreamWaitEvent(sA, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sB, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sC, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sD, eFork, 0));
    if (Kinv) PIPE_TRY(cudaMemsetAsync(Kinv, 0, sizeof(double) * (size_t)ld * ld, sD));
    GemmParams g{};
    g.lda = g.ldb = g.ldc = g.ldct = ld;
    g.bsA = g.bsB = g.bsC = g.bsCt = bs;
    // The inverse and K^-1 streams have no deadline on the chain, so they work on PAIRS of rows (K = 256 updates: half the C traffic
    // and pipeline fills per flop); the bulk stream feeds the chain and keeps the per-step K = 128 granularity.
    for (int k = 0; k < T; k++) {
        // ---- chain: diagonal tile k -> L_kk, L_kk^-1 (and its transpose)
        PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, k, info, 1, sA));
        PIPE_TRY(cudaEventRecord(eLeaf[k & 1], sA));
        const bool pair_end = (k & 1) || k == T - 1;      // rows (k-1, k) -- or the single last row of an odd T -- are complete after this step
        const int r0 = (k & 1) ? k - 1 : k;               // first row of the pair that ends here
        // ---- inverse stream: finish row k:  L^-1(k, j) = L_kk^-1 W(k, j)  for j < k  (+ transposed copy into L^-T(j, k))
        if (!(skip & 2)) {
            PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[k & 1], 0));
            if (k >= 1) {
                g.A = at(Li, k, k); g.B = at(Aw, 0, k); g.C = at(Li, k, 0); g.Ct = at(Lt, 0, k);
                g.mtiles = 1; g.ntiles = k; g.kblocks = KB_PER_TILE; g.krange = KR_LE_M; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
                g.tile_hint = 32;
                PIPE_TRY(launch_gemm_nt(g, 1, sC));
            }
            if (pair_end) PIPE_TRY(cudaEventRecord(eRow[(k >> 1) & 1], sC));
        }
        if (Kinv && !(skip & 4) && pair_end) {
            // K^-1[a, b] += sum_{rr in rows r0..k} L^-1[rr, a] L^-1[rr, b]  for b <= a <= k   (lower tiles; L^-T(a, rr) = 0 for a > rr)
            PIPE_TRY(cudaStreamWaitEvent(sD, (skip & 2) ? eLeaf[k & 1] : eRow[(k >> 1) & 1], 0));
            g.A = at(Lt, 0, r0); g.B = at(Lt, 0, r0); g.C = Kinv; g.Ct = nullptr;
            g.mtiles = g.ntiles = k + 1; g.kblocks = (k - r0 + 1) * KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 1; g.alpha = 1.0; g.beta = 1.0;
            g.tile_hint = pipe_tile((long long)(k + 1) * (k + 2) / 2);
            PIPE_TRY(launch_gemm_nt(g, 1, sD));
            g.lower_only = 0;
        }
        if (k == T - 1) break;
        // ---- chain: panel below the diagonal tile, L[i, k] = A[i, k] L_kk^-T
        g.A = at(Aw, k + 1, k); g.B = at(Li, k, k); g.C = at(L, k + 1, k); g.Ct = nullptr;
        g.mtiles = T - k - 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_LE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        g.tile_hint = 32;
        PIPE_TRY(launch_gemm_nt(g, 1, sA));
        PIPE_TRY(cudaEventRecord(ePanel[k & 1], sA));
        // ---- inverse stream: rows below take the finished rows:  W^T(j, i) -= L^-T(j, rows) L(i, rows)^T
        if (!(skip & 2)) {
            PIPE_TRY(cudaStreamWaitEvent(sC, ePanel[k & 1], 0));
            g.A = at(Lt, 0, k); g.B = at(L, k + 1, k); g.C = at(Aw, 0, k + 1); g.Ct = nullptr;
            g.krange = KR_FULL; g.lower_only = 0; g.alpha = -1.0; g.beta = 1.0;
            if (!(k & 1)) {
                // first row of a pair: only the NEXT row needs it now (so that it can be finished): tiles (j, k+1), j <= k, K = 128
                g.mtiles = k + 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.tile_hint = 32;
                PIPE_TRY(launch_gemm_nt(g, 1, sC));
            } else if (k + 1 < T) {
                // second row of a pair: all rows below take both rows at once: tiles (j, i), j <= k, i > k, K = 256
                g.A = at(Lt, 0, k - 1); g.B = at(L, k + 1, k - 1);
                g.mtiles = k + 1; g.ntiles = T - k - 1; g.kblocks = 2 * KB_PER_TILE;
                g.tile_hint = pipe_tile((long long)(k + 1) * (T - k - 1));
                PIPE_TRY(launch_gemm_nt(g, 1, sC));
            }
        }
        // ---- bulk: A[i, j] -= L[i, k] L[j, k]^T for k+2 <= j <= i
        if (k + 2 < T && !(skip & 1)) {
            PIPE_TRY(cudaStreamWaitEvent(sB, ePanel[k & 1], 0));
            g.A = at(L, k + 2, k); g.B = at(L, k + 2, k); g.C = at(Aw, k + 2, k + 2); g.Ct = nullptr;
            g.mtiles = g.ntiles = T - k - 2; g.kblocks = KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 1; g.alpha = -1.0; g.beta = 1.0;
            g.tile_hint = pipe_tile((long long)(T - k - 2) * (T - k - 1) / 2);
            PIPE_TRY(launch_gemm_nt(g, 1, sB));
            PIPE_TRY(cudaEventRecord(eBulk[k & 1], sB));
            g.lower_only = 0;
        }
        // ---- chain: column k+1 only, A[i, k+1] -= L[i, k] L[k+1, k]^T  (its earlier updates came from bulk(k-1): wait for it)
        if (k >= 1 && !(skip & 1)) PIPE_TRY(cudaStreamWaitEvent(sA, eBulk[(k - 1) & 1], 0));
        g.A = at(L, k + 1, k); g.B = at(L, k + 1, k); g.C = at(Aw, k + 1, k + 1); g.Ct = nullptr;
        g.mtiles = T - k - 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_FULL; g.alpha = -1.0; g.beta = 1.0;
        g.tile_hint = 32;
        PIPE_TRY(launch_gemm_nt(g, 1, sA));
    }
    // join: the caller's stream continues when all side streams are done
    cudaStream_t side[4] = {sA, sB, sC, sD};
    for (int i = 0; i < 4; i++) {
        PIPE_TRY(cudaEventRecord(cs.ev[9 + (i % 3)], side[i]));
        PIPE_TRY(cudaStreamWaitEvent(origin, cs.ev[9 + (i % 3)], 0));
    }
    return cudaSuccess;
}

}  // namespace hdbo
