// Pipelined ("look-ahead") Cholesky factor + inverse (+ K^-1) of ONE padded SPD matrix over caller-owned side streams.
//
// Why: the recursive factor-then-invert of kernels_chol.cu is a dependency chain -- at n = 4096 its 32 leaves (one SM each) and the
// small GEMMs between them take 2.7 ms during which 147 SMs idle, next to 2.4 ms of large GEMMs (profiles/r01_launches_fit_*.csv).
// Here the chain runs on its own stream and everything that is not on it runs beside it, ALL of it right-looking -- rank updates that
// become available the moment a panel exists -- and in PAIRS of 128-columns (c0, c1 = c0 + 1), so that the heavy updates have K = 256:
//
//   chain stream:  leaf(c0) -> panel(c0) -> column c1 -= panel c0 (K = 128) -> leaf(c1) -> panel(c1)
//                  -> columns c1+1, c1+2 (the next pair) -= panels c0, c1 (K = 256)
//   near stream:   forward substitution of the identity, right-looking: row k of L^-1 is finished by one scaling with L_kk^-1
//                  (L^-1(k, j) = L_kk^-1 W(k, j)); row c1 takes row c0 at once (K = 128), rows c1+1, c1+2 take both rows (K = 256) so that
//                  the next pair can be finished without waiting for the side stream
//   side stream:   ONE persistent launch per pair (k_side_updates, grid capped below the SM count so that the chain always finds SMs):
//       bulk     A(i, j)    -= L(i, c0..c1) L(j, c0..c1)^T          c1+3 <= j <= i
//       inverse  W^T(j, i)  -= L^-T(j, c0..c1) L(i, c0..c1)^T       j <= c1, i >= c1+3      (W^T = the running, unscaled rows of L^-1,
//                                                                                            transposed, in the unused upper tiles of Awork)
//       K^-1     Kinv(a, b) += L^-T(a, c0..c1) L^-T(b, c0..c1)^T    b <= a <= c1            (MLL gradient only)
//
// The three side loads add up to ~T^2 / 2 tile updates for EVERY pair (a row-by-row inverse puts 58 % of trtri + lauum behind the last
// quarter of the chain: measured 0.7 ms of tail), so potrf's trailing updates, trtri and lauum (n^3 / 3 flops each) overlap the
// latency-bound chain evenly; and the chain only meets the side stream once per pair (side(p-1) must be done before the chain touches
// the columns of pair p+1), which gives it a whole pair of slack.
// Dependencies cross streams through caller-owned events (stream-ordered record -> wait; an event is re-recorded only after the
// wait on its previous record has been enqueued).  Everything is enqueued asynchronously; the side streams are forked from and joined
// back to the caller's stream, so the whole sequence is capturable into a CUDA graph.
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
    cudaStream_t sA = cs.chain, sB = cs.bulk, sC = cs.inverse;      // chain, side, near  (cs.kinv is not needed by this schedule)
    // events: 0 fork, 1-2 leaf(c0) / leaf(c1) done, 3-4 panel(c0) / panel(c1) done, 5-6 side(p % 2) done, 7-8 row c0 / c1 of L^-1 done
    cudaEvent_t eFork = cs.ev[0], eJoin = cs.ev[11];
    cudaEvent_t eLeaf[2] = {cs.ev[1], cs.ev[2]}, ePanel[2] = {cs.ev[3], cs.ev[4]}, eSide[2] = {cs.ev[5], cs.ev[6]};
    cudaEvent_t eRow[2] = {cs.ev[7], cs.ev[8]};
    auto at = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    const long long bs = (long long)ld * ld;
    const int skip = pipe_skip_mask();
    static const int side_ctas = env_int("HDBO_PIPE_SIDE_CTAS", 120);
    // fork: the side streams start after everything already enqueued on the caller's stream (the kernel build; the upper tiles of
    // Awork -- W^T -- were zeroed before it)
    PIPE_TRY(cudaEventRecord(eFork, origin));
    PIPE_TRY(cudaStreamWaitEvent(sA, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sB, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sC, eFork, 0));
    if (Kinv) PIPE_TRY(cudaMemsetAsync(Kinv, 0, sizeof(double) * (size_t)ld * ld, sB));
    GemmParams g{};
    g.lda = g.ldb = g.ldc = g.ldct = ld;
    g.bsA = g.bsB = g.bsC = g.bsCt = bs;
    g.tile_hint = 32;                                    // every k_gemm_nt launch below is small and latency-bound
    SideParams sp{};
    sp.Aw = Aw; sp.L = L; sp.Lt = Lt; sp.Kinv = Kinv; sp.ld = ld; sp.T = T;
    // L^-1(k, j) = L_kk^-1 W(k, j) for j < k (+ transposed copy into L^-T(j, k)); row k of W must be complete
    auto scale_row = [&](int k) -> cudaError_t {
        if (k < 1 || (skip & 2)) return cudaSuccess;
        g.A = at(Li, k, k); g.B = at(Aw, 0, k); g.C = at(Li, k, 0); g.Ct = at(Lt, 0, k);
        g.mtiles = 1; g.ntiles = k; g.kblocks = KB_PER_TILE; g.krange = KR_LE_M; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        return launch_gemm_nt(g, 1, sC);
    };
    // L(i, k) = A(i, k) L_kk^-T for i > k
    auto panel = [&](int k) -> cudaError_t {
        g.A = at(Aw, k + 1, k); g.B = at(Li, k, k); g.C = at(L, k + 1, k); g.Ct = nullptr;
        g.mtiles = T - k - 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_LE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        return launch_gemm_nt(g, 1, sA);
    };
    // column c (rows >= c) -= L(., k0 .. k0+w-1) L(c, k0 .. k0+w-1)^T
    auto update_col = [&](int c, int k0, int w) -> cudaError_t {
        g.A = at(L, c, k0); g.B = at(L, c, k0); g.C = at(Aw, c, c); g.Ct = nullptr;
        g.mtiles = T - c; g.ntiles = 1; g.kblocks = w * KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 0; g.alpha = -1.0; g.beta = 1.0;
        return launch_gemm_nt(g, 1, sA);
    };
    // W^T(j, rows r .. r+nr-1) -= L^-T(j, k0 .. k0+w-1) L(rows, k0 .. k0+w-1)^T for j <= jmax
    auto near_rows = [&](int r, int nr, int jmax, int k0, int w) -> cudaError_t {
        if (skip & 2) return cudaSuccess;
        g.A = at(Lt, 0, k0); g.B = at(L, r, k0); g.C = at(Aw, 0, r); g.Ct = nullptr;
        g.mtiles = jmax + 1; g.ntiles = nr; g.kblocks = w * KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 0; g.alpha = -1.0; g.beta = 1.0;
        return launch_gemm_nt(g, 1, sC);
    };
    int p = 0;
    for (int c0 = 0; c0 < T; c0 += 2, p++) {
        const int c1 = c0 + 1;
        const bool has_c1 = c1 < T;
        const int last = has_c1 ? c1 : c0;               // last column of this pair
        // ---- chain + near, first column
        PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, c0, info, 1, sA));
        PIPE_TRY(cudaEventRecord(eLeaf[0], sA));
        PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[0], 0));
        PIPE_TRY(scale_row(c0));                         // (row c0 of W: side(<= p-2) is ordered before leaf(c0) through the chain's waits,
        PIPE_TRY(cudaEventRecord(eRow[0], sC));          //  the rows of pair p-1 came from near_rows of pair p-1)
        if (has_c1) {
            PIPE_TRY(panel(c0));
            PIPE_TRY(cudaEventRecord(ePanel[0], sA));
            PIPE_TRY(cudaStreamWaitEvent(sC, ePanel[0], 0));
            PIPE_TRY(near_rows(c1, 1, c0, c0, 1));       // row c1 takes row c0 (side(p-1) starts at row c1+1: no conflict)
            PIPE_TRY(update_col(c1, c0, 1));             // column c1 -= panel c0 (side(p-1) starts at column c1+1: no conflict)
            // ---- second column
            PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, c1, info, 1, sA));
            PIPE_TRY(cudaEventRecord(eLeaf[1], sA));
            PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[1], 0));
            PIPE_TRY(scale_row(c1));
            PIPE_TRY(cudaEventRecord(eRow[1], sC));
            if (c1 < T - 1) {
                PIPE_TRY(panel(c1));
                PIPE_TRY(cudaEventRecord(ePanel[1], sA));
            }
        }
        // ---- side: everything the pair makes possible beyond the next pair's rows / columns, one persistent launch
        const int w = has_c1 ? 2 : 1;
        const int first = last + 3;                      // columns / rows last+1, last+2 are updated by the chain / near stream below
        const int mb = T - first;
        sp.k0 = c0; sp.kblocks = w * KB_PER_TILE; sp.c0 = first; sp.r0 = first; sp.jmax = last;
        sp.n_bulk = (mb > 0 && !(skip & 1)) ? mb * (mb + 1) / 2 : 0;
        sp.n_inv = (mb > 0 && !(skip & 2)) ? (last + 1) * mb : 0;
        sp.n_kinv = (Kinv && !(skip & 4)) ? (last + 1) * (last + 2) / 2 : 0;
        if (last < T - 1) PIPE_TRY(cudaStreamWaitEvent(sB, ePanel[has_c1 ? 1 : 0], 0));
        PIPE_TRY(cudaStreamWaitEvent(sB, eRow[has_c1 ? 1 : 0], 0));
        PIPE_TRY(launch_side_updates(sp, side_ctas, sB));
        PIPE_TRY(cudaEventRecord(eSide[p & 1], sB));
        if (last >= T - 1) break;
        // ---- the next pair's rows / columns take this pair now (K = 256); side(p-1) also wrote them: ordered after it
        const int nn = (T - last - 1) < 2 ? (T - last - 1) : 2;
        PIPE_TRY(cudaStreamWaitEvent(sC, ePanel[1], 0));
        if (p >= 1) PIPE_TRY(cudaStreamWaitEvent(sC, eSide[(p - 1) & 1], 0));
        PIPE_TRY(near_rows(last + 1, nn, last, c0, w));
        if (p >= 1) PIPE_TRY(cudaStreamWaitEvent(sA, eSide[(p - 1) & 1], 0));
        for (int c = last + 1; c <= last + nn; c++) PIPE_TRY(update_col(c, c0, w));
    }
    // join: the caller's stream continues when all side streams are done
    cudaStream_t side[3] = {sA, sB, sC};
    for (int i = 0; i < 3; i++) {
        PIPE_TRY(cudaEventRecord(eJoin, side[i]));
        PIPE_TRY(cudaStreamWaitEvent(origin, eJoin, 0));
    }
    return cudaSuccess;
}

}  // namespace hdbo
