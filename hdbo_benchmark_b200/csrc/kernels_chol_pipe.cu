// Pipelined ("look-ahead") Cholesky factor + inverse (+ K^-1) of ONE padded SPD matrix over caller-owned side streams.
//
// Why: the recursive factor-then-invert of kernels_chol.cu is a dependency chain -- at n = 4096 its 32 leaves (one SM each) and the
// small GEMMs between them take 2.7 ms during which 147 SMs idle, next to 2.4 ms of large GEMMs (profiles/r01_launches_fit_*.csv).
// Here the chain runs on its own stream and carries ONLY what the next leaf needs -- single 128x128 tiles -- while everything else runs
// beside it, ALL of it right-looking (rank updates that become available the moment a panel exists) and in PAIRS of 128-columns
// (c0, c1 = c0 + 1; n0, n1 = the next pair; f = n1 + 1 the first "far" row), so that the heavy updates have K = 256:
//
//   chain stream:  leaf(c0) -> L(c1, c0) = A(c1, c0) L00^-T -> A(c1, c1) -= L(c1, c0) L(c1, c0)^T -> leaf(c1)
//                  -> L(n0..n1, c1) = A'(n0..n1, c1) L11^-T -> the next pair's diagonal block (3 tiles) -= L(., c0..c1) L(., c0..c1)^T
//                  -> leaf(n0) ...       (v5 had whole panels and column updates on the chain: 117 us leaf to leaf, the leaf being 50)
//   panel stream:  beside leaf(c1): L(n0..n1, c0) and A'(n0..n1, c1) = A - L(., c0) L(c1, c0)^T  (what the chain needs next);
//                  after leaf(c1): the REST of the pair's panel in one GEMM, L(i, c0..c1) = A(i, c0..c1) Linv(c0..c1, c0..c1)^T for
//                  i >= f -- through the pair's 256x256 inverse, so column c1 needs no in-pair update
//   near stream:   W^T(c0, c1) = -L00^-T L(c1, c0)^T beside leaf(c1), Linv(c1, c0) = L11^-1 W(c1, c0) after it: the off-diagonal tile of the
//                  pair's 256x256 inverse (W = the running, unscaled rows of L^-1; W^T lives in the unused upper tiles of Awork)
//   side stream:   TWO persistent launches per pair over one job list (k_side_updates, grid capped below the SM count so that the
//                  chain always finds SMs), the list ordered so that launch A (one round of the grid) contains the "head" -- the
//                  jobs on the next pair's tile columns / rows and on the diagonal block of the pair after it -- and an event after
//                  it releases the next pair's panel / rows and the chain's next diagonal-block update:
//       bulk     A(i, j)    -= L(i, c0..c1) L(j, c0..c1)^T          n0 <= j <= i, except the next pair's diagonal block
//       rows     Linv(c0..c1, j) = Linv(c0..c1, c0..c1) W(c0..c1, j)     j < c0: rows c0, c1 of L^-1 through the pair's inverse
//       inverse  W^T(j, i)  -= L^-T(j, c0..c1) L(i, c0..c1)^T       j <= c1, i >= n0
//       K^-1     Kinv(a, b) += L^-T(a, c0..c1) L^-T(b, c0..c1)^T    b <= a <= c1            (MLL gradient only)
//
// Every tile receives its updates in ascending order of the source pair, each ordered after the previous one through stream order or
// an event: no two writers of a tile are ever concurrent, and the summation order -- hence the result -- does not depend on timing.
// The side loads add up to T^2/2 + T/2 - 3 tile updates for EVERY pair (a row-by-row inverse puts 58 % of trtri + lauum behind the
// last quarter of the chain: measured 0.7 ms of tail), so potrf's trailing updates, trtri and lauum (n^3 / 3 flops each) overlap the
// latency-bound chain evenly.
// Dependencies cross streams through caller-owned events (stream-ordered record -> wait; every wait is enqueued before the event's
// next record).  Everything is enqueued asynchronously; the side streams are forked from and joined back to the caller's stream, so
// the whole sequence is capturable into a CUDA graph.
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

#define PIPE_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

// The caller builds only the first four tile columns of Khat (column-major tile order: what the first pair's chain, panel and the second
// pair's diagonal block read) before the fork; the rest of the build runs on the side stream beside the first pair's chain.
int pipe_build_head_tiles(int T) {
    int t = 0;
    for (int j = 0; j < 4 && j < T; j++) t += T - j;
    return t;
}

cudaError_t chol_inverse_pipelined(double* Aw, double* L, double* Li, double* Lt, double* Kinv, int ld, int T, int32_t* info,
                                   cudaStream_t origin, const ChainStreams& cs, const SymBuildParams* build_rest) {
    cudaStream_t sA = cs.chain, sB = cs.bulk, sC = cs.inverse, sP = cs.panel;      // chain, side, near, panel
    cudaEvent_t eFork = cs.ev[0], eJoin = cs.ev[11];
    cudaEvent_t eLeaf[2] = {cs.ev[1], cs.ev[2]};       // leaf(c0) / leaf(c1) done                                   chain -> panel, near
    cudaEvent_t eTrsm0 = cs.ev[3];                     // L(c1, c0) exists                                           chain -> panel, near
    cudaEvent_t eUpd0 = cs.ev[4];                      // L(n0..n1, c0), A'(n0..n1, c1) ready                        panel -> chain
    cudaEvent_t eInv10 = cs.ev[5];                     // Linv(c1, c0) exists (the pair's 256x256 inverse is whole)  near -> panel
    cudaEvent_t eChain = cs.ev[6];                     // L(n0..n1, c1) exists, next diagonal block updated          chain -> side, panel
    cudaEvent_t ePanel = cs.ev[7];                     // rest of the pair's panel (rows >= f) done                  panel -> side
    cudaEvent_t eHead = cs.ev[9];                      // launch A of side(p) (with the head) done                   side -> panel (-> chain)
    // (ev[8] and ev[10] are spare: earlier schedules needed separate row / side-done events)
    auto at = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    const long long bs = (long long)ld * ld;
    const int skip = pipe_skip_mask();
    static const int side_ctas = env_int("HDBO_PIPE_SIDE_CTAS", 140);
    static const int rest_tile = env_int("HDBO_PIPE_REST_TILE", 32);       // CTA tile of the panel-rest / row-pair GEMMs
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int cap = side_ctas < sms ? side_ctas : sms;
    // fork: the side streams start after everything already enqueued on the caller's stream (the head of the kernel build).  Neither
    // W^T (the upper tiles of Awork) nor Kinv needs zeroing: the first job on each of their tiles starts from zero accumulators
    PIPE_TRY(cudaEventRecord(eFork, origin));
    PIPE_TRY(cudaStreamWaitEvent(sA, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sB, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sC, eFork, 0));
    PIPE_TRY(cudaStreamWaitEvent(sP, eFork, 0));
    if (build_rest) {                                    // the far columns of Khat, beside the first pair's chain (capped like the side)
        SymBuildParams br = *build_rest;
        br.max_ctas = cap;
        PIPE_TRY(launch_sym_build(br, 1, sB));
    }
    GemmParams g{};
    g.lda = g.ldb = g.ldc = g.ldct = ld;
    g.bsA = g.bsB = g.bsC = g.bsCt = bs;
    SideParams sp{};
    sp.Aw = Aw; sp.L = L; sp.Li = Li; sp.Lt = Lt; sp.Kinv = Kinv; sp.ld = ld; sp.T = T;
    // L(r0 .. r0+nr-1, k .. k+w-1) = A(rows, k ..) Linv(k .. k+w-1, k .. k+w-1)^T   (w = 2: through the pair's 256x256 inverse)
    auto panel = [&](int k, int w, int r0, int nr, int tile, cudaStream_t st) -> cudaError_t {
        if (nr <= 0) return cudaSuccess;
        g.A = at(Aw, r0, k); g.B = at(Li, k, k); g.C = at(L, r0, k); g.Ct = nullptr;
        g.mtiles = nr; g.ntiles = w; g.kblocks = w * KB_PER_TILE; g.krange = KR_LE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        g.tile_hint = tile;
        return launch_gemm_nt(g, 1, st);
    };
    // A(r0 .. r0+nr-1, c) -= L(rows, k0 .. k0+w-1) L(c, k0 .. k0+w-1)^T
    auto update_col = [&](int c, int r0, int nr, int k0, int w, cudaStream_t st) -> cudaError_t {
        if (nr <= 0) return cudaSuccess;
        g.A = at(L, r0, k0); g.B = at(L, c, k0); g.C = at(Aw, r0, c); g.Ct = nullptr;
        g.mtiles = nr; g.ntiles = 1; g.kblocks = w * KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 0; g.alpha = -1.0; g.beta = 1.0;
        g.tile_hint = 32;
        return launch_gemm_nt(g, 1, st);
    };
    const int p_last = (T + 1) / 2 - 1;
    int p = 0;
    for (int c0 = 0; c0 < T; c0 += 2, p++) {
        const int c1 = c0 + 1;
        const bool has_c1 = c1 < T;
        const int last = has_c1 ? c1 : c0;               // last column of this pair
        const int w = has_c1 ? 2 : 1;
        const int n0 = last + 1;                         // first column of the next pair
        const int nn = (T - n0) < 2 ? (T - n0 > 0 ? T - n0 : 0) : 2;
        const int f = n0 + 2;                            // first far row
        const int mb = T - f > 0 ? T - f : 0;
        // ---- chain: first column.  The pair's diagonal block is final: its last update ran on the chain itself
        PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, c0, info, 1, sA));
        PIPE_TRY(cudaEventRecord(eLeaf[0], sA));
        if (has_c1) {
            PIPE_TRY(panel(c0, 1, c1, 1, 32, sA));
            PIPE_TRY(cudaEventRecord(eTrsm0, sA));
            PIPE_TRY(update_col(c1, c1, 1, c0, 1, sA));
            PIPE_TRY(launch_leaf(Aw, L, Li, Lt, ld, bs, c1, info, 1, sA));
            PIPE_TRY(cudaEventRecord(eLeaf[1], sA));
            // ---- panel stream, beside leaf(c1): what the chain needs right after it
            if (nn > 0) {
                PIPE_TRY(cudaStreamWaitEvent(sP, eLeaf[0], 0));
                if (p >= 1) PIPE_TRY(cudaStreamWaitEvent(sP, eHead, 0));              // A(n0.., c0..c1) carries every pair < p
                PIPE_TRY(panel(c0, 1, n0, nn, 32, sP));
                PIPE_TRY(cudaStreamWaitEvent(sP, eTrsm0, 0));
                PIPE_TRY(update_col(c1, n0, nn, c0, 1, sP));
                PIPE_TRY(cudaEventRecord(eUpd0, sP));
            }
            // ---- near stream: the off-diagonal tile of the pair's inverse.  W^T(c0, c1) = -L00^-T L(c1, c0)^T beside leaf(c1) ...
            if (!(skip & 2)) {
                PIPE_TRY(cudaStreamWaitEvent(sC, eTrsm0, 0));
                g.A = at(Lt, c0, c0); g.B = at(L, c1, c0); g.C = at(Aw, c0, c1); g.Ct = nullptr;
                g.mtiles = 1; g.ntiles = 1; g.kblocks = KB_PER_TILE; g.krange = KR_GE_M; g.lower_only = 0; g.alpha = -1.0; g.beta = 0.0;
                g.tile_hint = 32;
                PIPE_TRY(launch_gemm_nt(g, 1, sC));
                // ... Linv(c1, c0) = L11^-1 W(c1, c0) (+ transposed copy into L^-T(c0, c1)) after it
                PIPE_TRY(cudaStreamWaitEvent(sC, eLeaf[1], 0));
                g.A = at(Li, c1, c1); g.B = at(Aw, c0, c1); g.C = at(Li, c1, c0); g.Ct = at(Lt, c0, c1);
                g.krange = KR_LE_M; g.alpha = 1.0;
                PIPE_TRY(launch_gemm_nt(g, 1, sC));
            }
            PIPE_TRY(cudaEventRecord(eInv10, sC));
        }
        if (nn > 0) {
            // ---- chain: the next pair's diagonal block (has_c1 holds: a single last column has no next pair)
            PIPE_TRY(cudaStreamWaitEvent(sA, eUpd0, 0));
            PIPE_TRY(panel(c1, 1, n0, nn, 32, sA));
            g.A = at(L, n0, c0); g.B = at(L, n0, c0); g.C = at(Aw, n0, n0); g.Ct = nullptr;
            g.mtiles = g.ntiles = nn; g.kblocks = w * KB_PER_TILE; g.krange = KR_FULL; g.lower_only = 1; g.alpha = -1.0; g.beta = 1.0;
            g.tile_hint = 32;
            PIPE_TRY(launch_gemm_nt(g, 1, sA));         // (eHead of side(p-1): the panel stream waited for it before eUpd0)
            PIPE_TRY(cudaEventRecord(eChain, sA));
            // ---- panel stream: the rest of the pair's panel through the pair's inverse -- AFTER the chain's two launches (it has
            // 70 us of slack; started earlier, its few hundred small CTAs took the free SMs from the chain: 37 instead of 10 us)
            if (mb > 0) {
                PIPE_TRY(cudaStreamWaitEvent(sP, eInv10, 0));
                PIPE_TRY(cudaStreamWaitEvent(sP, eChain, 0));
                PIPE_TRY(panel(c0, w, f, mb, rest_tile, sP));
            }
            PIPE_TRY(cudaEventRecord(ePanel, sP));
        }
        // ---- side: every rank-256 update the pair makes possible, the pair's rows of L^-1 and the previous pair's K^-1 jobs; launch A =
        // one round of the grid (starts with the head; nothing in it reads this pair's rows of L^-1), launch B = the rest
        {
            sp.k0 = c0; sp.kblocks = w * KB_PER_TILE; sp.w = w; sp.n0 = n0; sp.nn = nn; sp.f = f; sp.jmax = last;
            sp.e_k0 = c0 - 2; sp.e_jmax = c0 - 1; sp.e_kblocks = 2 * KB_PER_TILE;
            const bool kinv_on = Kinv && !(skip & 4);
            sp.n_a = (skip & 1) ? 0 : mb * nn;
            sp.n_s = (skip & 2) ? 0 : w * c0;
            sp.n_c = (skip & 1) ? 0 : mb * (mb + 1) / 2;
            // K^-1 jobs run one pair late (they fill launch A, which must not read this pair's rows of L^-1) -- except at the end: the
            // last pair but one takes the previous pair's and its own as ONE list of K = 512 jobs (two lists on the same tiles in one
            // launch would race), the last pair its own, so that the tail after the last leaf stays short
            sp.n_e = (kinv_on && p >= 1 && p < p_last - 1) ? c0 * (c0 + 1) / 2 : 0;
            sp.n_b = (skip & 2) ? 0 : (last + 1) * nn;
            sp.n_d = (skip & 2) ? 0 : (last + 1) * mb;
            sp.n_e2 = (kinv_on && p >= p_last - 1) ? (last + 1) * (last + 2) / 2 : 0;
            const bool merged = p == p_last - 1 && p >= 1;                             // own + previous pair's columns in one K = 512 job
            sp.e2_k0 = merged ? c0 - 2 : c0; sp.e2_kblocks = (merged ? w + 2 : w) * KB_PER_TILE;
            const int n_l = sp.n_a + sp.n_s + sp.n_c + sp.n_e;                         // jobs that do not read this pair's rows of L^-1
            const int total = n_l + sp.n_b + sp.n_d + sp.n_e2;
            const int n_head = sp.n_a + sp.n_s + (sp.n_c < mb + 1 ? sp.n_c : mb + 1);  // ... up to tile (f+1, f+1) of region c
            // the last pair's list has the whole GPU: the chain is done
            const int ctas = nn > 0 ? cap : sms;
            int cut = n_head > ctas ? n_head : ctas;
            if (cut > n_l) cut = n_l;
            if (nn > 0) {
                PIPE_TRY(cudaStreamWaitEvent(sB, ePanel, 0));
                PIPE_TRY(cudaStreamWaitEvent(sB, eChain, 0));
            }
            PIPE_TRY(cudaStreamWaitEvent(sB, has_c1 ? eInv10 : eLeaf[0], 0));          // region s: the pair's 256x256 inverse is whole
            sp.job0 = 0; sp.job1 = cut;
            PIPE_TRY(launch_side_updates(sp, ctas, sB));
            PIPE_TRY(cudaEventRecord(eHead, sB));
            sp.job0 = cut; sp.job1 = total;
            PIPE_TRY(launch_side_updates(sp, ctas, sB));
        }
        if (nn == 0) break;
    }
    // join: the caller's stream continues when all side streams are done
    cudaStream_t side[4] = {sA, sB, sC, sP};
    for (int i = 0; i < 4; i++) {
        PIPE_TRY(cudaEventRecord(eJoin, side[i]));
        PIPE_TRY(cudaStreamWaitEvent(origin, eJoin, 0));
    }
    return cudaSuccess;
}

}  // namespace hdbo
