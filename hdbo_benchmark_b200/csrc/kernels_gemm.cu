// GEMM-shaped kernels of the GP hot path, all built on gemm_nt_mainloop (DMMA.8x8x4, cp.async ring).
//   k_gemm_nt    generic C = alpha*A*B^T + beta*C with triangular k-ranges   (potrf / trtri / K^-1 / beta phases)
//   k_cross      K*(Z, X): scaled-candidate x scaled-train GEMM + distance->kernel epilogue + partial K* alpha  (A1, A4)
//   k_sym_build  K(X, X) + noise I (lower tiles) and the squared-distance matrix                                 (A1, A2)
//   k_var        V = K* L^-T with the triangular k-range, row sum-of-squares epilogue                            (A4)
#include <cstdlib>

#include "gemm_core.cuh"
#include "kernels.cuh"

namespace hdbo {

// k-block range of tile (bm, bn) for tiles of TM x TN elements (kpm / kpn = k-blocks per tile edge)
__device__ __forceinline__ void krange_blocks(int mode, int bm, int bn, int kblocks, int kpm, int kpn, int& kb0, int& kb1) {
    kb0 = 0; kb1 = kblocks;
    switch (mode) {
        case KR_LE_N: kb1 = min(kblocks, (bn + 1) * kpn); break;
        case KR_LE_M: kb1 = min(kblocks, (bm + 1) * kpm); break;
        case KR_GE_N: kb0 = bn * kpn; break;
        case KR_GE_M: kb0 = bm * kpm; break;
        case KR_GE_MN: kb0 = max(bm * kpm, bn * kpn); break;
        default: break;
    }
}

__device__ __forceinline__ void lower_tile_decode(int t, int& bm, int& bn) {
    // t enumerates (bm, bn) with bm >= bn row by row: t = bm(bm+1)/2 + bn
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= t) r++;
    while (r * (r + 1) / 2 > t) r--;
    bm = r; bn = t - r * (r + 1) / 2;
}

// --------------------------------------------------------------------------------------------------------------
// p.mtiles / p.ntiles count tiles of G::TM x G::TN here (the launcher rescales the 128-tile counts).
template <class G>
__global__ void __launch_bounds__(G::THREADS) k_gemm_nt(GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    int bm, bn;
    if (p.lower_only) lower_tile_decode(blockIdx.x, bm, bn);
    else {
        // CTAs are dispatched in linear block order: put the tiles with the longest k-range first so the ragged
        // triangular work drains evenly over the SMs (heavy-first list scheduling).
        const int id = blockIdx.y * gridDim.x + blockIdx.x;
        switch (p.krange) {
            case KR_LE_N: bn = p.ntiles - 1 - id / p.mtiles; bm = id % p.mtiles; break;
            case KR_GE_N: bn = id / p.mtiles; bm = id % p.mtiles; break;
            case KR_LE_M: bm = p.mtiles - 1 - id / p.ntiles; bn = id % p.ntiles; break;
            case KR_GE_M: case KR_GE_MN: bm = id / p.ntiles; bn = id % p.ntiles; break;
            default: bm = blockIdx.x; bn = blockIdx.y; break;
        }
    }
    const int b = blockIdx.z;
    int kb0, kb1;
    int part = -1;                                              // KR_*_SPLIT: 0 / 1 = first / second part of a split column, -1 = whole
    if (!p.lower_only && (p.krange == KR_LE_N_SPLIT || p.krange == KR_GE_N_SPLIT)) {
        // virtual columns in order of decreasing k-range: the H first parts (H tiles each), then alternately the second part of a
        // split column and the unsplit column of the same length
        const int id = blockIdx.y * gridDim.x + blockIdx.x, v = id / p.mtiles, N = p.ntiles, H = N / 2;
        bm = id % p.mtiles;
        const bool le = p.krange == KR_LE_N_SPLIT;
        if (v < H) { bn = le ? N - 1 - v : v; part = 0; }
        else {
            const int u = v - H, k = H - u / 2;                 // k tiles of contraction left
            if ((u & 1) == 0) { bn = le ? k + H - 1 : H - k; part = 1; }
            else bn = le ? k - 1 : N - k;
        }
        const int kh = H * G::KB_PER_TN;
        if (le) { kb0 = part == 1 ? kh : 0; kb1 = part == 0 ? kh : (bn + 1) * G::KB_PER_TN; }
        else { kb0 = part == 0 ? kh : bn * G::KB_PER_TN; kb1 = part == 1 ? kh : p.kblocks; }
    } else {
        krange_blocks(p.krange, bm, bn, p.kblocks, G::KB_PER_TM, G::KB_PER_TN, kb0, kb1);
    }
    double* C = (part == 1 ? p.C2 : p.C) + b * p.bsC;
    const int ldc = part == 1 ? p.ldc2 : p.ldc;
    const int col0 = part == 1 ? p.c2_col0 : 0;                 // column origin of the output buffer
    double* Ct = p.Ct ? p.Ct + b * p.bsCt : nullptr;
    typename G::Frag acc;
    // beta != 0 (the rank-128 updates of the factorisation: K = 128, so a tile is 8 k-blocks long and the C round trip is a visible
    // part of it): C is read INTO the accumulators before the mainloop -- the loads overlap the pipeline fill instead of following the
    // last DMMA -- as (beta / alpha) C, so that the result is alpha (A B^T + (beta / alpha) C); exact for the +-1 factors used here
    if (p.beta != 0.0) {
        const double f = p.beta / p.alpha;
#pragma unroll
        for (int mt = 0; mt < G::MT; mt++) {
            const int i = bm * G::TM + G::frag_row(mt);
#pragma unroll
            for (int nt = 0; nt < G::NT; nt++) {
                const double2 c = *reinterpret_cast<const double2*>(C + (size_t)i * ldc + (bn * G::TN + G::frag_col(nt) - col0));
                acc.v[mt][nt][0] = f * c.x; acc.v[mt][nt][1] = f * c.y;
            }
        }
    } else {
        acc.zero();
    }
    G::mainloop(p.A + b * p.bsA + (size_t)bm * G::TM * p.lda, p.lda, p.B + b * p.bsB + (size_t)bn * G::TN * p.ldb, p.ldb,
                kb0, kb1, smem, acc);
#pragma unroll
    for (int mt = 0; mt < G::MT; mt++) {
        const int i = bm * G::TM + G::frag_row(mt);
#pragma unroll
        for (int nt = 0; nt < G::NT; nt++) {
            const int j = bn * G::TN + G::frag_col(nt);
            double2 o;
            o.x = p.alpha * acc.v[mt][nt][0];
            o.y = p.alpha * acc.v[mt][nt][1];
            *reinterpret_cast<double2*>(C + (size_t)i * ldc + (j - col0)) = o;
            if (Ct) { Ct[(size_t)j * p.ldct + i] = o.x; Ct[(size_t)(j + 1) * p.ldct + i] = o.y; }
        }
    }
}

template <class G>
static cudaError_t launch_gemm_cfg(GemmParams p, int batch, cudaStream_t stream) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_nt<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    p.mtiles *= 128 / G::TM;
    p.ntiles *= 128 / G::TN;
    dim3 grid;
    if (p.lower_only) grid = dim3(p.mtiles * (p.mtiles + 1) / 2, 1, batch);
    else if (p.krange == KR_LE_N_SPLIT || p.krange == KR_GE_N_SPLIT) grid = dim3(p.mtiles, p.ntiles + p.ntiles / 2, batch);
    else grid = dim3(p.mtiles, p.ntiles, batch);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    k_gemm_nt<G><<<grid, G::THREADS, G::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

// Tile-size choice (thresholds measured on the n=4096 fit: 6.51 -> 6.26 ms with 600/60 instead of 100/25): a launch with few
// 128-tiles is latency-bound or drains in ragged waves, so it is cut into 64x64 or 32x32 CTA
// tiles (4x / 16x more CTAs, each proportionally shorter) until the grid covers the 148 SMs.
cudaError_t launch_gemm_nt(const GemmParams& p, int batch, cudaStream_t stream) {
    const long long t128 = (p.lower_only ? (long long)p.mtiles * (p.mtiles + 1) / 2 : (long long)p.mtiles * p.ntiles) * batch;
    static const long long t128_min = [] { const char* e = getenv("HDBO_GEMM_T128_MIN"); return e ? atoll(e) : 600LL; }();
    static const long long t64_min = [] { const char* e = getenv("HDBO_GEMM_T64_MIN"); return e ? atoll(e) : 60LL; }();
    if (p.tile_hint == 128) return launch_gemm_cfg<G128>(p, batch, stream);
    if (p.tile_hint == 64) return launch_gemm_cfg<G64>(p, batch, stream);
    if (p.tile_hint == 32) return launch_gemm_cfg<G32>(p, batch, stream);
    if (t128 >= t128_min) return launch_gemm_cfg<G128>(p, batch, stream);
    if (t128 >= t64_min) return launch_gemm_cfg<G64>(p, batch, stream);
    return launch_gemm_cfg<G32>(p, batch, stream);
}

// --------------------------------------------------------------------------------------------------------------
// All rank-256 updates that one PAIR of columns (k0, k0+1) of the pipelined factorisation (kernels_chol_pipe.cu) makes available, as
// persistent launches over one job list (kblocks = 8: a single trailing column, K = 128):
//   bulk     A(i, j)    -= L(i, k0..) L(j, k0..)^T          n0 <= j <= i < T, except the next pair's diagonal block (the chain's)
//   inverse  W^T(j, i)  -= L^-T(j, k0..) L(i, k0..)^T       j <= jmax, n0 <= i < T  (running rows of L^-1, transposed, upper tiles of Awork)
//   K^-1     Kinv(a, b) += L^-T(a, k0..) L^-T(b, k0..)^T    b <= a <= jmax          (MLL gradient only)
// The three loads add up to T^2/2 + T/2 - 3 tile updates per pair.  What was measured on the way here (profiles/r02_fit_trace_*.jsonl,
// r02_pipe_prof_*.jsonl): as separate launches on separate streams they fought the chain's one-CTA leaf for SMs (a leaf needs a whole
// SM's shared memory: 10-40 us of queueing per step); as K = 128 updates they are L2 / HBM-bound, not DMMA-bound -- 8 flop per byte of
// tile traffic, 24 us per 4.2 MFLOP tile = 65 % of the SM's DMMA rate, and three independent 64x64 pipelines per SM did not change
// that -- hence K = 256 (pairs of columns: 16 flop per byte).  The grid is capped below the SM count: one 122 KB CTA per SM, so the
// remaining SMs stay free for the chain; every CTA walks a strided list of jobs, C is read into the accumulators before each mainloop.
// Job order: the jobs on the NEXT pair's tile columns / rows and on the diagonal block of the pair after it -- the "head" -- come
// first, so that the launcher can cut the list after one round of the grid (launch A = jobs [0, cut), launch B = the rest) and record an
// event between the two: the next pair's panel and the chain's diagonal-block update wait for the head only, not for the whole list.
__global__ void __launch_bounds__(NTHREADS, 1) k_side_updates(SideParams p) {
    extern __shared__ __align__(16) double smem[];
    const size_t ld = (size_t)p.ld;
    auto tile = [&](double* base, int r, int c) { return base + (size_t)r * 128 * ld + (size_t)c * 128; };
    const int mb = p.T - p.f;                                    // far tile columns (lower triangle of mb x mb tiles from f)
    const int e_a = p.n_a, e_s = e_a + p.n_s, e_c = e_s + p.n_c, e_e = e_c + p.n_e, e_b = e_e + p.n_b, e_d = e_b + p.n_d;
    for (int job = p.job0 + blockIdx.x; job < p.job1; job += gridDim.x) {
        const double *A, *B; double* C; double sign;
        double* Ct = nullptr;                                      // region s: C = A B^T (no accumulation) + transposed copy
        bool first = false;                                        // first update of this tile ever: nothing to read (no memset needed)
        int kb = p.kblocks;
        if (job < e_a) {
            const int j = p.n0 + job % p.nn, i = p.f + job / p.nn;
            A = tile(p.L, i, p.k0); B = tile(p.L, j, p.k0); C = tile(p.Aw, i, j); sign = -1.0;
        } else if (job < e_s) {
            const int q = job - e_a, r = q % p.w, j = q / p.w;
            A = tile(p.Li, p.k0 + r, p.k0); B = tile(p.Aw, j, p.k0); C = tile(p.Li, p.k0 + r, j); Ct = tile(p.Lt, j, p.k0 + r);
            sign = 1.0; kb = (r + 1) * KB_PER_TILE;
        } else if (job < e_c) {
            // column-major over the lower triangle (column f first): the row-major decode of the reversed index, mirrored
            int bi, bj; lower_tile_decode(p.n_c - 1 - (job - e_s), bi, bj);
            const int i = p.f + (mb - 1 - bj), j = p.f + (mb - 1 - bi);
            A = tile(p.L, i, p.k0); B = tile(p.L, j, p.k0); C = tile(p.Aw, i, j); sign = -1.0;
        } else if (job < e_e) {
            int a, b; lower_tile_decode(job - e_c, a, b);
            A = tile(p.Lt, a, p.e_k0); B = tile(p.Lt, b, p.e_k0); C = tile(p.Kinv, a, b); sign = 1.0; kb = p.e_kblocks;
            first = a >= p.e_k0;
        } else if (job < e_b) {
            // (j, i): j fastest so that neighbouring CTAs share B = L(i, k0..)
            const int q = job - e_e, j = q % (p.jmax + 1), i = p.n0 + q / (p.jmax + 1);
            A = tile(p.Lt, j, p.k0); B = tile(p.L, i, p.k0); C = tile(p.Aw, j, i); sign = -1.0; first = j >= p.k0;
        } else if (job < e_d) {
            const int q = job - e_b, j = q % (p.jmax + 1), i = p.f + q / (p.jmax + 1);
            A = tile(p.Lt, j, p.k0); B = tile(p.L, i, p.k0); C = tile(p.Aw, j, i); sign = -1.0; first = j >= p.k0;
        } else {
            int a, b; lower_tile_decode(job - e_d, a, b);
            A = tile(p.Lt, a, p.e2_k0); B = tile(p.Lt, b, p.e2_k0); C = tile(p.Kinv, a, b); sign = 1.0; kb = p.e2_kblocks;
            first = a >= p.e2_k0;
        }
        Frag acc;
        if (Ct || first) acc.zero();
        else {
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const double2 c = *reinterpret_cast<const double2*>(C + (size_t)frag_row(mt) * ld + frag_col(nt));
                    acc.v[mt][nt][0] = sign * c.x; acc.v[mt][nt][1] = sign * c.y;
                }
        }
        gemm_nt_mainloop(A, p.ld, B, p.ld, 0, kb, smem, acc);
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const double2 o = make_double2(sign * acc.v[mt][nt][0], sign * acc.v[mt][nt][1]);
                *reinterpret_cast<double2*>(C + (size_t)frag_row(mt) * ld + frag_col(nt)) = o;
                if (Ct) { Ct[(size_t)frag_col(nt) * ld + frag_row(mt)] = o.x; Ct[(size_t)(frag_col(nt) + 1) * ld + frag_row(mt)] = o.y; }
            }
    }
}

// jobs [p.job0, p.job1) of the pair's list on at most max_ctas CTAs
cudaError_t launch_side_updates(const SideParams& p, int max_ctas, cudaStream_t stream) {
    const int count = p.job1 - p.job0;
    if (count <= 0) return cudaSuccess;
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_side_updates, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    k_side_updates<<<count < max_ctas ? count : max_ctas, NTHREADS, GEMM_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K*(Z, X) tile: r2 = |z|^2 + |x|^2 - 2 z.x (clamped at 0), k(r2) * outputscale, masked beyond n; plus the partial
// dot with alpha over this tile's 128 columns (posterior mean), written to meanpart[bn][row].
template <int KIND>
__global__ void __launch_bounds__(NTHREADS, 1) k_cross(CrossParams p) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.x, bn = blockIdx.y, b = blockIdx.z;
    Frag acc; acc.zero();
    gemm_nt_mainloop(p.Zs + b * p.bsZ + (size_t)bm * BM * p.ldz, p.ldz, p.Xs + b * p.bsX + (size_t)bn * BN * p.ldx, p.ldx,
                     0, p.kblocks, smem, acc);
    const double os = p.hyp[b * 4 + 0];
    const double* zn = p.zn + b * p.bszn + bm * BM;
    const double* xn = p.xn + b * p.bsxn + bn * BN;
    const double* al = p.alpha + b * p.bsalpha + bn * BN;
    double* Kst = p.Kstar + b * p.bsK;
    double* DK = p.DK ? p.DK + b * p.bsK : nullptr;
    double xnv[4][2], alv[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) { int c = frag_col(nt) + e; xnv[nt][e] = xn[c]; alv[nt][e] = al[c]; }
    double rowval[8];
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int r = frag_row(mt);
        const double znr = zn[r];
        double s = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int c = frag_col(nt);
            double kv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                double r2 = fmax((znr + xnv[nt][e]) - 2.0 * acc.v[mt][nt][e], 0.0);
                double k, dk;
                kernel_eval(KIND, r2, k, dk);
                bool valid = ((bn * BN + c + e) < p.n) && ((bm * BM + r) < p.m);   // padding rows / columns are exact zeros
                kv[e] = valid ? os * k : 0.0;
                dv[e] = valid ? os * dk : 0.0;
                s += kv[e] * alv[nt][e];
            }
            {
                size_t off = (size_t)(bm * BM + r) * p.ldk + bn * BN + c;
                *reinterpret_cast<double2*>(Kst + off) = make_double2(kv[0], kv[1]);
                if (DK) *reinterpret_cast<double2*>(DK + off) = make_double2(dv[0], dv[1]);
            }
        }
        rowval[mt] = s;
    }
    double r = tile_row_reduce(rowval, smem);
    if (threadIdx.x < 128) p.meanpart[b * p.bsmp + (size_t)bn * (p.mtiles * BM) + bm * BM + threadIdx.x] = r;
}

cudaError_t launch_cross(const CrossParams& p, int batch, cudaStream_t stream) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_cross<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cross<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    dim3 grid(p.mtiles, p.ntiles, batch);
    if (p.kind == 0) k_cross<0><<<grid, NTHREADS, GEMM_SMEM_BYTES, stream>>>(p);
    else k_cross<1><<<grid, NTHREADS, GEMM_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// Train covariance, lower tiles only: Khat = outputscale*k(r2) + (noise + jitter) on the diagonal; rows/cols >= n are
// the identity so that the padded matrix stays SPD and its factor / inverse are block-diagonal with I.  When the MLL gradient is
// wanted, the `R2` state buffer gets dk/dr^2 (unit outputscale; padding 0) on the SAME lower tiles -- the only other function of the
// distances the gradient needs, already at hand here (the first version stored r^2, mirrored with strided 8-byte stores, and
// the gradient kernel evaluated exp / sqrt a second time for all n^2 entries).
// Templated on the tile configuration: few-tile problems (the small n of a BO loop: 6 tiles of 128x128 at n_pad = 384) run on
// 32x32 / 64x64 CTA tiles so that the FP64 kernel epilogue spreads over the SMs (p.tiles counts tiles of G::TM).
// (KIND is a template parameter: with a run-time kind both kernel functions were inlined behind selects for each of the 64 elements
// of a thread -- 15.6 k SASS instructions, 55 % of the tile time in an instruction-cache-bound epilogue, ncu r02_ncu_fit2)
template <class G, int KIND>
__global__ void __launch_bounds__(G::THREADS) k_sym_build(SymBuildParams p) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.z;
    const int t_all = p.tiles * (p.tiles + 1) / 2, t_end = p.t1 > 0 ? p.t1 : t_all;
    for (int t = p.t0 + blockIdx.x; t < t_end; t += gridDim.x) {
    int bm, bn;
    if (p.colmajor) { int bi, bj; lower_tile_decode(t_all - 1 - t, bi, bj); bm = p.tiles - 1 - bj; bn = p.tiles - 1 - bi; }
    else lower_tile_decode(t, bm, bn);
    typename G::Frag acc; acc.zero();
    const double* Xs = p.Xs + b * p.bsX;
    G::mainloop(Xs + (size_t)bm * G::TM * p.ldx, p.ldx, Xs + (size_t)bn * G::TN * p.ldx, p.ldx, 0, p.kblocks, smem, acc);
    const double os = p.hyp[b * 4 + 0], noise = p.hyp[b * 4 + 1], jitter = p.hyp[b * 4 + 3];
    const double* xn = p.xn + b * p.bsxn;
    double* Kh = p.Khat + b * p.bsK;
    double* R2 = p.R2 ? p.R2 + b * p.bsK : nullptr;
    // interior tiles (off the diagonal, inside the n x n block: all but ~2 T of the T^2 / 2) take a path without the per-element
    // diagonal / padding selects
    const bool interior = bm != bn && (bm + 1) * G::TM <= p.n;
    double xnj[G::NT][2];
#pragma unroll
    for (int nt = 0; nt < G::NT; nt++) { xnj[nt][0] = xn[bn * G::TN + G::frag_col(nt)]; xnj[nt][1] = xn[bn * G::TN + G::frag_col(nt) + 1]; }
    if (interior) {
#pragma unroll
        for (int mt = 0; mt < G::MT; mt++) {
            const int i = bm * G::TM + G::frag_row(mt);
            const double xi = xn[i];
            double* krow = Kh + (size_t)i * p.ldk + bn * G::TN;
            double* rrow = R2 ? R2 + (size_t)i * p.ldk + bn * G::TN : nullptr;
#pragma unroll
            for (int nt = 0; nt < G::NT; nt++) {
                double k0, d0, k1, d1;
                kernel_eval(KIND, fmax((xi + xnj[nt][0]) - 2.0 * acc.v[mt][nt][0], 0.0), k0, d0);
                kernel_eval(KIND, fmax((xi + xnj[nt][1]) - 2.0 * acc.v[mt][nt][1], 0.0), k1, d1);
                *reinterpret_cast<double2*>(krow + G::frag_col(nt)) = make_double2(os * k0, os * k1);
                if (rrow) *reinterpret_cast<double2*>(rrow + G::frag_col(nt)) = make_double2(d0, d1);
            }
        }
    } else {
#pragma unroll
    for (int mt = 0; mt < G::MT; mt++) {
        const int i = bm * G::TM + G::frag_row(mt);
        const double xi = xn[i];
#pragma unroll
        for (int nt = 0; nt < G::NT; nt++) {
            const int j0 = bn * G::TN + G::frag_col(nt);
            double kv[2], rv[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = j0 + e;
                double r2 = fmax((xi + xnj[nt][e]) - 2.0 * acc.v[mt][nt][e], 0.0);
                if (i == j) r2 = 0.0;
                double k, dk;
                kernel_eval(KIND, r2, k, dk);
                const bool valid = (i < p.n) && (j < p.n);
                kv[e] = valid ? os * k + ((i == j) ? (noise + jitter) : 0.0) : ((i == j) ? 1.0 : 0.0);
                rv[e] = valid ? dk : 0.0;
            }
            *reinterpret_cast<double2*>(Kh + (size_t)i * p.ldk + j0) = make_double2(kv[0], kv[1]);
            if (R2) *reinterpret_cast<double2*>(R2 + (size_t)i * p.ldk + j0) = make_double2(rv[0], rv[1]);
        }
    }
    }
    }
}

template <class G, int KIND>
static cudaError_t launch_sym_build_kind(SymBuildParams p, int batch, cudaStream_t stream) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_sym_build<G, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    p.tiles *= 128 / G::TM;
    int count = (p.t1 > 0 ? p.t1 : p.tiles * (p.tiles + 1) / 2) - p.t0;
    if (count <= 0) return cudaSuccess;
    if (p.max_ctas > 0 && count > p.max_ctas) count = p.max_ctas;
    dim3 grid(count, 1, batch);
    k_sym_build<G, KIND><<<grid, G::THREADS, G::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}
template <class G>
static cudaError_t launch_sym_build_cfg(SymBuildParams p, int batch, cudaStream_t stream) {
    return p.kind == 0 ? launch_sym_build_kind<G, 0>(p, batch, stream) : launch_sym_build_kind<G, 1>(p, batch, stream);
}

cudaError_t launch_sym_build(const SymBuildParams& p, int batch, cudaStream_t stream) {
    const long long t128 = (long long)p.tiles * (p.tiles + 1) / 2 * batch;
    if (t128 >= 148 || p.t1 > 0 || p.t0 > 0) return launch_sym_build_cfg<G128>(p, batch, stream);   // (tile ranges count 128-tiles)
    if (t128 >= 37) return launch_sym_build_cfg<G64>(p, batch, stream);
    return launch_sym_build_cfg<G32>(p, batch, stream);
}

// --------------------------------------------------------------------------------------------------------------
// V = K* L^-T : V[c,i] = sum_{j<=i} K*[c,j] Linv[i,j].  Column tile bn only needs k-tiles <= bn (n^2 flops per candidate,
// not 2n^2).  Heavy column tiles are scheduled first (blockIdx.y = 0 -> last column tile).  Epilogue: per-row sum of
// squares over this tile's 128 columns -> varpart[bn][row]; V itself is stored only for the gradient path.
// Mainloop configuration: same 2x4 warp grid / fragment layout as G128, but 32-wide k-blocks (3 x 72 KB ring = 221 KB):
// half the barriers; measured best of the variants in tools/var_bench.cu (35.0 vs 34.5 TFLOP/s for BK = 16).
using GVAR = G128W;
__global__ void __launch_bounds__(NTHREADS, 1) k_var(VarParams p) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.x, bn = p.ntiles - 1 - blockIdx.y, b = blockIdx.z;
    GVAR::Frag gacc; gacc.zero();
    GVAR::mainloop(p.Kstar + b * p.bsK + (size_t)bm * BM * p.ldk, p.ldk, p.Linv + b * p.bsL + (size_t)bn * BN * p.ldl,
                   p.ldl, 0, (bn + 1) * GVAR::KB_PER_TN, smem, gacc);
    auto& acc = gacc;
    double rowval[8];
    double* V = p.V ? p.V + b * p.bsV : nullptr;
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        double s = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            s = fma(acc.v[mt][nt][0], acc.v[mt][nt][0], s);
            s = fma(acc.v[mt][nt][1], acc.v[mt][nt][1], s);
            if (V)
                *reinterpret_cast<double2*>(V + (size_t)(bm * BM + frag_row(mt)) * p.ldv + bn * BN + frag_col(nt)) =
                    make_double2(acc.v[mt][nt][0], acc.v[mt][nt][1]);
        }
        rowval[mt] = s;
    }
    double r = tile_row_reduce(rowval, smem);
    if (threadIdx.x < 128) p.varpart[b * p.bsvp + (size_t)bn * (p.mtiles * BM) + bm * BM + threadIdx.x] = r;
}

cudaError_t launch_var(const VarParams& p, int batch, cudaStream_t stream) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_var, cudaFuncAttributeMaxDynamicSharedMemorySize, GVAR::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    dim3 grid(p.mtiles, p.ntiles, batch);
    k_var<<<grid, NTHREADS, GVAR::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hdbo
