// Gradient-path kernels.
//   MLL gradient (A3):        K^-1 = L^-T L^-1 (SYRK in the DMMA mainloop), G = (alpha alpha^T - K^-1) o (-2 os dk/dr2),
//                             lengthscale gradient from one n x n x d GEMM with a column-dot epilogue, scalar traces by
//                             warp-shuffle + block reductions.
//   posterior backward (A6):  dV -> dK* = dV L^-1 (triangular GEMM) -> C = dK* o dK*/dr2 -> dZ = 2/l o (z rowsum(C) - C Xs).
//   joint q x q covariance (A7) and its pairwise gradient term.
#include "gemm_core.cuh"
#include "kernels.cuh"
#include "elem.cuh"
#include "grad.cuh"

namespace hdbo {

__device__ __forceinline__ double warp_sum_g(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ---------------------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c]  (in: [rows][ld_in] using cols columns; out: [cols_pad][ld_out], zero padded rows >= cols)
__global__ void k_transpose(const double* __restrict__ in, int ld_in, long long bs_in, int rows, int cols,
                            double* __restrict__ out, int ld_out, long long bs_out, int cols_pad) {
    __shared__ double t[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        int r = r0 + i, c = c0 + tx;
        t[i][tx] = (r < rows && c < cols) ? in[b * bs_in + (size_t)r * ld_in + c] : 0.0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, r = r0 + tx;
        if (c < cols_pad && r < ld_out) out[b * bs_out + (size_t)c * ld_out + r] = t[tx][i];
    }
}

cudaError_t launch_transpose(const double* in, int ld_in, long long bs_in, int rows, int cols, double* out, int ld_out,
                             long long bs_out, int cols_pad, int batch, cudaStream_t st) {
    dim3 grid((ld_out + 31) / 32, (cols_pad + 31) / 32, batch);
    k_transpose<<<grid, 256, 0, st>>>(in, ld_in, bs_in, rows, cols, out, ld_out, bs_out, cols_pad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// G[i][j] = (alpha_i alpha_j - Kinv_ij) * (-2 os DK_ij) for i,j < n, else 0.  Kinv and DK (= dk/dr^2 for unit outputscale, left by the
// kernel build) hold lower tiles only; the upper half is read transposed through shared memory.  HBM-bound: no kernel evaluation here
// (the first version re-derived k and dk from the squared distances: 16.7 M FP64 exp / sqrt at n = 4096).
// Grid: (n_pad/32 row stripes, NSEG column segments, batch).  Per (stripe, segment): partial row sums of G and partial tr(W);
// sum(W o K), the other trace the gradient needs, follows from  sum(W o K) = r^T alpha - n - noise tr(W)  (Khat alpha = r).
constexpr int G_NSEG = 4;
__global__ void __launch_bounds__(256) k_mll_G(const double* __restrict__ Kinv, const double* __restrict__ DK, int ld,
                                               long long bsM, const double* __restrict__ alpha, long long bsv,
                                               const double* __restrict__ hyp, int n, int n_pad,
                                               double* __restrict__ G, double* __restrict__ rowpart /*[B][NSEG][n_pad]*/,
                                               double* __restrict__ scalpart /*[B][stripes*NSEG][2]*/) {
    __shared__ double t[32][33], td[32][33];
    __shared__ double red[8];
    const int b = blockIdx.z, stripe = blockIdx.x, seg = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double m2os = -2.0 * hyp[b * 4 + 0];
    const double* Ki = Kinv + b * bsM;
    const double* Dk = DK + b * bsM;
    const double* al = alpha + b * bsv;
    double* Gg = G + b * bsM;
    const int ntile = n_pad / 32;
    const int per = (ntile + G_NSEG - 1) / G_NSEG;
    const int t0 = seg * per, t1 = min(ntile, t0 + per);
    double rs[4] = {0, 0, 0, 0};
    double trw = 0.0;
    double ai[4];
#pragma unroll
    for (int r = 0; r < 4; r++) ai[r] = al[stripe * 32 + warp * 4 + r];
    for (int tj = t0; tj < t1; tj++) {
        const bool upper = tj > stripe;
        if (upper) {
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const size_t o = (size_t)(tj * 32 + warp * 4 + r) * ld + stripe * 32 + lane;
                t[warp * 4 + r][lane] = Ki[o];
                td[warp * 4 + r][lane] = Dk[o];
            }
            __syncthreads();
        }
        const int j = tj * 32 + lane;
        const double aj = al[j];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = stripe * 32 + warp * 4 + r;
            double kin, dk;
            if (upper) { kin = t[lane][warp * 4 + r]; dk = td[lane][warp * 4 + r]; }
            else if (tj < stripe || lane <= warp * 4 + r) { kin = Ki[(size_t)i * ld + j]; dk = Dk[(size_t)i * ld + j]; }
            else { kin = Ki[(size_t)j * ld + i]; dk = Dk[(size_t)j * ld + i]; }   // upper part of a diagonal 32x32 tile (stored tiles are full)
            double g = 0.0;
            if (i < n && j < n) {
                const double W = ai[r] * aj - kin;
                g = W * (m2os * dk);
                if (i == j) trw += W;
            }
            Gg[(size_t)i * ld + j] = g;
            rs[r] += g;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        double s = warp_sum_g(rs[r]);
        if (lane == 0) rowpart[((size_t)b * G_NSEG + seg) * n_pad + stripe * 32 + warp * 4 + r] = s;
    }
    trw = warp_sum_g(trw);
    __syncthreads();
    if (lane == 0) red[warp] = trw;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0;
        for (int w = 0; w < 8; w++) c += red[w];
        size_t o = ((size_t)b * gridDim.x * G_NSEG + (size_t)stripe * G_NSEG + seg) * 2;
        scalpart[o] = 0.0; scalpart[o + 1] = c;
    }
}

// T = G * Xs (as NT GEMM with B = Xs^T), epilogue: colpart[bm][k] = sum_{i in tile} Xs[i][k] (Xs[i][k] s_i - T[i][k])
struct GxaParams {
    const double* G; const double* XsT; const double* Xs; const double* rowpart; double* colpart;
    int ldg, ldt, ldx, n_pad, mtiles, ntiles, splits;   // split-K: blockIdx.y = bn + ntiles * split
    long long bsG, bsT, bsX, bsrow, bscol;
};
__global__ void __launch_bounds__(NTHREADS, 1) k_gxa(GxaParams p) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.x, bn = blockIdx.y % p.ntiles, sp = blockIdx.y / p.ntiles, b = blockIdx.z;
    const int kbt = p.n_pad / BK;
    Frag acc; acc.zero();
    gemm_nt_mainloop(p.G + b * p.bsG + (size_t)bm * BM * p.ldg, p.ldg, p.XsT + b * p.bsT + (size_t)bn * BN * p.ldt, p.ldt,
                     (int)((long long)kbt * sp / p.splits), (int)((long long)kbt * (sp + 1) / p.splits), smem, acc);
    const double* Xs = p.Xs + b * p.bsX;
    const double* rp = p.rowpart + b * p.bsrow;
    double colval[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int i = bm * BM + frag_row(mt);
        double s = 0.0;
        if (sp == 0) {   // the x^2 s_i term is added once, by the first k-split
#pragma unroll
            for (int sg = 0; sg < G_NSEG; sg++) s += rp[(size_t)sg * p.n_pad + i];
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = bn * BN + frag_col(nt) + e;
                const double x = (k < p.ldx) ? Xs[(size_t)i * p.ldx + k] : 0.0;
                colval[nt][e] += x * (x * s - acc.v[mt][nt][e]);
            }
    }
    double r = tile_col_reduce(colval, smem);
    if (threadIdx.x < 128)
        p.colpart[b * p.bscol + (size_t)(bm * p.splits + sp) * (p.ntiles * BN) + bn * BN + threadIdx.x] = r;
}

constexpr int GXA_MAX_SPLITS = 4;

__global__ void k_mll_grad_final(const double* __restrict__ colpart, long long bscol, int mtiles, int dpt,
                                 const double* __restrict__ scalpart, int nscal, const double* __restrict__ inv_ls,
                                 const double* __restrict__ hyp, const double* __restrict__ scal, int d, int n_true,
                                 double* __restrict__ grad) {
    const int b = blockIdx.x;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < mtiles; t++) s += colpart[b * bscol + (size_t)t * dpt + k];
        grad[(size_t)b * (d + 3) + k] = s * inv_ls[(size_t)b * d + k];
    }
    // the two trace-like scalars: block reduction over the nscal partial pairs (one thread summing them serially cost 80 us at
    // n = 4096: 512 dependent global loads)
    __shared__ double red[2][8];
    double a = 0.0, c = 0.0;
    for (int t = threadIdx.x; t < nscal; t += blockDim.x) { a += scalpart[((size_t)b * nscal + t) * 2]; c += scalpart[((size_t)b * nscal + t) * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0; c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; c += red[1][w]; }
        // sum(W o K) over the n x n block:  alpha^T K alpha - tr(Khat^-1 K)  with  K = Khat - s I  (s = noise + jitter), Khat alpha = r
        //   = r^T alpha - s alpha^T alpha - n + s tr(Khat^-1) = quad - n - s tr(W)
        a = scal[b * 8 + 1] - (double)n_true - (hyp[b * 4 + 1] + hyp[b * 4 + 3]) * c;
        grad[(size_t)b * (d + 3) + d] = 0.5 * a / hyp[b * 4 + 0];
        grad[(size_t)b * (d + 3) + d + 1] = 0.5 * c;
        grad[(size_t)b * (d + 3) + d + 2] = scal[b * 8 + 3];
    }
}

// K^-1 = L^-T L^-1 = U U^T (U = L^-T, upper): lower tiles, k-tile >= m-tile
cudaError_t kinv_from_linvt(const hdbo_gp* gp, double* Kinv, cudaStream_t st) {
    const int np = gp->n_pad, T = np / 128;
    const long long nn = (long long)np * np;
    GemmParams g{};
    g.A = gp->Linvt; g.B = gp->Linvt; g.C = Kinv; g.Ct = nullptr;
    g.lda = g.ldb = g.ldc = np; g.bsA = g.bsB = g.bsC = nn;
    g.mtiles = g.ntiles = T; g.kblocks = np / BK; g.krange = KR_GE_M; g.lower_only = 1; g.alpha = 1.0; g.beta = 0.0;
    return launch_gemm_nt(g, gp->batch, st);
}

cudaError_t mll_grad(const hdbo_gp* gp, double* Kinv, double* G, double* XsT, double* part, double* grad, cudaStream_t st) {
    const int np = gp->n_pad, dp = gp->d_pad, B = gp->batch, T = np / 128;
    const long long nn = (long long)np * np;
    const int dpt = (gp->d + 127) / 128 * 128;
    cudaError_t e;
    // Kinv == gp->Kinv: hdbo_gp_fit(need_r2) has already left K^-1 there (accumulated beside the factorisation when it ran pipelined)
    if (Kinv != gp->Kinv && (e = kinv_from_linvt(gp, Kinv, st)) != cudaSuccess) return e;
    // part layout per batch: rowpart [G_NSEG][np] | scalpart [np/32*G_NSEG][2] | colpart [T][dpt]
    const long long per_b = (long long)G_NSEG * np + (long long)(np / 32) * G_NSEG * 2 + (long long)T * dpt;
    double* rowpart = part;
    double* scalpart = part + (long long)B * G_NSEG * np;
    double* colpart = scalpart + (long long)B * (np / 32) * G_NSEG * 2;
    (void)per_b;
    dim3 gg(np / 32, G_NSEG, B);
    k_mll_G<<<gg, 256, 0, st>>>(Kinv, gp->R2, np, nn, gp->alpha, np, gp->hyp, gp->n, np, G, rowpart, scalpart);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_transpose(gp->Xs, dp, (long long)np * dp, np, dp, XsT, np, (long long)dpt * np, dpt, B, st)) != cudaSuccess)
        return e;
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        if ((e = cudaFuncSetAttribute(k_gxa, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES)) != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    GxaParams x{};
    x.G = G; x.XsT = XsT; x.Xs = gp->Xs; x.rowpart = rowpart; x.colpart = colpart;
    x.ldg = np; x.ldt = np; x.ldx = dp; x.n_pad = np; x.mtiles = T; x.ntiles = dpt / 128;
    // split-K factor: minimise (waves of 148 CTAs) x (k-blocks per CTA); e.g. 64 tiles: 2 splits = one 128-CTA wave of half the
    // k-range, whereas ceil(148 / 64) = 3 splits would run a second, 30 %-full wave
    const int tiles = T * (dpt / 128) * B, kblocks = np / BK;
    int splits = 1;
    long long best = -1;
    for (int sp = 1; sp <= GXA_MAX_SPLITS && sp <= kblocks; sp++) {
        const long long cost = (long long)((tiles * sp + 147) / 148) * ((kblocks + sp - 1) / sp);
        if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
    x.splits = splits;
    x.bsG = nn; x.bsT = (long long)dpt * np; x.bsX = (long long)np * dp; x.bsrow = (long long)G_NSEG * np;
    x.bscol = (long long)T * GXA_MAX_SPLITS * dpt;
    k_gxa<<<dim3(T, (dpt / 128) * splits, B), NTHREADS, GEMM_SMEM_BYTES, st>>>(x);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_mll_grad_final<<<B, 256, 0, st>>>(colpart, (long long)T * GXA_MAX_SPLITS * dpt, T * splits, dpt, scalpart, (np / 32) * G_NSEG, gp->inv_ls, gp->hyp,
                                        gp->scal, gp->d, gp->n, grad);
    return cudaGetLastError();
}

size_t mll_grad_part_doubles(const hdbo_gp* gp) {
    const size_t np = gp->n_pad, B = gp->batch, T = np / 128, dpt = (gp->d + 127) / 128 * 128;
    return B * (G_NSEG * np + (np / 32) * G_NSEG * 2 + T * GXA_MAX_SPLITS * dpt);
}

// ---------------------------------------------------------------------------------------------------------------
// posterior backward pieces
// dV[c][:] = -sum_b S[a][b] V[b][:], S = gSigma + gSigma^T within the q-group of row c (q == 1: S = 2 gvar).
__global__ void k_make_dV(const double* __restrict__ V, int ld, long long bsV, const double* __restrict__ gvar,
                          const double* __restrict__ gSigma, long long bsg, long long bsS, int m, int q, int rows_pad,
                          double* __restrict__ dV) {
    const int b = blockIdx.y;
    const int c = blockIdx.x;                       // one block per candidate row (including padding rows)
    double* o = dV + b * bsV + (size_t)c * ld;
    if (c >= m) { for (int j = threadIdx.x; j < ld; j += blockDim.x) o[j] = 0.0; return; }
    const int g = c / q, a = c - g * q;
    double S[8];
    for (int bb = 0; bb < q; bb++) {
        if (gSigma) S[bb] = gSigma[b * bsS + ((size_t)g * q + a) * q + bb] + gSigma[b * bsS + ((size_t)g * q + bb) * q + a];
        else S[bb] = 2.0 * gvar[b * bsg + c];
    }
    const double* Vg = V + b * bsV + (size_t)g * q * ld;
    for (int j = threadIdx.x; j < ld; j += blockDim.x) {
        double s = 0.0;
        for (int bb = 0; bb < q; bb++) s = fma(S[bb], Vg[(size_t)bb * ld + j], s);
        o[j] = -s;
    }
}

// C[c][j] = (gmean_c alpha_j + Bm[c][j]) * DK[c][j] in place over Bm; csum[c] = sum_j C[c][j].  One CTA per row, 16-byte accesses
// (the first version -- one warp per row, 8-byte accesses -- left 512-point calls on 64 CTAs: 64 us for 50 MB of traffic).
// add != nullptr: columns < add_cols of Bm first receive the second parts of a split-K product (add[row][col], leading dimension ldadd).
__global__ void __launch_bounds__(256) k_make_C(double* __restrict__ Bm, const double* __restrict__ DK, int ld, long long bsM,
                         const double* __restrict__ gmean, long long bsg, const double* __restrict__ alpha, long long bsa,
                         int m, int rows_pad, double* __restrict__ csum, long long bsc,
                         const double* __restrict__ add, int ldadd, int add_cols) {
    __shared__ double red[8];
    const int b = blockIdx.y, c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* row = reinterpret_cast<double2*>(Bm + b * bsM + (size_t)c * ld);
    const double2* dk = reinterpret_cast<const double2*>(DK + b * bsM + (size_t)c * ld);
    const double2* al = reinterpret_cast<const double2*>(alpha + b * bsa);
    const double2* ad = add ? reinterpret_cast<const double2*>(add + (size_t)c * ldadd) : nullptr;
    const bool live = c < m;
    const double gm = (live && gmean) ? gmean[b * bsg + c] : 0.0;
    double s = 0.0;
    for (int j = threadIdx.x; j < ld / 2; j += 256) {
        double2 v = make_double2(0.0, 0.0);
        if (live) {
            double2 r = row[j];
            const double2 k = dk[j], a = al[j];
            if (ad && 2 * j < add_cols) { const double2 x = ad[j]; r.x += x.x; r.y += x.y; }
            v.x = (gm * a.x + r.x) * k.x; v.y = (gm * a.y + r.y) * k.y;
        }
        row[j] = v;
        s += v.x + v.y;
    }
    s = warp_sum_g(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) csum[b * bsc + c] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// dZ[c][k] = 2 inv_ls[k] (Zs[c][k] csum[c] - (C Xs)[c][k])
struct DzParams {
    const double* Cm; const double* XsT; const double* Zs; const double* csum; const double* inv_ls; double* dZ;
    int ldc, ldt, ldz, lddz, n_pad, m, d;
    long long bsC, bsT, bsZ, bscs, bsdz;
};
__global__ void __launch_bounds__(NTHREADS, 1) k_dz(DzParams p) {
    extern __shared__ __align__(16) double smem[];
    const int bm = blockIdx.x, bn = blockIdx.y, b = blockIdx.z;
    Frag acc; acc.zero();
    gemm_nt_mainloop(p.Cm + b * p.bsC + (size_t)bm * BM * p.ldc, p.ldc, p.XsT + b * p.bsT + (size_t)bn * BN * p.ldt, p.ldt, 0,
                     p.n_pad / BK, smem, acc);
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int c = bm * BM + frag_row(mt);
        if (c >= p.m) continue;
        const double cs = p.csum[b * p.bscs + c];
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = bn * BN + frag_col(nt) + e;
                if (k < p.d)
                    p.dZ[b * p.bsdz + (size_t)c * p.lddz + k] =
                        2.0 * p.inv_ls[(size_t)b * p.d + k] * (p.Zs[b * p.bsZ + (size_t)c * p.ldz + k] * cs - acc.v[mt][nt][e]);
            }
    }
}

// the same epilogue on a product P = C Xs that the generic GEMM already formed (few-candidate path of posterior_bwd)
// (P may arrive as `nsplit` split-K partial products `split_stride` doubles apart: summed here in a fixed order)
__global__ void k_dz_finish(const double* __restrict__ P, int ldp, long long bsP, int nsplit, long long split_stride,
                            const double* __restrict__ Zs, int ldz, long long bsZ,
                            const double* __restrict__ csum, long long bscs, const double* __restrict__ inv_ls, int m, int d,
                            double* __restrict__ dZ, int lddz, long long bsdz) {
    const int b = blockIdx.y;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * d) return;
    const int c = (int)(idx / d), k = (int)(idx - (long long)c * d);
    const double* pp = P + b * bsP + (size_t)c * ldp + k;
    double pv = pp[0];
    for (int sidx = 1; sidx < nsplit; sidx++) pv += pp[sidx * split_stride];
    dZ[b * bsdz + (size_t)c * lddz + k] =
        2.0 * inv_ls[(size_t)b * d + k] * (Zs[b * bsZ + (size_t)c * ldz + k] * csum[b * bscs + c] - pv);
}

// joint covariance of each q-group: Sigma[a][b] = os k(r2(z_a,z_b)) - V_a.V_b (+ noise on the diagonal).  One block per group.
__global__ void k_joint_cov(const double* __restrict__ Zs, int ldz, long long bsZ, const double* __restrict__ V, int ldv,
                            long long bsV, const double* __restrict__ hyp, int q, int kind, int observation_noise,
                            double* __restrict__ Sigma, long long bsS) {
    const int b = blockIdx.y, g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double os = hyp[b * 4 + 0], noise = hyp[b * 4 + 1];
    for (int pr = warp; pr < q * q; pr += nw) {
        const int a = pr / q, c = pr - a * q;
        if (c > a) continue;
        const double* va = V + b * bsV + (size_t)(g * q + a) * ldv;
        const double* vc = V + b * bsV + (size_t)(g * q + c) * ldv;
        double dot = 0.0;
        for (int j = lane; j < ldv; j += 32) dot = fma(va[j], vc[j], dot);
        dot = warp_sum_g(dot);
        double r2 = 0.0;
        if (a != c) {
            const double* za = Zs + b * bsZ + (size_t)(g * q + a) * ldz;
            const double* zc = Zs + b * bsZ + (size_t)(g * q + c) * ldz;
            for (int k = lane; k < ldz; k += 32) { double t = za[k] - zc[k]; r2 = fma(t, t, r2); }
            r2 = warp_sum_g(r2);
        }
        if (lane == 0) {
            double k, dk;
            kernel_eval(kind, r2, k, dk);
            double s = os * k - dot;
            if (a == c && observation_noise) s += noise;
            Sigma[b * bsS + ((size_t)g * q + a) * q + c] = s;
            Sigma[b * bsS + ((size_t)g * q + c) * q + a] = s;
        }
    }
}

// dZ[a] += sum_{b != a} (gS_ab + gS_ba) os dk/dr2(r2_ab) 2 (za - zb) inv_ls   (k(Zq,Zq) term of the joint covariance)
__global__ void k_dz_pairs(const double* __restrict__ Zs, int ldz, long long bsZ, const double* __restrict__ gSigma,
                           long long bsS, const double* __restrict__ hyp, const double* __restrict__ inv_ls, int q, int d,
                           int kind, double* __restrict__ dZ, int lddz, long long bsdz) {
    const int b = blockIdx.y, g = blockIdx.x;
    __shared__ double coef[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double os = hyp[b * 4 + 0];
    for (int pr = warp; pr < q * q; pr += nw) {
        const int a = pr / q, c = pr - a * q;
        double cf = 0.0;
        if (a != c) {
            const double* za = Zs + b * bsZ + (size_t)(g * q + a) * ldz;
            const double* zc = Zs + b * bsZ + (size_t)(g * q + c) * ldz;
            double r2 = 0.0;
            for (int k = lane; k < ldz; k += 32) { double t = za[k] - zc[k]; r2 = fma(t, t, r2); }
            r2 = warp_sum_g(r2);
            double k, dk;
            kernel_eval(kind, r2, k, dk);
            cf = (gSigma[b * bsS + ((size_t)g * q + a) * q + c] + gSigma[b * bsS + ((size_t)g * q + c) * q + a]) * os * dk * 2.0;
        }
        if (lane == 0) coef[pr] = cf;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < q * d; e += blockDim.x) {
        const int a = e / d, k = e - a * d;
        const double za = Zs[b * bsZ + (size_t)(g * q + a) * ldz + k];
        double s = 0.0;
        for (int c = 0; c < q; c++)
            if (c != a) s = fma(coef[a * q + c], za - Zs[b * bsZ + (size_t)(g * q + c) * ldz + k], s);
        dZ[b * bsdz + (size_t)(g * q + a) * lddz + k] += s * inv_ls[(size_t)b * d + k];
    }
}

cudaError_t launch_joint_cov(const double* Zs, int ldz, long long bsZ, const double* V, int ldv, long long bsV,
                             const double* hyp, int groups, int q, int kind, int observation_noise, double* Sigma,
                             long long bsS, int batch, cudaStream_t st) {
    if (groups == 0) return cudaSuccess;
    k_joint_cov<<<dim3(groups, batch), 256, 0, st>>>(Zs, ldz, bsZ, V, ldv, bsV, hyp, q, kind, observation_noise, Sigma, bsS);
    return cudaGetLastError();
}

cudaError_t posterior_bwd(const hdbo_gp* gp, const BwdBuffers& w, int64_t rows_alloc, int64_t m, int q, const double* gmean,
                          const double* gvar, const double* gSigma, long long bsg, long long bsS, double* dZ, int lddz,
                          long long bsdz, cudaStream_t st) {
    const int np = gp->n_pad, dp = gp->d_pad, B = gp->batch;
    const int rp = (int)((m + 127) / 128 * 128), mt = rp / 128;
    const int dpt = (gp->d + 127) / 128 * 128;
    const long long bsM = (long long)rows_alloc * np;
    cudaError_t e;
    const bool have_v = (gvar != nullptr) || (gSigma != nullptr);
    const double* add = nullptr;
    // (1) dV into T, (2) Bm = dV * Linv into Kstar
    if (have_v) {
        k_make_dV<<<dim3(rp, B), 256, 0, st>>>(w.V, np, bsM, gvar, gSigma, bsg, bsS, (int)m, q, rp, w.T);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        GemmParams g{};
        g.A = w.T; g.B = gp->Linvt; g.C = w.Kstar; g.Ct = nullptr;
        g.lda = np; g.ldb = np; g.ldc = np; g.bsA = bsM; g.bsB = (long long)np * np; g.bsC = bsM;
        g.mtiles = mt; g.ntiles = np / 128; g.kblocks = np / BK; g.krange = KR_GE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        // few candidates, single GP: the long half of the columns is cut at K/2 (second parts -> P2, added back by k_make_C)
        if (B == 1 && (long long)mt * (np / 128) < 148 && (np / 128) % 2 == 0 && np / 128 >= 4 && w.P2) {
            g.krange = KR_GE_N_SPLIT; g.C2 = w.P2; g.ldc2 = np / 2; g.c2_col0 = 0; add = w.P2;
        }
        if ((e = launch_gemm_nt(g, B, st)) != cudaSuccess) return e;
    } else {
        if ((e = cudaMemsetAsync(w.Kstar, 0, sizeof(double) * (size_t)B * bsM, st)) != cudaSuccess) return e;
    }
    // (3) C and its row sums
    k_make_C<<<dim3(rp, B), 256, 0, st>>>(w.Kstar, w.DK, np, bsM, gmean, bsg, gp->alpha, np, (int)m, rp, w.csum, rows_alloc, add, np / 2, np / 2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // (4) dZ = 2/l o (z csum - C Xs)
    if ((e = launch_transpose(gp->Xs, dp, (long long)np * dp, np, dp, w.XsT, np, (long long)dpt * np, dpt, B, st)) != cudaSuccess) return e;
    if ((long long)mt * (dpt / 128) * B < 148 && dpt <= np) {
        // few candidates (optimize_acqf inner loop): 128x128 tiles would leave most SMs idle for the whole k = n range; the
        // generic GEMM cuts the product into 32x32 / 64x64 tiles, P lands in the (now free) dV scratch
        GemmParams g{};
        g.A = w.Kstar; g.B = w.XsT; g.C = w.T; g.Ct = nullptr;
        g.lda = np; g.ldb = np; g.ldc = dpt; g.bsA = bsM; g.bsB = (long long)dpt * np; g.bsC = bsM;
        g.mtiles = mt; g.ntiles = dpt / 128; g.kblocks = np / BK; g.krange = KR_FULL; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        // single GP: split K over the batch axis of the generic GEMM (the contraction index is contiguous, so a batch stride of
        // K / S doubles selects the s-th K window) -- 128 CTAs walking k = 4096 each measured 102 us, latency-bound per CTA
        int S = 1;
        const long long split_stride = (long long)mt * 128 * dpt;
        if (B == 1) {
            const long long ctas32 = (long long)mt * 4 * (dpt / 32);
            while (S < 8 && ctas32 * S < 444 && (long long)(2 * S) * dpt <= np && (np / BK) % (2 * S) == 0 && (np / BK) / (2 * S) >= 16) S *= 2;
            if (S > 1) { g.bsA = np / S; g.bsB = np / S; g.bsC = split_stride; g.kblocks = np / BK / S; }
        }
        if ((e = launch_gemm_nt(g, S > 1 ? S : B, st)) != cudaSuccess) return e;
        const long long tot = (long long)m * gp->d;
        k_dz_finish<<<dim3((unsigned)((tot + 255) / 256), B), 256, 0, st>>>(w.T, dpt, bsM, S, split_stride, w.Zs, dp, (long long)rows_alloc * dp,
                                                                             w.csum, rows_alloc, gp->inv_ls, (int)m, gp->d, dZ, lddz, bsdz);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (gSigma && q > 1) {
            k_dz_pairs<<<dim3((unsigned)(m / q), B), 256, 0, st>>>(w.Zs, dp, (long long)rows_alloc * dp, gSigma, bsS, gp->hyp, gp->inv_ls, q,
                                                                   gp->d, gp->kind, dZ, lddz, bsdz);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        if ((e = cudaFuncSetAttribute(k_dz, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES)) != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    DzParams p{};
    p.Cm = w.Kstar; p.XsT = w.XsT; p.Zs = w.Zs; p.csum = w.csum; p.inv_ls = gp->inv_ls; p.dZ = dZ;
    p.ldc = np; p.ldt = np; p.ldz = dp; p.lddz = lddz; p.n_pad = np; p.m = (int)m; p.d = gp->d;
    p.bsC = bsM; p.bsT = (long long)dpt * np; p.bsZ = (long long)rows_alloc * dp; p.bscs = rows_alloc; p.bsdz = bsdz;
    k_dz<<<dim3(mt, dpt / 128, B), NTHREADS, GEMM_SMEM_BYTES, st>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (gSigma && q > 1) {
        k_dz_pairs<<<dim3((unsigned)(m / q), B), 256, 0, st>>>(w.Zs, dp, (long long)rows_alloc * dp, gSigma, bsS, gp->hyp, gp->inv_ls, q,
                                                               gp->d, gp->kind, dZ, lddz, bsdz);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// cdist finish: D = sqrt(max(xn_i + xn_j - 2 G, 0)), exact 0 on the diagonal
__global__ void k_cdist_finish(double* __restrict__ D, int ld, const double* __restrict__ xn, int n_pad) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_pad * n_pad) return;
    const int i = (int)(idx / n_pad), j = (int)(idx - (size_t)i * n_pad);
    const double g = D[(size_t)i * ld + j];
    D[(size_t)i * ld + j] = (i == j) ? 0.0 : sqrt(fmax(xn[i] + xn[j] - 2.0 * g, 0.0));
}

cudaError_t launch_cdist_finish(double* D, int ld, const double* xn, int n_pad, cudaStream_t st) {
    const size_t tot = (size_t)n_pad * n_pad;
    k_cdist_finish<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(D, ld, xn, n_pad);
    return cudaGetLastError();
}

}  // namespace hdbo
