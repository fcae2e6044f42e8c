// HBM-bound kernels of the GP hot path: input centring/scaling, triangular mat-vecs (alpha, quadratic form),
// log-determinant, the acquisition epilogue (EI / LogEI / UCB, value and partial derivatives) and argmax.
// All reductions are deterministic (fixed shuffle trees, fixed partial order).
#include "kernels.cuh"
#include "elem.cuh"

namespace hdbo {

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in all threads
__device__ __forceinline__ double block_sum(double x, double* red /* 32 doubles */) {
    x = warp_sum(x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += red[w];
    return s;
}

// ------------------------------------------------------------------------------------------------------------
// Xs[i][k] = (X[i][k] - center[k]) * inv_ls[k]  (k < d), zero padding to d_pad and rows_pad; norms[i] = |Xs[i]|^2.
// One warp per row, lanes stride the row (coalesced).  A1 "centred inputs / lengthscale" step.
__global__ void k_scale_inputs(const double* __restrict__ X, int ldx, long long bsX, int nrows, int d,
                               const double* __restrict__ center, const double* __restrict__ inv_ls, long long bs_ls,
                               double* __restrict__ Xs, int d_pad, int rows_pad, long long bsXs,
                               double* __restrict__ norms, long long bsn) {
    const int b = blockIdx.y;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows_pad) return;
    const double* x = X + b * bsX + (size_t)row * ldx;
    const double* il = inv_ls + b * bs_ls;
    double* o = Xs + b * bsXs + (size_t)row * d_pad;
    double s = 0.0;
    for (int k = lane; k < d_pad; k += 32) {
        double v = 0.0;
        if (row < nrows && k < d) v = (x[k] - (center ? center[k] : 0.0)) * (inv_ls ? il[k] : 1.0);
        o[k] = v;
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) norms[b * bsn + row] = s;
}

cudaError_t launch_scale_inputs(const double* X, int ldx, long long bsX, int nrows, int d, const double* center,
                                const double* inv_ls, long long bs_ls, double* Xs, int d_pad, int rows_pad,
                                long long bsXs, double* norms, long long bsn, int batch, cudaStream_t stream) {
    dim3 grid((rows_pad + 7) / 8, batch);
    k_scale_inputs<<<grid, 256, 0, stream>>>(X, ldx, bsX, nrows, d, center, inv_ls, bs_ls, Xs, d_pad, rows_pad, bsXs,
                                             norms, bsn);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// out[i] = sum_{j in range(i)} M[i][j] v[j]; lower: j <= i, upper: j >= i.  One warp per row.
// rhs mode: v[j] = y[j] - mean (j < n) when y != nullptr, else v = vin.
// HBM-bound (n_pad^2 * 4 B per call): 16-byte loads, four independent row segments in flight per lane (the first version -- one 8-byte
// load per lane per iteration -- reached 0.22 of the copy bandwidth), long rows first (upper: row 0 first; lower: last row first).
__global__ void __launch_bounds__(256) k_tri_matvec(const double* __restrict__ M, int ld, long long bsM, int n_pad, int upper,
                             const double* __restrict__ vin, long long bsv, const double* __restrict__ y, int n,
                             const double* __restrict__ hyp, double* __restrict__ out, long long bso) {
    const int b = blockIdx.y;
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    if (!upper) row = n_pad - 1 - row;
    const double* m = M + b * bsM + (size_t)row * ld;
    const double mean = y ? hyp[b * 4 + 2] : 0.0;
    const double* vv = vin ? vin + b * bsv : nullptr;
    const int j0 = upper ? row : 0, j1 = upper ? n_pad : row + 1;
    auto vec = [&](int j) -> double { return y ? ((j < n) ? __ldg(y + j) - mean : 0.0) : __ldg(vv + j); };
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int ja = j0 & ~1;                       // 16-byte aligned start (ld, n_pad are even)
    const int jb = j1 & ~1;                       // end of the full pairs
    int j = ja + 2 * lane;
    for (; j + 192 < jb; j += 256) {
        const double2 a0 = *reinterpret_cast<const double2*>(m + j);
        const double2 a1 = *reinterpret_cast<const double2*>(m + j + 64);
        const double2 a2 = *reinterpret_cast<const double2*>(m + j + 128);
        const double2 a3 = *reinterpret_cast<const double2*>(m + j + 192);
        s0 = fma((j >= j0) ? a0.x : 0.0, (j >= j0) ? vec(j) : 0.0, s0); s0 = fma(a0.y, vec(j + 1), s0);
        s1 = fma(a1.x, vec(j + 64), s1); s1 = fma(a1.y, vec(j + 65), s1);
        s2 = fma(a2.x, vec(j + 128), s2); s2 = fma(a2.y, vec(j + 129), s2);
        s3 = fma(a3.x, vec(j + 192), s3); s3 = fma(a3.y, vec(j + 193), s3);
    }
    for (; j < jb; j += 64) {
        const double2 a0 = *reinterpret_cast<const double2*>(m + j);
        s0 = fma((j >= j0) ? a0.x : 0.0, (j >= j0) ? vec(j) : 0.0, s0); s0 = fma(a0.y, vec(j + 1), s0);
    }
    if (lane == 0 && jb < j1) s1 = fma(m[jb], vec(jb), s1);      // odd tail (lower rows of even index)
    double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[b * bso + row] = s;
}

cudaError_t launch_tri_matvec(const double* M, int ld, long long bsM, int n_pad, int upper, const double* vin,
                              long long bsv, const double* y, int n, const double* hyp, double* out, long long bso,
                              int batch, cudaStream_t stream) {
    dim3 grid((n_pad + 7) / 8, batch);
    k_tri_matvec<<<grid, 256, 0, stream>>>(M, ld, bsM, n_pad, upper, vin, bsv, y, n, hyp, out, bso);
    return cudaGetLastError();
}

// scal[b][0] = logdet = 2 sum log L_ii ; scal[b][1] = quad = w.w ; scal[b][2] = data term of n*mll:
// -1/2 quad - 1/2 logdet - n/2 log 2pi ; scal[b][3] = sum alpha.   One block per GP.
__global__ void k_fit_scalars(const double* __restrict__ L, int ld, long long bsL, const double* __restrict__ w,
                              const double* __restrict__ alpha, long long bsv, int n, int n_pad, double* __restrict__ scal) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    double sl = 0.0, sq = 0.0, sa = 0.0;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        sl += log(L[b * bsL + (size_t)i * ld + i]);
        double wi = w[b * bsv + i];
        sq = fma(wi, wi, sq);
        sa += alpha[b * bsv + i];
    }
    sl = block_sum(sl, red);
    sq = block_sum(sq, red);
    sa = block_sum(sa, red);
    if (threadIdx.x == 0) {
        scal[b * 8 + 0] = 2.0 * sl;
        scal[b * 8 + 1] = sq;
        scal[b * 8 + 2] = -0.5 * sq - sl - 0.5 * n * LOG_2PI;
        scal[b * 8 + 3] = sa;
    }
}

cudaError_t launch_fit_scalars(const double* L, int ld, long long bsL, const double* w, const double* alpha,
                               long long bsv, int n, int n_pad, double* scal, int batch, cudaStream_t stream) {
    k_fit_scalars<<<batch, 1024, 0, stream>>>(L, ld, bsL, w, alpha, bsv, n, n_pad, scal);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Per candidate: mean = m0 + sum_t meanpart[t][i]; var = outputscale - sum_t varpart[t][i] (+ noise);
// acquisition value (A5) and, if requested, d acq / d mean and d acq / d var.
// HBM-bound (8 B x (ntiles + ntiles_var) per candidate): 32 candidates x 8 tile slices per CTA so that a 18 944-candidate chunk is 592
// CTAs = 4 per SM and every thread issues its 4 + 8 loads at once (64 x 4: 296 CTAs, 0.11 of the copy bandwidth in a 21 us launch);
// the slice sums are combined through shared memory in a fixed order (deterministic, independent of the chunking).
constexpr int FIN_CAND = 32, FIN_SLICES = 8;
__global__ void __launch_bounds__(FIN_CAND * FIN_SLICES)
k_finalize_acq(const double* __restrict__ meanpart, const double* __restrict__ varpart, long long bsp,
               int ntiles, int ntiles_var, int rows_pad, int m, const double* __restrict__ hyp, int observation_noise,
               int acq_kind, double acq_param, double* __restrict__ mean_out,
               double* __restrict__ var_out, double* __restrict__ acq_out, long long bso) {
    __shared__ double red[2][FIN_SLICES][FIN_CAND];
    const int b = blockIdx.y;
    const int c = threadIdx.x % FIN_CAND, sl = threadIdx.x / FIN_CAND;
    const int i = blockIdx.x * FIN_CAND + c;
    const bool live = i < m;
    auto slice_sum = [&](const double* part, int nt) {
        const int per = (nt + FIN_SLICES - 1) / FIN_SLICES, t0 = sl * per, t1 = min(nt, t0 + per);
        const double* p = part + b * 2 * bsp + i;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int t = t0;
        for (; t + 3 < t1; t += 4) {
            s0 += p[(size_t)t * rows_pad]; s1 += p[(size_t)(t + 1) * rows_pad];
            s2 += p[(size_t)(t + 2) * rows_pad]; s3 += p[(size_t)(t + 3) * rows_pad];
        }
        for (; t < t1; t++) s0 += p[(size_t)t * rows_pad];
        return (s0 + s1) + (s2 + s3);
    };
    red[0][sl][c] = live ? slice_sum(meanpart, ntiles) : 0.0;
    red[1][sl][c] = live ? slice_sum(varpart, ntiles_var) : 0.0;
    __syncthreads();
    if (sl != 0 || !live) return;
    static_assert(FIN_SLICES == 8, "fixed-order combine below is written for 8 slices");
    const double sm = ((red[0][0][c] + red[0][1][c]) + (red[0][2][c] + red[0][3][c])) + ((red[0][4][c] + red[0][5][c]) + (red[0][6][c] + red[0][7][c]));
    const double sv = ((red[1][0][c] + red[1][1][c]) + (red[1][2][c] + red[1][3][c])) + ((red[1][4][c] + red[1][5][c]) + (red[1][6][c] + red[1][7][c]));
    const double os = hyp[b * 4 + 0], noise = hyp[b * 4 + 1], m0 = hyp[b * 4 + 2];
    const double mean = m0 + sm;
    double var = os - sv;
    if (observation_noise) var += noise;
    if (mean_out) mean_out[b * bso + i] = mean;
    if (var_out) var_out[b * bso + i] = var;
    if (acq_out) {
        double a, dm, dv;
        acq_eval(acq_kind, acq_param, mean, var, a, dm, dv);
        acq_out[b * bso + i] = a;
    }
}

cudaError_t launch_finalize_acq(const double* meanpart, const double* varpart, long long bsp, int ntiles, int ntiles_var, int rows_pad,
                                int m, const double* hyp, int observation_noise, int acq_kind, double acq_param,
                                double* mean_out, double* var_out, double* acq_out, long long bso, int batch,
                                cudaStream_t stream) {
    dim3 grid((m + FIN_CAND - 1) / FIN_CAND, batch);
    k_finalize_acq<<<grid, FIN_CAND * FIN_SLICES, 0, stream>>>(meanpart, varpart, bsp, ntiles, ntiles_var, rows_pad, m, hyp,
                                                              observation_noise, acq_kind, acq_param, mean_out, var_out, acq_out, bso);
    return cudaGetLastError();
}

// row sums of squares of V [rows][cols] (one CTA per row, 16-byte accesses) -> out[b*bso + row]   (small-m variance path).
// add != nullptr: columns >= add_from of V first receive the second parts of a split-K product (add[row][col - add_from], leading
// dimension ldadd)
// and are written back, so that V is whole for the gradient path.
__global__ void __launch_bounds__(256) k_row_sumsq(double* __restrict__ V, int ldv, long long bsV, int rows, int cols,
                                                   const double* __restrict__ add, int ldadd, int add_from, double* __restrict__ out, long long bso) {
    __shared__ double red[8];
    const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
    double2* src = reinterpret_cast<double2*>(V + b * bsV + (size_t)row * ldv);
    const double2* ad = add ? reinterpret_cast<const double2*>(add + (size_t)row * ldadd) : nullptr;
    double s = 0.0;
    for (int j = threadIdx.x; j < cols / 2; j += 256) {
        double2 v = src[j];
        if (ad && 2 * j >= add_from) { const double2 a = ad[j - add_from / 2]; v.x += a.x; v.y += a.y; src[j] = v; }
        s = fma(v.x, v.x, s); s = fma(v.y, v.y, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) out[b * bso + row] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

cudaError_t launch_row_sumsq(double* V, int ldv, long long bsV, int rows, int cols, const double* add, int ldadd, int add_from,
                             double* out, long long bso, int batch, cudaStream_t stream) {
    if (rows == 0) return cudaSuccess;
    k_row_sumsq<<<dim3(rows, batch), 256, 0, stream>>>(V, ldv, bsV, rows, cols, add, ldadd, add_from, out, bso);
    return cudaGetLastError();
}

// acquisition from given mean / var arrays: value + partials (used by the autograd path and tested on its own)
__global__ void k_acq_elem(const double* __restrict__ mean, const double* __restrict__ var, long long m, int acq_kind,
                           double acq_param, double* __restrict__ acq, double* __restrict__ dmean, double* __restrict__ dvar) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double a, dm, dv;
    acq_eval(acq_kind, acq_param, mean[i], var[i], a, dm, dv);
    if (acq) acq[i] = a;
    if (dmean) dmean[i] = dm;
    if (dvar) dvar[i] = dv;
}

cudaError_t launch_acq_elem(const double* mean, const double* var, long long m, int acq_kind, double acq_param, double* acq,
                            double* dmean, double* dvar, cudaStream_t stream) {
    if (m == 0) return cudaSuccess;
    k_acq_elem<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(mean, var, m, acq_kind, acq_param, acq, dmean, dvar);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// argmax with first-index tie-break; NaNs are ignored.  Two phases in one launch (last block reduces the partials).
__global__ void k_argmax(const double* __restrict__ v, long long m, double* __restrict__ pval, long long* __restrict__ pidx,
                         unsigned int* __restrict__ counter, double* __restrict__ out_val, long long* __restrict__ out_idx) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    __shared__ bool last;
    double best = -INFINITY; long long bi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        double x = v[i];
        if (x > best || (bi < 0 && x == x)) { best = x; bi = i; }
    }
    auto combine = [](double& a, long long& ai, double b, long long bidx) {
        if (bidx >= 0 && (ai < 0 || b > a || (b == a && bidx < ai))) { a = b; ai = bidx; }
    };
    auto block_reduce = [&](double& a, long long& ai) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, a, o);
            long long oi = __shfl_xor_sync(0xffffffffu, ai, o);
            combine(a, ai, ob, oi);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) { sv[warp] = a; si[warp] = ai; }
        __syncthreads();
        if (warp == 0) {
            a = (lane < (blockDim.x >> 5)) ? sv[lane] : -INFINITY;
            ai = (lane < (blockDim.x >> 5)) ? si[lane] : -1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, a, o);
                long long oi = __shfl_xor_sync(0xffffffffu, ai, o);
                combine(a, ai, ob, oi);
            }
        }
    };
    block_reduce(best, bi);
    if (threadIdx.x == 0) {
        pval[blockIdx.x] = best; pidx[blockIdx.x] = bi;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    best = -INFINITY; bi = -1;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) combine(best, bi, pval[i], pidx[i]);
    block_reduce(best, bi);
    if (threadIdx.x == 0) { *out_val = best; *out_idx = bi; *counter = 0u; }
}

cudaError_t launch_argmax(const double* v, long long m, void* work, double* out_val, long long* out_idx, cudaStream_t stream) {
    // work: [0,8) counter (must be zero on entry; reset on exit), then 296 doubles, then 296 int64
    const int nblocks = 296;
    unsigned int* counter = reinterpret_cast<unsigned int*>(work);
    double* pval = reinterpret_cast<double*>(reinterpret_cast<char*>(work) + 8);
    long long* pidx = reinterpret_cast<long long*>(pval + nblocks);
    k_argmax<<<nblocks, 256, 0, stream>>>(v, m, pval, pidx, counter, out_val, out_idx);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Scrambled-Sobol raw samples on the device (the pool `gen_batch_initial_conditions` scores), bit-identical to
// torch.quasirandom.SobolEngine: point k = shift ^ XOR_{b in bits(gray(k))} state[b], value = quasi 2^-30, then lower + range * u
// with separate roundings (what `lower + rng * u` does on the host).  One thread per dimension and run of 64 consecutive points:
// the start of a run costs popcount(gray) loads, every further point ONE xor with state[ctz(k)] (bits 0..5 live in registers);
// consecutive threads = consecutive dimensions, so every store instruction writes whole 128-byte lines.  HBM-bound: 8 B/element.
// torch quirk kept (first_fp32 != 0): point 0 is computed in torch's DEFAULT dtype when the engine is constructed
// (SobolEngine._first_point), i.e. rounded to FP32 unless the process runs with float64 as default dtype.
constexpr int SOBOL_RUN = 64;
__global__ void __launch_bounds__(128) k_sobol_draw(const int* __restrict__ state_t /* [30][dim] */, const int* __restrict__ shift,
                                                    int dim, long long k0, long long m, const double* __restrict__ lower,
                                                    const double* __restrict__ range, double* __restrict__ out, long long ld,
                                                    int first_fp32) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    const long long run0 = (k0 / SOBOL_RUN + blockIdx.x) * SOBOL_RUN;       // absolute index of this run's first point (aligned)
    int low[6];
#pragma unroll
    for (int b = 0; b < 6; b++) low[b] = state_t[(size_t)b * dim + j];
    unsigned long long g = (unsigned long long)run0 ^ ((unsigned long long)run0 >> 1);
    int quasi = shift[j];
    for (int b = 0; g != 0 && b < 30; b++, g >>= 1)
        if (g & 1ull) quasi ^= state_t[(size_t)b * dim + j];
    const double lo = lower ? lower[j] : 0.0, rg = range ? range[j] : 1.0;
#pragma unroll
    for (int i = 0; i < SOBOL_RUN; i++) {
        const long long k = run0 + i;
        // gray(k) ^ gray(k-1) = bit ctz(k) = ctz(i) inside an aligned run: a compile-time register index after the full unroll
        if (i > 0) quasi ^= low[(i & 1) ? 0 : (i & 2) ? 1 : (i & 4) ? 2 : (i & 8) ? 3 : (i & 16) ? 4 : 5];
        if (k >= k0 && k < k0 + m) {
            double u = (double)quasi * 9.313225746154785e-10;                 // 2^-30, exact
            if (k == 0 && first_fp32) u = (double)__fmul_rn(__int2float_rn(quasi), 9.313225746154785e-10f);
            out[(size_t)(k - k0) * ld + j] = __dadd_rn(lo, __dmul_rn(rg, u));
        }
    }
}

cudaError_t launch_sobol_draw(const int* state_t, const int* shift, int dim, long long k0, long long m, const double* lower,
                              const double* range, double* out, long long ld, int first_fp32, cudaStream_t stream) {
    if (m == 0) return cudaSuccess;
    const long long runs = (k0 + m - 1) / SOBOL_RUN - k0 / SOBOL_RUN + 1;
    dim3 grid((unsigned)runs, (dim + 127) / 128);
    k_sobol_draw<<<grid, 128, 0, stream>>>(state_t, shift, dim, k0, m, lower, range, out, ld, first_fp32);
    return cudaGetLastError();
}

}  // namespace hdbo
