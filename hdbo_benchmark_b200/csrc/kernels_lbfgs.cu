// Batched box-constrained L-BFGS on the device: R independent problems (restarts of optimize_acqf), one CTA each.
//
// Replaces the scipy L-BFGS-B host loop inside botorch's gen_candidates_scipy.  One iteration is
//   k_lbfgs_propose   projected gradient -> two-loop recursion over the (s, y) ring -> direction on the free variables ->
//                     T trial points  clip(x + 2^-j dir)  (a whole backtracking line search, evaluated in ONE batched f/g call)
//   [caller]          f and gradient at the R x T trial points (hdbo_posterior_fwd / hdbo_acq_elementwise / hdbo_posterior_bwd)
//   k_lbfgs_accept    best trial that satisfies the Armijo condition along the projected path (else the best decreasing one),
//                     curvature-checked ring update, convergence flags (projected-gradient sup-norm, relative decrease, maxiter)
// so the host enqueues a handful of launches per iteration and never reads a value back until it polls the `done` flags.
#include "kernels.cuh"
#include "lbfgs.cuh"

namespace hdbo {

namespace {
constexpr int LB_THREADS = 256;
constexpr int LB_EXPAND = 2;            // trial j has length step * 2^(LB_EXPAND - j)
constexpr double LB_C1 = 1e-4, LB_C2 = 0.9;   // Armijo / curvature constants (scipy L-BFGS-B: ftol-type 1e-3 / 0.9; Nocedal-Wright: 1e-4 / 0.9)

__device__ __forceinline__ double block_sum(double v, double* red) {   // red: >= 9 doubles of shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                       // previous users of `red` are done
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (l < LB_THREADS / 32) ? red[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) red[8] = t;
    }
    __syncthreads();
    return red[8];
}
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (l < LB_THREADS / 32) ? red[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (l == 0) red[8] = t;
    }
    __syncthreads();
    return red[8];
}
// gradient component that survives the projection onto the box at x
__device__ __forceinline__ double proj_grad(double x, double g, double lo, double hi) {
    if (x <= lo && g > 0.0) return 0.0;
    if (x >= hi && g < 0.0) return 0.0;
    return g;
}
}  // namespace

// meta[r] = {count, head, done (1 converged, 2 stalled), iters}
__global__ void __launch_bounds__(LB_THREADS) k_lbfgs_propose(LbfgsParams p) {
    extern __shared__ double sm[];
    double* q = sm;                 // [D] working vector of the two-loop recursion
    double* red = sm + p.D;         // [16]
    double* alpha = red + 16;       // [M]
    const int r = blockIdx.x, D = p.D, M = p.M, tid = threadIdx.x;
    const double* x = p.x + (size_t)r * D;
    const double* g = p.g + (size_t)r * D;
    int* meta = p.meta + 4 * r;
    double* Xt = p.Xt + (size_t)r * p.T * D;
    if (meta[2] > 0) {              // finished: keep the trial evaluations harmless
        for (int j = 0; j < p.T; j++)
            for (int i = tid; i < D; i += LB_THREADS) Xt[(size_t)j * D + i] = x[i];
        return;
    }
    const double* S = p.S + (size_t)r * M * D;
    const double* Y = p.Y + (size_t)r * M * D;
    const double* rho = p.rho + (size_t)r * M;
    int count = meta[0];
    const int head = meta[1];       // index of the newest pair
    double pg2 = 0.0;
    for (int i = tid; i < D; i += LB_THREADS) { const double v = proj_grad(x[i], g[i], p.lo[i], p.hi[i]); q[i] = v; pg2 += v * v; }
    pg2 = block_sum(pg2, red);
    double gamma = 1.0 / fmax(sqrt(pg2), 1e-300);          // first iteration: unit-length first trial step
    for (int pass = 0; pass < 2; pass++) {
        // ---- two-loop recursion: q <- H pg ----
        for (int k = 0; k < count; k++) {                   // newest -> oldest
            const int idx = (head - k + M) % M;
            const double* s = S + (size_t)idx * D; const double* y = Y + (size_t)idx * D;
            double d = 0.0;
            for (int i = tid; i < D; i += LB_THREADS) d += s[i] * q[i];
            const double a = rho[idx] * block_sum(d, red);
            if (tid == 0) alpha[k] = a;
            for (int i = tid; i < D; i += LB_THREADS) q[i] -= a * y[i];
        }
        if (count > 0) {
            const double* s = S + (size_t)head * D; const double* y = Y + (size_t)head * D;
            double sy = 0.0, yy = 0.0;
            for (int i = tid; i < D; i += LB_THREADS) { sy += s[i] * y[i]; yy += y[i] * y[i]; }
            sy = block_sum(sy, red); yy = block_sum(yy, red);
            gamma = sy / fmax(yy, 1e-300);
        }
        for (int i = tid; i < D; i += LB_THREADS) q[i] *= gamma;
        __syncthreads();
        for (int k = count - 1; k >= 0; k--) {              // oldest -> newest
            const int idx = (head - k + M) % M;
            const double* s = S + (size_t)idx * D; const double* y = Y + (size_t)idx * D;
            double d = 0.0;
            for (int i = tid; i < D; i += LB_THREADS) d += y[i] * q[i];
            const double b = rho[idx] * block_sum(d, red);
            const double a = alpha[k];
            for (int i = tid; i < D; i += LB_THREADS) q[i] += s[i] * (a - b);
        }
        // ---- direction = -H pg on the free variables; must be a descent direction, else restart from steepest descent ----
        double gd = 0.0;
        for (int i = tid; i < D; i += LB_THREADS) {
            const double v = (proj_grad(x[i], g[i], p.lo[i], p.hi[i]) != 0.0) ? -q[i] : 0.0;
            q[i] = v; gd += v * g[i];
        }
        gd = block_sum(gd, red);
        if (gd < 0.0 || count == 0) break;
        count = 0;                                          // drop the history, redo as steepest descent
        if (tid == 0) meta[0] = 0;
        gamma = 1.0 / fmax(sqrt(pg2), 1e-300);
        for (int i = tid; i < D; i += LB_THREADS) q[i] = proj_grad(x[i], g[i], p.lo[i], p.hi[i]);
        __syncthreads();
    }
    // ---- trial points of the line search: clip(x + step 2^(LB_EXPAND - j) dir): two expansions, the last accepted length, then
    //      halvings -- evaluated together, so that the accept kernel can pick a point satisfying the strong Wolfe conditions ----
    double t = p.step[r] * (double)(1 << LB_EXPAND);
    for (int j = 0; j < p.T; j++, t *= 0.5)
        for (int i = tid; i < D; i += LB_THREADS) Xt[(size_t)j * D + i] = fmin(fmax(x[i] + t * q[i], p.lo[i]), p.hi[i]);
}

__global__ void __launch_bounds__(LB_THREADS) k_lbfgs_accept(LbfgsParams p) {
    __shared__ double red[16];
    __shared__ int s_pick, s_head;
    const int r = blockIdx.x, D = p.D, M = p.M, T = p.T, tid = threadIdx.x;
    int* meta = p.meta + 4 * r;
    if (meta[2] > 0) return;
    // ring slot of the next curvature pair: read ONCE by thread 0 and shared, so that no late warp can see the meta[] that thread 0
    // rewrites further down (a warp that read the updated meta would have copied its part of (s, y) into the wrong slot)
    if (tid == 0) s_head = (meta[0] == 0) ? 0 : (meta[1] + 1) % M;
    double* x = p.x + (size_t)r * D;
    double* g = p.g + (size_t)r * D;
    const double* Xt = p.Xt + (size_t)r * T * D;
    const double* gt = p.gt + (size_t)r * T * D;
    const double* ft = p.ft + (size_t)r * T;
    const double f0 = p.f[r];
    // The T trials were evaluated together (values AND gradients), so the line search is a selection: the lowest trial that satisfies
    // the STRONG WOLFE conditions along the projected path p_j = x_j - x,
    //     f_j <= f0 + c1 g.p_j   (sufficient decrease)      |g_j.p_j| <= c2 |g.p_j|   (curvature),
    // what scipy's L-BFGS-B line search (dcsrch) terminates on; else the lowest Armijo trial (a first-acceptable backtracking search
    // took full steps with tiny decreases on badly scaled directions and stopped early on ftol).
    int pick = -1, pick_armijo = -1;
    double bestf = f0, bestf_armijo = f0;
    for (int j = 0; j < T; j++) {
        double d = 0.0, c = 0.0;
        for (int i = tid; i < D; i += LB_THREADS) {
            const double pj = Xt[(size_t)j * D + i] - x[i];
            d += g[i] * pj;
            c += gt[(size_t)j * D + i] * pj;
        }
        d = block_sum(d, red);
        c = block_sum(c, red);
        const double fj = ft[j];
        if (!(isfinite(fj) && isfinite(c)) || !(d < 0.0) || !(fj <= f0 + LB_C1 * d)) continue;
        if (fj < bestf_armijo) { bestf_armijo = fj; pick_armijo = j; }
        if (fabs(c) <= LB_C2 * fabs(d) && fj < bestf) { bestf = fj; pick = j; }
    }
    if (pick < 0) pick = pick_armijo;
    if (pick < 0) {                                          // fall back to the best strictly decreasing trial
        double best = f0;
        for (int j = 0; j < T; j++) if (isfinite(ft[j]) && ft[j] < best) { best = ft[j]; pick = j; }
    }
    if (tid == 0) s_pick = pick;
    __syncthreads();
    pick = s_pick;
    if (pick < 0) {      // no acceptable trial: keep backtracking next iteration from a clean steepest-descent state
        if (tid == 0) {
            const double st = p.step[r] * exp2((double)(LB_EXPAND - T));
            p.step[r] = st; meta[0] = 0;
            const int it = ++meta[3];
            if (st < 1e-18 || it >= p.maxiter) meta[2] = 2;                  // stalled (flag 2)
        }
        return;
    }
    const double* xn = Xt + (size_t)pick * D;
    const double* gn = gt + (size_t)pick * D;
    // curvature pair
    double sy = 0.0, yy = 0.0;
    for (int i = tid; i < D; i += LB_THREADS) { const double s = xn[i] - x[i], y = gn[i] - g[i]; sy += s * y; yy += y * y; }
    sy = block_sum(sy, red); yy = block_sum(yy, red);
    if (sy > 2.220446049250313e-16 * yy && sy > 0.0) {
        const int head = s_head;     // (published by the barriers of block_sum above)
        double* S = p.S + ((size_t)r * M + head) * D;
        double* Y = p.Y + ((size_t)r * M + head) * D;
        for (int i = tid; i < D; i += LB_THREADS) { S[i] = xn[i] - x[i]; Y[i] = gn[i] - g[i]; }
        if (tid == 0) { p.rho[(size_t)r * M + head] = 1.0 / sy; meta[1] = head; meta[0] = min(meta[0] + 1, M); }
    }
    double pgmax = 0.0;
    for (int i = tid; i < D; i += LB_THREADS) {
        const double xv = xn[i], gv = gn[i];
        x[i] = xv; g[i] = gv;
        pgmax = fmax(pgmax, fabs(proj_grad(xv, gv, p.lo[i], p.hi[i])));
    }
    pgmax = block_max(pgmax, red);
    if (tid == 0) {
        const double f1 = ft[pick];
        p.f[r] = f1;
        // next grid {4a, 2a, a, a/2, ...} is centred on the length a that was just accepted
        p.step[r] = fmin(fmax(p.step[r] * exp2((double)(LB_EXPAND - pick)), 1e-20), 1e6);
        const int it = ++meta[3];
        const double rel = (f0 - f1) / fmax(fmax(fabs(f0), fabs(f1)), 1.0);
        if (pgmax <= p.gtol || rel <= p.ftol || it >= p.maxiter) meta[2] = 1;
    }
}

cudaError_t launch_lbfgs_propose(const LbfgsParams& p, cudaStream_t st) {
    if (p.R == 0) return cudaSuccess;
    const size_t smem = (size_t)(p.D + 16 + p.M) * sizeof(double);
    if (smem > 48 * 1024) {   // per-device attribute, microseconds: set it whenever it is needed rather than remembering it per process
        cudaError_t e = cudaFuncSetAttribute(k_lbfgs_propose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_lbfgs_propose<<<p.R, LB_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_lbfgs_accept(const LbfgsParams& p, cudaStream_t st) {
    if (p.R == 0) return cudaSuccess;
    k_lbfgs_accept<<<p.R, LB_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace hdbo
