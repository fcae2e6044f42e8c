// extern "C" entry points of libhdbo_b200.so (declared in include/hdbo_b200.h).  Host-side orchestration only:
// argument checks, workspace carving, launch sequences on the caller's stream.  No allocation, no synchronisation.
#include "gemm_core.cuh"
#include "kernels.cuh"
#include "elem.cuh"
#include "grad.cuh"
#include "i8.cuh"
#include "mlp.cuh"
#include "lbfgs.cuh"

#include <cstdlib>
#include <vector>

using namespace hdbo;

#define TRY_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return (int)_e; } while (0)

static inline int64_t rup(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static int check_gp(const hdbo_gp* gp) {
    if (!gp) return -1;
    if (gp->n <= 0 || gp->d <= 0 || gp->batch <= 0) return -1;
    if (gp->n_pad != rup(gp->n, 128) || gp->d_pad != rup(gp->d, 16)) return -1;
    if (gp->kind != HDBO_KERNEL_MATERN52 && gp->kind != HDBO_KERNEL_RBF) return -1;
    return 0;
}

extern "C" int hdbo_version(void) { return 200; }

// ---- opt-in timing: CUDA events of a CALLER-OWNED timeline (hdbo_gp.timeline, may be NULL) recorded on the caller's stream around
// the phases of an entry point.  The library keeps no state of its own: the caller creates the events and reads them back.
static inline void tl_mark(const hdbo_gp* gp, int phase, cudaStream_t st) {
    hdbo_timeline* tl = gp->timeline;
    if (!tl || !tl->events || tl->used >= tl->capacity) return;
    if (tl->phase) tl->phase[tl->used] = phase;
    cudaEventRecord((cudaEvent_t)tl->events[tl->used++], st);
}


// ------------------------------------------------------------------------------------------------------------
extern "C" int hdbo_gp_fit(const hdbo_gp* gp, int need_r2, void* stream_) {
    return hdbo_gp_fit_downdated(gp, need_r2, nullptr, 0, 0, stream_);
}

extern "C" int hdbo_gp_fit_downdated(const hdbo_gp* gp, int need_r2, const double* V, int64_t ldv, int64_t k_pad,
                                     void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (need_r2 && !gp->R2) return -2;
    if (V && (gp->batch != 1 || ldv < k_pad || k_pad % 128 || ldv % 2)) return -4;
    cudaStream_t st = (cudaStream_t)stream_;
    const int np = gp->n_pad, dp = gp->d_pad, B = gp->batch;
    const long long nn = (long long)np * np;
    TRY_CUDA(cudaMemsetAsync(gp->info, 0, sizeof(int32_t) * B, st));
    static const int pipe_min_tiles = [] { const char* e = getenv("HDBO_PIPE_MIN_TILES"); return e ? atoi(e) : PIPE_MIN_TILES; }();   // (developer knob)
    bool piped = B == 1 && np / 128 >= pipe_min_tiles;
    for (int i = 0; i < 4 && piped; i++) piped = gp->aux_streams[i] != nullptr;
    for (int i = 0; i < 12 && piped; i++) piped = gp->aux_events[i] != nullptr;
    tl_mark(gp, HDBO_PHASE_FIT_BUILD, st);
    // (pipelined factorisation: the unused UPPER tiles of Awork hold the running rows of L^-1; their first update starts from zero
    // accumulators, so nothing is cleared here)
    TRY_CUDA(launch_scale_inputs(gp->X, (int)gp->ldx, 0, gp->n, gp->d, gp->center, gp->inv_ls, gp->d, gp->Xs, dp, np,
                                 (long long)np * dp, gp->xn, np, B, st));
    SymBuildParams sb{};
    sb.Xs = gp->Xs; sb.xn = gp->xn; sb.hyp = gp->hyp; sb.Khat = gp->Awork; sb.R2 = need_r2 ? gp->R2 : nullptr;
    sb.ldx = dp; sb.ldk = np; sb.bsX = (long long)np * dp; sb.bsxn = np; sb.bsK = nn;
    sb.n = gp->n; sb.tiles = np / 128; sb.kblocks = dp / BK; sb.kind = gp->kind;
    // pipelined factorisation: only the first four tile columns are built here; the rest of the build runs on the side stream beside
    // the first pair's chain (kernels_chol_pipe.cu)
    const bool split_build = piped && !V;
    if (split_build) { sb.colmajor = 1; sb.t0 = 0; sb.t1 = pipe_build_head_tiles(np / 128); }
    TRY_CUDA(launch_sym_build(sb, B, st));
    if (V) {   // Khat -= V V^T on the lower tiles (joint predictive covariance K(Z,Z) + noise I - V V^T)
        GemmParams g{};
        g.A = V; g.B = V; g.C = gp->Awork; g.Ct = nullptr;
        g.lda = g.ldb = (int)ldv; g.ldc = np; g.mtiles = g.ntiles = np / 128; g.kblocks = (int)(k_pad / BK);
        g.krange = KR_FULL; g.lower_only = 1; g.alpha = -1.0; g.beta = 1.0;
        TRY_CUDA(launch_gemm_nt(g, 1, st));
    }
    tl_mark(gp, HDBO_PHASE_FIT_BUILD, st);
    tl_mark(gp, HDBO_PHASE_FIT_FACTOR, st);
    double* kinv = (need_r2 && gp->Kinv) ? gp->Kinv : nullptr;
    if (piped) {
        ChainStreams cs{};
        cs.chain = (cudaStream_t)gp->aux_streams[0]; cs.bulk = (cudaStream_t)gp->aux_streams[1];
        cs.inverse = (cudaStream_t)gp->aux_streams[2]; cs.panel = (cudaStream_t)gp->aux_streams[3];
        for (int i = 0; i < 12; i++) cs.ev[i] = (cudaEvent_t)gp->aux_events[i];
        SymBuildParams rest = sb;
        rest.t0 = sb.t1; rest.t1 = 0;
        TRY_CUDA(chol_inverse_pipelined(gp->Awork, gp->L, gp->Linv, gp->Linvt, kinv, np, np / 128, gp->info, st, cs,
                                        split_build ? &rest : nullptr));
    } else {
        TRY_CUDA(chol_inverse(gp->Awork, gp->L, gp->Linv, gp->Linvt, np, nn, np / 128, gp->n, gp->info, B, st));
        if (kinv) TRY_CUDA(kinv_from_linvt(gp, kinv, st));
    }
    tl_mark(gp, HDBO_PHASE_FIT_FACTOR, st);
    tl_mark(gp, HDBO_PHASE_FIT_SOLVE, st);
    // w = Linv (y - m);  alpha = Linv^T w
    TRY_CUDA(launch_tri_matvec(gp->Linv, np, nn, np, 0, nullptr, 0, gp->y, gp->n, gp->hyp, gp->w, np, B, st));
    TRY_CUDA(launch_tri_matvec(gp->Linvt, np, nn, np, 1, gp->w, np, nullptr, gp->n, gp->hyp, gp->alpha, np, B, st));
    TRY_CUDA(launch_fit_scalars(gp->L, np, nn, gp->w, gp->alpha, np, gp->n, np, gp->scal, B, st));
    tl_mark(gp, HDBO_PHASE_FIT_SOLVE, st);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
namespace {
struct PostWs {
    double *Zs, *zn, *Kstar, *meanpart, *varpart, *DK, *V, *T, *XsT;
    double* P2;       // [rows][n_pad / 2]: second parts of the split-K triangular products of the few-candidate gradient path
    double* zscale;   // int8 mode: power-of-two row scales of the candidate digit images
    int8_t* Zdig;     // int8 mode: candidate digit images
    size_t bytes;
};
PostWs carve(const hdbo_gp* gp, int64_t rows, int keep, void* base) {
    const size_t B = gp->batch, np = gp->n_pad, dp = gp->d_pad, nt = np / 128;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += rup((int64_t)n, 16); return (double*)base + o; };
    PostWs w{};
    w.Zs = take(B * rows * dp);
    w.zn = take(B * rows);
    w.Kstar = take(B * rows * np);
    w.meanpart = take(B * 2 * nt * rows);  // the int8 kernels work on 64-column tiles: twice the partials
    w.varpart = take(B * 2 * nt * rows);
    w.zscale = take(B * rows);
    w.Zdig = reinterpret_cast<int8_t*>(take(B * (size_t)I8_S * rows * rup(gp->d, 32) / 8));
    if (keep) {
        w.DK = take(B * rows * np);
        w.V = take(B * rows * np);
        w.T = take(B * rows * np);
        w.XsT = take(B * (size_t)rup(gp->d, 128) * np);
        w.P2 = take(rows * np / 2);
    }
    w.bytes = off * sizeof(double);
    return w;
}

// state of the exact int8 variance path, carved from the caller's buffer (hdbo_i8_state_bytes)
struct I8State {
    int8_t* digits;   // I8_S n_pad^2 bytes: digit images of the row-scaled L^-1 (layout: kernels_i8.cu)
    double* scale;    // [n_pad] power-of-two row scales of L^-1
    int8_t* Xdig;     // I8_S n_pad roundup(d,32) bytes: digit images of the centred, scaled training inputs
    double* xscale;   // [n_pad] their row scales
    int32_t* flag;    // barrier-timeout flag of the tcgen05 kernels (0 = healthy)
    size_t bytes;
};
I8State carve_i8(const hdbo_gp* gp, void* base) {
    const size_t np = gp->n_pad, d32 = rup(gp->d, 32), B = gp->batch;   // per-GP blocks are contiguous inside each array
    I8State s{};
    char* p = (char*)base;
    s.digits = (int8_t*)p;  p += B * (size_t)I8_S * np * np;
    s.scale = (double*)p;   p += B * np * sizeof(double);
    s.Xdig = (int8_t*)p;    p += B * (size_t)I8_S * np * d32;
    s.xscale = (double*)p;  p += B * np * sizeof(double);
    s.flag = (int32_t*)p;   p += 16;
    s.bytes = (size_t)(p - (char*)base);
    return s;
}

// scale -> cross -> var for `mc` candidates starting at Z (rows_alloc = row capacity of the workspace arrays).
// i8 != nullptr: both GEMMs run as exact int8 digit-pair products on tcgen05 (kernels_i8.cu); K* only exists as digit images.
cudaError_t posterior_chunk(const hdbo_gp* gp, const PostWs& w, int64_t rows_alloc, const double* Z, int64_t ldz,
                            int64_t mc, int keep, cudaStream_t st, const I8State* i8 = nullptr, int* var_tiles = nullptr) {
    const int np = gp->n_pad, dp = gp->d_pad, B = gp->batch;
    const int rp = (int)rup(mc, 128), mt = rp / 128, nt = np / 128;
    cudaError_t e;
    if (i8) {
        // whole chunk on the int8 tensor cores: candidate digit images -> K* digit images + mean partials -> variance partials
        int8_t* Kimg = reinterpret_cast<int8_t*>(w.Kstar);   // tile images: 6 bytes per element of the 8-byte K* slab
        tl_mark(gp, HDBO_PHASE_PREP, st);
        e = launch_slice_rows(128, Z, ldz, (int)mc, rp, gp->d, gp->center, gp->inv_ls, w.Zdig, w.zscale, w.zn, rows_alloc, B, st);
        tl_mark(gp, HDBO_PHASE_PREP, st);
        if (e != cudaSuccess) return e;
        tl_mark(gp, HDBO_PHASE_CROSS, st);
        e = launch_cross_i8(gp->kind, w.Zdig, i8->Xdig, w.zscale, i8->xscale, w.zn, gp->xn, gp->alpha, gp->hyp, Kimg, w.meanpart,
                            (int)mc, gp->n, rp, (int)rows_alloc, np, gp->d, B, i8->flag, st);
        tl_mark(gp, HDBO_PHASE_CROSS, st);
        if (e != cudaSuccess) return e;
        tl_mark(gp, HDBO_PHASE_VAR, st);
        e = launch_var_i8(Kimg, rp, (int)rows_alloc, i8->digits, i8->scale, gp->hyp, np, w.varpart, B, i8->flag, st);
        tl_mark(gp, HDBO_PHASE_VAR, st);
        return e;
    }
    CrossParams c{};
    c.Zs = w.Zs; c.Xs = gp->Xs; c.zn = w.zn; c.xn = gp->xn; c.alpha = gp->alpha; c.hyp = gp->hyp;
    c.Kstar = w.Kstar; c.DK = keep ? w.DK : nullptr; c.meanpart = w.meanpart;
    c.ldz = dp; c.ldx = dp; c.ldk = np;
    c.bsZ = (long long)rows_alloc * dp; c.bsX = (long long)np * dp; c.bszn = rows_alloc; c.bsxn = np; c.bsalpha = np;
    c.bsK = (long long)rows_alloc * np; c.bsmp = 2LL * nt * rows_alloc;
    c.m = (int)mc; c.n = gp->n; c.mtiles = mt; c.ntiles = nt; c.kblocks = dp / BK; c.kind = gp->kind;
    tl_mark(gp, HDBO_PHASE_PREP, st);
    e = launch_scale_inputs(Z, (int)ldz, 0, (int)mc, gp->d, gp->center, gp->inv_ls, gp->d, w.Zs, dp, rp,
                            (long long)rows_alloc * dp, w.zn, rows_alloc, B, st);
    tl_mark(gp, HDBO_PHASE_PREP, st);
    if (e != cudaSuccess) return e;
    tl_mark(gp, HDBO_PHASE_CROSS, st);
    e = launch_cross(c, B, st);
    tl_mark(gp, HDBO_PHASE_CROSS, st);
    if (e != cudaSuccess) return e;
    if (var_tiles) *var_tiles = nt;
    if (keep && var_tiles && (long long)mt * nt * B < 148) {
        // Few candidates (the inner loop of optimize_acqf): a grid of 128x128 tiles would occupy a handful of SMs for the whole
        // k-range.  V comes from the generic GEMM (which cuts such launches into 64x64 / 32x32 tiles) + one row-reduction pass.
        GemmParams g{};
        g.A = w.Kstar; g.B = gp->Linv; g.C = w.V; g.Ct = nullptr;
        g.lda = g.ldb = g.ldc = np; g.bsA = (long long)rows_alloc * np; g.bsB = (long long)np * np; g.bsC = (long long)rows_alloc * np;
        g.mtiles = mt; g.ntiles = nt; g.kblocks = np / BK; g.krange = KR_LE_N; g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
        // single GP: the long half of the columns is cut at K/2 (second parts -> P2, added back by the row reduction): the launch
        // was as long as its longest k-range (402 us for 512 candidates at n = 4096 with the DMMA pipe 93 % busy WHILE ACTIVE)
        const bool split = B == 1 && nt % 2 == 0 && nt >= 4;
        const double* add = nullptr;
        if (split) { g.krange = KR_LE_N_SPLIT; g.C2 = w.P2; g.ldc2 = np / 2; g.c2_col0 = np / 2; add = w.P2; }
        tl_mark(gp, HDBO_PHASE_VAR, st);
        e = launch_gemm_nt(g, B, st);
        if (e == cudaSuccess)
            e = launch_row_sumsq(w.V, np, (long long)rows_alloc * np, rp, np, add, np / 2, np / 2, w.varpart, 2LL * nt * rows_alloc, B, st);
        tl_mark(gp, HDBO_PHASE_VAR, st);
        *var_tiles = 1;
        return e;
    }
    VarParams v{};
    v.Kstar = w.Kstar; v.Linv = gp->Linv; v.V = keep ? w.V : nullptr; v.varpart = w.varpart;
    v.ldk = np; v.ldl = np; v.ldv = np;
    v.bsK = (long long)rows_alloc * np; v.bsL = (long long)np * np; v.bsV = (long long)rows_alloc * np;
    v.bsvp = 2LL * nt * rows_alloc; v.mtiles = mt; v.ntiles = nt;
    tl_mark(gp, HDBO_PHASE_VAR, st);
    e = launch_var(v, B, st);
    tl_mark(gp, HDBO_PHASE_VAR, st);
    return e;
}
}  // namespace

extern "C" size_t hdbo_posterior_workspace_bytes(const hdbo_gp* gp, int64_t rows, int keep_grad_state) {
    if (check_gp(gp) || rows <= 0 || rows % 128) return 0;
    return carve(gp, rows, keep_grad_state, nullptr).bytes;
}

extern "C" double* hdbo_posterior_workspace_V(const hdbo_gp* gp, int64_t rows, void* workspace) {
    if (check_gp(gp) || rows <= 0 || rows % 128 || !workspace) return nullptr;
    return carve(gp, rows, 1, workspace).V;
}

static int posterior_acq_impl(const hdbo_gp* gp, const I8State* i8, const double* Z, int64_t ldz, int64_t m, int observation_noise,
                              int acq_kind, double acq_param, double* mean, double* var, double* acq, int64_t bso,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (m < 0) return -4;
    if (m == 0) return 0;   /* empty pool: nothing to do (Z may be NULL) */
    if (!Z || ldz < gp->d) return -2;
    if (!workspace) return -12;
    cudaStream_t st = (cudaStream_t)stream_;
    // largest multiple of 128 rows that fits the workspace
    int64_t rows = rup(m, 128);
    const size_t per128 = carve(gp, 128, 0, nullptr).bytes;
    if (per128 > workspace_bytes) return -13;
    while (rows > 128 && carve(gp, rows, 0, nullptr).bytes > workspace_bytes) {
        int64_t guess = (int64_t)(workspace_bytes / per128) * 128;
        rows = (guess < rows) ? guess : rows - 128;
        if (rows < 128) rows = 128;
    }
    PostWs w = carve(gp, rows, 0, workspace);
    const int nt = gp->n_pad / 128, B = gp->batch;
    for (int64_t off = 0; off < m; off += rows) {
        const int64_t mc = (m - off < rows) ? (m - off) : rows;
        TRY_CUDA(posterior_chunk(gp, w, rows, Z + off * ldz, ldz, mc, 0, st, i8));
        tl_mark(gp, HDBO_PHASE_FINALIZE, st);
        TRY_CUDA(launch_finalize_acq(w.meanpart, w.varpart, (long long)nt * rows, i8 ? 2 * nt : nt, i8 ? 2 * nt : nt, (int)rup(mc, 128), (int)mc, gp->hyp,
                                     observation_noise, acq_kind, acq_param, mean ? mean + off : nullptr,
                                     var ? var + off : nullptr, (acq && acq_kind >= 0) ? acq + off : nullptr, bso, B, st));
        tl_mark(gp, HDBO_PHASE_FINALIZE, st);
    }
    return 0;
}

extern "C" int hdbo_posterior_acq(const hdbo_gp* gp, const double* Z, int64_t ldz, int64_t m, int observation_noise,
                                  int acq_kind, double acq_param, double* mean, double* var, double* acq, int64_t bso,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    return posterior_acq_impl(gp, nullptr, Z, ldz, m, observation_noise, acq_kind, acq_param, mean, var, acq, bso, workspace,
                              workspace_bytes, stream_);
}

// ---- exact int8 (tcgen05) variance path: digit planes of L^-1 once per fit, then the pool evaluation ----
extern "C" size_t hdbo_i8_state_bytes(const hdbo_gp* gp) {
    if (check_gp(gp) || gp->n_pad > I8_MAX_N_PAD || gp->d > I8_MAX_D) return 0;   // int32 group sums: see i8.cuh
    return carve_i8(gp, nullptr).bytes;
}

extern "C" int32_t* hdbo_i8_state_flag(const hdbo_gp* gp, void* i8_state) {
    if (check_gp(gp) || !i8_state) return nullptr;
    return carve_i8(gp, i8_state).flag;
}

extern "C" int hdbo_gp_slice_linv(const hdbo_gp* gp, void* i8_state, void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (gp->n_pad > I8_MAX_N_PAD || gp->d > I8_MAX_D) return -1;
    if (!i8_state || ((uintptr_t)i8_state & 127)) return -2;
    cudaStream_t st = (cudaStream_t)stream_;
    I8State s = carve_i8(gp, i8_state);
    TRY_CUDA(cudaMemsetAsync(s.flag, 0, 16, st));
    TRY_CUDA(launch_slice_linv(gp->Linv, gp->n_pad, gp->n_pad, s.digits, s.scale, gp->batch, st));
    TRY_CUDA(launch_slice_rows(64, gp->X, gp->ldx, gp->n, gp->n_pad, gp->d, gp->center, gp->inv_ls, s.Xdig, s.xscale, nullptr,
                               gp->n_pad, gp->batch, st));
    return 0;
}

extern "C" int hdbo_posterior_acq_i8(const hdbo_gp* gp, void* i8_state, const double* Z, int64_t ldz, int64_t m,
                                     int observation_noise, int acq_kind, double acq_param, double* mean, double* var, double* acq,
                                     int64_t bso, void* workspace, size_t workspace_bytes, void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (gp->n_pad > I8_MAX_N_PAD || gp->d > I8_MAX_D) return -1;
    if (!i8_state || ((uintptr_t)i8_state & 127)) return -2;
    if (workspace && ((uintptr_t)workspace & 127)) return -13;
    I8State s = carve_i8(gp, i8_state);
    return posterior_acq_impl(gp, &s, Z, ldz, m, observation_noise, acq_kind, acq_param, mean, var, acq, bso, workspace,
                              workspace_bytes, stream_);
}

// ---- batched device L-BFGS ----
static int lbfgs_fill(LbfgsParams& p, const hdbo_lbfgs* s) {
    if (!s || s->R < 0 || s->D <= 0 || s->M <= 0 || s->M > 32 || s->T <= 0 || s->T > 16) return -1;
    if (!s->lo || !s->hi || !s->x || !s->f || !s->g || !s->S || !s->Y || !s->rho || !s->meta || !s->Xt || !s->step) return -1;
    if ((size_t)(s->D + 16 + s->M) * sizeof(double) > 200 * 1024) return -1;
    p.R = s->R; p.D = s->D; p.M = s->M; p.T = s->T; p.lo = s->lo; p.hi = s->hi; p.x = s->x; p.f = s->f; p.g = s->g;
    p.S = s->S; p.Y = s->Y; p.rho = s->rho; p.step = s->step; p.meta = s->meta; p.Xt = s->Xt; p.ft = s->ft; p.gt = s->gt;
    p.ftol = s->ftol; p.gtol = s->gtol; p.maxiter = s->maxiter;
    return 0;
}

extern "C" int hdbo_lbfgs_propose(const hdbo_lbfgs* s, void* stream_) {
    LbfgsParams p{};
    if (int e = lbfgs_fill(p, s)) return e;
    TRY_CUDA(launch_lbfgs_propose(p, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_lbfgs_accept(const hdbo_lbfgs* s, void* stream_) {
    LbfgsParams p{};
    if (int e = lbfgs_fill(p, s)) return e;
    if (!s->ft || !s->gt) return -1;
    TRY_CUDA(launch_lbfgs_accept(p, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_acq_elementwise(const double* mean, const double* var, int64_t m, int acq_kind, double acq_param,
                                    double* acq, double* dmean, double* dvar, void* stream_) {
    if (!mean || !var) return -1;
    if (m < 0) return -3;
    if (acq_kind < 0 || acq_kind > 4) return -4;
    TRY_CUDA(launch_acq_elem(mean, var, m, acq_kind, acq_param, acq, dmean, dvar, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_linear_f32(const float* X, int64_t ldx, int32_t b, int32_t K, const float* W, const float* bias, const float* scale,
                               const float* shift, int32_t relu, float* Y, int64_t ldy, int32_t N, void* stream_) {
    if (!X) return -1;
    if (b < 0) return -3;
    if (K <= 0 || ldx < K) return -4;
    if (!W) return -5;
    if (!Y) return -10;
    if (N < 0 || ldy < N) return -12;
    TRY_CUDA(launch_linear_f32(X, ldx, b, K, W, bias, scale, shift, relu, Y, ldy, N, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_argmax_tokens_f32(const float* logits, int64_t ld, int32_t b, int32_t L, int32_t A, int32_t* tokens, void* stream_) {
    if (!logits) return -1;
    if (b < 0) return -3;
    if (L <= 0 || A <= 0 || ld < (int64_t)L * A) return -2;
    if (!tokens) return -6;
    TRY_CUDA(launch_argmax_tokens_f32(logits, ld, b, L, A, tokens, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_argmax(const double* v, int64_t m, void* work, double* out_val, int64_t* out_idx, void* stream_) {
    if (!v || m <= 0) return -1;
    if (!work) return -3;
    TRY_CUDA(launch_argmax(v, m, work, out_val, (long long*)out_idx, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_dgemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                             int64_t m_pad, int64_t n_pad, int64_t k_pad, double alpha, double beta, int krange,
                             int lower_only, void* stream_) {
    if (!A || !B || !C) return -1;
    if (m_pad % 128 || n_pad % 128 || k_pad % 16 || lda % 2 || ldb % 2 || ldc % 2) return -7;
    if (krange < 0 || krange > KR_GE_MN) return -12;
    GemmParams g{};
    g.A = A; g.B = B; g.C = C; g.Ct = nullptr;
    g.lda = (int)lda; g.ldb = (int)ldb; g.ldc = (int)ldc; g.ldct = 0;
    g.mtiles = (int)(m_pad / 128); g.ntiles = (int)(n_pad / 128); g.kblocks = (int)(k_pad / 16);
    g.krange = krange; g.lower_only = lower_only; g.alpha = alpha; g.beta = beta;
    TRY_CUDA(launch_gemm_nt(g, 1, (cudaStream_t)stream_));
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
extern "C" int hdbo_mll_grad(const hdbo_gp* gp, double* Kinv, double* G, double* XsT, double* part, double* grad,
                             void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (!gp->R2) return -1;
    if (!Kinv) return -2;
    if (!G) return -3;
    if (!XsT) return -4;
    if (!part) return -5;
    if (!grad) return -6;
    tl_mark(gp, HDBO_PHASE_MLL_GRAD, (cudaStream_t)stream_);
    TRY_CUDA(mll_grad(gp, Kinv, G, XsT, part, grad, (cudaStream_t)stream_));
    tl_mark(gp, HDBO_PHASE_MLL_GRAD, (cudaStream_t)stream_);
    return 0;
}

extern "C" size_t hdbo_mll_grad_part_bytes(const hdbo_gp* gp) {
    if (check_gp(gp)) return 0;
    return mll_grad_part_doubles(gp) * sizeof(double);
}

extern "C" int hdbo_posterior_fwd(const hdbo_gp* gp, const double* Z, int64_t ldz, int64_t m, int observation_noise,
                                  double* mean, double* var, int64_t bso, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (!Z || ldz < gp->d) return -2;
    if (m <= 0) return -4;
    const int64_t rows = rup(m, 128);
    if (!workspace || carve(gp, rows, 1, nullptr).bytes > workspace_bytes) return -9;
    cudaStream_t st = (cudaStream_t)stream_;
    PostWs w = carve(gp, rows, 1, workspace);
    const int nt = gp->n_pad / 128;
    int vt = nt;
    TRY_CUDA(posterior_chunk(gp, w, rows, Z, ldz, m, 1, st, nullptr, &vt));
    TRY_CUDA(launch_finalize_acq(w.meanpart, w.varpart, (long long)nt * rows, nt, vt, (int)rows, (int)m, gp->hyp,
                                 observation_noise, HDBO_ACQ_NONE, 0.0, mean, var, nullptr, bso, gp->batch, st));
    return 0;
}

extern "C" int hdbo_posterior_joint_cov(const hdbo_gp* gp, int64_t m, int q, int observation_noise, double* Sigma,
                                        int64_t bsS, void* workspace, size_t workspace_bytes, void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (m <= 0) return -2;
    if (q < 1 || q > 8 || m % q) return -3;
    if (!Sigma) return -5;
    const int64_t rows = rup(m, 128);
    if (!workspace || carve(gp, rows, 1, nullptr).bytes > workspace_bytes) return -7;
    PostWs w = carve(gp, rows, 1, workspace);
    TRY_CUDA(launch_joint_cov(w.Zs, gp->d_pad, (long long)rows * gp->d_pad, w.V, gp->n_pad, (long long)rows * gp->n_pad,
                              gp->hyp, (int)(m / q), q, gp->kind, observation_noise, Sigma, bsS, gp->batch,
                              (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_posterior_bwd(const hdbo_gp* gp, int64_t m, int q, const double* gmean, const double* gvar,
                                  const double* gSigma, int64_t bsg, int64_t bsS, double* dZ, int64_t lddz, int64_t bsdz,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    if (int e = check_gp(gp)) return e;
    if (m <= 0) return -2;
    if (q < 1 || q > 8 || m % q) return -3;
    if (gvar && gSigma) return -5;
    if (!dZ || lddz < gp->d) return -9;
    const int64_t rows = rup(m, 128);
    if (!workspace || carve(gp, rows, 1, nullptr).bytes > workspace_bytes) return -12;
    PostWs w = carve(gp, rows, 1, workspace);
    BwdBuffers b{w.Zs, w.Kstar, w.DK, w.V, w.T, w.XsT, w.meanpart, w.P2};
    tl_mark(gp, HDBO_PHASE_BWD, (cudaStream_t)stream_);
    TRY_CUDA(posterior_bwd(gp, b, rows, m, q, gmean, gvar, gSigma, bsg, bsS, dZ, (int)lddz, bsdz, (cudaStream_t)stream_));
    tl_mark(gp, HDBO_PHASE_BWD, (cudaStream_t)stream_);
    return 0;
}

extern "C" int hdbo_qei(const double* mean, const double* Sigma, const double* base_samples, int64_t R, int q, int64_t S,
                        double best_f, double jitter, double* value, double* gmean, double* gSigma, int32_t* info,
                        void* stream_) {
    if (!mean) return -1;
    if (!Sigma) return -2;
    if (!base_samples) return -3;
    if (R < 0) return -4;
    if (q < 1 || q > 8) return -5;
    if (S <= 0) return -6;
    if (!value) return -9;
    TRY_CUDA(launch_qei(mean, Sigma, base_samples, R, q, S, best_f, jitter, value, gmean, gSigma, info, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_sobol_draw(const int32_t* state_t, const int32_t* shift, int dim, int64_t first, int64_t m, const double* lower,
                               const double* range, double* out, int64_t ld, int first_point_fp32, void* stream_) {
    if (!state_t) return -1;
    if (!shift) return -2;
    if (dim <= 0) return -3;
    if (first < 0 || first + m > (1LL << 30)) return -4;
    if (m < 0) return -5;
    if (m == 0) return 0;   /* empty draw: nothing to do (out may be NULL) */
    if (!out || ld < dim) return -8;
    TRY_CUDA(launch_sobol_draw(state_t, shift, dim, first, m, lower, range, out, ld, first_point_fp32, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_mvn_sample(const double* mean, const double* Sigma, const double* z, int64_t R, int q, int64_t S, int z_per_r,
                               double jitter, double* out, int32_t* info, void* stream_) {
    if (!mean) return -1;
    if (!Sigma) return -2;
    if (!z) return -3;
    if (R < 0) return -4;
    if (q < 1 || q > 8) return -5;
    if (S < 0) return -6;
    if (!out) return -9;
    TRY_CUDA(launch_mvn_sample(mean, Sigma, z, R, q, S, z_per_r, jitter, out, info, (cudaStream_t)stream_));
    return 0;
}

extern "C" int hdbo_cdist(const double* X, int64_t ldx, int n, int d, double* Xs, double* xn, double* D_pad, void* stream_) {
    if (!X || ldx < d) return -1;
    if (n <= 0 || d <= 0) return -3;
    if (!Xs || !xn || !D_pad) return -5;
    cudaStream_t st = (cudaStream_t)stream_;
    const int np = (int)rup(n, 128), dp = (int)rup(d, 16);
    TRY_CUDA(launch_scale_inputs(X, (int)ldx, 0, n, d, nullptr, nullptr, 0, Xs, dp, np, 0, xn, 0, 1, st));
    GemmParams g{};
    g.A = Xs; g.B = Xs; g.C = D_pad; g.Ct = nullptr;
    g.lda = g.ldb = dp; g.ldc = np; g.mtiles = g.ntiles = np / 128; g.kblocks = dp / BK; g.krange = KR_FULL;
    g.lower_only = 0; g.alpha = 1.0; g.beta = 0.0;
    TRY_CUDA(launch_gemm_nt(g, 1, st));
    TRY_CUDA(launch_cdist_finish(D_pad, np, xn, np, st));
    return 0;
}
