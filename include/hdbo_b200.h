/* libhdbo_b200.so -- C ABI of the B200-native GP-surrogate + acquisition hot path.
 *
 * This is the drop-in boundary.  The reference (MachineLearningLifeScience/hdbo_benchmark) has no FFI of its own:
 * its hot path is reached through the BoTorch / GPyTorch Python API (un-vendored), at the call sites
 *   /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-194   (solvers -> SingleTaskGP,
 *                                                  fit_gpytorch_mll, (Log)EI, optimize_acqf, SAAS)
 *   /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:124-145   (gpytorch kernel constructor)
 *   /root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:49-145    (ExactMarginalLogLikelihood,
 *                                                  model(train_x), eval-mode model(test_x))
 *   /root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:11-29     (pairwise cdist de-duplication)
 * Each entry point below names the upstream operation it replaces.  The Python host layer
 * (hdbo_benchmark_b200/) binds these with ctypes and re-exposes the BoTorch surface; see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to FP64 (or as typed), owned by the caller; the library never allocates,
 *    frees or retains memory and has no global state (kernel attributes are set per launch; timing events, when wanted,
 *    live in a caller-owned hdbo_timeline);
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit synchronisation;
 *  - return value: 0 ok; > 0 a cudaError_t from a launch; < 0 the (negated, 1-based) index of an invalid argument;
 *  - matrices are row-major.  Internal ("state") matrices are padded: n_pad = roundup(n,128), d_pad = roundup(d,16),
 *    leading dimension = padded width; the padding of Khat is the identity, everything else zero;
 *  - `batch` independent GPs (SAAS hyper-parameter samples) are laid out with explicit element strides.
 */
#ifndef HDBO_B200_H
#define HDBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDBO_KERNEL_MATERN52 0
#define HDBO_KERNEL_RBF 1

#define HDBO_ACQ_EI 0
#define HDBO_ACQ_LOGEI 1
#define HDBO_ACQ_UCB 2
#define HDBO_ACQ_MEAN 3
#define HDBO_ACQ_PI 4
#define HDBO_ACQ_NONE (-1)

/* Optional CALLER-OWNED timeline (bench.py's per-kernel roofline figures).  When hdbo_gp.timeline is non-NULL the entry points
 * record events[used], events[used+1] (begin / end) on the caller's stream around each phase they enqueue, write the phase id of
 * both slots to phase[] and advance `used`; they stop recording when the slots run out.  The caller creates the cudaEvent_t
 * handles (timing enabled), resets `used`, and reads the elapsed times itself: the library keeps no profiling state. */
#define HDBO_PHASE_CROSS 0      /* cross-kernel build K*(Z, X) (+ mean partials)                */
#define HDBO_PHASE_VAR 1        /* variance contraction V = K* L^-T                              */
#define HDBO_PHASE_PREP 2       /* candidate centring / scaling / digit slicing                  */
#define HDBO_PHASE_FINALIZE 3   /* partial sums -> mean, variance, acquisition                   */
#define HDBO_PHASE_FIT_BUILD 4  /* training-input scaling + K(X, X) + noise I                    */
#define HDBO_PHASE_FIT_FACTOR 5 /* Cholesky factor + inverse                                     */
#define HDBO_PHASE_FIT_SOLVE 6  /* w, alpha, logdet, quadratic form                              */
#define HDBO_PHASE_MLL_GRAD 7   /* K^-1, G, lengthscale / outputscale / noise / mean gradients   */
#define HDBO_PHASE_BWD 8        /* posterior backward (dZ)                                       */
typedef struct hdbo_timeline {
    void** events;          /* HOST array [capacity] of cudaEvent_t                                                    */
    int32_t* phase;         /* HOST array [capacity] (may be NULL)                                                     */
    int32_t capacity, used;
} hdbo_timeline;

/* One (batch of) exact GP(s).  hyp[b] = {outputscale, noise, constant mean, jitter}. */
typedef struct hdbo_gp {
    int32_t n, d, n_pad, d_pad, kind, batch;
    const double* X;        /* [n][d] raw training inputs (shared by the batch), leading dimension ldx          */
    int64_t ldx;
    const double* y;        /* [n] targets (shared by the batch)                                                */
    const double* center;   /* [d] centring vector (training mean)                                              */
    const double* inv_ls;   /* [batch][d] reciprocal ARD lengthscales                                           */
    const double* hyp;      /* [batch][4]                                                                       */
    double* Xs;             /* [batch][n_pad][d_pad] centred, scaled inputs                                     */
    double* xn;             /* [batch][n_pad]        squared norms of the rows of Xs                            */
    double* R2;             /* [batch][n_pad][n_pad] MLL-gradient state (may be NULL when no gradient is wanted): the    */
                            /* kernel build leaves dk/dr^2 (unit outputscale) on the lower tiles (historical field name) */
    double* Awork;          /* [batch][n_pad][n_pad] K + noise I, destroyed by the factorisation                */
    double* L;              /* [batch][n_pad][n_pad] Cholesky factor (lower)                                    */
    double* Linv;           /* [batch][n_pad][n_pad] L^-1 (lower)                                               */
    double* Linvt;          /* [batch][n_pad][n_pad] L^-T (upper)                                               */
    double* w;              /* [batch][n_pad]        L^-1 (y - m)                                               */
    double* alpha;          /* [batch][n_pad]        Khat^-1 (y - m)                                            */
    double* scal;           /* [batch][8]: logdet, quad, n*mll data term, sum(alpha), ...                       */
    int32_t* info;          /* [batch] 0 or 1-based index of the first non-positive pivot                       */
    hdbo_timeline* timeline;/* HOST pointer to a caller-owned timeline, or NULL (the default: nothing is recorded)      */
    double* Kinv;           /* [batch][n_pad][n_pad] or NULL: when set, hdbo_gp_fit(need_r2 != 0) also leaves the lower  */
                            /* tiles of Khat^-1 = L^-T L^-1 here (consumed by hdbo_mll_grad when passed as its Kinv)     */
    void* aux_streams[4];   /* optional CALLER-OWNED side streams (cudaStream_t): chain (create it with the highest      */
                            /* priority), side (lowest), near, panel.  All four + the 12 events set, batch == 1, n_pad >= 1024: */
                            /* hdbo_gp_fit runs the pipelined factorisation (csrc/kernels_chol_pipe.cu) forked from and   */
                            /* joined back to `stream`; otherwise everything runs on `stream` alone.                      */
    void* aux_events[12];   /* CALLER-OWNED cudaEvent_t (timing disabled) for the cross-stream dependencies               */
} hdbo_gp;

/* library / build identification */
int hdbo_version(void);

/* (hdbo_gp.Kinv / aux_streams / aux_events are optional: zero them when unused.)
 * Replaces: kernel build (gpytorch MaternKernel/RBFKernel/ScaleKernel forward on train x train),
 * GaussianLikelihood noise add, linear_operator psd_safe_cholesky (one attempt; the host applies the jitter
 * schedule by re-calling with hyp[3] = jitter), cholesky_solve for alpha, and the log-determinant.
 * need_r2 != 0 also keeps the squared-distance matrix for hdbo_mll_grad. */
int hdbo_gp_fit(const hdbo_gp* gp, int need_r2, void* stream);

/* Same, but factorises K + noise I - V V^T (V: [n_pad][k_pad] row-major, leading dimension ldv, padding rows zero;
 * batch == 1).  With X = test inputs, y = test targets - posterior mean and V from hdbo_posterior_fwd this yields the
 * joint predictive log-density that gpytorch's eval-mode `mll(model(test_x), test_y)` computes
 * (/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:77-94): scal[2] = log N(y | mu, Sigma). */
int hdbo_gp_fit_downdated(const hdbo_gp* gp, int need_r2, const double* V, int64_t ldv, int64_t k_pad, void* stream);

/* Replaces: autograd backward of gpytorch ExactMarginalLogLikelihood (data term; priors are added on the host).
 * Inputs: a fitted hdbo_gp with R2.  Kinv, G: [batch][n_pad][n_pad] scratch (Kinv may alias gp->L).
 * XsT: [batch][d_pad_t][n_pad] scratch, d_pad_t = roundup(d,128).  part: hdbo_mll_grad_part_bytes(gp) bytes of scratch.
 * grad: [batch][d + 3] = d(n*mll_data)/d{lengthscale[d], outputscale, noise, mean}. */
int hdbo_mll_grad(const hdbo_gp* gp, double* Kinv, double* G, double* XsT, double* part, double* grad, void* stream);
size_t hdbo_mll_grad_part_bytes(const hdbo_gp* gp);

/* Workspace (bytes) for hdbo_posterior_acq with `rows` (multiple of 128) candidates resident at a time. */
size_t hdbo_posterior_workspace_bytes(const hdbo_gp* gp, int64_t rows, int keep_grad_state);

/* Address of V = K* L^-T ([rows][n_pad], leading dimension n_pad) inside a keep_grad_state workspace of `rows` rows,
 * valid after hdbo_posterior_fwd. */
double* hdbo_posterior_workspace_V(const hdbo_gp* gp, int64_t rows, void* workspace);

/* Replaces: model.posterior(X).mean/.variance for X of shape [m,1,d] followed by an analytic acquisition
 * (botorch ExpectedImprovement / LogExpectedImprovement / UpperConfidenceBound forward), no-grad candidate-pool
 * evaluation as in optimize_acqf's raw-sample screening.  Z [m][d] (ldz), shared by the batch.
 * Outputs (each may be NULL): mean, var, acq: [batch][m] with stride bso.  Processes the pool in chunks that
 * fit `workspace`. */
int hdbo_posterior_acq(const hdbo_gp* gp, const double* Z, int64_t ldz, int64_t m, int observation_noise, int acq_kind,
                       double acq_param, double* mean, double* var, double* acq, int64_t bso, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Exact int8 mode (n_pad <= 8192, d <= 16384, any batch): the same call as hdbo_posterior_acq, with both GEMMs (Zs Xs^T and V = K* L^-T) computed
 * on the tcgen05 tensor cores from base-256 digit planes (6 x 6 digits, 21 exact int8 GEMMs accumulated in TMEM, FP64
 * recombination; csrc/kernels_i8.cu) instead of FP64 DMMA.  Results agree with hdbo_posterior_acq to ~1e-12 of the prior
 * variance (tests/test_gpu_i8.py); the absolute error grows like (outputscale / noise)^1/2, so for the 1e-6 RELATIVE bar on
 * candidates that coincide with training points use it when noise >= 1e-4 outputscale (the Python engine enforces this).
 *   hdbo_i8_state_bytes     size of the caller-owned state buffer (128-byte aligned): digit planes of L^-1 + row scales + flag
 *   hdbo_gp_slice_linv      after every hdbo_gp_fit: slices gp->Linv into the state
 *   hdbo_posterior_acq_i8   pool evaluation; `workspace` as for hdbo_posterior_acq (keep_grad_state = 0), 128-byte aligned
 *   hdbo_i8_state_flag      device int32 inside the state: non-zero if a pipeline wait timed out (outputs are then NaN) */
size_t hdbo_i8_state_bytes(const hdbo_gp* gp);
int hdbo_gp_slice_linv(const hdbo_gp* gp, void* i8_state, void* stream);
int hdbo_posterior_acq_i8(const hdbo_gp* gp, void* i8_state, const double* Z, int64_t ldz, int64_t m, int observation_noise,
                          int acq_kind, double acq_param, double* mean, double* var, double* acq, int64_t bso, void* workspace,
                          size_t workspace_bytes, void* stream);
int32_t* hdbo_i8_state_flag(const hdbo_gp* gp, void* i8_state);

/* Forward of the differentiable posterior for m <= rows candidates, keeping Kstar, dKstar_dr2 and V in the workspace
 * (layout fixed by hdbo_posterior_workspace_bytes(keep_grad_state=1)) for hdbo_posterior_bwd.
 * Replaces: model.posterior(X) under autograd. */
int hdbo_posterior_fwd(const hdbo_gp* gp, const double* Z, int64_t ldz, int64_t m, int observation_noise, double* mean,
                       double* var, int64_t bso, void* workspace, size_t workspace_bytes, void* stream);

/* Joint covariance for q-batches after hdbo_posterior_fwd: Z viewed as [m/q][q][d];
 * Sigma [batch][m/q][q][q] = k(Zq,Zq) - V V^T (+ noise I).  Replaces: posterior.mvn.covariance_matrix for q > 1. */
int hdbo_posterior_joint_cov(const hdbo_gp* gp, int64_t m, int q, int observation_noise, double* Sigma, int64_t bsS,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Backward: given d loss/d mean [batch][m], d loss/d var [batch][m] (gSigma NULL) or d loss/d Sigma
 * [batch][m/q][q][q] (q > 1; the diagonal carries d/d var), returns dZ [batch][m][d] (row stride lddz, batch stride
 * bsdz); the caller sums over the batch for mixture acquisitions.  Must follow hdbo_posterior_fwd on the same
 * workspace and m.  Replaces: autograd through posterior(X) (A6). */
int hdbo_posterior_bwd(const hdbo_gp* gp, int64_t m, int q, const double* gmean, const double* gvar, const double* gSigma,
                       int64_t bsg, int64_t bsS, double* dZ, int64_t lddz, int64_t bsdz, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Replaces: botorch analytic acquisition forward + autograd partials on given (mean, var). */
int hdbo_acq_elementwise(const double* mean, const double* var, int64_t m, int acq_kind, double acq_param, double* acq,
                         double* dacq_dmean, double* dacq_dvar, void* stream);

/* Replaces: botorch qExpectedImprovement.forward (+ backward): mean [R][q], Sigma [R][q][q], base samples [S][q]
 * (SobolQMCNormalSampler, fixed), jitter added to diag(Sigma) before its Cholesky.  q <= 8.
 * Outputs: value [R]; gmean [R][q], gSigma [R][q][q] (may be NULL).  info [R]: q x q Cholesky failure flag. */
int hdbo_qei(const double* mean, const double* Sigma, const double* base_samples, int64_t R, int q, int64_t S,
             double best_f, double jitter, double* value, double* gmean, double* gSigma, int32_t* info, void* stream);

/* Replaces: torch.quasirandom.SobolEngine(dim, scramble=True, seed).draw(m) + `lower + range * u` inside
 * botorch.utils.sampling.draw_sobol_samples, i.e. the raw-sample pool of gen_batch_initial_conditions (optimize_acqf),
 * generated in HBM instead of on the host: points first .. first+m-1 of the sequence, bit-identical to torch's engine.
 * state_t [30][dim] = the engine's (scrambled) direction numbers transposed, shift [dim] = its digital shift, both int32
 * (torch keeps them as int64 < 2^30); lower / range [dim] or NULL (unit cube); out [m][ld].  first_point_fp32 != 0 reproduces
 * torch's point 0, which SobolEngine computes in the process-wide default dtype (float32 unless changed) before casting. */
int hdbo_sobol_draw(const int32_t* state_t, const int32_t* shift, int dim, int64_t first, int64_t m, const double* lower,
                    const double* range, double* out, int64_t ld, int first_point_fp32, void* stream);

/* Replaces: GPyTorchPosterior.rsample_from_base_samples on a joint q-point posterior (q <= 8): psd_safe_cholesky of the q x q
 * covariance (+ jitter I) and  out[s][r][:] = mean[r][:] + L_r z[s][r][:]  (z_per_r != 0) or L_r z[s][:] (z shared by the R
 * groups).  mean [R][q], Sigma [R][q][q], z [S][R][q] or [S][q], out [S][R][q]; info [R] (may be NULL): first non-positive pivot. */
int hdbo_mvn_sample(const double* mean, const double* Sigma, const double* z, int64_t R, int q, int64_t S, int z_per_r,
                    double jitter, double* out, int32_t* info, void* stream);

/* argmax with first-index tie-break (torch.argmax semantics on a 1-d tensor).  work: >= 8192 bytes, zeroed once. */
int hdbo_argmax(const double* v, int64_t m, void* work, double* out_val, int64_t* out_idx, void* stream);

/* Generic FP64 C = alpha * A * B^T + beta * C on padded operands (rows multiples of 128, K multiple of 16), optional triangular
 * k-range.  Replaces: `mean + base_samples @ L_Sigma^T` of MultivariateNormal.rsample in the Thompson-sampling path of TuRBO / BAxUS
 * (hdbo_benchmark_b200/thompson.py; reference selection: load_solvers.py:224-246). */
int hdbo_dgemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t m_pad,
                  int64_t n_pad, int64_t k_pad, double alpha, double beta, int krange, int lower_only, void* stream);

/* Replaces: torch.cdist(x, x) in filter_close_points (training_gps.py:11-13): D [n][n] Euclidean distances.
 * Xs/xn/work as produced internally: caller passes scratch Xs [n_pad][d_pad], xn [n_pad], D_pad [n_pad][n_pad]. */
int hdbo_cdist(const double* X, int64_t ldx, int n, int d, double* Xs, double* xn, double* D_pad, void* stream);

/* Batched box-constrained L-BFGS on the device: R independent problems (the restarts of optimize_acqf), minimisation.
 * Replaces: scipy.optimize.minimize(method="L-BFGS-B") inside botorch.generation.gen_candidates_scipy, whose per-iteration
 * float64-ndarray round trip dominates `next_candidate()` once the acquisition itself takes microseconds.
 * One iteration = hdbo_lbfgs_propose (direction from the (s, y) ring, T trial points clip(x + step 2^(2-j) dir): two expansions, the
 * last accepted length, then halvings) -> the caller evaluates value + gradient at the R x T trial points (hdbo_posterior_fwd /
 * hdbo_acq_elementwise or hdbo_posterior_joint_cov + hdbo_qei / hdbo_posterior_bwd) into ft / gt -> hdbo_lbfgs_accept (the lowest
 * trial that satisfies the strong Wolfe conditions, else the lowest Armijo trial; ring update, convergence flags in meta).  No host
 * synchronisation is needed until the caller polls meta[r][2].  Initial state: x, f, g at the starting points, meta zeroed,
 * step = 1. */
typedef struct hdbo_lbfgs {
    int32_t R, D, M, T;          /* problems, variables each, history pairs (<= 32), trial steps (<= 16)                 */
    const double *lo, *hi;       /* [D] box shared by the problems (+-inf = unbounded)                                  */
    double *x, *f, *g;           /* [R][D], [R], [R][D]                                                                 */
    double *S, *Y, *rho;         /* [R][M][D], [R][M][D], [R][M]                                                        */
    double* step;                /* [R] first trial step length of the next line search (initialise to 1)              */
    int32_t* meta;               /* [R][4]: pairs stored, newest index, done (1 converged, 2 stalled), iterations       */
    double* Xt;                  /* [R][T][D] trial points                                                              */
    const double *ft, *gt;       /* [R][T], [R][T][D] value / gradient at the trial points                              */
    double ftol, gtol;           /* relative decrease / projected-gradient sup-norm tolerances (scipy: 2.2e-9, 1e-5)    */
    int32_t maxiter;
} hdbo_lbfgs;
int hdbo_lbfgs_propose(const hdbo_lbfgs* s, void* stream);
int hdbo_lbfgs_accept(const hdbo_lbfgs* s, void* stream);

/* ---- VAE decode next to the hot path -------------------------------------------------------------------
 * Replaces: one Linear (+ eval-mode BatchNorm1d + ReLU; Dropout is the identity in eval mode) of the decoder / encoder
 * nn.Sequential of the reference's latent-space models (src/hdbo_benchmark/generative_models/vae_selfies.py:48-81), FP32 like
 * the reference:   Y[r][n] = act((X[r] . W[n] + bias[n]) * scale[n] + shift[n]),  r < b, n < N
 * W [N][K] row-major (torch nn.Linear.weight), bias / scale / shift [N] or NULL (scale = gamma / sqrt(running_var + eps),
 * shift = beta - running_mean * scale: ATen's eval-mode batch norm), relu != 0 applies max(., 0).  X [b][ldx], Y [b][ldy]. */
int hdbo_linear_f32(const float* X, int64_t ldx, int32_t b, int32_t K, const float* W, const float* bias, const float* scale,
                    const float* shift, int32_t relu, float* Y, int64_t ldy, int32_t N, void* stream);

/* Replaces: np.argmax(logits.reshape(b, L, A), axis=-1) in VAESelfies._logits_to_string_array (vae_selfies.py:241-247):
 * tokens [b][L] (int32), first index on ties; logits [b][ld] finite FP32 with position p at columns p*A .. p*A+A-1. */
int hdbo_argmax_tokens_f32(const float* logits, int64_t ld, int32_t b, int32_t L, int32_t A, int32_t* tokens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HDBO_B200_H */
