// Declarations of the gradient-path kernels (MLL gradient, posterior backward, joint covariance, qEI, cdist).
#pragma once
#include "kernels.cuh"

namespace hdbo {

struct BwdBuffers {
    double *Zs, *Kstar, *DK, *V, *T, *XsT, *csum, *P2;
};

cudaError_t launch_transpose(const double* in, int ld_in, long long bs_in, int rows, int cols, double* out, int ld_out,
                             long long bs_out, int cols_pad, int batch, cudaStream_t st);
cudaError_t mll_grad(const hdbo_gp* gp, double* Kinv, double* G, double* XsT, double* part, double* grad, cudaStream_t st);
cudaError_t kinv_from_linvt(const hdbo_gp* gp, double* Kinv, cudaStream_t st);
size_t mll_grad_part_doubles(const hdbo_gp* gp);
cudaError_t launch_joint_cov(const double* Zs, int ldz, long long bsZ, const double* V, int ldv, long long bsV,
                             const double* hyp, int groups, int q, int kind, int observation_noise, double* Sigma,
                             long long bsS, int batch, cudaStream_t st);
cudaError_t posterior_bwd(const hdbo_gp* gp, const BwdBuffers& w, int64_t rows_alloc, int64_t m, int q, const double* gmean,
                          const double* gvar, const double* gSigma, long long bsg, long long bsS, double* dZ, int lddz,
                          long long bsdz, cudaStream_t st);
cudaError_t launch_cdist_finish(double* D, int ld, const double* xn, int n_pad, cudaStream_t st);
cudaError_t launch_qei(const double* mean, const double* Sigma, const double* base, long long R, int q, long long S,
                       double best_f, double jitter, double* value, double* gmean, double* gSigma, int32_t* info,
                       cudaStream_t st);

cudaError_t launch_mvn_sample(const double* mean, const double* Sigma, const double* z, long long R, int q, long long S,
                              int z_per_r, double jitter, double* out, int32_t* info, cudaStream_t st);

}  // namespace hdbo
