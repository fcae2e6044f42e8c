// Parameters of the batched device L-BFGS kernels (kernels_lbfgs.cu).  All arrays are device memory owned by the caller.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdbo {

struct LbfgsParams {
    int R, D, M, T;            // problems, variables per problem, history length, trial steps per line search
    const double *lo, *hi;     // [D] box (shared by the problems); +-inf for unbounded
    double *x, *f, *g;         // [R][D], [R], [R][D]   current iterate, value, gradient (minimisation)
    double *S, *Y, *rho;       // [R][M][D], [R][M][D], [R][M]   ring of curvature pairs
    double* step;              // [R] length of the first trial step (x the L-BFGS direction), initialised to 1
    int32_t* meta;             // [R][4] = {pairs stored, index of the newest pair, done flag (1 converged / 2 stalled), iterations}
    double* Xt;                // [R][T][D]  trial points (written by propose)
    const double *ft, *gt;     // [R][T], [R][T][D]  value / gradient at the trial points (read by accept)
    double ftol, gtol;
    int maxiter;
};

cudaError_t launch_lbfgs_propose(const LbfgsParams& p, cudaStream_t st);
cudaError_t launch_lbfgs_accept(const LbfgsParams& p, cudaStream_t st);

}  // namespace hdbo
