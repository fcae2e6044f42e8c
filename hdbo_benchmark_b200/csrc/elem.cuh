// Analytic acquisition functions on (mean, variance): value and partial derivatives  (SURVEY.md §8a rows A5/A6).
// Mirrors botorch.acquisition.analytic {ExpectedImprovement, LogExpectedImprovement, UpperConfidenceBound}:
// sigma = sqrt(clamp_min(var, 1e-12)), u = (mean - best_f)/sigma; LogEI uses the erfcx / log1mexp tail branches
// of botorch _log_ei_helper with the float64 thresholds (-1, -1e6).
#pragma once
#include "kernels.cuh"

namespace hdbo {

constexpr double MIN_VAR = 1e-12;

__device__ __forceinline__ double log1mexp_dev(double x) {   // log(1 - exp(x)), x < 0
    return (x > -0.69314718055994530942) ? log(-expm1(x)) : log1p(-exp(x));
}

// kind: 0 EI (param = best_f), 1 LogEI (param = best_f), 2 UCB (param = beta), 3 posterior mean, 4 -> PI (param = best_f)
__device__ __forceinline__ void acq_eval(int kind, double param, double mean, double var, double& a, double& dmean,
                                         double& dvar) {
    const bool clamped = !(var >= MIN_VAR);
    const double sigma = sqrt(fmax(var, MIN_VAR));
    const double dsig_dvar = clamped ? 0.0 : 0.5 / sigma;
    if (kind == 2) {
        const double sb = sqrt(param);
        a = mean + sb * sigma;
        dmean = 1.0;
        dvar = sb * dsig_dvar;
        return;
    }
    if (kind == 3) { a = mean; dmean = 1.0; dvar = 0.0; return; }
    const double u = (mean - param) / sigma;
    const double phi = INV_SQRT_2PI * exp(-0.5 * u * u);
    const double Phi = 0.5 * erfc(-INV_SQRT_2 * u);
    if (kind == 0) {
        a = sigma * (phi + u * Phi);
        dmean = Phi;
        dvar = phi * dsig_dvar;
        return;
    }
    if (kind == 4) {
        a = Phi;
        dmean = phi / sigma;
        dvar = -phi * u / sigma * dsig_dvar;
        return;
    }
    // LogEI
    double logh, dlogh;
    if (u > -1.0) {
        const double h = phi + u * Phi;
        logh = log(h);
        dlogh = Phi / h;
    } else {
        const double logphi = -0.5 * (u * u + LOG_2PI);
        if (u > -1e6) {
            const double R = erfcx(-INV_SQRT_2 * u);                 // Phi/phi = sqrt(pi/2) erfcx(-u/sqrt2)
            const double w = log(R * fabs(u)) + LOG_SQRT_PI_DIV_2;   // log(|u| Phi/phi) < 0
            logh = logphi + log1mexp_dev(w);
            dlogh = (R * 1.2533141373155002512078826424055) / (-expm1(w));
        } else {
            logh = logphi - 2.0 * log(fabs(u));
            dlogh = -u - 2.0 / u;
        }
    }
    a = logh + log(sigma);
    dmean = dlogh / sigma;
    const double dsig = (1.0 - dlogh * u) / sigma;
    dvar = dsig * dsig_dvar;
}

cudaError_t launch_tri_matvec(const double* M, int ld, long long bsM, int n_pad, int upper, const double* vin,
                              long long bsv, const double* y, int n, const double* hyp, double* out, long long bso,
                              int batch, cudaStream_t stream);
cudaError_t launch_fit_scalars(const double* L, int ld, long long bsL, const double* w, const double* alpha,
                               long long bsv, int n, int n_pad, double* scal, int batch, cudaStream_t stream);
cudaError_t launch_row_sumsq(double* V, int ldv, long long bsV, int rows, int cols, const double* add, int ldadd, int add_from,
                             double* out, long long bso, int batch, cudaStream_t stream);
cudaError_t launch_finalize_acq(const double* meanpart, const double* varpart, long long bsp, int ntiles, int ntiles_var, int rows_pad,
                                int m, const double* hyp, int observation_noise, int acq_kind, double acq_param,
                                double* mean_out, double* var_out, double* acq_out, long long bso, int batch,
                                cudaStream_t stream);
cudaError_t launch_acq_elem(const double* mean, const double* var, long long m, int acq_kind, double acq_param, double* acq,
                            double* dmean, double* dvar, cudaStream_t stream);
cudaError_t launch_argmax(const double* v, long long m, void* work, double* out_val, long long* out_idx, cudaStream_t stream);

cudaError_t launch_sobol_draw(const int* state_t, const int* shift, int dim, long long k0, long long m, const double* lower,
                              const double* range, double* out, long long ld, int first_fp32, cudaStream_t stream);

}  // namespace hdbo
