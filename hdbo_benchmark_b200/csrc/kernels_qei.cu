// qExpectedImprovement value and gradient w.r.t. the joint posterior (mean [q], Sigma [q x q])  (SURVEY.md §8a row A7).
// One CTA per restart: q x q Cholesky of Sigma + jitter I in shared memory, S base samples spread over the threads,
//   value = mean_s max_j relu(mu_j + (L z_s)_j - f*),
// the backward accumulates d/dmu and d/dL from the arg-max entry of each improving sample, then applies the Cholesky
// adjoint  Sigma_bar = sym( L^-T Phi(L^T L_bar) L^-1 )  (Phi = lower triangle, halved diagonal).
#include "kernels.cuh"
#include "grad.cuh"

namespace hdbo {

template <int Q>
__global__ void __launch_bounds__(128) k_qei(const double* __restrict__ mean, const double* __restrict__ Sigma,
                                             const double* __restrict__ base, long long S, double best_f, double jitter,
                                             double* __restrict__ value, double* __restrict__ gmean,
                                             double* __restrict__ gSigma, int32_t* __restrict__ info) {
    constexpr int NT = Q * (Q + 1) / 2;
    __shared__ double sL[Q][Q];
    __shared__ double smu[Q];
    __shared__ double red[4][1 + Q + NT];
    const long long r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < Q * Q) sL[tid / Q][tid % Q] = Sigma[r * Q * Q + tid];
    if (tid < Q) smu[tid] = mean[r * Q + tid];
    __syncthreads();
    if (tid == 0) {
        int bad = 0;
        for (int j = 0; j < Q; j++) {
            double djj = sL[j][j] + jitter;
            for (int k = 0; k < j; k++) djj -= sL[j][k] * sL[j][k];
            if (!(djj > 0.0)) { bad = j + 1; djj = 1.0; }
            const double l = sqrt(djj);
            sL[j][j] = l;
            for (int i = j + 1; i < Q; i++) {
                double s = sL[i][j];
                for (int k = 0; k < j; k++) s -= sL[i][k] * sL[j][k];
                sL[i][j] = s / l;
            }
            for (int k = j + 1; k < Q; k++) sL[j][k] = 0.0;
        }
        if (info) info[r] = bad;
    }
    __syncthreads();
    double L[Q][Q], mu[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) {
        mu[j] = smu[j] - best_f;
#pragma unroll
        for (int k = 0; k < Q; k++) L[j][k] = sL[j][k];
    }
    double acc = 0.0, gm[Q], gL[NT];
#pragma unroll
    for (int j = 0; j < Q; j++) gm[j] = 0.0;
#pragma unroll
    for (int t = 0; t < NT; t++) gL[t] = 0.0;
    for (long long s = tid; s < S; s += blockDim.x) {
        double z[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) z[k] = base[s * Q + k];
        double best = -INFINITY; int js = 0;
#pragma unroll
        for (int j = 0; j < Q; j++) {
            double v = mu[j];
#pragma unroll
            for (int k = 0; k <= j; k++) v = fma(L[j][k], z[k], v);
            if (v > best) { best = v; js = j; }
        }
        if (best > 0.0) {
            acc += best;
#pragma unroll
            for (int j = 0; j < Q; j++) {
                const double ind = (j == js) ? 1.0 : 0.0;
                gm[j] += ind;
#pragma unroll
                for (int k = 0; k <= j; k++) gL[j * (j + 1) / 2 + k] = fma(ind, z[k], gL[j * (j + 1) / 2 + k]);
            }
        }
    }
    auto wsum = [](double x) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    };
    acc = wsum(acc);
    if (lane == 0) red[warp][0] = acc;
#pragma unroll
    for (int j = 0; j < Q; j++) { double x = wsum(gm[j]); if (lane == 0) red[warp][1 + j] = x; }
#pragma unroll
    for (int t = 0; t < NT; t++) { double x = wsum(gL[t]); if (lane == 0) red[warp][1 + Q + t] = x; }
    __syncthreads();
    if (tid != 0) return;
    const double invS = 1.0 / (double)S;
    auto tot = [&](int i) { return ((red[0][i] + red[1][i]) + red[2][i]) + red[3][i]; };
    value[r] = tot(0) * invS;
    if (!gmean && !gSigma) return;
    if (gmean) for (int j = 0; j < Q; j++) gmean[r * Q + j] = tot(1 + j) * invS;
    if (!gSigma) return;
    double Lb[Q][Q], P[Q][Q], X[Q][Q], Gs[Q][Q];
    for (int j = 0; j < Q; j++)
        for (int k = 0; k < Q; k++) Lb[j][k] = (k <= j) ? tot(1 + Q + j * (j + 1) / 2 + k) * invS : 0.0;
    // P = Phi(L^T Lb)
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) {
            double s = 0.0;
            if (j <= i) for (int k = 0; k < Q; k++) s += L[k][i] * Lb[k][j];
            P[i][j] = (j < i) ? s : ((j == i) ? 0.5 * s : 0.0);
        }
    // X = L^-T P : solve L^T X = P (back substitution)
    for (int c = 0; c < Q; c++)
        for (int i = Q - 1; i >= 0; i--) {
            double s = P[i][c];
            for (int k = i + 1; k < Q; k++) s -= L[k][i] * X[k][c];
            X[i][c] = s / L[i][i];
        }
    // Gs = X L^-1 : solve Gs L = X  (columns right to left)
    for (int i = 0; i < Q; i++)
        for (int c = Q - 1; c >= 0; c--) {
            double s = X[i][c];
            for (int k = c + 1; k < Q; k++) s -= Gs[i][k] * L[k][c];
            Gs[i][c] = s / L[c][c];
        }
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) gSigma[r * Q * Q + i * Q + j] = 0.5 * (Gs[i][j] + Gs[j][i]);
}

cudaError_t launch_qei(const double* mean, const double* Sigma, const double* base, long long R, int q, long long S,
                       double best_f, double jitter, double* value, double* gmean, double* gSigma, int32_t* info,
                       cudaStream_t st) {
    if (R == 0) return cudaSuccess;
#define QCASE(Q) case Q: k_qei<Q><<<(unsigned)R, 128, 0, st>>>(mean, Sigma, base, S, best_f, jitter, value, gmean, gSigma, info); break;
    switch (q) {
        QCASE(1) QCASE(2) QCASE(3) QCASE(4) QCASE(5) QCASE(6) QCASE(7) QCASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef QCASE
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Joint posterior samples of q <= 8 points: out[s][r][:] = mean[r] + chol(Sigma[r] + jitter I) z[s][r or 0][:]
// (GPyTorchPosterior.rsample_from_base_samples on the q x q factor).  One thread per (s, r): the q x q factorisation is ~q^3/3
// flops, cheaper to redo in registers than to stage.
template <int Q>
__global__ void k_mvn_sample(const double* __restrict__ mean, const double* __restrict__ Sigma, const double* __restrict__ z,
                             long long R, long long S, int z_per_r, double jitter, double* __restrict__ out,
                             int32_t* __restrict__ info) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * S) return;
    const long long r = t % R, s = t / R;
    double L[Q][Q];
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j < Q; j++) L[i][j] = Sigma[(r * Q + i) * Q + j];
    int bad = 0;
#pragma unroll
    for (int j = 0; j < Q; j++) {
        double djj = L[j][j] + jitter;
#pragma unroll
        for (int k = 0; k < j; k++) djj -= L[j][k] * L[j][k];
        if (!(djj > 0.0)) { if (!bad) bad = j + 1; djj = 1.0; }
        const double l = sqrt(djj), il = 1.0 / l;
        L[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < Q; i++) {
            double v = L[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
            L[i][j] = v * il;
        }
    }
    if (info && s == 0) info[r] = bad;
    const double* zz = z + (z_per_r ? (s * R + r) : s) * Q;
    double zv[Q];
#pragma unroll
    for (int k = 0; k < Q; k++) zv[k] = zz[k];
#pragma unroll
    for (int j = 0; j < Q; j++) {
        double v = mean[r * Q + j];
#pragma unroll
        for (int k = 0; k <= j; k++) v = fma(L[j][k], zv[k], v);
        out[(s * R + r) * Q + j] = v;
    }
}

cudaError_t launch_mvn_sample(const double* mean, const double* Sigma, const double* z, long long R, int q, long long S,
                              int z_per_r, double jitter, double* out, int32_t* info, cudaStream_t st) {
    if (R * S == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((R * S + 127) / 128);
#define QCASE(Q) case Q: k_mvn_sample<Q><<<grid, 128, 0, st>>>(mean, Sigma, z, R, S, z_per_r, jitter, out, info); break;
    switch (q) {
        QCASE(1) QCASE(2) QCASE(3) QCASE(4) QCASE(5) QCASE(6) QCASE(7) QCASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef QCASE
    return cudaGetLastError();
}

}  // namespace hdbo
