// FP32 dense layers for the VAE decode step on either side of the GP hot path.
//
// Replaces: the eval-mode nn.Sequential(Linear, BatchNorm1d, ReLU, Dropout, ...) decoder of the reference's latent-space models
// (/root/reference/src/hdbo_benchmark/generative_models/vae_selfies.py:67-81: latent -> 256 -> 1024 -> 2048 -> L*A logits) and the
// arg-max over the alphabet that turns logits into token strings (vae_selfies.py:235-247).  A BO step decodes 1..10 latent
// points (the initial design up to 1000), so a layer is a skinny GEMM: every weight is read once per pass and used for <= 16
// batch rows -- bound by streaming the 46 MB of FP32 weights from HBM / L2, not by arithmetic.  FP32 FFMA on purpose: torch's
// default float32 matmul does not use TF32, and the arg-max must not move.
//
//   k_linear_f32<BT, NPW, UNR>   Y[r][n] = act((X[r] . W[n] + bias[n]) * scale[n] + shift[n])      (BatchNorm in eval mode = scale / shift)
//       one warp per NPW output neurons: lanes stride over k with 16-byte loads of the weight rows (full 512-byte lines per warp),
//       the BT activation rows come through L1 (every warp of the grid reads the same few KB), warp-shuffle reduction.
//   k_argmax_tokens_f32     tokens[r][p] = argmax_a logits[r][p * A + a], first index on ties (numpy.argmax)
#include "kernels.cuh"
#include "mlp.cuh"

namespace hdbo {

namespace {
constexpr int ML_WARPS = 8;

template <int BT, int NPW, int UNR, bool VEC>
__global__ void __launch_bounds__(ML_WARPS * 32) k_linear_f32(const float* __restrict__ X, long long ldx, int b, int K,
                                                               const float* __restrict__ W, const float* __restrict__ bias,
                                                               const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                               float* __restrict__ Y, long long ldy, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * ML_WARPS + warp) * NPW, r0 = blockIdx.y * BT;
    if (n0 >= N) return;
    float acc[NPW][BT];
#pragma unroll
    for (int j = 0; j < NPW; j++)
#pragma unroll
        for (int r = 0; r < BT; r++) acc[j][r] = 0.f;
    const float* xr[BT];
#pragma unroll
    for (int r = 0; r < BT; r++) xr[r] = X + (size_t)min(r0 + r, b - 1) * ldx;     // rows past the batch repeat the last one (not stored)
    const float* wr[NPW];
#pragma unroll
    for (int j = 0; j < NPW; j++) wr[j] = W + (size_t)min(n0 + j, N - 1) * K;
    if (VEC) {
        // UNR x NPW 16-byte weight loads are issued before the first FMA of an iteration: the layer is a stream of weights
        // (every byte used for <= BT rows), so bytes in flight per SM decide the speed, not instruction count
        const int K4 = K >> 2;
        for (int k0 = 0; k0 < K4; k0 += 32 * UNR) {
            float4 w[UNR][NPW];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = k0 + u * 32 + lane;
#pragma unroll
                for (int j = 0; j < NPW; j++) w[u][j] = (k < K4) ? __ldg(reinterpret_cast<const float4*>(wr[j]) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = min(k0 + u * 32 + lane, K4 - 1);                     // (weights are zero past the end)
#pragma unroll
                for (int r = 0; r < BT; r++) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(xr[r]) + k);
#pragma unroll
                    for (int j = 0; j < NPW; j++) {
                        acc[j][r] = fmaf(w[u][j].x, x.x, acc[j][r]); acc[j][r] = fmaf(w[u][j].y, x.y, acc[j][r]);
                        acc[j][r] = fmaf(w[u][j].z, x.z, acc[j][r]); acc[j][r] = fmaf(w[u][j].w, x.w, acc[j][r]);
                    }
                }
            }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
#pragma unroll
            for (int r = 0; r < BT; r++) {
                const float x = __ldg(xr[r] + k);
#pragma unroll
                for (int j = 0; j < NPW; j++) acc[j][r] = fmaf(__ldg(wr[j] + k), x, acc[j][r]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NPW; j++)
#pragma unroll
        for (int r = 0; r < BT; r++) {
            float v = acc[j][r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[j][r] = v;
        }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NPW; j++) {
            const int n = n0 + j;
            if (n >= N) break;
            const float bs = bias ? bias[n] : 0.f, sc = scale ? scale[n] : 1.f, sh = shift ? shift[n] : 0.f;
#pragma unroll
            for (int r = 0; r < BT; r++) {
                if (r0 + r >= b) break;
                float v = acc[j][r] + bs;
                if (scale || shift) v = fmaf(v, sc, sh);
                if (relu) v = fmaxf(v, 0.f);
                Y[(size_t)(r0 + r) * ldy + n] = v;
            }
        }
    }
}

// Batches of >= 128 rows (the initial design: up to 1000 latent points): a classic shared-memory tiled SGEMM, 64 x 64 outputs per
// CTA, 16-deep k-steps, 4 x 4 outputs per thread, global -> register prefetch of the next k-step while the current one is
// multiplied.  (The skinny kernel above would re-stream the weights once per 16 rows and sits at ~18 % of the FFMA rate there.)
constexpr int TL_BM = 64, TL_BN = 64, TL_BK = 16, TL_PAD = 4;
__global__ void __launch_bounds__(256) k_linear_tiled_f32(const float* __restrict__ X, long long ldx, int b, int K,
                                                          const float* __restrict__ W, const float* __restrict__ bias,
                                                          const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                          float* __restrict__ Y, long long ldy, int N) {
    __shared__ __align__(16) float Xs[2][TL_BK][TL_BM + TL_PAD];
    __shared__ __align__(16) float Ws[2][TL_BK][TL_BN + TL_PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.y * TL_BM, n0 = blockIdx.x * TL_BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;                  // this thread's 16-byte piece of the two operand tiles
    const bool xok = r0 + lrow < b, wok = n0 + lrow < N;
    const float* xp = X + (size_t)min(r0 + lrow, b - 1) * ldx + lk;
    const float* wp = W + (size_t)min(n0 + lrow, N - 1) * K + lk;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld = [&](const float* p, bool ok, int k0) { return (ok && k0 + lk < K) ? __ldg(reinterpret_cast<const float4*>(p + k0)) : z4; };   // K % 4 == 0
    float4 xv = ld(xp, xok, 0), wv = ld(wp, wok, 0);
    const int ksteps = (K + TL_BK - 1) / TL_BK;
    for (int s = 0; s < ksteps; s++) {
        const int buf = s & 1;
        Xs[buf][lk + 0][lrow] = xv.x; Xs[buf][lk + 1][lrow] = xv.y; Xs[buf][lk + 2][lrow] = xv.z; Xs[buf][lk + 3][lrow] = xv.w;
        Ws[buf][lk + 0][lrow] = wv.x; Ws[buf][lk + 1][lrow] = wv.y; Ws[buf][lk + 2][lrow] = wv.z; Ws[buf][lk + 3][lrow] = wv.w;
        __syncthreads();                                             // (two buffers: the next iteration's stores go to the other one)
        if (s + 1 < ksteps) { xv = ld(xp, xok, (s + 1) * TL_BK); wv = ld(wp, wok, (s + 1) * TL_BK); }
#pragma unroll
        for (int kk = 0; kk < TL_BK; kk++) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int n = n0 + tx * 4 + j;
        if (n >= N) break;
        const float bs = bias ? bias[n] : 0.f, sc = scale ? scale[n] : 1.f, sh = shift ? shift[n] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = r0 + ty * 4 + i;
            if (r >= b) break;
            float v = acc[i][j] + bs;
            if (scale || shift) v = fmaf(v, sc, sh);
            if (relu) v = fmaxf(v, 0.f);
            Y[(size_t)r * ldy + n] = v;
        }
    }
}

__global__ void k_argmax_tokens_f32(const float* __restrict__ logits, long long ld, int b, int L, int A, int* __restrict__ tokens) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)b * L) return;
    const int r = (int)(gw / L), p = (int)(gw % L);
    const float* src = logits + (size_t)r * ld + (size_t)p * A;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int a = lane; a < A; a += 32) { const float v = src[a]; if (v > best || bi == 0x7fffffff) { best = v; bi = a; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) tokens[(size_t)r * L + p] = bi;
}
}  // namespace

cudaError_t launch_linear_f32(const float* X, long long ldx, int b, int K, const float* W, const float* bias, const float* scale,
                              const float* shift, int relu, float* Y, long long ldy, int N, cudaStream_t st) {
    if (b == 0 || N == 0) return cudaSuccess;
    const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)W % 16 == 0);
    if (vec && b >= 128) {   // measured crossover against ceil(b / 16) passes of the skinny kernel: b = 100 0.205 vs 0.231 ms, b = 1000 1.63 vs 0.78 ms
        k_linear_tiled_f32<<<dim3((N + TL_BN - 1) / TL_BN, (b + TL_BM - 1) / TL_BM), 256, 0, st>>>(X, ldx, b, K, W, bias, scale, shift, relu, Y, ldy, N);
        return cudaGetLastError();
    }
#define HDBO_LAUNCH(BT, NPW, UNR)                                                                                                      \
    do {                                                                                                                               \
        const dim3 grid((N + ML_WARPS * NPW - 1) / (ML_WARPS * NPW), (b + BT - 1) / BT);                                               \
        if (vec) k_linear_f32<BT, NPW, UNR, true><<<grid, ML_WARPS * 32, 0, st>>>(X, ldx, b, K, W, bias, scale, shift, relu, Y, ldy, N); \
        else k_linear_f32<BT, NPW, 1, false><<<grid, ML_WARPS * 32, 0, st>>>(X, ldx, b, K, W, bias, scale, shift, relu, Y, ldy, N);    \
    } while (0)
    // few rows: one neuron per warp (N / 8 CTAs) unless the layer is wide enough to fill the machine twice over with two (4480 / 16 =
    // 280 CTAs) -- bytes in flight matter.  8-16 rows: every weight load is followed by BT activation loads, so 4 neurons per warp
    // amortise them (load : FMA = 20 : 256 instead of 17 : 64) even though the grid gets small.
    const bool wide = N >= 4096, mid = N >= 1024;
    if (b <= 1) { if (wide) HDBO_LAUNCH(1, 2, 4); else HDBO_LAUNCH(1, 1, 4); }
    else if (b <= 4) { if (wide) HDBO_LAUNCH(4, 2, 4); else HDBO_LAUNCH(4, 1, 4); }
    else if (b <= 8) { if (mid) HDBO_LAUNCH(8, 4, 2); else HDBO_LAUNCH(8, 1, 2); }
    else { if (mid) HDBO_LAUNCH(16, 4, 2); else HDBO_LAUNCH(16, 1, 2); }
#undef HDBO_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_argmax_tokens_f32(const float* logits, long long ld, int b, int L, int A, int* tokens, cudaStream_t st) {
    const long long warps = (long long)b * L;
    if (warps == 0) return cudaSuccess;
    k_argmax_tokens_f32<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(logits, ld, b, L, A, tokens);
    return cudaGetLastError();
}

}  // namespace hdbo
