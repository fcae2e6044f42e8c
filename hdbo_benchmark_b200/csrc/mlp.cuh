// FP32 dense-layer launchers (kernels_mlp.cu): the VAE decode / encode step next to the GP hot path.
#pragma once
#include <cuda_runtime.h>

namespace hdbo {
cudaError_t launch_linear_f32(const float* X, long long ldx, int b, int K, const float* W, const float* bias, const float* scale,
                              const float* shift, int relu, float* Y, long long ldy, int N, cudaStream_t st);
cudaError_t launch_argmax_tokens_f32(const float* logits, long long ld, int b, int L, int A, int* tokens, cudaStream_t st);
}  // namespace hdbo
