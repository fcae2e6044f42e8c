// Declarations of the gradient-path kernels (MLL gradient, posterior backward, qEI); see kernels_grad.cu.
#pragma once
#include "kernels.cuh"

namespace hdbo {
}  // namespace hdbo
