"""Quasi-random helpers with BoTorch's conventions (botorch.sampling.normal.SobolQMCNormalSampler,
botorch.utils.sampling.draw_sobol_samples / draw_sobol_normal_samples)."""
from __future__ import annotations

import math
from typing import Optional

import torch


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None) -> torch.Tensor:
    """n x q x d scrambled-Sobol points inside `bounds` (2 x d); one (q*d)-dimensional engine as in BoTorch."""
    bounds = torch.as_tensor(bounds)
    d = bounds.shape[-1]
    lower, rng = bounds[0], bounds[1] - bounds[0]
    eng = torch.quasirandom.SobolEngine(q * d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=torch.float64).reshape(n, q, d).to(bounds.device)
    return lower.to(torch.float64) + rng.to(torch.float64) * u


def draw_sobol_normal_samples(d: int, n: int, seed: Optional[int] = None) -> torch.Tensor:
    """n x d standard-normal QMC draws: Phi^-1 of scrambled Sobol points shrunk away from {0, 1}."""
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=torch.float64)
    v = 0.5 + (1.0 - torch.finfo(torch.float64).eps) * (u - 0.5)
    return torch.erfinv(2.0 * v - 1.0) * math.sqrt(2.0)


class SobolQMCNormalSampler:
    """Fixed base samples z[S, q] reused across calls (the MC acquisition is then a deterministic function of X)."""

    def __init__(self, sample_shape=torch.Size([512]), seed: Optional[int] = None, **kwargs):
        if isinstance(sample_shape, int):
            sample_shape = torch.Size([sample_shape])
        self.sample_shape = torch.Size(sample_shape)
        self.seed = seed if seed is not None else int(torch.randint(0, 1000000, (1,)).item())
        self._cache = {}

    def base_samples(self, q: int, device) -> torch.Tensor:
        key = (q, str(device))
        if key not in self._cache:
            S = int(self.sample_shape.numel())
            self._cache[key] = draw_sobol_normal_samples(q, S, seed=self.seed).to(device).contiguous()
        return self._cache[key]
