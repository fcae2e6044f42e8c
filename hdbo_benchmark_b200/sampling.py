"""Quasi-random helpers with BoTorch's conventions (botorch.sampling.normal.SobolQMCNormalSampler,
botorch.utils.sampling.draw_sobol_samples / draw_sobol_normal_samples)."""
from __future__ import annotations

import math
from typing import Optional

import torch


_ENGINES = {}
_SEEDED_POINTS = {}


def _cached(cache: dict, key, make, limit: int = 16):
    val = cache.get(key)
    if val is None:
        if len(cache) >= limit:   # bounded: callers that vary the key every iteration must not grow the cache
            cache.pop(next(iter(cache)))
        val = cache[key] = make()
    return val


def _sobol_unit_points(dim: int, n: int, seed: Optional[int], stream: Optional[int]) -> torch.Tensor:
    """n x dim scrambled-Sobol points in the unit cube (CPU float64).

    Constructing a scrambled SobolEngine costs 60-100 ms at dim >= 128 and `reset()` on a seeded one ~230 ms (measured
    here; it was half of the wall time of a BO step at d=128), so nothing is rebuilt per call:
      * seed given    -> the drawn points are cached by (dim, seed, n): identical points every call, as a fresh seeded
                         engine would give;
      * stream given  -> one engine per stream id, seeded with it and never reset: reproducible run to run, fresh points
                         on every call (solver loops);
      * neither       -> one unseeded engine per dimension that keeps drawing from its sequence."""
    if stream is not None:
        eng = _cached(_ENGINES, ("stream", dim, stream), lambda: torch.quasirandom.SobolEngine(dim, scramble=True, seed=stream))
        return eng.draw(n, dtype=torch.float64)
    if seed is not None:
        return _cached(_SEEDED_POINTS, (dim, seed, n),
                       lambda: torch.quasirandom.SobolEngine(dim, scramble=True, seed=seed).draw(n, dtype=torch.float64), limit=8)
    eng = _cached(_ENGINES, (dim, None), lambda: torch.quasirandom.SobolEngine(dim, scramble=True))
    return eng.draw(n, dtype=torch.float64)


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None, stream: Optional[int] = None) -> torch.Tensor:
    """n x q x d scrambled-Sobol points inside `bounds` (2 x d); one (q*d)-dimensional engine as in BoTorch."""
    bounds = torch.as_tensor(bounds)
    d = bounds.shape[-1]
    lower, rng = bounds[0], bounds[1] - bounds[0]
    u = _sobol_unit_points(q * d, n, seed, stream).reshape(n, q, d).to(bounds.device)
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
