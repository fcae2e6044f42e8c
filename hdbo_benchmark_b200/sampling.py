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


class DeviceSobol:
    """A torch.quasirandom.SobolEngine whose points are generated in HBM by `hdbo_sobol_draw` (bit-identical to the host engine).

    The (scrambled) direction numbers and the digital shift come from torch's own engine -- so seeds mean what they mean upstream --
    and are uploaded once (30 x dim int32); a draw of m points is then one HBM-bound kernel (8 B per element) instead of a
    single-threaded host loop plus a host-to-device copy: the 2^20 x 256 raw-sample pool of `optimize_acqf` takes seconds on the
    host.  Same stateful interface: `draw(n)` continues the sequence, `fast_forward(n)` skips, `reset()` rewinds."""

    def __init__(self, dimension: int, scramble: bool = True, seed: Optional[int] = None, device=None):
        eng = torch.quasirandom.SobolEngine(dimension, scramble=scramble, seed=seed)
        self.dimension = dimension
        self.device = torch.device(device if device is not None else "cuda")
        self.state_t = eng.sobolstate.t().contiguous().to(torch.int32).to(self.device)
        self.shift = eng.shift.to(torch.int32).to(self.device)
        # torch computes point 0 in the default dtype at construction time (FP32-rounded unless the default is float64)
        self.first_point_fp32 = int(eng._first_point.dtype == torch.float32)
        self.num_generated = 0

    def reset(self):
        self.num_generated = 0
        return self

    def fast_forward(self, n: int):
        self.num_generated += int(n)
        return self

    def draw(self, n: int = 1, lower: Optional[torch.Tensor] = None, rng: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """n x dimension float64 points on the device; `lower` / `rng` [dimension] map the unit cube to lower + rng * u."""
        from . import _lib
        from ._lib import check, ptr
        if out is None:
            out = torch.empty(n, self.dimension, dtype=torch.float64, device=self.device)
        assert out.is_cuda and out.dtype == torch.float64 and out.shape == (n, self.dimension) and out.stride(1) == 1
        with torch.cuda.device(self.device):
            check(_lib.load().hdbo_sobol_draw(ptr(self.state_t), ptr(self.shift), self.dimension, self.num_generated, n,
                                              ptr(lower), ptr(rng), ptr(out), out.stride(0), self.first_point_fp32,
                                              torch.cuda.current_stream().cuda_stream), "hdbo_sobol_draw")
        self.num_generated += n
        return out


_DEVICE_ENGINES = {}


def _device_sobol_points(dim: int, n: int, seed: Optional[int], stream: Optional[int], device, lower, rng) -> torch.Tensor:
    """Device twin of `_sobol_unit_points`: same seed / stream semantics, points generated in HBM (nothing to cache but the engine
    state: a seeded draw restarts the sequence at point 0)."""
    if stream is not None:
        eng = _cached(_DEVICE_ENGINES, ("stream", dim, stream, str(device)), lambda: DeviceSobol(dim, True, stream, device))
    elif seed is not None:
        eng = _cached(_DEVICE_ENGINES, ("seed", dim, seed, str(device)), lambda: DeviceSobol(dim, True, seed, device)).reset()
    else:
        eng = _cached(_DEVICE_ENGINES, (dim, None, str(device)), lambda: DeviceSobol(dim, True, None, device))
    return eng.draw(n, lower, rng)


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None, stream: Optional[int] = None,
                       device=None) -> torch.Tensor:
    """n x q x d scrambled-Sobol points inside `bounds` (2 x d); one (q*d)-dimensional engine as in BoTorch.  With a CUDA `device`
    the points are generated there (`DeviceSobol`), bit-identical to the host path."""
    bounds = torch.as_tensor(bounds)
    d = bounds.shape[-1]
    lower, rng = bounds[0], bounds[1] - bounds[0]
    if device is not None and torch.device(device).type == "cuda":
        lo = lower.to(device, torch.float64).repeat(q).contiguous()
        rg = rng.to(device, torch.float64).repeat(q).contiguous()
        return _device_sobol_points(q * d, n, seed, stream, torch.device(device), lo, rg).reshape(n, q, d)
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
