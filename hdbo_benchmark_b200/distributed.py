"""Multi-GPU sharding of the hot path: one process per GPU, independent units, one tiny exchange.

Candidate pools (A4/A5), multi-start restarts (A6-A8) and SAAS hyper-parameter samples (A10) are independent given
(X, y, theta): rank r of P takes the contiguous slice [r*m/P, (r+1)*m/P) -- for Sobol pools the same engine
fast-forwarded by the slice start -- evaluates it with the single-GPU path, and the only data-path collective is an
all-gather of each rank's (max value, global arg-max index) = 16 bytes per rank (NCCL over NVLink on GPUs, gloo in
the CPU tests), or of the per-rank acquisition values when the caller needs the full vector
(`gen_batch_initial_conditions` draws its Boltzmann sample from all of them).  Training data and
hyper-parameters are broadcast once per BO iteration; every rank refits redundantly (a few ms) instead of
receiving the 134 MB inverse factor.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice of `total` units for `rank` (first total % world ranks get one extra)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_training_data(X: torch.Tensor, y: torch.Tensor, src: int = 0, group=None):
    """Broadcast (X, y) from `src` once per BO iteration (<= 8.4 MB at n=4096, d=256)."""
    if is_distributed():
        dist.broadcast(X, src=src, group=group)
        dist.broadcast(y, src=src, group=group)
    return X, y


def global_argmax(local_val: torch.Tensor, local_idx: torch.Tensor, offset: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(best value, global index) over all ranks.  local_idx indexes the rank's slice; `offset` is the slice start.
    Ties resolve to the smallest global index (first-occurrence semantics of torch.argmax on the full vector)."""
    val = local_val.reshape(1).to(torch.float64)
    idx = (local_idx.reshape(1) + offset).to(torch.float64)  # exact below 2**53
    if not is_distributed():
        return val[0], idx[0].to(torch.int64)
    pair = torch.cat([val, idx])
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(gathered, pair, group=group)
    allp = torch.stack(gathered)  # [world, 2]
    vals, idxs = allp[:, 0], allp[:, 1]
    best = vals.max()
    cand = torch.where(vals == best, idxs, torch.full_like(idxs, float("inf")))
    return best, cand.min().to(torch.int64)


def gather_values(local_vals: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All ranks' acquisition values as one [total] vector (slices may differ by one element)."""
    if not is_distributed():
        return local_vals
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    pad = torch.full((width,), float("-inf"), dtype=local_vals.dtype, device=local_vals.device)
    pad[: local_vals.numel()] = local_vals
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    return torch.cat([g[: e - s] for g, (s, e) in zip(gathered, sizes)])


def sharded_pool_argmax(evaluate_slice, total: int, group=None):
    """evaluate_slice(start, end) -> (values [end-start], local best value, local best index).  Returns the global
    (value, index) on every rank."""
    rank = dist.get_rank(group) if is_distributed() else 0
    world = dist.get_world_size(group) if is_distributed() else 1
    s, e = shard_bounds(total, rank, world)
    _, v, i = evaluate_slice(s, e)
    return global_argmax(v, i, s, group)
