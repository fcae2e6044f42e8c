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


# ---- the other two shard axes: SAAS hyper-parameter samples and multi-start restarts ------------------------------
def strided_indices(total: int, rank: int, world: int) -> torch.Tensor:
    """Restart shard r::P: restart i is optimised by rank i mod P; no data-path exchange until the final arg-max."""
    return torch.arange(rank, total, world)


def mixture_moments(mean: torch.Tensor, var: torch.Tensor, num_samples_total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gaussian-mixture mean / variance over ALL S hyper-parameter samples from this rank's slice of them.

    mean / var: [S_local, m] per-sample posterior moments of the rank's samples (botorch fully-Bayesian posteriors keep the MCMC
    dimension; the mixture is  E = mean_s mu_s,  V = mean_s (var_s + mu_s^2) - E^2).  The exchange is ONE all-reduce (sum) of the
    stacked [2, m] partial sums -- 16 bytes per candidate, independent of S -- instead of an all-gather of S x m moments."""
    part = torch.stack([mean.sum(0), (var + mean * mean).sum(0)])
    if is_distributed():
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    e = part[0] / num_samples_total
    return e, part[1] / num_samples_total - e * e


def mixture_log_acq(log_acq: torch.Tensor, num_samples_total: int, group=None) -> torch.Tensor:
    """log mean_s exp(log_acq_s) over all S samples (LogEI of a fully Bayesian model) from the rank's [S_local, m] slice:
    all-reduce(max) for the stabiliser, all-reduce(sum) of the shifted exponentials."""
    mx = log_acq.max(dim=0).values
    if is_distributed():
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    safe = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    s = torch.exp(log_acq - safe).sum(0)
    if is_distributed():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return torch.log(s / num_samples_total) + safe


def best_restart(local_vals: torch.Tensor, local_X: torch.Tensor, local_ids: torch.Tensor, group=None):
    """Best (value, restart id, X [q, d]) over all ranks' restarts; ties -> smallest restart id.  One all-gather of
    (value, id) pairs (16 B / rank) + one broadcast of the winning X from its owner."""
    i = torch.argmax(local_vals) if local_vals.numel() else None
    val = local_vals[i].reshape(1).to(torch.float64) if i is not None else torch.full((1,), float("-inf"), dtype=torch.float64,
                                                                                      device=local_X.device)
    rid = local_ids[i].reshape(1).to(torch.float64).to(val.device) if i is not None else torch.full_like(val, float("inf"))
    if not is_distributed():
        return val[0], int(rid[0]), local_X[i]
    world = dist.get_world_size(group)
    pair = torch.cat([val, rid])
    gathered = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(gathered, pair, group=group)
    allp = torch.stack(gathered)
    best = allp[:, 0].max()
    ids = torch.where(allp[:, 0] == best, allp[:, 1], torch.full_like(allp[:, 1], float("inf")))
    win = int(ids.min().item())
    owner = win % world                      # strided shard: restart id -> owner rank
    X = local_X[i].clone() if (i is not None and int(rid[0]) == win) else torch.empty_like(local_X[0])
    dist.broadcast(X, src=owner, group=group)
    return best, win, X
