"""world_size-2 `gloo` test of the multi-GPU host logic (pool sharding + the final (max, argmax) all-gather),
run on the CPU.  The per-slice evaluation is replaced by a deterministic function of the global index so that the
sharded result can be compared with the single-process arg-max of the full vector."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _values(total):
    g = torch.Generator().manual_seed(7)
    v = torch.randn(total, generator=g, dtype=torch.float64)
    v[total // 3] = v.max() + 1.0          # unique maximum in rank 0's slice (world 2)
    return v


def _worker(rank, world, port, total, tie, out_q):
    from hdbo_benchmark_b200.distributed import gather_values, global_argmax, shard_bounds, sharded_pool_argmax

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _values(total)
    if tie:
        full[total - 5] = full.max()       # duplicate maximum in the last rank's slice: smallest index must win

    def evaluate_slice(s, e):
        vals = full[s:e]
        i = torch.argmax(vals)
        return vals, vals[i], i

    best, idx = sharded_pool_argmax(evaluate_slice, total)
    s, e = shard_bounds(total, rank, world)
    allv = gather_values(full[s:e].clone(), total)
    ok = bool(torch.equal(allv, full)) and int(idx) == int(torch.argmax(full)) and float(best) == float(full.max())
    # broadcast path
    from hdbo_benchmark_b200.distributed import broadcast_training_data
    X = torch.full((4, 2), float(rank), dtype=torch.float64)
    y = torch.full((4,), float(rank), dtype=torch.float64)
    broadcast_training_data(X, y, src=0)
    ok = ok and float(X.sum()) == 0.0 and float(y.sum()) == 0.0
    out_q.put((rank, ok, int(idx)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,tie", [(1001, False), (4096, True)])
def test_sharded_argmax_matches_single_process(total, tie):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, tie, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    assert len({idx for _, _, idx in results}) == 1  # every rank agrees


def test_single_process_fallthrough():
    from hdbo_benchmark_b200.distributed import gather_values, global_argmax, is_distributed

    assert not is_distributed()
    v, i = global_argmax(torch.tensor(3.5, dtype=torch.float64), torch.tensor(4), offset=10)
    assert float(v) == 3.5 and int(i) == 14
    x = torch.arange(5, dtype=torch.float64)
    assert gather_values(x, 5) is x


def _worker_axes(rank, world, port, out_q):
    """SAAS sample shard (mixture moments / mixture LogEI by all-reduce) and restart shard r::P (best restart), world 2 on gloo."""
    from hdbo_benchmark_b200.distributed import (best_restart, mixture_log_acq, mixture_moments, shard_bounds,
                                                 strided_indices)

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)
    S, m = 7, 50                                     # odd S: ragged shards
    mean = torch.randn(S, m, generator=g, dtype=torch.float64)
    var = torch.rand(S, m, generator=g, dtype=torch.float64) + 0.1
    lacq = torch.randn(S, m, generator=g, dtype=torch.float64) * 30.0
    lacq[:, 3] = float("-inf")                       # a candidate with zero improvement under every sample
    s0, s1 = shard_bounds(S, rank, world)
    e, v = mixture_moments(mean[s0:s1], var[s0:s1], S)
    e_ref = mean.mean(0)
    v_ref = (var + mean * mean).mean(0) - e_ref * e_ref
    la = mixture_log_acq(lacq[s0:s1], S)
    la_ref = torch.logsumexp(lacq, dim=0) - torch.log(torch.tensor(float(S), dtype=torch.float64))
    ok = bool(torch.allclose(e, e_ref, rtol=1e-13, atol=1e-13)) and bool(torch.allclose(v, v_ref, rtol=1e-12, atol=1e-12))
    fin = torch.isfinite(la_ref)
    ok = ok and bool(torch.allclose(la[fin], la_ref[fin], rtol=1e-12, atol=1e-12)) and bool((la[~fin] == la_ref[~fin]).all())
    # restarts r::P
    R, q, d = 9, 2, 3
    vals = torch.randn(R, generator=g, dtype=torch.float64)
    vals[5] = vals[2] = vals.max() + 1.0             # tie: the smaller restart id (2, owned by rank 0) must win
    Xs = torch.randn(R, q, d, generator=g, dtype=torch.float64)
    ids = strided_indices(R, rank, world)
    best, win, Xw = best_restart(vals[ids], Xs[ids], ids)
    ok = ok and win == 2 and float(best) == float(vals.max()) and bool(torch.equal(Xw, Xs[2]))
    ok = ok and sorted(torch.cat([strided_indices(R, r, world) for r in range(world)]).tolist()) == list(range(R))
    out_q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_and_restart_shards_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_axes, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def test_mixture_helpers_single_process():
    from hdbo_benchmark_b200.distributed import best_restart, mixture_log_acq, mixture_moments

    mean = torch.tensor([[1.0, 2.0], [3.0, 2.0]], dtype=torch.float64)
    var = torch.tensor([[0.5, 1.0], [0.5, 1.0]], dtype=torch.float64)
    e, v = mixture_moments(mean, var, 2)
    assert torch.allclose(e, torch.tensor([2.0, 2.0], dtype=torch.float64)) and torch.allclose(v, torch.tensor([1.5, 1.0], dtype=torch.float64))
    la = mixture_log_acq(torch.log(torch.tensor([[0.2], [0.6]], dtype=torch.float64)), 2)
    assert abs(float(la) - float(torch.log(torch.tensor(0.4)))) < 1e-7
    val, win, X = best_restart(torch.tensor([0.1, 0.9, 0.3], dtype=torch.float64), torch.arange(6.0).reshape(3, 1, 2), torch.arange(3))
    assert win == 1 and float(val) == 0.9 and X.tolist() == [[2.0, 3.0]]
