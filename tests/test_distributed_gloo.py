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
