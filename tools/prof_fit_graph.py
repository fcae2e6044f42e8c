"""Eager vs CUDA-graph replay of one MLL value + gradient (hdbo_gp_fit + hdbo_mll_grad) at n=4096, d=256, and the host time the eager
enqueue takes (the pipelined factorisation issues ~40 stream operations per pair of columns: is the host the limiter?).
    python tools/prof_fit_graph.py"""
import json, sys, time
sys.path.insert(0, ".")
import torch
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient
import bench

dev = torch.device("cuda")
n, d = 4096, 256
X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
gp.set_hyper(ls.to(dev), os_, nz, mu)


def evaluate():
    gp.fit_once(need_r2=True)
    return mll_gradient(gp)


def gpu_ms(fn, reps=7):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


for _ in range(3):
    ref = evaluate().clone()
torch.cuda.synchronize()
out = {}
out["eager_ms_min"], out["eager_ms_median"] = gpu_ms(evaluate)
host = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); evaluate(); host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
out["eager_host_enqueue_ms"] = min(host)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    res = evaluate()
graph.replay(); torch.cuda.synchronize()
out["graph_equals_eager"] = bool(torch.equal(res, ref))
out["graph_ms_min"], out["graph_ms_median"] = gpu_ms(graph.replay)
# fit alone (no gradient), both ways
def fit_only():
    gp.fit_once(need_r2=False)
fit_only(); torch.cuda.synchronize()
out["fit_only_eager_ms"], _ = gpu_ms(fit_only)
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    fit_only()
g2.replay(); torch.cuda.synchronize()
out["fit_only_graph_ms"], _ = gpu_ms(g2.replay)
print(json.dumps(out))
