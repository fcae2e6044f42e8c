"""Where does a scipy-driven fit at n=4096 spend its time?  Set-up (DeviceGP allocation, eager warm-up, graph capture) vs per-evaluation
cost of the fast closure, graph replay vs eager body.   python tools/prof_fast_closure.py [n] [d]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from hdbo_benchmark_b200.fit import _FastClosure, _named_raw_parameters
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator().manual_seed(0)
X = torch.rand(n, d, generator=g, dtype=torch.float64)
y = (torch.sin(3 * X[:, :4].sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
out = {"n": n, "d": d}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model = SingleTaskGP(X, y)
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    model.train()
    params = _named_raw_parameters(mll)
    fc = _FastClosure.build(mll, params)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    fc._capture()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    x = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in params])
    fc(x); torch.cuda.synchronize()
    t3 = time.perf_counter()
    for _ in range(20):
        fc(x)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    graph = fc.graph
    fc.graph = None
    fc(x); torch.cuda.synchronize(); t5 = time.perf_counter()
    for _ in range(20):
        fc(x)
    torch.cuda.synchronize(); t6 = time.perf_counter()
    out[f"rep{rep}"] = {"model_build_ms": (t1 - t0) * 1e3, "capture_ms": (t2 - t1) * 1e3, "eval_graph_ms": (t4 - t3) / 20 * 1e3,
                        "eval_eager_ms": (t6 - t5) / 20 * 1e3}
    del fc, mll, model, graph
    import gc; gc.collect(); torch.cuda.empty_cache()
print(json.dumps(out))
