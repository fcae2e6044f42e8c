"""MLL fit (+gradient) device time at a few sizes (CUDA events, best of 5)."""
import sys, json
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient
dev = torch.device("cuda")
for n, d in ((310, 128), (1000, 128), (4096, 256)):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, keep_r2=True); gp.set_hyper(ls.to(dev), os_, nz, mu)
    best_f, best_g = 1e9, 1e9
    for _ in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); gp.fit_once(need_r2=True); e[1].record(); mll_gradient(gp); e[2].record(); torch.cuda.synchronize()
        best_f = min(best_f, e[0].elapsed_time(e[1])); best_g = min(best_g, e[0].elapsed_time(e[2]))
    print(json.dumps({"n": n, "d": d, "fit_ms": best_f, "fit_plus_grad_ms": best_g}))
