"""One MLL value + gradient at n=4096, d=256 (pipelined factorisation) for ncu launch lists: python tools/prof_fit.py"""
import sys
sys.path.insert(0, ".")
import torch
import bench
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient

dev = torch.device("cuda")
n, d = 4096, 256
X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
gp.set_hyper(ls.to(dev), os_, nz, mu)
for _ in range(2):
    gp.fit_once(need_r2=True); mll_gradient(gp)
torch.cuda.synchronize()
print("ok")
