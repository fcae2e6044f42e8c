"""Timing only: int8-mode pool evaluation at the headline shape; prints cross / var milliseconds per 2^18 candidates."""
import sys, json, ctypes as C
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
dev = torch.device("cuda")
n, d, m = 4096, 256, 1 << 18
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
Z = torch.rand(m, d, dtype=torch.float64, device=dev)
lib = _lib.load()
gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, var_mode="i8")
gp.set_hyper(ls.to(dev), os_, nz, mu); gp.fit()
out = {"acq": torch.empty(1, m, dtype=torch.float64, device=dev)}
for _ in range(2):
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False, out=out)
lib.hdbo_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False, out=out)
e1.record(); torch.cuda.synchronize()
ms = (C.c_double * 2)(); cnt = (C.c_int64 * 2)()
lib.hdbo_profile_collect(ms, cnt); lib.hdbo_profile_enable(0)
print(json.dumps({"pool_ms": e0.elapsed_time(e1) / 3, "cross_ms": ms[0] / 3, "var_ms": ms[1] / 3, "flag": gp.i8_flag()}))
