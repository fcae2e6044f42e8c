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
gp.attach_timeline(4096)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False, out=out)
e1.record(); torch.cuda.synchronize()
tl = gp.collect_timeline()
print(json.dumps({"pool_ms": e0.elapsed_time(e1) / 3, "phases_ms": {k: v[0] / 3 for k, v in tl.items()}, "flag": gp.i8_flag()}))
