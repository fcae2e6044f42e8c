"""One chunk of the int8-mode pool evaluation at the headline shape (for ncu captures)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP

dev = torch.device("cuda")
n, d, m = 4096, 256, 148 * 128
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
Z = torch.rand(m, d, dtype=torch.float64, device=dev)
gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, var_mode=sys.argv[1] if len(sys.argv) > 1 else "i8")
gp.set_hyper(ls.to(dev), os_, nz, mu)
gp.fit()
for _ in range(3):
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False)
torch.cuda.synchronize()
print("ok", gp.i8_flag())
