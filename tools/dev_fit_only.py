import sys
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda")
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, keep_r2=True); gp.set_hyper(ls.to(dev), os_, nz, mu)
for _ in range(3):
    gp.fit_once(need_r2=True); mll_gradient(gp)
torch.cuda.synchronize()
print("ok")
