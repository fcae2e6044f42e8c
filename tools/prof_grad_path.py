"""A6 at the headline shape for ncu captures: LogEI value + d/dX for 512 points at n=4096, d=256 (hdbo_posterior_fwd ->
hdbo_acq_elementwise -> hdbo_posterior_bwd), FP64 DMMA path."""
import sys
sys.path.insert(0, ".")
import torch
import bench
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.sampling import DeviceSobol

dev = torch.device("cuda")
n, d, pts = 4096, 256, 512
X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52)
gp.set_hyper(ls.to(dev), os_, nz, mu)
gp.fit()
Z = DeviceSobol(d, scramble=True, seed=21, device=dev).draw(pts)
bufs = {}
for _ in range(3):
    gp.acq_value_grad(Z, _lib.ACQ_LOGEI, float(y.max()), 1.0, bufs)
torch.cuda.synchronize()
print("ok")
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gp.acq_value_grad(Z, _lib.ACQ_LOGEI, float(y.max()), 1.0, bufs); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("value+grad ms (min of 7):", min(ts))
