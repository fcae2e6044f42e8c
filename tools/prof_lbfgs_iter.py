"""A few iterations of the device L-BFGS at n=1024, d=128 (eager, no CUDA graph) for a per-kernel launch list."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import acquisition as A, gp_modules as G
from hdbo_benchmark_b200.models import SingleTaskGP
from hdbo_benchmark_b200.optim import optimize_acqf
torch.manual_seed(0)
n, d = 1024, 128
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-3, seed=1)
model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)), likelihood=G.GaussianLikelihood(), outcome_transform=None)
model.covar_module.base_kernel.lengthscale = ls; model.covar_module.outputscale = float(os_); model.likelihood.noise = nz; model.eval()
acq = A.LogExpectedImprovement(model, best_f=float(y.max()))
bounds = torch.stack([torch.zeros(d), torch.ones(d)]).double()
optimize_acqf(acq, bounds, q=1, num_restarts=20, raw_samples=512, options={"seed": 0, "method": "device-lbfgs", "cuda_graph": False, "maxiter": 30})
torch.cuda.synchronize(); print("ok")
