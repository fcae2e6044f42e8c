import sys
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import acquisition as A
from hdbo_benchmark_b200 import gp_modules as G
from hdbo_benchmark_b200.models import Normalize, SingleTaskGP, Standardize
from hdbo_benchmark_b200.optim import gen_candidates_device, gen_candidates_scipy
from hdbo_benchmark_b200.sampling import SobolQMCNormalSampler
dev = torch.device("cuda")
for q in (2, 4):
    n, d, R = 150, 8, 10
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "short", 1e-3, seed=2)
    bounds = torch.stack([-1.0 * torch.ones(d), 3.0 * torch.ones(d)]).double()
    model = SingleTaskGP(X * 4.0 - 1.0, 2.0 * y.unsqueeze(-1) - 3.0, covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=Standardize(m=1), input_transform=Normalize(d, bounds=bounds))
    model.covar_module.base_kernel.lengthscale = ls / 4.0 * 1.5
    model.covar_module.outputscale = 1.0
    model.likelihood.noise = 1e-3
    model.eval()
    acq = A.qExpectedImprovement(model, best_f=float((2.0 * y - 3.0).max()) - 0.3,
                                 sampler=SobolQMCNormalSampler(sample_shape=torch.Size([128]), seed=1))
    ics = (O.sobol_candidates(R * q, d, seed=8) * 4.0 - 1.0).reshape(R, q, d).to(dev)
    c_s, v_s = gen_candidates_scipy(ics, acq, bounds[0], bounds[1], options={"maxiter": 200})
    for T in (8, 12):
        c_d, v_d = gen_candidates_device(ics, acq, bounds[0], bounds[1], options={"maxiter": 200, "line_search_trials": T})
        info = gen_candidates_device.last_info
        print("q", q, "T", T, "v_d", [round(float(v), 4) for v in v_d], "iters", info.get("per_restart_iterations"), "flags", info.get("flags"))
    print("q", q, "v_s", [round(float(v), 4) for v in v_s])
    # projected gradient norms at the device / scipy answers
    for name, c in (("dev", c_d), ("scipy", c_s)):
        Xg = c.clone().to(dev).requires_grad_(True)
        v = acq(Xg)
        (g,) = torch.autograd.grad(v.sum(), Xg)
        lo, hi = -1.0, 3.0
        pg = g.clone()
        pg[(Xg <= lo + 1e-12) & (g < 0)] = 0.0     # maximisation: gradient pointing outside the box
        pg[(Xg >= hi - 1e-12) & (g > 0)] = 0.0
        print(" ", name, "max |proj grad| per restart", [float(f"{float(x):.2e}") for x in pg.reshape(R, -1).abs().max(dim=1).values])
