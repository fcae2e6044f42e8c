"""optimize_acqf wall clock: scipy L-BFGS-B host loop vs the batched device L-BFGS, at the reference's scale (n = 310, d = 128,
LogEI, 20 restarts, 512 raw samples) and at n = 1024.  Developer measurement for DESIGN.md."""
import json, sys, time
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import acquisition as A, gp_modules as G
from hdbo_benchmark_b200.models import SingleTaskGP
from hdbo_benchmark_b200.optim import optimize_acqf

torch.manual_seed(0)
for n, d in ((310, 128), (1024, 128)):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-3, seed=1)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = float(os_)
    model.likelihood.noise = nz
    model.eval()
    acq = A.LogExpectedImprovement(model, best_f=float(y.max()))
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).double()
    out = {"n": n, "d": d}
    from hdbo_benchmark_b200 import optim as _optim
    for name, opts in (("scipy", {"seed": 0}), ("device", {"seed": 0, "method": "device-lbfgs"}),
                       ("device_T8", {"seed": 0, "method": "device-lbfgs", "line_search_trials": 8}),
                       ("device_nograph", {"seed": 0, "method": "device-lbfgs", "cuda_graph": False})):
        optimize_acqf(acq, bounds, q=1, num_restarts=20, raw_samples=512, options=opts)      # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            cand, val = optimize_acqf(acq, bounds, q=1, num_restarts=20, raw_samples=512, options=opts)
        torch.cuda.synchronize()
        out[name + "_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        out[name + "_value"] = float(val)
        if name != "scipy":
            info = _optim.gen_candidates_device.last_info
            out[name + "_iters"] = info["iterations"]; out[name + "_flags"] = sorted(set(info["flags"]))
            out[name + "_median_restart_iters"] = sorted(info["per_restart_iterations"])[len(info["per_restart_iterations"]) // 2]
    print(json.dumps(out), flush=True)
