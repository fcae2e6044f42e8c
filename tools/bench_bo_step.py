"""Wall-clock anatomy of one BO solver step at the reference's own scale (BASELINE config #1 shape: d=128, n=310):
SingleTaskGP -> fit_gpytorch_mll (L-BFGS-B, one CUDA MLL value+gradient per iteration) -> LogEI -> optimize_acqf."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from hdbo_benchmark_b200 import acquisition as A
from hdbo_benchmark_b200.fit import fit_gpytorch_mll, fit_gpytorch_mll_scipy
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
from hdbo_benchmark_b200.optim import optimize_acqf, gen_batch_initial_conditions, gen_candidates_scipy

def sync_time(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, time.perf_counter() - t

for n, d in ((310, 128), (1000, 128)):
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = (torch.sin(3 * X[:, :4].sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
    for rep in range(2):
        model, t_ctor = sync_time(lambda: SingleTaskGP(X, y))
        mll = ExactMarginalLogLikelihood(model.likelihood, model)
        calls = {"n": 0}
        orig = mll.forward
        def counted(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)
        mll.forward = counted
        _, t_fit = sync_time(lambda: fit_gpytorch_mll(mll))
        model.eval()
        acqf = A.LogExpectedImprovement(model, best_f=float(y.max()))
        bounds = torch.stack([torch.zeros(d), torch.ones(d)]).double()
        ics, t_ics = sync_time(lambda: gen_batch_initial_conditions(acqf, bounds, 1, 20, 1024, {"seed": rep}))
        evals = {"n": 0}
        class Wrap:
            model = acqf.model
            def __call__(self, Xq):
                evals["n"] += 1
                return acqf(Xq)
        (cand, val), t_opt = sync_time(lambda: gen_candidates_scipy(ics, Wrap(), bounds[0], bounds[1]))
    print(json.dumps({"n": n, "d": d, "ctor_ms": t_ctor * 1e3, "fit_ms": t_fit * 1e3, "mll_evals": calls["n"],
                      "ms_per_mll_eval": t_fit * 1e3 / max(calls["n"], 1), "raw_samples_1024_ms": t_ics * 1e3,
                      "lbfgs_20_restarts_ms": t_opt * 1e3, "acqf_evals": evals["n"], "ms_per_acqf_eval": t_opt * 1e3 / max(evals["n"], 1),
                      "step_total_ms": (t_ctor + t_fit + t_ics + t_opt) * 1e3}))
