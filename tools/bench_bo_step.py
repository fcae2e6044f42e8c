"""Wall-clock anatomy of one BO solver step at the reference's own scale (BASELINE config #1 shape: d=128, n=310, and n=1000):
SingleTaskGP -> fit_gpytorch_mll (scipy L-BFGS-B; fast closure = numpy transforms + CUDA-graph replay of fit+gradient, vs the
generic autograd closure) -> LogEI -> optimize_acqf (device L-BFGS vs scipy)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from hdbo_benchmark_b200 import acquisition as A
from hdbo_benchmark_b200.fit import fit_gpytorch_mll_device, fit_gpytorch_mll_scipy
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
from hdbo_benchmark_b200.optim import optimize_acqf


def sync_time(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, (time.perf_counter() - t) * 1e3


import gc
sizes = [(int(a), 128) for a in sys.argv[1:]] or [(310, 128), (1000, 128)]
for n, d in sizes:
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = (torch.sin(3 * X[:, :4].sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
    out = {"n": n, "d": d}
    for rep in range(2):                                   # second repetition is the reported one (warm)
        for name, fast in (("fit_fast", True), ("fit_generic", False)):
            model = SingleTaskGP(X, y)
            mll = ExactMarginalLogLikelihood(model.likelihood, model)
            res, ms = sync_time(lambda: fit_gpytorch_mll_scipy(mll, options={"fast_closure": fast}))
            out[name + "_ms"], out[name + "_evals"], out[name + "_ms_per_eval"] = ms, int(res.nfev), ms / max(int(res.nfev), 1)
            out[name + "_neg_mll"] = float(res.fun)
            del mll
            gc.collect(); torch.cuda.empty_cache()
        # batched-trial L-BFGS on the device (8 line-search points per iteration), then the scipy polish from its answer
        m2 = SingleTaskGP(X, y)
        mll2 = ExactMarginalLogLikelihood(m2.likelihood, m2)
        res, ms = sync_time(lambda: fit_gpytorch_mll_device(mll2, options={"device_max_n": 10 ** 9}))   # (measure it at every n)
        out["fit_device_ms"], out["fit_device_iters"], out["fit_device_neg_mll"] = ms, int(res.nit), float(res.fun)
        res, ms2 = sync_time(lambda: fit_gpytorch_mll_scipy(mll2, options={}))
        out["fit_device_polish_ms"], out["fit_device_polish_evals"], out["fit_device_polished_neg_mll"] = ms2, int(res.nfev), float(res.fun)
        out["fit_device_total_ms"] = ms + ms2
        del mll2, m2
        gc.collect(); torch.cuda.empty_cache()
        model.eval()
        acqf = A.LogExpectedImprovement(model, best_f=float(y.max()))
        bounds = torch.stack([torch.zeros(d), torch.ones(d)]).double()
        for name, opts in (("acq_device", {"seed": rep, "method": "device-lbfgs"}), ("acq_scipy", {"seed": rep})):
            torch.manual_seed(rep)
            optimize_acqf(acqf, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts)     # warm (Sobol cache, kernel attributes)
            torch.manual_seed(rep)
            (cand, val), ms = sync_time(lambda: optimize_acqf(acqf, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts))
            out[name + "_ms"], out[name + "_value"] = ms, float(val)
    out["step_ms_device_paths"] = out["fit_device_total_ms"] + out["acq_device_ms"]
    out["step_ms_fast_paths"] = out["fit_fast_ms"] + out["acq_device_ms"]
    out["step_ms_generic_paths"] = out["fit_generic_ms"] + out["acq_scipy_ms"]
    print(json.dumps(out), flush=True)
