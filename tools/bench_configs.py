"""Secondary measurements for BASELINE.json configs #2, #4, #5 (the headline config #3 is bench.py).

Each line: GPU time through the package's public surface (CUDA events, 3 warm-ups, best of 5) next to the CPU oracle
(torch FP64, all host threads) on a bounded sample of the same workload.  Results are JSON lines on stdout.

    python tools/bench_configs.py [--skip-cpu]
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from hdbo_benchmark_b200 import _lib  # noqa: E402
from hdbo_benchmark_b200 import acquisition as A  # noqa: E402
from hdbo_benchmark_b200 import gp_modules as G  # noqa: E402
from hdbo_benchmark_b200.engine import DeviceGP  # noqa: E402
from hdbo_benchmark_b200.models import (SaasFullyBayesianSingleTaskGP, SingleTaskGP, mll_gradient,  # noqa: E402
                                        sample_saas_prior)
from hdbo_benchmark_b200.sampling import SobolQMCNormalSampler  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

dev = torch.device("cuda")


def gpu_ms(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def cpu_s(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()

    # ---- config #2: n=1024, d=128, Matern-5/2, EI over 65,536 raw samples ----
    n, d, m = 1024, 128, 65536
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    best_f = float(y.max())
    gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, keep_r2=True)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()
    Z = O.sobol_candidates(m, d)
    Zd = Z.to(dev)
    t_pool = gpu_ms(lambda: gp.posterior_acq(Zd, _lib.ACQ_EI, best_f, want_mean=False, want_var=False))
    t_mll = gpu_ms(lambda: (gp.fit_once(need_r2=True), mll_gradient(gp)))
    gp.fit()
    out = {"config": "#2 n=1024 d=128 EI over 65536 raw samples", "gpu_pool_ms": t_pool, "gpu_evals_per_s": m / t_pool * 1e3,
           "gpu_mll_value_grad_ms": t_mll}
    if not args.skip_cpu:
        st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
        with torch.no_grad():
            tc = cpu_s(lambda: O.expected_improvement(*O.posterior(st, Z[:8192]), best_f))

        def cpu_mll():
            l = ls.clone().requires_grad_(True)
            O.exact_mll(X, y, l, torch.tensor(os_, requires_grad=True), torch.tensor(nz, requires_grad=True),
                        torch.tensor(mu, requires_grad=True), O.MATERN52).backward()
        out.update({"cpu_evals_per_s": 8192 / tc, "cpu_mll_value_grad_ms": cpu_s(cpu_mll) * 1e3, "cpu_cores": cores,
                    "cpu_sample": "8192 candidates; MLL by torch autograd"})
    print(json.dumps(out), flush=True)

    # ---- config #5: qEI, R=512 restarts x q=4, 512 MC samples, same GP ----
    R, q, S = 512, 4, 512
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.eval()
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([S]), seed=6)
    qei = A.qExpectedImprovement(model, best_f=best_f, sampler=sampler)
    Zq = O.sobol_candidates(R * q, d, seed=5).reshape(R, q, d).to(dev)

    def qei_fg():
        Zg = Zq.clone().requires_grad_(True)
        v = qei(Zg)
        (g,) = torch.autograd.grad(v.sum(), Zg)
        return v, g
    t_q = gpu_ms(qei_fg)
    out = {"config": "#5 qEI value+grad, 512 restarts x q=4, 512 MC samples, n=1024 d=128", "gpu_ms": t_q}
    if not args.skip_cpu:
        st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
        base = sampler.base_samples(q, dev).cpu()
        sub = 64
        tc = cpu_s(lambda: O.qei_value_and_grad_autograd(st, Zq[:sub].cpu(), base, best_f), reps=1)
        out.update({"cpu_ms_extrapolated": tc * 1e3 * R / sub, "cpu_cores": cores, "cpu_sample": f"{sub} of {R} restarts, linear extrapolation"})
    print(json.dumps(out), flush=True)
    del model, gp

    # ---- config #4: SAAS-style batch, S=256 hyper-samples x n=512, d=1024 ----
    S, n, d, m = 256, 512, 1024, 4096
    X, y, *_ = O.synthetic_problem(n, d)
    smp = sample_saas_prior(S, d, seed=4)
    gpb = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, batch=S)
    gpb.set_hyper(smp["lengthscale"].to(dev), smp["outputscale"], smp["noise"], smp["mean"])
    t_fit = gpu_ms(lambda: gpb.fit_once(), reps=3, warm=2)
    gpb.fit()
    Z = O.sobol_candidates(m, d, seed=9)
    Zd = Z.to(dev)
    t_post = gpu_ms(lambda: gpb.posterior_acq(Zd, _lib.ACQ_LOGEI, float(y.max()), want_mean=False, want_var=False), reps=3, warm=2)
    out = {"config": "#4 SAAS batch S=256 x n=512, d=1024: batched build+Cholesky+inverse; posterior+LogEI over 4096 candidates",
           "gpu_fit_ms": t_fit, "gpu_posterior_ms": t_post, "gpu_candidates_per_s": m / t_post * 1e3,
           "jitter": gpb.jitter_used}
    if not args.skip_cpu:
        sub = 8
        def cpu_fit():
            return O.saas_fit_states(X, y, smp["lengthscale"][:sub], smp["outputscale"][:sub], smp["noise"][:sub], smp["mean"][:sub])
        tf = cpu_s(cpu_fit, reps=1)
        states = cpu_fit()
        with torch.no_grad():
            tp = cpu_s(lambda: O.saas_posterior(states, Z[:1024]), reps=1)
        out.update({"cpu_fit_ms_extrapolated": tf * 1e3 * S / sub, "cpu_candidates_per_s_extrapolated": 1024 / (tp * S / sub),
                    "cpu_cores": cores, "cpu_sample": f"{sub} of {S} hyper-samples (fit), 1024 candidates (posterior), linear extrapolation"})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
