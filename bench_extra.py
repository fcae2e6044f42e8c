"""Extra legs of bench.py (reported under "extra" in its JSON line; never inside the headline's timed region).

  saas_sharded     BASELINE.json configs[3]: 256 SAAS hyper-parameter samples x n=512, d=1024 -- the samples are sharded over the
                   ranks (`shard_bounds`), every rank runs ONE batched fit (kernel build + Cholesky + inverse, grid.z = its samples)
                   and the batched posterior + LogEI over 4096 candidates; the exchange is one all-reduce of the mixture sums
                   (`distributed.mixture_moments` / `mixture_log_acq`).  Reference consumer: load_solvers.py:183-194 (saas_bo).
  qei_multistart   configs[4]: 512 restarts x q=4, 512 base samples, GP of configs[1] (n=1024, d=128) -- restarts sharded r::P;
                   ms per batched qEI value+gradient and per full 50-iteration L-BFGS run.  Reference consumer: ax_solver.py:83.
  grad_path        A6 at the headline size: acquisition value + gradient for 512 points at n=4096, d=256 (the inner call of
                   gen_candidates), algorithmic flops 2n^2 + 4nd + 2n per point against the DMMA ceiling.
  optimize_acqf    the whole `optimize_acqf(acqf, bounds, q=1, num_restarts=20, raw_samples=2^20)` call at the headline size: raw
                   samples generated in HBM, scored, initial conditions picked, restarts optimised.
Every leg is wrapped: a failure is reported as {"error": ...} and never takes the headline line down.
"""
import math
import os
import time
import traceback

import torch


def _guard(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as e:  # noqa: BLE001 -- reported in the JSON line
        return {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc().splitlines()[-3:]}


def _problem(n, d, noise=1e-3, seed=0):
    import bench
    return bench.synthetic_problem(n, d, noise, seed)


def saas_sharded(args, rank, world, dev, timed, cpu=True):
    import torch.distributed as dist

    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.distributed import mixture_log_acq, mixture_moments, shard_bounds
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import sample_saas_prior
    from hdbo_benchmark_b200.sampling import DeviceSobol

    S, n, d, m = 256, 512, 1024, 4096
    X, y, *_ = _problem(n, d)
    smp = sample_saas_prior(S, d, seed=4)
    s0, s1 = shard_bounds(S, rank, world)
    best = float(y.max())
    gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, batch=s1 - s0)
    gp.set_hyper(smp["lengthscale"][s0:s1].to(dev), smp["outputscale"][s0:s1], smp["noise"][s0:s1], smp["mean"][s0:s1])
    Z = DeviceSobol(d, scramble=True, seed=9, device=dev).draw(m)
    state = {}

    def fit_only():
        gp.fit_once()

    def post_only():
        mean, var, acq = gp.posterior_acq(Z, _lib.ACQ_LOGEI, best)
        e, v = mixture_moments(mean, var, S)
        la = mixture_log_acq(acq, S)
        state["pick"] = torch.argmax(la)
        return e, v, la

    def step():
        fit_only()
        return post_only()

    gp.fit()
    for _ in range(2):
        step()
    ms_fit, _ = timed(fit_only, 3)
    gp.fitted = True
    ms_post, _ = timed(post_only, 3)
    ms_step, (e, v, la) = timed(step, 3)
    out = {"workload": f"S={S} hyper-samples x n={n}, d={d}; posterior + LogEI over {m} candidates; samples sharded {world}-way",
           "samples_per_rank": s1 - s0, "var_mode": gp.var_mode if gp._i8_ok else "f64 (noise guard)",
           "fit_ms": ms_fit / 3, "posterior_mixture_ms": ms_post / 3, "step_ms": ms_step / 3,
           "sample_candidates_per_s": S * m / (ms_step / 3 / 1e3), "collective": "all-reduce of [2, m] mixture sums + [m] max + [m] sums",
           "fit_tflops": S * (n ** 3 * 2.0 / 3.0 + float(n) * n * d) / (ms_fit / 3 / 1e3) / 1e12,
           "mixture_argmax": int(state["pick"]), "i8_flag": gp.i8_flag()}
    if cpu and rank == 0 and world == 1:
        from oracle import gp_oracle as O
        sub, mc = 4, 512
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        states = O.saas_fit_states(X, y, smp["lengthscale"][:sub], smp["outputscale"][:sub], smp["noise"][:sub], smp["mean"][:sub])
        tf = time.perf_counter() - t0
        Zc = Z[:mc].cpu()
        with torch.no_grad():
            t0 = time.perf_counter()
            O.saas_posterior(states, Zc)
            tp = time.perf_counter() - t0
        out["cpu"] = {"fit_ms_extrapolated": tf * 1e3 * S / sub, "step_ms_extrapolated": (tf * S / sub + tp * (S / sub) * (m / mc)) * 1e3,
                      "sample": f"{sub} of {S} samples, {mc} of {m} candidates, linear extrapolation", "cores": torch.get_num_threads()}
    del gp
    torch.cuda.empty_cache()
    return out


def qei_multistart(args, rank, world, dev, timed, cpu=True):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.distributed import best_restart, strided_indices
    from hdbo_benchmark_b200.models import SingleTaskGP
    from hdbo_benchmark_b200.optim import gen_candidates_scipy
    from hdbo_benchmark_b200.sampling import DeviceSobol, SobolQMCNormalSampler

    n, d, R, q, S = 1024, 128, 512, 4, 512
    X, y, ls, os_, nz, mu = _problem(n, d)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(nu=2.5, ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([S]), seed=6)
    best_f = float(y.max()) - 0.5
    qei = A.qExpectedImprovement(model, best_f=best_f, sampler=sampler)
    ids = strided_indices(R, rank, world).to(dev)
    Zall = DeviceSobol(q * d, scramble=True, seed=5, device=dev).draw(R).reshape(R, q, d)
    Zq = Zall[ids].contiguous()

    def value_grad():
        Zg = Zq.clone().requires_grad_(True)
        v = qei(Zg)
        (g,) = torch.autograd.grad(v.sum(), Zg)
        return v, g

    for _ in range(3):
        value_grad()
    ms_vg, _ = timed(value_grad, 5)
    out = {"workload": f"qEI: {R} restarts x q={q}, {S} base samples, n={n}, d={d}; restarts sharded r::{world}",
           "restarts_per_rank": int(ids.numel()), "value_grad_ms": ms_vg / 5}

    # full multi-start run, fixed 50 iterations, every rank on its restarts, then the best restart over all ranks
    lo, hi = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    from hdbo_benchmark_b200 import optim as OPT
    runs = {}
    def run_device():
        c, v = OPT.gen_candidates_device(Zq, qei, lo, hi, options={"maxiter": 50, "ftol": 0.0, "gtol": 0.0, "check_every": 50})
        return best_restart(v.reshape(-1), c, ids)
    run_device()
    ms_dev, (bv, win, _) = timed(run_device, 2)
    info = OPT.gen_candidates_device.last_info
    runs["device_lbfgs_50it_ms"] = ms_dev / 2
    runs["device_best_value"] = float(bv)
    runs["device_iterations"] = info["iterations"]
    runs["device_restarts_converged_or_stalled"] = int(sum(1 for f in info["flags"] if f > 0))

    def run_scipy():
        c, v = gen_candidates_scipy(Zq, qei, lo, hi, options={"maxiter": 50, "ftol": 0.0, "gtol": 0.0, "maxfun": 60})
        return best_restart(v.reshape(-1), c, ids)
    t0 = time.perf_counter()
    bv, win, _ = run_scipy()
    torch.cuda.synchronize()
    runs["scipy_host_loop_50it_ms"] = (time.perf_counter() - t0) * 1e3
    runs["scipy_best_value"] = float(bv)
    out.update(runs)
    if cpu and rank == 0 and world == 1:
        from oracle import gp_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
        base = sampler.base_samples(q, dev).cpu()
        sub = 32
        Zc = Zq[:sub].cpu()
        O.qei_value_and_grad_autograd(st, Zc[:4], base, best_f)
        t0 = time.perf_counter()
        O.qei_value_and_grad_autograd(st, Zc, base, best_f)
        tc = time.perf_counter() - t0
        out["cpu"] = {"value_grad_ms_extrapolated": tc * 1e3 * R / sub, "lbfgs_50it_ms_lower_bound": 50 * tc * 1e3 * R / sub,
                      "sample": f"{sub} of {R} restarts (torch autograd on the oracle), linear extrapolation; the 50-iteration figure is "
                                "50 value+gradient evaluations (a lower bound: L-BFGS-B line searches add evaluations)",
                      "cores": torch.get_num_threads()}
    return out


def grad_path(args, rank, world, dev, gp, event_ms, peak, best_f):
    """A6 at n=4096, d=256: acquisition value + gradient for 512 points (hdbo_posterior_fwd -> hdbo_acq_elementwise ->
    hdbo_posterior_bwd), the inner call of the multi-start optimiser."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.sampling import DeviceSobol

    n, d, pts = gp.n, gp.d, 512
    Z = DeviceSobol(d, scramble=True, seed=21, device=dev).draw(pts)
    bufs = {}
    ts = event_ms(lambda: gp.acq_value_grad(Z, _lib.ACQ_LOGEI, best_f, 1.0, bufs), 5, warm=3)
    gp.attach_timeline(64)
    gp.acq_value_grad(Z, _lib.ACQ_LOGEI, best_f, 1.0, bufs)
    tl = gp.collect_timeline()
    gp.detach_timeline()
    flops = pts * (2.0 * n * n + 4.0 * n * d + 2.0 * n)
    return {"workload": f"LogEI value + d/dX for {pts} points at n={n}, d={d} (A6)", "ms": min(ts), "samples_ms": [round(t, 3) for t in ts],
            "algorithmic_flops": flops, "achieved_tflops": flops / (min(ts) * 1e-3) / 1e12, "peak_tflops": peak,
            "frac_of_fp64_dmma_peak": flops / (min(ts) * 1e-3) / 1e12 / peak, "points_per_s": pts / (min(ts) * 1e-3),
            "phases_ms": {k: round(v[0], 4) for k, v in tl.items()}}


def optimize_acqf_leg(args, dev, acqf, pool):
    """The caller of the headline path: optimize_acqf with the 2^20 raw-sample pool of configs[2]."""
    from hdbo_benchmark_b200.optim import optimize_acqf
    d = args.d
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
    opts = {"method": "device-lbfgs", "maxiter": 50, "seed": 3, "init_batch_limit": pool}

    def run():
        torch.manual_seed(0)
        return optimize_acqf(acqf, bounds, q=1, num_restarts=20, raw_samples=pool, options=opts)
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        x, v = run()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    # the screening part alone (raw samples generated in HBM + scored), for the split
    from hdbo_benchmark_b200.optim import gen_batch_initial_conditions
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gen_batch_initial_conditions(acqf, bounds, 1, 20, pool, options=opts)
    torch.cuda.synchronize()
    t_ic = (time.perf_counter() - t0) * 1e3
    return {"call": f"optimize_acqf(LogEI, bounds, q=1, num_restarts=20, raw_samples={pool}, method=device-lbfgs, maxiter=50)",
            "ms": min(ts), "gen_batch_initial_conditions_ms": t_ic, "screening_evals_per_s": pool / (t_ic / 1e3),
            "acq_value": float(v), "timing": "host wall clock around the blocking call (perf_counter + synchronize)"}


def run_all(args, rank, world, dev, model, acqf, gp, timed, event_ms, peak, best_f):
    import torch.distributed as dist

    out = {}
    out["grad_path"] = _guard(grad_path, args, rank, world, dev, gp, event_ms, peak, best_f)
    if rank == 0:
        pool = args.pool if args.scaling == "weak" or world == 1 else args.pool // world
        out["optimize_acqf"] = _guard(optimize_acqf_leg, args, dev, acqf, pool)
    if world > 1:
        dist.barrier()
    model._gp = None           # release the n=4096 state before the other configurations allocate theirs
    torch.cuda.empty_cache()
    out["saas_sharded"] = _guard(saas_sharded, args, rank, world, dev, timed)
    out["qei_multistart"] = _guard(qei_multistart, args, rank, world, dev, timed)
    return out
