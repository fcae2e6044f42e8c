"""botorch.optim.optimize_acqf and its helpers on top of the CUDA acquisition path.

Host control flow only: Sobol raw samples -> batched no-grad acquisition on the device (the large candidate
pool) -> Boltzmann / non-negative heuristic pick of `num_restarts` initial conditions (always keeping the
arg-max) -> joint L-BFGS-B over all restarts (scipy; f = -sum acqf, g from the analytic CUDA backward) with box
bounds -> best restart.  Invoked once per solver `step()` by the poli-baselines solvers selected at
/root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-158 (in-repo statement of `step()`:
/root/reference/src/hdbo_benchmark/utils/latent_space_solver.py:42-63).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
from scipy.optimize import minimize

from .sampling import draw_sobol_samples


def _compute_device(acq_function) -> torch.device:
    model = getattr(acq_function, "model", None)
    dev = getattr(model, "_dev", None)
    return dev if dev is not None else torch.device("cuda")


def initialize_q_batch(X: torch.Tensor, Y: torch.Tensor, n: int, eta: float = 1.0, generator=None) -> torch.Tensor:
    """Boltzmann heuristic of botorch.optim.initializers.initialize_q_batch: weights exp(eta * zscore(Y)), sample n
    without replacement, always keep the arg-max."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X
    Ystd = Y.std()
    if Ystd == 0 or not torch.isfinite(Ystd):
        return X[torch.randperm(n_samples, generator=generator)[:n].to(X.device)]
    max_val, max_idx = torch.max(Y, dim=0)
    Z = (Y - Y.mean()) / Ystd
    etaZ = eta * Z
    weights = torch.exp(etaZ)
    while torch.isinf(weights).any():
        etaZ = etaZ * 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights.cpu(), n, generator=generator).to(X.device)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def initialize_q_batch_nonneg(X, Y, n, eta: float = 1.0, alpha: float = 1e-4, generator=None):
    """Heuristic for non-negative acquisition values (EI, qEI): drop points below alpha * max, weights ~ value."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X
    max_val, max_idx = torch.max(Y, dim=0)
    if torch.any(max_val <= 0):
        return X[torch.randperm(n_samples, generator=generator)[:n].to(X.device)]
    pos = Y >= alpha * max_val
    num_pos = int(pos.sum())
    if num_pos < n:
        keep = pos.nonzero().reshape(-1)
        rest = (~pos).nonzero().reshape(-1)
        extra = rest[torch.randperm(rest.numel(), generator=generator)[: n - num_pos].to(rest.device)]
        idcs = torch.cat([keep, extra])
    else:
        w = torch.where(pos, torch.exp(eta * (Y / max_val - 1.0)), torch.zeros_like(Y))
        idcs = torch.multinomial(w.cpu(), n, generator=generator).to(X.device)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def gen_batch_initial_conditions(acq_function, bounds, q: int, num_restarts: int, raw_samples: int,
                                 options: Optional[Dict] = None, generator=None) -> torch.Tensor:
    """num_restarts x q x d initial conditions from `raw_samples` Sobol q-batches scored on the device."""
    options = options or {}
    dev = _compute_device(acq_function)
    bounds = torch.as_tensor(bounds, dtype=torch.float64)
    seed = options.get("seed")
    # raw samples are generated on the device (bit-identical to torch's SobolEngine): no host loop, no 2 GB host-to-device copy
    X_rnd = draw_sobol_samples(bounds, n=raw_samples, q=q, seed=seed, stream=options.get("sobol_stream"), device=dev)
    batch_limit = options.get("init_batch_limit", options.get("batch_limit", 1 << 20))
    with torch.no_grad():
        Y_rnd = torch.cat([acq_function(X_rnd[i:i + batch_limit]).to(dev, torch.float64)
                           for i in range(0, raw_samples, batch_limit)])
    if options.get("nonnegative", False):
        return initialize_q_batch_nonneg(X_rnd, Y_rnd, num_restarts, eta=options.get("eta", 1.0),
                                         alpha=options.get("alpha", 1e-4), generator=generator)
    return initialize_q_batch(X_rnd, Y_rnd, num_restarts, eta=options.get("eta", 1.0), generator=generator)


def gen_candidates_scipy(initial_conditions: torch.Tensor, acquisition_function: Callable, lower_bounds=None,
                         upper_bounds=None, options: Optional[Dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """L-BFGS-B over all restarts jointly: minimise -sum_b acqf(X_b); returns (candidates, acq values per restart)."""
    options = dict(options or {})
    dev = initial_conditions.device
    shape = initial_conditions.shape
    x0 = initial_conditions.detach().to(torch.float64).cpu().numpy().reshape(-1)
    d = shape[-1]
    reps = int(np.prod(shape[:-1]))
    lb = None if lower_bounds is None else np.tile(torch.as_tensor(lower_bounds, dtype=torch.float64).cpu().numpy().reshape(-1), reps)
    ub = None if upper_bounds is None else np.tile(torch.as_tensor(upper_bounds, dtype=torch.float64).cpu().numpy().reshape(-1), reps)
    bnds = None if lb is None and ub is None else list(zip(
        lb if lb is not None else [None] * x0.size, ub if ub is not None else [None] * x0.size))

    def f_and_g(x):
        X = torch.from_numpy(x.reshape(shape)).to(dev).requires_grad_(True)
        with torch.enable_grad():
            loss = -acquisition_function(X).sum()
        (g,) = torch.autograd.grad(loss, X)
        fval = float(loss.item())
        gnp = g.detach().to(torch.float64).cpu().numpy().reshape(-1)
        if not np.isfinite(fval) or not np.all(np.isfinite(gnp)):
            raise RuntimeError("acquisition value or gradient is not finite")
        return fval, gnp

    res = minimize(f_and_g, x0, jac=True, method=options.pop("method", "L-BFGS-B"), bounds=bnds,
                   options={"maxiter": options.pop("maxiter", 200), **{k: v for k, v in options.items()
                                                                     if k in ("ftol", "gtol", "maxfun", "maxcor")}})
    cand = torch.from_numpy(res.x.reshape(shape)).to(dev)
    if lb is not None or ub is not None:
        lo = None if lower_bounds is None else torch.as_tensor(lower_bounds, dtype=torch.float64, device=dev)
        hi = None if upper_bounds is None else torch.as_tensor(upper_bounds, dtype=torch.float64, device=dev)
        cand = torch.clamp(cand, min=lo, max=hi) if (lo is not None or hi is not None) else cand
    with torch.no_grad():
        vals = acquisition_function(cand)
    return cand, vals


def _device_lbfgs_supported(acq_function, q: int) -> bool:
    """The device optimiser drives the C ABI directly: an analytic (q = 1) acquisition or qExpectedImprovement (q <= 8, no pending
    points) on a single GP with an affine input transform."""
    from .acquisition import AnalyticAcquisitionFunction, qExpectedImprovement
    from .models import Normalize

    model = getattr(acq_function, "model", None)
    if model is None:
        return False
    if isinstance(acq_function, qExpectedImprovement):
        if not (1 <= q <= 8) or getattr(acq_function, "X_pending", None) is not None:
            return False
    elif q != 1 or not isinstance(acq_function, AnalyticAcquisitionFunction):
        return False
    if getattr(model, "_is_fully_bayesian", False):
        return False
    it = getattr(model, "input_transform", None)
    return it is None or isinstance(it, Normalize)


def gen_candidates_device(initial_conditions: torch.Tensor, acquisition_function, lower_bounds=None, upper_bounds=None,
                          options: Optional[Dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched box-constrained L-BFGS on the device: every restart is an independent problem of q*d variables, the
    whole line search of an iteration is one batched value+gradient evaluation, and the host only enqueues
    `hdbo_lbfgs_propose -> hdbo_posterior_fwd -> hdbo_acq_elementwise (q = 1) | hdbo_posterior_joint_cov + hdbo_qei (qEI) ->
    hdbo_posterior_bwd -> hdbo_lbfgs_accept` per iteration, polling the convergence flags every `check_every` iterations.
    Same contract as `gen_candidates_scipy`; initial_conditions [R, q, d]."""
    import ctypes as C

    from . import _lib
    from ._lib import HdboLbfgs, check, ptr
    from .engine import _stream

    options = dict(options or {})
    model = acquisition_function.model
    gp = model._device_gp()
    dev = gp.device
    f64 = dict(dtype=torch.float64, device=dev)
    from .acquisition import qExpectedImprovement
    is_qei = isinstance(acquisition_function, qExpectedImprovement)
    shape = initial_conditions.shape                      # [R, q, d]
    d = shape[-1]
    q = int(shape[-2]) if (is_qei and len(shape) >= 3) else 1
    R = int(np.prod(shape[:-1])) // q
    D = q * d                                             # variables per restart
    M, T = int(options.get("maxcor", 10)), int(options.get("line_search_trials", 8))
    maxiter = int(options.get("maxiter", 200))
    # optimise in the model's input space (affine image of the caller's box)
    it = getattr(model, "input_transform", None)
    off = torch.zeros(d, **f64) if it is None else it.offset.to(dev, torch.float64)
    coef = torch.ones(d, **f64) if it is None else it.coefficient.to(dev, torch.float64)
    inf = float("inf")
    lo = torch.full((d,), -inf, **f64) if lower_bounds is None else torch.as_tensor(lower_bounds, dtype=torch.float64).to(dev).reshape(-1).expand(d)
    hi = torch.full((d,), inf, **f64) if upper_bounds is None else torch.as_tensor(upper_bounds, dtype=torch.float64).to(dev).reshape(-1).expand(d)
    lo_t, hi_t = ((lo - off) / coef).repeat(q).contiguous(), ((hi - off) / coef).repeat(q).contiguous()
    off, coef = off.repeat(q), coef.repeat(q)
    x = ((initial_conditions.detach().to(dev, torch.float64).reshape(R, D) - off) / coef)
    x = torch.minimum(torch.maximum(x, lo_t), hi_t).contiguous()
    if is_qei:
        ot = getattr(model, "outcome_transform", None)
        scale = float(ot.stdvs.reshape(())) if ot is not None else 1.0
        shift = float(ot.means.reshape(())) if ot is not None else 0.0
        base = acquisition_function.sampler.base_samples(q, dev)
        best_std, jit = (float(acquisition_function.best_f) - shift) / scale, acquisition_function.jitter

        def value_grad(P, bufs):                          # P [rows, q*d] -> (qEI [rows], d/dP [rows, q*d]) in standardised units
            v, dZ = gp.qei_value_grad(P.view(-1, d), q, base, best_std, jit, bufs)
            return v, dZ.view(-1, D)
    else:
        scale, shift = acquisition_function._scale_shift()
        sign = 1.0 if acquisition_function.maximize else -1.0
        kind, param = acquisition_function._kind, acquisition_function._std_param(scale, shift, sign)

        def value_grad(P, bufs):
            return gp.acq_value_grad(P, kind, param, sign, bufs)
    bufs0, bufs = {}, {}
    a0, g0 = value_grad(x, bufs0)
    f = (-a0).clone()
    g = (-g0).clone()
    S = torch.zeros(R, M, D, **f64)
    Y = torch.zeros(R, M, D, **f64)
    rho = torch.zeros(R, M, **f64)
    meta = torch.zeros(R, 4, dtype=torch.int32, device=dev)
    step = torch.ones(R, **f64)
    Xt = torch.empty(R * T, D, **f64)
    ft = torch.empty(R * T, **f64)
    gt = torch.empty(R * T, D, **f64)
    st = HdboLbfgs(R=R, D=D, M=M, T=T, lo=ptr(lo_t), hi=ptr(hi_t), x=ptr(x), f=ptr(f), g=ptr(g), S=ptr(S), Y=ptr(Y), rho=ptr(rho),
                   step=ptr(step), meta=ptr(meta), Xt=ptr(Xt), ft=ptr(ft), gt=ptr(gt), ftol=float(options.get("ftol", 2.220446049250313e-09)),
                   gtol=float(options.get("gtol", 1e-5)), maxiter=maxiter)
    lib = gp.lib
    check_every = int(options.get("check_every", 10))

    def iteration():
        check(lib.hdbo_lbfgs_propose(C.byref(st), _stream()), "hdbo_lbfgs_propose")
        a, dZ = value_grad(Xt, bufs)
        torch.neg(a, out=ft)
        torch.neg(dZ, out=gt)
        check(lib.hdbo_lbfgs_accept(C.byref(st), _stream()), "hdbo_lbfgs_accept")

    # Two eager iterations (they also do the one-time kernel-attribute set-up and size the buffers), then the iteration -- ~15
    # small dependent launches -- is captured once into a CUDA graph and replayed; the host only polls the flags between blocks.
    done_iters = 0
    for _ in range(min(2, maxiter)):
        iteration()
        done_iters += 1
    graph = None
    if options.get("cuda_graph", True) and maxiter > done_iters:
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                iteration()
            done_iters += 1
        except Exception:   # capture not possible (e.g. an outer capture is active): run eagerly
            graph = None
    while done_iters < maxiter:
        for _ in range(min(check_every, maxiter - done_iters)):
            graph.replay() if graph is not None else iteration()
            done_iters += 1
        if bool((meta[:, 2] > 0).all()):
            break
    if not bool(torch.isfinite(f).all()):
        raise RuntimeError("acquisition value is not finite")
    gen_candidates_device.last_info = {"iterations": done_iters, "per_restart_iterations": meta[:, 3].tolist(),
                                       "flags": meta[:, 2].tolist(), "cuda_graph": graph is not None}
    cand = (x * coef + off).reshape(shape)
    with torch.no_grad():
        vals = acquisition_function(cand)
    return cand, vals


def optimize_acqf(acq_function, bounds, q: int, num_restarts: int, raw_samples: Optional[int] = None,
                  options: Optional[Dict] = None, batch_initial_conditions: Optional[torch.Tensor] = None,
                  return_best_only: bool = True, sequential: bool = False, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (X* [q x d], acquisition value) like botorch.optim.optimize_acqf."""
    if sequential and q > 1:
        cands, vals = [], []
        base_pending = getattr(acq_function, "X_pending", None)
        for _ in range(q):
            c, v = optimize_acqf(acq_function, bounds, 1, num_restarts, raw_samples, options, None, True, False)
            cands.append(c)
            vals.append(v)
            pend = torch.cat(cands, dim=-2)
            acq_function.set_X_pending(pend if base_pending is None else torch.cat([base_pending, pend], dim=-2))
        acq_function.set_X_pending(base_pending)
        return torch.cat(cands, dim=-2), torch.stack(vals)
    options = options or {}
    bounds_t = torch.as_tensor(bounds, dtype=torch.float64)
    out_device, out_dtype = bounds_t.device, (torch.as_tensor(bounds).dtype if torch.as_tensor(bounds).dtype.is_floating_point else torch.float64)
    if batch_initial_conditions is None:
        if raw_samples is None:
            raise ValueError("Must specify raw_samples when batch_initial_conditions is None")
        ics = gen_batch_initial_conditions(acq_function, bounds_t, q, num_restarts, raw_samples, options)
    else:
        ics = torch.as_tensor(batch_initial_conditions).to(_compute_device(acq_function), torch.float64)
    batch_limit = options.get("batch_limit", num_restarts)
    cands, vals = [], []
    # options["method"] = "device-lbfgs": batched L-BFGS on the device (analytic q = 1 acquisitions); default = scipy L-BFGS-B
    on_device = options.get("method") == "device-lbfgs" and _device_lbfgs_supported(acq_function, q)
    for i in range(0, ics.shape[0], batch_limit):
        if on_device:
            c, v = gen_candidates_device(ics[i:i + batch_limit], acq_function, bounds_t[0], bounds_t[1],
                                         options={k: v for k, v in options.items()
                                                  if k in ("maxiter", "ftol", "gtol", "maxcor", "line_search_trials", "check_every", "cuda_graph")})
            cands.append(c)
            vals.append(v.to(c.device, torch.float64))
            continue
        c, v = gen_candidates_scipy(ics[i:i + batch_limit], acq_function, bounds_t[0], bounds_t[1],
                                    options={k: v for k, v in options.items() if k in ("maxiter", "ftol", "gtol", "maxfun", "method")
                                             and v != "device-lbfgs"})
        cands.append(c)
        vals.append(v.to(c.device, torch.float64))
    cand = torch.cat(cands)
    val = torch.cat(vals)
    if return_best_only:
        best = torch.argmax(val)
        cand, val = cand[best], val[best]
    return cand.to(out_device, out_dtype), val.to(out_device, out_dtype)
