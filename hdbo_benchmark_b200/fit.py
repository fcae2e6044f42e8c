"""botorch.fit.fit_gpytorch_mll: L-BFGS-B over the raw hyper-parameters.

The closure evaluates -MLL and its gradient with one `hdbo_gp_fit` + `hdbo_mll_grad` pair on the device; the
float64 ndarray round trip per iteration is the same host loop BoTorch's `fit_gpytorch_mll_scipy` runs.  On a
failed / non-converged attempt the hyper-parameters are re-initialised (from their priors where they have one)
and the fit is retried, up to `max_attempts` (BoTorch: 5); if every attempt fails the original state is restored.
"""
from __future__ import annotations

import copy
import math
import warnings
from typing import Dict, Optional

import numpy as np
import torch
from scipy.optimize import minimize

from .engine import NotPSDError


def _named_raw_parameters(mll):
    model = mll.model
    out = []
    for name, p in model.named_parameters():
        if p.requires_grad:
            out.append((name, p))
    return out


def _bounds_for(mll, names_and_params):
    model = mll.model
    table = {}
    for prefix, mod in (("covar_module.", model.covar_module), ("likelihood.", model.likelihood)):
        table.update(mod.raw_parameter_bounds(prefix))
    lo, hi = [], []
    for name, p in names_and_params:
        l, h = table.get(name, (-math.inf, math.inf))
        lo += [l] * p.numel()
        hi += [h] * p.numel()
    return [(None if math.isinf(l) else l, None if math.isinf(h) else h) for l, h in zip(lo, hi)]


def _sample_from_priors(mll, generator=None):
    """Re-initialise constrained values from their priors (botorch sample_all_priors), clamped into the constraint."""
    model = mll.model
    for mod in (model.covar_module, model.likelihood):
        stack = [mod]
        while stack:
            m = stack.pop()
            for name in getattr(m, "_hyper_names", []):
                prior = getattr(m, f"{name}_prior", None)
                cons = getattr(m, f"raw_{name}_constraint")
                raw = getattr(m, f"raw_{name}")
                if prior is None:
                    continue
                if hasattr(prior, "loc") and prior.__class__.__name__ == "LogNormalPrior":
                    v = torch.exp(prior.loc + prior.scale * torch.randn(raw.shape, dtype=torch.float64, generator=generator).to(raw.device))
                elif prior.__class__.__name__ == "GammaPrior":
                    v = torch.distributions.Gamma(prior.concentration.cpu(), prior.rate.cpu()).sample(raw.shape).to(raw.device)
                else:
                    continue
                lo = float(cons.lower_bound)
                v = v.clamp_min(lo * (1 + 1e-6) + 1e-12) if not math.isinf(lo) else v
                m._set(name, v)
            stack.extend(m.children())


def fit_gpytorch_mll_scipy(mll, options: Optional[Dict] = None, bounds=None):
    """One L-BFGS-B run; returns the scipy OptimizeResult."""
    options = dict(options or {})
    model = mll.model
    params = _named_raw_parameters(mll)
    bnds = bounds if bounds is not None else _bounds_for(mll, params)
    sizes = [p.numel() for _, p in params]

    def pack():
        return np.concatenate([p.detach().to(torch.float64).cpu().numpy().reshape(-1) for _, p in params])

    dev = params[0][1].device

    def unpack(x):
        # one host->device copy for all parameters, then device-side slices
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        off = 0
        with torch.no_grad():
            for (_, p), sz in zip(params, sizes):
                p.copy_(xd[off:off + sz].view_as(p))
                off += sz

    def closure(x):
        unpack(x)
        for _, p in params:
            p.grad = None
        try:
            loss = -mll(model(model.train_inputs[0]), model.train_targets)
            loss = loss.sum()
            loss.backward()
        except NotPSDError:
            return float("inf"), np.zeros_like(x)
        # one device->host copy (and one sync) for the loss and every gradient
        packed = torch.cat([loss.detach().reshape(1)] + [(p.grad if p.grad is not None else torch.zeros_like(p)).detach().reshape(-1)
                                                          for _, p in params]).cpu().numpy()
        f, g = float(packed[0]), packed[1:].copy()
        if not np.isfinite(f) or not np.all(np.isfinite(g)):
            return float("inf"), np.zeros_like(x)
        return f, g

    was_training = model.training
    model.train()
    res = minimize(closure, pack(), jac=True, method="L-BFGS-B", bounds=bnds,
                   options={"maxiter": options.get("maxiter", 2000), "ftol": options.get("ftol", 2.220446049250313e-09),
                            "gtol": options.get("gtol", 1e-05), "maxfun": options.get("maxfun", 15000)})
    unpack(res.x)
    model.train(was_training)
    return res


def fit_gpytorch_mll(mll, optimizer_kwargs: Optional[Dict] = None, max_attempts: int = 5, **kwargs):
    """Fit the model in place, return the mll in eval mode (BoTorch semantics)."""
    optimizer_kwargs = optimizer_kwargs or {}
    options = optimizer_kwargs.get("options", optimizer_kwargs)
    original = copy.deepcopy(mll.model.state_dict())
    best = None
    for attempt in range(max_attempts):
        if attempt > 0:
            mll.model.load_state_dict(original)
            _sample_from_priors(mll)
        try:
            res = fit_gpytorch_mll_scipy(mll, options=options)
        except (NotPSDError, RuntimeError) as e:  # pragma: no cover - numerical failure path
            warnings.warn(f"fit_gpytorch_mll attempt {attempt + 1} failed: {e}")
            continue
        if np.isfinite(res.fun) and (best is None or res.fun < best[0]):
            best = (float(res.fun), copy.deepcopy(mll.model.state_dict()))
        if res.success or res.status == 1 and np.isfinite(res.fun):  # converged, or hit maxiter with a finite value
            break
    if best is None:
        mll.model.load_state_dict(original)
        warnings.warn("fit_gpytorch_mll: all attempts failed; hyper-parameters restored")
    else:
        mll.model.load_state_dict(best[1])
    mll.eval()
    mll.model.eval()
    return mll
