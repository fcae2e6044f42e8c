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


# ----------------------------------------------------------------------------------------------------------------------
# Fast closure: the O(d) raw -> constrained transforms and prior densities in numpy on the host, ONE host->device copy of
# (lengthscale[d], outputscale, noise, mean), a CUDA-graph replay of {hyper-parameter upload, hdbo_gp_fit, hdbo_mll_grad, pack},
# ONE device->host copy of (n*mll, gradient[d+3], info).  The generic closure below does the same arithmetic through ~80 small
# torch launches + autograd per evaluation (1.4 ms at n = 310 of which 0.4 ms is GPU work); this one is bound by the GPU work.
# ----------------------------------------------------------------------------------------------------------------------
def _np_sigmoid(x):
    return np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x))))


def _np_softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def _constraint_np(c):
    """numpy (value, dvalue/draw) of a gp_modules.Interval; None if the constraint type is not known here."""
    from .gp_modules import Interval
    if c is None:
        return lambda r: (r, np.ones_like(r))
    if not isinstance(c, Interval):
        return None
    lo, hi, tr = float(c.lower_bound), float(c.upper_bound), c._transform
    if tr is None:
        return lambda r: (r, np.ones_like(r))
    if tr == "softplus":
        if math.isinf(hi):
            return lambda r: (_np_softplus(r) + lo, _np_sigmoid(r))
        return lambda r: (hi - _np_softplus(r), -_np_sigmoid(r))
    if tr == "sigmoid":
        def f(r):
            sg = _np_sigmoid(r)
            return sg * (hi - lo) + lo, (hi - lo) * sg * (1.0 - sg)
        return f
    return None


def _prior_np(prior):
    """numpy (log p(v) summed, dlogp/dv) of a gp_modules prior; None if unknown."""
    from .gp_modules import GammaPrior, HalfCauchyPrior, LogNormalPrior, NormalPrior
    if prior is None:
        return lambda v: (0.0, np.zeros_like(v))
    if isinstance(prior, GammaPrior):
        c, r = float(prior.concentration), float(prior.rate)
        k = c * math.log(r) - math.lgamma(c)
        return lambda v: (float(np.sum(k + (c - 1.0) * np.log(v) - r * v)), (c - 1.0) / v - r)
    if isinstance(prior, LogNormalPrior):
        mu, sg = float(prior.loc), float(prior.scale)
        k = -math.log(sg) - 0.5 * math.log(2 * math.pi)
        def f(v):
            lv = np.log(v)
            return float(np.sum(-lv + k - 0.5 * ((lv - mu) / sg) ** 2)), -1.0 / v - (lv - mu) / (sg * sg * v)
        return f
    if isinstance(prior, NormalPrior):
        mu, sg = float(prior.loc), float(prior.scale)
        k = -math.log(sg) - 0.5 * math.log(2 * math.pi)
        return lambda v: (float(np.sum(k - 0.5 * ((v - mu) / sg) ** 2)), -(v - mu) / (sg * sg))
    if isinstance(prior, HalfCauchyPrior):
        sg = float(prior.scale)
        k = math.log(2.0 / math.pi) - math.log(sg)
        return lambda v: (float(np.sum(k - np.log1p((v / sg) ** 2))), -2.0 * v / (sg * sg + v * v))
    return None


class _FastClosure:
    """f(x) = -mll, df/dx for SingleTaskGP-shaped models (ARD or isotropic lengthscale, optional ScaleKernel, homoskedastic
    noise, constant mean).  `build` returns None when the model uses anything it does not know (then the generic closure runs)."""

    @classmethod
    def build(cls, mll, params):
        from .gp_modules import ConstantMean, ScaleKernel
        from .models import ExactMarginalLogLikelihood, SingleTaskGP
        model = mll.model
        if not isinstance(mll, ExactMarginalLogLikelihood) or not isinstance(model, SingleTaskGP) or len(model.batch_shape):
            return None
        if not isinstance(model.mean_module, ConstantMean) or not torch.cuda.is_available():
            return None
        d = model.train_inputs[0].shape[-1]
        slots = {"raw_lengthscale": ("ls", model._base_kernel), "raw_noise": ("noise", None), "raw_constant": ("mean", model.mean_module)}
        if isinstance(model.covar_module, ScaleKernel):
            slots["raw_outputscale"] = ("os", model.covar_module)
        blocks, off = [], 0
        for name, p in params:
            leaf = name.rsplit(".", 1)[-1]
            if leaf not in slots:
                return None
            role = slots[leaf][0]
            owner = model
            for part in name.split(".")[:-1]:
                owner = getattr(owner, part)
            hname = leaf[len("raw_"):]
            cons = getattr(owner, f"raw_{hname}_constraint", None) if role != "mean" else None
            prior = getattr(owner, f"{hname}_prior", None)
            cf, pf = _constraint_np(cons), _prior_np(prior)
            if cf is None or pf is None or p.numel() not in ((1, d) if role == "ls" else (1,)):
                return None
            blocks.append((role, off, p.numel(), cf, pf))
            off += p.numel()
        roles = [b[0] for b in blocks]
        if len(set(roles)) != len(roles) or "ls" not in roles or "noise" not in roles:
            return None
        self = cls()
        self.mll, self.model, self.blocks, self.d, self.nx = mll, model, blocks, d, off
        self.has_os = "os" in roles
        self.fixed = {"os": 1.0, "mean": float(model.mean_module.constant.detach().reshape(-1)[0]) if "mean" not in roles else 0.0}
        self.n = model.train_targets.shape[-1]
        self.graph = None
        return self

    def _capture(self):
        from .engine import DeviceGP
        from .models import mll_gradient
        model, d = self.model, self.d
        gp = model._gp
        if gp is None or gp.batch != 1 or not gp.keep_r2:
            gp = model._gp = DeviceGP(model.train_inputs[0], model.train_targets, model._base_kernel.kind, batch=1, keep_r2=True)
        self.gp = gp
        model._fitted_key = None                 # the device state no longer mirrors the module parameters
        dev = gp.device
        self.h_host = torch.empty(d + 3, dtype=torch.float64).pin_memory()
        self.h_dev = torch.empty(d + 3, dtype=torch.float64, device=dev)
        self.out_dev = torch.empty(d + 5, dtype=torch.float64, device=dev)
        self.out_host = torch.empty(d + 5, dtype=torch.float64).pin_memory()

        def body():
            self.h_dev.copy_(self.h_host, non_blocking=True)
            torch.reciprocal(self.h_dev[:d], out=gp.inv_ls[0])
            gp.hyp[0, :3].copy_(self.h_dev[d:])
            gp.hyp[0, 3] = 0.0
            gp.fit_once(need_r2=True)
            grad = mll_gradient(gp)
            self.out_dev[0].copy_(gp.scal[0, 2])
            self.out_dev[1:d + 4].copy_(grad[0])
            self.out_dev[d + 4].copy_(gp.info[0].to(torch.float64))
            self.out_host.copy_(self.out_dev, non_blocking=True)

        self.body = body
        body()                                   # eager once: one-time kernel attributes, scratch buffers
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self.graph = g
        except Exception:                        # capture unavailable: the eager body still saves the torch launches
            self.graph = None
        self.captured = True

    def __call__(self, x):
        if not getattr(self, "captured", False):
            self._capture()
        d = self.d
        vals = dict(self.fixed)
        dv, lp, dlp = {}, 0.0, {}
        for role, off, sz, cf, pf in self.blocks:
            v, dvdr = cf(x[off:off + sz])
            if role != "mean" and np.any(v <= 0.0):
                return float("inf"), np.zeros_like(x)
            l, dl = pf(v)
            lp += l
            vals[role], dv[role], dlp[role] = v, dvdr, dl
        h = self.h_host.numpy()
        h[:d] = vals["ls"] if np.size(vals["ls"]) == d else np.full(d, float(np.ravel(vals["ls"])[0]))
        h[d], h[d + 1], h[d + 2] = float(np.ravel(vals["os"])[0]), float(np.ravel(vals["noise"])[0]), float(np.ravel(vals["mean"])[0])
        if self.graph is not None:
            self.graph.replay()
        else:
            self.body()
        torch.cuda.current_stream().synchronize()
        out = self.out_host.numpy()
        if out[d + 4] != 0.0 or not np.all(np.isfinite(out)):
            return None                          # not positive definite without jitter: let the generic closure decide
        gdata = {"ls": out[1:d + 1], "os": out[d + 1:d + 2], "noise": out[d + 2:d + 3], "mean": out[d + 3:d + 4]}
        g = np.zeros_like(x)
        for role, off, sz, cf, pf in self.blocks:
            gd = gdata[role] if sz == np.size(gdata[role]) else np.array([gdata[role].sum()])
            g[off:off + sz] = -(gd + dlp[role]) * dv[role] / self.n
        return -(out[0] + lp) / self.n, g


# ----------------------------------------------------------------------------------------------------------------------
# Device fit (the fit_gpytorch_mll half): the L-BFGS iteration itself runs on the device.  The T trial points of a
# backtracking line search are T hyper-parameter vectors = ONE batched hdbo_gp_fit + hdbo_mll_grad (grid.z = T; at BO sizes the fit is
# latency-bound, so T members cost about as much as one), the raw -> constrained transforms and prior densities are O(T d) torch
# ops, and hdbo_lbfgs_propose / hdbo_lbfgs_accept do the two-loop recursion and the Armijo pick.  The whole iteration is one CUDA
# graph; the host polls the `done` flag every few replays.  A trial that is not positive definite has an infinite objective
# (rejected by the line search), as in the scipy closure.
# ----------------------------------------------------------------------------------------------------------------------
def _constraint_t(c):
    from .gp_modules import Interval
    sp, sg = torch.nn.functional.softplus, torch.sigmoid
    if c is None:
        return lambda r: (r, torch.ones_like(r))
    if not isinstance(c, Interval):
        return None
    lo, hi, tr = float(c.lower_bound), float(c.upper_bound), c._transform
    if tr is None:
        return lambda r: (r, torch.ones_like(r))
    if tr == "softplus":
        if math.isinf(hi):
            return lambda r: (sp(r) + lo, sg(r))
        return lambda r: (hi - sp(r), -sg(r))
    if tr == "sigmoid":
        def f(r):
            s_ = sg(r)
            return s_ * (hi - lo) + lo, (hi - lo) * s_ * (1.0 - s_)
        return f
    return None


def _prior_t(prior):
    """torch (log p(v) summed over the last axis, dlogp/dv) on [T, size] tensors; None if unknown."""
    from .gp_modules import GammaPrior, HalfCauchyPrior, LogNormalPrior, NormalPrior
    if prior is None:
        return lambda v: (torch.zeros_like(v[:, 0]), torch.zeros_like(v))
    if isinstance(prior, GammaPrior):
        c, r = float(prior.concentration), float(prior.rate)
        k = c * math.log(r) - math.lgamma(c)
        return lambda v: ((k + (c - 1.0) * torch.log(v) - r * v).sum(1), (c - 1.0) / v - r)
    if isinstance(prior, LogNormalPrior):
        mu, sg = float(prior.loc), float(prior.scale)
        k = -math.log(sg) - 0.5 * math.log(2 * math.pi)
        def f(v):
            lv = torch.log(v)
            return (-lv + k - 0.5 * ((lv - mu) / sg) ** 2).sum(1), -1.0 / v - (lv - mu) / (sg * sg * v)
        return f
    if isinstance(prior, NormalPrior):
        mu, sg = float(prior.loc), float(prior.scale)
        k = -math.log(sg) - 0.5 * math.log(2 * math.pi)
        return lambda v: ((k - 0.5 * ((v - mu) / sg) ** 2).sum(1), -(v - mu) / (sg * sg))
    if isinstance(prior, HalfCauchyPrior):
        sg = float(prior.scale)
        k = math.log(2.0 / math.pi) - math.log(sg)
        return lambda v: ((k - torch.log1p((v / sg) ** 2)).sum(1), -2.0 * v / (sg * sg + v * v))
    return None


class _Result(dict):
    __getattr__ = dict.get


def fit_gpytorch_mll_device(mll, options: Optional[Dict] = None, bounds=None):
    """One batched-trial L-BFGS run on the device; returns an OptimizeResult-like object (fun, x, nit, nfev, success, status), or
    None when the model is not SingleTaskGP-shaped (the caller then uses the scipy path)."""
    import ctypes as C

    from ._lib import HdboLbfgs, check, ptr
    from .engine import DeviceGP, _stream
    from .models import mll_gradient

    options = dict(options or {})
    model = mll.model
    params = _named_raw_parameters(mll)
    meta_np = _FastClosure.build(mll, params)             # block layout / applicability test shared with the host closure
    if meta_np is None:
        return None
    blocks = []
    for (name, p), (role, off, sz, _, _) in zip(params, meta_np.blocks):
        owner = model
        for part in name.split(".")[:-1]:
            owner = getattr(owner, part)
        hname = name.rsplit(".", 1)[-1][len("raw_"):]
        cf = _constraint_t(getattr(owner, f"raw_{hname}_constraint", None) if role != "mean" else None)
        pf = _prior_t(getattr(owner, f"{hname}_prior", None))
        if cf is None or pf is None:
            return None
        blocks.append((role, off, sz, cf, pf))
    d, D, n = meta_np.d, meta_np.nx, meta_np.n
    # The batched line search pays off while one fit is latency-bound (8 members cost about one).  Measured at d = 128: n = 310 85 ms
    # vs 677 ms (scipy, fast closure), n = 1000 327 vs 869 ms, but n = 4096 -- compute-bound, every iteration costs T evaluations --
    # 2.2 s vs 0.9 s: above `device_max_n` the scipy path runs.
    if n > int(options.get("device_max_n", 1536)):
        return None
    T = int(options.get("line_search_trials", 8))   # measured: 2-4 trial steps stop early on ARD fits (n = 1000: -mll/n 0.56 vs 0.35)
    M = int(options.get("maxcor", 10))
    maxiter = int(options.get("maxiter", 2000))
    gp = DeviceGP(model.train_inputs[0], model.train_targets, model._base_kernel.kind, batch=T, keep_r2=True)
    dev = gp.device
    f64 = dict(dtype=torch.float64, device=dev)
    bnds = bounds if bounds is not None else _bounds_for(mll, params)
    inf = float("inf")
    lo = torch.tensor([-inf if b[0] is None else b[0] for b in bnds], **f64)
    hi = torch.tensor([inf if b[1] is None else b[1] for b in bnds], **f64)
    fixed_os = torch.ones(T, **f64)
    fixed_mean = torch.full((T,), float(meta_np.fixed["mean"]), **f64)

    def evaluate(Xr: torch.Tensor, f_out: torch.Tensor, g_out: torch.Tensor):
        """Xr [T, D] raw parameters -> f_out [T] = -mll (with priors) / n, g_out [T, D]."""
        vals, dv, dlp = {"os": fixed_os, "mean": fixed_mean}, {}, {}
        lp = torch.zeros(T, **f64)
        for role, off, sz, cf, pf in blocks:
            v, dvdr = cf(Xr[:, off:off + sz])
            l, dl = pf(v)
            lp = lp + l
            vals[role], dv[role], dlp[role] = v, dvdr, dl
        ls = vals["ls"] if vals["ls"].shape[1] == d else vals["ls"].expand(T, d)
        gp.set_hyper(ls, vals["os"].reshape(T), vals["noise"].reshape(T), vals["mean"].reshape(T))
        gp.fit_once(need_r2=True)
        grad = mll_gradient(gp)                            # [T, d + 3]: d(n mll)/d(ls, os, noise, mean)
        gdata = {"ls": grad[:, :d], "os": grad[:, d:d + 1], "noise": grad[:, d + 1:d + 2], "mean": grad[:, d + 2:d + 3]}
        bad = (gp.info.reshape(T) != 0)
        f = -(gp.scal[:, 2] + lp) / n
        bad = bad | ~torch.isfinite(f)
        f_out.copy_(torch.where(bad, torch.full_like(f, inf), f))
        for role, off, sz, cf, pf in blocks:
            gd = gdata[role] if sz == gdata[role].shape[1] else gdata[role].sum(1, keepdim=True)
            g_out[:, off:off + sz] = torch.where(bad.unsqueeze(1), torch.zeros_like(dv[role]), -(gd + dlp[role]) * dv[role] / n)

    x = torch.cat([p.detach().to(dev, torch.float64).reshape(-1) for _, p in params]).reshape(1, D)
    x = torch.minimum(torch.maximum(x, lo), hi).contiguous()
    Xt = x.expand(T, D).contiguous()
    ft = torch.empty(T, **f64)
    gt = torch.zeros(T, D, **f64)
    evaluate(Xt, ft, gt)
    f = ft[:1].clone()
    g = gt[:1].clone()
    if not bool(torch.isfinite(f).all()):
        return _Result(fun=inf, x=x.reshape(-1).cpu().numpy(), nit=0, nfev=T, success=False, status=2, message="not positive definite at the start")
    S = torch.zeros(1, M, D, **f64)
    Y = torch.zeros(1, M, D, **f64)
    rho = torch.zeros(1, M, **f64)
    meta = torch.zeros(1, 4, dtype=torch.int32, device=dev)
    step = torch.ones(1, **f64)
    st = HdboLbfgs(R=1, D=D, M=M, T=T, lo=ptr(lo), hi=ptr(hi), x=ptr(x), f=ptr(f), g=ptr(g), S=ptr(S), Y=ptr(Y), rho=ptr(rho), step=ptr(step),
                   meta=ptr(meta), Xt=ptr(Xt), ft=ptr(ft), gt=ptr(gt), ftol=float(options.get("ftol", 2.220446049250313e-09)),
                   gtol=float(options.get("gtol", 1e-05)), maxiter=maxiter)
    lib = gp.lib

    def iteration():
        check(lib.hdbo_lbfgs_propose(C.byref(st), _stream()), "hdbo_lbfgs_propose")
        evaluate(Xt, ft, gt)
        check(lib.hdbo_lbfgs_accept(C.byref(st), _stream()), "hdbo_lbfgs_accept")

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
        except Exception:
            graph = None
    check_every = int(options.get("check_every", 25))
    while done_iters < maxiter:
        for _ in range(min(check_every, maxiter - done_iters)):
            graph.replay() if graph is not None else iteration()
            done_iters += 1
        if int(meta[0, 2]) > 0:
            break
    flag, nit = int(meta[0, 2]), int(meta[0, 3])
    xs = x.reshape(-1)
    with torch.no_grad():
        off = 0
        for _, p in params:
            p.copy_(xs[off:off + p.numel()].view_as(p).to(p.device, p.dtype))
            off += p.numel()
    model._fitted_key = None
    fit_gpytorch_mll_device.last_info = {"replays": done_iters, "iterations": nit, "flag": flag, "trials": T, "cuda_graph": graph is not None}
    return _Result(fun=float(f[0]), x=xs.cpu().numpy(), nit=nit, nfev=(done_iters + 1) * T, success=flag in (1, 2) and nit < maxiter,
                   status=0 if flag in (1, 2) and nit < maxiter else 1, message={0: "maxiter", 1: "converged", 2: "line search stalled"}.get(flag, ""))


def fit_gpytorch_mll_scipy(mll, options: Optional[Dict] = None, bounds=None):
    """One L-BFGS-B run; returns the scipy OptimizeResult."""
    options = dict(options or {})
    model = mll.model
    params = _named_raw_parameters(mll)
    bnds = bounds if bounds is not None else _bounds_for(mll, params)
    sizes = [p.numel() for _, p in params]
    fast = _FastClosure.build(mll, params) if options.get("fast_closure", True) else None

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
        if fast is not None:
            r = fast(np.asarray(x, dtype=np.float64))
            if r is not None:
                return r
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
            res = None
            if options.get("method") == "device-lbfgs":
                res = fit_gpytorch_mll_device(mll, options=options)
                if res is not None and np.isfinite(res.fun) and options.get("polish", True):
                    # the device run does the bulk of the descent (Armijo best-of-T search); scipy's L-BFGS-B, started at its
                    # answer, applies the upstream stopping rule and usually needs a few dozen evaluations
                    res = fit_gpytorch_mll_scipy(mll, options=options)
            if res is None:
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
