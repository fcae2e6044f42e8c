"""BoTorch-shaped exact GP models whose arithmetic runs in libhdbo_b200.so.

`SingleTaskGP` mirrors botorch.models.SingleTaskGP (constructor, `.posterior`, `.train()/.eval()`,
`.parameters()`, `.state_dict()` with GPyTorch raw-parameter names, `.train_inputs`, `.train_targets`,
`.likelihood`, `.num_outputs`, `.batch_shape`) -- the surface the poli-baselines solvers selected at
/root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-194 use.  `model(train_x)` /
`ExactMarginalLogLikelihood` follow the GPyTorch call pattern of
/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:56,70-84.

The model keeps its training data and all factorisation state on the GPU (`DeviceGP`).  Inputs may live on
the host: they are copied to the device, results are returned on the caller's device / dtype.  There is no CPU
implementation: without CUDA every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr
from .engine import DeviceGP, NotPSDError, _rup, _stream
from .gp_modules import (ConstantMean, GammaPrior, GaussianLikelihood, GreaterThan, Kernel, LogNormalPrior,
                         MaternKernel, RBFKernel, ScaleKernel, ZeroMean)
from .posteriors import GPPosterior, TrainOutput

SQRT2, SQRT3 = math.sqrt(2.0), math.sqrt(3.0)


def compute_device() -> torch.device:
    """cuda unless HDBO_TORCH_DEVICE says otherwise (reference: utils/constants.py:6-12); never a CPU fallback."""
    name = os.environ.get("HDBO_TORCH_DEVICE", "cuda")
    dev = torch.device(name if name.startswith("cuda") else "cuda")
    if not torch.cuda.is_available():
        raise _lib.HdboError("the GP hot path needs a CUDA device (B200); there is no CPU fallback")
    return dev


# ------------------------------------------------------------------------------------------------------
# transforms (botorch.models.transforms)
# ------------------------------------------------------------------------------------------------------
class Standardize(nn.Module):
    """Outcome transform: zero mean / unit (unbiased) std per outcome, learned from the training targets."""

    def __init__(self, m: int = 1, min_stdv: float = 1e-8, **kwargs):
        super().__init__()
        if m != 1:
            raise NotImplementedError("single-output models only")
        self.register_buffer("means", torch.zeros(1, 1, dtype=torch.float64))
        self.register_buffer("stdvs", torch.ones(1, 1, dtype=torch.float64))
        self.register_buffer("_stdvs_sq", torch.ones(1, 1, dtype=torch.float64))      # BoTorch's buffer names: checkpoints round-trip
        self.register_buffer("_is_trained", torch.tensor(False))
        self._min_stdv = min_stdv

    def forward(self, Y):
        Y = Y.to(torch.float64)
        if self.training or not bool(self._is_trained):
            stdvs = Y.std(dim=-2, keepdim=True) if Y.shape[-2] > 1 else torch.ones_like(Y[..., :1, :])
            stdvs = stdvs.where(stdvs >= self._min_stdv, torch.ones_like(stdvs))
            self.means = Y.mean(dim=-2, keepdim=True).to(self.means.device)
            self.stdvs = stdvs.to(self.stdvs.device)
            self._stdvs_sq = self.stdvs.pow(2)
            self._is_trained = torch.tensor(True, device=self.means.device)
        return (Y - self.means.to(Y.device)) / self.stdvs.to(Y.device)

    def untransform_moments(self, mean, var):
        s = self.stdvs.reshape(()).to(mean.device)
        return mean * s + self.means.reshape(()).to(mean.device), var * s * s


class Normalize(nn.Module):
    """Input transform: affine map of `bounds` (2 x d) -- or the training range -- onto the unit cube."""

    def __init__(self, d: int, bounds: Optional[torch.Tensor] = None, **kwargs):
        super().__init__()
        self.d = d
        if bounds is not None:
            b = torch.as_tensor(bounds, dtype=torch.float64)
            self.register_buffer("offset", b[0].clone())
            self.register_buffer("coefficient", (b[1] - b[0]).clone())
            self._learn = False
        else:
            self.register_buffer("offset", torch.zeros(d, dtype=torch.float64))
            self.register_buffer("coefficient", torch.ones(d, dtype=torch.float64))
            self._learn = True

    def fit(self, X):
        if self._learn:
            X = X.to(torch.float64)
            lo, hi = X.min(dim=-2).values, X.max(dim=-2).values
            self.offset = lo
            self.coefficient = torch.where(hi > lo, hi - lo, torch.ones_like(lo))

    def forward(self, X):
        return (X - self.offset.to(X.device)) / self.coefficient.to(X.device)


_DEFAULT = object()


# ------------------------------------------------------------------------------------------------------
# autograd bridge: posterior(X) with gradients w.r.t. X  (A4 forward, A6 backward)
# ------------------------------------------------------------------------------------------------------
class _PosteriorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, model, q, observation_noise):
        gp: DeviceGP = model._device_gp()
        m = Z.shape[0]
        rows = _rup(m, 128)
        need = gp.lib.hdbo_posterior_workspace_bytes(C.byref(gp.desc), rows, 1)
        ws = torch.empty(need, dtype=torch.uint8, device=gp.device)
        f64 = dict(dtype=torch.float64, device=gp.device)
        mean = torch.empty(gp.batch, m, **f64)
        var = torch.empty(gp.batch, m, **f64)
        Zc = Z.detach().contiguous()
        check(gp.lib.hdbo_posterior_fwd(C.byref(gp.desc), ptr(Zc), Zc.stride(0), m, int(observation_noise), ptr(mean),
                                        ptr(var), m, ptr(ws), ws.numel(), _stream()), "hdbo_posterior_fwd")
        ctx.gp, ctx.ws, ctx.m, ctx.q, ctx.d = gp, ws, m, q, Z.shape[1]
        ctx.fit_version = model._fit_version
        ctx.model = model
        if q == 1:
            return mean, var
        Sigma = torch.empty(gp.batch, m // q, q, q, **f64)
        check(gp.lib.hdbo_posterior_joint_cov(C.byref(gp.desc), m, q, int(observation_noise), ptr(Sigma),
                                              Sigma.stride(0), ptr(ws), ws.numel(), _stream()), "hdbo_posterior_joint_cov")
        return mean, Sigma

    @staticmethod
    def backward(ctx, gmean, gsecond):
        gp: DeviceGP = ctx.gp
        if ctx.model._fit_version != ctx.fit_version:
            raise RuntimeError("the model was refitted between posterior() and backward()")
        m, q, d = ctx.m, ctx.q, ctx.d
        f64 = dict(dtype=torch.float64, device=gp.device)
        dZ = torch.empty(gp.batch, m, d, **f64)
        gmean_c = None if gmean is None else gmean.contiguous()
        gvar = gS = None
        if gsecond is not None:
            if q == 1:
                gvar = gsecond.contiguous()
            else:
                gS = gsecond.contiguous()
        if gmean_c is None and gvar is None and gS is None:
            return torch.zeros(m, d, **f64), None, None, None
        check(gp.lib.hdbo_posterior_bwd(C.byref(gp.desc), m, q, ptr(gmean_c), ptr(gvar), ptr(gS), m,
                                        0 if gS is None else gS.stride(0), ptr(dZ), d, m * d, ptr(ctx.ws),
                                        ctx.ws.numel(), _stream()), "hdbo_posterior_bwd")
        return dZ.sum(0) if gp.batch > 1 else dZ[0], None, None, None


# ------------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------------
class ExactGPBase(nn.Module):
    """Shared machinery of the exact-GP models: device state, hyper-parameter upload, posterior."""

    num_outputs = 1
    _is_fully_bayesian = False

    def _init_data(self, train_X, train_Y, outcome_transform, input_transform):
        if train_X.dim() != 2:
            raise NotImplementedError("train_X must be n x d (batched training data is not on the hdbo hot path)")
        if train_Y.dim() == 1:
            train_Y = train_Y.unsqueeze(-1)
        if train_Y.shape[-1] != 1:
            raise NotImplementedError("single-output models only")
        self._io_device = train_X.device
        self._io_dtype = train_X.dtype if train_X.dtype.is_floating_point else torch.float64
        self._dev = compute_device()
        self.input_transform = input_transform
        self.outcome_transform = outcome_transform
        X = train_X.detach().to(self._dev, torch.float64)
        Y = train_Y.detach().to(self._dev, torch.float64)
        self._train_X_raw = X
        if input_transform is not None:
            input_transform.to(self._dev)
            if hasattr(input_transform, "fit"):
                input_transform.fit(X)
            X = input_transform(X)
        if outcome_transform is not None:
            outcome_transform.to(self._dev)
            outcome_transform.train()
            Y = outcome_transform(Y)
            outcome_transform.eval()
        self.train_inputs = (X.contiguous(),)
        self.train_targets = Y.squeeze(-1).contiguous()
        self._gp: Optional[DeviceGP] = None
        self._fit_version = 0
        self._fitted_key = None

    # -- hyper-parameters as constrained tensors --------------------------------------------------------
    @property
    def _base_kernel(self) -> Kernel:
        cm = self.covar_module
        return cm.base_kernel if isinstance(cm, ScaleKernel) else cm

    def hyperparameters(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """(lengthscale [.., d], outputscale [..], noise [..], mean [..]) as differentiable tensors on the compute device."""
        d = self.train_inputs[0].shape[-1]
        ls = self._base_kernel.lengthscale
        ls = ls.reshape(*ls.shape[:-2], ls.shape[-1]).expand(*ls.shape[:-2], d)
        if isinstance(self.covar_module, ScaleKernel):
            os_ = self.covar_module.outputscale
        else:
            os_ = torch.ones(ls.shape[:-1], dtype=torch.float64, device=ls.device)
        noise = self.likelihood.noise.squeeze(-1)
        mean = self.mean_module.constant.to(ls.device)
        return ls, os_, noise, mean

    def _param_key(self):
        return tuple((id(p), p._version) for p in self.parameters())

    def _device_gp(self, need_r2: bool = False) -> DeviceGP:
        """The fitted device state for the current hyper-parameters (refits when any raw parameter changed)."""
        B = int(math.prod(self.batch_shape)) if len(self.batch_shape) else 1
        if self._gp is None or (need_r2 and not self._gp.keep_r2):
            self._gp = DeviceGP(self.train_inputs[0], self.train_targets, self._base_kernel.kind, batch=B,
                                keep_r2=need_r2 or (self._gp.keep_r2 if self._gp is not None else False))
            self._fitted_key = None
        key = self._param_key()
        if self._fitted_key != key or not self._gp.fitted:
            with torch.no_grad():
                ls, os_, noise, mean = self.hyperparameters()
                self._gp.set_hyper(ls.reshape(-1, ls.shape[-1]), os_.reshape(-1), noise.reshape(-1), mean.reshape(-1))
            self._gp.fit(need_r2=need_r2)
            self._fitted_key = key
            self._fit_version += 1
        return self._gp

    def train(self, mode: bool = True):
        return super().train(mode)

    # -- BoTorch surface --------------------------------------------------------------------------------
    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size([])

    @property
    def train_X(self):
        return self._train_X_raw

    def _prep_X(self, X):
        X = torch.as_tensor(X)
        out_device, out_dtype = X.device, (X.dtype if X.dtype.is_floating_point else torch.float64)
        Xd = X.to(self._dev, torch.float64)
        if self.input_transform is not None:
            Xd = self.input_transform(Xd)
        return Xd, out_device, out_dtype

    def posterior(self, X, observation_noise: bool = False, posterior_transform=None, **kwargs) -> GPPosterior:
        """X: [b, q, d] (or [q, d]) -> posterior with .mean / .variance of shape [b, q, 1]."""
        Xd, out_device, out_dtype = self._prep_X(X)
        post = self._posterior_from_transformed(Xd, bool(observation_noise), out_device, out_dtype)
        return posterior_transform(post) if posterior_transform is not None else post

    def _posterior_from_transformed(self, Xd, observation_noise: bool, out_device, out_dtype) -> GPPosterior:
        """Xd: already on the compute device and through the input transform."""
        squeeze_b = Xd.dim() == 2
        if squeeze_b:
            Xd = Xd.unsqueeze(0)
        if Xd.dim() != 3:
            raise ValueError("X must be b x q x d")
        b, q, d = Xd.shape
        gp = self._device_gp()
        flat = Xd.reshape(b * q, d)
        Sigma = None
        if torch.is_grad_enabled() and flat.requires_grad:
            if q > 8:
                raise NotImplementedError("differentiable joint posteriors support q <= 8")
            mean, second = _PosteriorFn.apply(flat, self, q, bool(observation_noise))
            if q == 1:
                var = second
            else:
                Sigma = second
                var = torch.diagonal(Sigma, dim1=-2, dim2=-1).reshape(gp.batch, b * q)
        else:
            mean, var, _ = gp.posterior_acq(flat, _lib.ACQ_NONE, 0.0, observation_noise=bool(observation_noise))
        return GPPosterior(self, flat, mean, var, Sigma, b, q, squeeze_b, out_device, out_dtype, bool(observation_noise))

    def forward(self, X):
        if self.training:
            return TrainOutput(self)
        return self.posterior(X)

    def __call__(self, X=None, *args, **kwargs):
        return self.forward(X)

    def condition_on_observations(self, X, Y, **kwargs):
        """A new model on the augmented data with the SAME hyper-parameters (botorch `Model.condition_on_observations`).
        BoTorch semantics: the new observations go through the EXISTING, already-trained transforms (eval mode: Standardize keeps
        its means / stdvs, a learned-bounds Normalize keeps its bounds); nothing is re-learned, so predictions of the conditioned
        model un-transform with the same statistics its targets were standardised with.  The transforms are deep-copied: the
        original model is never mutated."""
        import copy
        d = self._train_X_raw.shape[-1]
        Xraw_new = torch.as_tensor(X).to(self._dev, torch.float64).reshape(-1, d)
        Ynew = torch.as_tensor(Y).to(self._dev, torch.float64).reshape(-1, 1)
        Xt_new = self.input_transform(Xraw_new) if self.input_transform is not None else Xraw_new
        if self.outcome_transform is not None:
            m, s = self.outcome_transform.means.reshape(()), self.outcome_transform.stdvs.reshape(())
            Ynew = (Ynew - m) / s
        kw = self._ctor_kwargs()
        kw["outcome_transform"] = None
        kw["input_transform"] = None
        new = type(self)(torch.cat([self.train_inputs[0], Xt_new]), torch.cat([self.train_targets.unsqueeze(-1), Ynew]), **kw)
        new.outcome_transform = copy.deepcopy(self.outcome_transform)
        new.input_transform = copy.deepcopy(self.input_transform)
        new._train_X_raw = torch.cat([self._train_X_raw, Xraw_new])
        new._io_device, new._io_dtype = self._io_device, self._io_dtype
        hyper = {k: v for k, v in self.state_dict().items()
                 if not (k.startswith("outcome_transform.") or k.startswith("input_transform."))}
        new.load_state_dict(hyper, strict=False)
        return new

    def _ctor_kwargs(self):
        return {}


class SingleTaskGP(ExactGPBase):
    """Drop-in for botorch.models.SingleTaskGP (single outcome, inferred homoskedastic noise).

    Defaults follow BoTorch >= 0.12 (the reference pins no version, pyproject.toml:10): ARD RBF kernel without
    ScaleKernel, LogNormal(sqrt2 + log(d)/2, sqrt3) lengthscale prior, LogNormal(-4, 1) noise prior, noise >= 1e-4,
    `Standardize` outcome transform.  `legacy_defaults=True` gives the <= 0.11 model (Matern-5/2 ARD + ScaleKernel with
    Gamma(3,6) / Gamma(2,0.15) / Gamma(1.1,0.05) priors, softplus constraints)."""

    def __init__(self, train_X, train_Y, train_Yvar=None, likelihood=None, covar_module=None, mean_module=None,
                 outcome_transform=_DEFAULT, input_transform=None, legacy_defaults: bool = False):
        super().__init__()
        if train_Yvar is not None:
            raise NotImplementedError("fixed-noise models are not on the hdbo_benchmark hot path")
        if outcome_transform is _DEFAULT:
            outcome_transform = None if legacy_defaults else Standardize(m=1)
        self._legacy = legacy_defaults
        self._init_data(train_X, train_Y, outcome_transform, input_transform)
        d = self.train_inputs[0].shape[-1]
        if likelihood is None:
            if legacy_defaults:
                prior = GammaPrior(1.1, 0.05)
                likelihood = GaussianLikelihood(noise_prior=prior, noise_constraint=GreaterThan(
                    1e-4, transform="softplus", initial_value=float(prior.mode)))
            else:
                prior = LogNormalPrior(-4.0, 1.0)
                likelihood = GaussianLikelihood(noise_prior=prior, noise_constraint=GreaterThan(
                    1e-4, transform=None, initial_value=float(prior.mode)))
        if covar_module is None:
            if legacy_defaults:
                covar_module = ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=d, lengthscale_prior=GammaPrior(3.0, 6.0)),
                                           outputscale_prior=GammaPrior(2.0, 0.15))
            else:
                prior = LogNormalPrior(SQRT2 + 0.5 * math.log(d), SQRT3)
                covar_module = RBFKernel(ard_num_dims=d, lengthscale_prior=prior, lengthscale_constraint=GreaterThan(
                    2.5e-2, transform=None, initial_value=float(prior.mode)))
        self.likelihood = likelihood
        self.covar_module = covar_module
        self.mean_module = mean_module if mean_module is not None else ConstantMean()
        self.to(self._dev)

    def _ctor_kwargs(self):
        import copy
        return dict(likelihood=copy.deepcopy(self.likelihood), covar_module=copy.deepcopy(self.covar_module),
                    mean_module=copy.deepcopy(self.mean_module),
                    outcome_transform=copy.deepcopy(self.outcome_transform) if self.outcome_transform is not None else None,
                    input_transform=copy.deepcopy(self.input_transform) if self.input_transform is not None else None,
                    legacy_defaults=self._legacy)


# ------------------------------------------------------------------------------------------------------
# exact marginal log likelihood with analytic gradient (A3)
# ------------------------------------------------------------------------------------------------------
class _MLLDataTermFn(torch.autograd.Function):
    """(lengthscale [B,d], outputscale [B], noise [B], mean [B]) -> n*mll data term [B], gradient from hdbo_mll_grad."""

    @staticmethod
    def forward(ctx, ls, os_, noise, mean, model):
        need_grad = any(ctx.needs_input_grad[:4])
        B = ls.shape[0]
        gp = model._gp
        if gp is None or gp.batch != B or (need_grad and not gp.keep_r2):
            gp = model._gp = DeviceGP(model.train_inputs[0], model.train_targets, model._base_kernel.kind, batch=B,
                                      keep_r2=need_grad or (gp.keep_r2 if gp is not None else False))
        gp.set_hyper(ls, os_, noise, mean)
        gp.fit(need_r2=need_grad)
        model._fitted_key = model._param_key()
        model._fit_version += 1
        value = gp.mll_data_term.clone()
        if need_grad:
            ctx.grad = mll_gradient(gp)
        return value

    @staticmethod
    def backward(ctx, g):
        gr = ctx.grad  # [B, d+3]
        d = gr.shape[1] - 3
        g = g.reshape(-1, 1)
        return g * gr[:, :d], g[:, 0] * gr[:, d], g[:, 0] * gr[:, d + 1], g[:, 0] * gr[:, d + 2], None


def mll_gradient(gp: DeviceGP) -> torch.Tensor:
    """d(n*mll data term)/d(lengthscale, outputscale, noise, mean) for a fitted DeviceGP with R2; [B, d+3]."""
    f64 = dict(dtype=torch.float64, device=gp.device)
    B, npad = gp.batch, gp.n_pad
    dpt = _rup(gp.d, 128)
    scratch = getattr(gp, "_grad_scratch", None)
    if scratch is None:   # persistent scratch: an L-BFGS fit calls this 20-100 times on the same DeviceGP
        scratch = gp._grad_scratch = (
            torch.empty(B, npad, npad, **f64), torch.empty(B, dpt, npad, **f64),
            torch.empty(gp.lib.hdbo_mll_grad_part_bytes(C.byref(gp.desc)), dtype=torch.uint8, device=gp.device))
    G, XsT, part = scratch
    grad = torch.empty(B, gp.d + 3, **f64)
    # K^-1: already left in gp.Kinv by the fit (need_r2), else formed now over the Cholesky factor buffer (L itself is no longer
    # needed once L^-1 exists)
    kinv = gp.Kinv if getattr(gp, "_kinv_valid", False) else gp.L
    check(gp.lib.hdbo_mll_grad(C.byref(gp.desc), ptr(kinv), ptr(G), ptr(XsT), ptr(part), ptr(grad), _stream()),
          "hdbo_mll_grad")
    return grad


class MarginalLogLikelihood(nn.Module):
    pass


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    """gpytorch.mlls.ExactMarginalLogLikelihood: mll(model(train_x), train_y) -> scalar (divided by n), with
    log-priors added; `.backward()` fills the raw-parameter gradients (analytic kernel, autograd only for the
    O(d) raw->constrained transforms and prior densities)."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model

    def forward(self, output, target=None, *params, **kwargs):
        model = self.model
        if isinstance(output, TrainOutput):
            return self._train_mll()
        if isinstance(output, GPPosterior):
            return output.joint_log_prob_per_datum(target, self._log_prior())
        raise TypeError("ExactMarginalLogLikelihood expects model(train_x) or an eval-mode model(test_x)")

    def _log_prior(self):
        model = self.model
        total = torch.zeros((), dtype=torch.float64, device=model._dev)
        seen = set()
        for mod, prefix in ((model.covar_module, "covar_module."), (model.likelihood, "likelihood.")):
            for name, prior, value in mod.named_priors_with_values(prefix):
                if id(prior) in seen:
                    continue
                seen.add(id(prior))
                total = total + prior.log_prob(value).sum()
        mp = getattr(model.mean_module, "constant_prior", None)
        if mp is not None:
            total = total + mp.log_prob(model.mean_module.constant).sum()
        return total

    def _train_mll(self):
        model = self.model
        n = model.train_targets.shape[-1]
        ls, os_, noise, mean = model.hyperparameters()
        d = ls.shape[-1]
        data = _MLLDataTermFn.apply(ls.reshape(-1, d), os_.reshape(-1), noise.reshape(-1),
                                    mean.reshape(-1).expand(ls.reshape(-1, d).shape[0]), model)
        res = data.reshape(ls.shape[:-1]) if ls.dim() > 1 else data.reshape(())
        return (res + self._log_prior()) / n


def legacy_single_task_gp(train_X, train_Y, **kwargs) -> SingleTaskGP:
    return SingleTaskGP(train_X, train_Y, legacy_defaults=True, **kwargs)


# ------------------------------------------------------------------------------------------------------
# fully Bayesian (SAAS) model: S hyper-parameter samples = one batched device GP  (A10)
# ------------------------------------------------------------------------------------------------------
class SaasFullyBayesianSingleTaskGP(ExactGPBase):
    """Drop-in for botorch.models.fully_bayesian.SaasFullyBayesianSingleTaskGP once MCMC samples exist
    (the `saas_bo` solver, /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:183-194).

    `load_mcmc_samples({"lengthscale": [S,d], "outputscale": [S], "mean": [S], "noise": [S]})` installs a batch of S
    Matern-5/2 ARD GPs that share the training data; kernel build, Cholesky, inverse and posterior run batched on the
    device (grid.z = S).  `posterior(X)` returns a Gaussian-mixture posterior with the MCMC dimension at -3
    (mean: b x S x q x 1, plus `.mixture_mean` / `.mixture_variance`).  The NUTS sampler itself (Pyro control flow) is
    out of scope; `sample_saas_prior` draws hyper-parameters from the SAAS prior for synthetic workloads."""

    _is_fully_bayesian = True

    def __init__(self, train_X, train_Y, train_Yvar=None, outcome_transform=None, input_transform=None):
        super().__init__()
        if train_Yvar is not None:
            raise NotImplementedError("fixed-noise SAAS models are not on the hdbo_benchmark hot path")
        self._init_data(train_X, train_Y, outcome_transform, input_transform)
        self.mean_module = None
        self.covar_module = None
        self.likelihood = None
        self._num_samples = 0

    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size([self._num_samples]) if self._num_samples else torch.Size([])

    @property
    def num_mcmc_samples(self) -> int:
        return self._num_samples

    def load_mcmc_samples(self, mcmc_samples) -> None:
        ls = torch.as_tensor(mcmc_samples["lengthscale"], dtype=torch.float64)
        S, d = ls.shape[0], self.train_inputs[0].shape[-1]
        ls = ls.reshape(S, -1)
        if ls.shape[1] != d:
            raise ValueError(f"lengthscale samples must be S x {d}")
        bs = torch.Size([S])
        self.covar_module = ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=d, batch_shape=bs))
        self.mean_module = ConstantMean(batch_shape=bs)
        self.likelihood = GaussianLikelihood(batch_shape=bs, noise_constraint=GreaterThan(1e-6))
        self._num_samples = S
        self.to(self._dev)
        self.covar_module.base_kernel.lengthscale = ls.reshape(S, 1, d)
        self.covar_module.outputscale = torch.as_tensor(mcmc_samples["outputscale"], dtype=torch.float64).reshape(S)
        self.mean_module.constant = torch.as_tensor(mcmc_samples["mean"], dtype=torch.float64).reshape(S)
        noise = mcmc_samples.get("noise")
        noise = torch.full((S,), 1e-6, dtype=torch.float64) if noise is None else torch.as_tensor(noise, dtype=torch.float64)
        self.likelihood.noise = noise.reshape(S, 1).clamp_min(1e-6 * (1 + 1e-9) + 1e-12)
        self._gp = None
        self.eval()

    @property
    def median_lengthscale(self) -> torch.Tensor:
        return self.covar_module.base_kernel.lengthscale.detach().reshape(self._num_samples, -1).median(0).values

    def forward(self, X):
        if self._num_samples == 0:
            raise RuntimeError("load_mcmc_samples() must be called before evaluating the model")
        return super().forward(X)

    def posterior(self, X, observation_noise: bool = False, posterior_transform=None, **kwargs):
        if self._num_samples == 0:
            raise RuntimeError("load_mcmc_samples() must be called before evaluating the model")
        return super().posterior(X, observation_noise=observation_noise, posterior_transform=posterior_transform, **kwargs)


def sample_saas_prior(S: int, d: int, seed: int = 4):
    """Hyper-parameter draws from the SAAS prior: tau ~ HalfCauchy(0.1),
    1/l_k^2 ~ tau * HalfCauchy(1) clamped to [1e-4, 1e2], outputscale ~ Gamma(2, 0.15), noise ~ Gamma(0.9, 10) + 1e-4,
    mean ~ N(0, 1).  Returns the dict `load_mcmc_samples` expects (CPU float64 tensors)."""
    g = torch.Generator().manual_seed(seed)
    def half_cauchy(shape, scale):
        u = torch.rand(shape, generator=g, dtype=torch.float64)
        return scale * torch.tan(0.5 * math.pi * u)
    tau = half_cauchy((S, 1), 0.1)
    inv_l2 = (tau * half_cauchy((S, d), 1.0)).clamp(1e-4, 1e2)
    def gamma(shape, conc, rate):
        return torch._standard_gamma(torch.full(shape, conc, dtype=torch.float64), generator=g) / rate
    return {
        "lengthscale": inv_l2.rsqrt(),
        "outputscale": gamma((S,), 2.0, 0.15).clamp_min(1e-2),
        "noise": gamma((S,), 0.9, 10.0) + 1e-4,
        "mean": torch.randn(S, generator=g, dtype=torch.float64),
    }
