"""BoTorch-shaped acquisition functions evaluated by libhdbo_b200.

`acqf(X)` with X of shape [b, q, d] returns [b]; it is differentiable w.r.t. X through analytic CUDA backward
kernels (no tape replay).  Without gradients the analytic functions take the fused candidate-pool path
(`hdbo_posterior_acq`: cross-kernel GEMM -> variance GEMM -> acquisition epilogue).  The classes mirror
botorch.acquisition.analytic.{ExpectedImprovement, LogExpectedImprovement, UpperConfidenceBound,
ProbabilityOfImprovement, PosteriorMean} and botorch.acquisition.monte_carlo.qExpectedImprovement, which the
poli-baselines solvers selected at /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-194
instantiate with `best_f = y.max()`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import ACQ_EI, ACQ_LOGEI, ACQ_MEAN, ACQ_NONE, ACQ_PI, ACQ_UCB, check, ptr
from .engine import _stream
from .models import _PosteriorFn
from .sampling import SobolQMCNormalSampler


class _AcqElemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, var, kind, param):
        lib = _lib.load()
        mean_c, var_c = mean.contiguous(), var.contiguous()
        acq = torch.empty_like(mean_c)
        dmean = torch.empty_like(mean_c)
        dvar = torch.empty_like(mean_c)
        check(lib.hdbo_acq_elementwise(ptr(mean_c), ptr(var_c), mean_c.numel(), kind, float(param), ptr(acq), ptr(dmean),
                                       ptr(dvar), _stream()), "hdbo_acq_elementwise")
        ctx.save_for_backward(dmean, dvar)
        return acq

    @staticmethod
    def backward(ctx, g):
        dmean, dvar = ctx.saved_tensors
        return g * dmean, g * dvar, None, None


class _QEIFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, Sigma, base, best_f, jitter):
        lib = _lib.load()
        R, q = mean.shape
        mean_c, Sig_c = mean.contiguous(), Sigma.contiguous()
        value = torch.empty(R, dtype=torch.float64, device=mean.device)
        gmean = torch.empty_like(mean_c)
        gSigma = torch.empty_like(Sig_c)
        info = torch.zeros(R, dtype=torch.int32, device=mean.device)
        check(lib.hdbo_qei(ptr(mean_c), ptr(Sig_c), ptr(base), R, q, base.shape[0], float(best_f), float(jitter), ptr(value),
                           ptr(gmean), ptr(gSigma), ptr(info), _stream()), "hdbo_qei")
        ctx.save_for_backward(gmean, gSigma)
        return value

    @staticmethod
    def backward(ctx, g):
        gmean, gSigma = ctx.saved_tensors
        return g.unsqueeze(-1) * gmean, g.reshape(-1, 1, 1) * gSigma, None, None, None


def _logmeanexp(x, dim):
    return torch.logsumexp(x, dim=dim) - math.log(x.shape[dim])


class AcquisitionFunction(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending


class AnalyticAcquisitionFunction(AcquisitionFunction):
    _kind = ACQ_NONE
    _log = False

    def __init__(self, model, posterior_transform=None, maximize: bool = True):
        super().__init__(model)
        if posterior_transform is not None:
            raise NotImplementedError("posterior_transform is not used on the hdbo_benchmark hot path")
        self.maximize = maximize

    # parameter of the device kernel in the model's standardised units, and the map back to outcome units
    def _std_param(self, scale, shift, sign=1.0):
        raise NotImplementedError

    def _finish(self, a, scale, shift):
        return a

    def _scale_shift(self):
        ot = getattr(self.model, "outcome_transform", None)
        if ot is None:
            return 1.0, 0.0
        return float(ot.stdvs.reshape(())), float(ot.means.reshape(()))

    def forward(self, X):
        X = torch.as_tensor(X)
        if X.dim() == 2:
            X = X.unsqueeze(-2)
        if X.dim() != 3 or X.shape[-2] != 1:
            raise ValueError(f"analytic acquisition functions need X of shape b x 1 x d, got {tuple(X.shape)}")
        model = self.model
        scale, shift = self._scale_shift()
        sign = 1.0 if self.maximize else -1.0
        param = self._std_param(scale, shift, sign)
        if self._host_pool_supported(X, sign):
            return self._forward_host_pool(X, param, scale, shift * sign)
        Xd, out_device, out_dtype = model._prep_X(X)
        flat = Xd.reshape(X.shape[0], X.shape[-1])
        gp = model._device_gp()
        if torch.is_grad_enabled() and flat.requires_grad:
            mean, var = _PosteriorFn.apply(flat, model, 1, False)
            a = _AcqElemFn.apply(sign * mean if sign < 0 else mean, var, self._kind, param)
        elif sign > 0:
            _, _, a = gp.posterior_acq(flat, self._kind, param, want_mean=False, want_var=False)
        else:
            mean, var, _ = gp.posterior_acq(flat, ACQ_NONE, 0.0)
            a = _AcqElemFn.apply(-mean, var, self._kind, param)
        a = self._finish(a, scale, shift * sign)  # [B, b]
        a = (_logmeanexp(a, 0) if self._log else a.mean(0)) if a.shape[0] > 1 else a[0]
        return a.to(out_device, out_dtype)


    # -- large no-grad pools that live in HOST memory (the raw-sample screening of optimize_acqf driven from CPU tensors) ----------
    HOST_POOL_MIN = 1 << 15

    def _host_pool_supported(self, X, sign) -> bool:
        from .models import Normalize
        if X.is_cuda or X.dtype != torch.float64 or X.shape[0] < self.HOST_POOL_MIN or sign < 0:
            return False
        if torch.is_grad_enabled() and X.requires_grad:
            return False
        model = self.model
        it = getattr(model, "input_transform", None)
        return not getattr(model, "_is_fully_bayesian", False) and (it is None or isinstance(it, Normalize))

    def _forward_host_pool(self, X, param, scale, shift):
        """acqf(X) for a CPU tensor X [b, 1, d]: the candidates are streamed to the device slab by slab on a copy stream while the
        previous slab is evaluated (`DeviceGP.posterior_acq_host`), the values come back through pinned memory; the result is a CPU
        tensor like the input (BoTorch returns values on the input's device)."""
        model = self.model
        gp = model._device_gp()
        flat = X.detach().reshape(X.shape[0], X.shape[-1])
        if flat.stride(-1) != 1:
            flat = flat.contiguous()
        it = getattr(model, "input_transform", None)
        tf = None
        if it is not None:
            off, coef = it.offset.to(gp.device, torch.float64), it.coefficient.to(gp.device, torch.float64)
            tf = lambda slab: slab.sub_(off).div_(coef)
        a, _, _ = gp.posterior_acq_host(flat, self._kind, param, slab_transform=tf)
        a = self._finish(a.reshape(1, -1), scale, shift)[0]
        out = torch.empty(a.shape[0], dtype=torch.float64, pin_memory=True)
        out.copy_(a, non_blocking=True)
        torch.cuda.current_stream(gp.device).synchronize()
        return out


class ExpectedImprovement(AnalyticAcquisitionFunction):
    _kind = ACQ_EI

    def __init__(self, model, best_f, posterior_transform=None, maximize: bool = True):
        super().__init__(model, posterior_transform, maximize)
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=torch.float64).reshape(()).cpu())

    def _std_param(self, scale, shift, sign=1.0):
        return (sign * float(self.best_f) - sign * shift) / scale

    def _finish(self, a, scale, shift):
        return a * scale


class LogExpectedImprovement(ExpectedImprovement):
    _kind = ACQ_LOGEI
    _log = True

    def _finish(self, a, scale, shift):
        return a + math.log(scale)


class ProbabilityOfImprovement(ExpectedImprovement):
    _kind = ACQ_PI

    def _finish(self, a, scale, shift):
        return a


class UpperConfidenceBound(AnalyticAcquisitionFunction):
    _kind = ACQ_UCB

    def __init__(self, model, beta, posterior_transform=None, maximize: bool = True):
        super().__init__(model, posterior_transform, maximize)
        self.register_buffer("beta", torch.as_tensor(beta, dtype=torch.float64).reshape(()).cpu())

    def _std_param(self, scale, shift, sign=1.0):
        return float(self.beta)

    def _finish(self, a, scale, shift):
        return a * scale + shift


class PosteriorMean(AnalyticAcquisitionFunction):
    _kind = ACQ_MEAN

    def _std_param(self, scale, shift, sign=1.0):
        return 0.0

    def _finish(self, a, scale, shift):
        return a * scale + shift


class qExpectedImprovement(AcquisitionFunction):
    """mean_s max_j relu(sample_sj - best_f) over fixed Sobol-normal base samples (q <= 8)."""

    def __init__(self, model, best_f, sampler: Optional[SobolQMCNormalSampler] = None, objective=None,
                 posterior_transform=None, X_pending=None, jitter: float = 1e-8, **kwargs):
        super().__init__(model)
        if objective is not None or posterior_transform is not None:
            raise NotImplementedError("objectives / posterior transforms are not on the hdbo_benchmark hot path")
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=torch.float64).reshape(()).cpu())
        self.sampler = sampler if sampler is not None else SobolQMCNormalSampler(sample_shape=torch.Size([512]))
        self.jitter = jitter
        self.X_pending = X_pending

    def forward(self, X):
        X = torch.as_tensor(X)
        if X.dim() == 2:
            X = X.unsqueeze(0)
        model = self.model
        if self.X_pending is not None:
            Xp = torch.as_tensor(self.X_pending).to(X)
            X = torch.cat([X, Xp.expand(X.shape[0], *Xp.shape[-2:])], dim=-2)
        b, q, d = X.shape
        if q > 8:
            raise NotImplementedError("qExpectedImprovement supports q <= 8")
        ot = getattr(model, "outcome_transform", None)
        scale = float(ot.stdvs.reshape(())) if ot is not None else 1.0
        shift = float(ot.means.reshape(())) if ot is not None else 0.0
        Xd, out_device, out_dtype = model._prep_X(X)
        flat = Xd.reshape(b * q, d)
        gp = model._device_gp()
        base = self.sampler.base_samples(q, gp.device)  # [S, q]
        mean, second = _PosteriorFn.apply(flat, model, q, False)
        B = mean.shape[0]
        if q == 1:
            Sigma = second.reshape(B * b, 1, 1)
        else:
            Sigma = second.reshape(B * b, q, q)
        val = _QEIFn.apply(mean.reshape(B * b, q), Sigma, base, (float(self.best_f) - shift) / scale, self.jitter)
        val = val.reshape(B, b).mean(0) * scale
        return val.to(out_device, out_dtype)
