"""Posterior objects returned by `model.posterior(X)` / eval-mode `model(X)` (botorch GPyTorchPosterior surface).

Holds the device results of libhdbo_b200 (mean / variance / optional joint q x q covariance per t-batch) and exposes
`.mean`, `.variance` ([b, q, 1]), `.covariance_matrix`, `.rsample`, `.rsample_from_base_samples`.  Sampling works on
the tiny q x q factors only (torch ops on the device; the n-sized work is already done in CUDA).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib


class TrainOutput:
    """What `model(train_x)` returns in training mode: a handle the MLL consumes (GPyTorch returns a lazy prior MVN)."""

    def __init__(self, model):
        self.model = model


class GPPosterior:
    def __init__(self, model, flatZ, mean, var, Sigma, b, q, squeeze_b, out_device, out_dtype, observation_noise):
        self._model = model
        self._Z = flatZ
        self._B = mean.shape[0]
        self._b, self._q, self._squeeze_b = b, q, squeeze_b
        self._out_device, self._out_dtype = out_device, out_dtype
        self._observation_noise = observation_noise
        ot = getattr(model, "outcome_transform", None)
        self._scale = 1.0
        self._shift = 0.0
        if ot is not None:
            self._scale = ot.stdvs.reshape(()).to(mean.device)
            self._shift = ot.means.reshape(()).to(mean.device)
        self._mean_std = mean  # [B, b*q] in the model's (standardised) units
        self._var_std = var
        self._Sigma_std = Sigma  # [B, b, q, q] or None

    # ---- shapes -------------------------------------------------------------------------------------
    def _shape(self, t):
        """[B, b*q] -> [b, q, 1] (single model) or [b, B, q, 1] (fully Bayesian: MCMC dim is -3)."""
        t = t.reshape(self._B, self._b, self._q, 1)
        if getattr(self._model, "_is_fully_bayesian", False):
            t = t.permute(1, 0, 2, 3)
        else:
            t = t[0]
        if self._squeeze_b:
            t = t.squeeze(0)
        return t.to(self._out_device, self._out_dtype)

    @property
    def device(self):
        return self._out_device

    @property
    def dtype(self):
        return self._out_dtype

    @property
    def mean(self):
        return self._shape(self._mean_std * self._scale + self._shift)

    @property
    def variance(self):
        return self._shape(self._var_std * (self._scale * self._scale))

    @property
    def stddev(self):
        return self.variance.clamp_min(0).sqrt()

    # raw device views used by the fused acquisition paths (standardised units, [B, b*q])
    @property
    def device_mean(self):
        return self._mean_std

    @property
    def device_variance(self):
        return self._var_std

    def _joint_cov_std(self):
        if self._Sigma_std is None:
            if self._q == 1:
                self._Sigma_std = self._var_std.reshape(self._B, self._b, 1, 1)
            else:
                with torch.enable_grad():
                    Zg = self._Z.detach().requires_grad_(True)
                    from .models import _PosteriorFn
                    _, S = _PosteriorFn.apply(Zg, self._model, self._q, self._observation_noise)
                self._Sigma_std = S.detach()
        return self._Sigma_std

    @property
    def covariance_matrix(self):
        S = self._joint_cov_std() * (self._scale * self._scale)
        if getattr(self._model, "_is_fully_bayesian", False):
            S = S.permute(1, 0, 2, 3)
        else:
            S = S[0]
        if self._squeeze_b:
            S = S.squeeze(0)
        return S.to(self._out_device, self._out_dtype)

    @property
    def mvn(self):
        return self

    def with_observation_noise(self, noise=None):
        """The predictive distribution of y = f + eps: what `likelihood(model(x))` returns in GPyTorch.  Re-issues the posterior with
        the model's own noise added to the variances / the diagonal of the joint covariance."""
        if self._observation_noise:
            return self
        X = self._Z.reshape(self._b, self._q, -1)
        return self._model._posterior_from_transformed(X[0] if self._squeeze_b else X, True, self._out_device, self._out_dtype)

    # ---- sampling -----------------------------------------------------------------------------------
    def rsample_from_base_samples(self, sample_shape, base_samples):
        """base_samples: sample_shape x [b] x q x 1 standard normal draws -> samples sample_shape x b x q x 1.
        The q x q Cholesky factors (jitter 1e-8) and mean + L z run in `hdbo_mvn_sample` (q <= 8)."""
        from ._lib import check, load, ptr
        from .engine import _stream
        sample_shape = torch.Size(sample_shape)
        q, B, b = self._q, self._B, self._b
        Sig = self._joint_cov_std().reshape(B * b, q, q).contiguous()
        mean = self._mean_std.reshape(B * b, q).contiguous()
        dev = mean.device
        z = torch.as_tensor(base_samples).to(dev, torch.float64)
        if z.dim() and z.shape[-1] == 1 and z.dim() > len(sample_shape) + 1:
            z = z.squeeze(-1)
        S = int(math.prod(sample_shape)) if len(sample_shape) else 1
        lead = (B, b) if B > 1 else (b,)
        z = z.reshape(*sample_shape, *z.shape[len(sample_shape):])
        # broadcast the per-group dimensions that the caller left out or gave as 1
        while z.dim() < len(sample_shape) + len(lead) + 1:
            z = z.unsqueeze(len(sample_shape))
        shared = all(int(x) == 1 for x in z.shape[len(sample_shape):-1])
        if shared:
            zz = z.reshape(S, q).contiguous()
        else:
            zz = z.expand(*sample_shape, *lead, q).reshape(S, B * b, q).contiguous()
        out = torch.empty(S, B * b, q, dtype=torch.float64, device=dev)
        check(load().hdbo_mvn_sample(ptr(mean), ptr(Sig), ptr(zz), B * b, q, S, 0 if shared else 1, 1e-8, ptr(out), 0, _stream()),
              "hdbo_mvn_sample")
        samples = out.reshape(*sample_shape, *lead, q) * self._scale + self._shift
        return samples.unsqueeze(-1).to(self._out_device, self._out_dtype)

    def rsample(self, sample_shape: Optional[torch.Size] = None):
        sample_shape = torch.Size(sample_shape) if sample_shape is not None else torch.Size([1])
        dev = self._mean_std.device
        z = torch.randn(*sample_shape, self._b, self._q, 1, dtype=torch.float64, device=dev)
        return self.rsample_from_base_samples(sample_shape, z)

    sample = rsample

    # ---- mixture moments (fully Bayesian models) ----------------------------------------------------
    @property
    def mixture_mean(self):
        m = self._mean_std * self._scale + self._shift  # [B, b*q]
        out = m.mean(0).reshape(self._b, self._q, 1)
        return (out.squeeze(0) if self._squeeze_b else out).to(self._out_device, self._out_dtype)

    @property
    def mixture_variance(self):
        m = self._mean_std * self._scale + self._shift
        v = self._var_std * (self._scale * self._scale)
        mm = m.mean(0)
        out = ((v + m * m).mean(0) - mm * mm).reshape(self._b, self._q, 1)
        return (out.squeeze(0) if self._squeeze_b else out).to(self._out_device, self._out_dtype)

    # ---- eval-mode MLL (training_gps.py:83-84) ------------------------------------------------------
    def joint_log_prob_per_datum(self, target, log_prior):
        from .predictive import predictive_log_prob
        return predictive_log_prob(self._model, self._Z, target, log_prior)
