"""Posterior objects returned by `model.posterior(X)` / eval-mode `model(X)` (botorch GPyTorchPosterior surface).

Holds the device results of libhdbo_b200 (mean / variance / optional joint q x q covariance per t-batch) and exposes
`.mean`, `.variance` ([b, q, 1]), `.covariance_matrix`, `.rsample`, `.rsample_from_base_samples`.  Sampling works on
the tiny q x q factors only (torch ops on the device; the n-sized work is already done in CUDA).
"""
from __future__ import annotations

import ctypes as C
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

    # ---- sampling -----------------------------------------------------------------------------------
    def rsample_from_base_samples(self, sample_shape, base_samples):
        """base_samples: sample_shape x [b] x q x 1 standard normal draws -> samples sample_shape x b x q x 1."""
        S = self._joint_cov_std()[0] if self._B == 1 else self._joint_cov_std()
        mean = self._mean_std.reshape(self._B, self._b, self._q)
        mean = mean[0] if self._B == 1 else mean
        q = self._q
        L = torch.linalg.cholesky(S + 1e-8 * torch.eye(q, dtype=S.dtype, device=S.device))
        z = base_samples.to(S.device, torch.float64).squeeze(-1)
        while z.dim() < mean.dim() + len(sample_shape):
            z = z.unsqueeze(len(sample_shape))
        samples = mean + (L @ z.unsqueeze(-1)).squeeze(-1)
        samples = samples * self._scale + self._shift
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
