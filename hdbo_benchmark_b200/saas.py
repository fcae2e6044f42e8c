"""SAAS NUTS potential energy and gradient, batched over chains / trajectory points.

For `saas_bo` (solver selected at /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:183-194) the wall clock
is the MCMC: Pyro's NUTS differentiates the potential of botorch's SaasPyroModel 1e4-1e5 times per fit at n <= 512.  The sampler
itself (Pyro control flow) is out of scope; this module is the function it calls, evaluated for C states at once by ONE batched
kernel build + Cholesky + inverse (`hdbo_gp_fit`, grid.z = C) and ONE batched analytic gradient (`hdbo_mll_grad`).

    z = [log inv_length_sq (d), log tausq, log outputscale, log noise, mean]            (unconstrained NUTS coordinates)
    U(z) = -[ n * mll_data(theta(z)) + log prior(theta(z)) + log |d theta / d z| ]
    mean ~ N(0,1); outputscale ~ Gamma(2, 0.15); noise ~ Gamma(0.9, 10); tausq ~ HalfCauchy(0.1); inv_length_sq_k ~ HalfCauchy(1);
    lengthscale_k = (tausq * inv_length_sq_k)^(-1/2); Matern-5/2 ARD.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .engine import DeviceGP, NotPSDError
from .models import compute_device, mll_gradient

_LOG_2PI = math.log(2.0 * math.pi)


class SaasPotential:
    """`U, dU = SaasPotential(train_X, train_Y)(Z)` with Z [C, d + 4]; outputs on the compute device, [C] and [C, d + 4]."""

    def __init__(self, train_X: torch.Tensor, train_Y: torch.Tensor, min_noise: float = 0.0):
        self.dev = compute_device()
        self.X = train_X.detach().to(self.dev, torch.float64).contiguous()
        self.y = train_Y.detach().to(self.dev, torch.float64).reshape(-1).contiguous()
        self.n, self.d = self.X.shape
        self.min_noise = float(min_noise)
        self._gp = None

    def _gp_for(self, C: int) -> DeviceGP:
        if self._gp is None or self._gp.batch != C:
            self._gp = DeviceGP(self.X, self.y, _lib.KERNEL_MATERN52, batch=C, keep_r2=True)
        return self._gp

    def __call__(self, Z: torch.Tensor):
        Z = torch.as_tensor(Z).detach().to(self.dev, torch.float64)
        if Z.dim() == 1:
            Z = Z.unsqueeze(0)
        C, d = Z.shape[0], self.d
        assert Z.shape[1] == d + 4
        z_l, z_t, z_o, z_n, mean = Z[:, :d], Z[:, d], Z[:, d + 1], Z[:, d + 2], Z[:, d + 3]
        ils, tausq, os_, noise = torch.exp(z_l), torch.exp(z_t), torch.exp(z_o), torch.exp(z_n)
        ls = torch.rsqrt(tausq.unsqueeze(1) * ils)
        gp = self._gp_for(C)
        gp.set_hyper(ls, os_, noise + self.min_noise, mean)
        gp.fit_once(need_r2=True)
        bad = gp.info != 0                                        # not positive definite: infinite energy (NUTS rejects)
        ll = gp.mll_data_term.clone()                             # n * mll data term, [C]
        g = mll_gradient(gp)                                      # d ll / d (ls[d], os, noise, mean), [C, d+3]
        # priors and log-Jacobians of the exp transforms
        hc = lambda v, s: math.log(2.0 / math.pi) - math.log(s) - torch.log1p((v / s) ** 2)
        gam = lambda v, c, r: c * math.log(r) + (c - 1.0) * torch.log(v) - r * v - math.lgamma(c)
        lp = gam(os_, 2.0, 0.15) + gam(noise, 0.9, 10.0) + hc(tausq, 0.1) + hc(ils, 1.0).sum(1) - 0.5 * mean * mean - 0.5 * _LOG_2PI
        jac = z_l.sum(1) + z_t + z_o + z_n
        U = -(ll + lp + jac)
        # chain rule to z (d theta / d z = theta for the exp sites; ls_k = exp(-(z_t + z_lk) / 2))
        g_ls = g[:, :d] * ls * (-0.5)                              # d ll / d z_lk  (also each one's share of d ll / d z_t)
        dU = torch.empty_like(Z)
        dU[:, :d] = -(g_ls + (-2.0 * ils * ils / (1.0 + ils * ils)) + 1.0)
        dU[:, d] = -(g_ls.sum(1) + (-2.0 * (tausq / 0.1) ** 2 / (1.0 + (tausq / 0.1) ** 2)) + 1.0)
        dU[:, d + 1] = -(g[:, d] * os_ + (2.0 - 1.0) - 0.15 * os_ + 1.0)
        dU[:, d + 2] = -(g[:, d + 1] * noise + (0.9 - 1.0) - 10.0 * noise + 1.0)
        dU[:, d + 3] = -(g[:, d + 2] - mean)
        inf = torch.full_like(U, float("inf"))
        return torch.where(bad.to(U.device), inf, U), torch.where(bad.to(U.device).unsqueeze(1), torch.zeros_like(dU), dU)
