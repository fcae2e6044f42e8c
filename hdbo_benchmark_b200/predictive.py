"""Eval-mode marginal log likelihood: log N(y_test | mu(Z), K(Z,Z) - V V^T + noise I) / m  (+ log-priors / m).

This is what GPyTorch computes when the reference's trainer evaluates its test loss,
`mll(model(test_x), test_y)` in eval mode (/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:77-94):
the likelihood adds the noise to the joint predictive distribution over ALL test points and `log_prob` needs the
Cholesky factor of the m x m predictive covariance.  Everything heavy runs in libhdbo_b200: the differentiable-
posterior forward leaves V = K* L^-T in its workspace, then a second (test-side) factorisation is run on
K(Z,Z) + noise I - V V^T with the same kernels as the training fit (`hdbo_gp_fit_downdated`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .engine import DeviceGP, _rup, _stream


def predictive_log_prob(model, Z: torch.Tensor, target: torch.Tensor, log_prior) -> torch.Tensor:
    """Z: [m, d] on the compute device (already input-transformed); target: [m] in the model's outcome units."""
    gp: DeviceGP = model._device_gp()
    if gp.batch != 1:
        raise NotImplementedError("eval-mode MLL is defined for single (non fully-Bayesian) models")
    dev = gp.device
    m, d = Z.shape
    rows = _rup(m, 128)
    f64 = dict(dtype=torch.float64, device=dev)
    Zc = Z.detach().to(torch.float64).contiguous()
    t = torch.as_tensor(target).detach().to(dev, torch.float64).reshape(-1)
    ot = getattr(model, "outcome_transform", None)
    if ot is not None and getattr(model, "_targets_are_raw", False):
        t = (t - ot.means.reshape(()).to(dev)) / ot.stdvs.reshape(()).to(dev)
    need = gp.lib.hdbo_posterior_workspace_bytes(C.byref(gp.desc), rows, 1)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    mean = torch.empty(1, m, **f64)
    var = torch.empty(1, m, **f64)
    check(gp.lib.hdbo_posterior_fwd(C.byref(gp.desc), ptr(Zc), Zc.stride(0), m, 0, ptr(mean), ptr(var), m, ptr(ws), ws.numel(),
                                    _stream()), "hdbo_posterior_fwd")
    V = gp.lib.hdbo_posterior_workspace_V(C.byref(gp.desc), rows, ptr(ws))
    if not V:
        raise _lib.HdboError("hdbo_posterior_workspace_V rejected its arguments")
    # test-side GP: same hyper-parameters, inputs Z, targets y - mu (constant mean 0), covariance downdated by V V^T
    tgp = DeviceGP(Zc, t - mean[0], gp.kind)
    tgp.center.copy_(gp.center)
    tgp.inv_ls.copy_(gp.inv_ls)
    tgp.hyp.copy_(gp.hyp)
    tgp.hyp[:, 2] = 0.0
    tries = [0.0, 1e-8, 1e-7, 1e-6]
    for jitter in tries:
        tgp.hyp[:, 3] = jitter
        check(gp.lib.hdbo_gp_fit_downdated(C.byref(tgp.desc), 0, V, gp.n_pad, gp.n_pad, _stream()), "hdbo_gp_fit_downdated")
        if int(tgp.info.item()) == 0:
            break
    else:
        from .engine import NotPSDError
        raise NotPSDError("predictive covariance not positive definite after the jitter schedule")
    logp = tgp.scal[0, 2]
    lp = log_prior if torch.is_tensor(log_prior) else torch.as_tensor(float(log_prior), **f64)
    return (logp + lp.detach().to(dev)) / m
