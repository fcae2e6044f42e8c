"""Joint posterior sampling over a large candidate set (Thompson sampling) -- SURVEY §8(f) rank 2.

TuRBO / BAxUS draw `batch_size` joint samples of the GP posterior over <= 5000 Sobol candidates and take the arg-max of each
(botorch.generation.sampling.MaxPosteriorSampling; solvers selected at
/root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:224-246).  Upstream that is `model.posterior(X_cand)` with
q = m, i.e. the full m x m predictive covariance, its Cholesky factor and `mean + L z`.

Here the n- and m-sized work stays in libhdbo_b200:
    1. `hdbo_posterior_fwd`         mean and V = K* L^-T (kept in the workspace)
    2. `hdbo_gp_fit_downdated`      K(Z,Z) + jitter I - V V^T -> its Cholesky factor (the recursive DMMA factorisation of the fit)
    3. `hdbo_dgemm_nt`              samples = mean + base_samples . L_Sigma^T (triangular k-range)
    4. `hdbo_argmax`                per sample
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr
from .engine import DeviceGP, NotPSDError, _rup, _stream

KR_LE_N = 1   # csrc/kernels.cuh: B lower triangular (k-tile <= n-tile)


def joint_posterior_samples(gp: DeviceGP, Z: torch.Tensor, base_samples: torch.Tensor, observation_noise: bool = False,
                            jitters=(1e-8, 1e-7, 1e-6, 1e-5, 1e-4)):
    """Z [m, d] (device, model input space), base_samples [S, m] standard normal -> (samples [S, m], mean [m], jitter used),
    in the GP's (standardised) outcome units.  The jitter schedule mirrors psd_safe_cholesky on the m x m covariance."""
    if gp.batch != 1:
        raise NotImplementedError("joint sampling is defined for single (non fully-Bayesian) models")
    assert gp.fitted, "call fit() first"
    dev = gp.device
    f64 = dict(dtype=torch.float64, device=dev)
    Zc = Z.detach().to(dev, torch.float64).contiguous()
    m = Zc.shape[0]
    S = base_samples.shape[0]
    assert base_samples.shape[1] == m
    rows = _rup(m, 128)
    need = gp.lib.hdbo_posterior_workspace_bytes(C.byref(gp.desc), rows, 1)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    mean = torch.empty(1, m, **f64)
    var = torch.empty(1, m, **f64)
    check(gp.lib.hdbo_posterior_fwd(C.byref(gp.desc), ptr(Zc), Zc.stride(0), m, 0, ptr(mean), ptr(var), m, ptr(ws), ws.numel(),
                                    _stream()), "hdbo_posterior_fwd")
    V = gp.lib.hdbo_posterior_workspace_V(C.byref(gp.desc), rows, ptr(ws))
    if not V:
        raise _lib.HdboError("hdbo_posterior_workspace_V rejected its arguments")
    # candidate-side factorisation: same kernel hyper-parameters, inputs Z, covariance downdated by V V^T
    tgp = DeviceGP(Zc, torch.zeros(m, **f64), gp.kind)
    tgp.center.copy_(gp.center)
    tgp.inv_ls.copy_(gp.inv_ls)
    tgp.hyp.copy_(gp.hyp)
    tgp.hyp[:, 2] = 0.0
    if not observation_noise:
        tgp.hyp[:, 1] = 0.0
    used = None
    for jitter in jitters:
        tgp.hyp[:, 3] = jitter
        check(gp.lib.hdbo_gp_fit_downdated(C.byref(tgp.desc), 0, V, gp.n_pad, gp.n_pad, _stream()), "hdbo_gp_fit_downdated")
        if int(tgp.info.item()) == 0:
            used = jitter
            break
    if used is None:
        raise NotPSDError("joint predictive covariance not positive definite after the jitter schedule")
    # samples = mean + z L^T : NT GEMM with the lower-triangular factor as the B operand
    Sp = _rup(S, 128)
    zs = torch.zeros(Sp, rows, **f64)
    zs[:S, :m] = base_samples.to(dev, torch.float64)
    out = torch.empty(Sp, rows, **f64)
    check(gp.lib.hdbo_dgemm_nt(ptr(zs), rows, ptr(tgp.L), rows, ptr(out), rows, Sp, rows, rows, 1.0, 0.0, KR_LE_N, 0, _stream()),
          "hdbo_dgemm_nt")
    samples = out[:S, :m] + mean[0]
    return samples, mean[0], used


class MaxPosteriorSampling:
    """botorch.generation.sampling.MaxPosteriorSampling (replacement=True/False) for single-output models.

    `__call__(X_cand [m, d], num_samples)` -> X_cand rows that maximise each joint posterior sample, [num_samples, d]."""

    def __init__(self, model, replacement: bool = True, seed: Optional[int] = None):
        self.model = model
        self.replacement = replacement
        self._gen = torch.Generator(device="cpu")
        if seed is not None:
            self._gen.manual_seed(seed)

    def __call__(self, X: torch.Tensor, num_samples: int = 1, observation_noise: bool = False) -> torch.Tensor:
        model = self.model
        gp = model._device_gp()
        Xd = X.reshape(-1, X.shape[-1])
        Z, _, _ = model._prep_X(Xd)                     # device, float64, input-transformed
        m = Z.shape[0]
        z = torch.randn(num_samples, m, dtype=torch.float64, generator=self._gen)
        samples, _, _ = joint_posterior_samples(gp, Z, z, observation_noise=observation_noise)
        if self.replacement:
            idx = torch.stack([gp.argmax(samples[s])[1] for s in range(num_samples)]).reshape(-1)
        else:   # sequentially exclude chosen candidates (upstream `_get_max_sample_without_replacement`)
            taken = torch.zeros(m, dtype=torch.bool, device=samples.device)
            picks = []
            for s in range(num_samples):
                row = samples[s].masked_fill(taken, float("-inf"))
                i = gp.argmax(row)[1].reshape(())
                taken[i] = True
                picks.append(i)
            idx = torch.stack(picks)
        return Xd[idx.to(Xd.device)]
