"""GPyTorch-shaped hyper-parameter containers: constraints, priors, kernels, mean, likelihood.

These classes hold raw parameters, constraints and priors with the same names GPyTorch uses
(`covar_module.base_kernel.raw_lengthscale`, `covar_module.raw_outputscale`,
`likelihood.noise_covar.raw_noise`, `mean_module.raw_constant`) so that `state_dict()` /
`load_state_dict()`, `model.parameters()` + Adam (reference:
/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:53,101-108) and the kernel
constructor of the reference (/root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:131-138:
`ScaleKernel(RBFKernel(ard_num_dims=d, lengthscale_prior=LogNormalPrior(log(d)/2, 1)))`) work unchanged.
They contain no kernel-matrix arithmetic: that lives in libhdbo_b200.so.  Only the tiny
raw -> constrained transforms and log-prior densities (O(d) work) are torch ops here.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from ._lib import KERNEL_MATERN52, KERNEL_RBF


# ----------------------------------------------------------------------------------------------
# constraints (gpytorch.constraints)
# ----------------------------------------------------------------------------------------------
def _softplus(x):
    return torch.nn.functional.softplus(x)


def _inv_softplus(x):
    x = torch.as_tensor(x, dtype=torch.float64)
    return x + torch.log(-torch.expm1(-x))


class Interval(nn.Module):
    """lower < value < upper.  transform='softplus' (GreaterThan/Positive default), 'sigmoid' (Interval default)
    or None (value == raw; the bounds are then enforced by the optimiser, as BoTorch >= 0.12 does)."""

    def __init__(self, lower_bound, upper_bound, transform="sigmoid", initial_value=None):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound), dtype=torch.float64))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound), dtype=torch.float64))
        self._transform = transform
        self.initial_value = initial_value

    @property
    def enforced(self) -> bool:
        return self._transform is not None

    def transform(self, raw):
        if self._transform is None:
            return raw
        if self._transform == "softplus":
            if math.isinf(float(self.upper_bound)):
                return _softplus(raw) + self.lower_bound
            return self.upper_bound - _softplus(raw)
        return torch.sigmoid(raw) * (self.upper_bound - self.lower_bound) + self.lower_bound

    def inverse_transform(self, value):
        value = torch.as_tensor(value, dtype=torch.float64)
        if self._transform is None:
            return value
        if self._transform == "softplus":
            if math.isinf(float(self.upper_bound)):
                return _inv_softplus(value - self.lower_bound)
            return _inv_softplus(self.upper_bound - value)
        p = (value - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(p) - torch.log1p(-p)

    def raw_bounds(self):
        """Bounds on the RAW parameter for L-BFGS-B (botorch get_parameters_and_bounds semantics)."""
        if self._transform is None:
            return float(self.lower_bound), float(self.upper_bound)
        return -math.inf, math.inf


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform="softplus", initial_value=None):
        super().__init__(lower_bound, math.inf, transform=transform, initial_value=initial_value)


class Positive(GreaterThan):
    def __init__(self, transform="softplus", initial_value=None):
        super().__init__(0.0, transform=transform, initial_value=initial_value)


class LessThan(Interval):
    def __init__(self, upper_bound, transform="softplus", initial_value=None):
        super().__init__(-math.inf, upper_bound, transform=transform, initial_value=initial_value)


# ----------------------------------------------------------------------------------------------
# priors (gpytorch.priors): log densities, summed by the MLL
# ----------------------------------------------------------------------------------------------
class Prior(nn.Module):
    def log_prob(self, x):  # pragma: no cover - interface
        raise NotImplementedError

    @property
    def mode(self):
        raise NotImplementedError


class GammaPrior(Prior):
    def __init__(self, concentration, rate):
        super().__init__()
        self.register_buffer("concentration", torch.as_tensor(float(concentration), dtype=torch.float64))
        self.register_buffer("rate", torch.as_tensor(float(rate), dtype=torch.float64))

    def log_prob(self, x):
        c, r = self.concentration, self.rate
        return c * torch.log(r) + (c - 1.0) * torch.log(x) - r * x - torch.lgamma(c)

    @property
    def mode(self):
        return torch.clamp((self.concentration - 1.0) / self.rate, min=0.0)


class LogNormalPrior(Prior):
    def __init__(self, loc, scale):
        super().__init__()
        self.register_buffer("loc", torch.as_tensor(float(loc), dtype=torch.float64))
        self.register_buffer("scale", torch.as_tensor(float(scale), dtype=torch.float64))

    def log_prob(self, x):
        lx = torch.log(x)
        return -lx - torch.log(self.scale) - 0.5 * math.log(2 * math.pi) - 0.5 * ((lx - self.loc) / self.scale) ** 2

    @property
    def mode(self):
        return torch.exp(self.loc - self.scale ** 2)


class NormalPrior(Prior):
    def __init__(self, loc, scale):
        super().__init__()
        self.register_buffer("loc", torch.as_tensor(float(loc), dtype=torch.float64))
        self.register_buffer("scale", torch.as_tensor(float(scale), dtype=torch.float64))

    def log_prob(self, x):
        return -torch.log(self.scale) - 0.5 * math.log(2 * math.pi) - 0.5 * ((x - self.loc) / self.scale) ** 2

    @property
    def mode(self):
        return self.loc


class HalfCauchyPrior(Prior):
    def __init__(self, scale):
        super().__init__()
        self.register_buffer("scale", torch.as_tensor(float(scale), dtype=torch.float64))

    def log_prob(self, x):
        return math.log(2.0 / math.pi) - torch.log(self.scale) - torch.log1p((x / self.scale) ** 2)

    @property
    def mode(self):
        return torch.zeros_like(self.scale)


# ----------------------------------------------------------------------------------------------
# parameter plumbing shared by kernels / likelihood / mean
# ----------------------------------------------------------------------------------------------
class _HyperModule(nn.Module):
    """nn.Module with (raw parameter, constraint, optional prior) triples registered under GPyTorch's names."""

    def _register(self, name: str, shape, constraint: Interval, prior: Optional[Prior], init=None):
        raw = nn.Parameter(torch.zeros(shape, dtype=torch.float64))
        self.register_parameter(f"raw_{name}", raw)
        self.add_module(f"raw_{name}_constraint", constraint)
        if prior is not None:
            self.add_module(f"{name}_prior", prior)
        if not hasattr(self, "_hyper_names"):
            self._hyper_names = []
        self._hyper_names.append(name)
        if init is None:
            if constraint.initial_value is not None:
                init = constraint.initial_value
            elif constraint._transform is None:
                init = max(float(constraint.lower_bound), 0.0) + 1.0 if math.isinf(float(constraint.upper_bound)) else \
                    0.5 * (float(constraint.lower_bound) + float(constraint.upper_bound))
        if init is not None:
            self._set(name, init)

    def _get(self, name):
        return getattr(self, f"raw_{name}_constraint").transform(getattr(self, f"raw_{name}"))

    def _set(self, name, value):
        raw = getattr(self, f"raw_{name}")
        c = getattr(self, f"raw_{name}_constraint")
        v = torch.as_tensor(value, dtype=torch.float64, device=raw.device)
        with torch.no_grad():
            raw.copy_(c.inverse_transform(v.to(raw.device)).expand_as(raw))

    def named_priors_with_values(self, prefix=""):
        """Yields (qualified name, prior, constrained value) like gpytorch Module.named_priors."""
        for name in getattr(self, "_hyper_names", []):
            prior = getattr(self, f"{name}_prior", None)
            if prior is not None:
                yield f"{prefix}{name}_prior", prior, self._get(name)
        for child_name, child in self.named_children():
            if isinstance(child, _HyperModule):
                yield from child.named_priors_with_values(prefix=f"{prefix}{child_name}.")

    def raw_parameter_bounds(self, prefix=""):
        """{qualified raw parameter name: (lower, upper)} for the optimiser."""
        out = {}
        for name in getattr(self, "_hyper_names", []):
            out[f"{prefix}raw_{name}"] = getattr(self, f"raw_{name}_constraint").raw_bounds()
        for child_name, child in self.named_children():
            if isinstance(child, _HyperModule):
                out.update(child.raw_parameter_bounds(prefix=f"{prefix}{child_name}."))
        return out


# ----------------------------------------------------------------------------------------------
# kernels (gpytorch.kernels)
# ----------------------------------------------------------------------------------------------
class Kernel(_HyperModule):
    has_lengthscale = True
    kind = None

    def __init__(self, ard_num_dims: Optional[int] = None, batch_shape=torch.Size([]), lengthscale_prior=None,
                 lengthscale_constraint=None, active_dims=None, eps=1e-6, **kwargs):
        super().__init__()
        if active_dims is not None:
            raise NotImplementedError("active_dims is not supported by the B200 hot path")
        self.ard_num_dims = ard_num_dims
        self.batch_shape = torch.Size(batch_shape)
        nd = 1 if ard_num_dims is None else ard_num_dims
        constraint = lengthscale_constraint if lengthscale_constraint is not None else Positive()
        init = None
        if constraint.initial_value is None and constraint.enforced:
            init = math.log(2.0)  # raw = 0 under softplus, GPyTorch's default initialisation
        self._register("lengthscale", (*self.batch_shape, 1, nd), constraint, lengthscale_prior, init=init)

    @property
    def lengthscale(self):
        return self._get("lengthscale")

    @lengthscale.setter
    def lengthscale(self, value):
        self._set("lengthscale", value)


class MaternKernel(Kernel):
    kind = KERNEL_MATERN52

    def __init__(self, nu: float = 2.5, **kwargs):
        if nu != 2.5:
            raise NotImplementedError("only Matern-5/2 is on the hdbo_benchmark hot path")
        super().__init__(**kwargs)
        self.nu = nu


class RBFKernel(Kernel):
    kind = KERNEL_RBF


class ScaleKernel(_HyperModule):
    def __init__(self, base_kernel: Kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        super().__init__()
        self.base_kernel = base_kernel
        constraint = outputscale_constraint if outputscale_constraint is not None else Positive()
        init = math.log(2.0) if (constraint.initial_value is None and constraint.enforced) else None
        self._register("outputscale", tuple(base_kernel.batch_shape), constraint, outputscale_prior, init=init)

    @property
    def outputscale(self):
        return self._get("outputscale")

    @outputscale.setter
    def outputscale(self, value):
        self._set("outputscale", value)


# ----------------------------------------------------------------------------------------------
# mean (gpytorch.means.ConstantMean) and likelihood (gpytorch.likelihoods.GaussianLikelihood)
# ----------------------------------------------------------------------------------------------
class ConstantMean(nn.Module):
    def __init__(self, constant_prior=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.batch_shape = torch.Size(batch_shape)
        self.raw_constant = nn.Parameter(torch.zeros(self.batch_shape, dtype=torch.float64))
        if constant_prior is not None:
            self.constant_prior = constant_prior

    @property
    def constant(self):
        return self.raw_constant

    @constant.setter
    def constant(self, value):
        with torch.no_grad():
            self.raw_constant.copy_(torch.as_tensor(value, dtype=torch.float64).expand_as(self.raw_constant))


class ZeroMean(nn.Module):
    @property
    def constant(self):
        return torch.zeros((), dtype=torch.float64)


class HomoskedasticNoise(_HyperModule):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([])):
        super().__init__()
        constraint = noise_constraint if noise_constraint is not None else GreaterThan(1e-4)
        init = None
        if constraint.initial_value is None and constraint.enforced:
            init = float(constraint.lower_bound) + math.log(2.0)
        self._register("noise", (*torch.Size(batch_shape), 1), constraint, noise_prior, init=init)

    @property
    def noise(self):
        return self._get("noise")

    @noise.setter
    def noise(self, value):
        self._set("noise", value)


class GaussianLikelihood(_HyperModule):
    """Homoskedastic Gaussian noise; `likelihood.noise_covar.raw_noise` as in GPyTorch
    (constructed directly by the reference at training_gps.py:34)."""

    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior, noise_constraint, batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    def forward(self, dist, *args, **kwargs):
        """likelihood(model(x)): adds the observation noise to a predictive distribution."""
        return dist.with_observation_noise(self.noise) if hasattr(dist, "with_observation_noise") else dist
