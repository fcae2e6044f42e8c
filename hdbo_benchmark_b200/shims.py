"""Optional import shims: `install()` registers module trees named `botorch` and `gpytorch` that resolve to this
package's classes, for code written against those names -- the reference's kernel constructor
(`import gpytorch; gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(..., lengthscale_prior=
gpytorch.priors.LogNormalPrior(...)))`, /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:125-138),
its trainer (`gpytorch.mlls.ExactMarginalLogLikelihood`, `gpytorch.likelihoods.GaussianLikelihood`, training_gps.py:3,34,56)
and the BoTorch calls inside the poli-baselines solvers.  Nothing is registered when the real packages import,
unless `force=True`.
"""
from __future__ import annotations

import importlib
import sys
import types


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__hdbo_b200_shim__ = True
    return m


def _real(name: str) -> bool:
    if name in sys.modules:
        return not getattr(sys.modules[name], "__hdbo_b200_shim__", False)
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install(force: bool = False) -> bool:
    """Returns True when the shims were registered."""
    if not force and (_real("botorch") or _real("gpytorch")):
        return False
    from . import acquisition as A
    from . import fit as F
    from . import gp_modules as G
    from . import models as M
    from . import optim as O
    from . import sampling as S
    from .engine import NotPSDError

    gp = _module("gpytorch")
    gp.kernels = _module("gpytorch.kernels", Kernel=G.Kernel, MaternKernel=G.MaternKernel, RBFKernel=G.RBFKernel, ScaleKernel=G.ScaleKernel)
    gp.priors = _module("gpytorch.priors", Prior=G.Prior, GammaPrior=G.GammaPrior, LogNormalPrior=G.LogNormalPrior,
                        NormalPrior=G.NormalPrior, HalfCauchyPrior=G.HalfCauchyPrior)
    gp.constraints = _module("gpytorch.constraints", Interval=G.Interval, GreaterThan=G.GreaterThan, Positive=G.Positive, LessThan=G.LessThan)
    gp.likelihoods = _module("gpytorch.likelihoods", GaussianLikelihood=G.GaussianLikelihood)
    gp.means = _module("gpytorch.means", ConstantMean=G.ConstantMean, ZeroMean=G.ZeroMean)
    gp.mlls = _module("gpytorch.mlls", ExactMarginalLogLikelihood=M.ExactMarginalLogLikelihood, MarginalLogLikelihood=M.MarginalLogLikelihood)
    gp.models = _module("gpytorch.models", ExactGP=M.ExactGPBase)
    gp.utils = _module("gpytorch.utils")
    gp.utils.errors = _module("gpytorch.utils.errors", NotPSDError=NotPSDError)

    bo = _module("botorch", fit_gpytorch_mll=F.fit_gpytorch_mll)
    bo.models = _module("botorch.models", SingleTaskGP=M.SingleTaskGP, SaasFullyBayesianSingleTaskGP=M.SaasFullyBayesianSingleTaskGP)
    bo.models.transforms = _module("botorch.models.transforms", Standardize=M.Standardize, Normalize=M.Normalize)
    bo.models.transforms.outcome = _module("botorch.models.transforms.outcome", Standardize=M.Standardize)
    bo.models.transforms.input = _module("botorch.models.transforms.input", Normalize=M.Normalize)
    bo.models.fully_bayesian = _module("botorch.models.fully_bayesian", SaasFullyBayesianSingleTaskGP=M.SaasFullyBayesianSingleTaskGP)
    bo.fit = _module("botorch.fit", fit_gpytorch_mll=F.fit_gpytorch_mll, fit_gpytorch_mll_scipy=F.fit_gpytorch_mll_scipy)
    analytic = dict(ExpectedImprovement=A.ExpectedImprovement, LogExpectedImprovement=A.LogExpectedImprovement,
                    UpperConfidenceBound=A.UpperConfidenceBound, ProbabilityOfImprovement=A.ProbabilityOfImprovement,
                    PosteriorMean=A.PosteriorMean, AnalyticAcquisitionFunction=A.AnalyticAcquisitionFunction)
    bo.acquisition = _module("botorch.acquisition", AcquisitionFunction=A.AcquisitionFunction, qExpectedImprovement=A.qExpectedImprovement, **analytic)
    bo.acquisition.analytic = _module("botorch.acquisition.analytic", **analytic)
    bo.acquisition.monte_carlo = _module("botorch.acquisition.monte_carlo", qExpectedImprovement=A.qExpectedImprovement)
    bo.optim = _module("botorch.optim", optimize_acqf=O.optimize_acqf, gen_batch_initial_conditions=O.gen_batch_initial_conditions)
    bo.optim.optimize = _module("botorch.optim.optimize", optimize_acqf=O.optimize_acqf)
    bo.generation = _module("botorch.generation", gen_candidates_scipy=O.gen_candidates_scipy)
    bo.sampling = _module("botorch.sampling", SobolQMCNormalSampler=S.SobolQMCNormalSampler)
    bo.sampling.normal = _module("botorch.sampling.normal", SobolQMCNormalSampler=S.SobolQMCNormalSampler)
    bo.utils = _module("botorch.utils")
    bo.utils.sampling = _module("botorch.utils.sampling", draw_sobol_samples=S.draw_sobol_samples,
                                draw_sobol_normal_samples=S.draw_sobol_normal_samples)

    def register(mod):
        sys.modules[mod.__name__] = mod
        for v in list(mod.__dict__.values()):
            if isinstance(v, types.ModuleType) and getattr(v, "__hdbo_b200_shim__", False) and v.__name__.startswith(mod.__name__ + "."):
                register(v)

    register(gp)
    register(bo)
    return True
