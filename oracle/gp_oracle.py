"""CPU FP64 oracle for the GP-surrogate + acquisition hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain torch float64 on the CPU, the arithmetic that the reference
(`MachineLearningLifeScience/hdbo_benchmark`) reaches through its un-vendored, unpinned
dependencies BoTorch / GPyTorch / linear_operator.  The
reference repository itself contains none of this arithmetic; its call sites are

* kernel constructor:          /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:124-145
* solvers that consume it:     /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-194
* GPyTorch driven directly:    /root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:49-145

PARITY STATUS: *parity unpinned by the reference* -- the reference's only test
(/root/reference/test_main_entry_point.py:18,42-58) asserts no numeric value, and
botorch/gpytorch are not installable here (no network, not in /opt/wheelhouse).  The oracle
is therefore pinned against independent implementations that ARE installed:
scikit-learn's GaussianProcessRegressor (MLL, MLL gradient, predictive mean/variance),
scipy.stats/scipy.special (EI, LogEI pieces) and mpmath at 50 digits (small-n GP, LogEI
tails); see tests/test_oracle_pins.py and tests/golden/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (hdbo_benchmark_b200/) never does.

Upstream algorithm pins (so that "the reference FP64 path" is well defined):
float64 everywhere; exact Cholesky at every n (equivalent of gpytorch.settings
max_cholesky_size(inf), fast_computations(False, False, False)); squared distance by the
centred GEMM expansion, clamp at 0, then clamp at 1e-30 before sqrt (Matern); noise lower
bound 1e-4; variance clamp 1e-12 inside the acquisition; MLL divided by n; jitter schedule
1e-8 * 10**i, i < 3 (linear_operator.psd_safe_cholesky in float64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

MATERN52 = 0
RBF = 1

SQRT5 = math.sqrt(5.0)
LOG_2PI = math.log(2.0 * math.pi)
INV_SQRT_2 = 1.0 / math.sqrt(2.0)
INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)
LOG_SQRT_PI_DIV_2 = 0.5 * math.log(math.pi / 2.0)
MIN_VAR = 1e-12  # botorch.acquisition.analytic: _mean_and_sigma(min_var=1e-12)
MIN_NOISE = 1e-4  # botorch.models.utils.gpytorch_modules.MIN_INFERRED_NOISE_LEVEL


# ----------------------------------------------------------------------------------------
# A1  kernels  (gpytorch.kernels.{MaternKernel(nu=2.5), RBFKernel}, ScaleKernel)
# ----------------------------------------------------------------------------------------
def scaled_sq_dist(x1, x2, lengthscale, center=None, x1_eq_x2=False):
    """r^2 between rows of x1 [m,d] and x2 [n,d] after dividing by the ARD lengthscale [d].

    Follows gpytorch.kernels.kernel.sq_dist: subtract a common centre (GPyTorch uses the mean
    of x1; any common shift is exact in real arithmetic, we use `center` = the training mean
    so that train/train and test/train blocks share it), expand as
    |a|^2 + |b|^2 - 2 a.b with one matmul, force the diagonal to exactly 0 when x1 is x2,
    clamp at 0.
    """
    if center is None:
        center = x1.mean(dim=-2, keepdim=True)
    a = (x1 - center) / lengthscale
    b = (x2 - center) / lengthscale
    a2 = (a * a).sum(-1, keepdim=True)
    b2 = (b * b).sum(-1, keepdim=True)
    r2 = a2 + b2.transpose(-1, -2) - 2.0 * (a @ b.transpose(-1, -2))
    if x1_eq_x2:
        r2 = r2 - torch.diag_embed(torch.diagonal(r2, dim1=-2, dim2=-1))
    return r2.clamp_min(0.0)


def kernel_from_sq_dist(r2, kind):
    """Unit-outputscale kernel value from r^2 (GPyTorch MaternKernel.forward / postprocess_rbf)."""
    if kind == MATERN52:
        r = r2.clamp_min(1e-30).sqrt()
        return (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r) * torch.exp(-SQRT5 * r)
    if kind == RBF:
        return torch.exp(-0.5 * r2)
    raise ValueError(kind)


def dkernel_dsq_dist(r2, kind):
    """d k / d r^2 (unit outputscale).  Matern-5/2: -(5/6)(1+sqrt5 r) exp(-sqrt5 r); RBF: -k/2."""
    if kind == MATERN52:
        r = r2.clamp_min(1e-30).sqrt()
        return -(5.0 / 6.0) * (1.0 + SQRT5 * r) * torch.exp(-SQRT5 * r)
    if kind == RBF:
        return -0.5 * torch.exp(-0.5 * r2)
    raise ValueError(kind)


def kernel_matrix(x1, x2, lengthscale, outputscale, kind, center=None, x1_eq_x2=False):
    """sigma_f^2 * k(x1, x2); ScaleKernel wraps the base kernel (outputscale=1 when absent)."""
    return outputscale * kernel_from_sq_dist(scaled_sq_dist(x1, x2, lengthscale, center, x1_eq_x2), kind)


# ----------------------------------------------------------------------------------------
# A2  noise + Cholesky with jitter retry  (GaussianLikelihood, linear_operator.psd_safe_cholesky)
# ----------------------------------------------------------------------------------------
class NotPSDError(RuntimeError):
    pass


def psd_safe_cholesky(K, max_tries=3):
    """cholesky(K); on failure retry with jitter 1e-8 * 10**i added to the diagonal (float64)."""
    L, info = torch.linalg.cholesky_ex(K)
    if not bool((info > 0).any()):
        return L, 0.0
    prev = 0.0
    Kj = K.clone()
    for i in range(max_tries):
        jitter = 1e-8 * (10 ** i)
        Kj.diagonal(dim1=-2, dim2=-1).add_(jitter - prev)
        prev = jitter
        L, info = torch.linalg.cholesky_ex(Kj)
        if not bool((info > 0).any()):
            return L, jitter
    raise NotPSDError(f"Matrix not positive definite after adding jitter up to {prev:.1e}")


@dataclass
class GPState:
    """Everything the posterior needs once hyper-parameters are fixed (prediction-strategy cache)."""

    X: torch.Tensor  # [n, d]
    y: torch.Tensor  # [n]
    center: torch.Tensor  # [1, d]
    lengthscale: torch.Tensor  # [d]
    outputscale: float
    noise: float
    mean_const: float
    kind: int
    L: torch.Tensor  # [n, n] lower Cholesky factor of K + noise I
    Linv: torch.Tensor  # [n, n] lower, L^{-1}
    alpha: torch.Tensor  # [n] = (K + noise I)^{-1} (y - m)
    logdet: float  # log det (K + noise I) = 2 sum log L_ii


def fit_state(X, y, lengthscale, outputscale, noise, mean_const, kind) -> GPState:
    X = X.double()
    y = y.double().reshape(-1)
    n = X.shape[0]
    center = X.mean(dim=0, keepdim=True)
    K = kernel_matrix(X, X, lengthscale, outputscale, kind, center, x1_eq_x2=True)
    K = K + noise * torch.eye(n, dtype=torch.float64)
    L, _ = psd_safe_cholesky(K)
    r = y - mean_const
    alpha = torch.cholesky_solve(r.unsqueeze(-1), L).squeeze(-1)
    Linv = torch.linalg.solve_triangular(L, torch.eye(n, dtype=torch.float64), upper=False)
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum().item()
    return GPState(X, y, center, lengthscale.double(), float(outputscale), float(noise), float(mean_const),
                   kind, L, Linv, alpha, logdet)


# ----------------------------------------------------------------------------------------
# A3  exact marginal log likelihood  (gpytorch.mlls.ExactMarginalLogLikelihood) + analytic gradient
# ----------------------------------------------------------------------------------------
def exact_mll(X, y, lengthscale, outputscale, noise, mean_const, kind, log_prior=0.0):
    """[-1/2 r^T Khat^-1 r - sum log L_ii - n/2 log 2pi + log_prior] / n.  Differentiable (autograd)."""
    n = X.shape[0]
    center = X.mean(dim=0, keepdim=True)
    K = kernel_matrix(X, X, lengthscale, outputscale, kind, center, x1_eq_x2=not torch.is_grad_enabled())
    K = K + noise * torch.eye(n, dtype=X.dtype)
    L = torch.linalg.cholesky(K)
    r = (y.reshape(-1) - mean_const).unsqueeze(-1)
    alpha = torch.cholesky_solve(r, L)
    quad = (r * alpha).sum()
    logdet_half = torch.log(torch.diagonal(L)).sum()
    return (-0.5 * quad - logdet_half - 0.5 * n * LOG_2PI + log_prior) / n


def exact_mll_and_grad(X, y, lengthscale, outputscale, noise, mean_const, kind):
    """Value and analytic gradient of n*mll/n w.r.t. (lengthscale[d], outputscale, noise, mean), no priors.

    W = alpha alpha^T - Khat^-1;  G = W o (-2 sigma_f^2 dk/dr^2);
    d/dl_k = [ (Xc o Xc)^T rowsum(G) - diag(Xc^T G Xc) ]_k / l_k^3 ;  d/dsigma_f^2 = 1/2 sum(W o K)/sigma_f^2 ;
    d/dsigma^2 = 1/2 tr W ;  d/dm = sum alpha ; everything divided by n.   (SURVEY.md Appendix C.)
    """
    n = X.shape[0]
    center = X.mean(dim=0, keepdim=True)
    Xc = X - center
    r2 = scaled_sq_dist(X, X, lengthscale, center, x1_eq_x2=True)
    K = outputscale * kernel_from_sq_dist(r2, kind)
    Khat = K + noise * torch.eye(n, dtype=X.dtype)
    L = torch.linalg.cholesky(Khat)
    r = (y.reshape(-1) - mean_const)
    alpha = torch.cholesky_solve(r.unsqueeze(-1), L).squeeze(-1)
    Kinv = torch.cholesky_inverse(L)
    value = (-0.5 * (r * alpha).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * LOG_2PI) / n
    W = torch.outer(alpha, alpha) - Kinv
    G = W * (-2.0 * outputscale * dkernel_dsq_dist(r2, kind))
    g_ls = ((Xc * Xc).T @ G.sum(1) - ((Xc.T @ G) * Xc.T).sum(1)) / lengthscale ** 3
    g_os = 0.5 * (W * K).sum() / outputscale
    g_noise = 0.5 * torch.diagonal(W).sum()
    g_mean = alpha.sum()
    return value, g_ls / n, g_os / n, g_noise / n, g_mean / n


# priors (gpytorch.priors): log densities summed over elements, added before the division by n
def gamma_log_prob(x, concentration, rate):
    x = torch.as_tensor(x, dtype=torch.float64)
    return (concentration * math.log(rate) + (concentration - 1.0) * torch.log(x) - rate * x
            - math.lgamma(concentration)).sum()


def lognormal_log_prob(x, loc, scale):
    x = torch.as_tensor(x, dtype=torch.float64)
    lx = torch.log(x)
    return (-lx - math.log(scale) - 0.5 * LOG_2PI - 0.5 * ((lx - loc) / scale) ** 2).sum()


# ----------------------------------------------------------------------------------------
# A4  posterior  (botorch GPyTorchModel.posterior -> gpytorch DefaultPredictionStrategy)
# ----------------------------------------------------------------------------------------
def posterior(state: GPState, Z, observation_noise=False) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean [m], variance [m] for independent candidates Z [m,d] (q=1).

    mu = m + K* alpha ;  V = K* L^-T ;  var = sigma_f^2 - rowsum(V o V)  (+ noise)."""
    Ks = kernel_matrix(Z, state.X, state.lengthscale, state.outputscale, state.kind, state.center)
    mean = state.mean_const + Ks @ state.alpha
    V = Ks @ state.Linv.T
    var = state.outputscale - (V * V).sum(-1)
    if observation_noise:
        var = var + state.noise
    return mean, var


def posterior_joint(state: GPState, Zq, observation_noise=False):
    """Joint posterior over q points: Zq [b,q,d] -> mean [b,q], Sigma [b,q,q] = K** - V V^T."""
    b, q, d = Zq.shape
    Zf = Zq.reshape(b * q, d)
    Ks = kernel_matrix(Zf, state.X, state.lengthscale, state.outputscale, state.kind, state.center)
    mean = (state.mean_const + Ks @ state.alpha).reshape(b, q)
    V = (Ks @ state.Linv.T).reshape(b, q, -1)
    a = (Zq - state.center) / state.lengthscale
    a2 = (a * a).sum(-1, keepdim=True)
    r2 = (a2 + a2.transpose(-1, -2) - 2.0 * a @ a.transpose(-1, -2))
    r2 = (r2 - torch.diag_embed(torch.diagonal(r2, dim1=-2, dim2=-1))).clamp_min(0.0)
    Kss = state.outputscale * kernel_from_sq_dist(r2, state.kind)
    Sigma = Kss - V @ V.transpose(-1, -2)
    if observation_noise:
        Sigma = Sigma + state.noise * torch.eye(q, dtype=Sigma.dtype)
    return mean, Sigma


# ----------------------------------------------------------------------------------------
# A5  analytic acquisition functions  (botorch.acquisition.analytic)
# ----------------------------------------------------------------------------------------
def _phi(u):
    return INV_SQRT_2PI * torch.exp(-0.5 * u * u)


def _Phi(u):
    return 0.5 * torch.erfc(-INV_SQRT_2 * u)


def _log_phi(u):
    return -0.5 * (u * u + LOG_2PI)


def _log1mexp(x):
    """log(1 - exp(x)) for x < 0 (botorch.utils.safe_math.log1mexp)."""
    return torch.where(x > -math.log(2.0), torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def _log_ei_helper(u):
    """log(phi(u) + u Phi(u)), accurate in the far negative tail (botorch _log_ei_helper, float64 bounds)."""
    bound = -1.0
    neg_inv_sqrt_eps = -1e6
    u_upper = u.clamp_min(bound)
    log_upper = torch.log(_phi(u_upper) + u_upper * _Phi(u_upper))
    u_lower = u.clamp_max(bound)
    u_eps = u_lower.clamp_min(neg_inv_sqrt_eps)
    w = torch.log(torch.special.erfcx(-INV_SQRT_2 * u_eps) * u_eps.abs()) + LOG_SQRT_PI_DIV_2
    log_lower = _log_phi(u) + torch.where(u > neg_inv_sqrt_eps, _log1mexp(w), -2.0 * torch.log(u_lower.abs()))
    return torch.where(u > bound, log_upper, log_lower)


def expected_improvement(mean, var, best_f):
    sigma = var.clamp_min(MIN_VAR).sqrt()
    u = (mean - best_f) / sigma
    return sigma * (_phi(u) + u * _Phi(u))


def log_expected_improvement(mean, var, best_f):
    sigma = var.clamp_min(MIN_VAR).sqrt()
    u = (mean - best_f) / sigma
    return _log_ei_helper(u) + torch.log(sigma)


def upper_confidence_bound(mean, var, beta):
    sigma = var.clamp_min(MIN_VAR).sqrt()
    return mean + math.sqrt(beta) * sigma


ACQ_EI, ACQ_LOGEI, ACQ_UCB = 0, 1, 2


def acquisition(mean, var, acq_kind, param):
    if acq_kind == ACQ_EI:
        return expected_improvement(mean, var, param)
    if acq_kind == ACQ_LOGEI:
        return log_expected_improvement(mean, var, param)
    if acq_kind == ACQ_UCB:
        return upper_confidence_bound(mean, var, param)
    raise ValueError(acq_kind)


# ----------------------------------------------------------------------------------------
# A6  acquisition value + gradient w.r.t. the candidate  (autograd of A5 o A4 o A1 upstream)
# ----------------------------------------------------------------------------------------
def acquisition_value_and_grad_autograd(state: GPState, Z, acq_kind, param):
    Z = Z.detach().clone().requires_grad_(True)
    mean, var = posterior(state, Z)
    a = acquisition(mean, var, acq_kind, param)
    (g,) = torch.autograd.grad(a.sum(), Z)
    return a.detach(), g


def ei_value_and_grad_analytic(state: GPState, Z, best_f):
    """Closed form of SURVEY.md Appendix C (A6): beta = L^-T v, c_j = (Phi(u) alpha_j - phi(u)/sigma beta_j) dk/dr^2,
    grad_z = (2/l^2) o (z sum_j c_j - sum_j c_j x_j); zero variance-gradient where var was clamped."""
    r2 = scaled_sq_dist(Z, state.X, state.lengthscale, state.center)
    Ks = state.outputscale * kernel_from_sq_dist(r2, state.kind)
    dK = state.outputscale * dkernel_dsq_dist(r2, state.kind)
    mean = state.mean_const + Ks @ state.alpha
    V = Ks @ state.Linv.T
    var = state.outputscale - (V * V).sum(-1)
    clamped = var < MIN_VAR
    sigma = var.clamp_min(MIN_VAR).sqrt()
    u = (mean - best_f) / sigma
    ei = sigma * (_phi(u) + u * _Phi(u))
    beta = V @ state.Linv
    dsig = torch.where(clamped, torch.zeros_like(sigma), _phi(u) / sigma)
    C = (_Phi(u).unsqueeze(-1) * state.alpha.unsqueeze(0) - dsig.unsqueeze(-1) * beta) * dK
    Zc = Z - state.center
    Xc = state.X - state.center
    grad = (2.0 / state.lengthscale ** 2) * (Zc * C.sum(-1, keepdim=True) - C @ Xc)
    return ei, grad


# ----------------------------------------------------------------------------------------
# A7  qExpectedImprovement  (botorch.acquisition.monte_carlo + SobolQMCNormalSampler base samples)
# ----------------------------------------------------------------------------------------
def qei_from_posterior(mean, Sigma, base_samples, best_f, jitter=1e-8):
    """mean [b,q], Sigma [b,q,q], base_samples [S,q] -> qEI [b] = mean_s max_j relu(mu + L z_s - f*)."""
    q = mean.shape[-1]
    Lq = torch.linalg.cholesky(Sigma + jitter * torch.eye(q, dtype=Sigma.dtype))
    samples = mean.unsqueeze(-2) + base_samples.unsqueeze(0) @ Lq.transpose(-1, -2)  # [b,S,q]
    return (samples - best_f).clamp_min(0.0).amax(dim=-1).mean(dim=-1)


def qei_value_and_grad_autograd(state: GPState, Zq, base_samples, best_f, jitter=1e-8):
    Zq = Zq.detach().clone().requires_grad_(True)
    mean, Sigma = posterior_joint(state, Zq)
    v = qei_from_posterior(mean, Sigma, base_samples, best_f, jitter)
    (g,) = torch.autograd.grad(v.sum(), Zq)
    return v.detach(), g


# ----------------------------------------------------------------------------------------
# A10  batched (fully Bayesian / SAAS) posterior: S hyper-parameter samples, mixture moments
# ----------------------------------------------------------------------------------------
def saas_fit_states(X, y, lengthscales, outputscales, noises, means, kind=MATERN52):
    return [fit_state(X, y, lengthscales[s], float(outputscales[s]), float(noises[s]), float(means[s]), kind)
            for s in range(lengthscales.shape[0])]


def saas_posterior(states, Z):
    """Per-sample mean/var [S,m] and the Gaussian-mixture mean / variance [m]
    (botorch GaussianMixturePosterior: mixture_mean = mean_s mu_s; mixture_var = mean_s(var_s + mu_s^2) - mixture_mean^2)."""
    mus, vs = zip(*[posterior(st, Z) for st in states])
    mu = torch.stack(mus)
    var = torch.stack(vs)
    mmean = mu.mean(0)
    mvar = (var + mu * mu).mean(0) - mmean * mmean
    return mu, var, mmean, mvar


# ----------------------------------------------------------------------------------------
# synthetic workloads of BASELINE.md §2 / SURVEY.md §8d (seeded; shared by tests and bench.py)
# ----------------------------------------------------------------------------------------
def synthetic_problem(n, d, ls_mode="prior", noise=1e-3, seed=0):
    """X ~ U[0,1]^{n x d} (seed), y = standardise(sin(3 sum_{k<4} x_k) + 0.1 eps) (seed+1),
    l_k = sqrt(d) exp(0.25 xi_k) (seed+2) or the stress sets 1.5 sqrt(d) / 0.15 sqrt(d)."""
    g0 = torch.Generator().manual_seed(seed)
    g1 = torch.Generator().manual_seed(seed + 1)
    g2 = torch.Generator().manual_seed(seed + 2)
    X = torch.rand(n, d, generator=g0, dtype=torch.float64)
    y = torch.sin(3.0 * X[:, : min(4, d)].sum(-1)) + 0.1 * torch.randn(n, generator=g1, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    if ls_mode == "prior":
        ls = math.sqrt(d) * torch.exp(0.25 * torch.randn(d, generator=g2, dtype=torch.float64))
    elif ls_mode == "long":
        ls = torch.full((d,), 1.5 * math.sqrt(d), dtype=torch.float64)
    elif ls_mode == "short":
        ls = torch.full((d,), 0.15 * math.sqrt(d), dtype=torch.float64)
    else:
        raise ValueError(ls_mode)
    return X, y, ls, 1.0, noise, 0.0


def sobol_candidates(m, d, seed=3, skip=0):
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    if skip:
        eng.fast_forward(skip)
    return eng.draw(m, dtype=torch.float64)


# ----------------------------------------------------------------------------------------
# SAAS NUTS potential: what Pyro's NUTS differentiates ~1e4-1e5 times per `saas_bo` fit
# (solver selected at /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:183-194).
# Model of botorch.models.fully_bayesian.SaasPyroModel [UPSTREAM]: mean ~ N(0,1); outputscale ~ Gamma(2, 0.15);
# noise ~ Gamma(0.9, 10); tausq ~ HalfCauchy(0.1); inv_length_sq_k ~ HalfCauchy(1); lengthscale_k = (tausq inv_length_sq_k)^-1/2;
# Matern-5/2.  Positive sites are sampled in log space (ExpTransform): potential U(z) = -[log p(y | theta(z)) +
# sum log prior(theta(z)) + sum log |d theta / d z|].  z = [log inv_length_sq (d), log tausq, log outputscale, log noise, mean].
# ----------------------------------------------------------------------------------------
def saas_potential(X, y, z, min_noise=0.0):
    """z [d + 4] (differentiable) -> scalar potential energy."""
    d = X.shape[1]
    n = X.shape[0]
    z_l, z_t, z_o, z_n, mean = z[:d], z[d], z[d + 1], z[d + 2], z[d + 3]
    ils, tausq, os_, noise = torch.exp(z_l), torch.exp(z_t), torch.exp(z_o), torch.exp(z_n)
    ls = torch.rsqrt(tausq * ils)
    ll = exact_mll(X, y, ls, os_, noise + min_noise, mean, MATERN52) * n
    def half_cauchy(v, s):
        return math.log(2.0 / math.pi) - math.log(s) - torch.log1p((v / s) ** 2)
    lp = gamma_log_prob(os_, 2.0, 0.15) + gamma_log_prob(noise, 0.9, 10.0) + half_cauchy(tausq, 0.1) + half_cauchy(ils, 1.0).sum()
    lp = lp - 0.5 * mean * mean - 0.5 * LOG_2PI
    jac = z_l.sum() + z_t + z_o + z_n
    return -(ll + lp + jac)
