"""GPU tests of the callers either side of the kernels: the fully Bayesian (SAAS) batched model, the eval-mode MLL of
the reference's trainer, `train_exact_gp_model` / `filter_close_points`, the solver `step()` contract and the
botorch / gpytorch import shims."""
import math
import sys

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-6


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def test_saas_batched_posterior_matches_per_sample_oracle(cuda_device):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.models import SaasFullyBayesianSingleTaskGP, sample_saas_prior

    n, d, S, m = 96, 40, 6, 37
    X, y, *_ = O.synthetic_problem(n, d)
    samples = sample_saas_prior(S, d, seed=4)
    model = SaasFullyBayesianSingleTaskGP(X, y.unsqueeze(-1))
    with pytest.raises(RuntimeError):
        model.posterior(X[:2].unsqueeze(1))
    model.load_mcmc_samples(samples)
    assert model.batch_shape == torch.Size([S]) and model.num_mcmc_samples == S
    states = O.saas_fit_states(X, y, samples["lengthscale"], samples["outputscale"], samples["noise"], samples["mean"])
    Z = O.sobol_candidates(m, d, seed=9)
    mu_o, var_o, mmean_o, mvar_o = O.saas_posterior(states, Z)
    post = model.posterior(Z.unsqueeze(1))
    assert post.mean.shape == (m, S, 1, 1)  # MCMC dimension at -3
    assert _rel(post.mean.squeeze(-1).squeeze(-1).T, mu_o) < REL
    assert _rel(post.variance.squeeze(-1).squeeze(-1).T, var_o) < REL
    assert _rel(post.mixture_mean.reshape(-1), mmean_o) < REL and _rel(post.mixture_variance.reshape(-1), mvar_o) < REL
    # analytic acquisition averages over the ensemble (logmeanexp for LogEI); gradient flows through every member
    best_f = float(y.max())
    ei = A.ExpectedImprovement(model, best_f=best_f)(Z.unsqueeze(1))
    assert _rel(ei, O.expected_improvement(mu_o, var_o, best_f).mean(0)) < REL
    lei_o = torch.logsumexp(O.log_expected_improvement(mu_o, var_o, best_f), 0) - math.log(S)
    Zg = Z.clone().unsqueeze(1).requires_grad_(True)
    lei = A.LogExpectedImprovement(model, best_f=best_f)(Zg)
    assert float((lei.detach() - lei_o).abs().max()) < REL * max(1.0, float(lei_o.abs().max()))
    (g,) = torch.autograd.grad(lei.sum(), Zg)
    Zo = Z.clone().requires_grad_(True)
    vals = torch.stack([O.log_expected_improvement(*O.posterior(st, Zo), best_f) for st in states])
    (g_o,) = torch.autograd.grad((torch.logsumexp(vals, 0) - math.log(S)).sum(), Zo)
    assert _rel(g.squeeze(1), g_o) < 1e-5


def test_batched_mll_matches_per_sample_oracle(cuda_device):
    """Config #4 shape family: S MLL values + gradients from one batched build / Cholesky."""
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import mll_gradient, sample_saas_prior

    n, d, S = 130, 24, 5
    X, y, *_ = O.synthetic_problem(n, d)
    smp = sample_saas_prior(S, d, seed=2)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, batch=S, keep_r2=True)
    gp.set_hyper(smp["lengthscale"].to(cuda_device), smp["outputscale"], smp["noise"], smp["mean"])
    gp.fit()
    grad = mll_gradient(gp).cpu()
    for s in range(S):
        v, g_ls, g_os, g_nz, g_mu = O.exact_mll_and_grad(X, y, smp["lengthscale"][s], float(smp["outputscale"][s]),
                                                         float(smp["noise"][s]), float(smp["mean"][s]), O.MATERN52)
        assert abs(gp.mll_data_term[s].item() / n - v.item()) < REL * abs(v.item())
        assert float((grad[s, :d] / n - g_ls).abs().max()) < REL * max(float(g_ls.abs().max()), 1e-6)
        assert abs(grad[s, d] / n - g_os) < REL * max(abs(float(g_os)), 1e-3)
        assert abs(grad[s, d + 1] / n - g_nz) < REL * max(abs(float(g_nz)), 1e-3)


def _plain_model(X, y, ls, os_, nz, mu):
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP

    d = X.shape[1]
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    return model


def test_eval_mode_mll_is_the_joint_predictive_log_density(cuda_device):
    from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood

    X, y, ls, os_, nz, mu = O.synthetic_problem(150, 10, "prior", 1e-2)
    os_, mu = 1.2, 0.1
    model = _plain_model(X, y, ls, os_, nz, mu)
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    g = torch.Generator().manual_seed(5)
    Xt = torch.rand(70, 10, generator=g, dtype=torch.float64)
    yt = torch.sin(3 * Xt[:, :4].sum(-1))
    model.eval(); mll.eval()
    with torch.no_grad():
        got = mll(model(Xt), yt)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    mean, Sigma = O.posterior_joint(st, Xt.unsqueeze(0), observation_noise=True)
    want = torch.distributions.MultivariateNormal(mean[0], covariance_matrix=Sigma[0]).log_prob(yt) / 70
    assert abs(float(got) - float(want)) < REL * abs(float(want))


def test_reference_trainer_protocol(cuda_device, tmp_path):
    from hdbo_benchmark_b200.distances import filter_close_points, pairwise_distances
    from harness.adam_trainer import fit_with_adam

    g = torch.Generator().manual_seed(0)
    x = torch.rand(60, 5, generator=g, dtype=torch.float64)
    x[10] = x[3] + 1e-4
    x[40] = x[3] - 1e-4
    yv = torch.arange(60, dtype=torch.float64)
    D = pairwise_distances(x)
    assert _rel(D, torch.cdist(x, x)) < 1e-7 and float(D.diagonal().abs().max()) == 0.0
    xf, yf = filter_close_points(x, yv, tolerance=1e-2)
    keep = [i for i in range(60) if i not in (10, 40)]
    assert torch.equal(yf, yv[keep]) and torch.equal(xf, x[keep])
    # Adam + early stopping loop of the reference on a small problem
    X, y, ls, os_, nz, mu = O.synthetic_problem(80, 5, "prior", 1e-2)
    model = _plain_model(X[:60], y[:60], torch.ones(5), 1.0, 0.05, 0.0)
    ckpt = tmp_path / "model_model_npoints_60_ndim_5.pt"
    it, diverged = fit_with_adam(model, (X[:60], y[:60]), (X[60:], y[60:]), patience=5, max_iterations=25, checkpoint=ckpt)
    assert not diverged and it >= 5 and not model.training
    assert ckpt.exists() and set(torch.load(ckpt)) == set(model.state_dict())


def test_solver_step_contract_runs_end_to_end(cuda_device):
    from hdbo_benchmark_b200 import gp_modules as G
    from harness.poli_solvers import LineBO, VanillaBayesianOptimization

    d = 4
    opt = np.full(d, 0.3)

    def black_box(x):
        return -np.sum((np.asarray(x) - opt) ** 2, axis=-1, keepdims=True)

    rng = np.random.default_rng(0)
    x0 = rng.random((8, d))
    y0 = black_box(x0)
    torch.manual_seed(0)
    kernel = G.ScaleKernel(G.RBFKernel(ard_num_dims=d, lengthscale_prior=G.LogNormalPrior(math.log(d) / 2, 1.0)))  # load_solvers.py:131-138
    solver = VanillaBayesianOptimization(black_box, x0, y0, bounds=(0.0, 1.0), kernel=kernel, num_restarts=4, raw_samples=64, seed=0)
    solver.solve(max_iter=4)
    assert solver.iteration == 4 and len(solver.history["x"]) == 12
    assert solver.get_best_performance() >= float(y0.max())
    for kind in ("random", "coordinate"):
        lb = LineBO(black_box, x0, y0, bounds=(0.0, 1.0), type_of_line=kind, seed=1)
        x, y = lb.step()
        assert x.shape == (1, d) and (x >= 0).all() and (x <= 1).all() and lb.iteration == 1


def test_import_shims_resolve_reference_style_code(cuda_device):
    from hdbo_benchmark_b200 import shims

    installed = shims.install()
    try:
        import botorch  # noqa: F401
        import gpytorch
        from botorch.acquisition import LogExpectedImprovement
        from botorch.fit import fit_gpytorch_mll  # noqa: F401
        from botorch.models import SingleTaskGP
        from botorch.optim import optimize_acqf  # noqa: F401

        n_dimensions = 6
        kernel = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(
            ard_num_dims=n_dimensions, lengthscale_prior=gpytorch.priors.LogNormalPrior(np.log(n_dimensions) / 2, 1.0)))
        X, y, *_ = O.synthetic_problem(30, n_dimensions)
        model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=kernel)
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(model.likelihood, model)
        model.train()
        assert torch.isfinite(mll(model(X), y))
        model.eval()
        assert LogExpectedImprovement(model, best_f=float(y.max()))(X[:3].unsqueeze(1)).shape == (3,)
    finally:
        if installed:
            for k in [k for k, v in sys.modules.items() if getattr(v, "__hdbo_b200_shim__", False)]:
                del sys.modules[k]


def test_joint_posterior_samples_match_oracle_and_thompson_picks(cuda_device):
    """joint posterior samples over a candidate set (m x m covariance, its Cholesky factor, mean + L z) against the
    oracle with the same base samples and jitter; MaxPosteriorSampling returns the arg-max candidates of each sample."""
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP
    from hdbo_benchmark_b200.thompson import MaxPosteriorSampling, joint_posterior_samples

    n, d, m, S = 150, 6, 333, 5
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "short", 1e-2, seed=3)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(m, d, seed=8)
    z = torch.randn(S, m, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = float(os_)
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    gp = model._device_gp()
    samples, mean, jitter = joint_posterior_samples(gp, Z.to(cuda_device), z)
    mean_o, Sigma_o = O.posterior_joint(st, Z.unsqueeze(0))
    Lo = torch.linalg.cholesky(Sigma_o[0] + jitter * torch.eye(m, dtype=torch.float64))
    samples_o = mean_o[0] + z @ Lo.T
    assert float((mean.cpu() - mean_o[0]).abs().max()) < 1e-9
    # the factor of a (nearly low-rank) covariance amplifies rounding: compare at 1e-6 of the sample scale
    assert float((samples.cpu() - samples_o).abs().max()) < 1e-6 * float(samples_o.abs().max())
    picker = MaxPosteriorSampling(model, replacement=False, seed=1)
    Xn = picker(Z, num_samples=4)
    assert Xn.shape == (4, d)
    rows = {tuple(r.tolist()) for r in Xn}
    assert len(rows) == 4 and all(any(torch.equal(r, zc) for zc in Z) for r in Xn)


def test_saas_nuts_potential_and_gradient_match_oracle_autograd(cuda_device):
    """the potential energy Pyro's NUTS differentiates for `saas_bo`, batched over chain states."""
    from hdbo_benchmark_b200.saas import SaasPotential

    n, d, Cn = 70, 9, 5
    X, y, *_ = O.synthetic_problem(n, d)
    g = torch.Generator().manual_seed(3)
    Z = torch.randn(Cn, d + 4, generator=g, dtype=torch.float64) * 0.7
    Z[:, d] -= 2.0          # tausq around exp(-2)
    Z[:, d + 2] -= 3.0      # noise around exp(-3)
    U, dU = SaasPotential(X, y)(Z)
    for c in range(Cn):
        z = Z[c].clone().requires_grad_(True)
        u = O.saas_potential(X, y, z)
        (gz,) = torch.autograd.grad(u, z)
        assert abs(float(U[c]) - float(u.detach())) < REL * max(1.0, abs(float(u.detach())))
        assert float((dU[c].cpu() - gz).abs().max()) < REL * max(1.0, float(gz.abs().max()))
