"""GPU tests of the differentiable paths and of the BoTorch-shaped surface (model.posterior, acqf(X) with autograd,
fit_gpytorch_mll, optimize_acqf, qExpectedImprovement) against the CPU oracle."""
import math

import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-6


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _legacy_model(dev, n=200, d=12, kind=O.MATERN52, seed=0):
    """SingleTaskGP (Matern/RBF + ScaleKernel, no outcome transform) with fixed hyper-parameters + the oracle state."""
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-3, seed=seed)
    os_, mu = 1.4, 0.1
    base = (G.MaternKernel if kind == O.MATERN52 else G.RBFKernel)(ard_num_dims=d)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(base), likelihood=G.GaussianLikelihood(),
                         outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    return model, st, (X, y, ls, os_, nz, mu)


@pytest.mark.parametrize("kind", [O.MATERN52, O.RBF])
def test_posterior_surface_shapes_and_values(cuda_device, kind):
    model, st, (X, y, *_r) = _legacy_model(cuda_device, kind=kind)
    Z = O.sobol_candidates(50, X.shape[1])
    post = model.posterior(Z.unsqueeze(1))  # b x 1 x d, host tensor in -> host tensor out
    assert post.mean.shape == (50, 1, 1) and post.variance.shape == (50, 1, 1) and post.mean.device.type == "cpu"
    m_o, v_o = O.posterior(st, Z)
    assert _rel(post.mean.reshape(-1), m_o) < REL and _rel(post.variance.reshape(-1), v_o) < REL
    # joint q-batch covariance
    Zq = Z[:48].reshape(12, 4, -1)
    pq = model.posterior(Zq)
    mq_o, S_o = O.posterior_joint(st, Zq)
    assert pq.mean.shape == (12, 4, 1)
    assert _rel(pq.mean.squeeze(-1), mq_o) < REL
    assert _rel(pq.covariance_matrix, S_o) < REL
    s = pq.rsample(torch.Size([16]))
    assert s.shape == (16, 12, 4, 1) and torch.isfinite(s).all()


@pytest.mark.parametrize("kind", [O.MATERN52, O.RBF])
@pytest.mark.parametrize("acq", ["ei", "logei", "ucb"])
def test_analytic_acquisition_value_and_gradient(cuda_device, kind, acq):
    from hdbo_benchmark_b200 import acquisition as A

    model, st, (X, y, *_r) = _legacy_model(cuda_device, kind=kind)
    best_f = float(y.max()) - 0.3
    d = X.shape[1]
    Z = O.sobol_candidates(40, d, seed=7)
    if acq == "ei":
        f, kind_id, prm = A.ExpectedImprovement(model, best_f=best_f), O.ACQ_EI, best_f
    elif acq == "logei":
        f, kind_id, prm = A.LogExpectedImprovement(model, best_f=best_f), O.ACQ_LOGEI, best_f
    else:
        f, kind_id, prm = A.UpperConfidenceBound(model, beta=3.0), O.ACQ_UCB, 3.0
    a_o, g_o = O.acquisition_value_and_grad_autograd(st, Z, kind_id, prm)
    with torch.no_grad():
        a_fused = f(Z.unsqueeze(1))
    Zg = Z.clone().unsqueeze(1).requires_grad_(True)
    a = f(Zg)
    (g,) = torch.autograd.grad(a.sum(), Zg)
    assert a.shape == (40,) and g.shape == (40, 1, d)
    tol = REL * max(1.0, float(a_o.abs().max()))
    assert float((a.detach() - a_o).abs().max()) < tol and float((a_fused - a_o).abs().max()) < tol
    assert _rel(g.squeeze(1), g_o) < REL


def test_minimize_flag_and_standardize_equivariance(cuda_device):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.models import SingleTaskGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(150, 8, "prior", 1e-3)
    y_raw = 5.0 + 3.0 * y
    model = SingleTaskGP(X, y_raw.unsqueeze(-1))  # default: Standardize outcome transform, RBF ARD
    model.covar_module.lengthscale = ls
    model.likelihood.noise = nz
    model.eval()
    ys = (y_raw - y_raw.mean()) / y_raw.std()
    st = O.fit_state(X, ys, ls, 1.0, nz, 0.0, O.RBF)
    Z = O.sobol_candidates(30, 8, seed=11)
    m_o, v_o = O.posterior(st, Z)
    s, sh = float(y_raw.std()), float(y_raw.mean())
    post = model.posterior(Z.unsqueeze(1))
    assert _rel(post.mean.reshape(-1), m_o * s + sh) < REL and _rel(post.variance.reshape(-1), v_o * s * s) < REL
    best = float(y_raw.max())
    ei = A.ExpectedImprovement(model, best_f=best)(Z.unsqueeze(1))
    assert _rel(ei, O.expected_improvement(m_o * s + sh, v_o * s * s, best)) < 1e-5
    lei = A.LogExpectedImprovement(model, best_f=best)(Z.unsqueeze(1))
    assert float((lei - O.log_expected_improvement(m_o * s + sh, v_o * s * s, best)).abs().max()) < 1e-6 * 50
    ucb = A.UpperConfidenceBound(model, beta=2.0)(Z.unsqueeze(1))
    assert _rel(ucb, O.upper_confidence_bound(m_o * s + sh, v_o * s * s, 2.0)) < REL
    ei_min = A.ExpectedImprovement(model, best_f=float(y_raw.min()), maximize=False)(Z.unsqueeze(1))
    want = O.expected_improvement(-(m_o * s + sh), v_o * s * s, -float(y_raw.min()))
    assert _rel(ei_min, want) < 1e-5


@pytest.mark.parametrize("q", [1, 2, 4])
def test_qei_value_and_gradient(cuda_device, q):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.sampling import SobolQMCNormalSampler

    model, st, (X, y, *_r) = _legacy_model(cuda_device, n=160, d=10)
    d = X.shape[1]
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([256]), seed=6)
    best_f = float(y.mean())
    f = A.qExpectedImprovement(model, best_f=best_f, sampler=sampler)
    Zq = O.sobol_candidates(6 * q, d, seed=5).reshape(6, q, d)
    base = sampler.base_samples(q, cuda_device).cpu()
    v_o, g_o = O.qei_value_and_grad_autograd(st, Zq, base, best_f, jitter=1e-8)
    Zg = Zq.clone().requires_grad_(True)
    v = f(Zg)
    (g,) = torch.autograd.grad(v.sum(), Zg)
    assert _rel(v, v_o) < REL
    assert _rel(g, g_o) < 1e-5


def test_mll_surface_matches_oracle_and_adam_loop_runs(cuda_device):
    """ExactMarginalLogLikelihood(model(train_x), y).backward() as in the reference's training_gps.py:53-74."""
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(180, 9, "prior", 1e-2)
    d = X.shape[1]
    ls_prior = G.LogNormalPrior(math.log(d) / 2, 1.0)  # the reference's kernel: load_solvers.py:131-138
    kernel = G.ScaleKernel(G.RBFKernel(ard_num_dims=d, lengthscale_prior=ls_prior))
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=kernel, likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = 1.3
    model.likelihood.noise = nz
    model.mean_module.constant = 0.2
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    model.train()
    loss = -mll(model(model.train_inputs[0]), model.train_targets)
    loss.backward()
    # oracle: same quantity with torch autograd through the softplus transforms and the prior
    raw_ls = model.covar_module.base_kernel.raw_lengthscale.detach().cpu().clone().requires_grad_(True)
    raw_os = model.covar_module.raw_outputscale.detach().cpu().clone().requires_grad_(True)
    raw_nz = model.likelihood.noise_covar.raw_noise.detach().cpu().clone().requires_grad_(True)
    raw_mu = model.mean_module.raw_constant.detach().cpu().clone().requires_grad_(True)
    sp = torch.nn.functional.softplus
    ls_t = sp(raw_ls).reshape(-1)
    lp = O.lognormal_log_prob(ls_t, math.log(d) / 2, 1.0) if False else (
        -torch.log(ls_t) - math.log(1.0) - 0.5 * O.LOG_2PI - 0.5 * ((torch.log(ls_t) - math.log(d) / 2) / 1.0) ** 2).sum()
    want = -O.exact_mll(X, y, ls_t, sp(raw_os), sp(raw_nz).reshape(()) + 1e-4, raw_mu, O.RBF, log_prior=lp)
    want.backward()
    assert abs(loss.item() - want.item()) < REL * abs(want.item())
    assert _rel(model.covar_module.base_kernel.raw_lengthscale.grad, raw_ls.grad) < REL
    assert _rel(model.covar_module.raw_outputscale.grad, raw_os.grad) < REL
    assert _rel(model.likelihood.noise_covar.raw_noise.grad, raw_nz.grad) < REL
    assert _rel(model.mean_module.raw_constant.grad, raw_mu.grad) < REL
    # Adam loop (reference trainer) decreases the loss
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    first = None
    for it in range(15):
        opt.zero_grad()
        loss = -mll(model(model.train_inputs[0]), model.train_targets)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < first
    sd = model.state_dict()
    for key in ("covar_module.base_kernel.raw_lengthscale", "covar_module.raw_outputscale",
                "likelihood.noise_covar.raw_noise", "mean_module.raw_constant"):
        assert key in sd


def test_fit_gpytorch_mll_and_optimize_acqf_end_to_end(cuda_device):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.fit import fit_gpytorch_mll
    from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
    from hdbo_benchmark_b200.optim import optimize_acqf

    torch.manual_seed(0)
    d = 6
    X = torch.rand(40, d, dtype=torch.float64)
    y = -((X - 0.4) ** 2).sum(-1, keepdim=True) + 0.01 * torch.randn(40, 1, dtype=torch.float64)
    model = SingleTaskGP(X, y)
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    model.train()
    before = mll(model(model.train_inputs[0]), model.train_targets).item()
    fit_gpytorch_mll(mll)
    model.train()
    after = mll(model(model.train_inputs[0]), model.train_targets).item()
    model.eval()
    assert after > before
    # the fitted hyper-parameters reproduce the oracle's MLL and posterior
    ys = model.train_targets.cpu()
    ls = model.covar_module.lengthscale.detach().cpu().reshape(-1)
    nz = float(model.likelihood.noise.detach().cpu())
    mu = float(model.mean_module.constant.detach().cpu())
    lp = O.lognormal_log_prob(ls, math.sqrt(2) + 0.5 * math.log(d), math.sqrt(3)) + O.lognormal_log_prob(torch.tensor(nz), -4.0, 1.0)
    want = O.exact_mll(X, ys, ls, 1.0, nz, mu, O.RBF, log_prior=lp).item()
    assert abs(after - want) < REL * abs(want)
    acqf = A.LogExpectedImprovement(model, best_f=float(y.max()))
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).double()
    cand, val = optimize_acqf(acqf, bounds, q=1, num_restarts=8, raw_samples=256, options={"seed": 1})
    assert cand.shape == (1, d) and (cand >= 0).all() and (cand <= 1).all()
    # the optimiser's answer is at least as good as the best raw sample, evaluated by the oracle
    st = O.fit_state(X, ys, ls, 1.0, nz, mu, O.RBF)
    s, sh = float(y.std()), float(y.mean())
    def oracle_logei(Z):
        m, v = O.posterior(st, Z)
        return O.log_expected_improvement(m * s + sh, v * s * s, float(y.max()))
    from hdbo_benchmark_b200.sampling import draw_sobol_samples
    raw = draw_sobol_samples(bounds, 256, 1, seed=1).squeeze(1)
    assert oracle_logei(cand)[0] >= oracle_logei(raw).max() - 1e-9
    assert abs(float(val) - float(oracle_logei(cand)[0])) < 1e-6 * max(1.0, abs(float(val)))


@pytest.mark.parametrize("flavour", ["botorch_default", "legacy", "in_repo_rbf_scale", "isotropic"])
def test_fast_mll_closure_matches_generic_closure(cuda_device, flavour):
    """fit.py's fast closure (numpy transforms/priors + CUDA-graph replay of fit + gradient) against the generic autograd closure:
    same -mll and raw-parameter gradient at several points, and fit_gpytorch_mll reaches the same optimum either way."""
    import copy

    import numpy as np

    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.fit import _FastClosure, _bounds_for, _named_raw_parameters, fit_gpytorch_mll
    from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP

    n, d = 90, 7
    X, y, *_ = O.synthetic_problem(n, d, "prior", 1e-3, seed=6)
    Y = 2.0 * y.unsqueeze(-1) - 1.0

    def make():
        if flavour == "botorch_default":
            return SingleTaskGP(X, Y)
        if flavour == "legacy":
            return SingleTaskGP(X, Y, legacy_defaults=True)
        if flavour == "in_repo_rbf_scale":     # /root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:131-138
            k = G.ScaleKernel(G.RBFKernel(ard_num_dims=d, lengthscale_prior=G.LogNormalPrior(0.5 * math.log(d), 1.0)))
            return SingleTaskGP(X, Y, covar_module=k)
        return SingleTaskGP(X, Y, covar_module=G.ScaleKernel(G.MaternKernel(lengthscale_prior=G.GammaPrior(3.0, 6.0)),
                                                             outputscale_prior=G.GammaPrior(2.0, 0.15)))

    model = make()
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    model.train()
    params = _named_raw_parameters(mll)
    fast = _FastClosure.build(mll, params)
    assert fast is not None
    x0 = np.concatenate([p.detach().cpu().numpy().reshape(-1) for _, p in params])
    rng = np.random.default_rng(0)
    lo = np.array([(-np.inf if b[0] is None else b[0] + 1e-3) for b in _bounds_for(mll, params)])
    hi = np.array([(np.inf if b[1] is None else b[1] - 1e-3) for b in _bounds_for(mll, params)])
    for k in range(4):
        x = np.clip(x0 + 0.3 * k * rng.standard_normal(x0.shape), lo, hi)   # (untransformed parameters are box-bounded)
        f_fast, g_fast = fast(x)
        off = 0
        with torch.no_grad():
            for _, p in params:
                p.copy_(torch.from_numpy(x[off:off + p.numel()]).view_as(p).to(p.device)); off += p.numel()
        for _, p in params:
            p.grad = None
        loss = -mll(model(model.train_inputs[0]), model.train_targets).sum()
        loss.backward()
        g_ref = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for _, p in params])
        assert abs(f_fast - float(loss)) < 1e-9 * max(1.0, abs(float(loss)))
        assert np.abs(g_fast - g_ref).max() < 1e-8 * max(1.0, np.abs(g_ref).max())
    torch.manual_seed(0)
    m_fast, m_slow = make(), make()
    fit_gpytorch_mll(ExactMarginalLogLikelihood(m_fast.likelihood, m_fast))
    fit_gpytorch_mll(ExactMarginalLogLikelihood(m_slow.likelihood, m_slow), optimizer_kwargs={"options": {"fast_closure": False}})
    # both runs stop inside the same ftol-flat neighbourhood (tiny rounding differences change the L-BFGS path): equal optimum
    # value to 1e-5, parameters to a few 1e-3
    vals = []
    for mdl in (m_fast, m_slow):
        mdl.train()
        mm = ExactMarginalLogLikelihood(mdl.likelihood, mdl)
        with torch.no_grad():
            vals.append(float(mm(mdl(mdl.train_inputs[0]), mdl.train_targets).sum()))
    assert abs(vals[0] - vals[1]) < 1e-5 * max(1.0, abs(vals[1])), vals
    for (na, pa), (nb, pb) in zip(m_fast.named_parameters(), m_slow.named_parameters()):
        assert na == nb and float((pa - pb).abs().max()) < 2e-2 * max(1.0, float(pb.abs().max())), na


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,m", [(1024, 64, 48), (1500, 40, 300), (640, 24, 5)])
def test_split_k_gradient_path_matches_oracle_autograd(cuda_device, n, d, m):
    """Few candidates against one large GP: the two triangular products of the forward-with-state / backward pair run with the long
    half of their columns cut at K/2 (KR_LE_N_SPLIT / KR_GE_N_SPLIT, second parts added back by k_row_sumsq / k_make_C) and the dZ
    product splits K over the batch axis.  n_pad = 1024 and 1536 take the split, 640 (an odd number of 128-tiles) does not."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.fit()
    Z = O.sobol_candidates(m, d, seed=5)
    Z[: min(3, m)] = X[: min(3, m)] + 1e-3                     # near training points: the variance term matters
    best_f = float(y.max())
    a_o, g_o = O.acquisition_value_and_grad_autograd(st, Z, O.ACQ_LOGEI, best_f)
    bufs = {}
    for _ in range(2):                                          # twice: the scratch of the split products is reused
        a, g = gp.acq_value_grad(Z.to(cuda_device).contiguous(), _lib.ACQ_LOGEI, best_f, 1.0, bufs)
    assert float((a.cpu() - a_o).abs().max()) < REL * max(1.0, float(a_o.abs().max()))
    assert _rel(g, g_o) < REL
    # the forward state V kept for joint covariances is whole (second parts folded in): variance from V equals the oracle's
    mo, vo = O.posterior(st, Z)
    assert float(((bufs["var"].cpu() - vo).abs() / vo.abs().clamp_min(1e-8)).max()) < REL
