"""Pins the CPU oracle (oracle/gp_oracle.py) against independent implementations installed in this image:
scikit-learn GaussianProcessRegressor, scipy.stats / scipy.special, mpmath (50 digits), torch autograd.
The reference itself pins nothing (its only test asserts "no exception":
/root/reference/test_main_entry_point.py:18,42-58), so these are the known-answer tests of the path."""
import math

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

torch.set_default_dtype(torch.float64)


def _problem(n=200, d=12, seed=0, noise=1e-2):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g)
    y = torch.sin(3 * X[:, :3].sum(-1)) + 0.1 * torch.randn(n, generator=g)
    y = (y - y.mean()) / y.std()
    ls = 0.5 + torch.rand(d, generator=g)
    return X, y, ls, 1.7, noise, 0.0


@pytest.mark.parametrize("kind", [O.MATERN52, O.RBF])
def test_mll_and_posterior_match_sklearn(kind):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

    X, y, ls, os_, noise, mean = _problem()
    base = Matern(length_scale=ls.numpy(), nu=2.5) if kind == O.MATERN52 else RBF(length_scale=ls.numpy())
    kern = ConstantKernel(os_) * base + WhiteKernel(noise)
    gpr = GaussianProcessRegressor(kernel=kern, optimizer=None, alpha=0.0).fit(X.numpy(), y.numpy())
    lml, grad = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    n = X.shape[0]
    v, g_ls, g_os, g_noise, g_mean = O.exact_mll_and_grad(X, y, ls, os_, noise, mean, kind)
    assert abs(v.item() * n - lml) <= 1e-11 * abs(lml)
    # sklearn differentiates w.r.t. log(theta), theta = [constant, length_scale..., noise]
    ours = np.concatenate([[g_os.item() * os_], (g_ls * ls).numpy(), [g_noise.item() * noise]]) * n
    np.testing.assert_allclose(ours, grad, rtol=1e-8, atol=1e-9)
    # predictive mean / std (sklearn's std includes the WhiteKernel noise)
    Z = torch.rand(64, X.shape[1], generator=torch.Generator().manual_seed(5))
    st = O.fit_state(X, y, ls, os_, noise, mean, kind)
    m_o, v_o = O.posterior(st, Z, observation_noise=True)
    m_s, s_s = gpr.predict(Z.numpy(), return_std=True)
    np.testing.assert_allclose(m_o.numpy(), m_s, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(v_o.numpy(), s_s ** 2, rtol=1e-8)


@pytest.mark.parametrize("kind", [O.MATERN52, O.RBF])
def test_mll_analytic_gradient_matches_autograd(kind):
    X, y, ls, os_, noise, mean = _problem(n=150, d=9, seed=3)
    ls_t = ls.clone().requires_grad_(True)
    os_t = torch.tensor(os_, requires_grad=True)
    nz_t = torch.tensor(noise, requires_grad=True)
    mu_t = torch.tensor(0.3, requires_grad=True)
    val = O.exact_mll(X, y, ls_t, os_t, nz_t, mu_t, kind)
    val.backward()
    v, g_ls, g_os, g_noise, g_mean = O.exact_mll_and_grad(X, y, ls, os_, noise, 0.3, kind)
    assert abs(v.item() - val.item()) < 1e-13
    torch.testing.assert_close(g_ls, ls_t.grad, rtol=1e-9, atol=1e-13)
    assert abs(g_os - os_t.grad) < 1e-12 and abs(g_noise - nz_t.grad) < 1e-10 and abs(g_mean - mu_t.grad) < 1e-12


def test_kernel_properties():
    X, y, ls, os_, noise, mean = _problem(n=60, d=7)
    for kind in (O.MATERN52, O.RBF):
        K = O.kernel_matrix(X, X, ls, os_, kind, x1_eq_x2=True)
        torch.testing.assert_close(K, K.T, rtol=0, atol=1e-14)
        assert torch.all(torch.diagonal(K) == os_)
        assert torch.linalg.eigvalsh(K).min() > -1e-10
        shift = torch.full((1, X.shape[1]), 3.25)
        K2 = O.kernel_matrix(X + shift, X + shift, ls, os_, kind, x1_eq_x2=True)
        torch.testing.assert_close(K, K2, rtol=0, atol=1e-12)  # translation invariance
        # against the direct-difference definition
        diff = (X[:, None, :] - X[None, :, :]) / ls
        r2 = (diff * diff).sum(-1)
        torch.testing.assert_close(K, os_ * O.kernel_from_sq_dist(r2, kind), rtol=0, atol=1e-12)


def test_analytic_acquisitions_match_scipy():
    from scipy.special import log_ndtr
    from scipy.stats import norm

    g = torch.Generator().manual_seed(1)
    mean = torch.randn(500, generator=g) * 2
    var = torch.rand(500, generator=g) * 2 + 1e-6
    var[:5] = 1e-15  # exercises the 1e-12 clamp
    best_f = 0.7
    sig = np.sqrt(np.maximum(var.numpy(), 1e-12))
    u = (mean.numpy() - best_f) / sig
    ei = sig * (norm.pdf(u) + u * norm.cdf(u))
    # phi(u) + u Phi(u) cancels for u << 0 (that is why LogEI exists); compare tightly where it is well conditioned
    got_ei = O.expected_improvement(mean, var, best_f).numpy()
    wc = u > -4
    np.testing.assert_allclose(got_ei[wc], ei[wc], rtol=1e-11)
    np.testing.assert_allclose(got_ei[~wc], ei[~wc], rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(O.upper_confidence_bound(mean, var, 2.5).numpy(), mean.numpy() + math.sqrt(2.5) * sig,
                               rtol=1e-14)
    ok = ei > 1e-250
    np.testing.assert_allclose(O.log_expected_improvement(mean, var, best_f).numpy()[ok], np.log(ei[ok]), rtol=1e-9)


def test_log_ei_tails_match_mpmath():
    import mpmath as mp

    mp.mp.dps = 60
    us = [-1e7, -2e6, -1e6 + 1, -3e4, -500.0, -40.0, -7.5, -1.0 - 1e-9, -1.0, -0.5, 0.0, 3.0, 30.0]
    u = torch.tensor(us)
    got = O._log_ei_helper(u).numpy()
    for ui, gi in zip(us, got):
        x = mp.mpf(ui)
        h = mp.npdf(x) + x * mp.ncdf(x)
        want = float(mp.log(h))
        assert abs(gi - want) <= 2e-9 * abs(want) + 1e-12, (ui, gi, want)


def test_small_gp_matches_mpmath():
    import mpmath as mp

    mp.mp.dps = 50
    X, y, ls, os_, noise, mean = _problem(n=24, d=3, seed=11, noise=1e-4)
    st = O.fit_state(X, y, ls, os_, noise, mean, O.MATERN52)
    Z = torch.rand(6, 3, generator=torch.Generator().manual_seed(2))
    Z[0] = X[0] + 1e-3  # near a training point: variance cancellation
    m_o, v_o = O.posterior(st, Z)

    def k(a, b):
        r2 = sum(((mp.mpf(float(ai)) - mp.mpf(float(bi))) / mp.mpf(float(li))) ** 2 for ai, bi, li in zip(a, b, ls))
        r = mp.sqrt(r2)
        return mp.mpf(os_) * (1 + mp.sqrt(5) * r + mp.mpf(5) / 3 * r2) * mp.exp(-mp.sqrt(5) * r)

    n = X.shape[0]
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = k(X[i], X[j]) + (mp.mpf(noise) if i == j else 0)
    yv = mp.matrix([mp.mpf(float(t)) for t in y])
    alpha = mp.lu_solve(K, yv)
    for c in range(Z.shape[0]):
        ks = mp.matrix([k(Z[c], X[j]) for j in range(n)])
        mu = sum(ks[j] * alpha[j] for j in range(n))
        sol = mp.lu_solve(K, ks)
        var = mp.mpf(os_) - sum(ks[j] * sol[j] for j in range(n))
        assert abs(m_o[c].item() - float(mu)) <= 1e-9 * max(1.0, abs(float(mu)))
        assert abs(v_o[c].item() - float(var)) <= 1e-8 * abs(float(var)) + 1e-13
    # log marginal likelihood
    sign_logdet = mp.log(mp.det(K))
    lml = -mp.mpf(1) / 2 * sum(yv[i] * alpha[i] for i in range(n)) - sign_logdet / 2 - mp.mpf(n) / 2 * mp.log(2 * mp.pi)
    got = O.exact_mll(X, y, ls, os_, noise, mean, O.MATERN52).item() * n
    assert abs(got - float(lml)) <= 1e-9 * abs(float(lml))


def test_ei_analytic_gradient_matches_autograd():
    X, y, ls, os_, noise, mean = _problem(n=120, d=10, seed=4, noise=1e-3)
    for kind in (O.MATERN52, O.RBF):
        st = O.fit_state(X, y, ls, os_, noise, mean, kind)
        Z = torch.rand(32, 10, generator=torch.Generator().manual_seed(9))
        best_f = float(y.max())
        a1, g1 = O.acquisition_value_and_grad_autograd(st, Z, O.ACQ_EI, best_f)
        a2, g2 = O.ei_value_and_grad_analytic(st, Z, best_f)
        torch.testing.assert_close(a1, a2, rtol=1e-10, atol=1e-300)
        torch.testing.assert_close(g1, g2, rtol=1e-8, atol=1e-14)


def test_qei_reduces_to_ei_for_q1_and_is_differentiable():
    X, y, ls, os_, noise, mean = _problem(n=80, d=6, seed=8, noise=1e-3)
    st = O.fit_state(X, y, ls, os_, noise, mean, O.MATERN52)
    Zq = torch.rand(5, 1, 6, generator=torch.Generator().manual_seed(3))
    eng = torch.quasirandom.SobolEngine(1, scramble=True, seed=6)
    base = torch.erfinv(2 * eng.draw(4096, dtype=torch.float64).clamp(1e-9, 1 - 1e-9) - 1) * math.sqrt(2)
    best_f = float(y.mean())
    mean_q, Sig = O.posterior_joint(st, Zq)
    q1 = O.qei_from_posterior(mean_q, Sig, base, best_f, jitter=0.0)
    m1, v1 = O.posterior(st, Zq[:, 0])
    torch.testing.assert_close(mean_q[:, 0], m1)
    torch.testing.assert_close(Sig[:, 0, 0], v1, rtol=1e-9, atol=1e-14)
    ei = O.expected_improvement(m1, v1, best_f)
    torch.testing.assert_close(q1, ei, rtol=5e-3, atol=1e-5)  # QMC error only
    v, g = O.qei_value_and_grad_autograd(st, torch.rand(4, 3, 6, generator=torch.Generator().manual_seed(1)),
                                         base[:256].repeat(1, 3) * torch.tensor([1.0, -1.0, 0.5]), best_f)
    assert torch.isfinite(v).all() and torch.isfinite(g).all() and g.abs().max() > 0


def test_psd_safe_cholesky_jitter_schedule():
    A = torch.ones(4, 4)  # rank one: fails without jitter
    L, jit = O.psd_safe_cholesky(A)
    assert jit == 1e-8
    torch.testing.assert_close(L @ L.T, A + jit * torch.eye(4), rtol=0, atol=1e-12)
    with pytest.raises(O.NotPSDError):
        O.psd_safe_cholesky(-torch.eye(3))


def test_priors_match_torch_distributions():
    x = torch.tensor([0.3, 1.7, 4.0])
    want = torch.distributions.Gamma(3.0, 6.0).log_prob(x).sum()
    assert abs(O.gamma_log_prob(x, 3.0, 6.0) - want) < 1e-12
    want = torch.distributions.LogNormal(1.2, 0.8).log_prob(x).sum()
    assert abs(O.lognormal_log_prob(x, 1.2, 0.8) - want) < 1e-12


def test_acquisition_properties_hypothesis():
    """EI >= 0, EI and LogEI monotone in the mean, LogEI = log EI where EI is representable, UCB monotone in both."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(mu=st.floats(-5, 5), dmu=st.floats(1e-6, 3), var=st.floats(1e-10, 25), best=st.floats(-5, 5), beta=st.floats(0, 9))
    def check(mu, dmu, var, best, beta):
        m = torch.tensor([mu, mu + dmu], dtype=torch.float64)
        v = torch.tensor([var, var], dtype=torch.float64)
        ei = O.expected_improvement(m, v, best)
        lei = O.log_expected_improvement(m, v, best)
        assert bool((ei >= 0).all()) and ei[1] >= ei[0] and lei[1] >= lei[0] - 1e-12
        if ei[0] > 1e-250:
            assert abs(float(lei[0]) - math.log(float(ei[0]))) < 1e-8 * max(1.0, abs(float(lei[0])))
        u = O.upper_confidence_bound(m, v, beta)
        u2 = O.upper_confidence_bound(m, 2 * v, beta)
        assert u[1] > u[0] and bool((u2 >= u - 1e-15).all())

    check()


def test_posterior_variance_properties_hypothesis():
    """Variance is non-negative, at most the prior variance, and ~noise-limited at the training points; invariant to a shift of
    inputs and candidates."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(seed=st.integers(0, 10_000), kind=st.sampled_from([O.MATERN52, O.RBF]), lognoise=st.floats(-4, -1))
    def check(seed, kind, lognoise):
        g = torch.Generator().manual_seed(seed)
        n, d = 30, 5
        X = torch.rand(n, d, generator=g, dtype=torch.float64)
        y = torch.randn(n, generator=g, dtype=torch.float64)
        ls = 0.5 + torch.rand(d, generator=g, dtype=torch.float64)
        os_, noise = 1.7, 10.0 ** lognoise
        st_ = O.fit_state(X, y, ls, os_, noise, 0.1, kind)
        Z = torch.rand(40, d, generator=g, dtype=torch.float64)
        m, v = O.posterior(st_, Z)
        assert bool((v >= -1e-10).all()) and bool((v <= os_ + 1e-10).all())
        _, vt = O.posterior(st_, X)
        assert bool((vt <= noise + 1e-8).all())            # var(f | y) at a training point is below the noise level
        sh = torch.full((1, d), 2.5, dtype=torch.float64)
        m2, v2 = O.posterior(O.fit_state(X + sh, y, ls, os_, noise, 0.1, kind), Z + sh)
        torch.testing.assert_close(m, m2, rtol=0, atol=1e-8)
        torch.testing.assert_close(v, v2, rtol=0, atol=1e-8)

    check()
