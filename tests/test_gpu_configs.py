"""Parity at the FULL sizes BASELINE.json names (what bench.py times is what is checked here), all through the C ABI:
  configs[2]  n=4096, d=256 pool evaluation -- the default (int8 digit-slice) mode AND the FP64 DMMA mode, each directly against
              the oracle on a 3 000-candidate slice that contains training points, same LogEI arg-max
  configs[3]  SAAS batch S=256 x n=512 x d=1024 against the per-sample oracle on 8 of the samples
  configs[4]  qEI R=512 restarts x q=4 x 512 base samples: value and gradient against oracle autograd on 64 restarts
plus the entry points added in round 2 (device Sobol pool, joint-sample kernel, caller-owned timeline, host-pool acqf path)."""
import math

import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-6   # north_star: posterior mean / variance and MLL within relative 1e-6 of the FP64 reference path


@pytest.fixture(scope="module")
def config3_oracle():
    n, d = 4096, 256
    prob = O.synthetic_problem(n, d)
    X, y, ls, os_, nz, mu = prob
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(3000, d)
    Z[:200] = X[::20][:200]          # candidates ON training points: variance ~ noise, the hardest case for relative error
    m_o, v_o = O.posterior(st, Z)
    a_o = O.log_expected_improvement(m_o, v_o, float(y.max()))
    return prob, st, Z, m_o, v_o, a_o


@pytest.mark.parametrize("mode", ["default", "i8", "f64"])
def test_config3_pool_modes_directly_against_oracle(cuda_device, config3_oracle, mode):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    (X, y, ls, os_, nz, mu), st, Z, m_o, v_o, a_o = config3_oracle
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, var_mode=None if mode == "default" else mode)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.fit()
    mean, var, lei = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, float(y.max()))
    if mode != "f64":
        assert gp.var_mode == "i8" and gp._i8_state is not None and gp.i8_flag() == 0     # the tcgen05 path really ran
    else:
        assert gp._i8_state is None
    assert float((mean[0].cpu() - m_o).abs().max()) < REL * float(m_o.abs().max())
    assert float(((var[0].cpu() - v_o).abs() / v_o).max()) < REL
    assert float((lei[0].cpu() - a_o).abs().max()) < REL * float(a_o.abs().max())
    assert int(lei[0].argmax()) == int(a_o.argmax())
    assert abs(gp.logdet.item() - st.logdet) < 1e-9 * abs(st.logdet)


def test_config4_saas_batch_fullsize_against_per_sample_oracle(cuda_device):
    """S=256 hyper-samples x n=512 x d=1024 (BASELINE configs[3]); the oracle refits 8 of the samples one by one."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import sample_saas_prior

    S, n, d, m = 256, 512, 1024, 2048
    X, y, *_ = O.synthetic_problem(n, d)
    smp = sample_saas_prior(S, d, seed=4)
    Z = O.sobol_candidates(m, d, seed=9)
    Z[:64] = X[:64]
    best = float(y.max())
    ratio = smp["noise"] / smp["outputscale"]
    redo = [int(i) for i in (ratio < 1e-4).nonzero().reshape(-1)]          # members below the int8 noise-ratio guard
    assert 1 <= len(redo) < S // 2
    picks = sorted({0, 1, 37, 128, 254, 255, redo[0], redo[-1]})
    states = O.saas_fit_states(X, y, smp["lengthscale"][picks], smp["outputscale"][picks], smp["noise"][picks], smp["mean"][picks])
    mu_o, var_o, _, _ = O.saas_posterior(states, Z)
    a_o = O.log_expected_improvement(mu_o, var_o, best)
    for mode in ("i8", "f64"):
        gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, batch=S, var_mode=mode)
        gp.set_hyper(smp["lengthscale"].to(cuda_device), smp["outputscale"], smp["noise"], smp["mean"])
        gp.fit()
        mean, var, acq = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, best)
        assert mean.shape == (S, m) and gp.i8_flag() == 0
        if mode == "i8":
            # the batch stays on the tcgen05 path; the few members whose noise / outputscale is below 1e-4 are redone in FP64
            assert gp._i8_ok and gp._i8_state is not None and gp._i8_redo == redo, gp._i8_redo
        mean, var, acq = mean[picks].cpu(), var[picks].cpu(), acq[picks].cpu()
        assert float((mean - mu_o).abs().max()) < REL * float(mu_o.abs().max())
        assert float(((var - var_o).abs() / var_o.abs()).max()) < REL
        assert float((acq - a_o).abs().max()) < REL * max(1.0, float(a_o.abs().max()))
        assert torch.equal(acq.argmax(dim=1), a_o.argmax(dim=1))
        for j, s in enumerate(picks):
            assert abs(float(gp.logdet[s]) - states[j].logdet) < 1e-9 * max(1.0, abs(states[j].logdet))
        del gp
        torch.cuda.empty_cache()


def test_config5_qei_fullsize_value_and_gradient(cuda_device):
    """R=512 restarts x q=4, 512 Sobol-normal base samples, GP of configs[1] (n=1024, d=128): one batched value + gradient call;
    64 of the restarts are checked against the oracle's autograd."""
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP
    from hdbo_benchmark_b200.sampling import SobolQMCNormalSampler

    n, d, R, q, S = 1024, 128, 512, 4, 512
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([S]), seed=6)
    best_f = float(y.max()) - 0.5      # a reachable incumbent so that most restarts have improving samples
    f = A.qExpectedImprovement(model, best_f=best_f, sampler=sampler)
    Zq = O.sobol_candidates(R * q, d, seed=5).reshape(R, q, d)
    Zg = Zq.clone().requires_grad_(True)
    v = f(Zg)
    (g,) = torch.autograd.grad(v.sum(), Zg)
    assert v.shape == (R,) and g.shape == (R, q, d)
    pick = torch.arange(0, R, 8)       # 64 restarts
    base = sampler.base_samples(q, cuda_device).cpu()
    v_o, g_o = O.qei_value_and_grad_autograd(st, Zq[pick], base, best_f, jitter=1e-8)
    assert float(v_o.max()) > 0.0
    assert float((v[pick].cpu() - v_o).abs().max()) < REL * float(v_o.abs().max())
    assert float((g[pick].cpu() - g_o).abs().max()) < 1e-5 * float(g_o.abs().max())


# ---------------------------------------------------------------------------------------------------------------------------
# entry points added in round 2
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,seed", [(1, 0), (7, 3), (256, 3), (515, 11)])
def test_device_sobol_is_bit_identical_to_torch(cuda_device, dim, seed):
    """hdbo_sobol_draw against torch.quasirandom.SobolEngine: integer work, so the bar is bit-exact -- including torch's FP32-rounded
    first point, continued draws, fast_forward and ragged run boundaries."""
    from hdbo_benchmark_b200.sampling import DeviceSobol

    ref = torch.quasirandom.SobolEngine(dim, scramble=True, seed=seed)
    dev = DeviceSobol(dim, scramble=True, seed=seed, device=cuda_device)
    for n in (1, 63, 64, 1000, 4097):
        got, want = dev.draw(n).cpu(), ref.draw(n, dtype=torch.float64)
        bad = (got != want).nonzero()
        assert bad.numel() == 0, (dim, n, bad[:4].tolist(), [float(got[tuple(b)]).hex() for b in bad[:4]],
                                  [float(want[tuple(b)]).hex() for b in bad[:4]])
    ref.fast_forward(12345); dev.fast_forward(12345)
    assert torch.equal(dev.draw(777).cpu(), ref.draw(777, dtype=torch.float64))
    # empty draw, unscrambled engine
    assert dev.draw(0).shape == (0, dim)
    ref_u = torch.quasirandom.SobolEngine(dim, scramble=False)
    dev_u = DeviceSobol(dim, scramble=False, device=cuda_device)
    assert torch.equal(dev_u.draw(130).cpu(), ref_u.draw(130, dtype=torch.float64))


def test_draw_sobol_samples_device_matches_host(cuda_device):
    from hdbo_benchmark_b200.sampling import draw_sobol_samples

    bounds = torch.stack([torch.linspace(-5.0, -1.0, 6, dtype=torch.float64), torch.linspace(0.5, 7.0, 6, dtype=torch.float64)])
    for q in (1, 3):
        host = draw_sobol_samples(bounds, n=500, q=q, seed=17)
        devs = draw_sobol_samples(bounds, n=500, q=q, seed=17, device=cuda_device)
        assert devs.is_cuda and torch.equal(devs.cpu(), host)
        assert torch.equal(draw_sobol_samples(bounds, n=500, q=q, seed=17, device=cuda_device).cpu(), host)   # seeded: repeatable


def test_mvn_sample_matches_torch_cholesky(cuda_device):
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for q in (1, 2, 5, 8):
        R, S = 37, 19
        A_ = torch.randn(R, q, q + 2, generator=g, dtype=torch.float64)
        Sig = A_ @ A_.transpose(-1, -2) + 0.1 * torch.eye(q, dtype=torch.float64)
        mean = torch.randn(R, q, generator=g, dtype=torch.float64)
        L = torch.linalg.cholesky(Sig + 1e-8 * torch.eye(q, dtype=torch.float64))
        for per_r in (0, 1):
            z = torch.randn(S, R, q, generator=g, dtype=torch.float64) if per_r else torch.randn(S, q, generator=g, dtype=torch.float64)
            want = mean + torch.einsum("rjk,srk->srj", L, z if per_r else z[:, None, :].expand(S, R, q))
            out = torch.empty(S, R, q, dtype=torch.float64, device=cuda_device)
            info = torch.zeros(R, dtype=torch.int32, device=cuda_device)
            md, sd, zd = mean.to(cuda_device), Sig.to(cuda_device), z.to(cuda_device)
            _lib.check(lib.hdbo_mvn_sample(md.data_ptr(), sd.data_ptr(), zd.data_ptr(), R, q, S, per_r, 1e-8, out.data_ptr(),
                                           info.data_ptr(), torch.cuda.current_stream().cuda_stream), "hdbo_mvn_sample")
            assert float((out.cpu() - want).abs().max()) < 1e-12 * max(1.0, float(want.abs().max()))
            assert int(info.abs().max()) == 0
    bad = -torch.eye(3, dtype=torch.float64, device=cuda_device).reshape(1, 3, 3)
    info = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    out = torch.empty(1, 1, 3, dtype=torch.float64, device=cuda_device)
    z = torch.zeros(1, 3, dtype=torch.float64, device=cuda_device)
    mean = torch.zeros(1, 3, dtype=torch.float64, device=cuda_device)
    _lib.check(lib.hdbo_mvn_sample(mean.data_ptr(), bad.data_ptr(), z.data_ptr(), 1, 3, 1, 0, 0.0, out.data_ptr(), info.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "hdbo_mvn_sample")
    assert int(info[0]) == 1                                       # non-positive first pivot reported, no NaN
    assert lib.hdbo_mvn_sample(mean.data_ptr(), bad.data_ptr(), z.data_ptr(), 1, 9, 1, 0, 0.0, out.data_ptr(), 0, None) < 0


def test_posterior_rsample_uses_device_factor_and_matches_oracle_joint(cuda_device):
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP

    n, d, b, q = 90, 6, 5, 3
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    Zq = O.sobol_candidates(b * q, d, seed=2).reshape(b, q, d)
    post = model.posterior(Zq)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(11, b, q, 1, generator=g, dtype=torch.float64)
    got = post.rsample_from_base_samples(torch.Size([11]), z)
    mean_o, Sig_o = O.posterior_joint(st, Zq)
    L = torch.linalg.cholesky(Sig_o + 1e-8 * torch.eye(q, dtype=torch.float64))
    want = mean_o.unsqueeze(0) + (L.unsqueeze(0) @ z).squeeze(-1)
    assert got.shape == (11, b, q, 1)
    assert float((got.squeeze(-1) - want).abs().max()) < 1e-7
    shared = torch.randn(4, 1, q, 1, generator=g, dtype=torch.float64)      # base samples shared by the t-batches
    got2 = post.rsample_from_base_samples(torch.Size([4]), shared)
    want2 = mean_o.unsqueeze(0) + (L.unsqueeze(0) @ shared).squeeze(-1)
    assert float((got2.squeeze(-1) - want2).abs().max()) < 1e-7
    # likelihood(model(x)): predictive distribution = latent variance + noise (GPyTorch idiom)
    pred = model.likelihood(post)
    assert float((pred.variance - post.variance - nz).abs().max()) < 1e-9


def test_condition_on_observations_equals_fresh_model_with_same_hyperparameters(cuda_device):
    from hdbo_benchmark_b200.models import Normalize, SingleTaskGP

    g = torch.Generator().manual_seed(0)
    n, d = 40, 4
    X = torch.rand(n, d, generator=g, dtype=torch.float64) * 3.0 - 1.0
    y = (torch.sin(X.sum(-1)) * 5.0 + 20.0).unsqueeze(-1)
    Xn = torch.rand(6, d, generator=g, dtype=torch.float64) * 3.0 - 1.0
    yn = (torch.sin(Xn.sum(-1)) * 5.0 + 20.0).unsqueeze(-1)
    model = SingleTaskGP(X, y, input_transform=Normalize(d))
    model.covar_module.lengthscale = torch.full((d,), 0.7, dtype=torch.float64)
    model.likelihood.noise = 2e-3
    model.eval()
    offset_before = model.input_transform.offset.clone()
    cond = model.condition_on_observations(Xn, yn)
    assert torch.equal(model.input_transform.offset, offset_before)                      # the original is not mutated
    assert cond.input_transform is not model.input_transform
    assert cond.train_targets.shape[0] == n + 6 and model.train_targets.shape[0] == n
    # reference: exact GP on the same standardised / normalised data with the same hyper-parameters
    ot, it = model.outcome_transform, model.input_transform
    Xa = it(torch.cat([X, Xn]).to(it.offset.device)).cpu()
    ya = ((torch.cat([y, yn]) - ot.means.cpu().reshape(())) / ot.stdvs.cpu().reshape(())).squeeze(-1)
    st = O.fit_state(Xa, ya, torch.full((d,), 0.7, dtype=torch.float64), 1.0, 2e-3, float(model.mean_module.constant), O.RBF)
    Zraw = torch.rand(50, d, generator=g, dtype=torch.float64) * 3.0 - 1.0
    m_o, v_o = O.posterior(st, it(Zraw.to(it.offset.device)).cpu())
    s, sh = float(ot.stdvs.reshape(())), float(ot.means.reshape(()))
    post = cond.posterior(Zraw.unsqueeze(1))
    assert float((post.mean.reshape(-1) - (m_o * s + sh)).abs().max()) < REL * float((m_o * s + sh).abs().max())
    assert float(((post.variance.reshape(-1) - v_o * s * s).abs() / (v_o * s * s)).max()) < REL
    # the new observations are interpolated (up to the noise) in ORIGINAL units: wrong un-standardisation would miss by ~20
    at_new = cond.posterior(Xn.unsqueeze(1)).mean.reshape(-1)
    assert float((at_new - yn.reshape(-1)).abs().max()) < 1.0


def test_acqf_host_pool_path_equals_device_path(cuda_device):
    """LogExpectedImprovement(model)(X) for a CPU tensor X streams the pool through pinned slabs; same values as the device path."""
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.models import Normalize, SingleTaskGP

    n, d, m = 300, 20, (1 << 15) + 77
    X, y, *_ = O.synthetic_problem(n, d)
    model = SingleTaskGP(X * 4.0 - 2.0, y.unsqueeze(-1) * 3.0 + 1.0, input_transform=Normalize(d))
    model.eval()
    acqf = A.LogExpectedImprovement(model, best_f=float(y.max()) * 3.0 + 1.0)
    Z = O.sobol_candidates(m, d, seed=4) * 4.0 - 2.0
    with torch.no_grad():
        a_host = acqf(Z.unsqueeze(1))
        a_dev = acqf(Z.to(cuda_device).unsqueeze(1))
    assert not a_host.is_cuda and a_dev.is_cuda and a_host.shape == (m,)
    assert torch.equal(a_host, a_dev.cpu())
    assert model._gp.var_mode == "i8" and model._gp._i8_state is not None      # the default surface takes the tcgen05 path


def test_caller_owned_timeline_records_phases(cuda_device):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import mll_gradient

    X, y, ls, os_, nz, mu = O.synthetic_problem(700, 40)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, keep_r2=True)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.attach_timeline(256)
    gp.fit(need_r2=True)
    mll_gradient(gp)
    Z = O.sobol_candidates(5000, 40).to(cuda_device)
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
    tl = gp.collect_timeline()
    assert set(tl) >= {"fit_build", "fit_factor", "fit_solve", "mll_grad", "prep", "cross", "var", "finalize"}
    assert all(ms > 0.0 and cnt >= 1 for ms, cnt in tl.values())
    gp.detach_timeline()
    gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
    assert gp.collect_timeline() == {}
