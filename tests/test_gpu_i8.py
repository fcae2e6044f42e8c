"""GPU parity of the exact int8 (tcgen05 digit-sliced) variance mode: hdbo_gp_slice_linv + hdbo_posterior_acq_i8.

The mode must be indistinguishable from the FP64 DMMA mode at the parity bar of the path (rel 1e-6 on mean / variance, same
arg-max): compared here with the CPU oracle, with the FP64 mode (absolute agreement, in units of the prior variance), on ragged
shapes, multi-chunk workspaces, candidates that coincide with training points (k = 1 exactly: top digit 64) and through the
host-buffer entry.  The pipeline-timeout flag of the kernel must stay 0."""
import ctypes as C

import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-6           # parity bar of the path (BASELINE.json north_star)
ABS_VS_F64 = 5e-12   # |var_i8 - var_f64| / outputscale: digit quantisation 2^-49 x conditioning, measured <= 5e-13


def _gp(X, y, kind, hyper, dev, mode, force_i8=True):
    from hdbo_benchmark_b200.engine import DeviceGP

    ls, os_, nz, mu = hyper
    gp = DeviceGP(X.to(dev), y.to(dev), kind, var_mode=mode)
    if force_i8:
        gp.i8_min_noise_ratio = 0.0      # these tests exercise the int8 kernels themselves, also below the engine's noise-ratio guard
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()
    return gp


@pytest.mark.parametrize("n,d,m,kind,ls_mode,noise", [
    (300, 20, 1000, O.MATERN52, "prior", 1e-3),
    (300, 20, 1000, O.RBF, "prior", 1e-3),
    (1024, 128, 4096, O.MATERN52, "long", 1e-4),
    (700, 33, 515, O.RBF, "short", 1e-3),
    (129, 15, 127, O.MATERN52, "prior", 1e-2),
    (2, 3, 5, O.MATERN52, "prior", 1e-2),
    (64, 1280, 33, O.RBF, "prior", 1e-2),
])
def test_i8_mode_matches_oracle_and_f64_mode(cuda_device, n, d, m, kind, ls_mode, noise):
    from hdbo_benchmark_b200 import _lib

    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, ls_mode, noise)
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    Z = O.sobol_candidates(m, d)
    k = min(m, n) // 2
    Z[:k] = X[:k]                                    # candidates ON training points: variance ~ noise, worst cancellation
    m_o, v_o = O.posterior(st, Z)
    best = float(y.max())
    a_o = O.acquisition(m_o, v_o, _lib.ACQ_LOGEI, best)
    out = {}
    for mode in ("f64", "i8"):
        gp = _gp(X, y, kind, (ls, os_, nz, mu), cuda_device, mode)
        mean, var, acq = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, best)
        out[mode] = (mean[0].cpu(), var[0].cpu(), acq[0].cpu())
        assert gp.i8_flag() == 0
    mean, var, acq = out["i8"]
    assert not torch.isnan(var).any()
    assert float((mean - m_o).abs().max()) < REL * float(m_o.abs().max())
    assert float(((var - v_o).abs() / v_o.abs()).max()) < REL
    assert float((var - out["f64"][1]).abs().max()) < ABS_VS_F64 * float(os_)
    assert float((mean - out["f64"][0]).abs().max()) < 1e-9 * float(m_o.abs().max())
    assert int(acq.argmax()) == int(a_o.argmax())
    assert float((acq - a_o).abs().max()) < REL * max(1.0, float(a_o.abs().max()))


def test_i8_single_training_point(cuda_device):
    from hdbo_benchmark_b200 import _lib

    X = torch.tensor([[0.2, 0.7, 0.4]], dtype=torch.float64)
    y = torch.tensor([0.3], dtype=torch.float64)
    ls = torch.tensor([0.5, 0.8, 1.1], dtype=torch.float64)
    st = O.fit_state(X, y, ls, 0.8, 1e-2, -0.3, O.MATERN52)
    Z = torch.cat([X, torch.rand(4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))])
    m_o, v_o = O.posterior(st, Z)
    gp = _gp(X, y, O.MATERN52, (ls, 0.8, 1e-2, -0.3), cuda_device, "i8")
    mean, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE)
    assert float((mean[0].cpu() - m_o).abs().max()) < REL * float(m_o.abs().max())
    assert float(((var[0].cpu() - v_o).abs() / v_o.abs()).max()) < REL
    assert gp.i8_flag() == 0


def test_i8_small_noise_on_training_points(cuda_device):
    """noise 1e-6: the posterior variance at training points is ~1e-6 of the prior -- the relative bar still holds."""
    from hdbo_benchmark_b200 import _lib

    n, d = 1000, 64
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-6)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = torch.cat([X[:500], O.sobol_candidates(500, d)])
    _, v_o = O.posterior(st, Z)
    gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, "i8")
    _, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE)
    assert float(((var[0].cpu() - v_o).abs() / v_o.abs()).max()) < REL
    assert gp.i8_flag() == 0


def test_i8_multi_chunk_workspace_and_refit(cuda_device):
    """Small chunk_rows forces several chunks (one ragged); a refit with new hyper-parameters re-slices L^-1."""
    from hdbo_benchmark_b200 import _lib

    n, d, m = 520, 24, 1500
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    Z = O.sobol_candidates(m, d, seed=5)
    gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, "i8")
    for scale in (1.0, 0.6):
        gp.set_hyper((ls * scale).to(cuda_device), os_ * scale, nz, mu)
        gp.fit()
        st = O.fit_state(X, y, ls * scale, os_ * scale, nz, mu, O.MATERN52)
        m_o, v_o = O.posterior(st, Z)
        for chunk in (128, 384, 148 * 128):
            mean, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE, chunk_rows=chunk)
            assert float((mean[0].cpu() - m_o).abs().max()) < REL * float(m_o.abs().max())
            assert float(((var[0].cpu() - v_o).abs() / v_o.abs()).max()) < REL
    assert gp.i8_flag() == 0


def test_i8_host_buffer_entry(cuda_device):
    from hdbo_benchmark_b200 import _lib

    n, d, m = 384, 32, 5000
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(m, d, seed=9)
    a_o = O.acquisition(*O.posterior(st, Z), _lib.ACQ_EI, float(y.max()))
    gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, "i8")
    Zh = Z.pin_memory()
    acq, val, idx = gp.posterior_acq_host(Zh, _lib.ACQ_EI, float(y.max()), copy_rows=2048)
    torch.cuda.synchronize()
    assert int(idx.item()) == int(a_o.argmax())
    assert float((acq.cpu() - a_o).abs().max()) < REL * float(a_o.abs().max())


def test_i8_c_abi_argument_checks(cuda_device):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    lib = _lib.load()
    X, y, ls, os_, nz, mu = O.synthetic_problem(100, 5)
    gpb = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, batch=2, var_mode="i8")
    gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, "i8")
    need = lib.hdbo_i8_state_bytes(C.byref(gp.desc))
    assert need == 6 * 128 * 128 + 8 * 128 + 6 * 128 * 32 + 8 * 128 + 16   # L^-1 images + scales, X images + scales, flag
    assert lib.hdbo_i8_state_bytes(C.byref(gpb.desc)) == 2 * (need - 16) + 16   # batched (SAAS) GPs: per-GP blocks
    state = torch.empty(need + 256, dtype=torch.uint8, device=cuda_device)
    assert lib.hdbo_gp_slice_linv(C.byref(gp.desc), state.data_ptr() + 8, 0) == -2      # misaligned state
    assert lib.hdbo_gp_slice_linv(C.byref(gp.desc), 0, 0) == -2
    Z = torch.rand(10, 5, dtype=torch.float64, device=cuda_device)
    out = torch.empty(10, dtype=torch.float64, device=cuda_device)
    ws = gp.workspace(128)
    args = (Z.data_ptr(), 5, 10, 0, _lib.ACQ_NONE, 0.0, out.data_ptr(), 0, 0, 10)
    assert lib.hdbo_posterior_acq_i8(C.byref(gp.desc), 0, *args, ws.data_ptr(), ws.numel(), 0) == -2
    assert lib.hdbo_posterior_acq_i8(C.byref(gp.desc), state.data_ptr(), *args, 0, 0, 0) == -12
    assert lib.hdbo_posterior_acq_i8(C.byref(gp.desc), state.data_ptr(), *args, ws.data_ptr() + 8, ws.numel() - 8, 0) == -13
    torch.cuda.synchronize()


def test_i8_full_size_against_f64_mode(cuda_device):
    """BASELINE.json headline shape (n = 4096, d = 256): the oracle is too slow for a big pool, so the int8 mode is checked
    against the FP64 DMMA mode (itself oracle-checked at this size in test_gpu_fullsize.py) on a 2^15 pool."""
    from hdbo_benchmark_b200 import _lib

    n, d, m = 4096, 256, 1 << 15
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    Z = torch.rand(m, d, dtype=torch.float64, generator=torch.Generator().manual_seed(11)).to(cuda_device)
    Z[:100] = X[:100].to(cuda_device)
    res = {}
    for mode in ("f64", "i8"):
        gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, mode)
        mean, var, acq = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
        res[mode] = (mean[0].clone(), var[0].clone(), acq[0].clone())
        assert gp.i8_flag() == 0
        del gp
    assert float((res["i8"][0] - res["f64"][0]).abs().max()) < 1e-9 * float(res["f64"][0].abs().max())
    assert float((res["i8"][1] - res["f64"][1]).abs().max()) < ABS_VS_F64 * float(os_)
    assert float(((res["i8"][1] - res["f64"][1]).abs() / res["f64"][1].abs()).max()) < REL
    assert int(res["i8"][2].argmax()) == int(res["f64"][2].argmax())
    assert (res["i8"][1] > 0).all()


def test_i8_mode_through_the_botorch_surface(cuda_device, monkeypatch):
    """HDBO_VAR_MODE=i8: SingleTaskGP + LogExpectedImprovement under no_grad (the raw-sample screening of optimize_acqf) runs on the
    int8 path, the autograd path stays FP64; both agree with the oracle, and a hyper-parameter change re-slices lazily."""
    monkeypatch.setenv("HDBO_VAR_MODE", "i8")
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import SingleTaskGP

    n, d = 260, 12
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-3, seed=2)
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    Z = O.sobol_candidates(700, d, seed=4)
    best = float(y.max()) - 0.2
    for scale in (1.0, 0.7):
        model.covar_module.base_kernel.lengthscale = ls * scale
        model.covar_module.outputscale = float(os_) * scale
        model.eval()
        st = O.fit_state(X, y, ls * scale, float(os_) * scale, nz, mu, O.MATERN52)
        a_o, g_o = O.acquisition_value_and_grad_autograd(st, Z[:30], O.ACQ_LOGEI, best)
        a_all = O.acquisition(*O.posterior(st, Z), O.ACQ_LOGEI, best)
        f = A.LogExpectedImprovement(model, best_f=best)
        with torch.no_grad():
            a = f(Z.unsqueeze(1))
        assert model._gp.var_mode == "i8" and model._gp._i8_state is not None and model._gp.i8_flag() == 0
        assert float((a.cpu() - a_all).abs().max()) < REL * max(1.0, float(a_all.abs().max()))
        Zg = Z[:30].clone().unsqueeze(1).requires_grad_(True)
        (g,) = torch.autograd.grad(f(Zg).sum(), Zg)
        assert float((g.squeeze(1).cpu() - g_o).abs().max()) < REL * float(g_o.abs().max())


def test_i8_saas_batch_matches_per_sample_oracle(cuda_device):
    """A10 on the int8 path: S GPs with their own lengthscales / outputscales / noises, one batched launch sequence."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import sample_saas_prior

    n, d, S, m = 150, 70, 5, 333
    X, y, *_ = O.synthetic_problem(n, d)
    smp = sample_saas_prior(S, d, seed=4)
    states = O.saas_fit_states(X, y, smp["lengthscale"], smp["outputscale"], smp["noise"], smp["mean"])
    Z = O.sobol_candidates(m, d, seed=9)
    Z[:50] = X[:50]
    mu_o, var_o, _, _ = O.saas_posterior(states, Z)
    best = float(y.max())
    out = {}
    for mode in ("f64", "i8"):
        gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, batch=S, var_mode=mode)
        gp.set_hyper(smp["lengthscale"].to(cuda_device), smp["outputscale"], smp["noise"], smp["mean"])
        gp.fit()
        for chunk in (128, 148 * 128):
            mean, var, acq = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, best, chunk_rows=chunk)
            assert mean.shape == (S, m)
            assert float((mean.cpu() - mu_o).abs().max()) < REL * float(mu_o.abs().max())
            assert float(((var.cpu() - var_o).abs() / var_o.abs()).max()) < REL
        out[mode] = acq.cpu()
        assert gp.i8_flag() == 0
    a_o = O.log_expected_improvement(mu_o, var_o, best)
    assert float((out["i8"] - a_o).abs().max()) < REL * max(1.0, float(a_o.abs().max()))
    assert torch.equal(out["i8"].argmax(dim=1), out["f64"].argmax(dim=1))


def test_i8_largest_supported_size_and_fallback_beyond_it(cuda_device):
    """n_pad = 8192 is the largest size whose int32 digit-group sums are guaranteed exact (csrc/i8.cuh); beyond it the engine
    silently keeps the FP64 DMMA path."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    d, m = 16, 300
    g = torch.Generator().manual_seed(0)
    for n, expect_i8 in ((8192, True), (8200, False)):
        X = torch.rand(n, d, generator=g, dtype=torch.float64)
        y = torch.sin(3.0 * X[:, :3].sum(-1)) + 0.05 * torch.randn(n, generator=g, dtype=torch.float64)
        ls = torch.full((d,), 0.9, dtype=torch.float64)
        Z = torch.rand(m, d, generator=g, dtype=torch.float64).to(cuda_device)
        res = {}
        for mode in ("f64", "i8"):
            gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, var_mode=mode)
            gp.set_hyper(ls.to(cuda_device), 1.0, 1e-2, 0.0)
            gp.fit()
            mean, var, _ = gp.posterior_acq(Z, _lib.ACQ_NONE)
            res[mode] = (mean[0].clone(), var[0].clone())
            if mode == "i8":
                assert (gp._i8_state is not None) == expect_i8 and gp.i8_flag() == 0
            del gp
        assert float((res["i8"][0] - res["f64"][0]).abs().max()) < 1e-9 * float(res["f64"][0].abs().max())
        assert float(((res["i8"][1] - res["f64"][1]).abs() / res["f64"][1].abs()).max()) < REL
        torch.cuda.empty_cache()


def test_i8_results_are_bitwise_deterministic_and_chunk_invariant(cuda_device):
    """Integer accumulation + a fixed recombination order: the int8 path returns identical bits whatever the chunking of the pool,
    and on repeated calls (size-independent property, also checked at the headline shape)."""
    from hdbo_benchmark_b200 import _lib

    for n, d, m, chunks in ((600, 40, 3000, (128, 512, 1024, 148 * 128)), (4096, 256, 40000, (4096, 148 * 128, 2 * 148 * 128))):
        X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
        gp = _gp(X, y, O.MATERN52, (ls, os_, nz, mu), cuda_device, "i8")
        Z = torch.rand(m, d, dtype=torch.float64, generator=torch.Generator().manual_seed(n)).to(cuda_device)
        ref = None
        for chunk in chunks + (chunks[0],):
            mean, var, acq = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()), chunk_rows=chunk)
            cur = (mean.clone(), var.clone(), acq.clone())
            if ref is None:
                ref = cur
            else:
                assert all(torch.equal(a, b) for a, b in zip(cur, ref)), chunk
        assert gp.i8_flag() == 0
        del gp
        torch.cuda.empty_cache()


def test_i8_random_hyperparameter_sweep_against_f64_mode(cuda_device):
    """24 random (n, d, lengthscale scale, outputscale, noise, kernel) draws, including very short / very long lengthscales and
    large outputscales, noise down to 1e-5 of the prior variance.  At noise / outputscale >= 1e-4 (the engine's guard; BoTorch's own
    noise floor) the int8 path tracks the FP64 path to 1e-9 of the mean scale and, in the variance, to the relative parity bar (1e-6,
    also where the variance collapses to the noise level on training points) and 5e-11 of the prior variance; below the guard the
    engine silently keeps the FP64 path (the error grows like 2^-48 ||L^-1 row|| ~ (outputscale / noise)^1/2)."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    g = torch.Generator().manual_seed(123)
    for trial in range(24):
        n = int(torch.randint(3, 700, (1,), generator=g))
        d = int(torch.randint(1, 90, (1,), generator=g))
        m = int(torch.randint(1, 600, (1,), generator=g))
        kind = O.MATERN52 if trial % 2 == 0 else O.RBF
        ls = (10.0 ** (torch.rand(d, generator=g, dtype=torch.float64) * 3.0 - 1.5)) * max(1.0, d ** 0.5 / 3)   # 0.03 .. 30 x sqrt(d)/3
        os_ = float(10.0 ** (torch.rand(1, generator=g, dtype=torch.float64) * 4.0 - 2.0))                    # 0.01 .. 100
        nz = float(10.0 ** (torch.rand(1, generator=g, dtype=torch.float64) * 4.0 - 5.0)) * os_               # 1e-5 .. 1e-1 of the prior
        guarded = nz < 1e-4 * os_                     # below the engine's noise-ratio guard the int8 request falls back to FP64
        X = torch.rand(n, d, generator=g, dtype=torch.float64)
        y = torch.randn(n, generator=g, dtype=torch.float64) * os_ ** 0.5
        Z = torch.rand(m, d, generator=g, dtype=torch.float64)
        Z[: min(m, n) // 3] = X[: min(m, n) // 3]
        res = {}
        for mode in ("f64", "i8"):
            gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), kind, var_mode=mode)
            gp.set_hyper(ls.to(cuda_device), os_, nz, 0.1)
            gp.fit()
            mean, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE)
            res[mode] = (mean[0].clone(), var[0].clone())
            assert gp.i8_flag() == 0
            if mode == "i8":
                assert (gp._i8_state is None) == guarded, (trial, nz / os_)
        if guarded:
            assert torch.equal(res["i8"][1], res["f64"][1]) and torch.equal(res["i8"][0], res["f64"][0])
            continue
        scale_m = float(res["f64"][0].abs().max().clamp_min(os_ ** 0.5))
        assert float((res["i8"][0] - res["f64"][0]).abs().max()) < 1e-9 * scale_m, (trial, n, d)
        dv = (res["i8"][1] - res["f64"][1]).abs()
        assert float(dv.max()) < 5e-11 * os_, (trial, n, d, float(dv.max()) / os_, nz / os_)
        assert float((dv / res["f64"][1].abs()).max()) < REL, (trial, n, d, nz / os_)


def test_nan_candidate_gives_nan_mean_and_acquisition(cuda_device):
    """The integer clamps of the cross epilogue must not turn a NaN coordinate into a finite prediction."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(200, 12)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52, var_mode="i8")
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu); gp.fit()
    Z = torch.rand(300, 12, dtype=torch.float64, device=cuda_device)
    Z[17, 3] = float("nan")
    mean, var, acq = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
    assert bool(torch.isnan(mean[0, 17])) and bool(torch.isnan(acq[0, 17]))
    keep = torch.ones(300, dtype=torch.bool, device=cuda_device); keep[17] = False
    assert bool(torch.isfinite(mean[0, keep]).all()) and bool(torch.isfinite(acq[0, keep]).all()) and gp.i8_flag() == 0
