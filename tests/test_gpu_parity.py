"""GPU parity tests: the CUDA path (through the C ABI, libhdbo_b200.so) against the CPU oracle on the same seeded
inputs.  Tolerances are the ones BASELINE.json's north_star states: relative 1e-6 on posterior mean / variance and
the MLL (we assert tighter where the arithmetic allows), same arg-max candidate for the acquisition."""
import ctypes as C
import math

import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-6


def rel_max(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def rel_elem(a, b, floor=1e-300):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def make_gp(dev, n, d, kind, ls_mode="prior", noise=1e-3, keep_r2=False):
    from hdbo_benchmark_b200.engine import DeviceGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, ls_mode, noise)
    os_, mu = 1.3, 0.2
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    gp = DeviceGP(X.to(dev), y.to(dev), kind, keep_r2=keep_r2)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()
    return gp, st, (X, y, ls, os_, nz, mu)


# ------------------------------------------------------------------------------------------------------------
def test_dgemm_mainloop_matches_torch(cuda_device):
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(384, 256, generator=g, dtype=torch.float64)
    B = torch.randn(256, 256, generator=g, dtype=torch.float64)
    Ad, Bd = A.to(cuda_device), B.to(cuda_device)
    st = torch.cuda.current_stream().cuda_stream
    Cd = torch.full((384, 256), 7.0, dtype=torch.float64, device=cuda_device)
    _lib.check(lib.hdbo_dgemm_nt(Ad.data_ptr(), 256, Bd.data_ptr(), 256, Cd.data_ptr(), 256, 384, 256, 256, 2.0, 0.5,
                                 _lib.KR_FULL, 0, st), "dgemm")
    want = 2.0 * A @ B.T + 3.5
    assert rel_max(Cd, want) < 1e-13
    # triangular k-range: B lower triangular (k-tile <= n-tile) gives the same result as the full product
    Bl = torch.tril(B[:256, :256])
    Cd.zero_()
    _lib.check(lib.hdbo_dgemm_nt(Ad.data_ptr(), 256, Bl.to(cuda_device).data_ptr(), 256, Cd.data_ptr(), 256, 384, 256, 256,
                                 1.0, 0.0, _lib.KR_LE_N, 0, st), "dgemm")
    assert rel_max(Cd, A @ Bl.T) < 1e-13
    # invalid arguments are reported, not executed
    assert lib.hdbo_dgemm_nt(Ad.data_ptr(), 256, Bd.data_ptr(), 256, Cd.data_ptr(), 256, 100, 256, 256, 1.0, 0.0, 0, 0, st) < 0


@pytest.mark.parametrize("n,d,kind,ls_mode,noise", [
    (300, 20, O.MATERN52, "prior", 1e-3),      # ragged n and d (padding paths)
    (129, 7, O.RBF, "prior", 1e-2),
    (1024, 128, O.MATERN52, "prior", 1e-3),    # BASELINE config #2 shape
    (1024, 128, O.MATERN52, "long", 1e-4),     # stress: long lengthscales, |alpha| large
    (640, 33, O.RBF, "short", 1e-3),
])
def test_fit_state_matches_oracle(cuda_device, n, d, kind, ls_mode, noise):
    gp, st, _ = make_gp(cuda_device, n, d, kind, ls_mode, noise)
    assert int(gp.info.item()) == 0
    assert rel_max(gp.L[0, :n, :n].tril(), st.L) < 1e-10
    assert rel_max(gp.Linv[0, :n, :n], st.Linv) < 1e-9
    assert rel_max(gp.Linvt[0, :n, :n], st.Linv.T) < 1e-9
    assert rel_max(gp.alpha[0, :n], st.alpha) < 1e-8
    assert abs(gp.logdet.item() - st.logdet) < 1e-10 * abs(st.logdet)
    # padding of the inverse is the identity / zero, so padded candidates contribute nothing
    if gp.n_pad > n:
        pad = gp.Linv[0, n:, n:]
        assert torch.equal(pad, torch.eye(gp.n_pad - n, dtype=torch.float64, device=cuda_device))
        assert float(gp.Linv[0, n:, :n].abs().max()) == 0.0
        assert float(gp.alpha[0, n:].abs().max()) == 0.0


@pytest.mark.parametrize("n,d,m,kind,ls_mode,noise", [
    (300, 20, 1000, O.MATERN52, "prior", 1e-3),
    (300, 20, 257, O.RBF, "prior", 1e-3),
    (1024, 128, 4096, O.MATERN52, "prior", 1e-3),
    (1024, 128, 2048, O.MATERN52, "long", 1e-4),
    (1024, 128, 2048, O.RBF, "short", 1e-3),
])
def test_posterior_and_acquisition_match_oracle(cuda_device, n, d, m, kind, ls_mode, noise):
    from hdbo_benchmark_b200 import _lib

    gp, st, (X, y, *_rest) = make_gp(cuda_device, n, d, kind, ls_mode, noise)
    Z = O.sobol_candidates(m, d)
    Z[0] = X[3] + 1e-4  # next to a training point: variance cancellation
    mean_o, var_o = O.posterior(st, Z)
    best_f = float(y.max())
    Zd = Z.to(cuda_device)
    for kind_id, param in ((_lib.ACQ_EI, best_f), (_lib.ACQ_LOGEI, best_f), (_lib.ACQ_UCB, 2.0)):
        mean, var, acq = gp.posterior_acq(Zd, kind_id, param)
        assert rel_max(mean[0], mean_o) < REL
        big = var_o > 1e-8  # relative variance parity only where var >> eps * cond
        assert rel_elem(var[0].cpu()[big], var_o[big]) < REL
        assert float((var[0].cpu() - var_o).abs().max()) < 1e-9
        a_o = O.acquisition(mean_o, var_o, kind_id, param)
        if kind_id == _lib.ACQ_EI:
            assert rel_max(acq[0], a_o) < REL
        else:
            assert float((acq[0].cpu() - a_o).abs().max()) < 1e-6 * max(1.0, float(a_o.abs().max()))
        assert int(acq[0].argmax()) == int(a_o.argmax())
        val, idx = gp.argmax(acq[0])
        assert int(idx) == int(a_o.argmax())
    # observation noise adds sigma^2
    _, var_n, _ = gp.posterior_acq(Zd, _lib.ACQ_NONE, 0.0, observation_noise=True)
    assert rel_max(var_n[0], var_o + st.noise) < REL


def test_posterior_chunking_is_invisible(cuda_device):
    from hdbo_benchmark_b200 import _lib

    gp, st, (X, y, *_r) = make_gp(cuda_device, 300, 20, O.MATERN52)
    Z = O.sobol_candidates(1000, 20).to(cuda_device)
    m1, v1, a1 = gp.posterior_acq(Z, _lib.ACQ_LOGEI, 0.5, chunk_rows=128)
    gp._ws = None
    m2, v2, a2 = gp.posterior_acq(Z, _lib.ACQ_LOGEI, 0.5, chunk_rows=4096)
    assert torch.equal(m1, m2) and torch.equal(v1, v2) and torch.equal(a1, a2)
    # empty pool
    m0, v0, a0 = gp.posterior_acq(Z[:0], _lib.ACQ_EI, 0.5)
    assert m0.shape == (1, 0)


@pytest.mark.parametrize("n,d,kind,ls_mode", [
    (200, 12, O.MATERN52, "prior"), (200, 12, O.RBF, "prior"), (700, 40, O.MATERN52, "long"), (1024, 128, O.RBF, "prior"),
])
def test_mll_value_and_gradient_match_oracle(cuda_device, n, d, kind, ls_mode):
    from hdbo_benchmark_b200.models import mll_gradient

    gp, st, (X, y, ls, os_, nz, mu) = make_gp(cuda_device, n, d, kind, ls_mode, keep_r2=True)
    v, g_ls, g_os, g_nz, g_mu = O.exact_mll_and_grad(X, y, ls, os_, nz, mu, kind)
    assert abs(gp.mll_data_term.item() / n - v.item()) < REL * abs(v.item())
    grad = mll_gradient(gp)[0].cpu() / n
    scale = float(g_ls.abs().max())
    assert float((grad[:d] - g_ls).abs().max()) < REL * scale
    assert abs(grad[d] - g_os) < REL * max(abs(float(g_os)), 1e-3)
    assert abs(grad[d + 1] - g_nz) < REL * max(abs(float(g_nz)), 1e-3)
    assert abs(grad[d + 2] - g_mu) < REL * max(abs(float(g_mu)), 1e-3)


def test_branch_free_sqrt_exp_match_the_library(cuda_device):
    """csrc/fast_math.cuh (used by every distance -> kernel epilogue) against torch's float64 sqrt / exp."""
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load_dev()   # test hooks live in libhdbo_b200_dev.so, not in the product library
    g = torch.Generator().manual_seed(0)
    x = torch.cat([
        torch.rand(100000, generator=g, dtype=torch.float64) * 50.0,
        10.0 ** (torch.rand(50000, generator=g, dtype=torch.float64) * 60.0 - 30.0),   # 1e-30 .. 1e30
        torch.tensor([0.0, 1e-30, 1.0, 2.0, 4.0, 708.0, 739.9, 745.0, 1e4, 1e30], dtype=torch.float64),
    ]).to(cuda_device)
    s = torch.empty_like(x)
    e = torch.empty_like(x)
    _lib.check(lib.hdbo_test_math(x.data_ptr(), x.numel(), s.data_ptr(), e.data_ptr(), torch.cuda.current_stream().cuda_stream), "math")
    want_s = x.clamp(1e-30, 1e30).sqrt()
    assert float(((s - want_s).abs() / want_s).max()) < 4e-16
    want_e = torch.exp(-x)
    big = want_e > 1e-300
    assert float(((e - want_e).abs() / want_e)[big].max()) < 1e-15
    assert float(e[~big].abs().max()) < 1e-299 and bool((e >= 0).all())


def test_fp64_lean_sqrt_exp_of_the_int8_cross_kernel(cuda_device):
    """sqrt_pos_cubic / exp_neg_tab_inrange (csrc/fast_math.cuh; 5 and 10 FP64 instructions) against torch's float64 sqrt / exp."""
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load_dev()
    g = torch.Generator().manual_seed(1)
    x = torch.cat([
        torch.rand(200000, generator=g, dtype=torch.float64) * 60.0,
        10.0 ** (torch.rand(100000, generator=g, dtype=torch.float64) * 330.0 - 30.0),   # 1e-30 .. 1e300
        torch.tensor([0.0, 1e-30, 2.13e-14, 1.0, 2.0, 4.0, 632.4, 699.9, 700.0, 4e5, 1e300], dtype=torch.float64),
    ]).to(cuda_device)
    s = torch.empty_like(x)
    e = torch.empty_like(x)
    _lib.check(lib.hdbo_test_math_lean(x.data_ptr(), x.numel(), s.data_ptr(), e.data_ptr(), torch.cuda.current_stream().cuda_stream), "math")
    want_s = x.clamp(1e-30, 1e300).sqrt()
    assert float(((s - want_s).abs() / want_s).max()) < 4e-16
    want_e = torch.exp(-x.clamp(max=700.0))
    assert float(((e - want_e).abs() / want_e).max()) < 1e-15
    assert bool((e > 0).all()) and bool((e <= 1).all())


def test_cholesky_failure_reports_info_and_jitter_recovers(cuda_device):
    from hdbo_benchmark_b200.engine import DeviceGP, NotPSDError

    g = torch.Generator().manual_seed(0)
    X = torch.rand(1, 3, generator=g, dtype=torch.float64).repeat(40, 1)  # identical rows: K = 1 1^T, exactly singular
    y = torch.randn(40, generator=g, dtype=torch.float64)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.RBF)
    gp.set_hyper(torch.full((3,), 0.7), 1.0, 0.0, 0.0)
    gp.fit_once()
    assert int(gp.info.item()) > 0  # LAPACK-style: 1-based index of the failing pivot
    jit = gp.fit()  # jitter schedule 1e-8 * 10**i
    assert jit in (1e-8, 1e-7, 1e-6) and int(gp.info.item()) == 0
    gp.set_hyper(torch.full((3,), 0.7), -1.0, 0.0, 0.0)  # negative definite: no jitter helps
    with pytest.raises(NotPSDError):
        gp.fit()
