"""The pipelined factorisation (csrc/kernels_chol_pipe.cu: chain / bulk / inverse streams, K^-1 accumulated beside the chain) against
the oracle and against the single-stream recursion, including ragged sizes, the jitter path and CUDA-graph capture."""
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-6


def _gp(dev, n, d, kind, pipeline, keep_r2=True, noise=1e-3, ls_mode="prior", monkeypatch=None):
    from hdbo_benchmark_b200.engine import DeviceGP

    if monkeypatch is not None:
        monkeypatch.setenv("HDBO_FIT_PIPELINE", "1" if pipeline else "0")
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, ls_mode, noise)
    gp = DeviceGP(X.to(dev), y.to(dev), kind, keep_r2=keep_r2, var_mode="f64")
    assert (gp._aux is not None) == bool(pipeline and gp.n_pad >= 1024)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    return gp, (X, y, ls, os_, nz, mu)


def _rel(a, b):
    return float((a.cpu() - b.cpu()).abs().max() / b.abs().max())


@pytest.mark.parametrize("n,d,kind", [(1024, 64, O.MATERN52), (1100, 33, O.RBF), (2000, 128, O.MATERN52)])
def test_pipelined_fit_and_gradient_match_oracle(cuda_device, monkeypatch, n, d, kind):
    from hdbo_benchmark_b200.models import mll_gradient

    gp, (X, y, ls, os_, nz, mu) = _gp(cuda_device, n, d, kind, True, monkeypatch=monkeypatch)
    gp.fit(need_r2=True)
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    assert int(gp.info.item()) == 0
    assert _rel(gp.L[0, :n, :n].tril(), st.L) < 1e-10
    assert _rel(gp.Linv[0, :n, :n], st.Linv) < 1e-9
    assert _rel(gp.Linvt[0, :n, :n], st.Linv.T) < 1e-9
    assert _rel(gp.alpha[0, :n], st.alpha) < 1e-8
    assert abs(gp.logdet.item() - st.logdet) < 1e-10 * abs(st.logdet)
    Kinv = st.Linv.T @ st.Linv
    assert gp._kinv_valid
    assert _rel(gp.Kinv[0, :n, :n].tril(), Kinv.tril()) < 1e-9          # accumulated beside the chain, lower tiles
    if gp.n_pad > n:                                                     # padding: identity block, zero coupling
        assert torch.equal(gp.Linv[0, n:, n:], torch.eye(gp.n_pad - n, dtype=torch.float64, device=cuda_device))
        assert float(gp.Linv[0, n:, :n].abs().max()) == 0.0
    v, g_ls, g_os, g_nz, g_mu = O.exact_mll_and_grad(X, y, ls, os_, nz, mu, kind)
    grad = mll_gradient(gp)[0].cpu() / n
    assert abs(gp.mll_data_term.item() / n - v.item()) < REL * abs(v.item())
    assert float((grad[:d] - g_ls).abs().max()) < REL * float(g_ls.abs().max())
    assert abs(grad[d] - g_os) < REL * max(abs(float(g_os)), 1e-3)
    assert abs(grad[d + 1] - g_nz) < REL * max(abs(float(g_nz)), 1e-3)
    assert abs(grad[d + 2] - g_mu) < REL * max(abs(float(g_mu)), 1e-3)


def test_pipelined_equals_recursive_and_is_deterministic(cuda_device, monkeypatch):
    from hdbo_benchmark_b200.models import mll_gradient

    n, d = 1500, 40
    res = {}
    for pipeline in (True, False):
        gp, _ = _gp(cuda_device, n, d, O.MATERN52, pipeline, monkeypatch=monkeypatch)
        gp.fit(need_r2=True)
        g1 = mll_gradient(gp).clone()
        Linv1 = gp.Linv.clone()
        gp.fit(need_r2=True)
        g2 = mll_gradient(gp)
        assert torch.equal(g1, g2) and torch.equal(Linv1, gp.Linv)          # run-to-run bitwise, both schedules
        res[pipeline] = (gp.Linv.clone(), gp.alpha.clone(), gp.scal.clone(), g1)
    for a, b in zip(res[True], res[False]):                                  # different summation orders: equal to rounding
        assert float((a - b).abs().max()) <= 1e-10 * float(b.abs().max())


def test_pipelined_fit_without_gradient_state_and_posterior(cuda_device, monkeypatch):
    from hdbo_benchmark_b200 import _lib

    n, d = 1300, 20
    gp, (X, y, ls, os_, nz, mu) = _gp(cuda_device, n, d, O.MATERN52, True, keep_r2=False, monkeypatch=monkeypatch)
    gp.fit()
    assert gp.Kinv is None
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(700, d)
    mean, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE, 0.0)
    m_o, v_o = O.posterior(st, Z)
    assert _rel(mean[0], m_o) < REL and float(((var[0].cpu() - v_o).abs() / v_o).max()) < REL


def test_pipelined_fit_reports_non_psd_and_recovers_with_jitter(cuda_device, monkeypatch):
    from hdbo_benchmark_b200.engine import DeviceGP, NotPSDError

    monkeypatch.setenv("HDBO_FIT_PIPELINE", "1")
    g = torch.Generator().manual_seed(0)
    n, d = 1100, 3
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    X[700] = X[10]                                   # duplicate points + (almost) no noise: singular to working precision
    X[900] = X[10]
    y = torch.randn(n, generator=g, dtype=torch.float64)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.RBF, var_mode="f64")
    assert gp._aux is not None
    gp.set_hyper(torch.full((d,), 5.0, dtype=torch.float64), 1.0, 1e-17, 0.0)
    try:
        jitter = gp.fit()
        assert jitter > 0.0 and int(gp.info.item()) == 0
    except NotPSDError:
        pass                                          # also acceptable: reported, not hung / NaN-silent
    gp.set_hyper(torch.full((d,), 0.5, dtype=torch.float64), 1.0, 1e-2, 0.0)
    assert gp.fit() == 0.0 and bool(torch.isfinite(gp.alpha).all())


def test_pipelined_fit_is_cuda_graph_capturable(cuda_device, monkeypatch):
    from hdbo_benchmark_b200.models import mll_gradient

    n, d = 1200, 16
    gp, _ = _gp(cuda_device, n, d, O.MATERN52, True, monkeypatch=monkeypatch)

    def evaluate():
        gp.fit_once(need_r2=True)
        return mll_gradient(gp)
    for _ in range(2):
        eager = evaluate().clone()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = evaluate()
    gp.alpha.zero_(); gp.Linv.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager) and float(gp.alpha.abs().max()) > 0.0
    # new hyper-parameters through the same graph
    gp.inv_ls.mul_(1.1)
    graph.replay()
    torch.cuda.synchronize()
    replayed = out.clone()
    fresh = evaluate()
    assert torch.equal(replayed, fresh)
