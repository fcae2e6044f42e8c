"""GPU edge cases: tiny / ragged / wide shapes, single candidates, float32 callers, duplicated training points, pool sizes
around the 128-row tile and 18 944-row chunk boundaries.  Every case is compared with the CPU oracle."""
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-6


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("n,d,m", [(1, 3, 5), (2, 1, 1), (5, 2, 130), (127, 16, 128), (128, 17, 129), (129, 15, 127),
                                   (64, 1280, 33)])   # d = 1280: one-hot Ehrlich width (64 x 20)
@pytest.mark.parametrize("kind", [O.MATERN52, O.RBF])
def test_small_ragged_and_wide_shapes(cuda_device, n, d, m, kind):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP
    from hdbo_benchmark_b200.models import mll_gradient

    g = torch.Generator().manual_seed(n * 1000 + d)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = torch.randn(n, generator=g, dtype=torch.float64)
    ls = (0.5 + torch.rand(d, generator=g, dtype=torch.float64)) * max(1.0, d ** 0.5 / 2)
    os_, nz, mu = 0.8, 1e-2, -0.3
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), kind, keep_r2=True)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.fit()
    Z = torch.rand(m, d, generator=g, dtype=torch.float64)
    mean, var, acq = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, float(y.max()))
    m_o, v_o = O.posterior(st, Z)
    assert _rel(mean[0], m_o) < REL and _rel(var[0], v_o) < REL
    a_o = O.log_expected_improvement(m_o, v_o, float(y.max()))
    assert float((acq[0].cpu() - a_o).abs().max()) < REL * max(1.0, float(a_o.abs().max()))
    v, g_ls, g_os, g_nz, g_mu = O.exact_mll_and_grad(X, y, ls, os_, nz, mu, kind)
    assert abs(gp.mll_data_term.item() / n - v.item()) < REL * max(abs(v.item()), 1e-3)
    grad = mll_gradient(gp)[0].cpu() / n
    assert float((grad[:d] - g_ls).abs().max()) < REL * max(float(g_ls.abs().max()), 1e-6)
    assert abs(grad[d + 1] - g_nz) < REL * max(abs(float(g_nz)), 1e-3)


@pytest.mark.parametrize("m", [18943, 18944, 18945, 2 * 18944 + 1])
def test_pool_sizes_around_chunk_boundaries(cuda_device, m):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(200, 10)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.fit()
    Z = O.sobol_candidates(m, 10, seed=1)
    mean, var, ei = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_EI, float(y.max()))
    assert mean.shape == (1, m)
    tail = slice(m - 300, m)   # the ragged end of the last chunk
    m_o, v_o = O.posterior(st, Z[tail])
    assert _rel(mean[0, tail], m_o) < REL and _rel(var[0, tail], v_o) < REL
    a_o = O.expected_improvement(*O.posterior(st, Z), float(y.max()))
    assert int(ei[0].argmax()) == int(a_o.argmax())


def test_float32_callers_and_host_tensors(cuda_device):
    """The reference's scripts run in float32 on the host (benchmark_on_pmo/run.py:36): inputs are accepted as they are,
    computed in FP64 on the device, results come back in the caller's dtype / device."""
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200.models import SingleTaskGP

    X, y, ls, *_ = O.synthetic_problem(60, 6)
    model = SingleTaskGP(X.float(), y.float().unsqueeze(-1))
    model.eval()
    Z = torch.rand(7, 1, 6, dtype=torch.float32)   # explicit: other test modules change the process-wide default dtype
    post = model.posterior(Z)
    assert post.mean.dtype == torch.float32 and post.mean.device.type == "cpu" and post.mean.shape == (7, 1, 1)
    acq = A.LogExpectedImprovement(model, best_f=float(y.max()))(Z)
    assert acq.dtype == torch.float32 and acq.shape == (7,)
    Zg = Z.clone().requires_grad_(True)
    A.ExpectedImprovement(model, best_f=float(y.max()))(Zg).sum().backward()
    assert Zg.grad is not None and Zg.grad.shape == Z.shape and torch.isfinite(Zg.grad).all()


def test_duplicate_training_points_and_candidates_at_training_points(cuda_device):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    X, y, ls, os_, nz, mu = O.synthetic_problem(90, 5)
    X[10] = X[3]
    X[50] = X[3]
    st = O.fit_state(X, y, ls, os_, 1e-3, mu, O.MATERN52)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52)
    gp.set_hyper(ls.to(cuda_device), os_, 1e-3, mu)
    assert gp.fit() == 0.0   # noise keeps the duplicated rows positive definite: no jitter needed
    Z = X[:20].clone()       # candidates ON training points: r2 = 0 exactly, variance cancellation
    mean, var, _ = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_NONE, 0.0)
    m_o, v_o = O.posterior(st, Z)
    assert _rel(mean[0], m_o) < REL
    assert float((var[0].cpu() - v_o).abs().max()) < 1e-9 and float(var[0].min()) > -1e-9
