"""Full-size parity (BASELINE.json configs[2]: n=4096, d=256) through size-independent properties plus an oracle slice:
interpolation identity mean(X_i) = y_i - noise * alpha_i, variance bounds, linearity of the posterior mean in y,
chunking invariance and run-to-run determinism (bitwise), arg-max agreement on a pool slice."""
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(cuda_device):
    from hdbo_benchmark_b200.engine import DeviceGP

    n, d = 4096, 256
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    gp = DeviceGP(X.to(cuda_device), y.to(cuda_device), O.MATERN52)
    gp.set_hyper(ls.to(cuda_device), os_, nz, mu)
    gp.fit()
    return gp, (X, y, ls, os_, nz, mu)


def test_fullsize_slice_matches_oracle_and_same_argmax(cuda_device, big):
    from hdbo_benchmark_b200 import _lib

    gp, (X, y, ls, os_, nz, mu) = big
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(3000, X.shape[1])
    best_f = float(y.max())
    mean, var, lei = gp.posterior_acq(Z.to(cuda_device), _lib.ACQ_LOGEI, best_f)
    m_o, v_o = O.posterior(st, Z)
    a_o = O.log_expected_improvement(m_o, v_o, best_f)
    assert float((mean[0].cpu() - m_o).abs().max()) < 1e-6 * float(m_o.abs().max())
    assert float(((var[0].cpu() - v_o).abs() / v_o).max()) < 1e-6
    assert float((lei[0].cpu() - a_o).abs().max()) < 1e-6 * float(a_o.abs().max())
    assert int(lei[0].argmax()) == int(a_o.argmax())
    assert abs(gp.logdet.item() - st.logdet) < 1e-9 * abs(st.logdet)


def test_fullsize_interpolation_identity_and_variance_bounds(cuda_device, big):
    from hdbo_benchmark_b200 import _lib

    gp, (X, y, ls, os_, nz, mu) = big
    n = X.shape[0]
    mean, var, _ = gp.posterior_acq(gp.X, _lib.ACQ_NONE, 0.0)
    want = gp.y - nz * gp.alpha[0, :n]            # (K + noise I) alpha = y - m  =>  m + K alpha = y - noise alpha
    assert float((mean[0] - want).abs().max()) < 1e-8
    assert float(var[0].min()) > -1e-9 and float(var[0].max()) < nz * (1 + 1e-6)   # 0 <= var(x_i) <= noise at training points
    Z = O.sobol_candidates(40000, X.shape[1], seed=11).to(cuda_device)            # 2 full chunks + a ragged one
    _, v, _ = gp.posterior_acq(Z, _lib.ACQ_NONE, 0.0)
    assert float(v.min()) > 0.0 and float(v.max()) <= os_ * (1 + 1e-12)


def test_fullsize_chunking_invariance_and_determinism(cuda_device, big):
    from hdbo_benchmark_b200 import _lib

    gp, (X, y, *_r) = big
    Z = O.sobol_candidates(20000, X.shape[1], seed=5).to(cuda_device)
    a = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
    b = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()))
    assert all(torch.equal(u, v) for u, v in zip(a, b))                            # bitwise run-to-run
    gp._ws = None
    c = gp.posterior_acq(Z, _lib.ACQ_LOGEI, float(y.max()), chunk_rows=4096)
    assert all(torch.equal(u, v) for u, v in zip(a, c))                            # bitwise across chunk sizes
    gp._ws = None


def test_fullsize_posterior_mean_is_linear_in_targets(cuda_device):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.engine import DeviceGP

    n, d = 4096, 256
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    g = torch.Generator().manual_seed(3)
    y2 = torch.randn(n, generator=g, dtype=torch.float64)
    Z = O.sobol_candidates(1024, d, seed=2).to(cuda_device)
    means = []
    for target in (y, y2, y + 2.0 * y2):
        gp = DeviceGP(X.to(cuda_device), target.to(cuda_device), O.MATERN52)
        gp.set_hyper(ls.to(cuda_device), os_, nz, 0.0)
        gp.fit()
        means.append(gp.posterior_acq(Z, _lib.ACQ_NONE, 0.0)[0][0])
    lin = means[0] + 2.0 * means[1]
    assert float((means[2] - lin).abs().max()) < 1e-8 * float(lin.abs().max())
