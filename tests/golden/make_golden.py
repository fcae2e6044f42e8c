"""Generates tests/golden/gp_golden_{matern52,rbf}.npz.

The reference pins no numeric result (its only test is a no-exception smoke test) and BoTorch/GPyTorch cannot be
installed offline, so the golden vectors come from (a) the CPU oracle restatement (oracle/gp_oracle.py) and (b) an
INDEPENDENT implementation that is installed: scikit-learn's GaussianProcessRegressor (log marginal likelihood, its
gradient, predictive mean/std), stored side by side.  tests/test_golden.py checks the oracle against both sets on
the CPU and the CUDA path against them on the GPU.

    python tests/golden/make_golden.py        # rewrites the .npz files (deterministic)
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import gp_oracle as O  # noqa: E402

torch.set_default_dtype(torch.float64)


def make(kind, name, n=96, d=6, m=64, q=3, R=5, S=128):
    g = torch.Generator().manual_seed(1234 + kind)
    X = torch.rand(n, d, generator=g)
    y = torch.sin(3 * X[:, :3].sum(-1)) + 0.1 * torch.randn(n, generator=g)
    y = (y - y.mean()) / y.std()
    ls = 0.4 + torch.rand(d, generator=g)
    os_, noise, mean = 1.6, 5e-3, 0.15
    Z = torch.rand(m, d, generator=g)
    Z[0] = X[5] + 1e-4
    best_f = float(y.max()) - 0.2
    st = O.fit_state(X, y, ls, os_, noise, mean, kind)
    mu, var = O.posterior(st, Z)
    mll, g_ls, g_os, g_nz, g_mu = O.exact_mll_and_grad(X, y, ls, os_, noise, mean, kind)
    ei_v, ei_g = O.acquisition_value_and_grad_autograd(st, Z[:16], O.ACQ_EI, best_f)
    lei_v, lei_g = O.acquisition_value_and_grad_autograd(st, Z[:16], O.ACQ_LOGEI, best_f)
    ucb_v, ucb_g = O.acquisition_value_and_grad_autograd(st, Z[:16], O.ACQ_UCB, 2.5)
    Zq = torch.rand(R, q, d, generator=g)
    eng = torch.quasirandom.SobolEngine(q, scramble=True, seed=17)
    u = eng.draw(S, dtype=torch.float64)
    base = torch.erfinv(2.0 * (0.5 + (1 - 2.3e-16) * (u - 0.5)) - 1.0) * math.sqrt(2.0)
    mq, Sq = O.posterior_joint(st, Zq)
    qv, qg = O.qei_value_and_grad_autograd(st, Zq, base, best_f, jitter=1e-8)
    # independent pins: scikit-learn
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

    basek = Matern(length_scale=ls.numpy(), nu=2.5) if kind == O.MATERN52 else RBF(length_scale=ls.numpy())
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(os_) * basek + WhiteKernel(noise), optimizer=None, alpha=0.0)
    gpr.fit(X.numpy(), (y - mean).numpy())
    lml, lml_grad = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    sk_mean, sk_std = gpr.predict(Z.numpy(), return_std=True)
    np.savez_compressed(
        Path(__file__).parent / f"gp_golden_{name}.npz",
        kind=kind, X=X.numpy(), y=y.numpy(), lengthscale=ls.numpy(), outputscale=os_, noise=noise, mean_const=mean,
        Z=Z.numpy(), best_f=best_f, ucb_beta=2.5,
        alpha=st.alpha.numpy(), logdet=st.logdet, post_mean=mu.numpy(), post_var=var.numpy(),
        ei=O.expected_improvement(mu, var, best_f).numpy(), logei=O.log_expected_improvement(mu, var, best_f).numpy(),
        ucb=O.upper_confidence_bound(mu, var, 2.5).numpy(),
        mll=mll.item(), mll_g_ls=g_ls.numpy(), mll_g_os=g_os.item(), mll_g_noise=g_nz.item(), mll_g_mean=g_mu.item(),
        ei_grad=ei_g.numpy(), logei_grad=lei_g.numpy(), ucb_grad=ucb_g.numpy(),
        Zq=Zq.numpy(), base_samples=base.numpy(), joint_mean=mq.numpy(), joint_cov=Sq.numpy(), qei=qv.numpy(), qei_grad=qg.numpy(),
        sk_lml=lml, sk_lml_grad_logtheta=lml_grad, sk_mean=sk_mean + mean, sk_var_with_noise=sk_std ** 2,
    )


if __name__ == "__main__":
    make(O.MATERN52, "matern52")
    make(O.RBF, "rbf")
    print("golden vectors written")
