"""GPU tests of the batched device L-BFGS (hdbo_lbfgs_propose / hdbo_lbfgs_accept and
optimize_acqf(options={"method": "device-lbfgs"})) against scipy's L-BFGS-B, which is what botorch's gen_candidates_scipy runs."""
import ctypes as C

import numpy as np
import pytest
import torch
from scipy.optimize import minimize

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def _run_device_lbfgs(fg, x0, lo, hi, M=10, T=8, maxiter=300, ftol=2.2e-9, gtol=1e-5):
    """Drive the two kernels with a torch value/gradient callback (fg: [m, D] -> ([m], [m, D]))."""
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200._lib import HdboLbfgs, check, ptr

    lib = _lib.load()
    dev = x0.device
    R, D = x0.shape
    f64 = dict(dtype=torch.float64, device=dev)
    x = x0.clone().contiguous()
    f, g = fg(x)
    f, g = f.clone(), g.clone()
    S = torch.zeros(R, M, D, **f64); Y = torch.zeros(R, M, D, **f64); rho = torch.zeros(R, M, **f64)
    meta = torch.zeros(R, 4, dtype=torch.int32, device=dev)
    step = torch.ones(R, **f64)
    Xt = torch.empty(R * T, D, **f64); ft = torch.empty(R * T, **f64); gt = torch.empty(R * T, D, **f64)
    st = HdboLbfgs(R=R, D=D, M=M, T=T, lo=ptr(lo), hi=ptr(hi), x=ptr(x), f=ptr(f), g=ptr(g), S=ptr(S), Y=ptr(Y), rho=ptr(rho),
                   step=ptr(step), meta=ptr(meta), Xt=ptr(Xt), ft=ptr(ft), gt=ptr(gt), ftol=ftol, gtol=gtol, maxiter=maxiter)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(maxiter):
        check(lib.hdbo_lbfgs_propose(C.byref(st), stream), "propose")
        a, b = fg(Xt)
        ft.copy_(a); gt.copy_(b)
        check(lib.hdbo_lbfgs_accept(C.byref(st), stream), "accept")
        if bool((meta[:, 2] > 0).all()):
            break
    return x, f, meta


@pytest.mark.parametrize("D", [3, 40, 300])
def test_device_lbfgs_box_constrained_quadratics_match_scipy(cuda_device, D):
    g = torch.Generator().manual_seed(D)
    R = 6
    A = torch.randn(R, D, D, generator=g, dtype=torch.float64)
    A = A @ A.transpose(-1, -2) / D + 0.05 * torch.eye(D, dtype=torch.float64)
    c = torch.rand(R, D, generator=g, dtype=torch.float64) * 1.6 - 0.3          # some optima outside the unit box
    x0 = torch.rand(R, D, generator=g, dtype=torch.float64)
    lo, hi = torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64)
    Ad, cd = A.to(cuda_device), c.to(cuda_device)

    def fg(X):                                   # X [R*T or R, D], problem index = row // (rows / R)
        per = X.shape[0] // R
        Xr = X.reshape(R, per, D) - cd[:, None, :]
        AX = torch.einsum("rij,rpj->rpi", Ad, Xr)
        return (0.5 * (Xr * AX).sum(-1)).reshape(-1), AX.reshape(-1, D)

    x, f, meta = _run_device_lbfgs(fg, x0.to(cuda_device), lo.to(cuda_device), hi.to(cuda_device))
    assert bool((meta[:, 2] > 0).all())
    for r in range(R):
        Ar, cr = A[r].numpy(), c[r].numpy()
        res = minimize(lambda v: (0.5 * (v - cr) @ Ar @ (v - cr), Ar @ (v - cr)), x0[r].numpy(), jac=True, method="L-BFGS-B",
                       bounds=[(0.0, 1.0)] * D, options={"maxiter": 2000, "ftol": 1e-14, "gtol": 1e-10})
        assert float(f[r]) <= res.fun + 1e-7 * max(1.0, abs(res.fun))
        assert np.abs(x[r].cpu().numpy() - res.x).max() < 2e-3
        assert float(x[r].min()) >= 0.0 and float(x[r].max()) <= 1.0


def test_optimize_acqf_device_lbfgs_matches_scipy_path(cuda_device):
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import Normalize, SingleTaskGP, Standardize
    from hdbo_benchmark_b200.optim import gen_batch_initial_conditions, gen_candidates_device, gen_candidates_scipy, optimize_acqf

    torch.manual_seed(0)                                  # the Boltzmann pick of the initial conditions uses the global generator
    n, d = 120, 10
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "short", 1e-3, seed=4)
    Xw = X * 4.0 - 2.0                                    # a box that is not the unit cube -> exercises Normalize
    bounds = torch.stack([-2.0 * torch.ones(d), 2.0 * torch.ones(d)]).double()
    model = SingleTaskGP(Xw, 3.0 * y.unsqueeze(-1) + 1.0, covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=Standardize(m=1), input_transform=Normalize(d, bounds=bounds))
    model.covar_module.base_kernel.lengthscale = ls / 4.0 * 1.5
    model.covar_module.outputscale = 1.0
    model.likelihood.noise = 1e-3
    model.eval()
    for acq in (A.LogExpectedImprovement(model, best_f=float((3.0 * y + 1.0).max())), A.UpperConfidenceBound(model, beta=4.0)):
        ics = gen_batch_initial_conditions(acq, bounds, 1, 12, 512, options={"seed": 0})
        c_s, v_s = gen_candidates_scipy(ics, acq, bounds[0], bounds[1], options={"maxiter": 300})
        c_d, v_d = gen_candidates_device(ics, acq, bounds[0], bounds[1], options={"maxiter": 300})
        assert c_d.shape == c_s.shape and v_d.shape == v_s.shape
        assert float(c_d.min()) >= -2.0 and float(c_d.max()) <= 2.0
        v_s, v_d = v_s.cpu().double(), v_d.cpu().double()
        # same starting points, same objective: the best restart must be as good.  Individual restarts may settle in different
        # basins of the multi-modal surface (scipy optimises the SUM over restarts with one shared L-BFGS memory and step
        # length, here every restart has its own), so per restart the criterion is local optimality: monotone improvement and
        # a vanishing projected gradient.
        assert float(v_d.max()) >= float(v_s.max()) - 1e-6 * max(1.0, abs(float(v_s.max())))
        assert bool((v_d >= acq(ics).cpu().double() - 1e-12).all())            # never worse than the start
        Xg = c_d.detach().clone().requires_grad_(True)
        (grad,) = torch.autograd.grad(acq(Xg).sum(), Xg)
        grad, xc = grad.reshape(-1, d).cpu().double(), c_d.reshape(-1, d).cpu().double()
        pg = torch.where((xc <= -2.0 + 1e-12) & (grad < 0), torch.zeros_like(grad), grad)       # ascent: blocked at the lower bound
        pg = torch.where((xc >= 2.0 - 1e-12) & (grad > 0), torch.zeros_like(pg), pg)
        Xs = c_s.detach().clone().requires_grad_(True)
        (gs,) = torch.autograd.grad(acq(Xs).sum(), Xs)
        gs, xs = gs.reshape(-1, d).cpu().double(), c_s.reshape(-1, d).cpu().double()
        pgs = torch.where((xs <= -2.0 + 1e-12) & (gs < 0), torch.zeros_like(gs), gs)
        pgs = torch.where((xs >= 2.0 - 1e-12) & (gs > 0), torch.zeros_like(pgs), pgs)
        assert float(pg.abs().max()) <= max(1e-3, 2.0 * float(pgs.abs().max())), (pg.abs().max(), pgs.abs().max())
    cand, val = optimize_acqf(acq, bounds, q=1, num_restarts=8, raw_samples=256, options={"method": "device-lbfgs", "seed": 1})
    cand2, val2 = optimize_acqf(acq, bounds, q=1, num_restarts=8, raw_samples=256, options={"seed": 1})
    from hdbo_benchmark_b200.sampling import draw_sobol_samples
    with torch.no_grad():
        raw_best = float(acq(draw_sobol_samples(bounds, n=256, q=1, seed=1)).max())
    # the arg-max raw sample is always among the initial conditions and the optimiser is monotone; the scipy path reaches a
    # nearby (not necessarily the same) optimum from the same initial conditions
    assert cand.shape == (1, d) and float(val) >= raw_best - 1e-12 and float(val2) >= raw_best - 1e-12
    assert abs(float(val) - float(val2)) <= 0.05 * max(1.0, abs(float(val2)))


@pytest.mark.parametrize("flavour", ["botorch_default", "legacy", "in_repo_rbf_scale"])
def test_device_fit_reaches_the_scipy_optimum(cuda_device, flavour):
    """fit_gpytorch_mll(..., options={"method": "device-lbfgs"}): batched-trial L-BFGS on the device (every line search = one batched
    hdbo_gp_fit + hdbo_mll_grad) against the scipy L-BFGS-B path from the same start."""
    import math

    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.fit import fit_gpytorch_mll, fit_gpytorch_mll_device
    from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
    from oracle import gp_oracle as O

    n, d = 90, 7
    X, y, *_ = O.synthetic_problem(n, d, "prior", 1e-3, seed=6)
    Y = 2.0 * y.unsqueeze(-1) - 1.0

    def make():
        if flavour == "botorch_default":
            return SingleTaskGP(X, Y)
        if flavour == "legacy":
            return SingleTaskGP(X, Y, legacy_defaults=True)
        k = G.ScaleKernel(G.RBFKernel(ard_num_dims=d, lengthscale_prior=G.LogNormalPrior(0.5 * math.log(d), 1.0)))
        return SingleTaskGP(X, Y, covar_module=k)

    m_dev, m_ref = make(), make()
    fit_gpytorch_mll(ExactMarginalLogLikelihood(m_dev.likelihood, m_dev), optimizer_kwargs={"options": {"method": "device-lbfgs"}})
    info = fit_gpytorch_mll_device.last_info
    assert info["flag"] in (1, 2) and info["cuda_graph"] and info["iterations"] > 3
    fit_gpytorch_mll(ExactMarginalLogLikelihood(m_ref.likelihood, m_ref))
    vals = []
    for mdl in (m_dev, m_ref):
        mdl.train()
        mm = ExactMarginalLogLikelihood(mdl.likelihood, mdl)
        with torch.no_grad():
            vals.append(float(mm(mdl(mdl.train_inputs[0]), mdl.train_targets).sum()))
        mdl.eval()
    # device descent + scipy polish from its answer vs scipy from the start: same ftol-flat neighbourhood
    assert vals[0] > vals[1] - 1e-4 * max(1.0, abs(vals[1])), vals
    Zt = torch.rand(64, d, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    pd, pr = m_dev.posterior(Zt), m_ref.posterior(Zt)
    assert float((pd.mean - pr.mean).abs().max()) < 5e-2 * float(pr.mean.abs().max())
    for p in m_dev.parameters():
        assert bool(torch.isfinite(p).all())


@pytest.mark.parametrize("q", [2, 4])
def test_device_lbfgs_on_qei_matches_scipy_path(cuda_device, q):
    """q > 1: every restart is a (q*d)-variable problem; value + gradient of qExpectedImprovement come from
    hdbo_posterior_fwd -> hdbo_posterior_joint_cov -> hdbo_qei -> hdbo_posterior_bwd without the autograd tape."""
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.models import Normalize, SingleTaskGP, Standardize
    from hdbo_benchmark_b200.optim import gen_candidates_device, gen_candidates_scipy, optimize_acqf
    from hdbo_benchmark_b200.sampling import SobolQMCNormalSampler

    n, d, R = 150, 8, 10
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "short", 1e-3, seed=2)
    bounds = torch.stack([-1.0 * torch.ones(d), 3.0 * torch.ones(d)]).double()
    model = SingleTaskGP(X * 4.0 - 1.0, 2.0 * y.unsqueeze(-1) - 3.0, covar_module=G.ScaleKernel(G.MaternKernel(ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=Standardize(m=1), input_transform=Normalize(d, bounds=bounds))
    model.covar_module.base_kernel.lengthscale = ls / 4.0 * 1.5
    model.covar_module.outputscale = 1.0
    model.likelihood.noise = 1e-3
    model.eval()
    acq = A.qExpectedImprovement(model, best_f=float((2.0 * y - 3.0).max()) - 0.3,
                                 sampler=SobolQMCNormalSampler(sample_shape=torch.Size([128]), seed=1))
    ics = (O.sobol_candidates(R * q, d, seed=8) * 4.0 - 1.0).reshape(R, q, d).to(cuda_device)
    # the tape-free evaluator equals the autograd surface
    gp = model._device_gp()
    Zt = ((ics.reshape(-1, d) + 1.0) / 4.0).contiguous()
    base = acq.sampler.base_samples(q, cuda_device)
    ot = model.outcome_transform
    sc, sh = float(ot.stdvs.reshape(())), float(ot.means.reshape(()))
    v_raw, dZ = gp.qei_value_grad(Zt, q, base, (float(acq.best_f) - sh) / sc, acq.jitter)
    Xg = ics.clone().requires_grad_(True)
    v_auto = acq(Xg)
    (g_auto,) = torch.autograd.grad(v_auto.sum(), Xg)
    assert float((v_raw * sc - v_auto).abs().max()) < 1e-12 * max(1.0, float(v_auto.abs().max()))
    assert float((dZ.reshape(R, q, d) * sc / 4.0 - g_auto).abs().max()) < 1e-10 * max(1e-3, float(g_auto.abs().max()))
    # optimiser: same starts, same objective
    c_s, v_s = gen_candidates_scipy(ics, acq, bounds[0], bounds[1], options={"maxiter": 200})
    c_d, v_d = gen_candidates_device(ics, acq, bounds[0], bounds[1], options={"maxiter": 200})
    assert c_d.shape == (R, q, d) and v_d.shape == (R,)
    assert float(c_d.min()) >= -1.0 and float(c_d.max()) <= 3.0
    with torch.no_grad():
        v0 = acq(ics)
    assert bool((v_d >= v0 - 1e-12).all())                                    # never worse than the start
    # Both optimisers are LOCAL on a piecewise-smooth MC objective: from the same start they may settle in different basins (measured:
    # scipy's own answer for one start moved between 0.23 and 0.54 across runs that differ in the last bit of alpha), so the values
    # are not comparable restart by restart.  What must hold: every device answer is a stationary point of the box-constrained
    # problem -- which scipy's ftol stop does not even reach (its projected gradients end at 1e-3 .. 1e-2) -- and the two sets of
    # answers are of the same quality overall.
    Xs_ = c_d.clone().to(cuda_device).requires_grad_(True)
    (g_d,) = torch.autograd.grad(acq(Xs_).sum(), Xs_)
    lo_b, hi_b = bounds[0].to(cuda_device), bounds[1].to(cuda_device)
    pg = g_d.clone()
    pg[(Xs_ <= lo_b + 1e-12) & (g_d < 0)] = 0.0                               # maximisation: ascent direction leaves the box
    pg[(Xs_ >= hi_b - 1e-12) & (g_d > 0)] = 0.0
    info = gen_candidates_device.last_info
    iters = torch.as_tensor(info["per_restart_iterations"])
    conv = iters < 200
    assert bool(conv.any()) and float(pg.reshape(R, -1).abs().max(dim=1).values[conv.to(pg.device)].max()) < 1e-3
    assert float(v_d.mean()) >= 0.8 * float(v_s.mean()) and float(v_d.max()) >= 0.8 * float(v_s.max())
    assert info["cuda_graph"] and max(info["per_restart_iterations"]) <= 200
    # and through optimize_acqf
    xq, vq = optimize_acqf(acq, bounds, q=q, num_restarts=6, raw_samples=64, options={"method": "device-lbfgs", "seed": 3, "maxiter": 100})
    assert xq.shape == (q, d) and float(vq) >= 0.0
