"""Developer GPU check: CUDA path vs oracle on a few shapes + first timings.  Not part of the test suite."""
import sys, time, json
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP

torch.manual_seed(0)
dev = torch.device("cuda")

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))

def check(n, d, m, kind, ls_mode="prior", noise=1e-3):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, ls_mode, noise)
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    Z = O.sobol_candidates(m, d)
    gp = DeviceGP(X.to(dev), y.to(dev), kind)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()
    res = {"n": n, "d": d, "m": m, "kind": kind, "ls": ls_mode}
    res["L"] = rel(gp.L[0, :n, :n].cpu().tril(), st.L)
    res["Linv"] = rel(gp.Linv[0, :n, :n].cpu(), st.Linv)
    res["alpha"] = rel(gp.alpha[0, :n].cpu(), st.alpha)
    res["logdet"] = abs(gp.logdet.item() - st.logdet) / abs(st.logdet)
    mean_o, var_o = O.posterior(st, Z)
    best_f = float(y.max())
    for kname, kk, prm in (("ei", _lib.ACQ_EI, best_f), ("logei", _lib.ACQ_LOGEI, best_f), ("ucb", _lib.ACQ_UCB, 4.0)):
        mean, var, acq = gp.posterior_acq(Z.to(dev), kk, prm)
        a_o = O.acquisition(mean_o, var_o, kk, prm)
        res["mean"] = rel(mean[0].cpu(), mean_o)
        res["var"] = float(((var[0].cpu() - var_o).abs() / var_o.abs().clamp_min(1e-10)).max())
        res[kname] = float(((acq[0].cpu() - a_o).abs() / a_o.abs().clamp_min(1e-300)).max()) if kname != "ei" else rel(acq[0].cpu(), a_o)
        res[kname + "_argmax_same"] = int(acq[0].argmax().item() == a_o.argmax().item())
        v, i = gp.argmax(acq[0]); res[kname + "_argmax_kernel_ok"] = int(i.item() == acq[0].argmax().item())
    print(json.dumps(res))

# dgemm mainloop vs torch
lib = _lib.load()
A = torch.randn(256, 64, dtype=torch.float64, device=dev); B = torch.randn(384, 64, dtype=torch.float64, device=dev)
Cm = torch.zeros(256, 384, dtype=torch.float64, device=dev)
_lib.check(lib.hdbo_dgemm_nt(A.data_ptr(), 64, B.data_ptr(), 64, Cm.data_ptr(), 384, 256, 384, 64, 1.0, 0.0, 0, 0, torch.cuda.current_stream().cuda_stream), "dgemm")
print("dgemm rel err", rel(Cm, A @ B.T))

check(300, 20, 1000, O.MATERN52)
check(300, 20, 1000, O.RBF)
check(1024, 128, 4096, O.MATERN52)
check(1024, 128, 4096, O.MATERN52, "long", 1e-4)
check(700, 33, 515, O.RBF, "short")

# timing at the headline shape
n, d, m = 4096, 256, 1 << 17
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52)
gp.set_hyper(ls.to(dev), os_, nz, mu)
for _ in range(2): gp.fit_once()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); gp.fit_once(); e1.record(); torch.cuda.synchronize()
print("fit n=4096 d=256 ms:", e0.elapsed_time(e1), "info", gp.info.tolist())
gp.fit()
Z = O.sobol_candidates(m, d).to(dev)
best_f = float(y.max())
for _ in range(2): gp.posterior_acq(Z, _lib.ACQ_LOGEI, best_f)
torch.cuda.synchronize()
e0.record(); mean, var, acq = gp.posterior_acq(Z, _lib.ACQ_LOGEI, best_f); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = m * (2.0 * n * d + n * n + 2 * n)
print(f"posterior+LogEI m={m}: {ms:.2f} ms  {m/ms*1e3:.0f} evals/s  {fl/ms/1e9:.2f} TFLOP/s ({fl/ms/1e9/37.0:.3f} of DMMA peak)")
# check a slice against the oracle
st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
mo, vo = O.posterior(st, Z[:2048].cpu())
print("n=4096 mean rel", rel(mean[0, :2048].cpu(), mo), "var rel", float(((var[0, :2048].cpu() - vo).abs() / vo).max()),
      "Linv rel", rel(gp.Linv[0].cpu(), st.Linv), "alpha rel", rel(gp.alpha[0].cpu(), st.alpha))
