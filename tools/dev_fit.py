"""Developer GPU check of the fit path: parity of L / L^-1 vs the oracle and timings at several n."""
import sys, json
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient
dev = torch.device("cuda")
def rel(a, b): return float((a.cpu() - b).abs().max() / b.abs().max())
for n, d in ((100, 8), (384, 32), (1024, 128)):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, keep_r2=True); gp.set_hyper(ls.to(dev), os_, nz, mu); gp.fit()
    print(n, d, "L", rel(gp.L[0, :n, :n].tril(), st.L), "Linv", rel(gp.Linv[0, :n, :n], st.Linv), "Linvt", rel(gp.Linvt[0, :n, :n], st.Linv.T), "alpha", rel(gp.alpha[0, :n], st.alpha), "info", gp.info.tolist())
for n, d in ((310, 128), (1024, 128), (4096, 256)):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
    gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, keep_r2=True); gp.set_hyper(ls.to(dev), os_, nz, mu)
    def t(fn, reps=5):
        for _ in range(2): fn()
        torch.cuda.synchronize(); best = 1e9
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        return best
    f = t(lambda: gp.fit_once(need_r2=True))
    fg = t(lambda: (gp.fit_once(need_r2=True), mll_gradient(gp)))
    print(json.dumps({"n": n, "d": d, "fit_ms": f, "fit_plus_grad_ms": fg}))
