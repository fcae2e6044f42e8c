"""Developer GPU check of the exact int8 (tcgen05) variance mode: parity vs the FP64 DMMA mode and the oracle, then timing."""
import sys, json
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP

dev = torch.device("cuda")


def check(n, d, m, kind, ls_mode="prior", noise=1e-3, chunk_rows=148 * 128):
    X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, ls_mode, noise)
    st = O.fit_state(X, y, ls, os_, nz, mu, kind)
    Z = O.sobol_candidates(m, d)
    Z[: min(m, n) // 2] = X[: min(m, n) // 2]          # candidates ON training points: worst-case cancellation in the variance
    mean_o, var_o = O.posterior(st, Z)
    res = {"n": n, "d": d, "m": m, "kind": kind, "ls": ls_mode, "noise": noise}
    outs = {}
    for mode in ("f64", "i8"):
        gp = DeviceGP(X.to(dev), y.to(dev), kind, var_mode=mode)
        gp.set_hyper(ls.to(dev), os_, nz, mu)
        gp.fit()
        mean, var, acq = gp.posterior_acq(Z.to(dev), _lib.ACQ_LOGEI, float(y.max()), chunk_rows=chunk_rows)
        torch.cuda.synchronize()
        outs[mode] = (mean[0].cpu(), var[0].cpu(), acq[0].cpu())
        res[mode + "_flag"] = gp.i8_flag()
        res[mode + "_mean_rel"] = float((mean[0].cpu() - mean_o).abs().max() / mean_o.abs().max())
        res[mode + "_var_rel"] = float(((var[0].cpu() - var_o).abs() / var_o.abs()).max())
        res[mode + "_var_abs_over_os"] = float((var[0].cpu() - var_o).abs().max() / float(os_))
    res["i8_vs_f64_var_abs_over_os"] = float((outs["i8"][1] - outs["f64"][1]).abs().max() / float(os_))
    res["argmax_same"] = int(outs["i8"][2].argmax() == outs["f64"][2].argmax())
    res["nan"] = int(torch.isnan(outs["i8"][1]).any())
    print(json.dumps(res), flush=True)


check(300, 20, 1000, O.MATERN52)
check(300, 20, 1000, O.RBF)
check(1024, 128, 4096, O.MATERN52)
check(1024, 128, 4096, O.MATERN52, "long", 1e-4)
check(700, 33, 515, O.RBF, "short")
check(1000, 64, 3000, O.MATERN52, "prior", 1e-6, chunk_rows=1024)

n, d, m = 4096, 256, 1 << 18
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d)
Z = torch.rand(m, d, dtype=torch.float64, device=dev)
lib = _lib.load()
for mode in ("f64", "i8"):
    gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52, var_mode=mode)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); gp.fit(); e[1].record()
    out = {"acq": torch.empty(1, m, dtype=torch.float64, device=dev)}
    for _ in range(2):
        gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False, out=out)
    gp.attach_timeline(4096)
    e[2].record()
    for _ in range(3):
        gp.posterior_acq(Z, _lib.ACQ_LOGEI, 1.0, want_mean=False, want_var=False, out=out)
    e[3].record(); torch.cuda.synchronize()
    tl = gp.collect_timeline()
    t = e[2].elapsed_time(e[3]) / 3
    print(json.dumps({"mode": mode, "fit_ms": e[0].elapsed_time(e[1]), "pool_ms": t, "evals_per_s": m / t * 1e3,
                      "phases_ms": {k: v[0] / 3 for k, v in tl.items()}, "flag": gp.i8_flag()}), flush=True)
