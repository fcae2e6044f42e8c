"""Thompson-sampling step at TuRBO scale (n = 1000, d = 128, 5000 Sobol candidates, 10 joint samples): GPU path vs CPU oracle."""
import json, sys, time
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.thompson import joint_posterior_samples

dev = torch.device("cuda")
n, d, m, S = 1000, 128, 5000, 10
X, y, ls, os_, nz, mu = O.synthetic_problem(n, d, "prior", 1e-3, seed=1)
gp = DeviceGP(X.to(dev), y.to(dev), O.MATERN52)
gp.set_hyper(ls.to(dev), os_, nz, mu); gp.fit()
Z = O.sobol_candidates(m, d, seed=2)
z = torch.randn(S, m, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
Zd = Z.to(dev)
for _ in range(2):
    s, mean, jit = joint_posterior_samples(gp, Zd, z)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    s, mean, jit = joint_posterior_samples(gp, Zd, z)
torch.cuda.synchronize(); gpu_ms = (time.perf_counter() - t0) / 3 * 1e3
st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
t0 = time.perf_counter()
mean_o, Sig = O.posterior_joint(st, Z.unsqueeze(0))
Lo = torch.linalg.cholesky(Sig[0] + jit * torch.eye(m, dtype=torch.float64))
so = mean_o[0] + z @ Lo.T
cpu_ms = (time.perf_counter() - t0) * 1e3
err = float((s.cpu() - so).abs().max() / so.abs().max())
print(json.dumps({"config": "Thompson sampling n=1000 d=128 m=5000 candidates, 10 joint samples", "gpu_ms": gpu_ms, "cpu_ms": cpu_ms,
                  "jitter": jit, "max_rel_sample_diff": err, "same_argmax": bool((s.cpu().argmax(1) == so.argmax(1)).all())}))
