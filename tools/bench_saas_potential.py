"""SAAS NUTS potential + gradient: C chain states per call on the GPU vs one state at a time through CPU autograd (oracle)."""
import json, sys, time
sys.path.insert(0, ".")
import torch
from oracle import gp_oracle as O
from hdbo_benchmark_b200.saas import SaasPotential

for n, d, C in ((256, 128, 64), (512, 1024, 64)):
    X, y, *_ = O.synthetic_problem(n, d)
    g = torch.Generator().manual_seed(0)
    Z = torch.randn(C, d + 4, generator=g, dtype=torch.float64) * 0.5
    Z[:, d] -= 2.0; Z[:, d + 2] -= 3.0
    pot = SaasPotential(X, y)
    for _ in range(3):
        U, dU = pot(Z)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        U, dU = pot(Z)
    torch.cuda.synchronize(); gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    for c in range(2):
        z = Z[c].clone().requires_grad_(True)
        u = O.saas_potential(X, y, z); torch.autograd.grad(u, z)
    cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
    print(json.dumps({"config": f"SAAS potential+grad n={n} d={d}", "states_per_call": C, "gpu_ms_per_call": gpu_ms,
                      "gpu_us_per_state": gpu_ms / C * 1e3, "cpu_ms_per_state": cpu_ms}))
