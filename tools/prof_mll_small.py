"""A few fast-closure MLL evaluations (fit + gradient) at n=310, d=128 for a per-kernel launch list."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from hdbo_benchmark_b200.fit import _FastClosure, _named_raw_parameters
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 310, 128
g = torch.Generator().manual_seed(0)
X = torch.rand(n, d, generator=g, dtype=torch.float64)
y = (torch.sin(3 * X[:, :4].sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
model = SingleTaskGP(X, y); mll = ExactMarginalLogLikelihood(model.likelihood, model); model.train()
params = _named_raw_parameters(mll)
fast = _FastClosure.build(mll, params)
x0 = np.concatenate([p.detach().cpu().numpy().reshape(-1) for _, p in params])
fast.graph_disabled = True
for _ in range(3):
    fast(x0)
torch.cuda.synchronize(); print("ok")
