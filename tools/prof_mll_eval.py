"""cProfile of repeated MLL value+gradient evaluations through the public surface (host overhead hunt)."""
import cProfile, pstats, sys, time, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
n, d = 310, 128
g = torch.Generator().manual_seed(0)
X = torch.rand(n, d, generator=g, dtype=torch.float64)
y = (torch.sin(3 * X[:, :4].sum(-1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
model = SingleTaskGP(X, y); mll = ExactMarginalLogLikelihood(model.likelihood, model); model.train()
params = [p for p in model.parameters()]
def one():
    for p in params: p.grad = None
    with torch.no_grad(): params[0].add_(1e-9)
    loss = -mll(model(model.train_inputs[0]), model.train_targets)
    loss.backward()
    return torch.cat([loss.detach().reshape(1)] + [p.grad.reshape(-1) for p in params]).cpu()
for _ in range(5): one()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(50): one()
torch.cuda.synchronize(); print("ms per eval:", (time.perf_counter() - t) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): one()
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
