"""TEST HARNESS (not product code).  Stand-ins for the poli-baselines solver classes the reference selects
(/root/reference/src/hdbo_benchmark/utils/experiments/load_solvers.py:118-123 `vanilla_bo`, :124-145
`vanilla_bo_with_lognormal_prior`, :159-182 `random_line_bo` / `coordinate_line_bo`).

poli / poli-baselines are not installable offline, so these ~100 lines restate the solver contract the drop-in must
survive -- `solve(max_iter)` loops `step()`; `step()` is `x = next_candidate(); y = black_box(x); update(x, y);
post_update(x, y); iteration += 1` (the in-repo statement of it:
/root/reference/src/hdbo_benchmark/utils/latent_space_solver.py:42-63) -- and build their GP / acquisition /
optimiser exactly through the BoTorch-shaped surface of this package, so `step()` exercises the whole hot path:
SingleTaskGP -> fit_gpytorch_mll -> (Log)ExpectedImprovement(best_f = y.max()) -> optimize_acqf.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from hdbo_benchmark_b200.acquisition import ExpectedImprovement, LogExpectedImprovement, UpperConfidenceBound
from hdbo_benchmark_b200.fit import fit_gpytorch_mll
from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood, SingleTaskGP
from hdbo_benchmark_b200.optim import optimize_acqf


class AbstractSolver:
    def __init__(self, black_box: Callable, x0: np.ndarray, y0: np.ndarray):
        self.black_box = black_box
        self.x0, self.y0 = np.asarray(x0, dtype=np.float64), np.asarray(y0, dtype=np.float64).reshape(len(x0), -1)
        self.history = {"x": [self.x0[i:i + 1] for i in range(len(self.x0))], "y": [self.y0[i:i + 1] for i in range(len(self.y0))]}
        self.iteration = 0

    def next_candidate(self) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def update(self, x, y):
        self.history["x"].append(np.asarray(x, dtype=np.float64).reshape(1, -1))
        self.history["y"].append(np.asarray(y, dtype=np.float64).reshape(1, -1))

    def post_update(self, x, y):
        pass

    def step(self):
        x = self.next_candidate()
        y = self.black_box(x)
        self.update(x, y)
        self.post_update(x, y)
        self.iteration += 1
        return x, y

    def solve(self, max_iter: int = 100, verbose: bool = False):
        for _ in range(max_iter):
            x, y = self.step()
            if verbose:
                print(f"iteration {self.iteration}: y = {float(np.asarray(y).reshape(-1)[0]):.4f}, best = {self.get_best_performance():.4f}")
        return self.get_best_solution()

    def get_best_performance(self) -> float:
        return float(np.nanmax(np.concatenate(self.history["y"])))

    def get_best_solution(self) -> np.ndarray:
        ys = np.concatenate(self.history["y"]).reshape(-1)
        return np.concatenate(self.history["x"])[int(np.nanargmax(ys))]


class BaseBayesianOptimization(AbstractSolver):
    def __init__(self, black_box, x0, y0, bounds=(0.0, 1.0), kernel=None, acq_function: str = "log_expected_improvement",
                 penalize_nans_with: float = -10.0, num_restarts: int = 20, raw_samples: int = 100, seed: Optional[int] = None):
        super().__init__(black_box, x0, y0)
        d = self.x0.shape[1]
        lo, hi = bounds
        self.bounds = torch.tensor([[float(lo)] * d, [float(hi)] * d], dtype=torch.float64)
        self.kernel = kernel
        self.acq_name = acq_function
        self.penalize_nans_with = penalize_nans_with
        self.num_restarts, self.raw_samples, self.seed = num_restarts, raw_samples, seed
        self.model = None

    def _training_data(self):
        X = torch.from_numpy(np.concatenate(self.history["x"]))
        y = np.concatenate(self.history["y"]).copy()
        y[np.isnan(y)] = self.penalize_nans_with
        return X, torch.from_numpy(y)

    def _fit_model(self):
        import copy
        X, Y = self._training_data()
        kwargs = {} if self.kernel is None else {"covar_module": copy.deepcopy(self.kernel)}
        model = SingleTaskGP(X, Y, **kwargs)
        mll = ExactMarginalLogLikelihood(model.likelihood, model)
        fit_gpytorch_mll(mll)
        model.eval()
        self.model = model
        return model, float(Y.max())

    def _instantiate_acquisition_function(self, model, best_f):
        name = self.acq_name.lower()
        if name in ("expected_improvement", "ei"):
            return ExpectedImprovement(model, best_f=best_f)
        if name in ("log_expected_improvement", "logei", "log_ei"):
            return LogExpectedImprovement(model, best_f=best_f)
        if name in ("ucb", "upper_confidence_bound"):
            return UpperConfidenceBound(model, beta=2.0)
        raise ValueError(f"unknown acquisition function {self.acq_name}")


class VanillaBayesianOptimization(BaseBayesianOptimization):
    def next_candidate(self) -> np.ndarray:
        model, best_f = self._fit_model()
        acqf = self._instantiate_acquisition_function(model, best_f)
        opts = {"nonnegative": isinstance(acqf, ExpectedImprovement) and not isinstance(acqf, LogExpectedImprovement)}
        if self.seed is not None:
            opts["sobol_stream"] = self.seed   # one cached scrambled-Sobol sequence per solver: reproducible, fresh points each step
        cand, _ = optimize_acqf(acqf, self.bounds, q=1, num_restarts=self.num_restarts, raw_samples=self.raw_samples, options=opts)
        return cand.detach().cpu().numpy().reshape(1, -1)


class LineBO(BaseBayesianOptimization):
    """Acquisition maximised on a 1-d grid along a random or coordinate line through the incumbent."""

    def __init__(self, *args, type_of_line: str = "random", n_points_in_acq_grid: int = 100, **kwargs):
        super().__init__(*args, **kwargs)
        self.type_of_line = type_of_line
        self.n_grid = n_points_in_acq_grid
        self._rng = np.random.default_rng(self.seed)

    def next_candidate(self) -> np.ndarray:
        model, best_f = self._fit_model()
        acqf = self._instantiate_acquisition_function(model, best_f)
        d = self.x0.shape[1]
        x_best = self.get_best_solution()
        if self.type_of_line == "coordinate":
            direction = np.zeros(d)
            direction[self._rng.integers(d)] = 1.0
        else:
            direction = self._rng.normal(size=d)
            direction /= np.linalg.norm(direction)
        lo, hi = self.bounds[0].numpy(), self.bounds[1].numpy()
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (lo - x_best) / direction
            t2 = (hi - x_best) / direction
        tmin = np.nanmax(np.where(direction != 0, np.minimum(t1, t2), -np.inf))
        tmax = np.nanmin(np.where(direction != 0, np.maximum(t1, t2), np.inf))
        ts = np.linspace(tmin, tmax, self.n_grid)
        grid = np.clip(x_best[None, :] + ts[:, None] * direction[None, :], lo, hi)
        with torch.no_grad():
            vals = acqf(torch.from_numpy(grid).unsqueeze(1))
        return grid[int(torch.argmax(vals))].reshape(1, -1)
