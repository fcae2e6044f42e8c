"""TEST HARNESS (not product code): an Adam driver for the exact MLL with hold-out early stopping.

It exercises the same sequence of model calls as the reference's in-repo trainer
(/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py:53-131) -- train-mode `mll(model(x), y).backward()`,
eval-mode `mll(model(x_test), y_test)` every iteration, best `state_dict` kept -- so that the GPU tests cover the call pattern the
drop-in has to survive.  On a machine with gpytorch's names shimmed (`hdbo_benchmark_b200.shims.install()`) the reference's own
function is what runs; this file exists because the reference tree is not available on the GPU test box.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import torch

from hdbo_benchmark_b200.models import ExactMarginalLogLikelihood


@dataclass
class HoldOutMonitor:
    """Tracks the best hold-out loss; `stale` counts the evaluations since it last improved."""

    patience: int
    best: float = float("inf")
    stale: int = 0
    snapshot: dict = field(default_factory=dict)

    def observe(self, value: float, modules: dict) -> bool:
        """Returns True when `value` is a new best (the modules' states are then snapshotted)."""
        if value < self.best:
            self.best, self.stale = value, 0
            self.snapshot = {name: copy.deepcopy(mod.state_dict()) for name, mod in modules.items()}
            return True
        self.stale += 1
        return False

    @property
    def exhausted(self) -> bool:
        return self.stale > self.patience

    def restore(self, modules: dict) -> None:
        for name, mod in modules.items():
            if name in self.snapshot:
                mod.load_state_dict(self.snapshot[name])


def _set_mode(training: bool, *modules) -> None:
    for m in modules:
        m.train(training)


def _holdout_loss(mll, model, x, y) -> float:
    """Negative eval-mode MLL per datum on the hold-out set; +inf when the evaluation fails for any reason."""
    _set_mode(False, model, model.likelihood, mll)
    try:
        with torch.no_grad():
            return float(-mll(model(x), y))
    except Exception:
        return float("inf")
    finally:
        _set_mode(True, model, model.likelihood, mll)


def fit_with_adam(model, train, holdout, *, lr: float = 0.05, patience: int = 50, max_iterations: int = 1000,
                  checkpoint: Optional[Path] = None):
    """train / holdout: (x, y) pairs.  Returns (iterations run, diverged flag); the model ends in eval mode holding its best state."""
    likelihood = model.likelihood
    mll = ExactMarginalLogLikelihood(likelihood, model)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    monitor = HoldOutMonitor(patience)
    tracked = {"model": model, "likelihood": likelihood}
    _set_mode(True, model, likelihood, mll)
    diverged, done = False, 0
    for done in range(1, max_iterations + 1):
        opt.zero_grad()
        loss = -mll(model(train[0]), train[1])
        loss.backward()
        if monitor.observe(_holdout_loss(mll, model, *holdout), tracked) and checkpoint is not None:
            Path(checkpoint).parent.mkdir(parents=True, exist_ok=True)
            torch.save(monitor.snapshot["model"], checkpoint)
        if monitor.exhausted:
            break
        if not torch.isfinite(loss):
            diverged = True
            break
        opt.step()
    monitor.restore(tracked)
    _set_mode(False, model, likelihood)
    return done, diverged
