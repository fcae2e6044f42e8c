"""Host-side mirror of the reference's only in-repo GPyTorch driver,
/root/reference/src/hdbo_benchmark/utils/experiments/training_gps.py: `filter_close_points` (:11-29) and
`train_exact_gp_model` (:32-147), same names, arguments, return values and error behaviour, with the arithmetic on
the device: the pairwise distances come from `hdbo_cdist` (the same GEMM-expansion the kernel build uses), the
train loss from `ExactMarginalLogLikelihood` (analytic CUDA gradient) and the eval-mode test loss from the joint
predictive log-density (`predictive.py`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr
from .engine import _rup, _stream
from .models import ExactMarginalLogLikelihood, compute_device


def pairwise_distances(x: torch.Tensor) -> torch.Tensor:
    """torch.cdist(x, x) on the device through hdbo_cdist; returns [n, n] on x's device."""
    lib = _lib.load()
    dev = compute_device()
    xd = x.detach().to(dev, torch.float64).contiguous()
    n, d = xd.shape
    npad, dpad = _rup(n, 128), _rup(d, 16)
    Xs = torch.empty(npad, dpad, dtype=torch.float64, device=dev)
    xn = torch.empty(npad, dtype=torch.float64, device=dev)
    D = torch.empty(npad, npad, dtype=torch.float64, device=dev)
    check(lib.hdbo_cdist(ptr(xd), xd.stride(0), n, d, ptr(Xs), ptr(xn), ptr(D), _stream()), "hdbo_cdist")
    return D[:n, :n].to(x.device, x.dtype if x.dtype.is_floating_point else torch.float64)


def filter_close_points(x, y, tolerance=1e-2):
    """Keep one representative of every group of points closer than `tolerance` (reference semantics: a point is
    dropped when an earlier-indexed point lies within the tolerance)."""
    dist = pairwise_distances(x)
    mask = torch.triu(dist < tolerance, diagonal=1)
    dropped = mask.any(dim=0)  # column j is dropped if some i < j is close to it
    keep = (~dropped).nonzero().reshape(-1)
    return x[keep], y[keep]


def train_exact_gp_model(model, likelihood, train_x, train_y, test_x, test_y, n_points: int, n_dimensions: int,
                         patience: int = 50, lr: float = 0.05, model_name: str = "model", max_nr_iterations: int = 1000,
                         verbose: bool = True, models_dir: Optional[Path] = None) -> Tuple[object, object, int, bool]:
    """Adam + early stopping on the exact MLL, test MLL every step, best state_dict checkpointed."""
    model.train()
    likelihood.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    mll = ExactMarginalLogLikelihood(likelihood, model)
    training_had_nans = False
    best_model_state, best_lik_state = None, None
    best_loss = float("inf")
    patience_counter = 0
    i = 0
    for i in range(max_nr_iterations):
        optimizer.zero_grad()
        output = model(train_x)
        loss = -mll(output, train_y)
        loss.backward()
        with torch.no_grad():
            model.eval(); likelihood.eval(); mll.eval()
            try:
                test_output = model(test_x)
                test_loss = -mll(test_output, test_y)
            except Exception as e:  # the reference tolerates any failure of the test loss
                if verbose:
                    print(f"Exception {e} occurred at test iteration {i + 1}.")
                test_loss = torch.tensor(float("inf"))
            model.train(); likelihood.train(); mll.train()
        if test_loss < best_loss:
            best_loss = test_loss
            patience_counter = 0
            best_model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
            best_lik_state = {k: v.detach().clone() for k, v in likelihood.state_dict().items()}
            if models_dir is not None:
                Path(models_dir).mkdir(parents=True, exist_ok=True)
                torch.save(best_model_state, Path(models_dir) / f"model_{model_name}_npoints_{n_points}_ndim_{n_dimensions}.pt")
        else:
            patience_counter += 1
        if patience_counter > patience:
            if verbose:
                print(f"Patience reached at iteration {i + 1}. Stopping training early.")
            break
        if torch.isnan(loss):
            if verbose:
                print(f"Loss is NaN at iteration {i + 1}.")
            training_had_nans = True
            break
        optimizer.step()
        if verbose:
            print(f"Nr. points: {n_points} - Dim: {n_dimensions} - Iteration {i + 1}/{max_nr_iterations} - "
                  f"Train loss: {loss.item()} - Test loss: {float(test_loss)}")
    if best_model_state is not None:
        model.load_state_dict(best_model_state)
        likelihood.load_state_dict(best_lik_state)
    model.eval()
    likelihood.eval()
    return model, likelihood, i, training_had_nans
