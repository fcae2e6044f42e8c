"""CPU oracle for the VAE decode / encode step next to the GP hot path.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy, the eval-mode forward of the reference's MLP VAEs and the logits -> tokens -> strings step:

* decoder / encoder stacks:  /root/reference/src/hdbo_benchmark/generative_models/vae_selfies.py:48-81
      nn.Sequential(Linear, BatchNorm1d, ReLU, Dropout(0.2), ..., Linear)   -- Dropout is the identity in eval mode
* decode():                  vae_selfies.py:103-115   logits.reshape(-1, max_sequence_length, len(alphabet))
* decode_to_string_array():  vae_selfies.py:235-247   np.argmax(logits, axis=-1) -> "".join(alphabet_i_to_s[i])
* encode() mean:             vae_selfies.py:93-101, encode_from_string_array :205-233 (z = q(z|x).mean)

The arithmetic itself is torch.nn (float32).  PARITY STATUS: torch IS installed here, so this restatement is pinned against
torch.nn.Sequential in eval mode on the CPU (tests/test_vae_oracle.py) and the committed fixture tests/golden/vae_decode_small.npz
was generated from torch.nn (tests/golden/make_vae_golden.py).  The reference's own tests hold no value for this step
(test_main_entry_point.py only checks that an iteration runs), and its trained weights / alphabets are not in the repository, so
the fixture uses random weights with non-trivial BatchNorm running statistics.

Accumulation is float64 on float32 parameters (the "true" value of the float32 network); torch's float32 result and the CUDA
kernels' float32 result both sit within a few float32 ulps of the logit scale of it, and the parity tests compare the arg-max only
where the top-2 gap exceeds that tolerance.

Only tests/, __graft_entry__.smoke() and bench tools' CPU legs may import this module.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

BN_EPS = 1e-5   # torch.nn.BatchNorm1d default


class Layer:
    """One Linear (+ optional eval-mode BatchNorm1d) (+ optional ReLU)."""

    def __init__(self, weight, bias=None, bn_weight=None, bn_bias=None, bn_mean=None, bn_var=None, eps: float = BN_EPS, relu: bool = False):
        self.weight = np.asarray(weight, dtype=np.float32)            # [N, K] (nn.Linear.weight)
        self.bias = None if bias is None else np.asarray(bias, dtype=np.float32)
        self.bn = None
        if bn_mean is not None:
            n = self.weight.shape[0]
            g = np.ones(n, np.float32) if bn_weight is None else np.asarray(bn_weight, np.float32)
            b = np.zeros(n, np.float32) if bn_bias is None else np.asarray(bn_bias, np.float32)
            self.bn = (g, b, np.asarray(bn_mean, np.float32), np.asarray(bn_var, np.float32), float(eps))
        self.relu = bool(relu)


def mlp_forward(layers: Sequence[Layer], x: np.ndarray) -> np.ndarray:
    """Eval-mode forward, float64 accumulation.  x [b, K0] -> [b, N_last] (float64)."""
    h = np.asarray(x, dtype=np.float32).astype(np.float64)
    for ly in layers:
        h = h @ ly.weight.astype(np.float64).T
        if ly.bias is not None:
            h = h + ly.bias.astype(np.float64)
        if ly.bn is not None:
            g, b, mean, var, eps = ly.bn
            h = (h - mean.astype(np.float64)) / np.sqrt(var.astype(np.float64) + eps) * g.astype(np.float64) + b.astype(np.float64)
        if ly.relu:
            h = np.maximum(h, 0.0)
    return h


def logits_to_tokens(logits: np.ndarray, seq_len: int, n_tokens: int) -> np.ndarray:
    """vae_selfies.py:241-242: argmax over the alphabet axis of logits.reshape(b, L, A) (first index on ties)."""
    return np.argmax(np.asarray(logits).reshape(-1, seq_len, n_tokens), axis=-1).astype(np.int32)


def tokens_to_strings(tokens: np.ndarray, alphabet_i_to_s) -> np.ndarray:
    """vae_selfies.py:243-247."""
    return np.array(["".join(alphabet_i_to_s[int(i)] for i in row) for row in tokens])


def top2_gap(logits: np.ndarray, seq_len: int, n_tokens: int) -> np.ndarray:
    """[b, L] difference between the two largest logits of every position (where an arg-max is well defined numerically)."""
    s = np.sort(np.asarray(logits).reshape(-1, seq_len, n_tokens), axis=-1)
    return s[..., -1] - s[..., -2]


def layers_from_torch_sequential(seq) -> List[Layer]:
    """Walk a torch nn.Sequential of Linear / BatchNorm1d / ReLU / Dropout modules (the reference's decoder / encoder pattern)."""
    import torch.nn as nn

    out: List[Layer] = []
    cur: Optional[dict] = None
    for m in seq:
        if isinstance(m, nn.Linear):
            if cur is not None:
                out.append(Layer(**cur))
            cur = dict(weight=m.weight.detach().numpy(), bias=None if m.bias is None else m.bias.detach().numpy())
        elif isinstance(m, nn.BatchNorm1d):
            cur.update(bn_weight=None if m.weight is None else m.weight.detach().numpy(),
                       bn_bias=None if m.bias is None else m.bias.detach().numpy(),
                       bn_mean=m.running_mean.numpy(), bn_var=m.running_var.numpy(), eps=m.eps)
        elif isinstance(m, nn.ReLU):
            cur["relu"] = True
        elif isinstance(m, nn.Dropout):
            pass
        else:
            raise TypeError(f"unsupported module {type(m).__name__}")
    if cur is not None:
        out.append(Layer(**cur))
    return out
