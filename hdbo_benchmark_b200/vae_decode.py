"""Latent -> token strings on the GPU: the decode step on either side of the GP hot path.

Mirrors the decode / encode surface of the reference's MLP VAEs
(/root/reference/src/hdbo_benchmark/generative_models/vae_selfies.py):

    VAESelfies.decode(z).logits                 ->  DeviceVAE.decode_logits(z)            [b, L, A] float32
    VAESelfies.decode_to_string_array(z)        ->  DeviceVAE.decode_to_string_array(z)   np.ndarray of str   (:235-247)
    VAESelfies.encode(x).mean                   ->  DeviceVAE.encode_mean(x)              [b, latent]         (:93-101)

The reference runs torch's float32 nn.Sequential (Linear, BatchNorm1d, ReLU, Dropout) in eval mode, copies the [b, L*A] logits to
the host and takes np.argmax there.  Here every Linear(+BatchNorm+ReLU) is one `hdbo_linear_f32` launch (FP32, weights streamed
once per <= 16 batch rows) and the arg-max is `hdbo_argmax_tokens_f32`; only z (512 B per point) goes up and the tokens
(4 L bytes per point) come back.  For a fixed batch size the launches are replayed as one CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class DeviceMLP:
    """A stack of Linear (+ eval BatchNorm1d) (+ ReLU) layers resident on the device (float32)."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("hdbo_benchmark_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device or "cuda")
        self.lib = _lib.load()
        self.layers: List[dict] = []

    def add_linear(self, weight, bias=None, bn_weight=None, bn_bias=None, bn_mean=None, bn_var=None, eps: float = 1e-5, relu: bool = False):
        f32 = dict(dtype=torch.float32, device=self.device)
        W = torch.as_tensor(weight).detach().to(**f32).contiguous()
        n = W.shape[0]
        ly = dict(W=W, bias=None if bias is None else torch.as_tensor(bias).detach().to(**f32).contiguous(), scale=None, shift=None,
                  relu=int(bool(relu)), N=n, K=W.shape[1])
        if bn_mean is not None:
            # ATen's eval-mode batch norm: y = x * alpha + beta with alpha = gamma / sqrt(var + eps), beta = bias - mean * alpha (float32)
            mean = torch.as_tensor(bn_mean).detach().to(**f32)
            var = torch.as_tensor(bn_var).detach().to(**f32)
            g = torch.ones(n, **f32) if bn_weight is None else torch.as_tensor(bn_weight).detach().to(**f32)
            b = torch.zeros(n, **f32) if bn_bias is None else torch.as_tensor(bn_bias).detach().to(**f32)
            alpha = g / torch.sqrt(var + eps)
            ly["scale"] = alpha.contiguous()
            ly["shift"] = (b - mean * alpha).contiguous()
        if self.layers and self.layers[-1]["N"] != ly["K"]:
            raise ValueError(f"layer input width {ly['K']} does not match the previous output width {self.layers[-1]['N']}")
        self.layers.append(ly)
        return self

    @classmethod
    def from_sequential(cls, seq, device=None) -> "DeviceMLP":
        """From a torch nn.Sequential of Linear / BatchNorm1d / ReLU / Dropout modules (the reference's encoder / decoder pattern)."""
        import torch.nn as nn

        mlp = cls(device)
        cur = None
        for m in seq:
            if isinstance(m, nn.Linear):
                if cur is not None:
                    mlp.add_linear(**cur)
                cur = dict(weight=m.weight, bias=m.bias)
            elif isinstance(m, nn.BatchNorm1d):
                if cur is None:
                    raise TypeError("BatchNorm1d before the first Linear")
                cur.update(bn_weight=m.weight, bn_bias=m.bias, bn_mean=m.running_mean, bn_var=m.running_var, eps=m.eps)
            elif isinstance(m, nn.ReLU):
                cur["relu"] = True
            elif isinstance(m, nn.Dropout):
                pass   # identity in eval mode
            else:
                raise NotImplementedError(f"unsupported module in the stack: {type(m).__name__}")
        if cur is not None:
            mlp.add_linear(**cur)
        return mlp

    @property
    def in_features(self) -> int:
        return self.layers[0]["K"]

    @property
    def out_features(self) -> int:
        return self.layers[-1]["N"]

    def weight_bytes(self) -> int:
        return sum(4 * ly["N"] * ly["K"] for ly in self.layers)

    def _run(self, x: torch.Tensor, bufs: Sequence[torch.Tensor]) -> torch.Tensor:
        b = x.shape[0]
        h = x
        for ly, out in zip(self.layers, bufs):
            check(self.lib.hdbo_linear_f32(ptr(h), h.stride(0), b, ly["K"], ptr(ly["W"]), ptr(ly["bias"]), ptr(ly["scale"]), ptr(ly["shift"]),
                                           ly["relu"], ptr(out), out.stride(0), ly["N"], _stream()), "hdbo_linear_f32")
            h = out
        return h

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [b, in_features] (any float dtype / device) -> [b, out_features] float32 on the device."""
        x = torch.as_tensor(x).detach().to(device=self.device, dtype=torch.float32)
        if x.dim() != 2 or x.shape[1] != self.in_features:
            raise ValueError(f"expected [b, {self.in_features}], got {tuple(x.shape)}")
        x = x.contiguous()
        bufs = [torch.empty(x.shape[0], ly["N"], dtype=torch.float32, device=self.device) for ly in self.layers]
        if x.shape[0] == 0:
            return bufs[-1]
        return self._run(x, bufs)

    __call__ = forward


class DeviceVAE:
    """decode / encode of an MLP VAE over a token alphabet, on the device.  `alphabet_s_to_i` as in the reference's VAEs."""

    def __init__(self, decoder: DeviceMLP, alphabet_s_to_i: Dict[str, int], max_sequence_length: int, encoder: Optional[DeviceMLP] = None,
                 encoder_mu: Optional[DeviceMLP] = None):
        self.decoder, self.encoder, self.encoder_mu = decoder, encoder, encoder_mu
        self.alphabet_s_to_i = dict(alphabet_s_to_i)
        self.alphabet_i_to_s = {v: k for k, v in self.alphabet_s_to_i.items()}
        self.max_sequence_length = int(max_sequence_length)
        self.n_tokens = len(self.alphabet_s_to_i)
        if decoder.out_features != self.max_sequence_length * self.n_tokens:
            raise ValueError("decoder output width must be max_sequence_length * len(alphabet)")
        self.device = decoder.device
        self.lib = decoder.lib
        self._graph = {}

    @classmethod
    def from_module(cls, vae, device=None) -> "DeviceVAE":
        """From a reference-style module with .decoder (and optionally .encoder / .encoder_mu) nn.Sequential / nn.Linear members,
        .alphabet_s_to_i and .max_sequence_length (VAESelfies and friends); the module must be in eval() mode semantics."""
        import torch.nn as nn

        dec = DeviceMLP.from_sequential(vae.decoder, device)
        enc = DeviceMLP.from_sequential(vae.encoder, device) if hasattr(vae, "encoder") else None
        mu = DeviceMLP.from_sequential(nn.Sequential(vae.encoder_mu), device) if hasattr(vae, "encoder_mu") else None
        return cls(dec, vae.alphabet_s_to_i, vae.max_sequence_length, enc, mu)

    # ---- decode -------------------------------------------------------------------------------------------------------------
    def decode_logits(self, z) -> torch.Tensor:
        """[b, latent] -> logits [b, L, A] (float32, device): VAESelfies.decode(z).logits before the Categorical normalisation."""
        return self.decoder(z).reshape(-1, self.max_sequence_length, self.n_tokens)

    def decode_tokens(self, z) -> torch.Tensor:
        """[b, latent] -> arg-max tokens [b, L] (int32, device)."""
        z = torch.as_tensor(z).detach().to(device=self.device, dtype=torch.float32).contiguous()
        b = z.shape[0]
        tokens = torch.empty(b, self.max_sequence_length, dtype=torch.int32, device=self.device)
        if b == 0:
            return tokens
        g = self._graph.get(b)
        if g is None:
            g = self._capture(b)
        graph, zin, tok = g
        zin.copy_(z)
        graph.replay()
        tokens.copy_(tok)
        return tokens

    def _capture(self, b: int):
        dec = self.decoder
        zin = torch.zeros(b, dec.in_features, dtype=torch.float32, device=self.device)
        bufs = [torch.empty(b, ly["N"], dtype=torch.float32, device=self.device) for ly in dec.layers]
        tok = torch.empty(b, self.max_sequence_length, dtype=torch.int32, device=self.device)

        def run():
            logits = dec._run(zin, bufs)
            check(self.lib.hdbo_argmax_tokens_f32(ptr(logits), logits.stride(0), b, self.max_sequence_length, self.n_tokens, ptr(tok), _stream()),
                  "hdbo_argmax_tokens_f32")

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()                               # warm-up outside the capture
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        if len(self._graph) >= 8:               # a handful of batch sizes recur in a BO run (1, batch size, initial design)
            self._graph.pop(next(iter(self._graph)))
        self._graph[b] = (graph, zin, tok)
        return self._graph[b]

    def decode_to_string_array(self, z: np.ndarray) -> np.ndarray:
        """vae_selfies.py:235-247: latent points (numpy, [b, latent]) -> numpy array of b token strings."""
        tokens = self.decode_tokens(torch.as_tensor(np.asarray(z))).cpu().numpy()
        return np.array(["".join(self.alphabet_i_to_s[int(i)] for i in row) for row in tokens])

    # ---- encode -------------------------------------------------------------------------------------------------------------
    def encode_mean(self, onehot) -> torch.Tensor:
        """One-hot sequences [b, L, A] (or flattened) -> posterior mean [b, latent] (vae_selfies.py:93-101; what
        encode_from_string_array returns)."""
        if self.encoder is None or self.encoder_mu is None:
            raise RuntimeError("this DeviceVAE was built without an encoder")
        x = torch.as_tensor(onehot).detach().to(device=self.device, dtype=torch.float32)
        return self.encoder_mu(self.encoder(x.reshape(x.shape[0], -1)))

    def onehot_from_strings(self, token_lists: Sequence[Sequence[str]]) -> torch.Tensor:
        """Token sequences (already split and padded to max_sequence_length) -> one-hot [b, L, A] float32 on the device."""
        idx = torch.tensor([[self.alphabet_s_to_i[t] for t in seq] for seq in token_lists], dtype=torch.int64, device=self.device)
        if idx.shape[1] != self.max_sequence_length:
            raise ValueError("sequences must be padded to max_sequence_length")
        return torch.nn.functional.one_hot(idx, self.n_tokens).to(torch.float32)
