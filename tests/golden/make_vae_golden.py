"""Generates tests/golden/vae_decode_small.npz from torch.nn (CPU, float32, eval mode): the upstream arithmetic of the reference's
MLP VAE decoder / encoder (/root/reference/src/hdbo_benchmark/generative_models/vae_selfies.py:48-81) at a small size, with random
weights and non-trivial BatchNorm running statistics (the reference's trained weights are not in its repository).

    python tests/golden/make_vae_golden.py
"""
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

LATENT, SEQ, ALPHA = 8, 6, 11


def build(seed: int = 0):
    torch.manual_seed(seed)
    width = SEQ * ALPHA
    decoder = nn.Sequential(nn.Linear(LATENT, 16, dtype=torch.float32), nn.BatchNorm1d(16, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2),
                            nn.Linear(16, 40, dtype=torch.float32), nn.BatchNorm1d(40, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2),
                            nn.Linear(40, 52, dtype=torch.float32), nn.BatchNorm1d(52, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2),
                            nn.Linear(52, width, dtype=torch.float32))
    encoder = nn.Sequential(nn.Linear(width, 52, dtype=torch.float32), nn.BatchNorm1d(52, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2),
                            nn.Linear(52, 16, dtype=torch.float32), nn.BatchNorm1d(16, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2))
    encoder_mu = nn.Linear(16, LATENT, dtype=torch.float32)
    g = torch.Generator().manual_seed(seed + 1)
    for seq in (decoder, encoder):
        for m in seq:
            if isinstance(m, nn.BatchNorm1d):   # as after training: running statistics and affine parameters away from (0, 1, 1, 0)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g, dtype=torch.float32) * 0.3)
                m.running_var.copy_(torch.rand(m.num_features, generator=g, dtype=torch.float32) * 1.5 + 0.25)
                m.weight.data.copy_(torch.rand(m.num_features, generator=g, dtype=torch.float32) + 0.5)
                m.bias.data.copy_(torch.randn(m.num_features, generator=g, dtype=torch.float32) * 0.2)
    for m in (decoder, encoder, encoder_mu):
        m.eval()
    return decoder, encoder, encoder_mu


def main():
    decoder, encoder, encoder_mu = build()
    g = torch.Generator().manual_seed(7)
    z = torch.randn(23, LATENT, generator=g, dtype=torch.float32)
    with torch.no_grad():
        logits = decoder(z)
        tokens = logits.reshape(-1, SEQ, ALPHA).argmax(-1)
        onehot = torch.nn.functional.one_hot(tokens, ALPHA).to(torch.float32)
        mu = encoder_mu(encoder(onehot.flatten(start_dim=1)))
    out = {"z": z.numpy(), "logits": logits.numpy(), "tokens": tokens.numpy().astype(np.int32), "enc_mu": mu.numpy(),
           "latent": LATENT, "seq": SEQ, "alpha": ALPHA}
    for name, seq in (("dec", decoder), ("enc", encoder), ("mu", nn.Sequential(encoder_mu))):
        for k, v in seq.state_dict().items():
            out[f"{name}.{k}"] = v.numpy()
    path = Path(__file__).resolve().parent / "vae_decode_small.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
