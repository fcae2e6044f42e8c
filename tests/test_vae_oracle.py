"""CPU tests of the VAE decode oracle (oracle/vae_oracle.py): pinned against torch.nn (the upstream arithmetic of
/root/reference/src/hdbo_benchmark/generative_models/vae_selfies.py:48-81, 235-247) and against the committed fixture."""
import importlib.util
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from oracle import vae_oracle as V

GOLD = Path(__file__).resolve().parent / "golden"


def _load_maker():
    spec = importlib.util.spec_from_file_location("make_vae_golden", GOLD / "make_vae_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_matches_torch_nn_eval_mode_at_the_reference_shape():
    """latent 128 -> 256 -> 1024 -> 2048 -> 70 x 64 logits, random weights / running statistics."""
    torch.manual_seed(3)
    L, A = 70, 64
    dec = nn.Sequential(nn.Linear(128, 256, dtype=torch.float32), nn.BatchNorm1d(256, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2), nn.Linear(256, 1024, dtype=torch.float32), nn.BatchNorm1d(1024, dtype=torch.float32), nn.ReLU(),
                        nn.Dropout(0.2), nn.Linear(1024, 2048, dtype=torch.float32), nn.BatchNorm1d(2048, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2), nn.Linear(2048, L * A, dtype=torch.float32))
    for m in dec:
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.3, 1.7); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    dec.eval()
    z = torch.randn(9, 128, dtype=torch.float32)
    with torch.no_grad():
        want = dec(z).numpy()
    got = V.mlp_forward(V.layers_from_torch_sequential(dec), z.numpy())
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 2e-5 * scale                      # float32 network vs its float64-accumulated value
    tok_want = want.reshape(-1, L, A).argmax(-1)
    tok_got = V.logits_to_tokens(got, L, A)
    clear = V.top2_gap(got, L, A) > 1e-4 * scale
    assert clear.mean() > 0.99 and np.array_equal(tok_got[clear], tok_want[clear])
    alphabet = {i: f"[t{i}]" for i in range(A)}
    s = V.tokens_to_strings(tok_got, alphabet)
    assert s.shape == (9,) and s[0] == "".join(alphabet[int(i)] for i in tok_got[0])


def test_oracle_matches_the_committed_fixture_and_the_fixture_is_reproducible():
    g = np.load(GOLD / "vae_decode_small.npz")
    mk = _load_maker()
    decoder, encoder, encoder_mu = mk.build()
    L, A = int(g["seq"]), int(g["alpha"])
    for name, seq in (("dec", decoder), ("enc", encoder)):
        for k, v in seq.state_dict().items():
            assert np.array_equal(g[f"{name}.{k}"], v.numpy()), k    # same seeds -> same parameters
    logits = V.mlp_forward(V.layers_from_torch_sequential(decoder), g["z"])
    assert np.abs(logits - g["logits"]).max() < 1e-5 * np.abs(g["logits"]).max()
    assert np.array_equal(V.logits_to_tokens(logits, L, A), g["tokens"])
    onehot = np.eye(A, dtype=np.float32)[g["tokens"]].reshape(len(g["z"]), -1)
    mu = V.mlp_forward(V.layers_from_torch_sequential(nn.Sequential(*encoder, encoder_mu)), onehot)
    assert np.abs(mu - g["enc_mu"]).max() < 1e-5 * max(1.0, np.abs(g["enc_mu"]).max())


def test_argmax_takes_the_first_index_on_ties():
    logits = np.zeros((1, 2 * 4), np.float32)
    logits[0, 1] = logits[0, 3] = 2.0          # position 0: tie between tokens 1 and 3
    assert V.logits_to_tokens(logits, 2, 4).tolist() == [[1, 0]]
