"""GPU parity of the VAE decode / encode step (hdbo_linear_f32, hdbo_argmax_tokens_f32 through hdbo_benchmark_b200.vae_decode)
against oracle/vae_oracle.py and the torch.nn-generated fixture."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import vae_oracle as V

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
TOL = 2e-5          # float32 layers vs their float64-accumulated value, relative to the largest logit


def _decoder(latent, widths, out, seed):
    torch.manual_seed(seed)
    mods, k = [], latent
    for w in widths:
        mods += [nn.Linear(k, w, dtype=torch.float32), nn.BatchNorm1d(w, dtype=torch.float32), nn.ReLU(), nn.Dropout(0.2)]
        k = w
    mods.append(nn.Linear(k, out, dtype=torch.float32))
    seq = nn.Sequential(*mods)
    for m in seq:
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.3, 1.7); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    return seq.eval()


@pytest.mark.parametrize("b", [1, 3, 8, 17, 100, 160])
def test_decode_at_the_reference_shape_matches_the_oracle(cuda_device, b):
    from hdbo_benchmark_b200.vae_decode import DeviceMLP, DeviceVAE

    L, A = 70, 64
    dec = _decoder(128, [256, 1024, 2048], L * A, seed=11)
    alphabet = {f"[t{i}]": i for i in range(A)}
    vae = DeviceVAE(DeviceMLP.from_sequential(dec, cuda_device), alphabet, L)
    z = torch.randn(b, 128, generator=torch.Generator().manual_seed(b), dtype=torch.float32)
    want = V.mlp_forward(V.layers_from_torch_sequential(dec), z.numpy())
    got = vae.decode_logits(z).reshape(b, -1).cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < TOL * scale
    tok = vae.decode_tokens(z).cpu().numpy()
    assert tok.dtype == np.int32 and tok.shape == (b, L)
    assert np.array_equal(tok, got.reshape(b, L, A).argmax(-1))          # bit-exact index work on the kernel's own logits
    clear = V.top2_gap(want, L, A) > 4 * TOL * scale
    assert clear.mean() > 0.99 and np.array_equal(tok[clear], V.logits_to_tokens(want, L, A)[clear])
    strings = vae.decode_to_string_array(z.numpy())
    inv = {v: k for k, v in alphabet.items()}
    assert strings.shape == (b,) and strings[-1] == "".join(inv[int(i)] for i in tok[-1])
    with torch.no_grad():                                                  # and the upstream arithmetic itself (torch.nn, CPU float32)
        ref = dec(z).numpy()
    assert np.abs(got - ref).max() < 2 * TOL * scale


def test_fixture_generated_by_torch_nn(cuda_device):
    from hdbo_benchmark_b200.vae_decode import DeviceMLP, DeviceVAE

    g = np.load(GOLD / "vae_decode_small.npz")
    spec = importlib.util.spec_from_file_location("make_vae_golden", GOLD / "make_vae_golden.py")
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    decoder, encoder, encoder_mu = mk.build()
    L, A = int(g["seq"]), int(g["alpha"])

    class Ref:   # the members DeviceVAE.from_module reads from a reference-style VAE
        pass
    ref = Ref(); ref.decoder, ref.encoder, ref.encoder_mu = decoder, encoder, encoder_mu
    ref.alphabet_s_to_i = {chr(ord("a") + i): i for i in range(A)}; ref.max_sequence_length = L
    vae = DeviceVAE.from_module(ref, cuda_device)
    logits = vae.decode_logits(g["z"]).reshape(len(g["z"]), -1).cpu().numpy()      # widths 8/16/40/52/66: the unaligned (scalar) path too
    assert np.abs(logits - g["logits"]).max() < 1e-5 * np.abs(g["logits"]).max()
    assert np.array_equal(vae.decode_tokens(g["z"]).cpu().numpy(), g["tokens"])
    strings = vae.decode_to_string_array(g["z"])
    assert strings[0] == "".join(chr(ord("a") + int(i)) for i in g["tokens"][0])
    onehot = vae.onehot_from_strings([[chr(ord("a") + int(i)) for i in row] for row in g["tokens"]])
    mu = vae.encode_mean(onehot).cpu().numpy()
    assert np.abs(mu - g["enc_mu"]).max() < 1e-5 * max(1.0, np.abs(g["enc_mu"]).max())


def test_argmax_ties_empty_batch_and_argument_errors(cuda_device):
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.vae_decode import DeviceMLP, DeviceVAE

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    logits = torch.zeros(2, 3 * 40, dtype=torch.float32, device=cuda_device)
    logits[0, 7] = logits[0, 33] = 5.0; logits[1, 40 + 39] = 1.0; logits[1, 80:120] = -3.0; logits[1, 80 + 5] = -2.0
    tok = torch.empty(2, 3, dtype=torch.int32, device=cuda_device)
    _lib.check(lib.hdbo_argmax_tokens_f32(logits.data_ptr(), logits.stride(0), 2, 3, 40, tok.data_ptr(), st), "argmax")
    assert tok.cpu().tolist() == [[7, 0, 0], [0, 39, 5]]                  # first index on ties (numpy.argmax)
    assert lib.hdbo_argmax_tokens_f32(0, 120, 2, 3, 40, tok.data_ptr(), st) == -1
    assert lib.hdbo_argmax_tokens_f32(logits.data_ptr(), 100, 2, 3, 40, tok.data_ptr(), st) == -2
    assert lib.hdbo_linear_f32(0, 8, 1, 8, logits.data_ptr(), 0, 0, 0, 0, logits.data_ptr(), 8, 4, st) == -1
    assert lib.hdbo_linear_f32(logits.data_ptr(), 4, 1, 8, logits.data_ptr(), 0, 0, 0, 0, logits.data_ptr(), 8, 4, st) == -4
    dec = _decoder(4, [8], 6, seed=0)
    vae = DeviceVAE(DeviceMLP.from_sequential(dec, cuda_device), {"x": 0, "y": 1, "z": 2}, 2)
    assert vae.decode_tokens(torch.zeros(0, 4)).shape == (0, 2)
    assert vae.decode_to_string_array(np.zeros((0, 4))).shape == (0,)
    with pytest.raises(ValueError):
        vae.decode_logits(torch.zeros(2, 5))
    with pytest.raises(ValueError):
        DeviceVAE(DeviceMLP.from_sequential(dec, cuda_device), {"x": 0, "y": 1}, 2)   # 6 logits != 2 * 2


def test_linear_without_bias_or_batchnorm_and_strided_io(cuda_device):
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    b, K, N = 5, 36, 21
    Xp = torch.randn(b, K + 4, generator=g, dtype=torch.float32).to(cuda_device)              # ldx > K
    W = torch.randn(N, K, generator=g, dtype=torch.float32).to(cuda_device)
    Yp = torch.full((b, N + 3), 7.0, dtype=torch.float32, device=cuda_device)                 # ldy > N: the padding must stay untouched
    _lib.check(lib.hdbo_linear_f32(Xp.data_ptr(), Xp.stride(0), b, K, W.data_ptr(), 0, 0, 0, 1, Yp.data_ptr(), Yp.stride(0), N,
                                   torch.cuda.current_stream().cuda_stream), "linear")
    want = np.maximum(Xp[:, :K].cpu().numpy().astype(np.float64) @ W.cpu().numpy().astype(np.float64).T, 0.0)
    assert np.abs(Yp[:, :N].cpu().numpy() - want).max() < 1e-5 * np.abs(want).max()
    assert bool((Yp[:, N:] == 7.0).all())


@pytest.mark.parametrize("b,K,N", [(150, 36, 21), (128, 16, 64), (257, 100, 130)])
def test_tiled_layer_kernel_on_ragged_shapes(cuda_device, b, K, N):
    """>= 128 rows take the shared-memory tiled kernel: partial tiles in every dimension, K not a multiple of the 16-deep k-step."""
    from hdbo_benchmark_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(b + K + N)
    X = torch.randn(b, K, generator=g, dtype=torch.float32).to(cuda_device)
    W = torch.randn(N, K, generator=g, dtype=torch.float32).to(cuda_device)
    bias = torch.randn(N, generator=g, dtype=torch.float32).to(cuda_device)
    sc = (torch.rand(N, generator=g, dtype=torch.float32) + 0.5).to(cuda_device)
    sh = torch.randn(N, generator=g, dtype=torch.float32).to(cuda_device)
    Y = torch.full((b, N + 2), -5.0, dtype=torch.float32, device=cuda_device)
    _lib.check(lib.hdbo_linear_f32(X.data_ptr(), X.stride(0), b, K, W.data_ptr(), bias.data_ptr(), sc.data_ptr(), sh.data_ptr(), 1, Y.data_ptr(),
                                   Y.stride(0), N, torch.cuda.current_stream().cuda_stream), "linear")
    want = np.maximum((X.cpu().numpy().astype(np.float64) @ W.cpu().numpy().astype(np.float64).T + bias.cpu().numpy()) * sc.cpu().numpy()
                      + sh.cpu().numpy(), 0.0)
    assert np.abs(Y[:, :N].cpu().numpy() - want).max() < 1e-5 * max(1.0, np.abs(want).max())
    assert bool((Y[:, N:] == -5.0).all())
