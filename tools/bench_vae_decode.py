"""VAE decode (latent -> token strings) at the reference's shape: 128 -> 256 -> 1024 -> 2048 -> 70 x 64 logits + arg-max.
Per batch size: device time of decode_tokens (CUDA-graph replay, resident latents), end-to-end decode_to_string_array from a
numpy array, the bytes of weights streamed per second, and the reference's own path (torch.nn float32 on the host cores +
np.argmax + string join, vae_selfies.py:235-247) timed beside it."""
import json, sys, time, os
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.nn as nn
from hdbo_benchmark_b200.vae_decode import DeviceMLP, DeviceVAE

L, A, LATENT = 70, 64, 128
torch.manual_seed(0)
dec = nn.Sequential(nn.Linear(LATENT, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.2), nn.Linear(256, 1024), nn.BatchNorm1d(1024), nn.ReLU(),
                    nn.Dropout(0.2), nn.Linear(1024, 2048), nn.BatchNorm1d(2048), nn.ReLU(), nn.Dropout(0.2), nn.Linear(2048, L * A)).eval()
alphabet = {f"[t{i}]": i for i in range(A)}
inv = {v: k for k, v in alphabet.items()}
vae = DeviceVAE(DeviceMLP.from_sequential(dec, "cuda"), alphabet, L)
peaks = {}
if os.path.exists("MEASURED_PEAKS.json"):
    peaks = json.load(open("MEASURED_PEAKS.json"))
hbm = float(peaks.get("hbm_gbs", 6539.2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for b in (1, 10, 100, 1000):
    z = torch.randn(b, LATENT)
    zd = z.cuda()
    for _ in range(3):
        vae.decode_tokens(zd)
    # cold: L2 flushed before every decode (one BO step decodes once; the weights come from HBM)
    cold = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vae.decode_tokens(zd); e1.record(); torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        vae.decode_tokens(zd)
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 50
    zn = z.numpy()
    vae.decode_to_string_array(zn)
    t0 = time.perf_counter()
    for _ in range(20):
        s = vae.decode_to_string_array(zn)
    e2e = (time.perf_counter() - t0) / 20 * 1e3
    with torch.no_grad():
        def ref():
            logits = dec(torch.tensor(zn, dtype=torch.float32)).reshape(-1, L, A).numpy()
            idx = np.argmax(logits, axis=-1)
            return np.array(["".join(inv[int(i)] for i in row) for row in idx])
        ref()
        t0 = time.perf_counter()
        reps = 20 if b <= 100 else 5
        for _ in range(reps):
            r = ref()
        cpu = (time.perf_counter() - t0) / reps * 1e3
    same = float(np.mean(r == s))
    wb = vae.decoder.weight_bytes()
    passes = 1 if b <= 16 else (b + 15) // 16
    print(json.dumps({"batch": b, "device_ms_cold_l2": round(float(np.median(cold)), 4), "device_ms_warm": round(warm, 4), "e2e_ms_from_numpy": round(e2e, 4),
                      "cpu_torch_nn_ms": round(cpu, 3), "cpu_threads": torch.get_num_threads(), "strings_equal_frac": same,
                      "weight_bytes": wb, "weight_stream_GBs_cold": round(wb / (float(np.median(cold)) * 1e-3) / 1e9, 1), "hbm_peak_GBs": hbm,
                      "frac_of_hbm_cold": round(wb / (float(np.median(cold)) * 1e-3) / 1e9 / hbm, 3), "weight_passes": passes,
                      "decodes_per_s_warm": round(b / (warm * 1e-3), 1)}), flush=True)
