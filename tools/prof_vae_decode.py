"""A few decode_tokens calls at the reference's shape for ncu launch lists (argv: batch sizes)."""
import sys
sys.path.insert(0, ".")
import torch
import torch.nn as nn
from hdbo_benchmark_b200.vae_decode import DeviceMLP, DeviceVAE

L, A = 70, 64
torch.manual_seed(0)
dec = nn.Sequential(nn.Linear(128, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.2), nn.Linear(256, 1024), nn.BatchNorm1d(1024), nn.ReLU(),
                    nn.Dropout(0.2), nn.Linear(1024, 2048), nn.BatchNorm1d(2048), nn.ReLU(), nn.Dropout(0.2), nn.Linear(2048, L * A)).eval()
vae = DeviceVAE(DeviceMLP.from_sequential(dec, "cuda"), {f"[t{i}]": i for i in range(A)}, L)
for b in [int(a) for a in sys.argv[1:]] or [1, 10]:
    z = torch.randn(b, 128).cuda()
    for _ in range(3):
        vae.decode_tokens(z)
    torch.cuda.synchronize()
print("ok")
