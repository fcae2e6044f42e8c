"""Probe the GPU box once: host cores, CPU FP64 GEMM rate, cuBLAS FP64/TF32 GEMM rate (denominators only)."""
import json, os, subprocess, time
import torch

out = {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}
try:
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception:
    pass
try:
    out["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines()
                    if any(k in l for k in ("Model name", "Socket", "Core(s)", "Thread(s)", "CPU(s):"))]
except Exception as e:
    out["lscpu"] = str(e)
a = torch.randn(4096, 4096, dtype=torch.float64); b = torch.randn(4096, 4096, dtype=torch.float64)
(a @ b)
t = time.perf_counter(); (a @ b); dt = time.perf_counter() - t
out["cpu_dgemm_gflops"] = 2 * 4096**3 / dt / 1e9
if torch.cuda.is_available():
    out["gpu"] = torch.cuda.get_device_name(0)
    for dt_, name in ((torch.float64, "cublas_fp64"), (torch.float32, "cublas_fp32")):
        n = 8192
        A = torch.randn(n, n, dtype=dt_, device="cuda"); B = torch.randn(n, n, dtype=dt_, device="cuda")
        for _ in range(2): A @ B
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); A @ B; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_tflops"] = 2 * n**3 / best / 1e9
    L = torch.randn(4096, 4096, dtype=torch.float64, device="cuda"); K = L @ L.T + 4096 * torch.eye(4096, dtype=torch.float64, device="cuda")
    torch.linalg.cholesky(K); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.linalg.cholesky(K); e1.record(); torch.cuda.synchronize()
    out["cusolver_potrf_4096_ms"] = e0.elapsed_time(e1)
print(json.dumps(out, indent=1))
