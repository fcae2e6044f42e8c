"""Kernel timeline of ONE pipelined fit (+ gradient) at n=4096, d=256 through torch.profiler (CUPTI): which stream ran what, when.
Writes a compact per-kernel table (name, stream, start us, duration us) as JSON lines + a per-stream summary.
    python tools/trace_fit_pipeline.py gpurun_out/r02_fit_trace.jsonl"""
import json, sys
sys.path.insert(0, ".")
import torch
from torch.profiler import ProfilerActivity, profile
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient
import bench

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_fit_trace.jsonl"
dev = torch.device("cuda")
n, d = 4096, 256
X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
gp.set_hyper(ls.to(dev), os_, nz, mu)
for _ in range(3):
    gp.fit_once(need_r2=True); mll_gradient(gp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gp.fit_once(need_r2=True)
    mll_gradient(gp)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/fit_trace.json")
tr = json.load(open("/tmp/fit_trace.json"))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
rows = [{"name": e["name"][:48], "stream": e["args"].get("stream"), "start_us": round(e["ts"] - t0, 2), "dur_us": round(e["dur"], 2),
         "grid": e["args"].get("grid")} for e in ks]
with open(out_path, "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
streams = {}
for r in rows:
    s = streams.setdefault(r["stream"], {"kernels": 0, "busy_us": 0.0, "first": r["start_us"], "last": 0.0})
    s["kernels"] += 1; s["busy_us"] += r["dur_us"]; s["last"] = max(s["last"], r["start_us"] + r["dur_us"])
leaf = [r for r in rows if "leaf" in r["name"]]
gaps = [round(leaf[i + 1]["start_us"] - leaf[i]["start_us"], 1) for i in range(len(leaf) - 1)]
print(json.dumps({"total_us": round(max(r["start_us"] + r["dur_us"] for r in rows), 1), "streams": streams,
                  "leaf_dur_us": [r["dur_us"] for r in leaf], "leaf_to_leaf_us": gaps}))
