"""Where does the pipelined factorisation spend its time?  Times hdbo_gp_fit at n=4096, d=256 with parts of the side-stream work
switched off (HDBO_PIPE_SKIP bit mask, read once per process: run this script once per mask).  Results with a non-zero mask are wrong
by construction -- timing only.   python tools/prof_fit_pipeline.py <mask>"""
import json, os, sys
sys.path.insert(0, ".")
mask = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["HDBO_PIPE_SKIP"] = str(mask)
import torch
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200.engine import DeviceGP
import bench
dev = torch.device("cuda")
n, d = 4096, 256
X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
gp.set_hyper(ls.to(dev), os_, nz, mu)
out = {"mask": mask, "pipelined": gp._aux is not None}
for need in (False, True):
    for _ in range(3):
        gp.fit_once(need_r2=need)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gp.fit_once(need_r2=need); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out["fit_ms_need_r2_%d" % need] = min(ts)
print(json.dumps(out))
