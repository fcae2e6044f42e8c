"""Pipelined factorisation vs the recursion: MLL fit + gradient device time (CUDA events, best of 8) over n, both schedules
(HDBO_FIT_PIPELINE is read per DeviceGP).   python tools/time_fit_crossover.py"""
import json, os, sys
os.environ["HDBO_PIPE_MIN_TILES"] = "4"
sys.path.insert(0, ".")
import torch
import bench
from hdbo_benchmark_b200 import _lib
from hdbo_benchmark_b200 import engine
from hdbo_benchmark_b200.engine import DeviceGP
from hdbo_benchmark_b200.models import mll_gradient

dev = torch.device("cuda")
for n, d in ((640, 128), (768, 128), (896, 128), (1024, 128), (1536, 128), (2048, 128), (3072, 256), (4096, 256)):
    X, y, ls, os_, nz, mu = bench.synthetic_problem(n, d)
    row = {"n": n, "d": d}
    for piped in (False, True):
        os.environ["HDBO_FIT_PIPELINE"] = "1" if piped else "0"
        old = engine.PIPELINE_MIN_TILES
        engine.PIPELINE_MIN_TILES = 1 if piped else old
        gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
        engine.PIPELINE_MIN_TILES = old
        gp.set_hyper(ls.to(dev), os_, nz, mu)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gp.fit_once(need_r2=True); mll_gradient(gp); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row["pipelined_ms" if piped else "recursion_ms"] = round(best, 4)
        row["piped_active" if piped else "_"] = gp._aux is not None
        del gp
    row.pop("_", None)
    print(json.dumps(row), flush=True)
