#!/usr/bin/env python
"""Headline benchmark: acquisition evaluations per second at n=4096, d=256 (BASELINE.json metric, configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference arm: CPU FP64 restatement on the host cores

One step = posterior mean + variance + LogEI for every candidate of ONE 2^20-point scrambled-Sobol pool against a fitted exact GP
(n=4096, d=256, ARD Matern-5/2; L^-1 and alpha resident) + the arg-max, reached through the BoTorch surface the reference's solvers
call: `LogExpectedImprovement(SingleTaskGP)(X)` under no_grad.  `value` times the pool resident in HBM; `e2e` times the same call
with X a pinned HOST tensor (H2D of the candidates and D2H of the acquisition values inside the timed region).

Multi-GPU (BASELINE.json configs[2]: "pool sharded across 1/2/4/8 B200"): STRONG scaling by default -- the 2^20 pool is split into
contiguous Sobol slices, one per rank; the only collective is a 16-byte/rank all-gather of (max, argmax) per step.
`--scaling weak` keeps 2^20 candidates per GPU (round-1 behaviour).  The other two shard axes of SURVEY 8e are timed as extra legs
in the same line: SAAS hyper-parameter samples (configs[3], one all-reduce of the mixture sums) and qEI restarts r::P (configs[4]).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

METRIC = "acq evals/sec at n=4096,d=256"
UNIT = "evals/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=4096)
    p.add_argument("--d", type=int, default=256)
    p.add_argument("--pool", type=int, default=1 << 20, help="candidates per step: in total (strong scaling) or per GPU (weak)")
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--cpu-sample", type=int, default=0, help="candidates per CPU-baseline batch (0 = sweep 1k..16k, keep the best)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] / optimize_acqf / gradient legs")
    p.add_argument("--no-mll-batched", action="store_true", help="skip the batched (4 hyper-parameter vectors) MLL timing")
    p.add_argument("--var-mode", default="i8", choices=["i8", "f64"],
                   help="pool contraction: exact int8 digit slices on tcgen05 (the package default) or FP64 DMMA")
    return p.parse_args()


def synthetic_problem(n, d, noise=1e-3, seed=0):
    """The synthetic workload of SURVEY.md 8d (same draws as oracle.gp_oracle.synthetic_problem(ls_mode="prior"); kept here so that
    the measured arm does not touch oracle/): X ~ U[0,1]^{n x d}, y = standardise(sin(3 sum_{k<4} x_k) + 0.1 eps),
    l_k = sqrt(d) exp(0.25 xi_k), outputscale 1, noise 1e-3, mean 0."""
    g0, g1, g2 = (torch.Generator().manual_seed(seed + i) for i in range(3))
    X = torch.rand(n, d, generator=g0, dtype=torch.float64)
    y = torch.sin(3.0 * X[:, : min(4, d)].sum(-1)) + 0.1 * torch.randn(n, generator=g1, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    ls = math.sqrt(d) * torch.exp(0.25 * torch.randn(d, generator=g2, dtype=torch.float64))
    return X, y, ls, 1.0, noise, 0.0


def workload_config(args, world):
    total = args.pool if args.scaling == "strong" else args.pool * world
    return {
        "workload": f"synthetic GP n={args.n}, d={args.d}, ARD Matern-5/2, posterior+LogEI+argmax over a 2^{int(math.log2(args.pool))}"
                    f"-candidate scrambled-Sobol pool (BASELINE.json configs[2])",
        "n_train": args.n, "d": args.d, "pool_total": total, "scaling_mode": args.scaling,
        "acquisition": "LogExpectedImprovement", "kernel": "matern52_ard", "noise": 1e-3,
        "var_mode": getattr(args, "var_mode", "i8"),
        "cache": "inputs larger than L2 (per launch: K* chunk 620 MB FP64 / 543 MB int8 digit images, L^-1 67 MB / 117 MB; 126 MB L2)",
    }


def cpu_model_name():
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms under the load of the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------------
# the CPU reference: oracle/gp_oracle.py posterior + LogEI + argmax (torch FP64, MKL), all host threads
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_eval_once(O, st, Z, best_f):
    m, v = O.posterior(st, Z)
    a = O.log_expected_improvement(m, v, best_f)
    return int(a.argmax())


def cpu_pick_batch(O, st, Zpool, best_f, requested):
    """SURVEY 8d protocol: sweep the batch size 1k..16k and keep the one with the best throughput (one timed pass each)."""
    if requested:
        return requested, {}
    sweep = {}
    with torch.no_grad():
        for b in (1024, 2048, 4096, 8192, 16384):
            if b > Zpool.shape[0]:
                break
            cpu_eval_once(O, st, Zpool[:b], best_f)
            t0 = time.perf_counter()
            cpu_eval_once(O, st, Zpool[:b], best_f)
            sweep[b] = b / (time.perf_counter() - t0)
    return max(sweep, key=sweep.get), {str(k): round(v, 1) for k, v in sweep.items()}


def cpu_reference_rate(O, st, Zpool, best_f, requested_batch, budget_s=12.0):
    torch.set_num_threads(os.cpu_count() or 1)
    batch, sweep = cpu_pick_batch(O, st, Zpool, best_f, requested_batch)
    Zs = Zpool[:batch]
    with torch.no_grad():
        cpu_eval_once(O, st, Zs, best_f)
        t0 = time.perf_counter()
        reps = 0
        while True:
            cpu_eval_once(O, st, Zs, best_f)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or reps >= 200:
                break
    return reps * batch / dt, reps, dt, batch, sweep


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gp_oracle as O

    X, y, ls, os_, nz, mu = O.synthetic_problem(args.n, args.d)
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Zpool = O.sobol_candidates(16384, args.d)
    best_f = float(y.max())
    batch, sweep = cpu_pick_batch(O, st, Zpool, best_f, args.cpu_sample)
    Z = Zpool[:batch]
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            cpu_eval_once(O, st, Z, best_f)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_eval_once(O, st, Z, best_f)
        dt = time.perf_counter() - t0
    value = args.steps * batch / dt
    cores = torch.get_num_threads()
    sample = (f"each step evaluates {batch} Sobol candidates (a bounded sample of the 2^{int(math.log2(args.pool))} pool; batch size "
              f"= best of the 1k..16k sweep {sweep}), oracle/gp_oracle.py posterior+LogEI+argmax, torch FP64 / MKL, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
        "evaluated_candidates_per_step": batch, "cpu_model": cpu_model_name(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu_model": cpu_model_name(),
                         "batch_sweep_evals_per_s": sweep},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "botorch/gpytorch are not installable offline (not in /opt/wheelhouse); the reference arm is the CPU FP64 "
                "restatement of their exact path (oracle port), all host threads",
    }))


def _measured_peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch.distributed as dist

    import bench_extra
    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200 import acquisition as A
    from hdbo_benchmark_b200 import gp_modules as G
    from hdbo_benchmark_b200.distributed import global_argmax, shard_bounds
    from hdbo_benchmark_b200.engine import DEFAULT_CHUNK_ROWS, DeviceGP
    from hdbo_benchmark_b200.models import SingleTaskGP, mll_gradient
    from hdbo_benchmark_b200.sampling import DeviceSobol

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # a default run takes well under a minute of GPU time: anything still alive after 7 minutes is stuck (a collective waiting for a
    # rank that left) -- leave loudly instead of holding N GPUs until somebody's limit kills the job
    watchdog = threading.Timer(420.0, lambda: (sys.stderr.write("bench.py: watchdog expired after 420 s, aborting\n"), sys.stderr.flush(), os._exit(3)))
    watchdog.daemon = True
    watchdog.start()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a rank that dies must not leave the others spinning in a collective until the box is reclaimed
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _lib.load()
    devlib = _lib.load_dev()      # tensor-pipe ceiling probes (roofline denominators): developer library, not the product one
    peaks = _measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6539.2))

    n, d = args.n, args.d
    X, y, ls, os_, nz, mu = synthetic_problem(n, d)
    best_f = float(y.max())

    # ---- the model, through the BoTorch surface (what a poli-baselines solver constructs; hyper-parameters = the synthetic fit) ----
    model = SingleTaskGP(X, y.unsqueeze(-1), covar_module=G.ScaleKernel(G.MaternKernel(nu=2.5, ard_num_dims=d)),
                         likelihood=G.GaussianLikelihood(), outcome_transform=None)
    model.covar_module.base_kernel.lengthscale = ls
    model.covar_module.outputscale = os_
    model.likelihood.noise = nz
    model.mean_module.constant = mu
    model.eval()
    acqf = A.LogExpectedImprovement(model, best_f=best_f)
    gp = model._device_gp()            # the fitted device state behind the surface (package default mode: int8 digit slices)
    assert gp.var_mode == "i8" and gp._i8_ok, "the default surface must take the tcgen05 path on this workload"
    if args.var_mode == "f64":
        gp.var_mode = "f64"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    def event_ms(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        out = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    # ---- secondary metric: exact MLL value + full hyper-gradient (kernel build -> potrf -> inverse -> gradient) ----
    gpf = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode="f64")
    gpf.set_hyper(ls.to(dev), os_, nz, mu)

    def mll_once():
        gpf.fit_once(need_r2=True)
        return mll_gradient(gpf)
    mll_eager_ms = event_ms(mll_once, 5)
    # the product's fit loop (fit.py: the fast closure of fit_gpytorch_mll) replays exactly this pair of ABI calls as a CUDA graph:
    # that is the headline figure; the eager enqueue (~170 stream operations of the pipelined factorisation) is reported beside it
    torch.cuda.synchronize()
    mll_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(mll_graph):
        mll_once()
    mll_ms = event_ms(mll_graph.replay, 7)
    del mll_graph
    gpf.attach_timeline(64)
    mll_once()
    fit_tl = gpf.collect_timeline()
    gpf.detach_timeline()
    mll_batched = None
    if rank == 0 and not args.no_mll_batched:
        Bm = 4
        gpb = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, batch=Bm, keep_r2=True, var_mode="f64")
        scale = torch.linspace(0.9, 1.1, Bm, dtype=torch.float64).unsqueeze(1)
        gpb.set_hyper((ls.unsqueeze(0) * scale).to(dev), os_, nz, mu)

        def mll_b():
            gpb.fit_once(need_r2=True)
            return mll_gradient(gpb)
        tb = event_ms(mll_b, 3)
        mll_batched = {"batch": Bm, "ms_per_batch": min(tb), "ms_per_member": min(tb) / Bm}
        del gpb
    del gpf
    torch.cuda.empty_cache()

    # ---- this rank's slice of the pool: scrambled Sobol (seed 3), generated in HBM by hdbo_sobol_draw (bit-identical to torch) ----
    if args.scaling == "strong":
        s0, s1 = shard_bounds(args.pool, rank, world)
        pool_total = args.pool
    else:
        s0, s1 = rank * args.pool, (rank + 1) * args.pool
        pool_total = args.pool * world
    m = s1 - s0
    Z_dev = DeviceSobol(d, scramble=True, seed=3, device=dev).fast_forward(s0).draw(m)
    Z_host = torch.empty(m, d, dtype=torch.float64).pin_memory()
    Z_host.copy_(Z_dev)
    Xq_dev, Xq_host = Z_dev.unsqueeze(1), Z_host.unsqueeze(1)      # b x q x d with q = 1, the shape acqf(X) takes

    def step_resident():
        with torch.no_grad():
            a = acqf(Xq_dev)                                       # BoTorch surface, device tensor in / out
        v, i = gp.argmax(a)
        return global_argmax(v, i, s0) if world > 1 else (v, i)

    def step_e2e():
        with torch.no_grad():
            a = acqf(Xq_host)                                      # BoTorch surface, pinned host tensor in, host tensor out
        v, i = torch.max(a, dim=0)
        if world > 1:
            return global_argmax(v.to(dev), i.to(dev), s0)
        return v, i

    def step_e2e_engine():
        _, v, i = gp.posterior_acq_host(Z_host, _lib.ACQ_LOGEI, best_f)
        return global_argmax(v, i, s0) if world > 1 else (v, i)

    # ---- tensor-pipe ceilings, measured live (denominators of the roofline) ----
    stream = lambda: torch.cuda.current_stream().cuda_stream
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fl = C.c_double(0.0)
    peak = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(devlib.hdbo_fp64_peak_probe(scratch.data_ptr(), 16384, sms, C.byref(fl), stream()), "peak")
        e1.record(); torch.cuda.synchronize()
        peak = max(peak, fl.value / e0.elapsed_time(e1) / 1e9)
    flag = torch.zeros(4, dtype=torch.int32, device=dev)
    ops = C.c_double(0.0)
    i8_peak = 0.0          # burst: best of 6 short launches
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(devlib.hdbo_i8_peak_probe(flag.data_ptr(), 4096, sms, C.byref(ops), stream()), "i8 peak")
        e1.record(); torch.cuda.synchronize()
        i8_peak = max(i8_peak, ops.value / e0.elapsed_time(e1) / 1e9)
    # sustained: ~1.5 s of back-to-back launches, the second half timed (power-capped clocks, like the kernel inside a long step)
    reps = max(8, int(1.5 / (ops.value / (i8_peak * 1e12) * 8)))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for r in range(reps):
        if r == reps // 2:
            evs[0].record()
        _lib.check(devlib.hdbo_i8_peak_probe(flag.data_ptr(), 32768, sms, C.byref(ops), stream()), "i8 peak")
    evs[1].record(); torch.cuda.synchronize()
    i8_peak_sustained = ops.value * (reps - reps // 2) / evs[0].elapsed_time(evs[1]) / 1e9
    assert int(flag[0].item()) == 0, "int8 peak probe timed out"

    # ---- the FP64 DMMA mode of the same step (>= 5 timed steps) + the in-run guard of the int8 mode against it ----
    chunks = math.ceil(m / DEFAULT_CHUNK_ROWS)
    f64_steps = max(5, min(args.steps, 5))
    gp.var_mode = "f64"
    Zc = Z_dev[: 1 << 14]
    _, var_f64, acq_f64 = gp.posterior_acq(Zc, _lib.ACQ_LOGEI, best_f, want_mean=False)
    for _ in range(2):
        step_resident()
    gp.attach_timeline(f64_steps * chunks * 8 + 64)
    ms_f64, (bv_f64, bi_f64) = timed(step_resident, f64_steps)
    tl_f64 = gp.collect_timeline()
    gp.detach_timeline()
    f64_mode = {"value": f64_steps * pool_total / (ms_f64 / 1e3), "unit": UNIT, "steps": f64_steps, "ms_per_step": ms_f64 / f64_steps}
    if "var" in tl_f64:
        f64_mode["k_var_tflops"] = f64_steps * m * float(n) * n / (tl_f64["var"][0] / 1e3) / 1e12
        f64_mode["k_var_frac_of_dmma_peak"] = f64_mode["k_var_tflops"] / peak
    mode_check = None
    gp.var_mode = args.var_mode
    if args.var_mode == "i8":
        _, var_i8, acq_i8 = gp.posterior_acq(Zc, _lib.ACQ_LOGEI, best_f, want_mean=False)
        mode_check = {"candidates": int(Zc.shape[0]),
                      "max_abs_var_diff_over_outputscale": float((var_i8 - var_f64).abs().max() / float(os_)),
                      "max_rel_var_diff": float(((var_i8 - var_f64).abs() / var_f64.abs()).max()),
                      "same_argmax": bool(int(acq_i8[0].argmax()) == int(acq_f64[0].argmax())), "timeout_flag": gp.i8_flag()}
        assert mode_check["max_rel_var_diff"] < 1e-6 and mode_check["same_argmax"] and mode_check["timeout_flag"] == 0, mode_check

    # ---- headline: W warm-up steps, K timed steps, CUDA events + barrier on both sides, max over ranks ----
    # clocks / throttle reasons are sampled from the warm-up on (the same load as the timed steps): at N = 8 the timed region of a
    # strong-scaled run is 0.1 s, shorter than nvidia-smi's start-up + one 100 ms sampling period
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        step_resident()
    gp.attach_timeline(args.steps * chunks * 8 + 64)
    ms_total, (best_v, best_i) = timed(step_resident, args.steps)
    tl = gp.collect_timeline()
    gp.detach_timeline()
    # untimed: keep the same load up until the sampler has data.  Every rank decides for ITSELF how long that takes, so these steps
    # must not contain a collective (the first version called step_resident() here: ranks left the loop at different times and the
    # arg-max all-gather of the slower ones hung an 8-GPU run)
    t_more = time.perf_counter()
    while len(clocks.lines) < 4 and time.perf_counter() - t_more < 3.0:
        with torch.no_grad():
            gp.argmax(acqf(Xq_dev))
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    clk = clocks.stop()
    clk["window"] = "warm-up + timed steps (+ untimed identical steps until >= 4 samples)"
    value = args.steps * pool_total / (ms_total / 1e3)
    if args.var_mode == "i8":
        assert int(best_i) == int(bi_f64), "int8 and FP64 modes disagree on the arg-max of the full pool"

    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ms_e2e, (best_v2, best_i2) = timed(step_e2e, args.steps)
        assert int(best_i2) == int(best_i), "host path and resident path disagree on the arg-max"
        for _ in range(2):
            step_e2e_engine()
        ms_eng, (_, best_i3) = timed(step_e2e_engine, args.steps)
        assert int(best_i3) == int(best_i)
        e2e = {"value": args.steps * pool_total / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": pool_total * d * 8, "d2h_bytes_per_step": pool_total * 8,
               "api": "LogExpectedImprovement(SingleTaskGP)(X) under torch.no_grad(), X a pinned HOST tensor [pool, 1, d] "
                      "(BoTorch surface; candidates streamed H2D slab by slab, values returned as a host tensor) + torch.max",
               "engine_level": {"value": args.steps * pool_total / (ms_eng / 1e3), "ms_per_step": ms_eng / args.steps,
                                "api": "DeviceGP.posterior_acq_host"},
               "surface_over_engine": ms_eng / ms_e2e,
               "host_read_gbs": pool_total * d * 8 / (ms_e2e / args.steps / 1e3) / 1e9}

    # ---- the other scaling mode, short (same process, same GP): both curves in every line when N > 1 ----
    other = None
    if world > 1:
        if args.scaling == "strong":
            o0, o1, o_total = rank * args.pool, (rank + 1) * args.pool, args.pool * world
        else:
            (o0, o1), o_total = shard_bounds(args.pool, rank, world), args.pool
        Zo = DeviceSobol(d, scramble=True, seed=3, device=dev).fast_forward(o0).draw(o1 - o0).unsqueeze(1)

        def step_other():
            with torch.no_grad():
                a = acqf(Zo)
            v, i = gp.argmax(a)
            return global_argmax(v, i, o0)
        for _ in range(2):
            step_other()
        ms_o, _ = timed(step_other, 3)
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "value": 3 * o_total / (ms_o / 1e3), "unit": UNIT,
                 "ms_per_step": ms_o / 3, "pool_total": o_total, "steps": 3}
        del Zo

    # ---- extra legs (other shard axes, gradient path, optimize_acqf) -- outside every timed region above ----
    extra = {}
    if not args.no_extra:
        del Z_dev, Xq_dev
        torch.cuda.empty_cache()
        extra = bench_extra.run_all(args, rank, world, dev, model, acqf, gp, timed, event_ms, peak, best_f)

    if rank == 0:
        launches = (4 * chunks + 1) * world    # prep, cross, var, finalize per chunk + argmax, every rank
        pairs = 21.0
        var_ms, var_n = tl.get("var", (0.0, 0))
        cross_ms, _ = tl.get("cross", (0.0, 0))
        alg_var = float(args.steps) * m * float(n) * n          # n^2 flops per candidate (triangular V = K* L^-T), this rank
        alg_cross = float(args.steps) * m * 2.0 * n * d
        traffic = None
        ncu_path = ROOT / "profiles" / "ncu_summary.json"
        key = "k_var_i8" if args.var_mode == "i8" else "k_var"
        if ncu_path.exists():
            try:
                traffic = json.loads(ncu_path.read_text()).get(key, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        whole = args.steps * m * (2.0 * n * d + float(n) * n + 2 * n) / (ms_total / 1e3) / 1e12
        # HBM-bound phases: algorithmic bytes / measured time against the measured copy bandwidth (MEASURED_PEAKS.json)
        d32 = 32 * math.ceil(d / 32)
        nt2 = 2 * (gp.n_pad // 128)
        hbm = {}

        def hbm_entry(name, kernel, bytes_total, ms, what):
            if ms > 0:
                gbs = bytes_total / (ms / 1e3) / 1e9
                hbm[name] = {"kernel": kernel, "achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "bytes": what}
        if args.var_mode == "i8":
            hbm_entry("prep", "k_slice_rows<128>", args.steps * m * (d * 8.0 + 6.0 * d32 + 16), tl.get("prep", (0, 0))[0],
                      "d*8 read + 6*roundup(d,32) digit bytes + 16 B scale/norm written per candidate")
            hbm_entry("finalize", "k_finalize_acq", args.steps * m * (2.0 * nt2 * 8 + 8), tl.get("finalize", (0, 0))[0],
                      "2 x (n_pad/64) partial sums of 8 B read + 8 B acquisition value written per candidate")
        hbm_entry("prep_f64", "k_scale_inputs", f64_steps * m * (d * 8.0 + gp.d_pad * 8.0 + 8), tl_f64.get("prep", (0, 0))[0],
                  "d*8 read + d_pad*8 + 8 B written per candidate (FP64 mode)")
        hbm_entry("fit_solve", "k_tri_matvec x2 + k_fit_scalars", float(gp.n_pad) * gp.n_pad * 8.0 + gp.n_pad * 40.0, fit_tl.get("fit_solve", (0, 0))[0],
                  "lower triangle of L^-1 + upper triangle of L^-T (n_pad^2 * 8 B together) + vectors")
        cross_ops = pairs * float(args.steps) * m * 2.0 * n * d32
        if args.var_mode == "i8":
            alg_ops = pairs * alg_var
            achieved = alg_ops / (var_ms / 1e3) / 1e12 if var_ms > 0 else None
            bf16 = {"burst": 2 * peaks.get("bf16_tflops", 0.0), "sustained": 2 * peaks.get("bf16_tflops_sustained", 0.0)}
            roof = {
                "kernel": "k_var_i8 (V = K* L^-T as 21 exact int8 digit-pair GEMMs on tcgen05.mma kind::i8, TMEM accumulators, "
                          "triangular k-range, FP64 recombination)",
                "bound": "tensor", "achieved": achieved, "peak": i8_peak_sustained, "unit": "TOP/s",
                "unit_note": "int8 tensor operations per second, 2 per multiply-accumulate; NOT a DMMA fraction -- the FP64-equivalent "
                             "TFLOP/s of the exact integer emulation is given beside it and exceeds the DMMA ceiling by design",
                "frac": (achieved / i8_peak_sustained) if achieved else None, "traffic": traffic, "launches": int(var_n),
                "avg_launch_ms": var_ms / max(int(var_n), 1), "share_of_step": var_ms / ms_total,
                "algorithmic_ops_per_launch": alg_ops / max(int(var_n), 1),
                "fp64_equivalent_tflops": alg_var / (var_ms / 1e3) / 1e12 if var_ms > 0 else None,
                "fp64_dmma_peak_tflops": peak, "peak_burst": i8_peak, "frac_of_burst": (achieved / i8_peak) if achieved else None,
                "peak_source": "tcgen05.mma kind::i8 issue ceiling measured live by hdbo_i8_peak_probe (libhdbo_b200_dev.so; resident "
                               "operands, M128 N256 K32): `peak` = sustained (second half of ~1.5 s back to back, power-capped clocks, "
                               "the kernel is timed inside a long step), `peak_burst` = best short launch",
                # MEASURED_PEAKS.json has no int8 entry: 2 x its cuBLAS bf16 figures is the only driver-measured proxy (it is a cuBLAS GEMM
                # throughput, not an int8 issue ceiling -- fractions above 1 against it are possible and mean nothing more than that)
                "vs_2x_measured_bf16": {"peak_tops": bf16, "frac_of_sustained": (achieved / bf16["sustained"]) if achieved and bf16["sustained"] else None,
                                        "frac_of_burst": (achieved / bf16["burst"]) if achieved and bf16["burst"] else None},
                "cross_kernel": {"kernel": "k_cross_i8 (Zs Xs^T as 21 int8 digit-pair GEMMs + distance -> kernel -> digit epilogue)",
                                 "achieved": cross_ops / (cross_ms / 1e3) / 1e12 if cross_ms > 0 else None, "unit": "TOP/s",
                                 "peak": i8_peak_sustained, "avg_launch_ms": cross_ms / max(int(var_n), 1),
                                 "share_of_step": cross_ms / ms_total},
                "whole_step_fp64_equivalent_tflops": whole, "hbm_bound_phases": hbm,
            }
        else:
            achieved = alg_var / (var_ms / 1e3) / 1e12 if var_ms > 0 else None
            roof = {
                "kernel": "k_var (V = K* L^-T, triangular k-range, DMMA.8x8x4)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "launches": int(var_n), "avg_launch_ms": var_ms / max(int(var_n), 1),
                "share_of_step": var_ms / ms_total,
                "algorithmic_flops_per_launch": alg_var / max(int(var_n), 1),
                "peak_source": "FP64 DMMA.8x8x4 issue ceiling measured live by hdbo_fp64_peak_probe (MEASURED_PEAKS.json has "
                               "no FP64 entry; tcgen05 has no f64 kind, so the bf16 figure there does not bound this path)",
                "cross_kernel": {"achieved": alg_cross / (cross_ms / 1e3) / 1e12 if cross_ms > 0 else None, "unit": "TFLOP/s (FP64 DMMA)",
                                 "peak": peak, "share_of_step": cross_ms / ms_total},
                "whole_step_tflops": whole, "hbm_bound_phases": hbm,
            }
        fl_m = float(n) ** 3 + 3.0 * float(n) ** 2 * d
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if args.var_mode == "f64" else "f64 (both GEMMs of the step as exact int8 digit slices on tcgen05, FP64 recombination and epilogues)",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clk,
            "api": "LogExpectedImprovement(SingleTaskGP)(X) under torch.no_grad() -- the BoTorch surface, default var_mode",
            "gpu_launches": launches * args.steps, "argmax_index": int(best_i),
            "roofline": roof,
            "mll_fit_grad_ms": min(mll_ms),
            # the second metric of BASELINE.json against its own roofline: n^3 + 3 n^2 d FP64 flops on the DMMA ceiling
            "mll_roofline": {"bound": "tensor", "unit": "TFLOP/s", "algorithmic_flops": fl_m,
                             "achieved": fl_m / (min(mll_ms) * 1e-3) / 1e12, "peak": peak,
                             "frac": fl_m / (min(mll_ms) * 1e-3) / 1e12 / peak, "samples_ms": [round(t, 3) for t in mll_ms],
                             "how": "hdbo_gp_fit(need_r2) + hdbo_mll_grad replayed as one CUDA graph, as the fit loop's fast closure runs them",
                             "eager_ms": min(mll_eager_ms), "eager_frac": fl_m / (min(mll_eager_ms) * 1e-3) / 1e12 / peak,
                             "phases_ms": {k: round(v[0], 3) for k, v in fit_tl.items()}},
        }
        if mll_batched is not None:
            mll_batched.update({"achieved_tflops": fl_m / (mll_batched["ms_per_member"] * 1e-3) / 1e12,
                                "frac_of_fp64_dmma_peak": fl_m / (mll_batched["ms_per_member"] * 1e-3) / 1e12 / peak,
                                "note": "the trial points of one line search of the device fit = one batched fit + gradient"})
            out["mll_fit_grad_batched"] = mll_batched
        if mode_check is not None:
            out["var_mode_check"] = mode_check
        out["fp64_dmma_mode"] = f64_mode
        if e2e is not None:
            out["e2e"] = e2e
        if other is not None:
            out["other_scaling"] = other
        if extra:
            out["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            from oracle import gp_oracle as O  # the cpu_baseline leg: the only place the measured arm's process touches oracle/
            st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
            rate, reps_c, dt, batch, sweep = cpu_reference_rate(O, st, Z_host[:16384].clone(), best_f, args.cpu_sample)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "cpu_model": cpu_model_name(), "batch_sweep_evals_per_s": sweep,
                                   "sample": f"{reps_c} x {batch} candidates of the same pool in {dt:.1f} s (batch = best of the 1k..16k "
                                             "sweep), oracle/gp_oracle.py posterior+LogEI+argmax (torch FP64, MKL)"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
