#!/usr/bin/env python
"""Headline benchmark: acquisition evaluations per second at n=4096, d=256 (BASELINE.json metric, config #3).

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference arm: CPU FP64 restatement on the host cores

One step = posterior mean + variance + LogEI for every candidate of a scrambled-Sobol pool (2^20 per GPU) against a
fitted exact GP (n=4096, d=256, ARD Matern-5/2; L^-1 and alpha resident) + the arg-max.  `value` times the pool
resident in HBM; `e2e` times the same work from pinned HOST memory (H2D of the candidates and D2H of the
acquisition values inside the timed region).  Multi-GPU: every rank owns an independent pool shard (weak scaling),
the only collective is a 16-byte/rank all-gather of (max, argmax) per step.
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
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=4096)
    p.add_argument("--no-mll-batched", action="store_true", help="skip the batched (4 hyper-parameter vectors) MLL timing")
    p.add_argument("--d", type=int, default=256)
    p.add_argument("--pool", type=int, default=1 << 20, help="candidates per GPU per step")
    p.add_argument("--cpu-sample", type=int, default=8192, help="candidates per CPU-baseline batch")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--var-mode", default="i8", choices=["i8", "f64"],
                   help="variance contraction: exact int8 digit slices on tcgen05 (default) or FP64 DMMA")
    return p.parse_args()


def synthetic_problem(n, d, noise=1e-3, seed=0):
    """The synthetic workload of SURVEY.md §8d (same draws as oracle.gp_oracle.synthetic_problem(ls_mode="prior"); kept here so that
    the measured arm does not touch oracle/): X ~ U[0,1]^{n x d}, y = standardise(sin(3 sum_{k<4} x_k) + 0.1 eps),
    l_k = sqrt(d) exp(0.25 xi_k), outputscale 1, noise 1e-3, mean 0."""
    g0, g1, g2 = (torch.Generator().manual_seed(seed + i) for i in range(3))
    X = torch.rand(n, d, generator=g0, dtype=torch.float64)
    y = torch.sin(3.0 * X[:, : min(4, d)].sum(-1)) + 0.1 * torch.randn(n, generator=g1, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    ls = math.sqrt(d) * torch.exp(0.25 * torch.randn(d, generator=g2, dtype=torch.float64))
    return X, y, ls, 1.0, noise, 0.0


def workload_config(args, world):
    return {
        "workload": f"synthetic GP n={args.n}, d={args.d}, ARD Matern-5/2, posterior+LogEI+argmax over a 2^{int(math.log2(args.pool))}"
                    f"-candidate scrambled-Sobol pool per GPU (BASELINE.json configs[2])",
        "n_train": args.n, "d": args.d, "pool_per_gpu": args.pool, "pool_total": args.pool * world,
        "acquisition": "LogExpectedImprovement", "kernel": "matern52_ard", "noise": 1e-3,
        "parallelism": f"pool-shard x{world}", "var_mode": getattr(args, "var_mode", "f64"),
        "cache": "inputs larger than L2 (per launch: K* chunk 620 MB FP64 / 543 MB int8 digit images, L^-1 67 MB / 117 MB; 126 MB L2)",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
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


def cpu_reference_rate(args, st, Z_sample, best_f, budget_s=12.0):
    """The reference path on the host cores: plain-torch FP64 restatement (oracle) of BoTorch's
    model.posterior + LogExpectedImprovement, no grad, all threads."""
    from oracle import gp_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        m, v = O.posterior(st, Z_sample)  # warm-up
        O.log_expected_improvement(m, v, best_f)
        t0 = time.perf_counter()
        reps = 0
        while True:
            m, v = O.posterior(st, Z_sample)
            a = O.log_expected_improvement(m, v, best_f)
            int(a.argmax())
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or reps >= 200:
                break
    return reps * Z_sample.shape[0] / dt, reps, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gp_oracle as O

    X, y, ls, os_, nz, mu = O.synthetic_problem(args.n, args.d)
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
    Z = O.sobol_candidates(args.cpu_sample, args.d)
    best_f = float(y.max())
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            m, v = O.posterior(st, Z); O.log_expected_improvement(m, v, best_f)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m, v = O.posterior(st, Z)
            a = O.log_expected_improvement(m, v, best_f)
            int(a.argmax())
        dt = time.perf_counter() - t0
    value = args.steps * args.cpu_sample / dt
    cores = torch.get_num_threads()
    sample = f"{args.cpu_sample} Sobol candidates per step (bounded sample of the 2^20 pool), oracle/gp_oracle.py posterior+LogEI+argmax"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "botorch/gpytorch are not installable offline (not in /opt/wheelhouse); the reference arm is the CPU FP64 "
                "restatement of their exact path (oracle port), all host threads",
    }))


def _measured_bf16x2():
    try:
        pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return {"burst": 2 * pk["bf16_tflops"], "sustained": 2 * pk["bf16_tflops_sustained"]}
    except Exception:
        return {"burst": None, "sustained": None}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch.distributed as dist

    from hdbo_benchmark_b200 import _lib
    from hdbo_benchmark_b200.distributed import global_argmax
    from hdbo_benchmark_b200.engine import DEFAULT_CHUNK_ROWS, DeviceGP, _rup
    from hdbo_benchmark_b200.models import mll_gradient

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    n, d, m = args.n, args.d, args.pool
    X, y, ls, os_, nz, mu = synthetic_problem(n, d)
    best_f = float(y.max())
    gp = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, keep_r2=True, var_mode=args.var_mode)
    gp.set_hyper(ls.to(dev), os_, nz, mu)
    gp.fit()

    # ---- secondary metric: exact MLL value + full hyper-gradient (kernel build -> potrf -> inverse -> gradient) ----
    def mll_once():
        gp.fit_once(need_r2=True)
        return mll_gradient(gp)
    for _ in range(2):
        mll_once()
    torch.cuda.synchronize()
    mll_ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mll_once(); e1.record(); torch.cuda.synchronize()
        mll_ms.append(e0.elapsed_time(e1))
    gp.fit(need_r2=False)
    # the same evaluation as the device fit runs it (fit.fit_gpytorch_mll_device): the trial points of one line search = one batched
    # fit + gradient, so the latency-bound leaf chains of the members overlap.  Reported per member; rank 0 only.
    mll_batched = None
    if rank == 0 and not args.no_mll_batched:
        Bm = 4
        gpb = DeviceGP(X.to(dev), y.to(dev), _lib.KERNEL_MATERN52, batch=Bm, keep_r2=True)
        scale = torch.linspace(0.9, 1.1, Bm, dtype=torch.float64).unsqueeze(1)
        gpb.set_hyper((ls.unsqueeze(0) * scale).to(dev), os_, nz, mu)
        def mll_b():
            gpb.fit_once(need_r2=True)
            return mll_gradient(gpb)
        for _ in range(2):
            mll_b()
        torch.cuda.synchronize()
        tb = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mll_b(); e1.record(); torch.cuda.synchronize()
            tb.append(e0.elapsed_time(e1))
        mll_batched = {"batch": Bm, "ms_per_batch": min(tb), "ms_per_member": min(tb) / Bm}
        del gpb
        torch.cuda.empty_cache()

    # ---- candidate pool of this rank: scrambled Sobol, fast-forwarded to the rank's shard ----
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=3)
    if rank:
        eng.fast_forward(rank * m)
    Z_host = torch.empty(m, d, dtype=torch.float64).pin_memory()
    eng.draw(m, out=Z_host, dtype=torch.float64)
    Z_dev = Z_host.to(dev)
    acq_host = torch.empty(m, dtype=torch.float64).pin_memory()

    def step_resident():
        _, _, acq = gp.posterior_acq(Z_dev, _lib.ACQ_LOGEI, best_f, want_mean=False, want_var=False)
        v, i = gp.argmax(acq[0])
        return global_argmax(v, i, rank * m) if world > 1 else (v, i)

    def step_e2e():
        _, v, i = gp.posterior_acq_host(Z_host, _lib.ACQ_LOGEI, best_f, out_host=acq_host)
        return global_argmax(v, i, rank * m) if world > 1 else (v, i)

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

    # ---- FP64 tensor peak, measured live (denominator of the roofline) ----
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fl = C.c_double(0.0)
    peak = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.hdbo_fp64_peak_probe(scratch.data_ptr(), 16384, sms, C.byref(fl), torch.cuda.current_stream().cuda_stream), "peak")
        e1.record(); torch.cuda.synchronize()
        peak = max(peak, fl.value / e0.elapsed_time(e1) / 1e9)

    # ---- int8 tensor ceiling, measured live, and the in-run parity guard of the int8 mode against the FP64 DMMA mode ----
    i8_peak, i8_peak_sustained, mode_check, f64_mode = None, None, None, None
    if args.var_mode == "i8":
        flag = torch.zeros(4, dtype=torch.int32, device=dev)
        ops = C.c_double(0.0)
        i8_peak = 0.0          # burst: best of 6 short launches
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.hdbo_i8_peak_probe(flag.data_ptr(), 4096, sms, C.byref(ops), torch.cuda.current_stream().cuda_stream), "i8 peak")
            e1.record(); torch.cuda.synchronize()
            i8_peak = max(i8_peak, ops.value / e0.elapsed_time(e1) / 1e9)
        # sustained: ~1.5 s of back-to-back launches, the second half timed (power-capped clocks, like the kernel inside a long step)
        reps = max(8, int(1.5 / (ops.value / (i8_peak * 1e12) * 8)))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for r in range(reps):
            if r == reps // 2:
                evs[0].record()
            _lib.check(lib.hdbo_i8_peak_probe(flag.data_ptr(), 32768, sms, C.byref(ops), torch.cuda.current_stream().cuda_stream), "i8 peak")
        evs[1].record(); torch.cuda.synchronize()
        i8_peak_sustained = ops.value * (reps - reps // 2) / evs[0].elapsed_time(evs[1]) / 1e9
        assert int(flag[0].item()) == 0, "int8 peak probe timed out"
        Zc = Z_dev[: 1 << 14]
        _, var_i8, acq_i8 = gp.posterior_acq(Zc, _lib.ACQ_LOGEI, best_f, want_mean=False)
        gp.var_mode = "f64"
        _, var_f64, acq_f64 = gp.posterior_acq(Zc, _lib.ACQ_LOGEI, best_f, want_mean=False)
        step_resident()
        ms_f64, _ = timed(step_resident, 1)
        gp.var_mode = "i8"
        f64_mode = {"value": m * world / (ms_f64 / 1e3), "unit": UNIT, "steps": 1}
        mode_check = {"candidates": int(Zc.shape[0]),
                      "max_abs_var_diff_over_outputscale": float((var_i8 - var_f64).abs().max() / float(os_)),
                      "max_rel_var_diff": float(((var_i8 - var_f64).abs() / var_f64.abs()).max()),
                      "same_argmax": bool(int(acq_i8[0].argmax()) == int(acq_f64[0].argmax())), "timeout_flag": gp.i8_flag()}
        assert mode_check["max_rel_var_diff"] < 1e-6 and mode_check["same_argmax"] and mode_check["timeout_flag"] == 0, mode_check

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    lib.hdbo_profile_enable(1)
    ms_total, (best_v, best_i) = timed(step_resident, args.steps)
    prof_ms = (C.c_double * 2)()
    prof_n = (C.c_int64 * 2)()
    lib.hdbo_profile_collect(prof_ms, prof_n)
    lib.hdbo_profile_enable(0)
    clk = clocks.stop()
    value = args.steps * m * world / (ms_total / 1e3)

    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ms_e2e, (best_v2, best_i2) = timed(step_e2e, args.steps)
        assert int(best_i2) == int(best_i), "host path and resident path disagree on the arg-max"
        e2e = {"value": args.steps * m * world / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": m * d * 8 * world, "d2h_bytes_per_step": (m * 8 + 16) * world,
               "api": "DeviceGP.posterior_acq_host (pinned host candidates in, pinned host acquisition values + argmax out)"}

    if rank == 0:
        chunks = math.ceil(m / DEFAULT_CHUNK_ROWS)
        launches = 4 * chunks + 1  # scale, cross, var, finalize per chunk + argmax
        var_ms, var_n = prof_ms[1], prof_n[1]
        cross_ms = prof_ms[0]
        alg_var = float(args.steps) * m * float(n) * n  # n^2 flops per candidate (triangular V = K* L^-T)
        alg_cross = float(args.steps) * m * 2.0 * n * d
        traffic = None
        ncu_path = ROOT / "profiles" / "ncu_summary.json"
        key = "k_var_i8" if args.var_mode == "i8" else "k_var"
        if ncu_path.exists():
            try:
                traffic = json.loads(ncu_path.read_text()).get(key, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cross = {"achieved": alg_cross / (cross_ms / 1e3) / 1e12 if cross_ms > 0 else None, "unit": "TFLOP/s (FP64 DMMA)",
                 "peak": peak, "share_of_step": cross_ms / ms_total}
        whole = args.steps * m * (2.0 * n * d + float(n) * n + 2 * n) / (ms_total / 1e3) / 1e12
        if args.var_mode == "i8":
            # 21 digit-pair products (6 x 6 base-256 digits, s + t <= 5) of the triangular contraction, 2 integer ops per MAC
            pairs = 21.0
            alg_ops = pairs * alg_var
            achieved = alg_ops / (var_ms / 1e3) / 1e12 if var_ms > 0 else None
            cross_ops = pairs * float(args.steps) * m * 2.0 * n * (32 * math.ceil(d / 32))
            cross = {"kernel": "k_cross_i8 (Zs Xs^T as 21 int8 digit-pair GEMMs + FP64 distance -> kernel -> digit epilogue)",
                     "achieved": cross_ops / (cross_ms / 1e3) / 1e12 if cross_ms > 0 else None, "unit": "TOP/s",
                     "peak": i8_peak_sustained, "share_of_step": cross_ms / ms_total,
                     "note": "tile time = MMA time + epilogue FP64 time (the FP64 pipe is starved while tcgen05.mma streams: "
                             "profiles/r01_fp64_vs_tensor.jsonl); 26 FP64 instructions per K* element, see DESIGN.md 2b"}
            roof = {
                "kernel": "k_var_i8 (V = K* L^-T as 21 exact int8 digit-pair GEMMs on tcgen05.mma kind::i8, TMEM accumulators, "
                          "triangular k-range, FP64 recombination)",
                "bound": "tensor", "achieved": achieved, "peak": i8_peak_sustained, "unit": "TOP/s",
                "unit_note": "int8 tensor operations per second, 2 per multiply-accumulate (the FP64-equivalent TFLOP/s is given beside it)",
                "frac": (achieved / i8_peak_sustained) if achieved else None, "traffic": traffic, "launches": int(var_n),
                "avg_launch_ms": var_ms / max(int(var_n), 1), "share_of_step": var_ms / ms_total,
                "algorithmic_ops_per_launch": alg_ops / max(int(var_n), 1),
                "fp64_equivalent_tflops": alg_var / (var_ms / 1e3) / 1e12 if var_ms > 0 else None,
                "fp64_dmma_peak_tflops": peak, "peak_burst": i8_peak, "frac_of_burst": (achieved / i8_peak) if achieved else None,
                "peak_source": "tcgen05.mma kind::i8 issue ceiling measured live by hdbo_i8_peak_probe (resident operands, M128 N256 "
                               "K32): `peak` = sustained (second half of ~1.5 s back to back, power-capped clocks, the kernel is timed "
                               "inside a long step), `peak_burst` = best short launch.  MEASURED_PEAKS.json has no int8 entry -- for "
                               "context 2 x its bf16 figures: " + json.dumps(_measured_bf16x2()),
                "cross_kernel": cross, "whole_step_fp64_equivalent_tflops": whole,
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
                "cross_kernel": cross, "whole_step_tflops": whole,
            }
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.var_mode == "f64" else "f64 (both GEMMs of the step as exact int8 digit slices on tcgen05, FP64 recombination and epilogues)",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clk,
            "gpu_launches": launches * args.steps, "argmax_index": int(best_i),
            "roofline": roof,
            "mll_fit_grad_ms": min(mll_ms),
            # the second metric of BASELINE.json against its own roofline: n^3 + 3 n^2 d FP64 flops on the DMMA ceiling
            "mll_roofline": {"bound": "tensor", "unit": "TFLOP/s", "algorithmic_flops": float(args.n) ** 3 + 3.0 * float(args.n) ** 2 * args.d,
                             "achieved": (float(args.n) ** 3 + 3.0 * float(args.n) ** 2 * args.d) / (min(mll_ms) * 1e-3) / 1e12, "peak": peak,
                             "frac": (float(args.n) ** 3 + 3.0 * float(args.n) ** 2 * args.d) / (min(mll_ms) * 1e-3) / 1e12 / peak,
                             "note": "a dependency chain, not a throughput limit: 32 leaves + 114 small GEMMs per evaluation run on a few SMs "
                                     "(2.7 ms) next to 2.4 ms of large GEMMs at 28-32 TF/s; DESIGN.md 8 (next steps)"},
        }
        if mll_batched is not None:
            fl_m = float(args.n) ** 3 + 3.0 * float(args.n) ** 2 * args.d
            mll_batched.update({"achieved_tflops": fl_m / (mll_batched["ms_per_member"] * 1e-3) / 1e12,
                                "frac_of_fp64_dmma_peak": fl_m / (mll_batched["ms_per_member"] * 1e-3) / 1e12 / peak,
                                "note": "the trial points of one line search of the device fit = one batched fit + gradient; the leaf chains overlap"})
            out["mll_fit_grad_batched"] = mll_batched
        if mode_check is not None:
            out["var_mode_check"] = mode_check
            out["fp64_dmma_mode"] = f64_mode
        if e2e is not None:
            out["e2e"] = e2e
        if not args.no_cpu_baseline:
            from oracle import gp_oracle as O  # the cpu_baseline leg: the only place the measured arm's process touches oracle/
            st = O.fit_state(X, y, ls, os_, nz, mu, O.MATERN52)
            Zs = Z_host[: args.cpu_sample].clone()
            rate, reps, dt = cpu_reference_rate(args, st, Zs, best_f)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{reps} x {args.cpu_sample} candidates of the same pool in {dt:.1f} s, "
                                             "oracle/gp_oracle.py posterior+LogEI+argmax (torch FP64, MKL)"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
