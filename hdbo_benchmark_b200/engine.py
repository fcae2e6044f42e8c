"""Device-resident exact-GP state and the calls into libhdbo_b200.so.

`DeviceGP` owns (through torch's caching allocator) every buffer the C ABI needs for one GP or a
batch of GPs that share training data but differ in hyper-parameters (fully Bayesian / SAAS
samples).  It is the thin layer between the BoTorch-shaped classes in `models.py` /
`acquisition.py` and the C ABI; it contains no arithmetic of its own.

Data layout in HBM (all FP64, row-major, padded: n_pad = roundup(n,128), d_pad = roundup(d,16)):
    Xs   [B][n_pad][d_pad]   centred inputs / lengthscale          xn    [B][n_pad]  row norms^2
    Awork[B][n_pad][n_pad]   K + noise I (consumed)                 R2    [B][n_pad][n_pad] (MLL gradient only)
    L, Linv (lower), Linvt (upper) [B][n_pad][n_pad]                w, alpha [B][n_pad]
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import HdboGp, check, ptr

DEFAULT_CHUNK_ROWS = 148 * 128  # one 128-row tile per SM per column tile
WORKSPACE_CAP_BYTES = 8 << 30
I8_MAX_N_PAD = 8192   # int8 mode: int32 group sums stay exact up to this size (csrc/i8.cuh); larger GPs use the FP64 path
# int8 mode: the variance error is ~2^-48 ||L^-1 row|| in units of the prior, i.e. it grows like (outputscale / noise)^1/2, and on a
# training point the variance itself is ~noise: relative error ~ c (outputscale / noise)^3/2 2^-48 -- measured 2.3e-6 at a ratio of
# 1.6e-5, 1.5e-7 at 1e-4.  Below this noise / outputscale ratio the engine keeps the FP64 path (BoTorch's own floor is 1e-4).
I8_MIN_NOISE_RATIO = 1e-4
I8_MAX_D = 16384      # int8 mode: the same bound for the cross GEMM (K = d)
I8_DEFAULT_MIN_POOL = 1024
PIPELINE_MIN_TILES = 8   # n_pad >= 1024: the factorisation runs pipelined over side streams (single GPs only)


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class NotPSDError(RuntimeError):
    """Raised when K + noise I is not positive definite even after the jitter schedule
    (linear_operator.utils.errors.NotPSDError upstream)."""


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_gp_device(fn):
    """Run a DeviceGP method with the GP's device current: kernel launches and `_stream()` follow the CUDA current device, which
    need not be the device the model lives on (HDBO_TORCH_DEVICE=cuda:1 in a process whose current device is cuda:0)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapped


class DeviceGP:
    def __init__(self, X: torch.Tensor, y: torch.Tensor, kind: int, batch: int = 1, keep_r2: bool = False,
                 var_mode: Optional[str] = None):
        """var_mode: "i8" (the default: both GEMMs of the no-grad pool evaluation as exact digit-sliced int8 contractions on the
        tcgen05 tensor cores; applies when n_pad <= 8192, d <= 16384 and noise >= 1e-4 outputscale -- BoTorch's own noise floor for
        standardised targets -- and silently keeps the FP64 DMMA path otherwise) or "f64" (always FP64 DMMA).  $HDBO_VAR_MODE
        overrides the default."""
        if not X.is_cuda:
            raise _lib.HdboError("DeviceGP needs CUDA tensors: the hot path has no CPU implementation")
        self.lib = _lib.load()
        self.device = torch.device("cuda", X.device.index if X.device.index is not None else torch.cuda.current_device())
        self.X = X.detach().to(torch.float64).contiguous()
        self.y = y.detach().to(torch.float64).reshape(-1).contiguous()
        self.n, self.d = self.X.shape
        assert self.y.numel() == self.n
        self.kind = int(kind)
        self.batch = int(batch)
        self.n_pad = _rup(self.n, 128)
        self.d_pad = _rup(self.d, 16)
        self.keep_r2 = keep_r2
        B, npad, dpad = self.batch, self.n_pad, self.d_pad
        f64 = dict(dtype=torch.float64, device=self.device)
        self.center = self.X.mean(dim=0).contiguous()
        self.inv_ls = torch.ones(B, self.d, **f64)
        self.hyp = torch.zeros(B, 4, **f64)
        self.Xs = torch.empty(B, npad, dpad, **f64)
        self.xn = torch.empty(B, npad, **f64)
        self.R2 = torch.empty(B, npad, npad, **f64) if keep_r2 else None
        self.Awork = torch.empty(B, npad, npad, **f64)
        self.L = torch.empty(B, npad, npad, **f64)
        # the strictly-upper tiles of Linv / strictly-lower tiles of Linvt are never written nor read
        self.Linv = torch.zeros(B, npad, npad, **f64)
        self.Linvt = torch.zeros(B, npad, npad, **f64)
        self.w = torch.empty(B, npad, **f64)
        self.alpha = torch.empty(B, npad, **f64)
        self.scal = torch.zeros(B, 8, **f64)
        self.info = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._ws: Optional[torch.Tensor] = None
        self._argmax_work = torch.zeros(8192, dtype=torch.uint8, device=self.device)
        self.fitted = False
        self.jitter_used = 0.0
        explicit = var_mode or os.environ.get("HDBO_VAR_MODE")
        self.var_mode = (explicit or "i8").lower()
        # defaulted mode: pools below this many candidates stay on the FP64 kernels (their small-tile variants win there, and a BO
        # loop that refits between a handful of posterior calls never pays for slicing L^-1 into digit planes)
        self.i8_min_pool = 0 if explicit else I8_DEFAULT_MIN_POOL
        if self.var_mode not in ("f64", "i8"):
            raise ValueError(f"var_mode must be 'f64' or 'i8', got {self.var_mode!r}")
        self._i8_state: Optional[torch.Tensor] = None
        self._i8_dirty = True
        self._i8_ok = True
        self._i8_redo = []
        self.i8_min_noise_ratio = I8_MIN_NOISE_RATIO
        self.desc = HdboGp(
            n=self.n, d=self.d, n_pad=npad, d_pad=dpad, kind=self.kind, batch=B,
            X=ptr(self.X), ldx=self.X.stride(0), y=ptr(self.y), center=ptr(self.center), inv_ls=ptr(self.inv_ls),
            hyp=ptr(self.hyp), Xs=ptr(self.Xs), xn=ptr(self.xn), R2=ptr(self.R2), Awork=ptr(self.Awork), L=ptr(self.L),
            Linv=ptr(self.Linv), Linvt=ptr(self.Linvt), w=ptr(self.w), alpha=ptr(self.alpha), scal=ptr(self.scal),
            info=ptr(self.info), timeline=None,
        )
        self._timeline = None
        # K^-1 for the MLL gradient: left by hdbo_gp_fit(need_r2) in its own buffer, so that the pipelined factorisation can
        # accumulate it beside the chain (the gradient kernels used to overwrite L with it afterwards)
        self.Kinv = torch.empty(B, npad, npad, **f64) if keep_r2 else None
        self.desc.Kinv = ptr(self.Kinv)
        self._kinv_valid = False
        self._aux = None
        if B == 1 and npad // 128 >= PIPELINE_MIN_TILES and os.environ.get("HDBO_FIT_PIPELINE", "1") != "0":
            self._attach_side_streams()

    # ---- caller-owned side streams / events of the pipelined factorisation (csrc/kernels_chol_pipe.cu) ----
    def _attach_side_streams(self) -> None:
        """Four side streams -- chain (highest priority: its one-CTA leaves and single-tile updates must never queue), side (the
        persistent rank-256 updates, lowest priority), near (rows of L^-1 next to the chain) and panel (the rest of each panel and
        the next pair's column updates, second priority: the chain waits for it once per column) -- and twelve events, owned here
        and handed to the library through the descriptor."""
        with torch.cuda.device(self.device):
            streams = [torch.cuda.Stream(device=self.device, priority=p) for p in (-3, 0, -1, -2)]   # chain, side, near, panel
            events = [torch.cuda.Event() for _ in range(12)]
            for e in events:
                e.record()          # forces the lazily created cudaEvent_t into existence
        self._aux = (streams, events)
        for i, st in enumerate(streams):
            self.desc.aux_streams[i] = st.cuda_stream
        for i, e in enumerate(events):
            self.desc.aux_events[i] = e.cuda_event

    # ---- caller-owned timeline: CUDA events recorded by the library around its phases (bench.py's per-kernel figures) ----
    @_on_gp_device
    def attach_timeline(self, capacity: int = 4096) -> None:
        """Create `capacity` timing events (owned here, i.e. by the caller of the C ABI) and hand them to the library."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(capacity)]
        for e in evs:
            e.record()      # torch creates the cudaEvent_t lazily: force it so that the raw handle exists
        handles = (C.c_void_p * capacity)(*[e.cuda_event for e in evs])
        phases = (C.c_int32 * capacity)()
        tl = _lib.HdboTimeline(events=C.cast(handles, C.POINTER(C.c_void_p)), phase=C.cast(phases, C.POINTER(C.c_int32)),
                               capacity=capacity, used=0)
        self._timeline = (tl, evs, handles, phases)
        self.desc.timeline = C.cast(C.pointer(tl), C.c_void_p)

    def detach_timeline(self) -> None:
        self.desc.timeline = None
        self._timeline = None

    def reset_timeline(self) -> None:
        if self._timeline is not None:
            self._timeline[0].used = 0

    def collect_timeline(self) -> dict:
        """{phase name: (total ms, launches)} of everything recorded since the last reset.  Synchronises."""
        out = {}
        if self._timeline is None:
            return out
        tl, evs, _, phases = self._timeline
        torch.cuda.synchronize(self.device)
        for i in range(0, tl.used - 1, 2):
            name = _lib.PHASE_NAMES[phases[i]]
            ms, cnt = out.get(name, (0.0, 0))
            out[name] = (ms + evs[i].elapsed_time(evs[i + 1]), cnt + 1)
        tl.used = 0
        return out

    # ---- hyper-parameters (device tensors; no host sync) ----
    def set_hyper(self, lengthscale, outputscale, noise, mean, jitter: float = 0.0) -> None:
        """lengthscale [d] or [B,d]; outputscale / noise / mean scalars or [B] (tensors or floats)."""
        f64 = dict(dtype=torch.float64, device=self.device)
        ls = torch.as_tensor(lengthscale, **f64).detach().reshape(-1, self.d)
        self.inv_ls.copy_((1.0 / ls).expand(self.batch, self.d))
        vals = (outputscale, noise, mean, jitter)
        if all(torch.is_tensor(v) and v.is_cuda for v in vals[:3]):   # fit loop: already on the device, one fused write
            cols = [v.detach().reshape(-1).expand(self.batch) for v in vals[:3]]
            cols.append(torch.full((self.batch,), float(jitter), **f64))
            self.hyp.copy_(torch.stack(cols, dim=1))
        else:
            host = torch.stack([torch.as_tensor(v, dtype=torch.float64).detach().cpu().reshape(-1).expand(self.batch) for v in vals], dim=1)
            self.hyp.copy_(host)
        self.fitted = False

    # ---- A1 + A2: kernel build, Cholesky, inverse, alpha, logdet ----
    @_on_gp_device
    def fit_once(self, need_r2: Optional[bool] = None) -> None:
        need = self.keep_r2 if need_r2 is None else need_r2
        check(self.lib.hdbo_gp_fit(C.byref(self.desc), int(bool(need)), _stream()), "hdbo_gp_fit")
        self._i8_dirty = True
        self._kinv_valid = bool(need) and self.Kinv is not None

    @_on_gp_device
    def fit(self, need_r2: Optional[bool] = None, max_tries: int = 3) -> float:
        """Factorise with the psd_safe_cholesky jitter schedule (1e-8 * 10**i, i < max_tries).  One host sync."""
        self.fit_once(need_r2)
        jitter = 0.0
        if int(self.info.max().item()) != 0:
            for i in range(max_tries):
                jitter = 1e-8 * (10 ** i)
                self.hyp[:, 3] = jitter
                self.fit_once(need_r2)
                if int(self.info.max().item()) == 0:
                    break
            else:
                self.hyp[:, 3] = 0.0
                raise NotPSDError(f"Matrix not positive definite after adding jitter up to {jitter:.1e}")
            self.hyp[:, 3] = 0.0
        self.fitted = True
        self.jitter_used = jitter
        self._i8_dirty = True      # int8 mode: the digit images are re-sliced lazily by the next pool evaluation
        if self.var_mode == "i8":  # (fit() has just synchronised on `info`: one more tiny read)
            hy = self.hyp[:, :2].cpu()
            good = (hy[:, 1] + jitter) >= self.i8_min_noise_ratio * hy[:, 0]
            # per-GP guard: a SAAS batch keeps the int8 path when only a few of its hyper-parameter samples fall below the noise
            # ratio -- those few are re-evaluated by the FP64 kernels (sub-batch descriptors) after the batched int8 pass
            self._i8_redo = [int(i) for i in (~good).nonzero().reshape(-1)]
            self._i8_ok = len(self._i8_redo) * 2 <= self.batch and len(self._i8_redo) < self.batch
        return jitter

    # ---- exact int8 variance mode: digit planes of L^-1, once per fit ----
    @_on_gp_device
    def slice_linv(self) -> None:
        if self._i8_state is None:
            need = self.lib.hdbo_i8_state_bytes(C.byref(self.desc))
            if need == 0:
                raise _lib.HdboError("hdbo_i8_state_bytes rejected its arguments")
            self._i8_state = torch.empty(need, dtype=torch.uint8, device=self.device)
        check(self.lib.hdbo_gp_slice_linv(C.byref(self.desc), ptr(self._i8_state), _stream()), "hdbo_gp_slice_linv")
        self._i8_dirty = False

    def i8_flag(self) -> int:
        """Non-zero if a pipeline wait of the int8 variance kernel ever timed out (its outputs are then NaN).  Host sync."""
        if self._i8_state is None:
            return 0
        return int(self._i8_state[-16:-12].view(torch.int32).item())

    def _pool_call(self, Zptr, ldz, m, observation_noise, acq_kind, acq_param, mean, var, acq, bso, ws, stream) -> None:
        if (self.var_mode == "i8" and self.n_pad <= I8_MAX_N_PAD and self.d <= I8_MAX_D and self._i8_ok
                and m * self.batch >= self.i8_min_pool):
            if self._i8_dirty or self._i8_state is None:
                self.slice_linv()
            check(self.lib.hdbo_posterior_acq_i8(C.byref(self.desc), ptr(self._i8_state), Zptr, ldz, m, int(observation_noise),
                                                 acq_kind, float(acq_param), mean, var, acq, bso, ptr(ws), ws.numel(), stream),
                  "hdbo_posterior_acq_i8")
            for b in self._i8_redo:          # members below the noise-ratio guard: FP64 kernels, same outputs, same stream
                off = lambda p: (p + 8 * b * bso) if p else 0
                check(self.lib.hdbo_posterior_acq(C.byref(self._sub_desc(b, b + 1)), Zptr, ldz, m, int(observation_noise), acq_kind,
                                                  float(acq_param), off(mean), off(var), off(acq), bso, ptr(ws), ws.numel(), stream),
                      "hdbo_posterior_acq")
        else:
            check(self.lib.hdbo_posterior_acq(C.byref(self.desc), Zptr, ldz, m, int(observation_noise), acq_kind,
                                              float(acq_param), mean, var, acq, bso, ptr(ws), ws.numel(), stream),
                  "hdbo_posterior_acq")

    def _sub_desc(self, b0: int, b1: int) -> HdboGp:
        """Descriptor of the members [b0, b1) of a batched GP (same training data, offset per-member arrays)."""
        npad, dpad, d = self.n_pad, self.d_pad, self.d
        o = lambda t, per: (t.data_ptr() + 8 * b0 * per) if t is not None else None
        return HdboGp(
            n=self.n, d=d, n_pad=npad, d_pad=dpad, kind=self.kind, batch=b1 - b0,
            X=ptr(self.X), ldx=self.X.stride(0), y=ptr(self.y), center=ptr(self.center), inv_ls=o(self.inv_ls, d),
            hyp=o(self.hyp, 4), Xs=o(self.Xs, npad * dpad), xn=o(self.xn, npad), R2=o(self.R2, npad * npad),
            Awork=o(self.Awork, npad * npad), L=o(self.L, npad * npad), Linv=o(self.Linv, npad * npad),
            Linvt=o(self.Linvt, npad * npad), w=o(self.w, npad), alpha=o(self.alpha, npad), scal=o(self.scal, 8),
            info=self.info.data_ptr() + 4 * b0, timeline=None,
        )

    # ---- workspace ----
    def workspace(self, rows: int, keep: bool = False) -> torch.Tensor:
        need = self.lib.hdbo_posterior_workspace_bytes(C.byref(self.desc), rows, int(keep))
        if need == 0:
            raise _lib.HdboError("hdbo_posterior_workspace_bytes rejected its arguments")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- A4 + A5: posterior mean / variance / acquisition over a candidate pool (no grad) ----
    @_on_gp_device
    def posterior_acq(self, Z: torch.Tensor, acq_kind: int = _lib.ACQ_NONE, acq_param: float = 0.0,
                      observation_noise: bool = False, want_mean: bool = True, want_var: bool = True,
                      chunk_rows: int = DEFAULT_CHUNK_ROWS, out: Optional[dict] = None):
        assert self.fitted, "call fit() first"
        Z = Z.detach()
        if Z.dtype != torch.float64 or Z.stride(-1) != 1:
            Z = Z.to(torch.float64).contiguous()
        m = Z.shape[0]
        rows = min(_rup(max(m, 1), 128), _rup(chunk_rows, 128))
        # keep the resident K* slab of a batched (fully Bayesian) GP below ~8 GiB
        per_row = 8 * self.batch * (self.n_pad + self.d_pad + 2 * (self.n_pad // 128) + 1)
        rows = max(128, min(rows, (WORKSPACE_CAP_BYTES // per_row) // 128 * 128))
        ws = self.workspace(rows)
        f64 = dict(dtype=torch.float64, device=self.device)
        out = out or {}
        mean = out.get("mean") if "mean" in out else (torch.empty(self.batch, m, **f64) if want_mean else None)
        var = out.get("var") if "var" in out else (torch.empty(self.batch, m, **f64) if want_var else None)
        acq = out.get("acq") if "acq" in out else (torch.empty(self.batch, m, **f64) if acq_kind >= 0 else None)
        self._pool_call(ptr(Z), Z.stride(0), m, observation_noise, acq_kind, acq_param, ptr(mean), ptr(var), ptr(acq), m, ws,
                        _stream())
        return mean, var, acq

    @_on_gp_device
    def posterior_acq_host(self, Z_host: torch.Tensor, acq_kind: int, acq_param: float, out_host: Optional[torch.Tensor] = None,
                           copy_rows: int = 2 * DEFAULT_CHUNK_ROWS, chunk_rows: int = DEFAULT_CHUNK_ROWS, slab_transform=None):
        """End-to-end pool evaluation from HOST memory (pinned for full overlap): candidates are streamed to the device in
        `copy_rows` slabs on a copy stream (double-buffered) while the previous slab is evaluated; the acquisition values
        [m] come back to `out_host` and the arg-max (value, index) is reduced on the device.  batch == 1 only.
        `slab_transform` (optional): in-place map applied to each slab on the device before it is evaluated (the model's input
        transform when the call comes through the BoTorch surface)."""
        assert self.fitted and self.batch == 1
        m, d = Z_host.shape
        f64 = dict(dtype=torch.float64, device=self.device)
        main = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slabs = None
        if self._slabs is None or self._slabs[0].shape[0] < min(copy_rows, m) or self._slabs[0].shape[1] != d:
            self._slabs = [torch.empty(min(copy_rows, m), d, **f64) for _ in range(2)]
        rows = min(_rup(min(copy_rows, m), 128), _rup(chunk_rows, 128))
        ws = self.workspace(rows)
        acq = torch.empty(m, **f64)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        for ev in free:
            ev.record(main)
        step = self._slabs[0].shape[0]
        # the first slab's copy is the only one nothing overlaps: keep it short (37 row tiles; the persistent kernels fill the SMs
        # from the column tiles) -- at N = 8 a strong-scaled rank has 3.5 slabs per step and a full first slab cost 10 % of it
        first = min(step, 37 * 128)
        pieces = [(0, min(first, m))] + [(off, min(step, m - off)) for off in range(first, m, step)]
        for i, (off, mc) in enumerate(pieces):
            slab = self._slabs[i % 2]
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(free[i % 2])
                slab[:mc].copy_(Z_host[off:off + mc], non_blocking=True)
                ready[i % 2].record(self._copy_stream)
            main.wait_event(ready[i % 2])
            if slab_transform is not None:
                slab_transform(slab[:mc])
            self._pool_call(ptr(slab), slab.stride(0), mc, 0, acq_kind, acq_param, 0, 0, acq.data_ptr() + off * 8, m, ws,
                            main.cuda_stream)
            free[i % 2].record(main)
        val, idx = self.argmax(acq)
        if out_host is not None:
            out_host.copy_(acq, non_blocking=True)
        return acq, val, idx

    # ---- A4 + A5 + A6 without the autograd tape: acquisition value and d/dZ for m points (device L-BFGS inner loop) ----
    @_on_gp_device
    def acq_value_grad(self, Z: torch.Tensor, acq_kind: int, acq_param: float, sign: float = 1.0, bufs: Optional[dict] = None):
        """Z [m, d] float64 contiguous on the device -> (acq [m], dacq/dZ [m, d]); `sign` = -1 evaluates the acquisition of
        the negated mean (minimisation problems).  `bufs` (a dict the caller keeps) caches workspace and outputs across calls."""
        assert self.fitted and self.batch == 1
        m, d = Z.shape
        bufs = bufs if bufs is not None else {}
        if bufs.get("m") != m:
            f64 = dict(dtype=torch.float64, device=self.device)
            rows = _rup(m, 128)
            need = self.lib.hdbo_posterior_workspace_bytes(C.byref(self.desc), rows, 1)
            bufs.update(m=m, ws=torch.empty(need, dtype=torch.uint8, device=self.device), mean=torch.empty(m, **f64),
                        var=torch.empty(m, **f64), acq=torch.empty(m, **f64), dmean=torch.empty(m, **f64),
                        dvar=torch.empty(m, **f64), dZ=torch.empty(m, d, **f64))
        b, st, lib = bufs, _stream(), self.lib
        check(lib.hdbo_posterior_fwd(C.byref(self.desc), ptr(Z), Z.stride(0), m, 0, ptr(b["mean"]), ptr(b["var"]), m, ptr(b["ws"]),
                                     b["ws"].numel(), st), "hdbo_posterior_fwd")
        if sign < 0:
            b["mean"].neg_()
        check(lib.hdbo_acq_elementwise(ptr(b["mean"]), ptr(b["var"]), m, acq_kind, float(acq_param), ptr(b["acq"]), ptr(b["dmean"]),
                                       ptr(b["dvar"]), st), "hdbo_acq_elementwise")
        if sign < 0:
            b["dmean"].neg_()
        check(lib.hdbo_posterior_bwd(C.byref(self.desc), m, 1, ptr(b["dmean"]), ptr(b["dvar"]), 0, m, 0, ptr(b["dZ"]), d, m * d,
                                     ptr(b["ws"]), b["ws"].numel(), st), "hdbo_posterior_bwd")
        return b["acq"], b["dZ"]

    # ---- A4 + A7 + A6 without the autograd tape: qEI value and d/dZ for R q-batches (device L-BFGS inner loop, q > 1) ----
    @_on_gp_device
    def qei_value_grad(self, Z: torch.Tensor, q: int, base: torch.Tensor, best_f: float, jitter: float = 1e-8,
                       bufs: Optional[dict] = None):
        """Z [R*q, d] float64 contiguous (R q-batches back to back), base [S, q] fixed standard-normal base samples ->
        (qEI [R], dqEI/dZ [R*q, d]) in the model's standardised units:
        hdbo_posterior_fwd -> hdbo_posterior_joint_cov -> hdbo_qei (value + adjoints of mean / Sigma) -> hdbo_posterior_bwd."""
        assert self.fitted and self.batch == 1 and 1 <= q <= 8
        m, d = Z.shape
        R = m // q
        assert R * q == m
        bufs = bufs if bufs is not None else {}
        if bufs.get("m") != m or bufs.get("q") != q:
            f64 = dict(dtype=torch.float64, device=self.device)
            need = self.lib.hdbo_posterior_workspace_bytes(C.byref(self.desc), _rup(m, 128), 1)
            bufs.update(m=m, q=q, ws=torch.empty(need, dtype=torch.uint8, device=self.device), mean=torch.empty(m, **f64),
                        var=torch.empty(m, **f64), Sigma=torch.empty(R, q, q, **f64), val=torch.empty(R, **f64),
                        gmean=torch.empty(m, **f64), gSigma=torch.empty(R, q, q, **f64), dZ=torch.empty(m, d, **f64),
                        info=torch.zeros(R, dtype=torch.int32, device=self.device))
        b, st, lib = bufs, _stream(), self.lib
        check(lib.hdbo_posterior_fwd(C.byref(self.desc), ptr(Z), Z.stride(0), m, 0, ptr(b["mean"]), ptr(b["var"]), m, ptr(b["ws"]),
                                     b["ws"].numel(), st), "hdbo_posterior_fwd")
        check(lib.hdbo_posterior_joint_cov(C.byref(self.desc), m, q, 0, ptr(b["Sigma"]), R * q * q, ptr(b["ws"]), b["ws"].numel(), st),
              "hdbo_posterior_joint_cov")
        check(lib.hdbo_qei(ptr(b["mean"]), ptr(b["Sigma"]), ptr(base), R, q, base.shape[0], float(best_f), float(jitter), ptr(b["val"]),
                           ptr(b["gmean"]), ptr(b["gSigma"]), ptr(b["info"]), st), "hdbo_qei")
        check(lib.hdbo_posterior_bwd(C.byref(self.desc), m, q, ptr(b["gmean"]), 0, ptr(b["gSigma"]), m, R * q * q, ptr(b["dZ"]), d,
                                     m * d, ptr(b["ws"]), b["ws"].numel(), st), "hdbo_posterior_bwd")
        return b["val"], b["dZ"]

    @_on_gp_device
    def argmax(self, v: torch.Tensor):
        v = v.reshape(-1)
        out_val = torch.empty(1, dtype=torch.float64, device=self.device)
        out_idx = torch.empty(1, dtype=torch.int64, device=self.device)
        check(self.lib.hdbo_argmax(ptr(v), v.numel(), ptr(self._argmax_work), ptr(out_val), ptr(out_idx), _stream()),
              "hdbo_argmax")
        return out_val, out_idx

    # ---- convenience accessors (unpadded views) ----
    @property
    def logdet(self) -> torch.Tensor:
        return self.scal[:, 0]

    @property
    def mll_data_term(self) -> torch.Tensor:
        """n * mll without priors: -1/2 r^T Khat^-1 r - 1/2 logdet - n/2 log 2pi."""
        return self.scal[:, 2]
