"""ctypes binding of libhdbo_b200.so (the C ABI declared in include/hdbo_b200.h).

There is no CPU fallback: if the shared library is missing, or a call returns non-zero, this
module raises.  The library is built in-tree by `__graft_entry__.build()` /
`make -C hdbo_benchmark_b200/csrc` into hdbo_benchmark_b200/lib/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libhdbo_b200.so"
DEV_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libhdbo_b200_dev.so"

KERNEL_MATERN52 = 0
KERNEL_RBF = 1
ACQ_EI, ACQ_LOGEI, ACQ_UCB, ACQ_MEAN, ACQ_PI, ACQ_NONE = 0, 1, 2, 3, 4, -1
KR_FULL, KR_LE_N, KR_LE_M, KR_GE_N, KR_GE_M, KR_GE_MN = range(6)

_dp = C.c_void_p  # device pointers travel as integers


class HdboGp(C.Structure):
    """Mirror of `struct hdbo_gp` (include/hdbo_b200.h)."""

    _fields_ = [
        ("n", C.c_int32), ("d", C.c_int32), ("n_pad", C.c_int32), ("d_pad", C.c_int32),
        ("kind", C.c_int32), ("batch", C.c_int32),
        ("X", _dp), ("ldx", C.c_int64), ("y", _dp), ("center", _dp), ("inv_ls", _dp), ("hyp", _dp),
        ("Xs", _dp), ("xn", _dp), ("R2", _dp), ("Awork", _dp), ("L", _dp), ("Linv", _dp), ("Linvt", _dp),
        ("w", _dp), ("alpha", _dp), ("scal", _dp), ("info", _dp),
        ("timeline", C.c_void_p),   # HOST pointer to a caller-owned HdboTimeline, or NULL
        ("Kinv", _dp),              # optional: lower tiles of Khat^-1 left by hdbo_gp_fit(need_r2)
        ("aux_streams", C.c_void_p * 4), ("aux_events", C.c_void_p * 12),   # optional caller-owned side streams / events
    ]


class HdboTimeline(C.Structure):
    """Mirror of `struct hdbo_timeline`: caller-owned CUDA events the entry points record around their phases."""

    _fields_ = [("events", C.POINTER(C.c_void_p)), ("phase", C.POINTER(C.c_int32)), ("capacity", C.c_int32), ("used", C.c_int32)]


PHASE_NAMES = ("cross", "var", "prep", "finalize", "fit_build", "fit_factor", "fit_solve", "mll_grad", "bwd")


class HdboError(RuntimeError):
    pass


class HdboLbfgs(C.Structure):
    """Mirror of `struct hdbo_lbfgs` (include/hdbo_b200.h)."""
    _fields_ = [("R", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("T", C.c_int32),
                ("lo", C.c_void_p), ("hi", C.c_void_p), ("x", C.c_void_p), ("f", C.c_void_p), ("g", C.c_void_p),
                ("S", C.c_void_p), ("Y", C.c_void_p), ("rho", C.c_void_p), ("step", C.c_void_p), ("meta", C.c_void_p), ("Xt", C.c_void_p),
                ("ft", C.c_void_p), ("gt", C.c_void_p), ("ftol", C.c_double), ("gtol", C.c_double), ("maxiter", C.c_int32)]


_SIGNATURES = {
    "hdbo_version": (C.c_int, []),
    "hdbo_gp_fit": (C.c_int, [C.POINTER(HdboGp), C.c_int, _dp]),
    "hdbo_gp_fit_downdated": (C.c_int, [C.POINTER(HdboGp), C.c_int, _dp, C.c_int64, C.c_int64, _dp]),
    "hdbo_posterior_workspace_V": (C.c_void_p, [C.POINTER(HdboGp), C.c_int64, _dp]),
    "hdbo_mll_grad":(C.c_int, [C.POINTER(HdboGp), _dp, _dp, _dp, _dp, _dp, _dp]),
    "hdbo_mll_grad_part_bytes": (C.c_size_t, [C.POINTER(HdboGp)]),
    "hdbo_posterior_workspace_bytes": (C.c_size_t, [C.POINTER(HdboGp), C.c_int64, C.c_int]),
    "hdbo_posterior_acq": (C.c_int, [C.POINTER(HdboGp), _dp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                     _dp, _dp, _dp, C.c_int64, _dp, C.c_size_t, _dp]),
    "hdbo_i8_state_bytes": (C.c_size_t, [C.POINTER(HdboGp)]),
    "hdbo_gp_slice_linv": (C.c_int, [C.POINTER(HdboGp), _dp, _dp]),
    "hdbo_posterior_acq_i8": (C.c_int, [C.POINTER(HdboGp), _dp, _dp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                        _dp, _dp, _dp, C.c_int64, _dp, C.c_size_t, _dp]),
    "hdbo_lbfgs_propose": (C.c_int, [C.POINTER(HdboLbfgs), _dp]),
    "hdbo_lbfgs_accept": (C.c_int, [C.POINTER(HdboLbfgs), _dp]),
    "hdbo_i8_state_flag": (C.c_void_p, [C.POINTER(HdboGp), _dp]),
    "hdbo_posterior_fwd": (C.c_int, [C.POINTER(HdboGp), _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64,
                                     _dp, C.c_size_t, _dp]),
    "hdbo_posterior_joint_cov": (C.c_int, [C.POINTER(HdboGp), C.c_int64, C.c_int, C.c_int, _dp, C.c_int64, _dp,
                                           C.c_size_t, _dp]),
    "hdbo_posterior_bwd": (C.c_int, [C.POINTER(HdboGp), C.c_int64, C.c_int, _dp, _dp, _dp, C.c_int64, C.c_int64, _dp,
                                     C.c_int64, C.c_int64, _dp, C.c_size_t, _dp]),
    "hdbo_acq_elementwise": (C.c_int, [_dp, _dp, C.c_int64, C.c_int, C.c_double, _dp, _dp, _dp, _dp]),
    "hdbo_qei": (C.c_int, [_dp, _dp, _dp, C.c_int64, C.c_int, C.c_int64, C.c_double, C.c_double, _dp, _dp, _dp, _dp,
                           _dp]),
    "hdbo_sobol_draw": (C.c_int, [_dp, _dp, C.c_int, C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int64, C.c_int, _dp]),
    "hdbo_mvn_sample": (C.c_int, [_dp, _dp, _dp, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_double, _dp, _dp, _dp]),
    "hdbo_argmax": (C.c_int, [_dp, C.c_int64, _dp, _dp, _dp, _dp]),
    "hdbo_dgemm_nt": (C.c_int, [_dp, C.c_int64, _dp, C.c_int64, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                C.c_double, C.c_double, C.c_int, C.c_int, _dp]),
    "hdbo_cdist": (C.c_int, [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
    "hdbo_linear_f32": (C.c_int, [_dp, C.c_int64, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, C.c_int32, _dp, C.c_int64, C.c_int32, _dp]),
    "hdbo_argmax_tokens_f32": (C.c_int, [_dp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

# libhdbo_b200_dev.so (include/hdbo_b200_dev.h): test / measurement hooks, never loaded by the product path
_DEV_SIGNATURES = {
    "hdbo_test_math": (C.c_int, [_dp, C.c_int64, _dp, _dp, _dp]),
    "hdbo_test_math_lean": (C.c_int, [_dp, C.c_int64, _dp, _dp, _dp]),
    "hdbo_fp64_peak_probe": (C.c_int, [_dp, C.c_int, C.c_int, C.POINTER(C.c_double), _dp]),
    "hdbo_i8_peak_probe": (C.c_int, [_dp, C.c_int, C.c_int, C.POINTER(C.c_double), _dp]),
}
DEV_EXPORTED_SYMBOLS = tuple(_DEV_SIGNATURES)
_dev = None

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises HdboError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("HDBO_B200_LIB", LIB_PATH))
    if not path.exists():
        raise HdboError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the CUDA hot path)"
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if os.environ.get("HDBO_B200_DEV_PARTIAL"):  # developer builds of a subset of the ABI only
                continue
            raise HdboError(f"{path} does not export {name}; rebuild it")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_dev() -> C.CDLL:
    """Load the developer-hook library (tests and bench.py only)."""
    global _dev
    if _dev is not None:
        return _dev
    path = Path(os.environ.get("HDBO_B200_DEV_LIB", DEV_LIB_PATH))
    if not path.exists():
        raise HdboError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(str(path))
    for name, (res, args) in _DEV_SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise HdboError(f"{path} does not export {name}; rebuild it")
        fn.restype = res
        fn.argtypes = args
    _dev = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc > 0:
        raise HdboError(f"{what}: CUDA error {rc}")
    raise HdboError(f"{what}: invalid argument #{-rc}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
