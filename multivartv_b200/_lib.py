"""ctypes binding of include/mvtv.h (libmvtv_b200.so).  No torch types cross this boundary.

The library is hand-written CUDA for sm_100a and has no CPU fallback: loading fails loudly when the
shared object has not been built, and every compute entry point fails with MVTV_ERR_CUDA when no
CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

MAXP = 4
OK, ERR_INVALID, ERR_CUDA, ERR_NOT_CONVERGED, ERR_DIM_MISMATCH, ERR_UNSUPPORTED, ERR_INNER_SOLVE = range(7)
MODE_CPP, MODE_RCPP, MODE_PY = 0, 1, 2
VARIANT_REFERENCE, VARIANT_INTENDED = 0, 1
F64, F32 = 64, 32
PRECOND_JACOBI, PRECOND_CHEB1, PRECOND_AUTO, PRECOND_CHEB2, PRECOND_CHEB3, PRECOND_CHEB4 = 0, 1, 2, 3, 4, 5
WARM_THETA_FROM_PLAN, WARM_U_FROM_PLAN = 1, 2
KC_NAMES = ["zu", "zu_init", "cg_init", "cg_step", "cg_update", "cg_prec"]
KC_N = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmvtv_b200.so")

SYMBOLS = [
    "mvtv_abi_version", "mvtv_last_error", "mvtv_device_count", "mvtv_nccl_unique_id", "mvtv_host_alloc", "mvtv_host_free", "mvtv_plan_create", "mvtv_plan_destroy",
    "mvtv_plan_info", "mvtv_plan_describe", "mvtv_plan_profile", "mvtv_plan_get_profile", "mvtv_plan_set_points", "mvtv_plan_set_points_strided", "mvtv_plan_set_points_dev", "mvtv_plan_get_cache", "mvtv_solve", "mvtv_solve_path", "mvtv_lambda_max",
    "mvtv_predict", "mvtv_apply_D", "mvtv_apply_Dt", "mvtv_apply_M", "mvtv_softthresh", "mvtv_adapt_step", "mvtv_nearest",
]


class PlanDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("p", C.c_int32), ("m", C.c_int64 * MAXP),
                ("dtype", C.c_int32), ("variant", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32), ("reserved", C.c_int32),
                ("deltas", C.POINTER(C.c_double)), ("nccl_unique_id", C.c_void_p)]


class SolveParams(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("mode", C.c_int32), ("lam", C.c_double),
                ("rho_init", C.c_double), ("rho_matrix0", C.c_double), ("tol", C.c_double),
                ("max_counter", C.c_int32), ("max_passes", C.c_int32), ("cg_rtol", C.c_double),
                ("cg_maxit", C.c_int32), ("precond", C.c_int32), ("flags", C.c_uint32), ("timing_skip_passes", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("counter", C.c_int32), ("passes", C.c_int32), ("status", C.c_int32), ("reserved", C.c_int32),
                ("rho", C.c_double), ("r_norm", C.c_double), ("s_norm", C.c_double), ("max_dtheta", C.c_double),
                ("inner_iters", C.c_int64), ("device_seconds", C.c_double), ("kernel_launches", C.c_int64),
                ("timed_inner_iters", C.c_int64), ("timed_kernel_launches", C.c_int64), ("timed_passes", C.c_int32),
                ("reserved2", C.c_int32)]


class MvtvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mvtv error %d: %s" % (code, msg))
        self.code = code


class NotConverged(MvtvError):
    """The reference's std::invalid_argument("Failed to converge!") (cpp-code/solvers.cpp:122-124)."""


_lib = None


def load():
    """Load libmvtv_b200.so (build it first with ``python -m multivartv_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m multivartv_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.mvtv_abi_version.restype = C.c_int
    L.mvtv_last_error.restype = C.c_char_p
    L.mvtv_device_count.argtypes = [C.POINTER(C.c_int)]
    L.mvtv_nccl_unique_id.argtypes = [vp]
    L.mvtv_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.mvtv_host_free.argtypes = [vp]
    L.mvtv_plan_create.argtypes = [C.POINTER(vp), C.POINTER(PlanDesc)]
    L.mvtv_plan_destroy.argtypes = [vp]
    L.mvtv_plan_info.argtypes = [vp, ip, ip, ip, ip]
    L.mvtv_plan_describe.argtypes = [vp, C.c_char_p, C.c_int64]
    L.mvtv_plan_profile.argtypes = [vp, C.c_int]
    L.mvtv_plan_get_profile.argtypes = [vp, dp, ip]
    L.mvtv_plan_set_points.argtypes = [vp, C.c_int64, dp, dp, dp]
    L.mvtv_plan_set_points_strided.argtypes = [vp, C.c_int64, dp, C.c_int64, C.c_int64, dp, dp]
    L.mvtv_plan_set_points_dev.argtypes = [vp, C.c_int64, vp, vp, vp]
    L.mvtv_plan_get_cache.argtypes = [vp, dp, dp, ip]
    L.mvtv_solve.argtypes = [vp, C.POINTER(SolveParams), dp, dp, dp, dp, C.POINTER(SolveResult)]
    L.mvtv_solve_path.argtypes = [vp, C.POINTER(SolveParams), C.c_int32, dp, dp, dp, C.POINTER(C.c_int32), dp, dp, dp, dp,
                                  C.POINTER(C.c_int32), C.POINTER(SolveResult)]
    L.mvtv_lambda_max.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int32)]
    L.mvtv_predict.argtypes = [vp, C.c_int64, dp, dp, dp, dp]
    L.mvtv_apply_D.argtypes = [vp, dp, dp]
    L.mvtv_apply_Dt.argtypes = [vp, dp, dp]
    L.mvtv_apply_M.argtypes = [vp, C.c_double, dp, dp]
    L.mvtv_softthresh.argtypes = [C.c_int64, dp, C.c_double, dp]
    L.mvtv_adapt_step.argtypes = [C.c_int, C.c_int64, dp, C.c_int64, dp, C.c_double, C.c_int64, dp, dp, dp]
    L.mvtv_nearest.argtypes = [C.c_int, ip, dp, C.c_int64, dp, ip]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("mvtv_last_error", "mvtv_abi_version"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(code, allow=()):
    if code == OK or code in allow:
        return code
    msg = load().mvtv_last_error().decode("utf-8", "replace")
    if code == ERR_NOT_CONVERGED:
        raise NotConverged(code, msg)
    raise MvtvError(code, msg)
