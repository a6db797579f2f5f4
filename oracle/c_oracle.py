"""ctypes front-end of the C oracle (oracle/c/mvtv_oracle.c).  TEST INFRASTRUCTURE ONLY: importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libmvtv_oracle.so")

MAXP, MAXK = 6, 63
MODE_CPP, MODE_RCPP, MODE_PY = 0, 1, 2
VARIANT_REFERENCE, VARIANT_INTENDED = 0, 1
SOLVER_BANDCHOL, SOLVER_PCG = 0, 1


class OraOp(C.Structure):
    _fields_ = [("p", C.c_int), ("K", C.c_int),
                ("m", C.c_int64 * MAXP), ("stride", C.c_int64 * MAXP), ("N", C.c_int64), ("R", C.c_int64),
                ("mask", C.c_int * MAXK), ("scale", C.c_double * MAXK),
                ("rows", C.c_int64 * MAXK), ("row_off", C.c_int64 * MAXK),
                ("rstride", (C.c_int64 * MAXP) * MAXK)]


class OraParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("variant", C.c_int), ("solver", C.c_int),
                ("lam", C.c_double), ("rho_init", C.c_double), ("rho_matrix0", C.c_double),
                ("tol", C.c_double), ("max_counter", C.c_int), ("max_passes", C.c_int),
                ("cg_rtol", C.c_double), ("cg_maxit", C.c_int), ("nthreads", C.c_int)]


class OraResult(C.Structure):
    _fields_ = [("counter", C.c_int), ("passes", C.c_int), ("status", C.c_int),
                ("rho", C.c_double), ("r_norm", C.c_double), ("s_norm", C.c_double),
                ("inner_iters", C.c_int64), ("seconds", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "c", "mvtv_oracle.c")
    hdr = os.path.join(_HERE, "c", "mvtv_oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c"), "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.ora_op_init.argtypes = [C.POINTER(OraOp), C.c_int, ip, dp, C.c_int]
        L.ora_op_init.restype = C.c_int
        L.ora_D_apply.argtypes = [C.POINTER(OraOp), dp, dp]
        L.ora_Dt_apply.argtypes = [C.POINTER(OraOp), dp, dp]
        L.ora_nearest.argtypes = [C.c_int, ip, dp, C.c_int64, dp, ip]
        L.ora_nearest_brute.argtypes = [C.c_int, ip, dp, C.c_int64, dp, ip]
        L.ora_scatter.argtypes = [C.c_int64, ip, dp, C.c_int64, dp, dp]
        L.ora_admm.argtypes = [C.POINTER(OraOp), dp, dp, C.c_double, C.POINTER(OraParams), dp, dp, dp, dp,
                               C.POINTER(OraResult)]
        L.ora_admm.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class Operator:
    def __init__(self, m, deltas=None, variant=VARIANT_REFERENCE):
        self.m = np.ascontiguousarray(m, dtype=np.int64)
        self.p = len(self.m)
        self.op = OraOp()
        d = None if deltas is None else np.ascontiguousarray(deltas, dtype=np.float64)
        rc = lib().ora_op_init(C.byref(self.op), self.p, _ip(self.m), _dp(d), variant)
        if rc != 0:
            raise ValueError("matrix multiplication: incompatible matrix dimensions (non-cubic mesh with "
                             "the reference mixedpartial quirk)" if rc == -1 else "bad p")
        self.N, self.R, self.K = int(self.op.N), int(self.op.R), int(self.op.K)

    def D(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.empty(self.R)
        lib().ora_D_apply(C.byref(self.op), _dp(theta), _dp(out))
        return out

    def Dt(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.empty(self.N)
        lib().ora_Dt_apply(C.byref(self.op), _dp(w), _dp(out))
        return out

    @property
    def masks(self):
        return [int(self.op.mask[b]) for b in range(self.K)]

    @property
    def scales(self):
        return [float(self.op.scale[b]) for b in range(self.K)]

    @property
    def block_rows(self):
        return [int(self.op.rows[b]) for b in range(self.K)]


def nearest(m, axes, data, brute=False):
    m = np.ascontiguousarray(m, dtype=np.int64)
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    n, p = data.shape
    dcm = np.ascontiguousarray(data.T).ravel()     # column-major n x p
    ax = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64) for a in axes]))
    idx = np.empty(n, dtype=np.int64)
    fn = lib().ora_nearest_brute if brute else lib().ora_nearest
    fn(p, _ip(m), _dp(ax), n, _dp(dcm), _ip(idx))
    return idx


def scatter(idx, y, N):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    Oty, cnt = np.empty(N), np.empty(N)
    lib().ora_scatter(len(idx), _ip(idx), _dp(y), N, _dp(Oty), _dp(cnt))
    return Oty, cnt


def admm(op: Operator, Oty, counts, mean_y, lam, mode=MODE_CPP, theta_init=None, u_init=None,
         rho_init=None, rho_matrix0=None, tol=0.0, max_counter=0, max_passes=0, solver=SOLVER_BANDCHOL,
         cg_rtol=1e-13, cg_maxit=0, nthreads=0):
    prm = OraParams()
    prm.mode, prm.variant, prm.solver = mode, 0, solver
    prm.lam = lam
    prm.rho_init = (lam / 5.0) if rho_init is None else rho_init
    prm.rho_matrix0 = lam if rho_matrix0 is None else rho_matrix0
    prm.tol, prm.max_counter, prm.max_passes = tol, max_counter, max_passes
    prm.cg_rtol, prm.cg_maxit, prm.nthreads = cg_rtol, cg_maxit, nthreads
    Oty = np.ascontiguousarray(Oty, dtype=np.float64)
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    th0 = None if theta_init is None else np.ascontiguousarray(theta_init, dtype=np.float64)
    u0 = None if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
    theta, u = np.empty(op.N), np.empty(op.R)
    res = OraResult()
    rc = lib().ora_admm(C.byref(op.op), _dp(Oty), _dp(counts), float(mean_y), C.byref(prm), _dp(th0), _dp(u0),
                        _dp(theta), _dp(u), C.byref(res))
    if rc != 0:
        raise RuntimeError("ora_admm failed rc=%d" % rc)
    return dict(theta=theta, u=u, rho=res.rho, counter=res.counter, passes=res.passes, status=res.status,
                r_norm=res.r_norm, s_norm=res.s_norm, inner_iters=int(res.inner_iters), seconds=res.seconds)


def mbs_one(data, y, m, axes, lam, mode=MODE_CPP, deltas=None, variant=VARIANT_REFERENCE, **kw):
    """mbs_one with cache==NULL (cpp-code/solvers.cpp:134-152): O, D, Oty, admm, fitted = O*theta."""
    y = np.asarray(y, dtype=np.float64).ravel()
    op = Operator(m, deltas, variant)
    idx = nearest(m, axes, data)
    Oty, counts = scatter(idx, y, op.N)
    out = admm(op, Oty, counts, float(np.mean(y)), lam, mode=mode, **kw)
    out["fitted"] = out["theta"][idx]
    out["idx"], out["Oty"], out["counts"] = idx, Oty, counts
    return out
