"""Host-side mirror of the reference's solver interface for the hot path, on top of the C ABI.

Names, argument meaning and outputs follow the reference:
  * ``mbs_one`` / ``mbs_predict`` / ``mbs_mse`` / ``softthresh``  -- code/solvers.py:9-89 (dict API) and
    cpp-code/solvers.hpp:22,85-96 (the C++ entry points they shadow)
  * ``create_mesh`` / ``mesh_axes`` / ``create_deltas`` / ``nearest1`` -- cpp-code/utils.cpp:271-330,
    rcpp-code/MultivarTV/src/utils.cpp:234-287, code/utils.py:153-193
Everything numerical runs in libmvtv_b200.so (CUDA, sm_100a); this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref

import numpy as np

from . import _lib
from ._lib import (F32, F64, MODE_CPP, MODE_PY, MODE_RCPP, PRECOND_AUTO, PRECOND_CHEB1, PRECOND_CHEB2, PRECOND_CHEB3, PRECOND_CHEB4,  # noqa: F401
                   PRECOND_JACOBI, VARIANT_INTENDED,
                   VARIANT_REFERENCE, WARM_THETA_FROM_PLAN, WARM_U_FROM_PLAN, MvtvError, NotConverged)

_MODES = {"cpp": MODE_CPP, "rcpp": MODE_RCPP, "py": MODE_PY, MODE_CPP: MODE_CPP, MODE_RCPP: MODE_RCPP,
          MODE_PY: MODE_PY}


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(data):
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    return np.ascontiguousarray(data.T).ravel(), data.shape[0], data.shape[1]


# ----------------------------------------------------------------------------------------------
# mesh helpers (host logic, O(sum m_k))
# ----------------------------------------------------------------------------------------------
def _arma_linspace(start, end, num):
    num = int(num)
    if num == 1:
        return np.array([end], dtype=np.float64)
    x = start + np.arange(num, dtype=np.float64) * ((end - start) / float(num - 1))
    x[-1] = end
    return x


def mesh_axes(data, m, mode="cpp"):
    """Per-axis knots of create_mesh: cpp-code/utils.cpp:281 (min+EPS..max+EPS, EPS=0.01, stored in a
    float matrix), rcpp utils.cpp:242 (min-EPS..max+EPS, EPS=1e-4, double), code/utils.py:184 (eps=0.01)."""
    mode = _MODES[mode]
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    axes = []
    for k in range(data.shape[1]):
        lo, hi = float(data[:, k].min()), float(data[:, k].max())
        if mode == MODE_CPP:
            ax = _arma_linspace(lo + 0.01, hi + 0.01, m[k]).astype(np.float32).astype(np.float64)
        elif mode == MODE_RCPP:
            ax = _arma_linspace(lo - 0.0001, hi + 0.0001, m[k])
        else:
            ax = np.linspace(lo - 0.01, hi + 0.01, int(m[k]))
        axes.append(ax)
    return axes


def create_mesh(data, m, mode="cpp"):
    """N x p coordinate table, row i = knots at vector2tensor(i) (cpp-code/utils.cpp:271-298)."""
    return mesh_from_axes(mesh_axes(data, m, mode))


def mesh_from_axes(axes):
    dims = [len(a) for a in axes]
    N = int(np.prod(dims))
    lin = np.arange(N, dtype=np.int64)
    mesh = np.empty((N, len(axes)), dtype=np.float64)
    stride = 1
    for k, ax in enumerate(axes):
        mesh[:, k] = np.asarray(ax)[(lin // stride) % dims[k]]
        stride *= dims[k]
    return mesh


def axes_from_mesh(mesh, m):
    """Recover the knot vectors from a tensor-product mesh matrix (the ``MAT mesh`` argument of
    cpp-code/solvers.hpp:89)."""
    mesh = np.asarray(mesh, dtype=np.float64)
    if mesh.ndim == 1:
        mesh = mesh[:, None]
    axes, stride = [], 1
    for k in range(len(m)):
        axes.append(np.ascontiguousarray(mesh[::stride, k][: int(m[k])]))
        stride *= int(m[k])
    return axes


def create_deltas(data, m, mode="cpp"):
    """cpp-code/utils.cpp:300-307 (EPS=0.01) ; rcpp utils.cpp:256-263 (EPS=1e-4)."""
    eps = 0.01 if _MODES[mode] == MODE_CPP else 0.0001
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    return np.array([(data[:, i].max() - data[:, i].min() + 2 * eps) / float(m[i]) for i in range(data.shape[1])])


def nearest1(data, mesh=None, m=None, axes=None):
    """cpp-code/utils.cpp:323-330 on a tensor-product mesh (device kernel k_bin)."""
    if axes is None:
        axes = axes_from_mesh(mesh, m)
    m = np.ascontiguousarray([len(a) for a in axes], dtype=np.int64)
    dcm, n, p = _colmajor(data)
    ax = _f64(np.concatenate(axes))
    out = np.empty(n, dtype=np.int64)
    _lib.check(_lib.load().mvtv_nearest(p, _ip(m), _dp(ax), n, _dp(dcm), _ip(out)))
    return out


def softthresh(z, lam):
    """cpp-code/solvers.hpp:22 / code/solvers.py:9-12."""
    z = _f64(z)
    out = np.empty_like(z)
    _lib.check(_lib.load().mvtv_softthresh(z.size, _dp(z.ravel()), float(lam), _dp(out.ravel())))
    return out


def adapt_step(r, s, rho, u, mode="cpp"):
    """cpp-code/solvers.hpp:77-82 (solvers.cpp:70-88) / rcpp solvers.cpp:77-94: returns (rho_next, u_next)."""
    r, s, u = _f64(np.asarray(r).ravel()), _f64(np.asarray(s).ravel()), _f64(np.asarray(u).ravel())
    u_next = np.empty_like(u)
    rho_next = C.c_double(0.0)
    _lib.check(_lib.load().mvtv_adapt_step(_MODES[mode], r.size, _dp(r), s.size, _dp(s), float(rho), u.size, _dp(u),
                                           C.byref(rho_next), _dp(u_next)))
    return rho_next.value, u_next


def pinned_empty(shape, dtype=np.float64, order="C"):
    """numpy array in page-locked host memory (mvtv_host_alloc): host<->device copies of arrays allocated here
    run as direct DMA transfers.  The memory is released when the last view of the array dies."""
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(v) for v in shape)
    dt = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape)) * dt.itemsize)
    L = _lib.load()
    ptr = C.c_void_p()
    _lib.check(L.mvtv_host_alloc(C.byref(ptr), nbytes))
    buf = (C.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, L.mvtv_host_free, ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)))
    return arr.reshape(shape, order=order)


def _out_buffer(out, size, name):
    """Caller-provided output array (e.g. from pinned_empty): must be contiguous float64 of the right size."""
    if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.size == size
            and out.flags.writeable):
        raise ValueError("%s must be a writeable C-contiguous float64 array of %d elements" % (name, size))
    return out


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (generate on rank 0, broadcast, pass to every rank's Plan)."""
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().mvtv_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


# ----------------------------------------------------------------------------------------------
# Plan = the cached operators of create_cache_objects (cpp-code/solvers.cpp:31-41) on the device
# ----------------------------------------------------------------------------------------------
class Plan:
    def __init__(self, m, deltas=None, dtype=F64, variant=VARIANT_REFERENCE, device=-1, rank=0, world=1,
                 nccl_unique_id: bytes | None = None):
        L = _lib.load()
        m = [int(v) for v in np.asarray(m).ravel()]
        self.m, self.p = m, len(m)
        d = _lib.PlanDesc()
        d.struct_size = C.sizeof(_lib.PlanDesc)
        d.p = self.p
        for k in range(min(self.p, _lib.MAXP)):
            d.m[k] = m[k]
        d.dtype, d.variant, d.device, d.rank, d.world = dtype, variant, device, rank, world
        self._deltas = None if deltas is None or len(deltas) == 0 else _f64(deltas)
        d.deltas = _dp(self._deltas)
        self._uid = None
        if nccl_unique_id is not None:
            self._uid = C.create_string_buffer(bytes(nccl_unique_id), 128)
            d.nccl_unique_id = C.cast(self._uid, C.c_void_p)
        self._h = C.c_void_p()
        if self.p < 1 or self.p > _lib.MAXP:
            raise MvtvError(_lib.ERR_INVALID, "p must be 1..4")
        _lib.check(L.mvtv_plan_create(C.byref(self._h), C.byref(d)))
        N, R, z0, nz = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(L.mvtv_plan_info(self._h, C.byref(N), C.byref(R), C.byref(z0), C.byref(nz)))
        self.N, self.R, self.z0, self.nz = N.value, R.value, z0.value, nz.value
        self.plane = self.N // m[-1] if self.p > 1 else self.N
        self.n_local = self.plane * self.nz if self.p > 1 else self.N
        self.dtype, self.world, self.rank = dtype, world, rank
        self.n = 0
        self.axes = None

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().mvtv_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- points (A = O) ----------------------------------------------------------------------
    def set_points(self, data, y, axes):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1:
            data = data[:, None]
        n, p = data.shape
        assert p == self.p, "data has %d columns, mesh has %d axes" % (p, self.p)
        # pass the matrix in the layout it already has (no host transpose): F-order = arma's column-major
        if data.flags.f_contiguous:
            ld_point, ld_axis = 1, n
        else:
            data = np.ascontiguousarray(data)
            ld_point, ld_axis = p, 1
        y = _f64(np.asarray(y).ravel())
        assert y.size == n
        self.axes = [np.asarray(a, dtype=np.float64) for a in axes]
        ax = _f64(np.concatenate(self.axes))
        assert ax.size == sum(self.m)
        _lib.check(_lib.load().mvtv_plan_set_points_strided(self._h, n, _dp(data), ld_point, ld_axis, _dp(y), _dp(ax)))
        self.n = n
        return self

    def set_points_dev(self, n, data_ptr, y_ptr, axes_ptr, axes=None):
        """Inputs already resident in HBM: raw device addresses (e.g. ``torch.Tensor.data_ptr()``)."""
        _lib.check(_lib.load().mvtv_plan_set_points_dev(self._h, int(n), C.c_void_p(data_ptr), C.c_void_p(y_ptr),
                                                        C.c_void_p(axes_ptr)))
        self.n = int(n)
        if axes is not None:
            self.axes = axes
        return self

    def cache(self):
        """(Oty, diag(crossO), nearest vertex per point) -- cpp-code/solvers.cpp:32,37,40."""
        Oty, cnt = np.empty(self.n_local), np.empty(self.n_local)
        vid = np.empty(self.n, dtype=np.int64)
        _lib.check(_lib.load().mvtv_plan_get_cache(self._h, _dp(Oty), _dp(cnt), _ip(vid)))
        return Oty, cnt, vid

    # -- the hot path ------------------------------------------------------------------------
    def solve(self, lam, mode="cpp", theta_init=None, u_init=None, rho_init=None, rho_matrix0=None, tol=None,
              max_counter=0, max_passes=0, cg_rtol=0.0, cg_maxit=0, precond=PRECOND_AUTO, flags=0,
              want_u=False, want_fitted=True, want_theta=True, raise_on_nonconvergence=True, theta_out=None,
              fitted_out=None, timing_skip_passes=0):
        """admm_update + fitted (mvtv_solve).  ``theta_out`` / ``fitted_out``: optional preallocated float64 arrays
        (``pinned_empty``) that receive the outputs instead of fresh numpy arrays."""
        L = _lib.load()
        mode = _MODES[mode]
        prm = _lib.SolveParams()
        prm.struct_size = C.sizeof(_lib.SolveParams)
        prm.mode, prm.lam = mode, float(lam)
        prm.rho_init = math.nan if rho_init is None else float(rho_init)
        prm.rho_matrix0 = math.nan if rho_matrix0 is None else float(rho_matrix0)
        prm.tol = math.nan if tol is None else float(tol)
        prm.max_counter, prm.max_passes = int(max_counter), int(max_passes)
        prm.cg_rtol, prm.cg_maxit, prm.precond, prm.flags = float(cg_rtol), int(cg_maxit), int(precond), int(flags)
        prm.timing_skip_passes = int(timing_skip_passes)
        th0 = None if theta_init is None else _f64(np.asarray(theta_init).ravel())
        if th0 is not None:
            assert th0.size == self.n_local
        u = None
        if u_init is not None:
            u = _f64(np.asarray(u_init).ravel()).copy()
            assert u.size == self.R
        elif want_u:
            if mode == MODE_RCPP and not (flags & WARM_U_FROM_PLAN):
                u = np.zeros(self.R)
            else:
                u = np.empty(self.R)  # CPP/PY ignore the input, only the output is used
        theta = (_out_buffer(theta_out, self.n_local, "theta_out") if theta_out is not None
                 else np.empty(self.n_local)) if want_theta else None
        fitted = (_out_buffer(fitted_out, self.n, "fitted_out") if fitted_out is not None
                  else np.empty(self.n)) if want_fitted else None
        res = _lib.SolveResult()
        allow = () if raise_on_nonconvergence else (_lib.ERR_NOT_CONVERGED,)
        if mode == MODE_RCPP:
            allow = (_lib.ERR_NOT_CONVERGED,)  # rcpp prints and breaks (rcpp solvers.cpp:129-132)
        code = L.mvtv_solve(self._h, C.byref(prm), _dp(th0), _dp(u), _dp(theta), _dp(fitted), C.byref(res))
        _lib.check(code, allow)
        return dict(theta=theta, fitted=fitted, u=u, rho=res.rho, counter=res.counter, passes=res.passes,
                    status=res.status, r_norm=res.r_norm, s_norm=res.s_norm, max_dtheta=res.max_dtheta,
                    inner_iters=int(res.inner_iters), device_seconds=res.device_seconds,
                    kernel_launches=int(res.kernel_launches), timed_passes=int(res.timed_passes),
                    timed_inner_iters=int(res.timed_inner_iters), timed_kernel_launches=int(res.timed_kernel_launches))

    def solve_path(self, lambdas, ftrue, mode="cpp", tol=None, max_counter=0, cg_rtol=0.0, cg_maxit=0,
                   rho_init=None, want_thetas=False, want_best=True, precond=PRECOND_AUTO):
        """mbs_path (cpp-code/solvers.cpp:196-217 ; rcpp solvers.cpp:204-222): warm-started lambda path that stays
        on the device between lambdas.  Returns MSEs, Counters and the first minimum-MSE model."""
        L = _lib.load()
        mode = _MODES[mode]
        lambdas = _f64(np.asarray(lambdas).ravel())
        ftrue = _f64(np.asarray(ftrue).ravel())
        assert ftrue.size == self.n
        prm = _lib.SolveParams()
        prm.struct_size = C.sizeof(_lib.SolveParams)
        prm.mode, prm.lam = mode, float(lambdas[0])
        prm.rho_init = math.nan if rho_init is None else float(rho_init)
        prm.rho_matrix0 = math.nan
        prm.tol = math.nan if tol is None else float(tol)
        prm.max_counter, prm.cg_rtol, prm.cg_maxit = int(max_counter), float(cg_rtol), int(cg_maxit)
        prm.precond = int(precond)
        nl = lambdas.size
        mses = np.empty(nl)
        counters = np.zeros(nl, dtype=np.int32)
        rhos = np.zeros(nl)
        thetas = np.empty((nl, self.n_local)) if want_thetas else None
        tb = np.empty(self.n_local) if want_best else None
        fb = np.empty(self.n) if want_best else None
        best = C.c_int32(0)
        res = _lib.SolveResult()
        code = L.mvtv_solve_path(self._h, C.byref(prm), nl, _dp(lambdas), _dp(ftrue), _dp(mses),
                                 counters.ctypes.data_as(C.POINTER(C.c_int32)), _dp(rhos), _dp(thetas), _dp(tb), _dp(fb),
                                 C.byref(best), C.byref(res))
        _lib.check(code)
        return dict(mses=mses, counters=counters, rhos=rhos, thetas=thetas, theta_best=tb, fitted_best=fb, best_index=best.value,
                    minmse=float(mses[best.value]), minmse_lambda=float(lambdas[best.value]), passes=res.passes,
                    inner_iters=int(res.inner_iters), device_seconds=res.device_seconds, rho=res.rho,
                    kernel_launches=int(res.kernel_launches), status=res.status)

    def lambda_max(self, mode="cpp"):
        """lam_max_pinv (cpp-code/utils.cpp:399-404 ; rcpp utils.cpp:351-355): returns (lambda_max, cg_iterations)."""
        lam = C.c_double(0.0)
        it = C.c_int32(0)
        _lib.check(_lib.load().mvtv_lambda_max(self._h, _MODES[mode], C.byref(lam), C.byref(it)))
        return lam.value, it.value

    def predict(self, data, theta=None, axes=None):
        """mbs_predict (cpp-code/solvers.cpp:154-158)."""
        dcm, n, p = _colmajor(data)
        axes = self.axes if axes is None else axes
        ax = _f64(np.concatenate(axes))
        th = None if theta is None else _f64(np.asarray(theta).ravel())
        out = np.empty(n)
        _lib.check(_lib.load().mvtv_predict(self._h, n, _dp(dcm), _dp(ax), _dp(th), _dp(out)))
        return out

    def describe(self):
        """Which kernels this plan runs (mvtv_plan_describe) as a dict."""
        import json
        buf = C.create_string_buffer(1024)
        _lib.check(_lib.load().mvtv_plan_describe(self._h, buf, 1024))
        return json.loads(buf.value.decode())

    def profile(self, enable=True):
        _lib.check(_lib.load().mvtv_plan_profile(self._h, 1 if enable else 0))

    def get_profile(self):
        """{kernel class: (milliseconds, launches)} accumulated since profile(True)."""
        ms = np.zeros(_lib.KC_N)
        cnt = np.zeros(_lib.KC_N, dtype=np.int64)
        _lib.check(_lib.load().mvtv_plan_get_profile(self._h, _dp(ms), _ip(cnt)))
        return {name: (float(ms[k]), int(cnt[k])) for k, name in enumerate(_lib.KC_NAMES)}

    # -- operator-level entry points -----------------------------------------------------------
    def apply_D(self, theta):
        theta = _f64(np.asarray(theta).ravel())
        out = np.empty(self.R)
        _lib.check(_lib.load().mvtv_apply_D(self._h, _dp(theta), _dp(out)))
        return out

    def apply_Dt(self, rows):
        rows = _f64(np.asarray(rows).ravel())
        out = np.empty(self.N)
        _lib.check(_lib.load().mvtv_apply_Dt(self._h, _dp(rows), _dp(out)))
        return out

    def apply_M(self, s, x):
        x = _f64(np.asarray(x).ravel())
        out = np.empty(self.N)
        _lib.check(_lib.load().mvtv_apply_M(self._h, float(s), _dp(x), _dp(out)))
        return out


# ----------------------------------------------------------------------------------------------
# reference-shaped free functions
# ----------------------------------------------------------------------------------------------
def mbs_one(data, y, m, theta_init=None, mesh=None, tune=1.0, eps=0.01, tol=None, cache=None, mode="cpp",
            deltas=None, u=None, rho=None, dtype=F64, variant=VARIANT_REFERENCE, **solve_kw):
    """Mesh-based TV solution at one tuning parameter.

    Mirrors ``mbs_one`` of code/solvers.py:15-78 (dict output, ``tune`` = lambda, ``cache`` reuse) and of
    cpp-code/solvers.hpp:89 / rcpp solvers.hpp:104 (``theta_init``, ``lambda``, ``u``, ``rho``).
    ``mode`` picks which sibling's loop semantics to reproduce ("cpp", "rcpp", "py").  ``cache`` is a
    ``Plan`` with points already set (the reference's mbs_cache); with ``cache=None`` the operators are
    built here exactly like the stand-alone path (deltas empty -> unit block scales,
    cpp-code/solvers.cpp:141-145)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    y = np.asarray(y, dtype=np.float64).ravel()
    m = [int(v) for v in np.asarray(m).ravel()]
    own = cache is None
    if own:
        if mesh is None:
            axes = mesh_axes(data, m, mode)
            mesh = None
        elif isinstance(mesh, (list, tuple)):
            axes = [np.asarray(a, dtype=np.float64) for a in mesh]
        else:
            axes = axes_from_mesh(mesh, m)
        plan = Plan(m, deltas=deltas, dtype=dtype, variant=variant)
        plan.set_points(data, y, axes)
    else:
        plan = cache
        axes = plan.axes
    try:
        out = plan.solve(tune, mode=mode, theta_init=theta_init, u_init=u, rho_init=rho, tol=tol,
                         want_u=(_MODES[mode] == MODE_RCPP), **solve_kw)
    finally:
        if own:
            plan.close()
    return {"mesh": mesh if mesh is not None else mesh_from_axes(axes), "theta.hat": out["theta"],
            "fitted": out["fitted"], "data": data, "y": y, "eps": eps, "m": m, "counter": out["counter"],
            "passes": out["passes"], "u.hat": out["u"], "rho.hat": out["rho"], "axes": axes,
            "r_norm": out["r_norm"], "s_norm": out["s_norm"], "inner_iters": out["inner_iters"]}


def mbs_predict(mbs_one_object, data):
    """code/solvers.py:80-83 / cpp-code/solvers.cpp:154-158."""
    m = mbs_one_object["m"]
    axes = mbs_one_object.get("axes") or axes_from_mesh(mbs_one_object["mesh"], m)
    with Plan(m) as plan:
        return plan.predict(data, theta=mbs_one_object["theta.hat"], axes=axes)


def mbs_mse(mbs_one_object, y):
    """code/solvers.py:85-89 / cpp-code/solvers.cpp:160-168."""
    yhat = np.asarray(mbs_one_object["fitted"]).ravel()
    ytrue = np.asarray(y).ravel()
    return float(np.sum((yhat - ytrue) ** 2) / ytrue.size)


# ----------------------------------------------------------------------------------------------
# model-selection driver (host orchestration of the device path)
# ----------------------------------------------------------------------------------------------
def create_lambdas(n_lambda, lambda_max, mode="cpp"):
    """cpp-code/solvers.cpp:185 (1e-5*lambda_max .. lambda_max) ; rcpp solvers.cpp:191 (1e-4*lambda_max ..)."""
    lo = 0.00001 if _MODES[mode] == MODE_CPP else 0.0001
    return np.flipud(np.exp(_arma_linspace(math.log(lambda_max * lo), math.log(lambda_max), n_lambda)))


def kfoldinds(n, k, seed=117):
    """rcpp utils.cpp:367-376: indices i % k, shuffled (our own seeded shuffle: arma's RNG stream is not reproducible)."""
    idx = np.arange(n) % k
    np.random.RandomState(seed).shuffle(idx)
    return idx


def mbs(data, y, m, mesh=None, n_lambda=100, ftrue=None, lambdas=None, folds=5, verbose=False, mode="rcpp",
        foldinds=None, seed=117, dtype=F64, variant=VARIANT_REFERENCE, devices=None, **solve_kw):
    """Cross-validated fit over a path of lambdas: ``mbs`` / ``mbs_impl`` (cpp-code/solvers.cpp:277-310,
    rcpp-code/MultivarTV/src/solvers.cpp:305-376).  Follows the rcpp semantics (per-fold operators, fresh path per
    fold; SURVEY section 3 lists the cpp-code CV bugs).  ``mode`` selects the ADMM loop ("cpp" / "rcpp").
    ``devices``: CUDA ordinals for the cross-validation folds -- the folds are independent solves (cpp-code/solvers.cpp:300-304,
    rcpp solvers.cpp:341-352), so each runs as its own plan, one host thread per plan, round-robin over the listed GPUs
    (replica-level parallelism: no collective on the data path); None = one after the other on the current device.
    Returns the rcpp list as a dict plus the Python prototype's keys (code/solvers.py:140)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    y = np.asarray(y, dtype=np.float64).ravel()
    m = [int(v) for v in np.asarray(m).ravel()]
    n = y.size
    deltas = create_deltas(data, m, mode)                      # inits.deltas (cpp :281 / rcpp :310)
    axes = mesh_axes(data, m, mode) if mesh is None else (
        [np.asarray(a, dtype=np.float64) for a in mesh] if isinstance(mesh, (list, tuple)) else axes_from_mesh(mesh, m))
    FTRUE = y if ftrue is None else np.asarray(ftrue, dtype=np.float64).ravel()
    plan = Plan(m, deltas=deltas, dtype=dtype, variant=variant)
    try:
        plan.set_points(data, y, axes)                         # create_cache_objects on the full data
        if lambdas is None:
            lam_max, _ = plan.lambda_max(mode)
            if verbose:
                print("lambda_max = %f" % lam_max)
            LAMBDAS = create_lambdas(n_lambda, lam_max, mode)
        else:
            LAMBDAS = np.asarray(lambdas, dtype=np.float64).ravel()
        nl = LAMBDAS.size
        mse_mat = np.zeros((nl, max(1, folds)))
        if folds <= 1:
            final = plan.solve_path(LAMBDAS, FTRUE, mode=mode, want_thetas=True, **solve_kw)
            idx_all = plan.cache()[2]
            for i in range(nl):                                 # test_mse on the training data itself (rcpp :330)
                mse_mat[i, 0] = mbs_mse({"fitted": final["thetas"][i][idx_all]}, y)
            mean_mses = mse_mat[:, 0].copy()
            best1 = int(np.argmin(mean_mses))
            # mbs_fit_optimal: refit at the best lambda; the cached system matrix is whatever mbs_path left behind
            if _MODES[mode] == MODE_RCPP:   # rcpp solvers.cpp:261-274: cold start, rho = lambdas[0]/5, stale matrix scalar
                stale = final["rhos"][nl - 2] if nl >= 2 else LAMBDAS[0] / 5.0
                refit = plan.solve(LAMBDAS[best1], mode=mode, rho_init=LAMBDAS[0] / 5.0, rho_matrix0=stale, **solve_kw)
            else:                           # cpp solvers.cpp:248-260: warm start from the path's theta, matrix of the last lambda
                refit = plan.solve(LAMBDAS[best1], mode=mode, theta_init=final["thetas"][best1],
                                   rho_matrix0=LAMBDAS[-1], **solve_kw)
        else:
            if foldinds is None:
                foldinds = kfoldinds(n, folds, seed)
            foldinds = np.asarray(foldinds)

            def run_fold(f, pl):
                tr, te = foldinds != f, foldinds == f
                pl.set_points(data[tr], y[tr], axes)            # per-fold operators (rcpp :347-348)
                path = pl.solve_path(LAMBDAS, y[tr], mode=mode, want_thetas=True, want_best=False, **solve_kw)
                idx_te = nearest1(data[te], axes=axes)
                for i in range(nl):                             # test_mse (rcpp :278-288)
                    mse_mat[i, f] = float(np.sum((path["thetas"][i][idx_te] - y[te]) ** 2) / te.sum())
                if verbose:
                    print("Fold Complete: %d" % f)

            if devices:
                from concurrent.futures import ThreadPoolExecutor
                devs = [int(d) for d in devices]

                def worker(k):                                  # one plan per host thread, folds k, k + len(devs), ...
                    with Plan(m, deltas=deltas, dtype=dtype, variant=variant, device=devs[k]) as pl:
                        for f in range(k, folds, len(devs)):
                            run_fold(f, pl)
                with ThreadPoolExecutor(max_workers=len(devs)) as ex:
                    for fut in [ex.submit(worker, k) for k in range(min(len(devs), folds))]:
                        fut.result()
            else:
                for f in range(folds):
                    run_fold(f, plan)
            plan.set_points(data, y, axes)                      # final path on the full data (rcpp :355-358)
            final = plan.solve_path(LAMBDAS, y, mode=mode, want_thetas=True, **solve_kw)
            idx_all = plan.cache()[2]
            mean_mses = mse_mat.mean(axis=1)                    # rowmean (rcpp :359)
        best = int(np.argmin(mean_mses))                        # index_min: first minimum
        theta_hat = refit["theta"] if folds <= 1 else final["thetas"][best]
        fitted = theta_hat[idx_all]
        models = [{"lambda": float(LAMBDAS[i]), "mse": float(final["mses"][i]), "theta_hat": final["thetas"][i],
                   "fitted": final["thetas"][i][idx_all]} for i in range(nl)]
    finally:
        plan.close()
    out = {"data": data, "fitted": fitted, "m": m, "mesh": mesh_from_axes(axes), "theta_hat": theta_hat, "y": y,
           "residuals": y - fitted, "models": models, "lambda_minmse_ind": best + 1, "cv.mses": mean_mses,
           "lambdas": LAMBDAS, "cv.mse_mat": mse_mat, "axes": axes, "counters": final["counters"]}
    out["minmse.fits"] = {"mesh": out["mesh"], "theta.hat": theta_hat, "fitted": fitted, "data": data, "y": y, "m": m,
                          "axes": axes}
    out["minmse"] = float(mean_mses[best])
    out["minmse.lam"] = float(LAMBDAS[best])
    return out
