"""ctypes binding of oracle/_ref/libmvtv_ref.so and libmvtv_ref_rcpp.so: the REFERENCE's own compiled C++ solvers
(cpp-code/{utils,solvers}.cpp and rcpp-code/MultivarTV/src/{utils,solvers}.cpp, compiled where they lie under
/root/reference against the stand-ins oracle/arma_shim/{armadillo,RcppArmadillo.h}; recipe oracle/ref_shim/Makefile).
TEST INFRASTRUCTURE ONLY.  Module-level functions call the cpp-code library, the ``rcpp`` namespace the Rcpp-side one.

`available()` is False where neither the prebuilt library nor /root/reference exists (e.g. a fresh clone): callers skip.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libmvtv_ref.so")
LIB_RCPP = os.path.join(HERE, "_ref", "libmvtv_ref_rcpp.so")
REFERENCE = os.environ.get("MVTV_REFERENCE_DIR", "/root/reference")
_lib = None


def build(force=False):
    """make -C oracle/ref_shim (needs the reference sources; never copies them)."""
    if not os.path.isdir(os.path.join(REFERENCE, "cpp-code")):
        return os.path.exists(LIB) and os.path.exists(LIB_RCPP)
    cmd = ["make", "-C", os.path.join(HERE, "ref_shim"), "REF=" + REFERENCE]
    if force:
        subprocess.check_call(cmd + ["clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(cmd + ["all"], stdout=subprocess.DEVNULL)
    return os.path.exists(LIB) and os.path.exists(LIB_RCPP)


def available():
    return (os.path.exists(LIB) and os.path.exists(LIB_RCPP)) or build()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref/libmvtv_ref.so is missing and %s is not present" % REFERENCE)
        L = C.CDLL(LIB)
        L.ref_last_error.restype = C.c_char_p
        ll, i32, dbl = C.c_longlong, C.c_int, C.c_double
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)
        L.ref_tensor2vector.argtypes = [i32, ip, ip]
        L.ref_vector2tensor.argtypes = [i32, i32, ip, ip]
        L.ref_fd_binaries.argtypes = [i32, ip]
        L.ref_create_D_rows.argtypes = [i32, dp, dp, lp, lp]
        L.ref_create_D_dense.argtypes = [i32, dp, dp, dp]
        L.ref_create_mesh.argtypes = [ll, i32, dp, dp, dp]
        L.ref_create_deltas.argtypes = [ll, i32, dp, dp, dp]
        L.ref_nearest1.argtypes = [ll, i32, dp, ll, dp, lp]
        L.ref_softthresh.argtypes = [ll, dp, dbl, dp]
        L.ref_adapt_step.argtypes = [ll, dp, ll, dp, dbl, ll, dp, dp, dp]
        L.ref_mbs_one.argtypes = [ll, i32, dp, dp, dp, dp, dp, dbl, dp, dp, ip]
        L.ref_mbs_path.argtypes = [ll, i32, dp, dp, dp, i32, dp, dp, dp, dp, dp, ip, dp]
        L.ref_lambda_max.argtypes = [ll, i32, dp, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(data):
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    return np.ascontiguousarray(data.T).ravel(), data.shape[0], data.shape[1]


class RefError(RuntimeError):
    def __init__(self, code):
        super().__init__("reference raised (%d): %s" % (code, lib().ref_last_error().decode()))
        self.code = code


def _check(rc):
    if rc != 0:
        raise RefError(rc)


def tensor2vector(p, multi_ind, dims):
    return int(lib().ref_tensor2vector(int(p), _ip(np.ascontiguousarray(multi_ind, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(dims, dtype=np.int32))))


def vector2tensor(p, vec_ind, dims):
    out = np.zeros(int(p), dtype=np.int32)
    lib().ref_vector2tensor(int(p), int(vec_ind), _ip(np.ascontiguousarray(dims, dtype=np.int32)), _ip(out))
    return [int(v) for v in out]


def fd_binaries(p):
    out = np.zeros(((1 << p) - 1, p), dtype=np.int32)
    lib().ref_fd_binaries(int(p), _ip(out))
    return out.astype(np.int64)


def create_D(dims, deltas=None):
    """Dense copy of the reference's create_D(p, dims, deltas) (cpp-code/utils.cpp:245-269)."""
    d = _f(dims)
    dl = None if deltas is None else _f(deltas)
    r, c = C.c_longlong(0), C.c_longlong(0)
    _check(lib().ref_create_D_rows(len(d), _dp(d), _dp(dl), C.byref(r), C.byref(c)))
    out = np.zeros((r.value, c.value))
    _check(lib().ref_create_D_dense(len(d), _dp(d), _dp(dl), _dp(out)))
    return out


def create_mesh(data, dims):
    dcm, n, p = _colmajor(data)
    d = _f(dims)
    N = int(np.prod(d))
    out = np.zeros(N * p)
    _check(lib().ref_create_mesh(n, p, _dp(dcm), _dp(d), _dp(out)))
    return out.reshape(p, N).T.copy()


def create_deltas(data, dims):
    dcm, n, p = _colmajor(data)
    out = np.zeros(p)
    _check(lib().ref_create_deltas(n, p, _dp(dcm), _dp(_f(dims)), _dp(out)))
    return out


def nearest1(data, mesh):
    dcm, n, p = _colmajor(data)
    mcm, N, p2 = _colmajor(mesh)
    assert p == p2
    out = np.zeros(n, dtype=np.int64)
    _check(lib().ref_nearest1(n, p, _dp(dcm), N, _dp(mcm), out.ctypes.data_as(C.POINTER(C.c_longlong))))
    return out


def softthresh(z, lam):
    z = _f(z).ravel()
    out = np.empty_like(z)
    _check(lib().ref_softthresh(z.size, _dp(z), C.c_double(lam), _dp(out)))
    return out


def adapt_step(r, s, rho, u):
    r, s, u = _f(r).ravel(), _f(s).ravel(), _f(u).ravel()
    out = np.empty_like(u)
    rho_next = C.c_double(0.0)
    _check(lib().ref_adapt_step(r.size, _dp(r), s.size, _dp(s), C.c_double(rho), u.size, _dp(u), C.byref(rho_next), _dp(out)))
    return rho_next.value, out


def mbs_one(data, y, m, lam, mesh=None, theta_init=None):
    """The reference's stand-alone mbs_one(data, y, m, out, mesh, theta_init, lambda, NULL)."""
    dcm, n, p = _colmajor(data)
    y = _f(y).ravel()
    md = _f(m)
    N = int(np.prod(md))
    mcm = None if mesh is None else _colmajor(mesh)[0]
    th0 = None if theta_init is None else _f(theta_init).ravel()
    theta, fitted = np.zeros(N), np.zeros(n)
    counter = C.c_int(0)
    _check(lib().ref_mbs_one(n, p, _dp(dcm), _dp(y), _dp(md), _dp(mcm), _dp(th0), C.c_double(lam), _dp(theta), _dp(fitted),
                             C.byref(counter)))
    return {"theta": theta, "fitted": fitted, "counter": counter.value}


def mbs_path(data, y, m, n_lambda=None, lambdas=None, ftrue=None):
    """Operator set-up of the reference's mbs() + create_lambdas + mbs_path on the full data."""
    dcm, n, p = _colmajor(data)
    y = _f(y).ravel()
    md = _f(m)
    N = int(np.prod(md))
    lam_in = None if lambdas is None else _f(lambdas).ravel()
    nl = int(n_lambda if lam_in is None else lam_in.size)
    ft = None if ftrue is None else _f(ftrue).ravel()
    lam_out, thetas, mses = np.zeros(nl), np.zeros((nl, N)), np.zeros(nl)
    counters = np.zeros(nl, dtype=np.int32)
    lmax = C.c_double(0.0)
    _check(lib().ref_mbs_path(n, p, _dp(dcm), _dp(y), _dp(md), nl, _dp(lam_in), _dp(ft), _dp(lam_out), _dp(thetas), _dp(mses),
                              _ip(counters), C.byref(lmax)))
    return {"lambdas": lam_out, "thetas": thetas, "mses": mses, "counters": counters, "lambda_max": lmax.value}


def lambda_max(data, y, m):
    dcm, n, p = _colmajor(data)
    out = C.c_double(0.0)
    _check(lib().ref_lambda_max(n, p, _dp(dcm), _dp(_f(y).ravel()), _dp(_f(m)), C.byref(out)))
    return out.value


class _Rcpp:
    """The Rcpp-side sibling (rcpp-code/MultivarTV/src): double rho, residual stopping rule, EPS = 1e-4 mesh."""
    _lib = None

    def lib(self):
        if self._lib is None:
            if not available():
                raise ImportError("oracle/_ref/libmvtv_ref_rcpp.so is missing and %s is not present" % REFERENCE)
            L = C.CDLL(LIB_RCPP)
            L.rref_last_error.restype = C.c_char_p
            ll, i32, dbl = C.c_longlong, C.c_int, C.c_double
            dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)
            L.rref_create_mesh.argtypes = [ll, i32, dp, dp, dp]
            L.rref_create_deltas.argtypes = [ll, i32, dp, dp, dp]
            L.rref_adapt_step.argtypes = [ll, dp, ll, dp, dbl, ll, dp, dp, dp]
            L.rref_rows_of_D.argtypes = [i32, dp, lp]
            L.rref_mbs_one.argtypes = [ll, i32, dp, dp, dp, dp, dp, dp, dbl, dp, dp, ip]
            L.rref_mbs_path.argtypes = [ll, i32, dp, dp, dp, i32, dp, dp, dp, dp, dp, ip, dp]
            L.rref_lambda_max.argtypes = [ll, i32, dp, dp, dp, dp]
            L.rref_mbs_impl_folds1.argtypes = [ll, i32, dp, dp, dp, i32, dp, dp, dp, dp, ip]
            self._lib = L
        return self._lib

    def _check(self, rc):
        if rc != 0:
            e = RuntimeError("reference (rcpp) raised (%d): %s" % (rc, self.lib().rref_last_error().decode()))
            e.code = rc
            raise e

    def create_mesh(self, data, dims):
        dcm, n, p = _colmajor(data)
        d = _f(dims)
        N = int(np.prod(d))
        out = np.zeros(N * p)
        self._check(self.lib().rref_create_mesh(n, p, _dp(dcm), _dp(d), _dp(out)))
        return out.reshape(p, N).T.copy()

    def create_deltas(self, data, dims):
        dcm, n, p = _colmajor(data)
        out = np.zeros(p)
        self._check(self.lib().rref_create_deltas(n, p, _dp(dcm), _dp(_f(dims)), _dp(out)))
        return out

    def adapt_step(self, r, s, rho, u):
        r, s, u = _f(r).ravel(), _f(s).ravel(), _f(u).ravel()
        out = np.empty_like(u)
        rho_next = C.c_double(0.0)
        self._check(self.lib().rref_adapt_step(r.size, _dp(r), s.size, _dp(s), C.c_double(rho), u.size, _dp(u), C.byref(rho_next),
                                               _dp(out)))
        return rho_next.value, out

    def rows_of_D(self, dims):
        d = _f(dims)
        r = C.c_longlong(0)
        self._check(self.lib().rref_rows_of_D(len(d), _dp(d), C.byref(r)))
        return r.value

    def mbs_one(self, data, y, m, lam, theta_init=None, u_init=None, rho_init=None):
        """Stand-alone mbs_one(data, y, m, out, mesh, u, rho, theta_init, lambda, NULL, verbose) of the Rcpp side; defaults
        as mbs_path starts a path: theta = mean(y), u = 0, rho = lambda/5."""
        dcm, n, p = _colmajor(data)
        y = _f(y).ravel()
        md = _f(m)
        N, R = int(np.prod(md)), self.rows_of_D(md)
        th0 = np.full(N, y.mean()) if theta_init is None else _f(theta_init).ravel().copy()
        u = np.zeros(R) if u_init is None else _f(u_init).ravel().copy()
        rho = C.c_double(lam / 5.0 if rho_init is None else rho_init)
        theta, fitted = np.zeros(N), np.zeros(n)
        counter = C.c_int(0)
        self._check(self.lib().rref_mbs_one(n, p, _dp(dcm), _dp(y), _dp(md), _dp(th0), _dp(u), C.byref(rho), C.c_double(lam),
                                            _dp(theta), _dp(fitted), C.byref(counter)))
        return {"theta": theta, "fitted": fitted, "u": u, "rho": rho.value, "counter": counter.value}

    def mbs_path(self, data, y, m, n_lambda=None, lambdas=None):
        dcm, n, p = _colmajor(data)
        y = _f(y).ravel()
        md = _f(m)
        N = int(np.prod(md))
        lam_in = None if lambdas is None else _f(lambdas).ravel()
        nl = int(n_lambda if lam_in is None else lam_in.size)
        lam_out, thetas, mses, rhos = np.zeros(nl), np.zeros((nl, N)), np.zeros(nl), np.zeros(nl)
        counters = np.zeros(nl, dtype=np.int32)
        lmax = C.c_double(0.0)
        self._check(self.lib().rref_mbs_path(n, p, _dp(dcm), _dp(y), _dp(md), nl, _dp(lam_in), _dp(lam_out), _dp(thetas), _dp(mses),
                                             _dp(rhos), _ip(counters), C.byref(lmax)))
        return {"lambdas": lam_out, "thetas": thetas, "mses": mses, "rhos": rhos, "counters": counters, "lambda_max": lmax.value}

    def lambda_max(self, data, y, m):
        dcm, n, p = _colmajor(data)
        out = C.c_double(0.0)
        self._check(self.lib().rref_lambda_max(n, p, _dp(dcm), _dp(_f(y).ravel()), _dp(_f(m)), C.byref(out)))
        return out.value

    def mbs_impl_folds1(self, data, y, m, lambdas):
        dcm, n, p = _colmajor(data)
        y = _f(y).ravel()
        md = _f(m)
        lam = _f(lambdas).ravel()
        theta, fitted, cv = np.zeros(int(np.prod(md))), np.zeros(n), np.zeros(lam.size)
        best = C.c_int(0)
        self._check(self.lib().rref_mbs_impl_folds1(n, p, _dp(dcm), _dp(y), _dp(md), lam.size, _dp(lam), _dp(theta), _dp(fitted),
                                                    _dp(cv), C.byref(best)))
        return {"theta_hat": theta, "fitted": fitted, "cv.mses": cv, "lambda_minmse_ind": best.value}


rcpp = _Rcpp()
