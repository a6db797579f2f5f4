"""CPU ORACLE (test infrastructure, NOT product code) -- literal numpy/scipy restatement of
the MultivarTV mesh-based ADMM hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``multivartv_b200``) never does.

Every function cites the reference file:line it restates (paths relative to the upstream
repository root).  D is *materialised* as a scipy sparse matrix exactly the way the reference
materialises it (``build_diffmat`` -> ``mixedpartial`` -> ``create_D``) and the x-update uses
scipy's bundled SuperLU (``splu``), the same solver family as ``arma::spsolve``.

Pinning status (see DESIGN.md "Oracle" and oracle/README.md):
  * operators (index maps, masks, D incl. the mixedpartial direction-0 quirk, nearest, O, mesh)
    and the PY solver mode (code/solvers.py:54-76) are pinned against the reference's own Python
    prototype executed in this container (tests/golden/make_golden.py -> tests/golden/ref_py_golden.npz)
    and against the known answers of code/test_utils.py and cpp-code/utils_test.cpp;
  * the CPP and RCPP solver loops (cpp-code/solvers.cpp:90-130, rcpp solvers.cpp:96-136), create_D,
    create_mesh, nearest1, adapt_step, mbs_path and mbs_impl are pinned against the reference's own
    C++ sources compiled where they lie (oracle/_ref, built by oracle/ref_shim/Makefile against the
    Armadillo / Rcpp stand-ins of oracle/arma_shim): identical Counter, theta within 1e-10
    (tests/test_oracle_vs_reference.py, tests/test_golden_ref_cpp.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

MODE_CPP, MODE_RCPP, MODE_PY = 0, 1, 2
VARIANT_REFERENCE, VARIANT_INTENDED = 0, 1


# --------------------------------------------------------------------------------------
# index maps -- cpp-code/utils.cpp:16-71
# --------------------------------------------------------------------------------------
def prod(p, vec):
    """cpp-code/utils.cpp:16-22."""
    out = 1
    for i in range(p):
        out *= int(vec[i])
    return out


def tensor2vector(p, multi_ind, dims):
    """cpp-code/utils.cpp:40-52 -- column-major, axis 0 fastest."""
    vec_ind = int(multi_ind[0])
    for i in range(1, p):
        dims_prod = 1
        for j in range(i):
            dims_prod *= int(dims[j])
        vec_ind += int(multi_ind[i]) * dims_prod
    return vec_ind


def vector2tensor(p, vec_ind, dims):
    """cpp-code/utils.cpp:54-71 -- note the (float) division at :65 (exact only for N <= 2**24)."""
    multi_ind = [0] * p
    ind2 = vec_ind + 1
    for i in range(p, 0, -1):
        dims_prod = 1
        for j in range(i - 1):
            dims_prod *= int(dims[j])
        q = np.float32(ind2) / np.float32(dims_prod)
        multi_ind[i - 1] = max(1, int(math.ceil(float(q)))) - 1
        ind2 -= multi_ind[i - 1] * dims_prod
    return multi_ind


# --------------------------------------------------------------------------------------
# difference-operator masks -- cpp-code/utils.cpp:73-101
# --------------------------------------------------------------------------------------
def dec2binary(n, p):
    """cpp-code/utils.cpp:73-89 -- MSB first: axis p-1 is the least-significant bit."""
    out = [0] * p
    for k in range(p):
        out[p - 1 - k] = (n >> k) & 1
    return out


def fd_binaries(p):
    """cpp-code/utils.cpp:91-101 -- rows = binaries of 1..2^p-1."""
    return np.array([dec2binary(i + 1, p) for i in range((1 << p) - 1)], dtype=np.int64)


# --------------------------------------------------------------------------------------
# one-axis difference matrix -- cpp-code/utils.cpp:103-169
# --------------------------------------------------------------------------------------
def build_diffmat(p, dims, direction):
    """cpp-code/utils.cpp:103-169.  Row r (r-th vertex, in linear order, with
    ind[direction]+1 < dims[direction]) has +1 at ind and -1 at ind+e_direction."""
    dims = [int(d) for d in dims]
    n = prod(p, dims)
    lin = np.arange(n, dtype=np.int64)
    stride = 1
    for j in range(direction):
        stride *= dims[j]
    coord = (lin // stride) % dims[direction]
    keep = lin[coord + 1 < dims[direction]]
    nrow = keep.size
    rows = np.repeat(np.arange(nrow, dtype=np.int64), 2)
    cols = np.empty(2 * nrow, dtype=np.int64)
    cols[0::2] = keep
    cols[1::2] = keep + stride
    vals = np.tile(np.array([1.0, -1.0]), nrow)
    # sp_mat(locations, vals) sizes itself from the max indices (utils.cpp:167)
    ncol = int(cols.max()) + 1 if nrow else 0
    return sp.csc_matrix((vals, (rows, cols)), shape=(nrow, ncol))


def mixedpartial(p, dims, binary, variant=VARIANT_REFERENCE):
    """cpp-code/utils.cpp:171-221.  The first factor is ALWAYS built along direction 0
    (:187; the intended ``indices[j]`` is commented out at :186)."""
    indices = [i for i in range(p) if binary[i] == 1]
    armadims = [int(d) for d in dims]
    smats = []
    for j, ax in enumerate(indices):
        if j == 0:
            first_dir = 0 if variant == VARIANT_REFERENCE else ax
            smats.append(build_diffmat(p, armadims, first_dir))
        else:
            prev = indices[j - 1]
            armadims = list(armadims)
            armadims[prev] -= 1
            smats.append(build_diffmat(p, armadims, ax))
    if len(smats) == 1:
        return smats[0]
    D = smats[-1]
    for k in range(len(smats) - 2, -1, -1):
        if D.shape[1] != smats[k].shape[0]:
            raise ValueError(
                "matrix multiplication: incompatible matrix dimensions: %dx%d and %dx%d"
                % (D.shape + smats[k].shape))
        D = D @ smats[k]
    return D.tocsc()


def binary2diffmat(p, dims, binary, variant=VARIANT_REFERENCE):
    """cpp-code/utils.cpp:223-234."""
    if int(np.sum(binary)) == 1:
        direction = int(np.sum(np.arange(p) * np.asarray(binary)))
        return build_diffmat(p, dims, direction)
    return mixedpartial(p, dims, binary, variant)


def create_D(p, dims, deltas=None, variant=VARIANT_REFERENCE):
    """cpp-code/utils.cpp:245-269.  All-ones mask first (unscaled), then masks 1..K-1 in
    fd_binaries order, mask S scaled by prod_{k not in S} delta_k.  Empty ``deltas`` (the
    stand-alone mbs_one path, solvers.cpp:141-145) gives scale 1."""
    binaries = fd_binaries(p)
    n_mats = binaries.shape[0]
    blocks = [binary2diffmat(p, dims, binaries[n_mats - 1], variant)]
    for i in range(n_mats - 1):
        if deltas is None or len(deltas) == 0:
            delts = 1.0
        else:
            delts = 1.0
            for k in range(p):
                delts *= float(deltas[k]) ** float(1 - binaries[i, k])
        blocks.append(binary2diffmat(p, dims, binaries[i], variant) * delts)
    N = prod(p, dims)
    blocks = [_pad_cols(b, N) for b in blocks]
    return sp.vstack(blocks).tocsr()


def _pad_cols(b, ncol):
    b = b.tocoo()
    return sp.coo_matrix((b.data, (b.row, b.col)), shape=(b.shape[0], ncol)).tocsr()


def create_D_py(dims, deltas=None):
    """code/utils.py:138-149 -- the Python prototype's stacking: fd_binaries order (all-ones
    LAST); with ``deltas`` the all-ones block is dropped (:145)."""
    dims = np.asarray(dims)
    p = dims.shape[0]
    binaries = fd_binaries(p)
    N = prod(p, dims)
    blocks = []
    if deltas is None:
        for i in range(binaries.shape[0]):
            blocks.append(_pad_cols(binary2diffmat(p, dims, binaries[i]), N))
    else:
        for i in range(binaries.shape[0] - 1):
            delts = float(np.prod(np.asarray(deltas, dtype=float) ** (1 - binaries[i])))
            blocks.append(_pad_cols(binary2diffmat(p, dims, binaries[i]), N) * delts)
    return sp.vstack(blocks).tocsr()


# --------------------------------------------------------------------------------------
# mesh / deltas / nearest -- cpp-code/utils.cpp:271-352 ; rcpp utils.cpp:234-304
# --------------------------------------------------------------------------------------
def arma_linspace(start, end, num):
    """arma::linspace: x[i] = start + i*delta, x[N-1] = end."""
    num = int(num)
    if num == 1:
        return np.array([end], dtype=np.float64)
    delta = (end - start) / float(num - 1)
    x = start + np.arange(num, dtype=np.float64) * delta
    x[-1] = end
    return x


def mesh_axes(data, dims, mode=MODE_CPP):
    """Per-axis knot vectors of create_mesh.
    CPP : linspace(min+EPS, max+EPS, m_k), EPS=0.01, stored as float (cpp-code/utils.cpp:281,
          MAT=fmat utils.hpp:12) -> values are float32-rounded.
    RCPP: linspace(min-EPS, max+EPS, m_k), EPS=1e-4, double (rcpp utils.cpp:242).
    PY  : linspace(min-eps, max+eps, m_k), eps=0.01 (code/utils.py:179-186)."""
    data = np.asarray(data, dtype=np.float64)
    axes = []
    for k in range(data.shape[1]):
        lo, hi = float(data[:, k].min()), float(data[:, k].max())
        if mode == MODE_CPP:
            ax = arma_linspace(lo + 0.01, hi + 0.01, dims[k]).astype(np.float32).astype(np.float64)
        elif mode == MODE_RCPP:
            ax = arma_linspace(lo - 0.0001, hi + 0.0001, dims[k])
        else:
            ax = np.linspace(lo - 0.01, hi + 0.01, int(dims[k]))
        axes.append(ax)
    return axes


def create_mesh(data, dims, mode=MODE_CPP):
    """cpp-code/utils.cpp:271-298 -- N x p coordinate table, row i = knots at vector2tensor(i)."""
    axes = mesh_axes(data, dims, mode)
    p = len(axes)
    dims = [int(d) for d in dims]
    N = prod(p, dims)
    mesh = np.empty((N, p), dtype=np.float64)
    stride = 1
    lin = np.arange(N, dtype=np.int64)
    for k in range(p):
        mesh[:, k] = axes[k][(lin // stride) % dims[k]]
        stride *= dims[k]
    return mesh


def create_deltas(data, dims, mode=MODE_CPP):
    """cpp-code/utils.cpp:300-307 ; rcpp utils.cpp:256-263 (EPS differs)."""
    eps = 0.01 if mode == MODE_CPP else 0.0001
    data = np.asarray(data, dtype=np.float64)
    return np.array([(data[:, i].max() - data[:, i].min() + 2 * eps) / float(dims[i])
                     for i in range(data.shape[1])])


def nearest1_unit(target, choices):
    """cpp-code/utils.cpp:311-321 -- squared L2 to every mesh row, ties -> lowest index."""
    dists = np.sum((np.asarray(target, dtype=np.float64)[None, :] - choices) ** 2, axis=1)
    return int(np.flatnonzero(dists - dists.min() == 0)[0])


def nearest1(data, mesh):
    """cpp-code/utils.cpp:323-330 (O(n*N) brute force, literal)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    return np.array([nearest1_unit(data[i], mesh) for i in range(data.shape[0])], dtype=np.int64)


def nearest1_separable(data, axes, dims):
    """Same result as nearest1 on a tensor-product mesh, O(n p log m): per-axis nearest knot by
    the reference's own metric (x-knot)^2, ties -> lower index.  (Agreement with the literal
    brute force is asserted in tests/test_oracle_operators.py.)"""
    data = np.asarray(data, dtype=np.float64)
    n, p = data.shape
    idx = np.zeros(n, dtype=np.int64)
    stride = 1
    for k in range(p):
        ax = np.asarray(axes[k], dtype=np.float64)
        x = data[:, k]
        j = np.clip(np.searchsorted(ax, x), 1, len(ax) - 1) if len(ax) > 1 else np.zeros(n, dtype=np.int64)
        if len(ax) > 1:
            dl = (x - ax[j - 1]) ** 2
            dr = (x - ax[j]) ** 2
            jk = np.where(dl <= dr, j - 1, j)
        else:
            jk = j
        idx += jk * stride
        stride *= int(dims[k])
    return idx


def nearest_interp_matrix(data, mesh):
    """cpp-code/utils.cpp:332-352 -- n x N, one 1.0 per row."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    col = nearest1(data, mesh)
    n = data.shape[0]
    return sp.csc_matrix((np.ones(n), (np.arange(n), col)), shape=(n, mesh.shape[0]))


def interp_matrix_from_idx(idx, N):
    n = len(idx)
    return sp.csc_matrix((np.ones(n), (np.arange(n), idx)), shape=(n, N))


# --------------------------------------------------------------------------------------
# solver pieces -- cpp-code/solvers.cpp:15-130 ; rcpp solvers.cpp:20-136 ; code/solvers.py
# --------------------------------------------------------------------------------------
def softthresh(z, lam):
    """cpp-code/solvers.cpp:15-29 -- sign(z) % max(|z|-lam, 0); lam may be +inf."""
    return np.sign(z) * np.maximum(np.abs(z) - lam, 0.0)


def adapt_step_cpp(r, s, rho, u):
    """cpp-code/solvers.cpp:70-88."""
    r_norm = math.sqrt(float(np.dot(r, r)))
    s_norm = math.sqrt(float(np.dot(s, s)))
    if r_norm > 20 * s_norm:
        return 20 * rho, 0.05 * u
    if s_norm > 20 * r_norm:
        return 0.1 * rho, 10 * u
    return rho, u


def adapt_step_rcpp(r, s, rho, u):
    """rcpp-code/MultivarTV/src/solvers.cpp:77-94."""
    r_norm = float(np.linalg.norm(r))
    s_norm = float(np.linalg.norm(s))
    tau = 2.0
    if r_norm > 10 * s_norm:
        return tau * rho, 1.0 / tau * u
    if s_norm > 10 * r_norm:
        return 1.0 / tau * rho, tau * u
    return rho, u


@dataclass
class Inits:
    """mbs_one_inits (cpp-code/solvers.hpp:36-43)."""
    O: sp.spmatrix
    D: sp.spmatrix
    Oty: np.ndarray
    ntheta: int
    rowsD: int
    crossD: sp.spmatrix
    crossO: sp.spmatrix
    sp_crosses: sp.spmatrix | None = None
    Dt: sp.spmatrix = field(init=False)
    Ot: sp.spmatrix = field(init=False)

    def __post_init__(self):
        self.Dt = self.D.T.tocsr()
        self.Ot = self.O.T.tocsr()


def create_cache_objects(O, D, y):
    """cpp-code/solvers.cpp:31-41."""
    O = O.tocsr()
    D = D.tocsr()
    return Inits(O=O, D=D, Oty=np.asarray(O.T @ y).ravel(), ntheta=D.shape[1], rowsD=D.shape[0],
                 crossD=(D.T @ D).tocsc(), crossO=(O.T @ O).tocsc())


class _LU:
    """spsolve with a per-matrix-scalar factor cache (the reference re-factorises every
    iteration; the numbers are the same, this only saves oracle time)."""

    def __init__(self, crossO, crossD):
        self.crossO, self.crossD, self.cache = crossO, crossD, {}

    def solve(self, scal, b):
        lu = self.cache.get(scal)
        if lu is None:
            if len(self.cache) > 8:
                self.cache.clear()
            lu = splu((self.crossO + scal * self.crossD).tocsc())
            self.cache[scal] = lu
        return lu.solve(b)


def admm_update_cpp(y, inits, theta_init, lam, tol=0.001, max_counter=2000, trace=None):
    """cpp-code/solvers.cpp:90-130, literal -- including ``int rho = lambda`` (:108), the
    truncation at :126, the fixed system matrix crossO + lambda*crossD (:144/:209), the dual
    residual built from the *pre-update* u (:118) and the throw at :122-124."""
    D, Dt = inits.D, inits.Dt
    meany = float(np.mean(y))
    theta = np.full(inits.ntheta, meany) if theta_init is None else np.array(theta_init, dtype=np.float64)
    alpha = D @ theta
    u = np.full(inits.rowsD, 1.0 / lam)
    thetaold = np.full(inits.ntheta, meany - 0.1)
    counter = 1
    rho = int(lam)                                           # :108
    lu = _LU(inits.crossO, inits.crossD)
    r_norm = s_norm = float("nan")
    with np.errstate(divide="ignore", invalid="ignore"):
        while np.any(np.abs(theta - thetaold) > tol):        # :113
            thetaold = theta
            b = inits.Oty + rho * (Dt @ (alpha + u))          # :115
            theta = lu.solve(lam, b)                          # :116 (sp_crosses fixed at lambda)
            kappa = (lam / rho) if rho != 0 else math.inf     # :117 double/int -> +inf at rho==0
            Dtheta = D @ theta
            alpha = softthresh(Dtheta - u, kappa)
            dual_residual = rho * (Dt @ (alpha + u))          # :118 (u before update)
            primal_residual = alpha - Dtheta                  # :119
            u = u + primal_residual                           # :120
            counter += 1
            if counter > max_counter:
                raise ValueError("Failed to converge!")      # :122-124
            r_norm = math.sqrt(float(np.dot(primal_residual, primal_residual)))
            s_norm = math.sqrt(float(np.dot(dual_residual, dual_residual)))
            rho_next, u = adapt_step_cpp(primal_residual, dual_residual, rho, u)
            rho = int(rho_next)                               # :126 truncation
            if trace is not None:
                trace.append(dict(theta=theta.copy(), u=u.copy(), rho=rho, r=r_norm, s=s_norm))
    return dict(theta=theta, u=u, rho=float(rho), counter=counter, r_norm=r_norm, s_norm=s_norm)


def admm_update_rcpp(y, inits, theta_init, lam, u_init, rho_init, rho_matrix0=None, tol=0.0001,
                     max_counter=3000, trace=None):
    """rcpp-code/MultivarTV/src/solvers.cpp:96-136, literal.  ``rho_matrix0`` is the scalar of
    the cached sp_crosses used for the FIRST pass (mbs_path :213 -> rho_init; stand-alone
    mbs_one :150 -> lambda)."""
    D, Dt = inits.D, inits.Dt
    theta = np.array(theta_init, dtype=np.float64)
    u = np.array(u_init, dtype=np.float64)
    rho = float(rho_init)
    rho_mat = float(rho_init if rho_matrix0 is None else rho_matrix0)
    alpha = D @ theta
    counter = 1
    dual_norm = primal_norm = 1.0
    eps_dual = eps_primal = tol
    lu = _LU(inits.crossO, inits.crossD)
    while dual_norm > eps_dual or primal_norm > eps_primal:  # :110
        uold = u
        b = inits.Oty + rho * (Dt @ (alpha + u))              # :112
        theta = lu.solve(rho_mat, b)                          # :113
        Dtheta = D @ theta
        alpha = softthresh(Dtheta - u, lam / rho)             # :114
        primal_residual = alpha - Dtheta                      # :115
        u = u + primal_residual                               # :116
        dual_residual = rho * (Dt @ (u - uold))               # :117
        dual_norm = float(np.linalg.norm(dual_residual))
        primal_norm = float(np.linalg.norm(primal_residual))
        eps_dual = tol * (math.sqrt(inits.ntheta) + float(np.linalg.norm(Dt @ u)))            # :121
        eps_primal = tol * (math.sqrt(inits.rowsD)
                            + max(float(np.linalg.norm(Dtheta)), float(np.linalg.norm(alpha))))  # :122
        rho, u = adapt_step_rcpp(primal_residual, dual_residual, rho, u)  # :124-125
        rho_mat = rho                                         # :126
        counter += 1
        if trace is not None:
            trace.append(dict(theta=theta.copy(), u=u.copy(), rho=rho, r=primal_norm, s=dual_norm))
        if counter > max_counter:                             # :129-132 message + break
            break
    return dict(theta=theta, u=u, rho=rho, counter=counter, r_norm=primal_norm, s_norm=dual_norm)


def admm_update_py(y, inits, theta_init, lam, rho_matrix=None, tol=0.001, maxc=5000, trace=None):
    """code/solvers.py:54-76 -- fixed double rho = tune, factor once, u = 1/lambda,
    thetaold = mean(y)-1, stop on any(|dtheta|>tol).  The reference never increments
    ``counter`` (:65-76); we count passes separately and return the reference's value (1)
    as ``counter`` and the real number of passes as ``passes``."""
    D, Dt = inits.D, inits.Dt
    meany = float(np.mean(y))
    rho = float(lam)
    theta = np.full(inits.ntheta, meany) if theta_init is None else np.array(theta_init, dtype=np.float64)
    alpha = D @ theta
    u = np.full(inits.rowsD, 1.0 / lam)
    thetaold = np.full(inits.ntheta, meany - 1.0)
    lu = _LU(inits.crossO, inits.crossD)
    scal = float(lam if rho_matrix is None else rho_matrix)
    passes = 0
    while np.any(np.abs(theta - thetaold) > tol):
        thetaold = theta
        b = inits.Oty + rho * (Dt @ (alpha + u))
        theta = lu.solve(scal, b)
        Dtheta = D @ theta
        alpha = softthresh(Dtheta - u, lam / rho)
        u = u + alpha - Dtheta
        passes += 1
        if trace is not None:
            trace.append(dict(theta=theta.copy(), u=u.copy(), rho=rho))
        if passes > maxc:
            raise RuntimeError("Solver did not converge")
    return dict(theta=theta, u=u, rho=rho, counter=1, passes=passes)


def mbs_one(data, y, m, mesh_axes_=None, theta_init=None, lam=1.0, mode=MODE_CPP, deltas=None,
            u_init=None, rho_init=None, rho_matrix0=None, variant=VARIANT_REFERENCE,
            brute_force_nearest=False, tol=None, max_counter=None):
    """mbs_one with cache==NULL (cpp-code/solvers.cpp:134-152 ; rcpp solvers.cpp:140-159):
    build O, D (``deltas`` empty in the stand-alone path), sp_crosses = crossO + lambda*crossD,
    run admm_update, fitted = O*theta."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    y = np.asarray(y, dtype=np.float64).ravel()
    p = data.shape[1]
    dims = [int(v) for v in m]
    axes = mesh_axes(data, dims, mode) if mesh_axes_ is None else mesh_axes_
    N = prod(p, dims)
    if brute_force_nearest:
        stride, lin = 1, np.arange(N)
        mesh = np.empty((N, p))
        for k in range(p):
            mesh[:, k] = np.asarray(axes[k])[(lin // stride) % dims[k]]
            stride *= dims[k]
        idx = nearest1(data, mesh)
    else:
        idx = nearest1_separable(data, axes, dims)
    O = interp_matrix_from_idx(idx, N)
    D = create_D(p, dims, deltas, variant)
    inits = create_cache_objects(O, D, y)
    kw = {}
    if tol is not None:
        kw["tol"] = tol
    if mode == MODE_CPP:
        if max_counter is not None:
            kw["max_counter"] = max_counter
        out = admm_update_cpp(y, inits, theta_init, lam, **kw)
    elif mode == MODE_RCPP:
        if max_counter is not None:
            kw["max_counter"] = max_counter
        th0 = np.full(N, float(np.mean(y))) if theta_init is None else theta_init
        u0 = np.zeros(D.shape[0]) if u_init is None else u_init
        rho0 = lam / 5.0 if rho_init is None else rho_init
        out = admm_update_rcpp(y, inits, th0, lam, u0, rho0,
                               rho_matrix0=(lam if rho_matrix0 is None else rho_matrix0), **kw)
    else:
        out = admm_update_py(y, inits, theta_init, lam, rho_matrix=rho_matrix0, **kw)
    out["fitted"] = np.asarray(O @ out["theta"]).ravel()
    out["idx"] = idx
    out["Oty"] = inits.Oty
    out["counts"] = np.asarray(inits.crossO.diagonal()).ravel()
    out["D"] = D
    return out


def mbs_predict(theta, axes, dims, data):
    """cpp-code/solvers.cpp:154-158."""
    idx = nearest1_separable(np.asarray(data, dtype=np.float64), axes, dims)
    return np.asarray(theta)[idx]


def mse(fits, y):
    """cpp-code/solvers.cpp:160-163."""
    fits, y = np.asarray(fits).ravel(), np.asarray(y).ravel()
    return float(np.sum((fits - y) ** 2) / y.size)


# --------------------------------------------------------------------------------------
# lambda_max and the lambda grid -- cpp-code/utils.cpp:354-404, solvers.cpp:179-192 ; rcpp utils.cpp:306-355
# --------------------------------------------------------------------------------------
def cg_cpp(A, b):
    """cpp-code/utils.cpp:354-386 -- truncated CG from x0 = mean(b), abs tol 0.01, MAXIT 100 (500 if n < 400)."""
    b = np.asarray(b, dtype=np.float64)
    x = np.full(b.shape[0], float(np.mean(b)))
    r = b - A @ x
    p = r.copy()
    rsold = float(r @ r)
    rsnew = rsold + 1.0
    it = 0
    MAXIT = 500 if b.shape[0] < 400 else 100
    while math.sqrt(rsnew) >= 0.01:
        Ap = A @ p
        alpha = rsold / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rsnew = float(r @ r)
        it += 1
        if it == MAXIT:
            break
        p = r + (rsnew / rsold) * p
        rsold = rsnew
    return x, it


def cg_rcpp(A, b):
    """rcpp-code/MultivarTV/src/utils.cpp:306-340 -- CGNR, relative tolerance 1e-4, MAXIT min(n, 2000)."""
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros(b.shape[0])
    d = b - A @ x
    r = A.T @ d
    p = r.copy()
    rsold0 = float(np.linalg.norm(r))
    rsold = rsold0 ** 2
    rsnew = rsold + 1.0
    t = A @ p
    it = 0
    MAXIT = b.shape[0] if b.shape[0] < 2000 else 2000
    while math.sqrt(rsnew) >= 0.0001 * rsold0:
        alpha = rsold / float(np.linalg.norm(t)) ** 2
        x = x + alpha * p
        d = d - alpha * t
        r = A.T @ d
        rsnew = float(np.linalg.norm(r)) ** 2
        it += 1
        if it == MAXIT:
            break
        p = r + (rsnew / rsold) * p
        t = A @ p
        rsold = rsnew
    return x, it


def lam_max_pinv(D, Oty, mode=MODE_CPP):
    """cpp-code/utils.cpp:389-404 (max|D b|) ; rcpp utils.cpp:343-355 (5 * ||D b||_inf)."""
    ata = (D.T @ D).tocsr()
    if mode == MODE_RCPP:
        b, it = cg_rcpp(ata, Oty)
        return 5.0 * float(np.max(np.abs(D @ b))), it
    b, it = cg_cpp(ata, Oty)
    return float(np.max(np.abs(D @ b))), it


def create_lambdas(n_lambda, lambda_max, mode=MODE_CPP):
    """cpp-code/solvers.cpp:185 ; rcpp solvers.cpp:191."""
    lo = 0.00001 if mode == MODE_CPP else 0.0001
    return np.flipud(np.exp(arma_linspace(math.log(lambda_max * lo), math.log(lambda_max), n_lambda)))
