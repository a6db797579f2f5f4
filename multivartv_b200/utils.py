"""Host-side mirror of the reference's utils interface for the hot path (cpp-code/utils.hpp, code/utils.py).

Index maps and masks are integer bookkeeping and stay on the host; the operators ``create_D`` and
``nearest_interp_matrix`` return *device-backed* objects (D and O are never materialised: ``.dot`` runs
``mvtv_apply_D`` / ``mvtv_apply_Dt`` / ``mvtv_predict`` of libmvtv_b200.so), so the reference's own unit tests
(code/test_utils.py:10-64) can be written against this module unchanged.
"""
from __future__ import annotations

import numpy as np

from . import solvers as _s
from ._lib import VARIANT_REFERENCE


# ----------------------------------------------------------------------------------------------
# index maps -- cpp-code/utils.cpp:14-71 ; code/utils.py:12-59
# ----------------------------------------------------------------------------------------------
def prod(p, vec):
    """cpp-code/utils.cpp:14-22: product of the first p entries."""
    out = 1
    for i in range(int(p)):
        out *= int(vec[i])
    return out


def tensor2vector(p, multi_ind, dims):
    """cpp-code/utils.cpp:40-52: column-major flattening, axis 0 fastest: v = i0 + i1*m0 + i2*m0*m1 + ..."""
    v = int(multi_ind[0])
    for i in range(1, int(p)):
        v += int(multi_ind[i]) * prod(i, dims)
    return v


def vector2tensor(p, vec_ind, dims):
    """cpp-code/utils.cpp:54-71.  Upstream divides in ``float`` (exact only for N <= 2^24); integer arithmetic here
    gives the same multi-index wherever upstream's is exact, and the right one beyond."""
    out = [0] * int(p)
    ind2 = int(vec_ind) + 1
    for i in range(int(p), 0, -1):
        dp = prod(i - 1, dims)
        out[i - 1] = max(1, -(-ind2 // dp)) - 1
        ind2 -= out[i - 1] * dp
    return out


def t2v_unit(ind, dims):
    """code/utils.py:12-17."""
    return tensor2vector(len(dims), ind, dims)


def v2t_unit(ind, dims):
    """code/utils.py:32-41."""
    return np.asarray(vector2tensor(len(dims), ind, dims))


def t2v(dims):
    """code/utils.py:19-28: closure over the mesh dims."""
    dims = np.asarray(dims)
    return lambda ind: t2v_unit(ind, dims)


def v2t(dims):
    """code/utils.py:43-55."""
    dims = np.asarray(dims)
    return lambda ind: v2t_unit(ind, dims)


def dec2binary(n, p):
    """cpp-code/utils.cpp:73-89: p binary digits of n, most significant first."""
    return [(int(n) >> (int(p) - 1 - j)) & 1 for j in range(int(p))]


def fd_binaries(p):
    """cpp-code/utils.cpp:91-101 / code/utils.py:63-69: the 2^p - 1 axis masks, binaries of 1..2^p-1, MSB first."""
    return np.array([dec2binary(i, p) for i in range(1, 1 << int(p))], dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# D -- cpp-code/utils.cpp:245-269 (create_D) as an operator that lives on the device
# ----------------------------------------------------------------------------------------------
class _Transposed:
    def __init__(self, op):
        self._op = op
        self.shape = (op.shape[1], op.shape[0])

    @property
    def T(self):
        return self._op

    def dot(self, x):
        return self._op._rdot(x)

    __matmul__ = dot


class DifferenceOperator:
    """``create_D``: R x N stack of scaled difference blocks in the reference's row order (all-ones mask first).
    ``D.dot(theta)`` and ``D.T.dot(rows)`` run on the GPU (k_apply_D / k_apply_Dt); D is never stored."""

    def __init__(self, dims, deltas=None, variant=VARIANT_REFERENCE):
        self.dims = [int(v) for v in np.asarray(dims).ravel()]
        self._plan = _s.Plan(self.dims, deltas=deltas, variant=variant)
        self.shape = (self._plan.R, self._plan.N)

    @property
    def T(self):
        return _Transposed(self)

    def dot(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        out = self._plan.apply_D(theta.ravel())
        return out.reshape(-1, 1) if theta.ndim == 2 else out

    def _rdot(self, rows):
        rows = np.asarray(rows, dtype=np.float64)
        out = self._plan.apply_Dt(rows.ravel())
        return out.reshape(-1, 1) if rows.ndim == 2 else out

    __matmul__ = dot

    def close(self):
        self._plan.close()


def create_D(dims, deltas=None, variant=VARIANT_REFERENCE):
    """cpp-code/utils.cpp:245-269 (``create_D(p, dims, deltas)``; empty deltas = unit block scales)."""
    return DifferenceOperator(dims, deltas=deltas, variant=variant)


# ----------------------------------------------------------------------------------------------
# A = O -- cpp-code/utils.cpp:311-352 ; code/utils.py:153-177
# ----------------------------------------------------------------------------------------------
def _mesh_dims(mesh):
    """Mesh dims of a tensor-product mesh matrix (N x p, axis 0 fastest): the only meshes the reference builds."""
    mesh = np.asarray(mesh, dtype=np.float64)
    if mesh.ndim == 1:
        mesh = mesh[:, None]
    m = [int(np.unique(mesh[:, k]).size) for k in range(mesh.shape[1])]
    if int(np.prod(m)) != mesh.shape[0]:
        raise ValueError("mesh is not a tensor-product mesh: %d rows, per-axis distinct knots %s" % (mesh.shape[0], m))
    return mesh, m


def nearest1_unit(target, choices):
    """cpp-code/utils.cpp:311-321 / code/utils.py:153-157: index of the nearest mesh row (ties -> lowest index)."""
    mesh, m = _mesh_dims(choices)
    data = np.asarray(target, dtype=np.float64).reshape(1, -1)
    return int(_s.nearest1(data, mesh=mesh, m=m)[0])


def nearest1(data, mesh):
    """cpp-code/utils.cpp:323-330 / code/utils.py:159-166 (device kernel k_bin, O(n p log m) instead of O(n N))."""
    mesh, m = _mesh_dims(mesh)
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(-1, mesh.shape[1])
    return _s.nearest1(data, mesh=mesh, m=m)


class InterpolationOperator:
    """``nearest_interp_matrix``: n x N with a single 1.0 per row.  ``O.dot(theta)`` is a device gather
    (mvtv_predict); ``O.T.dot(y)`` the sorted segmented reduction of mvtv_plan_set_points (= Oty)."""

    def __init__(self, data, mesh):
        self._mesh, self.m = _mesh_dims(mesh)
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1:
            data = data.reshape(-1, self._mesh.shape[1])
        self._data = data
        self._axes = _s.axes_from_mesh(self._mesh, self.m)
        self._plan = _s.Plan(self.m)
        self.shape = (data.shape[0], self._plan.N)

    @property
    def T(self):
        return _Transposed(self)

    def dot(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        out = self._plan.predict(self._data, theta=theta.ravel(), axes=self._axes)
        return out.reshape(-1, 1) if theta.ndim == 2 else out

    def _rdot(self, y):
        y = np.asarray(y, dtype=np.float64)
        self._plan.set_points(self._data, y.ravel(), self._axes)
        out = self._plan.cache()[0]
        return out.reshape(-1, 1) if y.ndim == 2 else out

    __matmul__ = dot

    def close(self):
        self._plan.close()


def nearest_interp_matrix(data, mesh):
    """cpp-code/utils.cpp:332-352 / code/utils.py:168-177."""
    return InterpolationOperator(data, mesh)


def mesh_coords(data, mesh_dims):
    """code/utils.py:179-193: knots linspace(min-eps, max+eps, m_k) with eps = 0.01, deltas = knot spacing.
    The flattening follows the solver (axis 0 fastest, cpp-code/utils.cpp:40-52)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    axes = _s.mesh_axes(data, mesh_dims, "py")
    deltas = [float(np.diff(a)[0]) if len(a) > 1 else 0.0 for a in axes]
    return {"mesh": _s.mesh_from_axes(axes), "deltas": deltas, "axes": axes}
