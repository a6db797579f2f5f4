"""Generate tests/golden/ref_py_golden.npz by EXECUTING the reference's own Python prototype
(/root/reference/code/utils.py, solvers.py) in this container.

The prototype is Python 2 and imports cvxopt (absent).  Nothing is copied into this repository:
the two source files are read where they lie, three py2->py3 compatibility substitutions are applied
IN MEMORY (none of them touches arithmetic), and the modules are exec'd:

  1. ``import cvxopt as cvxopt``              -> removed (never used by the functions we call)
  2. ``vals = range(dims.shape[0])``          -> ``list(range(...))``   (py2 range() was a list)
  3. ``bins.append( map(int,list(s)))``       -> ``list(map(...))``     (py2 map() returned a list)
  4. the module-level name ``csc_matrix`` is rebound to a wrapper that casts the (float-valued)
     row/column index lists to int64 -- scipy 0.x accepted float indices, scipy 1.18 does not.

Run:  python tests/golden/make_golden.py      (needs /root/reference; the GPU box never runs this)
"""
import os
import re
import sys
import types

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

REF = os.environ.get("MVTV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _csc_compat(arg, shape=None, **kw):
    if isinstance(arg, tuple) and len(arg) == 2 and isinstance(arg[1], tuple):
        val, (r, c) = arg
        r = np.asarray(r).astype(np.int64)
        c = np.asarray(c).astype(np.int64)
        arg = (np.asarray(val, dtype=float), (r, c))
    if shape is None:
        return sp.csc_matrix(arg, **kw)
    return sp.csc_matrix(arg, shape, **kw)


def load_reference():
    src = open(os.path.join(REF, "code", "utils.py")).read()
    src = src.replace("import cvxopt as cvxopt", "")
    src = re.sub(r"vals = range\(dims\.shape\[0\]\)", "vals = list(range(dims.shape[0]))", src)
    src = src.replace("bins.append( map(int,list(s)))", "bins.append( list(map(int,list(s))))")
    u = types.ModuleType("utils")
    exec(compile(src, "reference:code/utils.py", "exec"), u.__dict__)
    u.csc_matrix = _csc_compat
    sys.modules["utils"] = u
    s = types.ModuleType("solvers")
    exec(compile(open(os.path.join(REF, "code", "solvers.py")).read(), "reference:code/solvers.py", "exec"),
         s.__dict__)
    return u, s


class CountingLU:
    """Wraps the splu object handed to the reference's mbs_one as cache[0] so the number of
    passes is observable (the reference never increments its own counter, solvers.py:65-76)."""

    def __init__(self, lu):
        self.lu, self.calls = lu, 0

    def solve(self, b):
        self.calls += 1
        return self.lu.solve(b)


def synth(seed, n, p):
    rng = np.random.RandomState(seed)
    x = rng.uniform(-1, 1, (n, p))
    f = np.sin(3 * x[:, 0])
    for k in range(1, p):
        f = f + (k + 1) * (x[:, k] > 0.2)
    y = f + 0.3 * rng.normal(size=n)
    return x, y


def main():
    u, s = load_reference()
    out = {}

    # ---- index maps (code/utils.py:12-55) -------------------------------------------------
    for dims in ([3, 3, 3], [3, 2, 3], [4, 5], [2, 3, 4, 2]):
        d = np.array(dims)
        N = int(np.prod(d))
        v2t = np.array([u.v2t_unit(i, d) for i in range(N)])
        t2v = np.array([u.t2v_unit(v2t[i], d) for i in range(N)])
        key = "x".join(map(str, dims))
        out["v2t_" + key] = v2t
        out["t2v_" + key] = t2v

    # ---- masks (code/utils.py:63-69) ------------------------------------------------------
    for p in (1, 2, 3, 4):
        out["fd_binaries_%d" % p] = np.array(u.fd_binaries(p))

    # ---- D, dense (code/utils.py:85-149) --------------------------------------------------
    for dims in ([5], [3, 3], [4, 3], [3, 5], [3, 3, 3], [4, 4, 4], [3, 3, 4], [3, 3, 3, 3]):
        key = "x".join(map(str, dims))
        out["D_" + key] = u.create_D(np.array(dims), None).toarray()
    out["D_4x3_deltas"] = u.create_D(np.array([4, 3]), np.array([0.25, 0.5])).toarray()
    out["D_3x3x3_deltas"] = u.create_D(np.array([3, 3, 3]), np.array([0.25, 0.5, 2.0])).toarray()
    out["deltas_4x3"] = np.array([0.25, 0.5])
    out["deltas_3x3x3"] = np.array([0.25, 0.5, 2.0])
    for dims in ([3, 4, 5], [4, 3, 3]):
        try:
            u.create_D(np.array(dims), None)
            ok = 1
        except Exception:
            ok = 0
        out["D_noncubic_ok_" + "x".join(map(str, dims))] = np.array(ok)

    # ---- mesh_coords / nearest (code/utils.py:153-193) ------------------------------------
    r = u.mesh_coords(np.linspace(0.01, 0.99, 10).reshape(10, 1), mesh_dims=np.array([6]))
    out["mesh_coords_1d_mesh"] = np.asarray(r["mesh"])
    out["mesh_coords_1d_deltas"] = np.asarray(r["deltas"])
    x, y = synth(117, 400, 2)
    m = np.array([7, 5])
    mo = u.mesh_coords(x, m)
    out["near2_x"], out["near2_m"], out["near2_mesh"] = x, m, np.asarray(mo["mesh"])
    out["near2_idx"] = np.array(u.nearest1(x, mo["mesh"]))
    out["near2_deltas"] = np.asarray(mo["deltas"])

    # ---- the Python solver loop (code/solvers.py:15-78), cache supplied ---------------------
    def run_case(name, x, y, m, mesh, lam, rho_matrix, theta_init=None):
        n = y.size
        O = u.nearest_interp_matrix(x, mesh)
        D = u.create_D(m, None)
        Ot, Dt = O.transpose(), D.transpose()
        lu = CountingLU(splu((Ot.dot(O) + rho_matrix * Dt.dot(D)).tocsc()))
        cache = [lu, Ot.dot(y.reshape(n, 1)), D, Dt, D.shape[0], O, Ot, mesh, int(np.prod(m))]
        res = s.mbs_one(data=x, y=y, m=m, mesh=mesh, tune=lam, theta_init=theta_init, cache=cache)
        out[name + "_x"], out[name + "_y"], out[name + "_m"] = x, y, m
        out[name + "_mesh"] = np.asarray(mesh)
        out[name + "_lam"], out[name + "_rho_matrix"] = np.array(lam), np.array(rho_matrix)
        out[name + "_theta"] = np.asarray(res["theta.hat"]).ravel()
        out[name + "_fitted"] = np.asarray(res["fitted"]).ravel()
        out[name + "_passes"] = np.array(lu.calls)
        out[name + "_counter"] = np.array(res["counter"])
        out[name + "_idx"] = np.array(u.nearest1(x, mesh))
        return res

    x, y = synth(117, 300, 2)
    m = np.array([6, 5])
    mesh = u.mesh_coords(x, m)["mesh"]
    r1 = run_case("solve2a", x, y, m, mesh, 0.7, 0.7)
    run_case("solve2b", x, y, m, mesh, 0.3, 0.7, theta_init=r1["theta.hat"])   # warm start, stale matrix
    run_case("solve2c", x, y, m, mesh, 2.5, 2.5)

    x, y = synth(118, 250, 3)
    m = np.array([4, 4, 4])
    # meshgrid-based mesh_coords is only consistent with t2v for p<=2: build the p=3 mesh by hand
    axes = [np.linspace(x[:, k].min() - 0.01, x[:, k].max() + 0.01, m[k]) for k in range(3)]
    N = int(np.prod(m))
    mesh = np.array([[axes[k][u.v2t_unit(i, m)[k]] for k in range(3)] for i in range(N)])
    run_case("solve3a", x, y, m, mesh, 0.4, 0.4)

    x, y = synth(119, 60, 1)
    m = np.array([12])
    mesh = u.mesh_coords(x, m)["mesh"]
    run_case("solve1a", x, y, m, mesh, 0.5, 0.5)

    # ---- softthresh (code/solvers.py:9-12) --------------------------------------------------
    z = np.array([-2.0, -0.9, -0.3, 0.0, 0.3, 0.9, 2.0])
    out["soft_z"], out["soft_out_0p9"] = z, s.softthresh(z, 0.9)

    path = os.path.join(HERE, "ref_py_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "keys:", len(out))


if __name__ == "__main__":
    main()
