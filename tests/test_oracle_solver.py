"""Oracle solver loops: pinned against the reference Python prototype's golden solves (PY mode),
the reference's lambda_max property test (code/test_solvers.py:24-29), and cross-checked between
the two independent restatements (scipy/SuperLU literal vs matrix-free C) in CPP and RCPP modes.
CPU-only."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po
from tests.helpers import synth


def _axes_from_mesh(mesh, m):
    axes, stride = [], 1
    for k in range(len(m)):
        axes.append(np.ascontiguousarray(mesh[::stride, k][: m[k]]))
        stride *= int(m[k])
    return axes


CASES = ["solve2a", "solve2b", "solve2c", "solve3a", "solve1a"]


@pytest.mark.parametrize("name", CASES)
def test_py_mode_golden(golden, name):
    """code/solvers.py:54-76 executed by the reference itself vs both oracles: identical number of
    passes, theta within 1e-11 (LU vs banded Cholesky rounding), fitted likewise."""
    x, y, m, mesh = golden[name + "_x"], golden[name + "_y"], golden[name + "_m"], golden[name + "_mesh"]
    lam, rm = float(golden[name + "_lam"]), float(golden[name + "_rho_matrix"])
    th0 = golden["solve2a_theta"] if name == "solve2b" else None
    axes = _axes_from_mesh(mesh, m)
    # python oracle, python stacking order (the golden run used deltas=None -> all blocks, unscaled)
    idx = po.nearest1(x, mesh)
    assert np.array_equal(idx, golden[name + "_idx"])
    inits = po.create_cache_objects(po.interp_matrix_from_idx(idx, mesh.shape[0]), po.create_D_py(m), y)
    a = po.admm_update_py(y, inits, th0, lam, rho_matrix=rm)
    assert a["passes"] == int(golden[name + "_passes"])
    assert a["counter"] == int(golden[name + "_counter"]) == 1
    assert np.abs(a["theta"] - golden[name + "_theta"]).max() < 1e-12
    # C oracle (cpp stacking order: a row permutation, theta is order independent)
    b = co.mbs_one(x, y, m, axes, lam, mode=co.MODE_PY, theta_init=th0, rho_matrix0=rm)
    assert np.array_equal(b["idx"], golden[name + "_idx"])
    assert b["passes"] == int(golden[name + "_passes"])
    assert np.abs(b["theta"] - golden[name + "_theta"]).max() < 1e-11
    assert np.abs(b["fitted"] - golden[name + "_fitted"]).max() < 1e-11
    # and with the iterative x-update used at large N
    c = co.mbs_one(x, y, m, axes, lam, mode=co.MODE_PY, theta_init=th0, rho_matrix0=rm,
                   solver=co.SOLVER_PCG, cg_rtol=1e-13)
    assert c["passes"] == int(golden[name + "_passes"])
    assert np.abs(c["theta"] - golden[name + "_theta"]).max() < 1e-10


def test_lambda_max_property():
    """code/test_solvers.py:24-29: at a very large lambda the fit is the constant mean(y)."""
    x, y = synth(117, 10000, 2)
    m = [10, 10]
    axes = po.mesh_axes(x, m, po.MODE_PY)
    out = co.mbs_one(x, y, m, axes, 1e4, mode=co.MODE_PY)
    a, b, c = np.round(np.mean(out["theta"]), 3), np.round(np.mean(out["fitted"]), 3), np.round(np.mean(y), 3)
    assert a == b == c


@pytest.mark.parametrize("mode", [po.MODE_CPP, po.MODE_RCPP, po.MODE_PY])
@pytest.mark.parametrize("lam", [0.2, 1.0, 1.5, 7.3])
def test_two_oracles_agree_config1(mode, lam):
    """BASELINE config 1 (2-D, n=1000, 32x32) and the explicit lambdas of
    rcpp-code/test_mbs_cpp2r.R:50 (+ one lambda>2 so CPP mode's int rho is >1)."""
    x, y = synth(117, 1000, 2)
    m = [32, 32]
    axes = po.mesh_axes(x, m, mode)
    a = po.mbs_one(x, y, m, axes, lam=lam, mode=mode)
    b = co.mbs_one(x, y, m, axes, lam, mode=mode)
    c = co.mbs_one(x, y, m, axes, lam, mode=mode, solver=co.SOLVER_PCG, cg_rtol=1e-13)
    assert a["counter"] == b["counter"] == c["counter"]
    assert np.abs(a["theta"] - b["theta"]).max() < 1e-10
    assert np.abs(a["theta"] - c["theta"]).max() < 1e-9
    assert np.abs(a["u"] - b["u"]).max() < 1e-10
    assert a["rho"] == b["rho"]
    if mode == po.MODE_CPP and lam < 1:
        assert a["counter"] == 3          # int rho == 0 -> exits at Counter = 3
    if mode == po.MODE_CPP:
        assert a["rho"] == 0.0            # truncation drives rho to 0


@pytest.mark.parametrize("mode", [po.MODE_CPP, po.MODE_RCPP])
@pytest.mark.parametrize("dims,deltas", [([6, 6, 6], None), ([5, 5, 5], "auto"), ([4, 4, 4, 4], None),
                                          ([9, 7], "auto"), ([16], None)])
def test_two_oracles_agree_p134(mode, dims, deltas):
    p = len(dims)
    x, y = synth(7 + p, 400, p, 0.0, 1.0, 0.5)
    axes = po.mesh_axes(x, dims, mode)
    d = po.create_deltas(x, dims, mode) if deltas == "auto" else None
    for lam in (0.5, 3.0):
        a = po.mbs_one(x, y, dims, axes, lam=lam, mode=mode, deltas=d)
        b = co.mbs_one(x, y, dims, axes, lam, mode=mode, deltas=d)
        assert a["counter"] == b["counter"]
        assert np.abs(a["theta"] - b["theta"]).max() < 1e-9
        assert np.abs(a["fitted"] - b["fitted"]).max() < 1e-9


def test_rcpp_warm_start_path():
    """rcpp mbs_path (solvers.cpp:204-222): theta, u, rho carried along the lambda path; the cached
    matrix for the first pass of each solve uses the carried rho (:213)."""
    x, y = synth(5, 500, 2)
    m = [12, 12]
    axes = po.mesh_axes(x, m, po.MODE_RCPP)
    lams = [4.0, 1.0, 0.25]
    th_a = th_b = None
    u_a = u_b = None
    rho_a = rho_b = lams[0] / 5.0
    for lam in lams:
        a = po.mbs_one(x, y, m, axes, lam=lam, mode=po.MODE_RCPP, theta_init=th_a, u_init=u_a,
                       rho_init=rho_a, rho_matrix0=rho_a)
        b = co.mbs_one(x, y, m, axes, lam, mode=co.MODE_RCPP, theta_init=th_b, u_init=u_b,
                       rho_init=rho_b, rho_matrix0=rho_b)
        assert a["counter"] == b["counter"]
        assert np.abs(a["theta"] - b["theta"]).max() < 1e-9
        th_a, u_a, rho_a = a["theta"], a["u"], a["rho"]
        th_b, u_b, rho_b = b["theta"], b["u"], b["rho"]
        assert rho_a == rho_b


def test_cpp_nonconvergence_status():
    """cpp-code/solvers.cpp:122-124 throws when counter > max_counter."""
    x, y = synth(117, 1000, 2)
    axes = po.mesh_axes(x, [32, 32], po.MODE_CPP)
    with pytest.raises(ValueError):
        po.mbs_one(x, y, [32, 32], axes, lam=7.3, mode=po.MODE_CPP, max_counter=3)
    b = co.mbs_one(x, y, [32, 32], axes, 7.3, mode=co.MODE_CPP, max_counter=3)
    assert b["status"] == 1
