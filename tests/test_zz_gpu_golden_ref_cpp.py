"""The CUDA path against golden vectors produced by the REFERENCE's own compiled C++ solver
(tests/golden/ref_cpp_golden.npz, see tests/golden/make_golden_ref_cpp.py): identical Counter, max|theta - theta_ref| <= 1e-9.
Runs through the C ABI (multivartv_b200.Plan -> libmvtv_b200.so); nothing here touches /root/reference."""
import os

import numpy as np
import pytest

from tests.helpers import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_cpp_golden.npz"))


@pytest.fixture(scope="module")
def mv():
    import multivartv_b200 as m
    from multivartv_b200 import build
    build.build()
    return m


def test_cpp_mode_solves_match_compiled_reference(mv, gold):
    for k, row in enumerate(gold["solve_cases"]):
        seed, n, p, lam = int(row[0]), int(row[1]), int(row[2]), float(row[3])
        m = [int(v) for v in row[4:4 + p]]
        x, y = synth(seed, n, p)
        out = mv.mbs_one(x, y, m, tune=lam, mode="cpp")          # stand-alone mbs_one: mesh of create_mesh, unit block scales
        assert out["counter"] == int(gold["solve%d_counter" % k]), (m, lam)
        assert np.abs(out["theta.hat"] - gold["solve%d_theta" % k]).max() <= 1e-9
        assert np.abs(out["fitted"] - gold["solve%d_fitted" % k]).max() <= 1e-9


def test_cpp_mode_lambda_path_matches_compiled_reference(mv, gold):
    seed, n = int(gold["path_case"][0]), int(gold["path_case"][1])
    m = [int(v) for v in gold["path_case"][2:]]
    x, y = synth(seed, n, len(m))
    with mv.Plan(m, deltas=mv.create_deltas(x, m, "cpp")) as pl:   # mbs()'s operators (cpp-code/solvers.cpp:279-287)
        pl.set_points(x, y, mv.mesh_axes(x, m, "cpp"))
        out = pl.solve_path(gold["path_lambdas"], y, mode="cpp", want_thetas=True, cg_rtol=1e-14)
    assert list(out["counters"]) == [int(c) for c in gold["path_counters"]]
    assert np.allclose(out["mses"], gold["path_mses"], rtol=1e-9, atol=0)
    for i in range(len(gold["path_lambdas"])):
        assert np.abs(out["thetas"][i] - gold["path_thetas"][i]).max() <= 1e-9


def test_operator_rows_match_compiled_reference(mv, gold):
    from multivartv_b200 import utils
    for key in [k for k in gold.files if k.startswith("D_")]:
        dims = [int(v) for v in key.split("_")[1].split("x")]
        deltas = [0.3, 0.5, 2.0] if key.endswith("_deltas") else None
        D = utils.create_D(dims, deltas)
        rng = np.random.RandomState(4)
        th, w = rng.normal(size=gold[key].shape[1]), rng.normal(size=gold[key].shape[0])
        assert D.shape == gold[key].shape
        assert np.abs(D.dot(th) - gold[key] @ th).max() <= 1e-13
        assert np.abs(D.T.dot(w) - gold[key].T @ w).max() <= 1e-13
        D.close()


@pytest.fixture(scope="module")
def rgold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_rcpp_golden.npz"))


def test_rcpp_mode_solves_match_compiled_reference(mv, rgold):
    """RCPP mode against the Rcpp-side reference compiled from rcpp-code/MultivarTV/src (tests/golden/ref_rcpp_golden.npz):
    identical Counter, theta <= 1e-9, u <= 1e-7, same final rho."""
    for k, row in enumerate(rgold["solve_cases"]):
        seed, n, p, lam = int(row[0]), int(row[1]), int(row[2]), float(row[3])
        m = [int(v) for v in row[4:4 + p]]
        x, y = synth(seed, n, p)
        with mv.Plan(m) as pl:
            pl.set_points(x, y, mv.mesh_axes(x, m, "rcpp"))
            out = pl.solve(lam, mode="rcpp", want_u=True)
        assert out["counter"] == int(rgold["solve%d_counter" % k]), (m, lam)
        assert np.abs(out["theta"] - rgold["solve%d_theta" % k]).max() <= 1e-9
        assert np.abs(out["fitted"] - rgold["solve%d_fitted" % k]).max() <= 1e-9
        assert np.abs(out["u"] - rgold["solve%d_u" % k]).max() <= 1e-7
        assert abs(out["rho"] - float(rgold["solve%d_rho" % k])) <= 1e-12 * abs(out["rho"])
