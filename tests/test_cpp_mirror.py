"""The header-only C++ mirror of cpp-code/solvers.hpp (multivartv_b200/host/mvtv_solvers.hpp) compiles against
include/mvtv.h and links to libmvtv_b200.so (CPU check); on the GPU it reproduces the oracle (gpu check)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPPDIR = os.path.join(ROOT, "tests", "cpp")


def _build(tmp_path, name="mbs_one_cli"):
    from multivartv_b200 import build
    lib = build.build()
    exe = str(tmp_path / name)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(CPPDIR, name + ".cpp"),
                           "-o", exe, "-L", os.path.dirname(lib), "-lmvtv_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


@pytest.mark.parametrize("name", ["mbs_one_cli", "mbs_cli", "solvers_test"])
def test_cpp_mirror_compiles_and_links(tmp_path, name):
    exe = _build(tmp_path, name)
    assert os.path.exists(exe)
    if name != "solvers_test":
        # usage error path runs without touching CUDA
        assert subprocess.run([exe]).returncode == 2


def test_cpp_utils_known_answers(tmp_path):
    """cpp-code/utils_test.cpp's printed known answers, asserted (host index maps: runs without a GPU)."""
    r = subprocess.run([_build(tmp_path, "utils_test")], capture_output=True, text=True)
    assert r.returncode == 0 and "utils_test ok" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback behind the C++ interface either: without a device the first call throws."""
    import ctypes
    from multivartv_b200 import _lib
    n = ctypes.c_int(0)
    if _lib.load().mvtv_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    exe = _build(tmp_path, "solvers_test")
    r = subprocess.run([exe, "50", "2"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "mvtv error 2" in (r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,lam", [(0, 1.5), (0, 0.2), (1, 1.0)])
def test_cpp_mirror_matches_oracle(tmp_path, mode, lam):
    from oracle import c_oracle as co
    from oracle import py_oracle as po
    from tests.helpers import synth
    exe = _build(tmp_path)
    x, y = synth(117, 1000, 2)
    m = [20, 20]                                     # the shape of cpp-code/solvers_test.cpp:16-24
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qqqd", x.shape[0], 2, mode, lam))
        f.write(struct.pack("<qq", *m))
        f.write(np.ascontiguousarray(x.T).tobytes())
        f.write(np.ascontiguousarray(y).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(fout, "rb").read()
    counter, N, n = struct.unpack("<qqq", raw[:24])
    theta = np.frombuffer(raw[24:24 + 8 * N])
    fitted = np.frombuffer(raw[24 + 8 * N:24 + 8 * N + 8 * n])
    axes = po.mesh_axes(x, m, mode)
    # the CLI's rcpp branch goes through a cache, i.e. mbs_path's convention: first-pass matrix = crossO + rho_init*crossD
    ref = co.mbs_one(x, y, m, axes, lam, mode=mode, rho_matrix0=(lam / 5.0 if mode == 1 else None))
    assert counter == ref["counter"]
    assert np.abs(theta - ref["theta"]).max() <= 1e-9
    assert np.abs(fitted - ref["fitted"]).max() <= 1e-9
    assert ("Counter = %d" % ref["counter"]) in r.stdout


def _run_mbs_cli(tmp_path, x, y, m, mode, folds, n_lambda, lambdas, foldinds):
    exe = _build(tmp_path, "mbs_cli")
    fin, fout = str(tmp_path / "cv_in.bin"), str(tmp_path / "cv_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qqqqqq", x.shape[0], x.shape[1], mode, folds, n_lambda, 0 if lambdas is None else 1))
        f.write(struct.pack("<%dq" % len(m), *m))
        f.write(np.ascontiguousarray(x.T).tobytes())
        f.write(np.ascontiguousarray(y).tobytes())
        if lambdas is not None:
            f.write(np.ascontiguousarray(lambdas, dtype=np.float64).tobytes())
        if folds > 1:
            f.write(np.ascontiguousarray(foldinds, dtype=np.int64).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(fout, "rb").read()
    nl, N, n, best = struct.unpack("<qqqq", raw[:32])
    a = np.frombuffer(raw[32:])
    ncol = max(folds, 1)
    o = 0
    out = {"best": best, "stdout": r.stdout}
    for key, cnt in (("lambdas", nl), ("cv_mses", nl), ("mse_mat", nl * ncol), ("theta", N), ("fitted", n)):
        out[key] = a[o:o + cnt]
        o += cnt
    out["mse_mat"] = out["mse_mat"].reshape(ncol, nl).T      # column-major n_lambda x folds
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("mode,folds,grid", [("cpp", 3, False), ("rcpp", 3, False), ("rcpp", 1, False), ("cpp", 1, False),
                                             ("cpp", 2, True)])
def test_cpp_mbs_matches_python_driver(tmp_path, mode, folds, grid):
    """mbs (cpp-code/solvers.hpp:129) / rcpp::mbs_impl of the C++ mirror against multivartv_b200.mbs (itself checked
    against the oracle in test_gpu_parity.test_mbs_cv_driver) on the same injected folds."""
    import multivartv_b200 as mv
    from tests.helpers import synth
    x, y = synth(12, 600, 2)
    m = [8, 8]
    lambdas = None if grid else np.array([3.0, 1.0, 0.3, 0.1])
    n_lambda = 5 if grid else 4
    foldinds = mv.kfoldinds(600, folds, seed=5) if folds > 1 else None
    got = _run_mbs_cli(tmp_path, x, y, m, 0 if mode == "cpp" else 1, folds, n_lambda, lambdas, foldinds)
    ref = mv.mbs(x, y, m, n_lambda=n_lambda, lambdas=lambdas, folds=folds, mode=mode, foldinds=foldinds)
    assert np.allclose(got["lambdas"], ref["lambdas"], rtol=1e-12)
    assert np.allclose(got["mse_mat"], ref["cv.mse_mat"], rtol=1e-10, atol=1e-13)
    assert np.allclose(got["cv_mses"], ref["cv.mses"], rtol=1e-10, atol=1e-13)
    assert got["best"] == ref["lambda_minmse_ind"]
    assert np.abs(got["theta"] - ref["theta_hat"]).max() <= 1e-10
    assert np.abs(got["fitted"] - ref["fitted"]).max() <= 1e-10
    assert "MBS BEGINS: ntheta = 64" in got["stdout"]      # cpp-code/solvers.cpp:283


@pytest.mark.gpu
def test_cpp_solvers_test_flow(tmp_path):
    """The flow of cpp-code/solvers_test.cpp (softthresh, mbs with default folds on 10000 uniform points, 20x20 mesh,
    mse) with the assertions upstream leaves out; the lambda grid is cut to 12 values to keep the test short."""
    exe = _build(tmp_path, "solvers_test")
    r = subprocess.run([exe, "10000", "12"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "solvers_test ok" in r.stdout and "Tuned model Training MSE" in r.stdout
    assert r.stdout.count("Fold = ") == 5                   # cpp-code/solvers.hpp:129 folds = 5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["cpp", "rcpp"])
def test_adapt_step_matches_oracle(mode):
    import multivartv_b200 as mv
    from oracle import py_oracle as po
    rng = np.random.RandomState(3)
    u = rng.normal(size=500)
    for scale_r, scale_s in ((1.0, 1.0), (100.0, 1.0), (1.0, 100.0)):
        r, s = scale_r * rng.normal(size=700), scale_s * rng.normal(size=300)
        rho_next, u_next = mv.adapt_step(r, s, 1.5, u, mode=mode)
        ref = (po.adapt_step_cpp if mode == "cpp" else po.adapt_step_rcpp)(r, s, 1.5, u)
        assert rho_next == ref[0]
        assert np.array_equal(u_next, ref[1])
