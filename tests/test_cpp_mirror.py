"""The header-only C++ mirror of cpp-code/solvers.hpp (multivartv_b200/host/mvtv_solvers.hpp) compiles against
include/mvtv.h and links to libmvtv_b200.so (CPU check); on the GPU it reproduces the oracle (gpu check)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mbs_one_cli.cpp")


def _build(tmp_path):
    from multivartv_b200 import build
    lib = build.build()
    exe = str(tmp_path / "mbs_one_cli")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", SRC, "-o", exe, "-L", os.path.dirname(lib),
                           "-lmvtv_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    # usage error path runs without touching CUDA
    assert subprocess.run([exe]).returncode == 2


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
