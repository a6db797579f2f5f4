"""The reference's own unit tests (code/test_utils.py:10-64) and the known answers printed by
cpp-code/utils_test.cpp:64-97, written against multivartv_b200.utils -- same names, same calls, same expected
values.  Index maps are host logic (CPU tests); the operator tests run D and O on the GPU."""
import numpy as np
import pytest

from multivartv_b200 import utils
from oracle import py_oracle as po

mydim = np.array([3, 3, 3])


# ---- code/test_utils.py:10-31 / cpp-code/utils_test.cpp:64-71 ---------------------------------------------
def test_t2v_0():
    assert utils.t2v_unit(dims=mydim, ind=np.array([0, 0, 0])) == 0


def test_v2t_000():
    assert all(utils.v2t_unit(dims=mydim, ind=0) == np.array([0, 0, 0]))


def test_t2v_26():
    assert utils.t2v_unit(dims=mydim, ind=np.array([2, 2, 2])) == 26


def test_v2t_222():
    assert all(utils.v2t_unit(dims=mydim, ind=26) == np.array([2, 2, 2]))


def test_cpp_signatures_and_closures():
    assert utils.tensor2vector(3, [0, 0, 0], [3, 3, 3]) == 0           # "0 is correct"
    assert utils.tensor2vector(3, [2, 2, 2], [3, 3, 3]) == 26          # "27-1 is correct"
    assert utils.vector2tensor(3, 0, [3, 3, 3]) == [0, 0, 0]
    assert utils.vector2tensor(3, 26, [3, 3, 3]) == [2, 2, 2]
    assert [utils.vector2tensor(3, i, [3, 2, 3]) for i in range(3)] == [[0, 0, 0], [1, 0, 0], [2, 0, 0]]   # utils_test.cpp:73-82
    assert utils.t2v(np.array([3, 3]))([2, 2]) == 8                     # code/utils.py:21 example
    assert list(utils.v2t(np.array([3, 3]))(8)) == [2, 2]
    assert utils.prod(2, [3, 4, 5]) == 12


@pytest.mark.parametrize("dims", [[3, 3, 3], [3, 2, 3], [4, 5], [7], [2, 3, 2, 3]])
def test_index_maps_round_trip_and_match_oracle(dims):
    p, N = len(dims), int(np.prod(dims))
    for v in range(N):
        mi = utils.vector2tensor(p, v, dims)
        assert mi == list(po.vector2tensor(p, v, dims))
        assert utils.tensor2vector(p, mi, dims) == v == po.tensor2vector(p, mi, dims)


def test_binaries_match_reference_print_and_golden(golden):
    # cpp-code/utils_test.cpp:84-97 prints 001 010 011 100 101 110 111 for p = 3
    assert ["".join(map(str, utils.dec2binary(i, 3))) for i in range(1, 8)] == ["001", "010", "011", "100", "101", "110", "111"]
    for p in (1, 2, 3, 4):
        assert np.array_equal(utils.fd_binaries(p), po.fd_binaries(p))
        assert np.array_equal(utils.fd_binaries(p), golden["fd_binaries_%d" % p])   # executed code/utils.py:63-69


@pytest.mark.parametrize("key", ["3x3x3", "3x2x3", "4x5", "2x3x4x2"])
def test_index_maps_match_golden_from_reference_python(golden, key):
    """tests/golden/make_golden.py executed code/utils.py:12-41 (t2v_unit / v2t_unit) over whole meshes."""
    d = np.array([int(v) for v in key.split("x")])
    v2t, t2v = golden["v2t_" + key], golden["t2v_" + key]
    for i in range(int(np.prod(d))):
        assert list(utils.v2t_unit(i, d)) == list(v2t[i])
        assert utils.t2v_unit(v2t[i], d) == t2v[i]


# ---- code/test_utils.py:58-64 -------------------------------------------------------------------------------
def test_mesh_coords(golden):
    data = np.linspace(0.01, 0.99, 10)
    result = utils.mesh_coords(data.reshape((10, 1)), mesh_dims=np.array([6]))
    assert np.round(result["deltas"], 2) == 0.20
    assert np.array_equal(result["mesh"], golden["mesh_coords_1d_mesh"])              # executed code/utils.py:179-193
    assert np.array_equal(np.asarray(result["deltas"]), golden["mesh_coords_1d_deltas"])


# ---- code/test_utils.py:33-56: operators, on the device ------------------------------------------------------
@pytest.mark.gpu
def test_create_D():
    D = utils.create_D(dims=np.array([3, 3]), deltas=None)
    theta = np.tile([1, -1, 1], 3)
    assert np.sum(D.dot(theta)) == 0.0
    assert D.shape == po.create_D(2, [3, 3]).shape
    # D and D^T against the materialised reference operator, and the adjoint identity
    Dref = po.create_D(2, [3, 3])
    rng = np.random.RandomState(0)
    th, w = rng.normal(size=9), rng.normal(size=D.shape[0])
    assert np.abs(D.dot(th) - Dref.dot(th)).max() <= 1e-14
    assert np.abs(D.T.dot(w) - Dref.T.dot(w)).max() <= 1e-14
    assert abs(np.dot(D @ th, w) - np.dot(th, D.T @ w)) <= 1e-12
    D.close()


@pytest.mark.gpu
def test_nearest1_unit():
    target = np.array(0.1)
    choices = np.array([[0], [0.5], [1.0]])
    assert utils.nearest1_unit(target, choices) == 0


@pytest.mark.gpu
def test_nearest1():
    data = np.array([0.1, 0.9])
    mesh = np.array([[0], [0.5], [1.0]])
    assert all(utils.nearest1(data, mesh) == np.array([0, 2]))


@pytest.mark.gpu
def test_nearest_interp_matrix():
    data = np.array([0.1, 0.9])
    mesh = np.array([[0], [0.5], [1.0]])
    O = utils.nearest_interp_matrix(data, mesh)
    theta = mesh * np.array([[1], [2], [3]])
    assert all(O.dot(theta) == np.array([[0], [3]]))
    assert O.shape == (2, 3)
    assert list(O.T.dot(np.array([5.0, 7.0]))) == [5.0, 0.0, 7.0]      # Ot*y, cpp-code/solvers.cpp:40
    O.close()
