"""Oracle vs the reference's own known answers and golden vectors (operators).
CPU-only.  Known answers: code/test_utils.py:10-64, cpp-code/utils_test.cpp:64-97.
Golden vectors: tests/golden/ref_py_golden.npz, produced by executing the reference Python
prototype (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po


# ---- known answers of code/test_utils.py ---------------------------------------------------
def test_t2v_known_answers():                      # code/test_utils.py:10-29
    assert po.tensor2vector(3, [0, 0, 0], [3, 3, 3]) == 0
    assert po.tensor2vector(3, [2, 2, 2], [3, 3, 3]) == 26
    assert po.vector2tensor(3, 0, [3, 3, 3]) == [0, 0, 0]
    assert po.vector2tensor(3, 26, [3, 3, 3]) == [2, 2, 2]


def test_create_D_null_vector():                   # code/test_utils.py:33-36
    theta = np.tile([1, -1, 1], 3)
    assert np.sum(po.create_D_py([3, 3], None) @ theta) == 0.0
    assert np.sum(po.create_D(2, [3, 3], None) @ theta) == 0.0
    assert np.sum(co.Operator([3, 3]).D(theta.astype(float))) == 0.0


def test_nearest_known_answers():                  # code/test_utils.py:40-57
    mesh = np.array([[0], [0.5], [1.0]])
    assert po.nearest1_unit(np.array([0.1]), mesh) == 0
    assert list(po.nearest1(np.array([0.1, 0.9]), mesh)) == [0, 2]
    O = po.nearest_interp_matrix(np.array([0.1, 0.9]), mesh)
    theta = mesh * np.array([[1], [2], [3]])
    assert np.all(O @ theta == np.array([[0], [3]]))
    axes = [np.array([0, 0.5, 1.0])]
    assert list(co.nearest([3], axes, np.array([0.1, 0.9]))) == [0, 2]
    assert list(co.nearest([3], axes, np.array([0.1, 0.9]), brute=True)) == [0, 2]


def test_mesh_delta_known_answer(golden):          # code/test_utils.py:61-64
    data = np.linspace(0.01, 0.99, 10).reshape(10, 1)
    ax = po.mesh_axes(data, [6], po.MODE_PY)[0]
    assert np.round(np.diff(ax)[0], 2) == 0.20
    assert np.allclose(ax, golden["mesh_coords_1d_mesh"].ravel(), rtol=0, atol=1e-15)


def test_binaries_known_answer():                  # cpp-code/utils_test.cpp:90-97 ("7 binaries for p=3")
    assert ["".join(map(str, po.dec2binary(i, 3))) for i in range(1, 8)] == \
        ["001", "010", "011", "100", "101", "110", "111"]


# ---- golden vectors -----------------------------------------------------------------------
@pytest.mark.parametrize("dims", [[3, 3, 3], [3, 2, 3], [4, 5], [2, 3, 4, 2]])
def test_index_maps_golden(golden, dims):
    key = "x".join(map(str, dims))
    N = int(np.prod(dims))
    v2t = np.array([po.vector2tensor(len(dims), i, dims) for i in range(N)])
    assert np.array_equal(v2t, golden["v2t_" + key])
    t2v = np.array([po.tensor2vector(len(dims), v2t[i], dims) for i in range(N)])
    assert np.array_equal(t2v, golden["t2v_" + key])
    assert np.array_equal(t2v, np.arange(N))


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_fd_binaries_golden(golden, p):
    assert np.array_equal(po.fd_binaries(p), golden["fd_binaries_%d" % p])


DIMS = [[5], [3, 3], [4, 3], [3, 5], [3, 3, 3], [4, 4, 4], [3, 3, 4], [3, 3, 3, 3]]


@pytest.mark.parametrize("dims", DIMS)
def test_D_golden_bit_exact(golden, dims):
    """Materialised D (python stacking order) is bit-identical to the reference's, including the
    mixedpartial direction-0 quirk for p>=3."""
    key = "x".join(map(str, dims))
    assert np.array_equal(po.create_D_py(dims).toarray(), golden["D_" + key])


def test_D_golden_deltas(golden):
    assert np.array_equal(po.create_D_py([4, 3], golden["deltas_4x3"]).toarray(), golden["D_4x3_deltas"])
    assert np.array_equal(po.create_D_py([3, 3, 3], golden["deltas_3x3x3"]).toarray(), golden["D_3x3x3_deltas"])


def _cpp_order_from_py_order(Dpy, rows_per_block):
    """cpp stacking = all-ones block first, then the python order without its last block
    (cpp-code/utils.cpp:258-267 vs code/utils.py:138-149)."""
    offs = np.concatenate([[0], np.cumsum(rows_per_block)])
    K = len(rows_per_block)
    order = [K - 1] + list(range(K - 1))
    return np.vstack([Dpy[offs[b]:offs[b + 1]] for b in order])


@pytest.mark.parametrize("dims", DIMS)
def test_D_cpp_order_and_matrix_free(golden, dims):
    """cpp-order D == block permutation of the golden D, and the matrix-free C oracle applies
    exactly that matrix (D and D^T)."""
    key = "x".join(map(str, dims))
    op = co.Operator(dims)
    # C block 0 is all-ones; python order puts it last
    rows_py = op.block_rows[1:] + op.block_rows[:1]
    Dcpp = _cpp_order_from_py_order(golden["D_" + key], rows_py)
    assert np.array_equal(po.create_D(len(dims), dims).toarray(), Dcpp)
    rng = np.random.RandomState(1)
    th, w = rng.normal(size=op.N), rng.normal(size=op.R)
    assert np.abs(Dcpp @ th - op.D(th)).max() < 1e-13
    assert np.abs(Dcpp.T @ w - op.Dt(w)).max() < 1e-13


def test_D_noncubic_p3_fails_like_reference(golden):
    assert int(golden["D_noncubic_ok_3x4x5"]) == 0 and int(golden["D_noncubic_ok_4x3x3"]) == 0
    for dims in ([3, 4, 5], [4, 3, 3]):
        with pytest.raises(ValueError):
            po.create_D(3, dims)
        with pytest.raises(ValueError):
            co.Operator(dims)
    # the intended operator (utils.cpp:186) works on any mesh
    op = co.Operator([3, 4, 5], variant=co.VARIANT_INTENDED)
    D = po.create_D(3, [3, 4, 5], variant=po.VARIANT_INTENDED)
    th = np.random.RandomState(2).normal(size=op.N)
    assert np.abs(D @ th - op.D(th)).max() < 1e-13


def test_quirk_is_axis0_replacement():
    """for p=3 the block of mask [0,1,1] equals d0 d2 (not d1 d2)."""
    m = 4
    op = co.Operator([m, m, m])
    assert op.masks == [0b111, 0b100, 0b010, 0b101, 0b001, 0b101, 0b011]


def test_nearest_golden(golden):
    x, m, mesh = golden["near2_x"], golden["near2_m"], golden["near2_mesh"]
    idx_ref = golden["near2_idx"]
    assert np.array_equal(po.nearest1(x, mesh), idx_ref)
    axes = [mesh[: m[0], 0], mesh[:: m[0], 1]]
    assert np.array_equal(po.nearest1_separable(x, axes, m), idx_ref)
    assert np.array_equal(co.nearest(m, axes, x), idx_ref)
    assert np.array_equal(co.nearest(m, axes, x, brute=True), idx_ref)


def test_nearest_ties_go_to_lowest_index():
    axes = [np.array([0.0, 1.0, 2.0]), np.array([0.0, 1.0])]
    x = np.array([[0.5, 0.5], [1.5, 0.5], [1.0, 0.25]])
    mesh = np.array([[a, b] for b in axes[1] for a in axes[0]])
    ref = po.nearest1(x, mesh)
    assert list(ref) == [0, 1, 1]
    assert np.array_equal(co.nearest([3, 2], axes, x), ref)
    assert np.array_equal(po.nearest1_separable(x, axes, [3, 2]), ref)


def test_softthresh_golden(golden):
    assert np.array_equal(po.softthresh(golden["soft_z"], 0.9), golden["soft_out_0p9"])
    assert np.all(po.softthresh(golden["soft_z"], np.inf) == 0.0)


def test_create_mesh_modes():
    rng = np.random.RandomState(3)
    x = rng.uniform(0, 1, (50, 2))
    ax = po.mesh_axes(x, [4, 5], po.MODE_CPP)
    assert ax[0][0] == np.float32(x[:, 0].min() + 0.01) and ax[0][-1] == np.float32(x[:, 0].max() + 0.01)
    ax = po.mesh_axes(x, [4, 5], po.MODE_RCPP)
    assert ax[1][0] == x[:, 1].min() - 1e-4 and ax[1][-1] == x[:, 1].max() + 1e-4
    d = po.create_deltas(x, [4, 5], po.MODE_CPP)
    assert np.allclose(d, (x.max(0) - x.min(0) + 0.02) / np.array([4, 5]))
    mesh = po.create_mesh(x, [4, 5], po.MODE_RCPP)
    assert mesh.shape == (20, 2) and mesh[5, 0] == ax[0][1] and mesh[5, 1] == ax[1][1]
