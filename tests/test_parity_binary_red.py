"""GPU parity for BINARY_RED (array_equal / allclose): the CUDA kernel behind `cnb_binary_red`
against the reference's BinaryOp<EQUAL|ISCLOSE> functors folded by the CPU loop of
binary/binary_red.cc:38-47 (oracle.ref.binary_red).  The result is a bool, so parity is exact.
Cases follow tests/integration/test_array_equal.py and test_allclose.py of the reference:
equal / differing arrays per dtype, a single mismatch anywhere (first, last, every tile position),
NaNs, broadcast operands, shape mismatch, empty arrays, strided views."""
import numpy as np
import pytest

from oracle import ref

import parity_utils as pu

pytestmark = pytest.mark.gpu

ALL = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
       np.float16, np.float32, np.float64, np.complex64, np.complex128]


def _name(d):
    return np.dtype(d).name


@pytest.mark.parametrize("dt", ALL, ids=_name)
@pytest.mark.parametrize("n", [1, 31, 4096, 100003])
def test_array_equal_dtypes(dt, n):
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    rng = pu.rng_for("array_equal", dt.name, n)
    a = pu.make_input(dt, n, rng, "small")
    if dt.kind in "fc":
        a = np.where(np.isnan(a), np.ones_like(a), a)
    A = cn.array(a)
    assert bool(cn.array_equal(A, cn.array(a.copy()))) is ref.binary_red("EQUAL", a, a.copy()) is True
    for pos in sorted({0, n // 2, n - 1}):
        b = a.copy()
        b[pos] = (not b[pos]) if dt.kind == "b" else b[pos] + dt.type(1)
        exp = ref.binary_red("EQUAL", a, b)
        assert bool(cn.array_equal(A, cn.array(b))) is exp
        assert exp is False


def test_array_equal_every_position_in_a_tile():
    """one mismatch at each lane/unroll slot of the first tiles and of the ragged tail"""
    import cunumeric_b200 as cn

    n = 3 * 8192 + 77
    a = np.arange(n, dtype=np.float32)
    A = cn.array(a)
    B = cn.array(a.copy())
    for pos in list(range(0, 2200, 37)) + list(range(n - 90, n)):
        B[pos] = -1.0
        assert bool(cn.array_equal(A, B)) is False, pos
        B[pos] = float(a[pos])
        assert bool(cn.array_equal(A, B)) is True, pos


@pytest.mark.parametrize("dt", [np.float16, np.float32, np.float64, np.complex64, np.complex128], ids=_name)
def test_nan_never_equal(dt):
    import cunumeric_b200 as cn

    a = np.ones(1000, dtype=dt)
    a[777] = np.nan
    assert ref.binary_red("EQUAL", a, a.copy()) is False
    assert bool(cn.array_equal(cn.array(a), cn.array(a.copy()))) is False
    assert bool(cn.allclose(cn.array(a), cn.array(a.copy()))) is ref.binary_red("ISCLOSE", a, a.copy())
    with pytest.raises(NotImplementedError):
        cn.array_equal(cn.array(a), cn.array(a), equal_nan=True)


def test_array_equal_shape_mismatch_and_empty():
    import cunumeric_b200 as cn

    assert cn.array_equal(cn.zeros((3, 4)), cn.zeros((4, 3))) is False
    assert cn.array_equal(cn.zeros((3, 4)), cn.zeros((3,))) is False
    # empty rect: nothing folds, the pre-filled True stands (binary_red_template.inl:48-51)
    assert bool(cn.array_equal(cn.zeros((0,)), cn.zeros((0,)))) is True
    assert bool(cn.array_equal(cn.zeros((0, 5)), cn.zeros((0, 5)))) is True


def test_array_equal_mixed_dtypes_and_scalars():
    import cunumeric_b200 as cn

    a = np.arange(100, dtype=np.int32)
    assert bool(cn.array_equal(cn.array(a), cn.array(a.astype(np.float64)))) is True
    assert bool(cn.array_equal(cn.array(a), a.tolist())) is True
    assert bool(cn.array_equal(cn.array(np.float32(3.0)), cn.array(np.float32(3.0)))) is True
    assert bool(cn.array_equal(cn.array(np.float32(3.0)), cn.array(np.float32(4.0)))) is False


def test_array_equal_views():
    import cunumeric_b200 as cn

    rng = pu.rng_for("array_equal_views")
    g = rng.integers(-5, 5, size=(130, 257)).astype(np.int64)
    G, H = cn.array(g), cn.array(g.copy())
    assert bool(cn.array_equal(G[1:-1, 3:-2], H[1:-1, 3:-2])) is True
    assert bool(cn.array_equal(G.T, H.T)) is True
    assert bool(cn.array_equal(G[::2, ::3], H[::2, ::3])) is True
    H[64, 100] += 1
    assert bool(cn.array_equal(G[1:-1, 3:-2], H[1:-1, 3:-2])) is False
    assert bool(cn.array_equal(G.T, H.T)) is False
    assert bool(cn.array_equal(G[::2, ::3], H[::2, ::3])) is np.array_equal(g[::2, ::3], H.__array__()[::2, ::3])
    assert bool(cn.array_equal(G[:64], H[:64])) is True


@pytest.mark.parametrize("dt", [np.float16, np.float32, np.float64, np.int32, np.complex64, np.complex128],
                         ids=_name)
@pytest.mark.parametrize("rtol,atol", [(1e-5, 1e-8), (1e-2, 0.0), (0.0, 1e-3), (0.0, 0.0)])
def test_allclose(dt, rtol, atol):
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    rng = pu.rng_for("allclose", dt.name, rtol, atol)
    n = 50000
    a = pu.make_input(dt, n, rng, "small")
    if dt.kind in "fc":
        a = np.where(np.isnan(a) | np.isinf(a), np.ones_like(a), a)
    for scale in (0.0, 1e-9, 1e-6, 1e-3, 1e-1):
        if dt.kind in "iu":
            b = a + dt.type(int(scale * 10))
        else:
            b = (a * (1 + scale)).astype(dt)
        exp = ref.binary_red("ISCLOSE", a, b, rtol, atol)
        got = bool(cn.allclose(cn.array(a), cn.array(b), rtol=rtol, atol=atol))
        assert got is exp, (scale, got, exp)


def test_allclose_broadcast_and_single_outlier():
    import cunumeric_b200 as cn

    a = np.full((64, 1000), 2.0, dtype=np.float64)
    row = np.full((1000,), 2.0 + 1e-9, dtype=np.float64)
    assert bool(cn.allclose(cn.array(a), cn.array(row))) is True
    assert bool(cn.allclose(cn.array(a), 2.0)) is True
    a[63, 999] = 2.1
    exp = ref.binary_red("ISCLOSE", a, np.broadcast_to(row, a.shape).copy())
    assert bool(cn.allclose(cn.array(a), cn.array(row))) is exp is False


def test_binary_red_rejects_other_ops():
    """binary_op_util.h:149-161: reduce_op_dispatch only knows EQUAL and ISCLOSE"""
    import ctypes

    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib
    from cunumeric_b200.config import BinaryOpCode

    a = cn.ones((10,), dtype=np.float32)
    out = cn.ones((1,), dtype=np.bool_)
    d_out, d_a = out._thunk.base.descriptor(), a._thunk.base.descriptor()
    rc = cn.runtime.lib.cnb_binary_red(int(BinaryOpCode.ADD), ctypes.byref(d_out), ctypes.byref(d_a),
                                       ctypes.byref(d_a), None, cn.runtime.stream)
    assert rc == _lib.CNB_ERR_INVALID_OP if hasattr(_lib, "CNB_ERR_INVALID_OP") else rc == -1
