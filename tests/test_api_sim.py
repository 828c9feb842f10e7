"""Host-side API logic on CPU (no GPU): the ufunc type resolution of cunumeric_b200/_ufunc against
NumPy for every dtype pair and scalar kind, through the host-memory stand-in for the CUDA library
(tests/sim_backend.py).  The values come from the stand-in, so only dtypes, shapes and error
behaviour are asserted here; values are the business of the -m gpu parity suites."""
import numpy as np
import pytest

from test_fusion_sim import sim  # noqa: F401  (fixture)

DTYPES = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
          np.float16, np.float32, np.float64, np.complex64, np.complex128]
SCALARS = [2, 2.5, 1 + 2j, True, np.float32(2.5), np.int8(3), np.float64(1.5), np.uint64(7)]


def _dtype_or_error(fn):
    try:
        return fn().dtype
    except TypeError:
        return "TypeError"


@pytest.mark.parametrize("op", ["add", "multiply", "maximum", "greater", "less", "subtract"])
def test_binary_result_dtypes_match_numpy_for_every_pair(sim, op):  # noqa: F811
    """cunumeric/_ufunc/ufunc.py:618-781 (`binary_ufunc.__call__`, `_resolve_dtype`): ordered type
    tables + first castable signature, weak Python scalars."""
    import cunumeric_b200 as cn

    for da in DTYPES:
        a = np.ones((3,), dtype=da)
        for db in DTYPES:
            if op == "subtract" and da == np.bool_ and db == np.bool_:
                continue   # the reference's table accepts bool - bool (math.py:38-43); NumPy raises
            b = np.ones((3,), dtype=db)
            exp = _dtype_or_error(lambda: getattr(np, op)(a, b))
            got = _dtype_or_error(lambda: getattr(cn, op)(cn.array(a), cn.array(b)))
            assert got == exp, (op, np.dtype(da).name, np.dtype(db).name, got, exp)
            cn.flush()
        for sc in SCALARS:
            if op == "subtract" and da == np.bool_ and isinstance(sc, bool):
                continue
            exp = _dtype_or_error(lambda: getattr(np, op)(a, sc))
            got = _dtype_or_error(lambda: getattr(cn, op)(cn.array(a), sc))
            assert got == exp, (op, np.dtype(da).name, repr(sc), got, exp)
            got_r = _dtype_or_error(lambda: getattr(cn, op)(sc, cn.array(a)))
            exp_r = _dtype_or_error(lambda: getattr(np, op)(sc, a))
            assert got_r == exp_r, (op, repr(sc), np.dtype(da).name, got_r, exp_r)
            cn.flush()


def test_out_and_casting_rules(sim):  # noqa: F811
    """ufunc.py:262-315: `out=` must be castable from the computed type under `casting`."""
    import cunumeric_b200 as cn

    a = cn.array(np.arange(6, dtype=np.float64))
    out64, out32, outi = cn.empty((6,)), cn.empty((6,), dtype=np.float32), cn.empty((6,), dtype=np.int32)
    assert cn.add(a, a, out=out64) is out64
    assert cn.add(a, a, out=out32) is out32              # same_kind (the default): float64 -> float32
    with pytest.raises(TypeError):
        cn.add(a, a, out=out32, casting="safe")
    with pytest.raises(TypeError):
        cn.add(a, a, out=outi)                           # float64 -> int32 is not same_kind
    cn.add(a, a, out=outi, casting="unsafe")
    assert np.array_equal(np.array(outi), (np.arange(6) * 2).astype(np.int32))
    with pytest.raises(ValueError):
        cn.add(a, a, out=cn.empty((5,)))
    assert cn.add(a, a, dtype=np.float32).dtype == np.float32
    with pytest.raises(NotImplementedError):
        cn.add(a, a, where=cn.array(np.ones(6, dtype=bool)))   # ufunc.py:338-341


def test_broadcasting_and_shape_errors(sim):  # noqa: F811
    import cunumeric_b200 as cn

    a = cn.array(np.ones((4, 1, 5)))
    b = cn.array(np.ones((3, 1)))
    assert cn.add(a, b).shape == (4, 3, 5)
    assert (a + 2).shape == (4, 1, 5) and (2 + a).shape == (4, 1, 5)
    with pytest.raises(ValueError):
        cn.add(cn.array(np.ones((4, 3))), cn.array(np.ones((5, 3))))
    r = cn.maximum(cn.array(np.arange(5.0)), cn.array(np.array(2.0)))    # 0-d operand
    assert np.array_equal(np.array(r), np.maximum(np.arange(5.0), 2.0))


@pytest.mark.parametrize("op", ["negative", "absolute", "square"])
def test_unary_result_dtypes_match_numpy(sim, op):  # noqa: F811
    """ufunc.py:349-520 (`unary_ufunc.__call__`): first signature the input can be cast to safely."""
    import cunumeric_b200 as cn

    for dt in DTYPES:
        if dt == np.bool_ and op in ("negative", "square"):
            continue   # NumPy refuses -bool; the reference's tables start at the integers for these
        a = np.ones((4,), dtype=dt)
        exp = _dtype_or_error(lambda: getattr(np, op)(a))
        got = _dtype_or_error(lambda: getattr(cn, op)(cn.array(a)))
        assert got == exp, (op, np.dtype(dt).name, got, exp)
        cn.flush()
    a = cn.array(np.arange(4, dtype=np.int32))
    assert cn.negative(a, dtype=np.float64).dtype == np.float64
    out = cn.empty((4,), dtype=np.int64)
    assert cn.negative(a, out=out) is out and np.array_equal(np.array(out), -np.arange(4))
