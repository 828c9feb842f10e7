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


def _random_key(rng, shape):
    key = []
    nd = len(shape)
    n_idx = int(rng.integers(0, nd + 1))
    for d in range(n_idx):
        n = shape[d]
        r = rng.random()
        if r < 0.25 and n > 0:
            key.append(int(rng.integers(-n, n)))
        elif r < 0.9:
            start = None if rng.random() < 0.3 else int(rng.integers(-n - 1, n + 2))
            stop = None if rng.random() < 0.3 else int(rng.integers(-n - 1, n + 2))
            step = None if rng.random() < 0.5 else int(rng.choice([-3, -2, -1, 1, 2, 3]))
            key.append(slice(start, stop, step))
        else:
            key.append(slice(None))
    if rng.random() < 0.2 and n_idx < nd:
        key.insert(int(rng.integers(0, len(key) + 1)), Ellipsis)
    return tuple(key)


@pytest.mark.parametrize("seed", range(4))
def test_basic_indexing_matches_numpy(sim, seed):  # noqa: F811
    """Store view algebra behind ndarray.__getitem__ / __setitem__ (deferred.py:_basic_index; the
    reference's get_item / set_item, deferred.py:927-1094): random integer / slice / Ellipsis keys incl.
    negative steps, out-of-range bounds and empty results; views of views; assignment of arrays and
    scalars through a view."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(100 + seed)
    for _ in range(150):
        shape = tuple(int(rng.integers(1, 6)) for _ in range(int(rng.integers(1, 5))))
        a = rng.integers(-50, 50, size=shape).astype(np.int64)
        A = cn.array(a)
        key = _random_key(rng, shape)
        try:
            exp = a[key]
        except IndexError:      # an integer that lands on a shorter axis behind an Ellipsis
            with pytest.raises(IndexError):
                A[key]
            continue
        got = A[key]
        g = np.array(got)
        assert g.shape == exp.shape and np.array_equal(g, exp), (shape, key)
        key2 = _random_key(rng, exp.shape) if exp.ndim else ()
        try:
            exp2 = exp[key2]
        except IndexError:
            exp2 = None
        if exp2 is not None:
            assert np.array_equal(np.array(got[key2]), exp2), (shape, key, key2)
        val = rng.integers(-9, 9, size=exp.shape).astype(np.int64)
        a2, A2 = a.copy(), cn.array(a)
        a2[key] = val
        A2[key] = cn.array(val) if val.ndim else int(val)
        assert np.array_equal(np.array(A2), a2), (shape, key)
        a2[key] = 7
        A2[key] = 7
        assert np.array_equal(np.array(A2), a2), (shape, key)
    with pytest.raises(IndexError):
        cn.array(np.arange(5))[7]
    with pytest.raises(IndexError):
        cn.array(np.ones((2, 3)))[0, 1, 2]


def test_shape_manipulation_matches_numpy(sim):  # noqa: F811
    import cunumeric_b200 as cn

    rng = np.random.default_rng(9)
    for _ in range(60):
        shape = tuple(int(rng.integers(1, 5)) for _ in range(int(rng.integers(1, 5))))
        a = rng.integers(-50, 50, size=shape).astype(np.int64)
        A = cn.array(a)
        perm = tuple(int(p) for p in rng.permutation(len(shape)))
        assert np.array_equal(np.array(A.transpose(perm)), a.transpose(perm))
        assert np.array_equal(np.array(A.T), a.T)
        i, j = int(rng.integers(len(shape))), int(rng.integers(len(shape)))
        assert np.array_equal(np.array(A.swapaxes(i, j)), a.swapaxes(i, j))
        assert np.array_equal(np.array(A.ravel()), a.ravel())
        assert np.array_equal(np.array(A.T.flatten()), a.T.flatten())
        assert np.array_equal(np.array(A.reshape(-1)), a.reshape(-1))
        assert np.array_equal(np.array(A.T.reshape(a.size)), a.T.reshape(a.size))   # needs a copy
        assert np.array_equal(np.array(cn.squeeze(A)), np.squeeze(a))
        assert A.T.shape == a.T.shape and A.size == a.size and A.ndim == a.ndim
    with pytest.raises(ValueError):
        cn.array(np.arange(6)).reshape(4, 2)


def test_reduction_api_matches_numpy(sim):  # noqa: F811
    """ndarray.sum / prod / max / min / all / any (array.py:_perform_unary_reduction; reference
    array.py:4323-4418): axis (negative too) / keepdims / initial / out= on arrays and on stepped or
    reversed views; int64 data, so every association order gives the same bits."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(3)
    for _ in range(250):
        nd = int(rng.integers(1, 5))
        shape = tuple(int(rng.integers(1, 5)) for _ in range(nd))
        a = rng.integers(-5, 6, size=shape).astype(np.int64)
        A = cn.array(a)
        if rng.random() < 0.4:
            key = tuple(slice(None, None, int(rng.choice([1, -1, 2]))) for _ in range(nd))
            a, A = a[key], A[key]
        axis = None if rng.random() < 0.25 else int(rng.integers(-nd, nd))
        keep = bool(rng.random() < 0.5)
        name = ["sum", "prod", "max", "min", "all", "any"][rng.integers(6)]
        kw = {}
        if name in ("sum", "prod", "max", "min") and rng.random() < 0.3:
            kw["initial"] = int(rng.integers(-3, 4))
        exp = getattr(a, name)(axis=axis, keepdims=keep, **kw)
        g = np.array(getattr(A, name)(axis=axis, keepdims=keep, **kw))
        assert g.shape == np.shape(exp) and g.dtype == np.asarray(exp).dtype, (shape, name, axis, keep)
        assert np.array_equal(g, exp), (shape, name, axis, keep, kw)
        if name in ("sum", "max") and rng.random() < 0.3:
            out = cn.empty(np.shape(exp), dtype=np.int64)
            assert getattr(A, name)(axis=axis, keepdims=keep, out=out, **kw) is out
            assert np.array_equal(np.array(out), exp)
    A = cn.array(np.ones((3, 4), dtype=np.int64))
    with pytest.raises(ValueError):
        A.sum(axis=0, out=cn.empty((3,), dtype=np.int64))
    with pytest.raises((ValueError, np.exceptions.AxisError)):
        A.sum(axis=2)
    assert cn.array(np.array([True, False, True])).sum().dtype == np.int32   # array.py:3146 (bool -> int32)


def test_where_and_astype_dtypes_match_numpy(sim):  # noqa: F811
    """module.where (reference module.py:3493-3541, array.py:4455-4470: common type of the two
    branches, mask broadcast) and ndarray.astype for every dtype pair."""
    import warnings

    import cunumeric_b200 as cn

    rng = np.random.default_rng(0)
    for da in DTYPES:
        a = np.arange(6).reshape(2, 3).astype(da)
        for db in DTYPES:
            b = (np.arange(3) + 10).astype(db)
            m = rng.random((2, 3)) < 0.5
            exp = np.where(m, a, b)
            g = np.array(cn.where(cn.array(m), cn.array(a), cn.array(b)))
            assert g.dtype == exp.dtype and np.array_equal(g, exp), (np.dtype(da).name, np.dtype(db).name)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")      # complex -> real discards the imaginary part
                exp = a.astype(db)
                g = np.array(cn.array(a).astype(db))
            assert g.dtype == exp.dtype and np.array_equal(g, exp), (np.dtype(da).name, np.dtype(db).name)
        for sc in (2, 2.5, True, 1j):
            m = rng.random((2, 3)) < 0.5
            exp = np.where(m, a, sc)
            g = np.array(cn.where(cn.array(m), cn.array(a), sc))
            assert g.dtype == exp.dtype and np.array_equal(g, exp), (np.dtype(da).name, sc)
    x = cn.array(np.arange(4.0))
    with pytest.raises(TypeError):
        x.astype(np.int32, casting="safe")
    assert x.astype(np.float64, copy=False) is x
