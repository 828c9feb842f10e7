"""GPU tests of the NumPy-level API (ndarray / ufunc objects / module functions), written like the
reference's own integration tests: differential against NumPy on the same inputs
(tests/integration/utils/comparisons.py allclose rtol=1e-5, atol=1e-8, check_dtype=True)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def allclose(a, b, rtol=1e-5, atol=1e-8, equal_nan=False, check_dtype=True):
    a, b = np.asarray(a), np.asarray(b)
    if check_dtype and a.dtype != b.dtype:
        return False
    return np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=equal_nan)


def test_jacobi_matches_numpy():
    """tests/integration/test_jacobi.py:23-57."""
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import stencil_init, stencil_run

    g, g_np = stencil_init(8, np.float32, xp=cn), stencil_init(8, np.float32, xp=np)
    w, w_np = stencil_run(g, 2), stencil_run(g_np, 2)
    assert allclose(w, w_np)
    assert allclose(cn.sum(cn.absolute(w - g[1:-1, 1:-1])), np.sum(np.absolute(w_np - g_np[1:-1, 1:-1])))


def test_black_scholes_matches_oracle_and_numpy():
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs
    from oracle import refnp

    from cunumeric_b200 import fusion

    S, X, T = black_scholes_inputs(50000, np.float32, seed=3)
    old = fusion.set_mode("0")  # op-by-op: one kernel per task, scalars converted on the host
    try:
        dS, dX, dT = cn.array(S), cn.array(X), cn.array(T)
        launches = cn.runtime.launch_count()
        c, p = black_scholes(dS, dX, dT, 0.02, 0.3)
        assert cn.runtime.launch_count() - launches == 63
        fusion.set_mode("always")  # the same program as ONE fused kernel, bit-identical
        launches = cn.runtime.launch_count()
        cf, pf = black_scholes(dS, dX, dT, 0.02, 0.3)
        cn.flush()
        assert cn.runtime.launch_count() - launches == 1
        assert np.array_equal(cf.__array__(), c.__array__()) and np.array_equal(pf.__array__(), p.__array__())
    finally:
        fusion.set_mode(old)
    assert c.dtype == np.float32 and p.dtype == np.float32
    co, po = black_scholes(refnp.array(S), refnp.array(X), refnp.array(T), 0.02, 0.3, xp=refnp)
    cn_, pn = black_scholes(S, X, T, 0.02, 0.3, xp=np)
    for got, ora, npy in ((c, co.a, cn_), (p, po.a, pn)):
        got = got.__array__()
        assert np.max(np.abs(got - ora) / (np.abs(ora) + 1)) < 1e-5  # 63 chained ops
        assert np.allclose(got, npy, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("op", ["add", "subtract", "multiply", "true_divide", "maximum", "minimum",
                                "equal", "less", "logical_and", "power", "floor_divide", "remainder"])
def test_binary_ufunc_result_dtypes(op):
    """tests/integration/test_binary_ufunc.py:96-231: values AND result dtypes, arrays x arrays and
    arrays x scalars."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(11)
    arrs = [rng.integers(1, 5, (4, 5)).astype("I"), rng.random((4, 5)).astype("e") + 1,
            rng.random((4, 5)).astype("f") + 1, rng.random((4, 5)) + 1]
    scalars = [2, 1.5, np.float32(2.5)]
    for a in arrs:
        for b in arrs + scalars:
            if op in ("floor_divide", "remainder", "power") and (a.dtype == np.float16 or getattr(b, "dtype", None) == np.float16):
                continue
            with np.errstate(all="ignore"):
                exp = getattr(np, op)(a, b)
            A = cn.array(a)
            B = cn.array(b) if isinstance(b, np.ndarray) else b
            got = getattr(cn, op)(A, B)
            rtol = 1e-2 if exp.dtype == np.float16 or a.dtype == np.float16 else 1e-5
            assert got.dtype == exp.dtype, (op, a.dtype, getattr(b, "dtype", type(b)), got.dtype, exp.dtype)
            assert np.allclose(got.__array__(), exp, rtol=rtol, atol=1e-6), (op, a.dtype, b)


def test_unary_ufuncs_out_and_dtype_kwargs():
    import cunumeric_b200 as cn

    a = np.linspace(0.1, 2.0, 37)
    A = cn.array(a)
    out = cn.empty(a.shape)
    r = cn.exp(A, out=out)
    assert r is out and allclose(out, np.exp(a))
    out32 = cn.empty(a.shape, dtype=np.float32)
    with pytest.raises(TypeError):
        cn.exp(A, out=out32, casting="safe")
    cn.exp(A, out=out32)  # same_kind: allowed, converts on the way out
    assert allclose(out32, np.exp(a).astype(np.float32))
    host_out = np.empty(a.shape)
    cn.sqrt(A, out=host_out)  # NumPy array as `out` (ufunc.py:262-272)
    assert np.allclose(host_out, np.sqrt(a))
    assert cn.sqrt(cn.array(np.arange(5))).dtype == np.float64  # ints promote to the first float sig
    assert cn.negative(cn.array(np.arange(5, dtype=np.int8))).dtype == np.int8
    with pytest.raises(NotImplementedError):
        cn.exp(A, where=cn.array(a > 1))
    m, e = cn.frexp(A)
    mn, en = np.frexp(a)
    assert allclose(m, mn) and np.array_equal(e.__array__(), en)
    assert allclose(cn.add.reduce(A), np.add.reduce(a)) and allclose(cn.maximum.reduce(A), a.max())
    assert bool(cn.logical_and.reduce(cn.array(a > 0)))


def test_astype_and_casting_errors():
    """tests/integration/test_astype.py:21-121."""
    import cunumeric_b200 as cn

    tv = np.array([0, 0, 1, 2, 3, 0, 1, 2, 3])
    for s in "?bBhHiIlLefd":
        for d in "?bBhHiIlLefdFD":
            a = tv.astype(s)
            assert np.array_equal(cn.array(a).astype(d).__array__(), a.astype(d)), (s, d)
    with pytest.raises(TypeError):
        cn.array(tv.astype("d")).astype("i", casting="safe")
    c = cn.array(tv.astype("D") + 1j)
    with pytest.warns(np.exceptions.ComplexWarning):
        assert np.array_equal(c.astype("d").__array__(), tv.astype("d"))


def test_where_fixtures():
    """tests/integration/test_where.py:22-125."""
    import cunumeric_b200 as cn

    x, y = np.array([[1, 2], [3, 4]]), np.array([[9, 8], [7, 6]])
    for cond in ([[True, False], [True, True]], [[True, False]], [True, False], [False, True]):
        assert np.array_equal(cn.where(cn.array(cond), cn.array(x), cn.array(y)).__array__(),
                              np.where(cond, x, y))
    a = np.arange(10, dtype=np.int32)
    r = cn.where(cn.array(a > 4), cn.array(a), cn.array(a.astype(np.float32) * 0.5))
    assert r.dtype == np.float64 and np.array_equal(r.__array__(), np.where(a > 4, a, a.astype(np.float32) * 0.5))
    f = np.linspace(0, 1, 11, dtype=np.float32)
    assert cn.where(cn.array(f > 0.5), cn.array(f), 0.0).dtype == np.float32


def test_reduction_api():
    """tests/integration/test_reduction.py, test_amax_amin.py, test_arg_reduce.py,
    test_logical_reduction.py."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(7)
    for dt in ("l", "L", "f", "d", "F", "D"):
        a = (rng.random((5, 5, 5)) * 10).astype(dt)
        A = cn.array(a)
        assert allclose(A.sum(), a.sum(), check_dtype=False)
        for axis in range(-2, 3):
            assert allclose(A.sum(axis=axis), a.sum(axis=axis))
            assert allclose(A.sum(axis=axis, keepdims=True), a.sum(axis=axis, keepdims=True))
            if dt != "D":  # the reference marks PROD invalid for complex128 (unary_red_util.h:229)
                assert allclose(A.prod(axis=axis), a.prod(axis=axis), rtol=1e-4)
            if dt not in "FD":
                assert np.array_equal(A.max(axis=axis).__array__(), a.max(axis=axis))
                assert np.array_equal(A.argmin(axis=axis).__array__(), a.argmin(axis=axis))
                assert np.array_equal(A.argmax(axis=axis, keepdims=True).__array__(),
                                      a.argmax(axis=axis, keepdims=True))
    a = rng.random((6, 7))
    A = cn.array(a)
    assert allclose(A.sum(initial=5.0), a.sum(initial=5.0), check_dtype=False)
    assert allclose(A.max(axis=0, initial=0.5), a.max(axis=0, initial=0.5))
    out = cn.empty((7,))
    assert A.sum(axis=0, out=out) is out and allclose(out, a.sum(axis=0))
    assert allclose(A.sum(dtype=np.float32), a.sum(dtype=np.float32), check_dtype=False)
    assert cn.array(np.array([True, False])).sum().dtype == np.int32  # bool -> int32 first
    assert allclose(A.mean(axis=1), a.mean(axis=1)) and allclose(A.mean(), a.mean(), check_dtype=False)
    # several axes: the reference raises (deferred.py:3259-3262); here the separable reductions run
    # one UNARY_RED per axis
    t = rng.random((2, 3, 4))
    T = cn.array(t)
    assert allclose(T.sum(axis=(0, 1)), t.sum(axis=(0, 1)))
    assert allclose(T.sum(axis=(0, 2), keepdims=True), t.sum(axis=(0, 2), keepdims=True))
    assert np.array_equal(T.max(axis=(1, 2)).__array__(), t.max(axis=(1, 2)))
    assert np.array_equal(T.min(axis=(-1, 0)).__array__(), t.min(axis=(-1, 0)))
    assert np.array_equal((T > 0.5).any(axis=(0, 1)).__array__(), (t > 0.5).any(axis=(0, 1)))
    assert np.array_equal(cn.count_nonzero(T > 0.5, axis=(1, 2)).__array__(),
                          np.count_nonzero(t > 0.5, axis=(1, 2)))
    with pytest.raises(ValueError):
        T.argmax(axis=(0, 1))   # "axis must be an integer" (array.py:3466 in the reference)
    with pytest.raises(NotImplementedError):
        cn.array(a + 1j).max()
    assert cn.zeros((0,)).sum().__array__() == 0.0
    assert int(cn.count_nonzero(cn.array(a > 0.5))) == np.count_nonzero(a > 0.5)
    b = a.copy()
    b[2, 3] = np.nan
    B = cn.array(b)
    assert allclose(cn.nansum(B), np.nansum(b), check_dtype=False)
    assert allclose(cn.nanmax(B, axis=0), np.nanmax(b, axis=0))
    assert int(cn.nanargmin(B)) == np.nanargmin(b)


def test_views_and_setitem():
    import cunumeric_b200 as cn

    a = np.arange(60, dtype=np.float64).reshape(3, 4, 5)
    A = cn.array(a)
    for key in [(1,), (slice(None), 2), (Ellipsis, 3), (slice(0, 2), slice(1, 3), slice(None, None, 2)),
                (None, 1, slice(None), None), (-1, -2, -3)]:
        assert np.array_equal(A[key].__array__(), a[key]), key
    A[1, :, 2] = -1.0
    a[1, :, 2] = -1.0
    A[0] = A[2]
    a[0] = a[2]
    A[:, 1:3, :] = cn.array(np.ones((2, 5)))
    a[:, 1:3, :] = np.ones((2, 5))
    assert np.array_equal(A.__array__(), a)
    assert np.array_equal(A.T.__array__(), a.T) and np.array_equal(A.swapaxes(0, 2).__array__(), a.swapaxes(0, 2))
    assert np.array_equal(A.reshape(12, 5).__array__(), a.reshape(12, 5))
    assert np.array_equal(A.T.reshape(-1).__array__(), a.T.reshape(-1))
    c = cn.array(a[..., 0] + 1j * a[..., 1])
    assert np.array_equal(c.real.__array__(), a[..., 0]) and np.array_equal(c.imag.__array__(), a[..., 1])
    with pytest.raises(NotImplementedError):
        A[[0, 1]]
    with pytest.raises(IndexError):
        A[5]
    assert np.array_equal(cn.clip(A, 3, 20).__array__(), np.clip(a, 3, 20))


def test_async_copies_and_map_chunks():
    import cunumeric_b200 as cn

    n = 300_000
    rng = np.random.default_rng(2)
    xh, yh = cn.pinned_empty(n, np.float32), cn.pinned_empty(n, np.float32)
    xh[...] = rng.random(n, dtype=np.float32)
    yh[...] = rng.random(n, dtype=np.float32)
    X = cn.from_host(xh, blocking=False)
    Y = cn.from_host(yh, blocking=False)
    fut = (X * Y + 1.0).to_host(blocking=False)
    assert np.array_equal(fut.wait(), xh * yh + np.float32(1.0))
    o1, o2 = cn.pinned_empty(n, np.float32), cn.pinned_empty(n, np.float32)
    cn.map_chunks(lambda a, b: (a + b, cn.sqrt(a) * b), (xh, yh), (o1, o2), chunk=70_001)
    assert np.array_equal(o1, xh + yh)
    assert np.allclose(o2, np.sqrt(xh) * yh, rtol=1e-6)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=lambda d: np.dtype(d).name)
def test_var_matches_numpy(dt):
    """array.py:3234-3323 / tests/integration/test_stats.py: VARIANCE as one scalar reduction,
    `x - mu` + SUM_SQUARES along an axis, ddof, keepdims."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(21)
    a = (rng.normal(size=(37, 53)) * 5 + 3).astype(dt)
    A = cn.array(a)
    tol = dict(rtol=2e-5, atol=1e-6) if dt == np.float32 else dict(rtol=1e-12, atol=1e-12)
    assert np.allclose(float(A.var()), a.var(), **tol)
    assert np.allclose(float(cn.var(A, ddof=1)), a.var(ddof=1), **tol)
    for axis in (0, 1):
        got = A.var(axis=axis).__array__()
        assert got.shape == a.var(axis=axis).shape
        assert np.allclose(got, a.var(axis=axis), **tol)
        got = A.var(axis=axis, ddof=1, keepdims=True).__array__()
        assert np.allclose(got, a.var(axis=axis, ddof=1, keepdims=True), **tol)
    v = cn.array(a[:, 0].copy())
    assert np.allclose(float(v.var(axis=0)), a[:, 0].var(), **tol)  # 1-D: the scalar-reduction path
    with pytest.raises(NotImplementedError):
        A.var(axis=(0, 1))


@pytest.mark.parametrize("mode", ["0", "always"], ids=["op-by-op", "fused"])
def test_stencil_config_c1_matches_the_reference_arithmetic(mode):
    """BASELINE.json configs[0]: examples/stencil.py, fp64, N=1000, 100 iterations — the reference's
    own CPU-runnable case.  Every task is an IEEE add / multiply / copy, so the result must equal
    op-by-op NumPy (= the reference functors, tests/test_oracle.py::test_jacobi...) bit for bit."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from cunumeric_b200.workloads import stencil_init, stencil_run

    old = fusion.set_mode(mode)
    try:
        g = stencil_init(1000, np.float64, xp=cn)
        w = stencil_run(g, 100)
        got_g, got_w = g.__array__(), w.__array__()
    finally:
        fusion.set_mode(old)
    g_np = stencil_init(1000, np.float64, xp=np)
    w_np = stencil_run(g_np, 100)
    assert np.array_equal(got_w, w_np) and np.array_equal(got_g, g_np)


# ------------------------------------------------------------------ advisor findings (round 1)
@pytest.mark.gpu
def test_user_zero_d_arrays_do_not_alias_the_constant_cache():
    import cunumeric_b200 as cn

    """array.py `convert_to_cunumeric_ndarray`: a user-visible 0-d array owns its buffer; in-place
    writes to it must not change the cached constants later `arr * 2.0` operations read."""
    x = cn.array(2.0)
    x += 1
    assert float(x) == 3.0
    assert (np.array(cn.ones(4) * 2.0) == 2).all()
    acc = cn.array(0.0)
    acc += cn.ones(8).sum()
    assert float(acc) == 8.0
    assert (np.array(cn.ones(4) + 0.0) == 1).all()
    y = cn.asarray(np.float64(5.0))
    y.fill(7.0)
    assert float(y) == 7.0 and float(y.astype(np.float32)) == 7.0
    assert (np.array(cn.ones(3) * 5.0) == 5).all()
    cn.add(cn.ones(()), 1.0, out=x)
    assert float(x) == 2.0 and (np.array(cn.full(3, 4.0) / 2.0) == 2).all()


@pytest.mark.gpu
def test_negative_zero_scalar_is_not_confused_with_positive_zero():
    import cunumeric_b200 as cn

    a = cn.array(np.array([1.0, -2.0, 3.0], dtype=np.float32))
    assert np.array_equal(np.array(a + 0.0), np.array([1.0, -2.0, 3.0], dtype=np.float32))
    with np.errstate(divide="ignore"):
        got = np.array(a / -0.0)
        exp = np.array([1.0, -2.0, 3.0], dtype=np.float32) / np.float32(-0.0)
    assert np.array_equal(got, exp)
    assert np.array_equal(np.signbit(np.array(a * -0.0)), [True, False, True])
    assert np.array_equal(np.signbit(np.array(a * 0.0)), [False, True, False])


@pytest.mark.gpu
def test_deepcopy_and_pickle_go_through_device_copy_and_host_array():
    import cunumeric_b200 as cn

    import copy
    import pickle

    a = cn.array(np.arange(12.0).reshape(3, 4))
    b = copy.deepcopy(a)
    b += 1
    assert np.array_equal(np.array(a), np.arange(12.0).reshape(3, 4))
    assert np.array_equal(np.array(b), np.arange(12.0).reshape(3, 4) + 1)
    c = pickle.loads(pickle.dumps(a))
    assert np.array_equal(np.array(c), np.array(a))
    with pytest.raises(TypeError):
        copy.deepcopy(a._thunk.base)
