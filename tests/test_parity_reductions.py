"""GPU parity for SCALAR_UNARY_RED and UNARY_RED against the oracle.

Bars: bit-exact for integer / boolean / arg-index results and for MAX/MIN; floating-point SUM-like
results within n*eps (relative to sum |x|), n = reduction length, eps of the accumulation dtype."""
import numpy as np
import pytest

from oracle import ref

import parity_utils as pu

pytestmark = pytest.mark.gpu


def thunk_reduce(op, a, axis=None, where=None, initial=None, args=None, keepdims=False):
    """Drive DeferredArray.unary_reduction exactly like ndarray._perform_unary_reduction does."""
    from cunumeric_b200.config import UnaryRedCode

    code = UnaryRedCode[op]
    vdt = ref.red_val_dtype(op, a.dtype)
    res_dt = np.dtype(np.int64) if op in ref.ARG_RED else vdt
    axes = tuple(range(a.ndim)) if axis is None else (axis % a.ndim,)
    out_shape = tuple(n for d, n in enumerate(a.shape) if d not in axes)
    out = pu.new_thunk(out_shape, res_dt)
    w = None if where is None else pu.to_device(np.broadcast_to(where, a.shape).copy())
    out.unary_reduction(code, pu.to_device(a), w, axis, axes, keepdims, args, initial)
    return out.__numpy_array__()


def python_prefill(op, dt):
    """The value the reference's Python layer pre-fills the output with (deferred.py:213-238):
    finite finfo min/max, not the Legion identity."""
    from cunumeric_b200.config import UnaryRedCode
    from cunumeric_b200.deferred import _UNARY_RED_IDENTITIES

    return _UNARY_RED_IDENTITIES[UnaryRedCode[op]](np.dtype(dt))


def check_reduction(op, a, got, exp_val, n, what):
    exp = exp_val["arg"] if op in ref.ARG_RED else exp_val
    exp = np.asarray(exp).reshape(np.shape(got))
    if op in ref.ARG_RED:
        assert got.dtype == np.int64
        assert np.array_equal(got, exp), f"{what}: {got} vs {exp}"
        return
    assert got.dtype == exp.dtype, what
    if exp.dtype.kind in "biu" or op in ("MAX", "MIN", "NANMAX", "NANMIN"):
        assert np.array_equal(got, exp, equal_nan=True), f"{what}: {got} vs {exp}"
        return
    # floating SUM/PROD-like: |got - exp| <= n * eps * scale
    acc_dt = a.dtype if a.dtype.kind != "c" else a.real.dtype
    eps = np.finfo(acc_dt).eps
    if op in ("PROD", "NANPROD"):
        scale = np.abs(exp.astype(np.complex128))
    else:
        mag = np.abs(np.nan_to_num(a.astype(np.complex128 if a.dtype.kind == "c" else np.float64)))
        if op in ("SUM_SQUARES", "VARIANCE"):
            mag = mag ** 2
        scale = mag.sum()
    err = np.abs(got.astype(np.complex128) - exp.astype(np.complex128))
    bound = 2 * n * eps * np.maximum(scale, np.finfo(np.float64).tiny)
    assert np.all(err <= bound), f"{what}: err {err.max()} > bound {np.min(bound)}"


RED_NO_CONTAINS = [o for o in ref.RED_OPS if o != "CONTAINS"]


def red_input(op, dt, shape, rng):
    n = int(np.prod(shape))
    if op in ("PROD", "NANPROD"):
        if dt.kind in "iu":
            a = rng.integers(1, 3, n).astype(dt)
        elif dt.kind == "b":
            a = rng.random(n) < 0.999
        else:
            a = rng.uniform(0.98, 1.02, n).astype(dt)
            if dt.kind == "c":
                a = (rng.uniform(0.98, 1.02, n) * np.exp(1j * rng.uniform(-0.01, 0.01, n))).astype(dt)
    elif op in ("ALL", "ANY", "COUNT_NONZERO"):
        a = pu.make_input(dt, n, rng, "small")
        a.reshape(-1)[rng.random(n) < 0.4] = 0
    elif dt == np.float16:
        a = rng.uniform(-1, 1, n).astype(dt)
    else:
        a = pu.make_input(dt, n, rng, "small")
    if op.startswith("NAN") and dt.kind in "fc":
        a.reshape(-1)[rng.random(n) < 0.1] = np.nan
    if op in ("ARGMAX", "ARGMIN", "NANARGMAX", "NANARGMIN", "MAX", "MIN") and dt.kind == "f":
        # tie-free data, as in tests/integration/test_arg_reduce.py
        pass
    return a.reshape(shape)


@pytest.mark.parametrize("op", ref.RED_OPS)
@pytest.mark.parametrize("dt", pu.DTYPES, ids=lambda d: d.name)
def test_scalar_reduction(op, dt):
    try:
        ref.red_identity(op, dt)
    except ref.InvalidOp:
        pytest.skip("reference marks this (op, dtype) invalid")
    rng = pu.rng_for("sred", op, dt.name)
    n = 70001 if dt != np.float16 else 3001
    a = red_input(op, dt, (n,), rng)
    args = None
    extra = None
    if op == "CONTAINS":
        extra = a[n // 2]
        args = (extra,)
    if op == "VARIANCE":
        extra = np.array(a.mean() if dt.kind in "fc" else 1).astype(dt)
        args = (extra,)
    prefill = python_prefill(op, dt)
    exp = ref.scalar_unary_red(op, a, extra=extra, initial=prefill)
    got = thunk_reduce(op, a, args=args)
    check_reduction(op, a, got, exp, n, f"scalar {op}/{dt.name}")


@pytest.mark.parametrize("op", ["SUM", "MAX", "ARGMAX", "ARGMIN", "ALL", "COUNT_NONZERO", "PROD"])
@pytest.mark.parametrize("shape,axis", [((37, 53), 0), ((37, 53), 1), ((5, 7, 9), 0), ((5, 7, 9), 1),
                                        ((5, 7, 9), 2), ((3, 4, 5, 6), 2), ((2000, 3), 0),
                                        ((3, 2000), 1), ((300, 260), 0), ((260, 4100), 1),
                                        ((1, 50), 0), ((50, 1), 1)])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.float16, np.bool_, np.uint8],
                         ids=lambda d: np.dtype(d).name)
def test_axis_reduction(op, shape, axis, dt):
    dt = np.dtype(dt)
    try:
        ref.red_identity(op, dt)
    except ref.InvalidOp:
        pytest.skip("invalid")
    rng = pu.rng_for("ared", op, dt.name, shape, axis)
    a = red_input(op, dt, shape, rng)
    exp = ref.unary_red(op, a, axis, initial=python_prefill(op, dt))
    got = thunk_reduce(op, a, axis=axis)
    check_reduction(op, a, got, exp, shape[axis], f"axis {op}/{dt.name} {shape}@{axis}")


@pytest.mark.parametrize("op", RED_NO_CONTAINS)
@pytest.mark.parametrize("dt", pu.DTYPES, ids=lambda d: d.name)
def test_axis_reduction_all_pairs(op, dt):
    try:
        ref.red_identity(op, dt)
    except ref.InvalidOp:
        pytest.skip("invalid")
    for shape, axis in (((41, 67), 0), ((41, 67), 1)):
        rng = pu.rng_for("ared-all", op, dt.name, shape, axis)
        a = red_input(op, dt, shape, rng)
        exp = ref.unary_red(op, a, axis, initial=python_prefill(op, dt))
        got = thunk_reduce(op, a, axis=axis)
        check_reduction(op, a, got, exp, shape[axis], f"axis {op}/{dt.name} {shape}@{axis}")


# Long contiguous rows take the bulk-copy (TMA) pipeline of axis_red.inl when the row is >= 16 KB,
# 16-byte aligned and there are >= SMs/2 rows; the shapes below hit full chunks, ragged last
# chunks, several chunks per row, more rows than CTAs, and 3-D kept dims.
LONG_ROWS = [((80, 4096), 1, np.float32), ((150, 5004), 1, np.float32), ((333, 16388), 1, np.float32),
             ((76, 2048 + 6), 1, np.float64), ((90, 8192 + 8), 1, np.float16), ((100, 20000), 1, np.int8),
             ((80, 1030), 1, np.complex128), ((75, 33000), 1, np.uint8), ((5, 20, 6000), 2, np.int32),
             ((160, 4100), 1, np.int64), ((200, 16384 + 16), 1, np.bool_), ((600, 4096), 1, np.bool_)]


@pytest.mark.parametrize("op", ["SUM", "MAX", "MIN", "ARGMAX", "ARGMIN", "PROD", "ALL", "ANY",
                                "COUNT_NONZERO", "NANSUM", "NANARGMAX", "SUM_SQUARES"])
@pytest.mark.parametrize("shape,axis,dt", LONG_ROWS,
                         ids=lambda v: np.dtype(v).name if isinstance(v, type) else str(v))
def test_axis_reduction_long_rows(op, shape, axis, dt):
    dt = np.dtype(dt)
    try:
        ref.red_identity(op, dt)
    except ref.InvalidOp:
        pytest.skip("invalid")
    rng = pu.rng_for("ared-long", op, dt.name, shape, axis)
    a = red_input(op, dt, shape, rng)
    if op in ("ARGMAX", "ARGMIN", "NANARGMAX") and dt.kind in "biu":
        # many ties: the first occurrence must win across threads, warps and chunks
        a = (a.astype(np.int64) % 3).astype(dt)
    exp = ref.unary_red(op, a, axis, initial=python_prefill(op, dt))
    got = thunk_reduce(op, a, axis=axis)
    check_reduction(op, a, got, exp, shape[axis], f"long rows {op}/{dt.name} {shape}@{axis}")


def test_axis_reduction_long_rows_views():
    """pitched rows (16-byte aligned -> bulk-copy path), misaligned rows (-> LDG fallback), and the
    extreme at the start / end of a row and on chunk boundaries"""
    import cunumeric_b200 as cn

    rng = pu.rng_for("ared-long-views")
    a = rng.normal(size=(96, 3 * 4096 + 40)).astype(np.float32)
    A = cn.array(a)
    for sl in (np.s_[:, 8:8 + 8192], np.s_[:, 4:4 + 8192], np.s_[3:83, 16:], np.s_[:, :4096],
               np.s_[::2, 12:12 + 4100]):
        v, V = a[sl], A[sl]
        assert np.array_equal(V.max(axis=1).__array__(), v.max(axis=1)), sl
        assert np.array_equal(V.argmin(axis=1).__array__(), v.argmin(axis=1)), sl
        got = V.sum(axis=1).__array__()
        exp = ref.unary_red("SUM", np.ascontiguousarray(v), 1, initial=np.float32(0))
        assert np.allclose(got, exp, rtol=0, atol=v.shape[1] * np.finfo(np.float32).eps *
                           np.abs(v).sum(axis=1).max()), sl
    b = np.zeros((80, 12288), dtype=np.float32)
    for pos in (0, 1, 1023, 1024, 4095, 4096, 4097, 8191, 8192, 12287):
        b[:] = 0
        b[np.arange(80), pos] = 5.0
        b[np.arange(80), min(pos + 4096, 12287)] = 5.0  # a later tie must lose
        assert np.array_equal(cn.array(b).argmax(axis=1).__array__(), np.full(80, pos)), pos
    # pre-filled output is folded in, not overwritten (reduce-accessor semantics)
    assert np.array_equal(cn.array(b).max(axis=1, initial=7.0).__array__(), np.full(80, 7.0, np.float32))


def test_reductions_on_views_and_where():
    import cunumeric_b200 as cn

    rng = pu.rng_for("redviews")
    a = rng.normal(size=(64, 96)).astype(np.float64)
    A = cn.array(a)
    # transposed input: the kernel mode follows the memory layout
    for axis in (0, 1):
        got = A.T.sum(axis=axis).__array__()
        exp = ref.unary_red("SUM", a.T, axis, initial=0.0)
        assert np.allclose(got, exp, rtol=1e-13)
        got = A[3:-5, 2::3].max(axis=axis).__array__()
        assert np.array_equal(got, a[3:-5, 2::3].max(axis=axis))
        assert np.array_equal(A[::2].argmax(axis=axis).__array__(), a[::2].argmax(axis=axis))
    assert int(A[5:, 7:].argmin()) == int(a[5:, 7:].argmin())
    # where masks (tests/integration/test_reduction.py:154-158)
    x = cn.array(np.array([[1, 2], [3, 4]]))
    assert int(x.sum(where=np.array([False, True]))) == 6
    w = rng.random(a.shape) < 0.5
    assert np.allclose(A.sum(where=cn.array(w)).__array__(), a.sum(where=w))
    assert np.allclose(A.sum(axis=0, where=cn.array(w)).__array__(), a.sum(axis=0, where=w))
    assert np.allclose(A.sum(axis=1, where=cn.array(w)).__array__(), a.sum(axis=1, where=w))


def test_arg_reduction_ties_take_first_occurrence():
    import cunumeric_b200 as cn

    a = np.zeros(100000, dtype=np.float32)
    a[[17, 4099, 65000]] = 7.0
    assert int(cn.array(a).argmax()) == 17
    assert int(ref.scalar_unary_red("ARGMAX", a)["arg"]) == 17
    b = np.zeros((3000, 5), dtype=np.int32)
    b[[5, 2500], :] = 9
    assert np.array_equal(cn.array(b).argmax(axis=0).__array__(), np.full(5, 5))


def test_global_index_of_partitioned_rect():
    """Arg-reductions return GLOBAL flat indices when the rect is a tile of a larger array
    (unary_red_util.h:342-351): the C ABI takes origin + global shape."""
    import ctypes

    from cunumeric_b200 import _lib
    from cunumeric_b200.config import UnaryRedCode, argval_dtype
    from cunumeric_b200.runtime import runtime

    rng = pu.rng_for("global-index")
    full = rng.normal(size=(40, 30)).astype(np.float32)
    lo, hi = 12, 29
    tile = np.ascontiguousarray(full[lo:hi])
    exp = ref.scalar_unary_red("ARGMAX", tile, origin=(lo, 0), shape=full.shape)
    d_in = pu.to_device(tile)
    out = pu.new_thunk((1,), argval_dtype(np.float32))
    out.fill(np.array((np.iinfo(np.int64).min, -np.inf), dtype=out.dtype))
    origin = (ctypes.c_int64 * 2)(lo, 0)
    gshape = (ctypes.c_int64 * 2)(*full.shape)
    di, do = d_in.base.descriptor(), out.base.descriptor()
    _lib.check(runtime.lib.cnb_scalar_unary_red(int(UnaryRedCode.ARGMAX), ctypes.byref(do),
                                                ctypes.byref(di), None, origin, gshape, None,
                                                runtime.stream))
    got = out.__numpy_array__()[0]
    assert int(got["arg"]) == int(exp["arg"]) == lo * 30 + int(np.argmax(tile))


def test_fp16_sums_accumulate_in_fp32_a_documented_deviation():
    """DELIBERATE DEVIATION from the reference, stated here and in DESIGN.md §3.2: the reference folds
    fp16 SUM / PROD in fp16 (`unary_red_util.h:250-271`: VAL = __half, one rounding per addition); the
    CUDA path carries fp32 partials and rounds once at the end.  Both are inside the n * eps(fp16)
    contract, but they are NOT bit-identical: the sequential fp16 fold stalls once the running sum
    exceeds 2048 * |x| (every further addend is rounded away), the fp32 accumulation does not.  This
    test pins the behaviour: (1) the result is within n * eps of the oracle, (2) it is at least as
    close to the exact (float64) sum as the reference's own result."""
    rng = pu.rng_for("fp16-sum-deviation")
    a = rng.uniform(0.5, 1.5, 6000).astype(np.float16)    # exact sum ~6000: the fp16 fold saturates
    exp = ref.scalar_unary_red("SUM", a)
    got = thunk_reduce("SUM", a)
    exact = a.astype(np.float64).sum()
    n_eps = a.size * float(np.finfo(np.float16).eps)
    assert abs(float(got.reshape(())) - float(exp)) <= n_eps * exact
    assert abs(float(got.reshape(())) - exact) <= abs(float(exp) - exact) + float(np.spacing(np.float16(exact)))
    # axis reduction, same rule
    m = rng.uniform(0.5, 1.5, (3000, 8)).astype(np.float16)
    got0 = thunk_reduce("SUM", m, axis=0).astype(np.float64)
    exp0 = ref.unary_red("SUM", m, 0).astype(np.float64)
    exact0 = m.astype(np.float64).sum(axis=0)
    assert np.all(np.abs(got0 - exp0) <= m.shape[0] * float(np.finfo(np.float16).eps) * exact0)
    assert np.all(np.abs(got0 - exact0) <= np.abs(exp0 - exact0) + np.spacing(exact0.astype(np.float16)).astype(np.float64))


def test_more_scalar_reductions_in_flight_than_scratch_slots():
    """cnb_scalar_unary_red hands out 64 rotating {partials, ticket} scratch slots (INTEGRATION.md
    contract notes).  On ONE stream reductions are serialised, so re-using a slot after 64 launches is
    safe: 200 back-to-back reductions of different arrays, no synchronisation in between, all
    correct."""
    rng = pu.rng_for("slots")
    arrays = [rng.integers(-1000, 1000, 50000 + 17 * i).astype(np.int64) for i in range(200)]
    dev = [pu.to_device(a) for a in arrays]
    outs = []
    from cunumeric_b200.config import UnaryRedCode

    for d in dev:
        out = pu.new_thunk((), np.int64)
        out.unary_reduction(UnaryRedCode.SUM, d, None, None, (0,), False, None, None)
        outs.append(out)
    for a, o in zip(arrays, outs):
        assert int(o.__numpy_array__()) == int(a.sum())


@pytest.mark.parametrize("op", ["SUM", "PROD", "MAX", "MIN", "NANSUM", "NANMAX", "ALL", "ANY",
                                "COUNT_NONZERO", "ARGMAX"])
@pytest.mark.parametrize("shape,dt", [((3, 1 << 23), np.float32), ((2, 3, (1 << 22) + 77), np.float64),
                                      ((5, (1 << 25) + 3), np.int8), ((2, 1 << 22), np.complex64)],
                         ids=["3xf32", "2x3xf64", "5xi8", "2xc64"])
def test_few_very_long_rows_are_split_across_ctas(op, shape, dt):
    """ROW mode with fewer outputs than SMs / 2: value reductions run in two stages —
    S segments per row into a scratch of partials, then the scratch along S — so the whole chip works
    on 3 rows; arg-reductions (partials would be Argvals) keep one CTA per row.  Same bars as the
    rest of the file, against the oracle."""
    dt = np.dtype(dt)
    if ref.red_val_dtype(op, dt) is None:
        pytest.skip("invalid pair")
    try:
        ref.red_identity(op, dt)
    except ref.InvalidOp:
        pytest.skip("reference marks this (op, dtype) invalid")
    rng = pu.rng_for("long-rows", op, dt.name, shape)
    if op in ("PROD", "NANPROD"):
        a = np.ones(shape, dtype=dt)
        idx = rng.integers(0, shape[-1], 40)
        a[..., idx] = np.asarray(rng.uniform(0.5, 2.0, 40)).astype(dt)
    elif dt.kind == "c":
        a = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(dt)
    elif dt.kind == "i":
        a = rng.integers(-5, 6, size=shape).astype(dt)
    else:
        a = rng.normal(size=shape).astype(dt)
    if op.startswith("NAN") and dt.kind == "f":
        a[..., ::1001] = np.nan
    if op == "ALL":
        a[a == 0] = 1
        a[0, ..., shape[-1] // 2] = 0        # exactly one zero, in the middle segment of row 0
    axis = len(shape) - 1
    with np.errstate(all="ignore"):
        exp = ref.unary_red(op, a, axis)
    got = thunk_reduce(op, a, axis=axis)
    check_reduction(op, a, got, exp, shape[-1], f"{op}/{dt.name}/{shape}")


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("shape,axes", [((5, 7, 9), (0, 2)), ((5, 7, 9), (1, 2)), ((3, 4, 5, 6), (0, 1, 3)),
                                        ((64, 33, 130), (0, 1)), ((2, 300, 2, 70), (1, 3))])
def test_multi_axis_reductions(shape, axes, dt):
    """Several axes at once (ndarray._perform_unary_reduction; the reference raises,
    deferred.py:3259-3262): one UNARY_RED per axis, innermost first.  Expected = the oracle's UNARY_RED
    applied in the same order; exact for integers / MAX / MIN, n * eps for floating sums."""
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    rng = pu.rng_for("multi-axis", dt.name, shape, axes)
    a = red_input("SUM", dt, shape, rng)
    A = cn.array(a)
    n = int(np.prod([shape[x] for x in axes]))

    def oracle(op, follow):
        cur = a
        for i, ax in enumerate(sorted(axes, reverse=True)):
            cur = ref.unary_red(op if i == 0 else follow, cur, ax)
        return cur

    for op, fn in (("MAX", lambda x: x.max(axis=axes)), ("MIN", lambda x: x.min(axis=axes))):
        assert np.array_equal(fn(A).__array__(), oracle(op, op)), op
    got, exp = A.sum(axis=axes).__array__(), oracle("SUM", "SUM")
    assert got.dtype == exp.dtype and got.shape == exp.shape
    if dt.kind in "iu":
        assert np.array_equal(got, exp)
    else:
        bound = n * np.finfo(dt).eps * np.abs(a).sum(axis=axes, dtype=np.float64)
        assert np.all(np.abs(got.astype(np.float64) - exp.astype(np.float64)) <= bound)
    assert np.array_equal(cn.count_nonzero(A, axis=axes).__array__(),
                          np.count_nonzero(a, axis=axes))
    assert np.array_equal(A.sum(axis=axes, keepdims=True).__array__().shape,
                          a.sum(axis=axes, keepdims=True).shape)


PEEL_VIEWS = [
    ((1030, 1104), np.s_[1:-1, 1:-1]),        # head + body + tail
    ((1030, 1104), np.s_[:, 3:]),             # head only
    ((1030, 1104), np.s_[2:, :-5]),           # tail only
    ((4, 515, 1104), np.s_[:, 1:-1, 1:-7]),   # 3-D, two kept dims
]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8, np.float16, np.int64],
                         ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("case", range(len(PEEL_VIEWS)))
def test_reductions_of_misaligned_pitched_views(case, dt):
    """Views whose rows keep a 16-byte pitch but start / end inside a 16-byte word (x[1:-1, 1:-1]) are
    reduced as head | body | tail (capi.cu try_peel_axis / try_peel_scalar): every axis and the full
    reduction against the oracle on a contiguous copy of the view, arg-reductions with their global
    indices, ties across the pieces resolved to the first occurrence."""
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    shape, sl = PEEL_VIEWS[case]
    rng = pu.rng_for("peel", dt.name, case)
    for op in ("SUM", "MAX", "MIN", "ARGMAX", "ARGMIN", "COUNT_NONZERO", "ANY"):
        base = red_input(op, dt, shape, rng)
        view = np.ascontiguousarray(base[sl])
        assert view.size >= 1 << 20
        V = cn.array(base)[sl]
        n_eps = view.size
        for axis in list(range(view.ndim)) + [None]:
            if axis is None and op == "SUM" and dt == np.float16:
                continue   # 10^6 terms accumulated in fp16 by the oracle: not a usable expectation
            if axis is None:
                exp = ref.scalar_unary_red(op, view, initial=python_prefill(op, dt))
            else:
                exp = ref.unary_red(op, view, axis, initial=python_prefill(op, dt))
            if op in ref.ARG_RED:
                got = (V.argmax(axis=axis) if op == "ARGMAX" else V.argmin(axis=axis)).__array__()
            elif op == "COUNT_NONZERO":
                got = cn.count_nonzero(V, axis=axis).__array__()
                exp = np.count_nonzero(view, axis=axis)
                assert np.array_equal(got, exp), (op, axis)
                continue
            elif op == "ANY":
                got = V.any(axis=axis).__array__()
            else:
                got = getattr(V, op.lower())(axis=axis).__array__()
            check_reduction(op, view, np.asarray(got), exp, n_eps if axis is None else view.shape[axis],
                            f"peel {op}/{dt.name} case {case} axis {axis}")
    # ties: the same extreme in the head, the body and the tail of a row -> first occurrence wins
    t = np.zeros(shape, dtype=dt)
    tv = t[sl]
    tv[..., 0] = 3
    tv[..., tv.shape[-1] // 2] = 3
    tv[..., -1] = 3
    T = cn.array(t)[sl]
    assert np.array_equal(T.argmax(axis=tv.ndim - 1).__array__(), np.zeros(tv.shape[:-1], np.int64))
    assert int(T.argmax()) == 0
    tv[..., 0] = 0
    T = cn.array(t)[sl]
    assert np.array_equal(T.argmax(axis=tv.ndim - 1).__array__(),
                          np.full(tv.shape[:-1], tv.shape[-1] // 2, np.int64))
    assert int(T.argmax()) == tv.shape[-1] // 2
    # a pre-filled output is folded in once, not once per piece
    if dt.kind == "f":
        ones = cn.array(np.ones(shape, dtype=dt))[sl]
        got = ones.sum(axis=ones.ndim - 1, initial=5).__array__()
        assert np.array_equal(got, np.full(tv.shape[:-1], tv.shape[-1] + 5, dtype=dt)) or dt == np.float16
