"""CPU tests of the oracle itself (no GPU): it must reproduce every known-answer vector the
reference's own test-suite holds for this path, agree with NumPy the way the
reference's differential tests demand, and match the committed golden fixtures bit for bit."""
import os

import numpy as np
import pytest

from oracle import ref, refnp

import parity_utils as pu

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (make -C oracle)")

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz")


# ---- known-answer vectors of the reference tests ------------------------------------------------
def test_astype_test_vector():
    """tests/integration/test_astype.py:21 TEST_VECTOR x all dtype pairs, exact."""
    tv = [0, 0, 1, 2, 3, 0, 1, 2, 3]
    for s in ref.DTYPES:
        for d in ref.DTYPES:
            if s == d:
                continue
            a = np.array(tv).astype(s)
            with np.errstate(all="ignore"), pytest.warns() if False else np.testing.suppress_warnings() as sup:
                sup.filter(np.exceptions.ComplexWarning)
                exp = a.astype(d)
            assert np.array_equal(ref.convert(a, d), exp), (s, d)


def test_isfinite_isinf_isnan_vector():
    """tests/integration/test_unary_ufunc.py:342-346."""
    v = np.array([-np.inf, 0, 1, np.inf, np.nan])
    for dt in (np.float64, np.complex64, np.float16):
        x = v.astype(dt)
        assert np.array_equal(ref.unary_op("ISFINITE", x), np.isfinite(x))
        assert np.array_equal(ref.unary_op("ISINF", x), np.isinf(x))
        assert np.array_equal(ref.unary_op("ISNAN", x), np.isnan(x))


def test_complex_binary_vectors():
    """tests/integration/test_binary_op_complex.py:21-81."""
    xn = np.array([1 + 4j, 2 + 5j, 3 + 6j], np.complex64)
    yn = np.array([4 + 7j, 5 + 8j, 6 + 9j], np.complex64)
    for op, fn, tol in (("ADD", np.add, 1e-5), ("SUBTRACT", np.subtract, 1e-5),
                        ("MULTIPLY", np.multiply, 1e-5), ("DIVIDE", np.divide, 1e-5),
                        ("POWER", np.power, 1e-1)):
        assert np.max(np.abs(ref.binary_op(op, xn, yn) - fn(xn, yn))) < tol, op


def test_where_fixtures():
    """tests/integration/test_where.py:22-50."""
    x = np.array([[1, 2], [3, 4]])
    y = np.array([[9, 8], [7, 6]])
    for cond in ([[True, False], [True, True]], [[True, False]], [True, False], [False, True]):
        c = np.broadcast_to(np.array(cond), x.shape)
        assert np.array_equal(ref.where(c, x, y), np.where(c, x, y))


def test_map_reduce_min_and_where_sum():
    """test_map_reduce.py:22-31 (=21), test_min_on_gpu.py:21-23 (=1), test_reduction.py:154-158."""
    s = ref.binary_op("ADD", np.array([1, 2, 3]), np.array([4, 5, 6]))
    assert int(ref.scalar_unary_red("SUM", s)) == 21
    assert int(ref.scalar_unary_red("MIN", np.array([1, 2, 3]))) == 1
    a = np.array([[1, 2], [3, 4]])
    assert int(ref.scalar_unary_red("SUM", a, where=np.array([False, True]))) == 6


def test_jacobi_10x10():
    """tests/integration/test_jacobi.py:23-57: fp32 10x10, 2 iterations, vs NumPy."""
    from cunumeric_b200.workloads import stencil_init, stencil_run

    g_ref = stencil_init(8, np.float32, xp=refnp)
    g_np = stencil_init(8, np.float32, xp=np)
    w_ref = stencil_run(g_ref, 2)
    w_np = stencil_run(g_np, 2)
    assert np.allclose(w_ref.a, w_np, rtol=1e-5, atol=1e-8)
    assert np.allclose(np.abs(w_ref.a - g_ref.a[1:-1, 1:-1]).sum(),
                       np.abs(w_np - g_np[1:-1, 1:-1]).sum(), rtol=1e-5)


def test_survey_spot_values():
    """SURVEY §8c spot values of the shim-compiled functors."""
    assert ref.binary_op("FLOOR_DIVIDE", np.int32([-7]), np.int32([2]))[0] == -4
    assert ref.binary_op("MOD", np.float32([-7.5]), np.float32([2]))[0] == 0.5
    assert ref.binary_op("LOGADDEXP", np.float16([1]), np.float16([2]))[0] == np.float16(2.3125)
    assert ref.unary_op("SIGN", np.complex128([2j]))[0] == 1 + 0j
    assert ref.convert(np.float64([0.1]), np.float16)[0] == np.float16(0.0999756)
    assert int(ref.scalar_unary_red("ARGMAX", np.float32([1, 7, 3, 7, 2]))["arg"]) == 1
    nb = sum(ref.binary_out_dtype(o, d) is not None for o in ref.BINARY_OPS for d in ref.DTYPES)
    nu = sum(ref.unary_out_dtype(o, d) is not None for o in ref.UNARY_OPS
             if o not in ("FREXP", "MODF", "GETARG", "POSITIVE") for d in ref.DTYPES)
    assert (nb, nu) == (358, 333)


# ---- differential vs NumPy, the way tests/integration does it -----------------------------------
NUMPY_BINARY = {
    "ADD": np.add, "SUBTRACT": np.subtract, "MULTIPLY": np.multiply, "DIVIDE": np.true_divide,
    "FLOOR_DIVIDE": np.floor_divide, "MOD": np.remainder, "FMOD": np.fmod, "POWER": np.power,
    "EQUAL": np.equal, "NOT_EQUAL": np.not_equal, "LESS": np.less, "LESS_EQUAL": np.less_equal,
    "GREATER": np.greater, "GREATER_EQUAL": np.greater_equal, "LOGICAL_AND": np.logical_and,
    "LOGICAL_OR": np.logical_or, "LOGICAL_XOR": np.logical_xor, "MAXIMUM": np.maximum,
    "MINIMUM": np.minimum, "ARCTAN2": np.arctan2, "HYPOT": np.hypot, "COPYSIGN": np.copysign,
    "NEXTAFTER": np.nextafter, "LOGADDEXP": np.logaddexp, "LOGADDEXP2": np.logaddexp2,
    "BITWISE_AND": np.bitwise_and, "BITWISE_OR": np.bitwise_or, "BITWISE_XOR": np.bitwise_xor,
    "LEFT_SHIFT": np.left_shift, "RIGHT_SHIFT": np.right_shift, "GCD": np.gcd, "LCM": np.lcm,
}


@pytest.mark.parametrize("op", sorted(NUMPY_BINARY))
def test_binary_vs_numpy(op):
    import test_parity_elementwise as tpe

    tpe_n, tpe.N = tpe.N, 257
    try:
        for dt in (np.dtype(np.int32), np.dtype(np.uint32), np.dtype(np.float32),
                   np.dtype(np.float64), np.dtype(np.float16)):
            if ref.binary_out_dtype(op, dt) is None:
                continue
            if op == "POWER" and dt.kind in "iu":
                continue
            rng = pu.rng_for("np-binary", op, dt.name)
            a, b = tpe.binary_inputs(op, dt, rng)
            ok = np.isfinite(a.astype(np.float64)) & np.isfinite(b.astype(np.float64))
            a, b = a[ok], b[ok]
            with np.errstate(all="ignore"):
                exp = NUMPY_BINARY[op](a, b)
                got = ref.binary_op(op, a, b)
            rtol = 1e-2 if dt == np.float16 else 1e-5  # test_binary_ufunc.py:26-30
            assert np.allclose(got.astype(np.float64), exp.astype(np.float64), rtol=rtol,
                               atol=1e-8 if dt != np.float16 else 1e-3, equal_nan=True), (op, dt)
    finally:
        tpe.N = tpe_n


NUMPY_UNARY = {
    "ABSOLUTE": np.absolute, "EXP": np.exp, "EXP2": np.exp2, "EXPM1": np.expm1, "LOG": np.log,
    "LOG2": np.log2, "LOG10": np.log10, "LOG1P": np.log1p, "SQRT": np.sqrt, "CBRT": np.cbrt,
    "SIN": np.sin, "COS": np.cos, "TAN": np.tan, "ARCSIN": np.arcsin, "ARCCOS": np.arccos,
    "ARCTAN": np.arctan, "SINH": np.sinh, "COSH": np.cosh, "TANH": np.tanh, "ARCSINH": np.arcsinh,
    "ARCCOSH": np.arccosh, "ARCTANH": np.arctanh, "NEGATIVE": np.negative, "SQUARE": np.square,
    "RINT": np.rint, "CEIL": np.ceil, "FLOOR": np.floor, "TRUNC": np.trunc,
    "DEG2RAD": np.deg2rad, "RAD2DEG": np.rad2deg, "SIGNBIT": np.signbit, "CONJ": np.conjugate,
    "ISNAN": np.isnan, "LOGICAL_NOT": np.logical_not,
}


@pytest.mark.parametrize("op", sorted(NUMPY_UNARY))
def test_unary_vs_numpy(op):
    import test_parity_elementwise as tpe

    tpe_n, tpe.N = tpe.N, 257
    try:
        for dt in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.float16),
                   np.dtype(np.complex64), np.dtype(np.complex128)):
            if ref.unary_out_dtype(op, dt) is None:
                continue
            rng = pu.rng_for("np-unary", op, dt.name)
            a = tpe.unary_inputs(op, dt, rng)
            with np.errstate(all="ignore"):
                exp = NUMPY_UNARY[op](a)
                got = ref.unary_op(op, a)
            assert got.dtype == exp.dtype, (op, dt)
            rtol = 1e-2 if dt == np.float16 else 1e-5
            assert np.allclose(got, exp, rtol=rtol, atol=1e-3 if dt == np.float16 else 1e-7,
                               equal_nan=True), (op, dt)
    finally:
        tpe.N = tpe_n


def test_reductions_vs_numpy():
    rng = pu.rng_for("np-red")
    for dt in (np.int64, np.uint64, np.float32, np.float64, np.complex64, np.complex128):
        a = (rng.normal(size=(5, 6, 7)) * 3).astype(dt)
        assert np.allclose(ref.scalar_unary_red("SUM", a), a.sum(), rtol=1e-5)
        for axis in range(3):
            assert np.allclose(ref.unary_red("SUM", a, axis), a.sum(axis=axis), rtol=1e-5)
            if np.dtype(dt).kind != "c":
                assert np.array_equal(ref.unary_red("MAX", a, axis), a.max(axis=axis))
                assert np.array_equal(ref.unary_red("ARGMIN", a, axis)["arg"], a.argmin(axis=axis))
    b = rng.random((9, 11)) < 0.7
    assert ref.scalar_unary_red("ALL", b) == b.all() and ref.scalar_unary_red("ANY", b) == b.any()
    assert int(ref.scalar_unary_red("COUNT_NONZERO", b)) == np.count_nonzero(b)
    c = rng.normal(size=101)
    c[::9] = np.nan
    assert np.allclose(ref.scalar_unary_red("NANSUM", c), np.nansum(c))
    assert ref.scalar_unary_red("NANMAX", c) == np.nanmax(c)
    assert int(ref.scalar_unary_red("NANARGMIN", c)["arg"]) == np.nanargmin(c)


def test_openmp_variant_matches_sequential():
    rng = pu.rng_for("omp")
    a = rng.normal(size=100003).astype(np.float32)
    b = rng.normal(size=100003).astype(np.float32)
    assert np.array_equal(ref.binary_op("MULTIPLY", a, b, nthreads=4), ref.binary_op("MULTIPLY", a, b))
    assert np.array_equal(ref.unary_op("EXP", a, nthreads=4), ref.unary_op("EXP", a))
    assert ref.scalar_unary_red("MAX", a, nthreads=4) == ref.scalar_unary_red("MAX", a)
    assert int(ref.scalar_unary_red("ARGMAX", a, nthreads=4)["arg"]) == int(np.argmax(a))
    assert np.allclose(ref.scalar_unary_red("SUM", a, nthreads=4), a.sum(dtype=np.float64), rtol=1e-4)


# ---- golden fixtures pin this oracle build -------------------------------------------------------
def test_oracle_matches_golden_fixtures():
    g = np.load(GOLDEN)
    keys = sorted({k.rsplit("/", 1)[0] for k in g.files})
    checked = 0
    for k in keys:
        kind, *rest = k.split("/")
        with np.errstate(all="ignore"):
            if kind == "binary":
                got = ref.binary_op(rest[0], g[k + "/a"], g[k + "/b"], 1e-3, 1e-5)
                exp = g[k + "/out"]
            elif kind == "unary":
                a = g[k + "/a"]
                extra = None
                if rest[0] == "CLIP":
                    extra = tuple(np.array(v).astype(a.dtype) for v in (
                        (False, True) if a.dtype == np.bool_ else (-3, 5) if a.dtype.kind != "u" else (2, 9)))
                got, exp = ref.unary_op(rest[0], a, extra=extra), g[k + "/out"]
            elif kind == "convert":
                got, exp = ref.convert(g[k + "/a"], np.dtype(rest[2]), rest[0]), g[k + "/out"]
            elif kind == "binred":
                got = np.array([ref.binary_red(rest[0], g[k + "/a"], g[k + "/" + b], 1e-3, 1e-5)
                                for b in ("b_same", "b_diff", "b_near")])
                exp = g[k + "/out"]
            else:
                continue
        assert got.tobytes() == exp.tobytes(), k
        checked += 1
    assert checked > 1000


# ---- BINARY_RED pinned on the reference's own test vectors ---------------------------------------
def _binred(op, a, b, **kw):
    a, b = np.asarray(a), np.asarray(b)
    dt = np.result_type(a, b)
    shape = np.broadcast_shapes(a.shape, b.shape)
    a = np.ascontiguousarray(np.broadcast_to(a.astype(dt), shape))
    b = np.ascontiguousarray(np.broadcast_to(b.astype(dt), shape))
    return ref.binary_red(op, a, b, **kw)


def test_binary_red_known_answers_from_the_reference_tests():
    """tests/integration/test_array_equal.py:21-75 and test_allclose.py:21-125."""
    for arr in (1, [1], [[1, 2], [3, 4]]):
        assert _binred("EQUAL", arr, arr) is True
    for a, b in ((1, 2), ([1], [2]), ([1, 2], [1, 3])):
        assert _binred("EQUAL", a, b) is False and _binred("EQUAL", b, a) is False
    for d1, d2 in ((np.int32, np.float64), (np.float64, np.complex128)):
        assert _binred("EQUAL", np.array([1, 2, 3], d1), np.array([1, 2, 3], d2)) is True
    true_pairs = ((0, -1e-8), (1e10, 1.00001e10), (1 + 1j, 1 + 1.00001j), (np.inf, np.inf),
                  (-np.inf, -np.inf))
    false_pairs = ((0, -0.000001), (1e10, 1.0001e10), (1 + 1j, 1 + 1.0001j), (np.inf, -np.inf))
    for a, b in true_pairs:
        assert _binred("ISCLOSE", a, b) is np.allclose(a, b) is True
        assert _binred("ISCLOSE", b, a) is np.allclose(b, a) is True
    for a, b in false_pairs:
        assert _binred("ISCLOSE", a, b) is np.allclose(a, b) is False
        assert _binred("ISCLOSE", b, a) is np.allclose(b, a) is False
    for shape in ((1,), (6,), (1, 1), (2, 3), (2, 3, 4)):
        size = int(np.prod(shape))
        for pairs, expect in ((true_pairs[:3], True), (false_pairs[:3], False)):
            a = np.array([pairs[i % 3][0] for i in range(size)]).reshape(shape)
            b = np.array([pairs[i % 3][1] for i in range(size)]).reshape(shape)
            assert _binred("ISCLOSE", a, b) is bool(np.allclose(a, b)) is expect
    # test_allclose.py:128-150 rtol / atol arguments
    assert _binred("ISCLOSE", 1e10, 1.0001e10, rtol=1e-3) is True
    assert _binred("ISCLOSE", 0.0, -1e-6, atol=1e-5) is True
    # NaN never compares equal / close (equal_nan is unsupported in the reference)
    assert _binred("EQUAL", [1.0, np.nan], [1.0, np.nan]) is False
    assert _binred("ISCLOSE", [1.0, np.nan], [1.0, np.nan]) is False
