"""Shared helpers for the parity tests: seeded inputs that stay inside each op's defined domain,
ULP distance, and thin wrappers that drive the CUDA path through the thunk layer / C ABI."""
from __future__ import annotations

import numpy as np

from oracle import ref

DTYPES = ref.DTYPES
INT_DTYPES = [d for d in DTYPES if d.kind in "iu"]
FLOAT_DTYPES = [np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64)]
COMPLEX_DTYPES = [np.dtype(np.complex64), np.dtype(np.complex128)]

# ops whose floating-point result must be bit-exact (IEEE basic ops, selections, roundings)
EXACT_BINARY = {
    "ADD", "SUBTRACT", "MULTIPLY", "DIVIDE", "MAXIMUM", "MINIMUM", "COPYSIGN", "NEXTAFTER",
    "LDEXP", "FMOD", "MOD", "FLOOR_DIVIDE", "EQUAL", "NOT_EQUAL", "LESS", "LESS_EQUAL", "GREATER",
    "GREATER_EQUAL", "LOGICAL_AND", "LOGICAL_OR", "LOGICAL_XOR", "BITWISE_AND", "BITWISE_OR",
    "BITWISE_XOR", "LEFT_SHIFT", "RIGHT_SHIFT", "GCD", "LCM", "ISCLOSE",
}
EXACT_UNARY = {
    "ABSOLUTE", "NEGATIVE", "POSITIVE", "COPY", "CONJ", "SQUARE", "SIGN", "CEIL", "FLOOR", "TRUNC",
    "RINT", "SQRT", "RECIPROCAL", "REAL", "IMAG", "DEG2RAD", "RAD2DEG", "CLIP", "ISFINITE",
    "ISINF", "ISNAN", "LOGICAL_NOT", "SIGNBIT", "INVERT",
}


def rng_for(*key) -> np.random.Generator:
    import zlib

    return np.random.default_rng(zlib.crc32(repr(key).encode()))


def make_input(dtype, n, rng, kind="general"):
    """kind: general | positive | unit (|x|<=1) | small (|x|<=4) | nonzero | shift | pow_base |
    pow_exp | ge1"""
    dtype = np.dtype(dtype)
    if dtype == np.bool_:
        if kind in ("nonzero", "positive"):
            return np.ones(n, dtype=np.bool_)  # x / False is a division by zero in the functors
        return rng.random(n) < 0.5
    if dtype.kind in "iu":
        info = np.iinfo(dtype)
        if kind == "shift":
            return rng.integers(0, dtype.itemsize * 8 - 1, n).astype(dtype)
        if kind == "pow_base":
            lo = -3 if dtype.kind == "i" else 0
            return rng.integers(lo, 4, n).astype(dtype)
        if kind == "pow_exp":
            return rng.integers(0, 4, n).astype(dtype)
        if kind == "small":
            lo = -100 if dtype.kind == "i" else 0
            return rng.integers(lo, 101, n).astype(dtype)
        lo = max(info.min, -(2 ** 31)) if dtype.kind == "i" else 0
        hi = min(info.max, 2 ** 31)
        if dtype.kind == "i":
            lo += 1  # never INT_MIN: INT_MIN / -1 traps on the CPU oracle
        a = rng.integers(lo, hi, n, dtype=np.int64).astype(dtype)
        if kind in ("nonzero", "positive"):
            a[a == 0] = 1
        if kind == "positive" and dtype.kind == "i":
            a = np.abs(a).astype(dtype)
            a[a <= 0] = 1
        return a
    if dtype.kind == "f":
        if kind == "unit":
            a = rng.uniform(-1, 1, n)
        elif kind == "positive":
            a = rng.uniform(0.01, 50, n)
        elif kind == "ge1":
            a = rng.uniform(1, 50, n)
        elif kind in ("small", "pow_base", "pow_exp"):
            a = rng.uniform(-4, 4, n)
        elif kind == "nonzero":
            a = rng.uniform(-50, 50, n)
            a[a == 0] = 1
        else:
            a = rng.normal(0, 8, n)
        a = a.astype(dtype)
        if kind == "nonzero":
            a[a == 0] = 1
        return a
    if dtype.kind == "c":
        part = np.float32 if dtype == np.complex64 else np.float64
        re = make_input(part, n, rng, kind if kind not in ("positive", "ge1") else "small")
        im = make_input(part, n, rng, kind if kind not in ("positive", "ge1") else "small")
        return (re + 1j * im).astype(dtype)
    raise TypeError(dtype)


def with_specials(a, rng):
    """Sprinkle 0, -0, inf, -inf, nan into a float/complex array (in place, returns it)."""
    if a.dtype.kind not in "fc" or a.size < 16:
        return a
    idx = rng.choice(a.size, size=10, replace=False)
    specials = [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 0.5, 2.0, -2.0]
    flat = a.reshape(-1)
    for i, s in zip(idx, specials):
        flat[i] = s
    return a


def ulp_distance(got, exp):
    """Max distance in units of the last place between two float arrays of the same dtype (NaNs
    must coincide, infinities must match exactly)."""
    got = np.asarray(got)
    exp = np.asarray(exp)
    assert got.dtype == exp.dtype and got.shape == exp.shape, (got.dtype, exp.dtype)
    if got.dtype.kind == "c":
        return max(ulp_distance(got.real, exp.real), ulp_distance(got.imag, exp.imag))
    nan_g, nan_e = np.isnan(got), np.isnan(exp)
    if not np.array_equal(nan_g, nan_e):
        return np.inf
    ok = ~nan_g
    g, e = got[ok], exp[ok]
    if g.size == 0:
        return 0
    itype = {2: np.int16, 4: np.int32, 8: np.int64}[got.dtype.itemsize]
    gi = g.view(itype).astype(np.int64)
    ei = e.view(itype).astype(np.int64)
    mask = np.int64(np.iinfo(itype).max)
    # sign-magnitude bit patterns -> a monotone integer line (-0 and +0 coincide)
    gi = np.where(gi < 0, -(gi & mask), gi)
    ei = np.where(ei < 0, -(ei & mask), ei)
    d = np.abs(gi - ei)
    inf_mismatch = np.isinf(g) != np.isinf(e)
    if inf_mismatch.any():
        return np.inf
    return int(d.max())


def assert_close_scaled(got, exp, scale, k, what=""):
    """|got - exp| <= k * eps * scale element-wise (NaNs / infinities must coincide).  Used where a
    plain ulp count of the RESULT is the wrong yardstick: complex functions (error is relative to
    |z|, a tiny component next to a large one carries the large one's rounding) and sums that
    cancel (logaddexp of negative operands)."""
    got = np.asarray(got)
    exp = np.asarray(exp)
    assert got.dtype == exp.dtype and got.shape == exp.shape, f"{what}: {got.dtype} vs {exp.dtype}"
    part = got.real.dtype
    eps = np.finfo(part).eps
    bad_g = ~np.isfinite(got)
    bad_e = ~np.isfinite(exp)
    assert np.array_equal(bad_g, bad_e), f"{what}: non-finite positions differ"
    ok = ~bad_e
    if got.dtype.kind != "c":
        assert np.array_equal(got[bad_e], exp[bad_e], equal_nan=True), f"{what}: inf/nan mismatch"
    err = np.abs(got[ok].astype(np.complex128) - exp[ok].astype(np.complex128))
    scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), exp.shape)[ok]
    bound = k * eps * np.maximum(scale, float(np.finfo(part).tiny))
    worst = (err / bound).max() if err.size else 0.0
    assert worst <= 1.0, f"{what}: error {worst * k:.2f} eps*scale > {k}"


def assert_close_ulp(got, exp, max_ulp, what=""):
    got = np.asarray(got)
    exp = np.asarray(exp)
    assert got.dtype == exp.dtype, f"{what}: dtype {got.dtype} != {exp.dtype}"
    assert got.shape == exp.shape, f"{what}: shape {got.shape} != {exp.shape}"
    if got.dtype.kind in "fc":
        if max_ulp == 0:
            same = (got.view(np.uint8).reshape(-1) == exp.view(np.uint8).reshape(-1)).all()
            if not same:
                # +-0 and NaN payloads are not significant
                d = ulp_distance(got, exp)
                assert d == 0, f"{what}: not bit-exact, ulp distance {d}"
        else:
            d = ulp_distance(got, exp)
            assert d <= max_ulp, f"{what}: {d} ulp > {max_ulp}"
    else:
        assert np.array_equal(got, exp), f"{what}: mismatch at {np.flatnonzero(got != exp)[:5]}"


# ---------------------------------------------------------------------------------------------
# CUDA-side drivers (through the thunk layer, i.e. the C ABI)
def to_device(a):
    from cunumeric_b200.deferred import DeferredArray
    from cunumeric_b200.store import Store

    a = np.asarray(a)
    if a.ndim == 0:
        return DeferredArray(Store.from_scalar(a))
    return DeferredArray.from_numpy(a)


def new_thunk(shape, dtype):
    from cunumeric_b200.deferred import DeferredArray
    from cunumeric_b200.store import Store

    return DeferredArray(Store.empty(shape, dtype))


def gpu_binary(op, a, b, out_dtype, args=()):
    from cunumeric_b200.config import BinaryOpCode

    out = new_thunk(np.broadcast_shapes(np.shape(a), np.shape(b)), out_dtype)
    out.binary_op(BinaryOpCode[op], to_device(a), to_device(b), True, args)
    return out.__numpy_array__()


def gpu_unary(op, a, out_dtype, args=()):
    from cunumeric_b200.config import UnaryOpCode

    out = new_thunk(np.shape(a), out_dtype)
    out.unary_op(UnaryOpCode[op], to_device(a), True, args)
    return out.__numpy_array__()
