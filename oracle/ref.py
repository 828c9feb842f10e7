"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/_ref/libcunumeric_ref.so.

The shared object is the reference's own functor headers (binary_op_util.h, unary_op_util.h,
unary_red_util.h, convert_util.h, arg.h/arg.inl — included by path from /root/reference/src)
compiled with g++ against the stand-in ``oracle/shim/legate.h`` by ``oracle/Makefile``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``cunumeric_b200``) never does.

All entry points take/return dense C-order numpy arrays (the test harness applies views and
broadcasting with numpy before calling, which is what Legion's accessors do for the reference).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcunumeric_ref.so")

# legate::Type::Code order: ? b h i l B H I L e f d F D
DTYPES = [
    np.dtype(np.bool_), np.dtype(np.int8), np.dtype(np.int16), np.dtype(np.int32),
    np.dtype(np.int64), np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.uint32),
    np.dtype(np.uint64), np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64),
    np.dtype(np.complex64), np.dtype(np.complex128),
]
CODE_OF = {dt: i for i, dt in enumerate(DTYPES)}

# cunumeric_c.h:86-134 (CuNumericUnaryOpCode, from 1), :138-157, :161-197, :236-240
UNARY_OPS = (
    "ABSOLUTE ARCCOS ARCCOSH ARCSIN ARCSINH ARCTAN ARCTANH CBRT CEIL CLIP CONJ COPY COS COSH "
    "DEG2RAD EXP EXP2 EXPM1 FLOOR FREXP GETARG IMAG INVERT ISFINITE ISINF ISNAN LOG LOG10 LOG1P "
    "LOG2 LOGICAL_NOT MODF NEGATIVE POSITIVE RAD2DEG REAL RECIPROCAL RINT SIGN SIGNBIT SIN SINH "
    "SQRT SQUARE TAN TANH TRUNC"
).split()
RED_OPS = (
    "ALL ANY ARGMAX ARGMIN CONTAINS COUNT_NONZERO MAX MIN NANARGMAX NANARGMIN NANMAX NANMIN "
    "NANPROD NANSUM PROD SUM SUM_SQUARES VARIANCE"
).split()
BINARY_OPS = (
    "ADD ARCTAN2 BITWISE_AND BITWISE_OR BITWISE_XOR COPYSIGN DIVIDE EQUAL FLOAT_POWER "
    "FLOOR_DIVIDE FMOD GCD GREATER GREATER_EQUAL HYPOT ISCLOSE LCM LDEXP LEFT_SHIFT LESS "
    "LESS_EQUAL LOGADDEXP LOGADDEXP2 LOGICAL_AND LOGICAL_OR LOGICAL_XOR MAXIMUM MINIMUM MOD "
    "MULTIPLY NEXTAFTER NOT_EQUAL POWER RIGHT_SHIFT SUBTRACT"
).split()
CONVERT_OPS = "NOOP PROD SUM".split()
UNARY = {n: i + 1 for i, n in enumerate(UNARY_OPS)}
RED = {n: i + 1 for i, n in enumerate(RED_OPS)}
BINARY = {n: i + 1 for i, n in enumerate(BINARY_OPS)}
CONVERT = {n: i + 1 for i, n in enumerate(CONVERT_OPS)}

ARG_RED = {"ARGMAX", "ARGMIN", "NANARGMAX", "NANARGMIN"}

ERR_INVALID = -1


def argval_dtype(dt) -> np.dtype:
    """Argval<T> = {int64 arg; T arg_value}, 16 bytes for every T up to 8 bytes (arg.h:23-59)."""
    dt = np.dtype(dt)
    return np.dtype({"names": ["arg", "arg_value"], "formats": [np.int64, dt],
                     "offsets": [0, 8], "itemsize": 8 + max(8, dt.itemsize)})


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run `make -C oracle` where /root/reference is mounted")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.ref_binary_out_code.argtypes = [i32, i32]
        L.ref_binary_op.argtypes = [i32, i32, vp, vp, vp, sz, vp, i32]
        L.ref_binary_red.argtypes = [i32, i32, vp, vp, vp, sz, vp]
        L.ref_binary_op_strided.argtypes = [i32, i32, vp, sz, vp, sz, vp, sz, vp, i32]
        L.ref_unary_out_code.argtypes = [i32, i32]
        L.ref_unary_op.argtypes = [i32, i32, vp, vp, sz, vp, i32]
        L.ref_unary_multiout.argtypes = [i32, i32, vp, vp, vp, sz]
        L.ref_getarg.argtypes = [vp, vp, sz]
        L.ref_convert.argtypes = [i32, i32, i32, vp, vp, sz, i32]
        L.ref_where.argtypes = [i32, vp, vp, vp, vp, sz, i32]
        L.ref_scalar_unary_red.argtypes = [i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, i32]
        L.ref_unary_red.argtypes = [i32, i32, vp, vp, i32, vp, vp, i32, vp, i32]
        L.ref_red_identity.argtypes = [i32, i32, vp]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _dense(a, dtype=None) -> np.ndarray:
    a = np.asarray(a, dtype=dtype)
    # (np.ascontiguousarray would promote 0-d arrays to shape (1,))
    return a if a.ndim == 0 else np.ascontiguousarray(a)


def _i64(seq: Sequence[int]) -> np.ndarray:
    return np.asarray(list(seq), dtype=np.int64)


class InvalidOp(ValueError):
    """The reference marks this (op, dtype) `valid = false` (the task would assert(false))."""


def binary_out_dtype(op: str, dtype) -> Optional[np.dtype]:
    rc = lib().ref_binary_out_code(BINARY[op], CODE_OF[np.dtype(dtype)])
    return DTYPES[rc] if rc >= 0 else None


def unary_out_dtype(op: str, dtype) -> Optional[np.dtype]:
    if op in ("FREXP", "MODF", "GETARG"):
        raise ValueError(op)
    rc = lib().ref_unary_out_code(UNARY[op], CODE_OF[np.dtype(dtype)])
    return DTYPES[rc] if rc >= 0 else None


def binary_op(op: str, a, b, rtol: float = 1e-5, atol: float = 1e-8, nthreads: int = 1):
    a = _dense(a)
    rhs2_dtype = np.int32 if op == "LDEXP" else a.dtype  # binary_op_util.h:884-887
    b = _dense(b, rhs2_dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    odt = binary_out_dtype(op, a.dtype)
    if odt is None:
        raise InvalidOp(f"{op}/{a.dtype}")
    out = np.empty(a.shape, dtype=odt)
    extra = np.array([rtol, atol], dtype=np.float64)
    rc = lib().ref_binary_op(BINARY[op], CODE_OF[a.dtype], _ptr(a), _ptr(b), _ptr(out), a.size,
                             _ptr(extra), nthreads)
    assert rc == CODE_OF[odt], rc
    return out


def binary_red(op: str, a, b, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """BINARY_RED: all(op(a, b)) for op in EQUAL / ISCLOSE (binary/binary_red.cc:38-47)."""
    a = _dense(a)
    b = _dense(b, a.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    out = np.array([True])
    extra = np.array([rtol, atol], dtype=np.float64)
    rc = lib().ref_binary_red(BINARY[op], CODE_OF[a.dtype], _ptr(a), _ptr(b), _ptr(out), a.size,
                              _ptr(extra))
    if rc == ERR_INVALID:
        raise InvalidOp(f"{op}/{a.dtype}")
    assert rc == 0, rc
    return bool(out[0])


def binary_op_bcast(op: str, a, b, nthreads: int = 1, out=None):
    """binary_op where either operand may be a 0-d scalar read through a stride-0 accessor (no
    materialised broadcast) — what the reference's CPU loop does for Future-backed scalars."""
    a = _dense(a)
    b = _dense(b, np.int32 if op == "LDEXP" else a.dtype)
    shape = a.shape if a.ndim else b.shape
    n = int(np.prod(shape, dtype=np.int64))
    odt = binary_out_dtype(op, a.dtype)
    if odt is None:
        raise InvalidOp(f"{op}/{a.dtype}")
    if out is None:
        out = np.empty(shape, dtype=odt)
    extra = np.array([1e-5, 1e-8], dtype=np.float64)
    rc = lib().ref_binary_op_strided(BINARY[op], CODE_OF[a.dtype], _ptr(a), 1 if a.ndim else 0,
                                     _ptr(b), 1 if b.ndim else 0, _ptr(out), n, _ptr(extra),
                                     nthreads)
    assert rc == CODE_OF[odt], rc
    return out


def unary_op(op: str, a, extra=None, nthreads: int = 1):
    a = _dense(a)
    odt = unary_out_dtype(op, a.dtype)
    if odt is None:
        raise InvalidOp(f"{op}/{a.dtype}")
    out = np.empty(a.shape, dtype=odt)
    ex = None
    if op == "CLIP":
        ex = np.array(list(extra), dtype=a.dtype)
        assert ex.size == 2
    rc = lib().ref_unary_op(UNARY[op], CODE_OF[a.dtype], _ptr(a), _ptr(out), a.size, _ptr(ex),
                            nthreads)
    assert rc == CODE_OF[odt], rc
    return out


def unary_multiout(op: str, a):
    a = _dense(a)
    code2 = lib().ref_unary_multiout(UNARY[op], CODE_OF[a.dtype], None, None, None, 0)
    if code2 < 0:
        raise InvalidOp(f"{op}/{a.dtype}")
    out1 = np.empty(a.shape, dtype=a.dtype)
    out2 = np.empty(a.shape, dtype=DTYPES[code2])
    lib().ref_unary_multiout(UNARY[op], CODE_OF[a.dtype], _ptr(a), _ptr(out1), _ptr(out2), a.size)
    return out1, out2


def getarg(a):
    a = _dense(a)
    assert a.dtype.itemsize == 16
    out = np.empty(a.shape, dtype=np.int64)
    lib().ref_getarg(_ptr(a), _ptr(out), a.size)
    return out


def convert(a, dst, nan_op: str = "NOOP", nthreads: int = 1):
    a = _dense(a)
    dst = np.dtype(dst)
    out = np.empty(a.shape, dtype=dst)
    rc = lib().ref_convert(CONVERT[nan_op], CODE_OF[dst], CODE_OF[a.dtype], _ptr(a), _ptr(out),
                           a.size, nthreads)
    if rc < 0:
        raise InvalidOp(f"CONVERT {nan_op} {a.dtype}->{dst}")
    return out


def where(mask, a, b, nthreads: int = 1):
    mask = _dense(mask, np.bool_)
    a = _dense(a)
    b = _dense(b, a.dtype)
    assert mask.shape == a.shape == b.shape
    out = np.empty(a.shape, dtype=a.dtype)
    lib().ref_where(CODE_OF[a.dtype], _ptr(mask), _ptr(a), _ptr(b), _ptr(out), a.size, nthreads)
    return out


def red_val_dtype(op: str, dtype) -> np.dtype:
    dtype = np.dtype(dtype)
    if op in ARG_RED:
        return argval_dtype(dtype)
    if op in ("ALL", "ANY", "CONTAINS"):
        return np.dtype(np.bool_)
    if op == "COUNT_NONZERO":
        return np.dtype(np.uint64)
    return dtype


def red_identity(op: str, dtype) -> np.ndarray:
    """LG_OP::identity of the (op, dtype) reduction as a 0-d array of the VAL dtype."""
    vdt = red_val_dtype(op, dtype)
    out = np.zeros((), dtype=vdt)
    rc = lib().ref_red_identity(RED[op], CODE_OF[np.dtype(dtype)], _ptr(out))
    if rc < 0:
        raise InvalidOp(f"{op}/{dtype}")
    assert rc == vdt.itemsize, (rc, vdt)
    return out


def scalar_unary_red(op: str, a, where=None, initial=None, origin=None, shape=None, extra=None,
                     nthreads: int = 1):
    """Fold the whole (dense) array into one VAL. `initial` pre-fills the output (defaults to the
    reduction identity); `origin`/`shape` place the rect in a larger global array (arg-reductions
    return GLOBAL row-major flat indices)."""
    a = _dense(a)
    if a.ndim == 0:
        a = a.reshape(1)  # scalar_unary_red_template.inl:177-181
    vdt = red_val_dtype(op, a.dtype)
    out = red_identity(op, a.dtype) if initial is None else np.array(initial, dtype=vdt)
    w = None if where is None else _dense(np.broadcast_to(where, a.shape), np.bool_)
    ex = None if extra is None else np.array(extra, dtype=a.dtype)
    ext = _i64(a.shape)
    org = _i64(origin if origin is not None else [0] * a.ndim)
    shp = _i64(shape if shape is not None else a.shape)
    rc = lib().ref_scalar_unary_red(RED[op], CODE_OF[a.dtype], _ptr(a), _ptr(w), a.ndim, _ptr(ext),
                                    _ptr(org), _ptr(shp), _ptr(out), _ptr(ex), nthreads)
    if rc < 0:
        raise InvalidOp(f"{op}/{a.dtype}")
    return out


def unary_red(op: str, a, axis: int, where=None, initial=None, origin=None, nthreads: int = 1):
    """Reduce along one axis; returns the VAL array with `axis` removed."""
    a = _dense(a)
    assert a.ndim > 1  # unary_red_template.inl:35-37
    axis = axis % a.ndim
    vdt = red_val_dtype(op, a.dtype)
    oshape = a.shape[:axis] + a.shape[axis + 1:]
    out = np.empty(oshape, dtype=vdt)
    out[...] = red_identity(op, a.dtype) if initial is None else np.array(initial, dtype=vdt)
    w = None if where is None else _dense(np.broadcast_to(where, a.shape), np.bool_)
    ext = _i64(a.shape)
    org = _i64(origin if origin is not None else [0] * a.ndim)
    rc = lib().ref_unary_red(RED[op], CODE_OF[a.dtype], _ptr(a), _ptr(w), a.ndim, _ptr(ext),
                             _ptr(org), axis, _ptr(out), nthreads)
    if rc < 0:
        raise InvalidOp(f"{op}/{a.dtype}")
    return out
