"""Opcode enums of the hot path, numerically identical to the reference's C header
(src/cunumeric/cunumeric_c.h:27-240) and named like cunumeric/config.py:334-; the reference reads
them back from the cffi'd library, here they are checked against include/cunumeric_b200.h by
tests/test_abi.py."""
from __future__ import annotations

from enum import IntEnum, unique

import numpy as np


@unique
class CuNumericOpCode(IntEnum):
    BINARY_OP = 5
    BINARY_RED = 6
    CONVERT = 11
    FILL = 19
    SCALAR_UNARY_RED = 33
    UNARY_OP = 43
    UNARY_RED = 44
    WHERE = 49


_UNARY = (
    "ABSOLUTE ARCCOS ARCCOSH ARCSIN ARCSINH ARCTAN ARCTANH CBRT CEIL CLIP CONJ COPY COS COSH "
    "DEG2RAD EXP EXP2 EXPM1 FLOOR FREXP GETARG IMAG INVERT ISFINITE ISINF ISNAN LOG LOG10 LOG1P "
    "LOG2 LOGICAL_NOT MODF NEGATIVE POSITIVE RAD2DEG REAL RECIPROCAL RINT SIGN SIGNBIT SIN SINH "
    "SQRT SQUARE TAN TANH TRUNC"
)
_RED = (
    "ALL ANY ARGMAX ARGMIN CONTAINS COUNT_NONZERO MAX MIN NANARGMAX NANARGMIN NANMAX NANMIN "
    "NANPROD NANSUM PROD SUM SUM_SQUARES VARIANCE"
)
_BINARY = (
    "ADD ARCTAN2 BITWISE_AND BITWISE_OR BITWISE_XOR COPYSIGN DIVIDE EQUAL FLOAT_POWER "
    "FLOOR_DIVIDE FMOD GCD GREATER GREATER_EQUAL HYPOT ISCLOSE LCM LDEXP LEFT_SHIFT LESS "
    "LESS_EQUAL LOGADDEXP LOGADDEXP2 LOGICAL_AND LOGICAL_OR LOGICAL_XOR MAXIMUM MINIMUM MOD "
    "MULTIPLY NEXTAFTER NOT_EQUAL POWER RIGHT_SHIFT SUBTRACT"
)

# cunumeric_c.h:86-134 / :138-157 / :161-197 / :236-240 (all start at 1, alphabetical)
UnaryOpCode = unique(IntEnum("UnaryOpCode", _UNARY, start=1))
UnaryRedCode = unique(IntEnum("UnaryRedCode", _RED, start=1))
BinaryOpCode = unique(IntEnum("BinaryOpCode", _BINARY, start=1))
ConvertCode = unique(IntEnum("ConvertCode", "NOOP PROD SUM", start=1))


@unique
class CuNumericRedopCode(IntEnum):
    ARGMAX = 1
    ARGMIN = 2


# legate::Type::Code order (? b h i l B H I L e f d F D)
SUPPORTED_DTYPES = (
    np.dtype(np.bool_), np.dtype(np.int8), np.dtype(np.int16), np.dtype(np.int32),
    np.dtype(np.int64), np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.uint32),
    np.dtype(np.uint64), np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64),
    np.dtype(np.complex64), np.dtype(np.complex128),
)
DTYPE_CODE = {dt: i for i, dt in enumerate(SUPPORTED_DTYPES)}
ARGVAL_BASE = 32
MAX_DIM = 4


def argval_dtype(elem) -> np.dtype:
    """Struct dtype {int64 arg; T arg_value} padded to 16 bytes — what runtime.get_argred_type
    builds in the reference (cunumeric/runtime.py:125-134) for Argval<T> (src/cunumeric/arg.h)."""
    elem = np.dtype(elem)
    return np.dtype({"names": ["arg", "arg_value"], "formats": [np.int64, elem],
                     "offsets": [0, 8], "itemsize": 8 + max(8, elem.itemsize)})


def dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    code = DTYPE_CODE.get(dtype)
    if code is not None:
        return code
    if dtype.names == ("arg", "arg_value") and dtype.itemsize == 16:
        return ARGVAL_BASE + DTYPE_CODE[dtype.fields["arg_value"][0]]
    raise TypeError(f"cunumeric_b200 does not support dtype={dtype}")


def is_supported_dtype(dtype) -> bool:
    return np.dtype(dtype) in DTYPE_CODE
