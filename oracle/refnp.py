"""TEST INFRASTRUCTURE ONLY — a tiny NumPy-like facade over the oracle (oracle/ref.py), so that an
API-level program (cunumeric_b200/workloads.py: Black-Scholes, the Jacobi stencil) can be evaluated
task-by-task with the reference's own functors: every operator / function call below is exactly one
reference task (BINARY_OP, UNARY_OP, WHERE, SCALAR_UNARY_RED) on the CPU, with scalar operands read
through a stride-0 operand like Future-backed stores.  Used as the CPU baseline (`--impl reference`,
cpu_baseline) and by smoke()/tests as the checker for whole-workload parity."""
from __future__ import annotations

import numpy as np

from . import ref

NTHREADS = 1


def set_threads(n: int) -> None:
    global NTHREADS
    NTHREADS = max(1, int(n))


def _weak(x, dtype):
    """Python scalars are weak: they take the array operand's dtype when the kinds allow it."""
    if isinstance(x, RefArray):
        return x.a
    if isinstance(x, np.ndarray):
        return x
    return np.asarray(x, dtype=np.result_type(dtype, x))


class RefArray:
    __array_priority__ = 200.0

    def __init__(self, a) -> None:
        self.a = np.asarray(a)

    @property
    def shape(self):
        return self.a.shape

    @property
    def dtype(self):
        return self.a.dtype

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def _bin(self, op, other, reverse=False):
        o = _weak(other, self.a.dtype)
        x, y = (o, self.a) if reverse else (self.a, o)
        common = np.result_type(x.dtype, y.dtype)
        if x.dtype != common:
            x = ref.convert(x, common) if x.ndim else x.astype(common)
        if y.dtype != common:
            y = ref.convert(y, common) if y.ndim else y.astype(common)
        if x.ndim and y.ndim and x.shape != y.shape:
            x, y = (np.ascontiguousarray(v) for v in np.broadcast_arrays(x, y))
        if x.ndim == 0 and y.ndim == 0:
            return RefArray(ref.binary_op(op, x.reshape(1), y.reshape(1)).reshape(()))
        return RefArray(ref.binary_op_bcast(op, x, y, nthreads=NTHREADS))

    def __add__(self, o): return self._bin("ADD", o)
    def __radd__(self, o): return self._bin("ADD", o, True)
    def __sub__(self, o): return self._bin("SUBTRACT", o)
    def __rsub__(self, o): return self._bin("SUBTRACT", o, True)
    def __mul__(self, o): return self._bin("MULTIPLY", o)
    def __rmul__(self, o): return self._bin("MULTIPLY", o, True)
    def __truediv__(self, o): return self._bin("DIVIDE", o)
    def __rtruediv__(self, o): return self._bin("DIVIDE", o, True)
    def __gt__(self, o): return self._bin("GREATER", o)
    def __lt__(self, o): return self._bin("LESS", o)
    def __neg__(self): return RefArray(ref.unary_op("NEGATIVE", self.a, nthreads=NTHREADS))

    def __getitem__(self, key):
        return RefArray(self.a[key])  # NumPy views alias like legate store views

    def __setitem__(self, key, value):
        # view[:] = value is UNARY_OP(COPY) (deferred.py:392-401)
        v = value.a if isinstance(value, RefArray) else np.asarray(value, dtype=self.a.dtype)
        dst = self.a[key]
        if v.ndim == 0 or v.shape != dst.shape:
            dst[...] = v
        else:
            dst[...] = ref.unary_op("COPY", v, nthreads=NTHREADS)

    def sum(self):
        return RefArray(ref.scalar_unary_red("SUM", self.a, nthreads=NTHREADS))


def _un(op):
    def fn(x):
        return RefArray(ref.unary_op(op, x.a if isinstance(x, RefArray) else np.asarray(x),
                                     nthreads=NTHREADS))
    return fn


sqrt = _un("SQRT")
log = _un("LOG")
exp = _un("EXP")
absolute = _un("ABSOLUTE")


def where(m, x, y):
    m = m.a if isinstance(m, RefArray) else np.asarray(m)
    dt = x.dtype if isinstance(x, RefArray) else y.dtype
    x, y = _weak(x, dt), _weak(y, dt)
    x, y = (np.ascontiguousarray(np.broadcast_to(v, m.shape)) for v in (x, y))
    return RefArray(ref.where(m, x, y, nthreads=NTHREADS))


def zeros(shape, dtype=np.float64):
    return RefArray(np.zeros(shape, dtype=dtype))


def array(a):
    return RefArray(np.array(a))
