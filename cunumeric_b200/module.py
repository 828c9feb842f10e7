"""Module-level NumPy API for the hot path (the thin wrappers of the reference's
cunumeric/module.py that forward to ndarray: where :3104, sum :5476, amax :6514, argmax :7194,
all :5079, any :5137, prod :5395, plus array creation)."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from .array import convert_to_cunumeric_ndarray, ndarray
from .config import UnaryRedCode
from .deferred import DeferredArray
from .runtime import runtime
from .store import Store


def _is_weak_scalar(x: Any) -> bool:
    return isinstance(x, (bool, int, float, complex)) and not isinstance(x, np.generic)


# ---------------------------------------------------------------------- creation
def array(obj: Any, dtype=None, copy: bool = True, order="K", subok=False, ndmin: int = 0) -> ndarray:
    if isinstance(obj, ndarray):
        out = obj if dtype is None or np.dtype(dtype) == obj.dtype else obj.astype(dtype)
        return out.copy() if (copy and out is obj) else out
    # the host->device transfer IS the copy: no extra host-side duplicate.  (A pinned source from
    # pinned_empty() is read asynchronously; call synchronize() before overwriting it.)
    host = np.asarray(obj, dtype=dtype)
    if ndmin > host.ndim:
        host = host.reshape((1,) * (ndmin - host.ndim) + host.shape)
    return convert_to_cunumeric_ndarray(host)


def asarray(a: Any, dtype=None) -> ndarray:
    if isinstance(a, ndarray) and (dtype is None or np.dtype(dtype) == a.dtype):
        return a
    return array(a, dtype=dtype, copy=False)


def empty(shape, dtype=np.float64) -> ndarray:
    return ndarray(shape=shape, dtype=dtype)


def full(shape, value, dtype=None) -> ndarray:
    if dtype is None:
        dtype = np.asarray(value).dtype
    out = ndarray(shape=shape, dtype=dtype)
    out.fill(value)
    return out


def zeros(shape, dtype=np.float64) -> ndarray:
    return full(shape, 0, dtype)


def ones(shape, dtype=np.float64) -> ndarray:
    return full(shape, 1, dtype)


def empty_like(a, dtype=None, shape=None) -> ndarray:
    a = convert_to_cunumeric_ndarray(a)
    return ndarray(shape=a.shape if shape is None else shape, dtype=dtype or a.dtype)


def zeros_like(a, dtype=None, shape=None) -> ndarray:
    out = empty_like(a, dtype, shape)
    out.fill(0)
    return out


def ones_like(a, dtype=None, shape=None) -> ndarray:
    out = empty_like(a, dtype, shape)
    out.fill(1)
    return out


def full_like(a, value, dtype=None, shape=None) -> ndarray:
    out = empty_like(a, dtype, shape)
    out.fill(value)
    return out


def copy(a) -> ndarray:
    return convert_to_cunumeric_ndarray(a).copy()


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Page-locked HOST array (NumPy) for asynchronous transfers into/out of the device."""
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return runtime.pinned_empty(tuple(shape), dtype)


def from_host(host: np.ndarray, blocking: bool = True) -> ndarray:
    """Upload a host array.  blocking=False issues the copy on the copy stream (the source should
    be pinned, see pinned_empty, and must stay untouched until the result is first used) so that it
    overlaps with kernels already queued."""
    host = np.asarray(host)
    from .distributed import _partitioning

    if blocking or host.ndim == 0 or (runtime.world_size > 1 and _partitioning[0]):
        return convert_to_cunumeric_ndarray(host)
    return ndarray(shape=host.shape, dtype=host.dtype, thunk=DeferredArray.from_numpy_async(host))


def from_host_rows(block: np.ndarray, global_rows: int, blocking: bool = True) -> ndarray:
    """SPMD upload of a row-partitioned array: every rank passes ITS block of rows (the even split
    of `global_rows` over the ranks).  With one GPU the block is the whole array.  blocking=False
    issues the copy on the H2D stream (pinned source, untouched until the result is first used) so
    that it overlaps with kernels already queued."""
    block = np.asarray(block)
    if runtime.world_size == 1:
        if block.shape[0] != int(global_rows):
            raise ValueError("single-GPU job: the block must hold every row")
        return from_host(block, blocking=blocking)
    from .distributed import PartitionedArray

    thunk = PartitionedArray.from_local_rows(block, global_rows, blocking=blocking)
    return ndarray(shape=thunk.shape, dtype=thunk.dtype, thunk=thunk)


def map_chunks(fn, inputs, outputs, chunk: int) -> None:
    """Out-of-core evaluation of an elementwise program over HOST arrays: `fn(*device_inputs)`
    returns one device array per entry of `outputs`; the 1-D host `inputs` / `outputs` (ideally
    pinned) are processed in chunks of `chunk` elements with the upload of chunk i+1, the kernels
    of chunk i and the download of chunk i-1 in flight at the same time."""
    n = inputs[0].shape[0]
    starts = list(range(0, n, chunk))
    pending: list = []

    def upload(i):
        s = starts[i]
        return [from_host(a[s:s + chunk], blocking=False) for a in inputs]

    nxt = upload(0)
    for i, s in enumerate(starts):
        cur = nxt
        if i + 1 < len(starts):
            nxt = upload(i + 1)
        res = fn(*cur)
        if not isinstance(res, (tuple, list)):
            res = (res,)
        pending.append([r.to_host(o[s:s + chunk], blocking=False) for r, o in zip(res, outputs)])
        del res, cur
        if len(pending) > 2:
            for f in pending.pop(0):
                f.wait()
    for futs in pending:
        for f in futs:
            f.wait()


def flush() -> None:
    """Launch every captured-but-not-yet-issued elementwise task (fusion.py).  Reading a value,
    copying to the host and `synchronize()` do this implicitly."""
    from . import fusion

    fusion.flush()


def synchronize() -> None:
    runtime.synchronize()


# ---------------------------------------------------------------------- shape helpers (views)
def transpose(a, axes=None) -> ndarray:
    a = convert_to_cunumeric_ndarray(a)
    return a.transpose() if axes is None else a.transpose(axes)


def squeeze(a, axis=None) -> ndarray:
    return convert_to_cunumeric_ndarray(a).squeeze(axis)


def reshape(a, newshape, order="C") -> ndarray:
    return convert_to_cunumeric_ndarray(a).reshape(newshape)


def swapaxes(a, axis1, axis2) -> ndarray:
    return convert_to_cunumeric_ndarray(a).swapaxes(axis1, axis2)


def real(val) -> ndarray:
    return convert_to_cunumeric_ndarray(val).real


def imag(val) -> ndarray:
    return convert_to_cunumeric_ndarray(val).imag


def shape(a):
    return convert_to_cunumeric_ndarray(a).shape


def ndim(a) -> int:
    return convert_to_cunumeric_ndarray(a).ndim


# ---------------------------------------------------------------------- WHERE
def where(a, x=None, y=None) -> ndarray:
    """module.py:3104-3150 (three-argument form; the one-argument form is nonzero(), out of
    scope)."""
    if x is None or y is None:
        if x is not None or y is not None:
            raise ValueError("both 'x' and 'y' parameters must be specified together for where")
        raise NotImplementedError("where(condition) == nonzero(condition) is outside the "
                                  "hot-path scope (SURVEY §2.1 row 24)")
    mask = convert_to_cunumeric_ndarray(a)
    xs, ys = x, y
    x = convert_to_cunumeric_ndarray(x, share=True)
    y = convert_to_cunumeric_ndarray(y, share=True)
    # Python scalars are weak, exactly like numpy.where
    common = np.result_type(xs if _is_weak_scalar(xs) else x.dtype,
                            ys if _is_weak_scalar(ys) else y.dtype)
    x = x._maybe_convert(common)
    y = y._maybe_convert(common)
    return ndarray._perform_where(mask, x, y)


where_ = where


# ---------------------------------------------------------------------- dot (1-D inner product)
def dot(a, b, out=None):
    """module.py:4235 `dot`, the vector case: sum(a * b) without conjugation — the reference's DOT
    task (matrix/dot.cu:24-41: a block reduction of lhs * rhs).  Here it is MULTIPLY + SCALAR_UNARY_RED
    through the thunk layer, which the fusion layer runs as ONE map -> reduce kernel: the products
    never reach memory.  0-d operands multiply; matrix products (>= 2-D) belong to the BLAS part of
    the reference, outside the hot-path scope."""
    a, b = convert_to_cunumeric_ndarray(a), convert_to_cunumeric_ndarray(b)
    if a.ndim == 0 or b.ndim == 0:
        from ._ufunc import multiply

        return multiply(a, b, out=out)
    if a.ndim != 1 or b.ndim != 1:
        raise NotImplementedError("dot of arrays with more than one dimension (matrix products) is "
                                  "outside the hot-path scope (SURVEY §2.1: BLAS)")
    if a.shape != b.shape:
        raise ValueError(f"shapes {a.shape} and {b.shape} not aligned")
    common = ndarray.find_common_type(a, b)
    if common == np.bool_:
        result = (a & b).any()
    else:
        prod = a._maybe_convert(common) * b._maybe_convert(common)
        result = prod.sum()
    if out is not None:
        out._thunk.copy(result._thunk, deep=True)
        return out
    return result


# ---------------------------------------------------------------------- reductions
def sum(a, axis=None, dtype=None, out=None, keepdims=False, initial=None, where=None):
    return convert_to_cunumeric_ndarray(a).sum(axis=axis, dtype=dtype, out=out,
                                               keepdims=keepdims, initial=initial, where=where)


def prod(a, axis=None, dtype=None, out=None, keepdims=False, initial=None, where=None):
    return convert_to_cunumeric_ndarray(a).prod(axis=axis, dtype=dtype, out=out,
                                                keepdims=keepdims, initial=initial, where=where)


def amax(a, axis=None, out=None, keepdims=False, initial=None, where=None):
    return convert_to_cunumeric_ndarray(a).max(axis=axis, out=out, keepdims=keepdims,
                                               initial=initial, where=where)


def amin(a, axis=None, out=None, keepdims=False, initial=None, where=None):
    return convert_to_cunumeric_ndarray(a).min(axis=axis, out=out, keepdims=keepdims,
                                               initial=initial, where=where)


max = amax
min = amin


def argmax(a, axis=None, out=None, keepdims=False):
    return convert_to_cunumeric_ndarray(a).argmax(axis=axis, out=out, keepdims=keepdims)


def argmin(a, axis=None, out=None, keepdims=False):
    return convert_to_cunumeric_ndarray(a).argmin(axis=axis, out=out, keepdims=keepdims)


def all(a, axis=None, out=None, keepdims=False, where=None):
    return convert_to_cunumeric_ndarray(a).all(axis=axis, out=out, keepdims=keepdims, where=where)


def any(a, axis=None, out=None, keepdims=False, where=None):
    return convert_to_cunumeric_ndarray(a).any(axis=axis, out=out, keepdims=keepdims, where=where)


def mean(a, axis=None, dtype=None, out=None, keepdims=False):
    return convert_to_cunumeric_ndarray(a).mean(axis=axis, dtype=dtype, out=out, keepdims=keepdims)


def var(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    """module.py:7608 -> ndarray.var."""
    return convert_to_cunumeric_ndarray(a).var(axis=axis, dtype=dtype, out=out, ddof=ddof,
                                               keepdims=keepdims)


def count_nonzero(a, axis=None):
    a = convert_to_cunumeric_ndarray(a)
    return ndarray._perform_unary_reduction(UnaryRedCode.COUNT_NONZERO, a, axis=axis,
                                            res_dtype=np.dtype(np.uint64))


def numpy_compat() -> bool:
    """settings.py:79-89 `CUNUMERIC_NUMPY_COMPATIBILITY`: issue the additional tasks that make
    nanmin / nanmax / nanargmin / nanargmax behave like NumPy on all-NaN slices."""
    import os

    return os.environ.get("CUNUMERIC_NUMPY_COMPATIBILITY", "0").lower() in ("1", "true", "yes", "on")


def _nan_red(op: UnaryRedCode, fallback: UnaryRedCode):
    def fn(a, axis=None, out=None, keepdims=False, initial=None, where=None, dtype=None):
        a = convert_to_cunumeric_ndarray(a)
        code = op if a.dtype.kind in "fc" else fallback
        if a.dtype == np.bool_ and fallback in (UnaryRedCode.SUM, UnaryRedCode.PROD):
            a = a._astype(np.dtype(np.int32), True)
        kwargs = dict(axis=axis, out=out, keepdims=keepdims, initial=initial, where=where)
        if op in (UnaryRedCode.NANSUM, UnaryRedCode.NANPROD):
            kwargs["dtype"] = dtype
        result = ndarray._perform_unary_reduction(code, a, **kwargs)
        if op in (UnaryRedCode.NANMAX, UnaryRedCode.NANMIN) and numpy_compat() and a.dtype.kind == "f":
            # module.py:6022-6024 / 6118-6120: an all-NaN slice yields NaN, as in NumPy (the plain
            # reduction leaves the finite identity there); putmask == where(all_nan, nan, result)
            from ._ufunc import isnan

            all_nan = all(isnan(a), axis=axis, keepdims=keepdims, where=where)
            fixed = where_(all_nan, np.nan, result)
            if out is not None:
                out._thunk.copy(fixed._thunk, deep=True)
                return out
            return fixed
        return result

    return fn


nansum = _nan_red(UnaryRedCode.NANSUM, UnaryRedCode.SUM)
nanprod = _nan_red(UnaryRedCode.NANPROD, UnaryRedCode.PROD)
nanmax = _nan_red(UnaryRedCode.NANMAX, UnaryRedCode.MAX)
nanmin = _nan_red(UnaryRedCode.NANMIN, UnaryRedCode.MIN)


def _nan_argred(op: UnaryRedCode, fallback: UnaryRedCode):
    def fn(a, axis=None, out=None, keepdims=False):
        a = convert_to_cunumeric_ndarray(a)
        if a.size == 0:
            raise ValueError(f"attempt to get {'nanargmax' if op == UnaryRedCode.NANARGMAX else 'nanargmin'}"
                             " of an empty sequence")
        if numpy_compat() and a.dtype.kind == "f":
            # module.py:5850-5852 / 5918-5920
            from ._ufunc import isnan

            if bool(any(all(isnan(a), axis=axis))):
                raise ValueError("Array/Slice contains only NaNs")
        code = op if a.dtype.kind == "f" else fallback
        return a._argred(code, axis, out, keepdims)

    return fn


nanargmax = _nan_argred(UnaryRedCode.NANARGMAX, UnaryRedCode.ARGMAX)
nanargmin = _nan_argred(UnaryRedCode.NANARGMIN, UnaryRedCode.ARGMIN)


def clip(a, a_min, a_max, out=None):
    return convert_to_cunumeric_ndarray(a).clip(a_min, a_max, out=out)


def array_equal(a1, a2, equal_nan: bool = False):
    """module.py:5337-5375 -> BINARY_RED(EQUAL)."""
    from .config import BinaryOpCode

    if equal_nan:
        raise NotImplementedError("cuNumeric does not support `equal_nan` yet for `array_equal`")
    a1, a2 = convert_to_cunumeric_ndarray(a1), convert_to_cunumeric_ndarray(a2)
    if a1.shape != a2.shape:
        return False
    return ndarray._perform_binary_reduction(BinaryOpCode.EQUAL, a1, a2, np.dtype(np.bool_))


def allclose(a, b, rtol=1e-5, atol=1e-8, equal_nan: bool = False):
    """module.py `allclose` -> BINARY_RED(ISCLOSE) with rtol/atol as extra scalars."""
    from .config import BinaryOpCode

    if equal_nan:
        raise NotImplementedError("cuNumeric does not support `equal_nan` yet for allclose")
    a, b = convert_to_cunumeric_ndarray(a), convert_to_cunumeric_ndarray(b)
    return ndarray._perform_binary_reduction(BinaryOpCode.ISCLOSE, a, b, np.dtype(np.bool_),
                                             extra_args=(rtol, atol))


def isclose(a, b, rtol=1e-5, atol=1e-8, equal_nan=False) -> ndarray:
    """module.py `isclose` -> BINARY_OP(ISCLOSE) with rtol/atol as extra scalars."""
    if equal_nan:
        raise NotImplementedError("cuNumeric does not support `equal_nan` yet for isclose")
    a = convert_to_cunumeric_ndarray(a)
    b = convert_to_cunumeric_ndarray(b)
    common = ndarray.find_common_type(a, b)
    a, b = a._maybe_convert(common), b._maybe_convert(common)
    out = ndarray(shape=np.broadcast_shapes(a.shape, b.shape), dtype=np.bool_)
    out._thunk.isclose(a._thunk, b._thunk, rtol, atol, equal_nan)
    return out
