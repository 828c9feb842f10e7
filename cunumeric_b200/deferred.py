"""DeferredArray: the thunk the NumPy-level classes call into.

Keeps the reference's NumPyThunk method surface for the hot path (cunumeric/thunk.py:116-703,
cunumeric/deferred.py: unary_op :3139, unary_reduction :3170, binary_op :3302, where :3367,
convert :1348, copy :392, fill :1463, get_item :962, set_item :1024) but, instead of building a
Legate AutoTask, hands the same per-opcode contract (stores in order + scalar args) to the C ABI in
include/cunumeric_b200.h.  Every call is asynchronous on runtime.stream."""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib, fusion
from .config import BinaryOpCode, ConvertCode, UnaryOpCode, UnaryRedCode, MAX_DIM
from .runtime import runtime
from .store import Store

_ARG_REDS = (UnaryRedCode.ARGMAX, UnaryRedCode.ARGMIN, UnaryRedCode.NANARGMAX,
             UnaryRedCode.NANARGMIN)


def max_identity(ty: np.dtype):
    # deferred.py:183-195 — the value an output store is pre-filled with before a MAX reduction
    if ty.kind in "iu":
        return np.iinfo(ty).min
    if ty.kind == "f":
        return np.finfo(ty).min
    if ty.kind == "b":
        return False
    raise ValueError(f"Unsupported dtype: {ty}")


def min_identity(ty: np.dtype):
    if ty.kind in "iu":
        return np.iinfo(ty).max
    if ty.kind == "f":
        return np.finfo(ty).max
    if ty.kind == "b":
        return True
    raise ValueError(f"Unsupported dtype: {ty}")


# deferred.py:213-238
_UNARY_RED_IDENTITIES = {
    UnaryRedCode.SUM: lambda _: 0,
    UnaryRedCode.SUM_SQUARES: lambda _: 0,
    UnaryRedCode.VARIANCE: lambda _: 0,
    UnaryRedCode.PROD: lambda _: 1,
    UnaryRedCode.MIN: min_identity,
    UnaryRedCode.MAX: max_identity,
    UnaryRedCode.ARGMAX: lambda ty: (np.iinfo(np.int64).min, max_identity(ty)),
    UnaryRedCode.ARGMIN: lambda ty: (np.iinfo(np.int64).min, min_identity(ty)),
    UnaryRedCode.CONTAINS: lambda _: False,
    UnaryRedCode.COUNT_NONZERO: lambda _: 0,
    UnaryRedCode.ALL: lambda _: True,
    UnaryRedCode.ANY: lambda _: False,
    UnaryRedCode.NANARGMAX: lambda ty: (np.iinfo(np.int64).min, max_identity(ty)),
    UnaryRedCode.NANARGMIN: lambda ty: (np.iinfo(np.int64).min, min_identity(ty)),
    UnaryRedCode.NANMAX: max_identity,
    UnaryRedCode.NANMIN: min_identity,
    UnaryRedCode.NANPROD: lambda _: 1,
    UnaryRedCode.NANSUM: lambda _: 0,
}


def _host_scalars(values: Sequence[Any], dtype) -> Optional[np.ndarray]:
    if not values:
        return None
    return np.ascontiguousarray(np.array([np.asarray(v).reshape(()) for v in values], dtype=dtype))


def _vp(arr: Optional[np.ndarray]):
    return None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)


class HostFuture:
    """Completion handle of an asynchronous device->host copy.  Holds the device array alive until
    the copy has finished."""

    def __init__(self, event, keepalive, out: np.ndarray) -> None:
        self._event, self._keepalive, self.out = event, keepalive, out

    def wait(self) -> np.ndarray:
        if self._event is not None:
            _lib.check(runtime.lib.cnb_event_synchronize(self._event))
            runtime._recycle_event(self._event)
            self._event = None
            self._keepalive = None
        return self.out

    def __del__(self) -> None:
        try:
            self.wait()
        except Exception:
            pass


def launch_elementwise(kind: str, op: int, nan_op: int, lhs: Store, ins: Sequence[Store]) -> None:
    """One elementwise task through the per-task C ABI (fusion.py replays chains with this)."""
    lib = runtime.lib
    d_out = lhs.descriptor()
    d_in = [s.descriptor() for s in ins]
    if kind == "U":
        rc = lib.cnb_unary_op(op, ctypes.byref(d_out), None, ctypes.byref(d_in[0]), None, runtime.stream)
    elif kind == "B":
        rc = lib.cnb_binary_op(op, ctypes.byref(d_out), ctypes.byref(d_in[0]), ctypes.byref(d_in[1]),
                               None, runtime.stream)
    elif kind == "W":
        rc = lib.cnb_where(ctypes.byref(d_out), ctypes.byref(d_in[0]), ctypes.byref(d_in[1]),
                           ctypes.byref(d_in[2]), runtime.stream)
    elif kind == "C":
        rc = lib.cnb_convert(nan_op, ctypes.byref(d_out), ctypes.byref(d_in[0]), runtime.stream)
    else:
        raise ValueError(kind)
    _lib.check(rc)


def _rep(thunk):
    """Operand of a replicated task: a partitioned thunk (distributed.PartitionedArray) is gathered
    onto every rank first (collective)."""
    gather = getattr(thunk, "gather", None)
    return thunk if gather is None else gather()


def launch_scalar_red(op, out: Store, src: Store, where: Optional[Store], origin, gshape,
                      args) -> None:
    """SCALAR_UNARY_RED through the C ABI: folds `src` into the 1-element store `out`.  origin /
    gshape place the rect in the global array (arg-reductions return global flat indices)."""
    lhs = out
    while lhs.ndim > 1:
        lhs = lhs.project(0, 0)
    if lhs.ndim == 0:
        lhs = lhs.promote(0, 1)
    d_out, d_in = lhs.descriptor(), src.descriptor()
    d_w = None if where is None else where.descriptor()
    nd = max(src.ndim, 1)
    c_origin = c_shape = None
    if origin is not None:
        c_origin = (ctypes.c_int64 * nd)(*[int(v) for v in origin])
        c_shape = (ctypes.c_int64 * nd)(*[int(v) for v in gshape])
    extra = _host_scalars(args, src.dtype) if args else None
    _lib.check(runtime.lib.cnb_scalar_unary_red(
        int(op), ctypes.byref(d_out), ctypes.byref(d_in),
        None if d_w is None else ctypes.byref(d_w), c_origin, c_shape, _vp(extra),
        runtime.stream))


class DeferredArray:
    __slots__ = ("base", "host_scalar")

    def __init__(self, base: Store, host_scalar: Optional[np.ndarray] = None) -> None:
        self.base = base
        # 0-d operands that came from the host (Python / NumPy scalars) keep their host value, so
        # dtype conversions of scalars happen on the host — the counterpart of the reference
        # routing tiny arrays through its eager NumPy thunk (runtime.py:448-500) instead of
        # launching a one-element CONVERT task.
        self.host_scalar = host_scalar

    # ------------------------------------------------------------------ basic properties
    @property
    def shape(self):
        return self.base.shape

    @property
    def ndim(self) -> int:
        return self.base.ndim

    @property
    def size(self) -> int:
        return self.base.size

    @property
    def dtype(self) -> np.dtype:
        return self.base.dtype

    @property
    def scalar(self) -> bool:
        return self.base.size == 1

    # ------------------------------------------------------------------ host <-> device
    @staticmethod
    def from_numpy(array: np.ndarray) -> "DeferredArray":
        array = np.asarray(array)
        if array.ndim > MAX_DIM:
            raise NotImplementedError(f"at most {MAX_DIM} dimensions are supported")
        src = np.ascontiguousarray(array)
        thunk = DeferredArray(Store.empty(array.shape, array.dtype))
        # pageable sources are staged by the driver before cudaMemcpyAsync returns; pinned sources
        # (runtime.pinned_empty) are read asynchronously and must stay untouched until the stream
        # has passed this point
        runtime.copy_h2d(thunk.base.ptr, src)
        return thunk

    @staticmethod
    def from_numpy_async(array: np.ndarray) -> "DeferredArray":
        """H2D on the copy stream (source should be pinned): overlaps with kernels already queued on
        the compute stream; the first task that uses the result waits for the copy."""
        src = np.ascontiguousarray(array)
        thunk = DeferredArray(Store.empty(array.shape, array.dtype))
        runtime.copy_h2d_async(thunk.base.buffer, src)
        return thunk

    def to_host_async(self, out: np.ndarray) -> "HostFuture":
        """D2H on the copy stream into `out` (should be pinned); returns a future to wait on."""
        src = self
        if not self.base.is_c_contiguous:
            src = DeferredArray(Store.empty(self.shape, self.dtype))
            src.copy(self, deep=True)
        assert out.flags.c_contiguous and out.shape == tuple(self.shape) and out.dtype == self.dtype
        if src.base.buffer.ready_event is not None:
            runtime.wait_ready(src.base.buffer)
        return HostFuture(runtime.copy_d2h_async(out, src.base.ptr, src.base.buffer), src, out)

    def __numpy_array__(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Blocking device->host read (the reference blocks in get_scalar_array / inline mapping)."""
        src = self
        if not self.base.is_c_contiguous:
            src = DeferredArray(Store.empty(self.shape, self.dtype))
            src.copy(self, deep=True)
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        assert out.flags.c_contiguous and out.shape == tuple(self.shape) and out.dtype == self.dtype
        if src.base.buffer.ready_event is not None:
            runtime.wait_ready(src.base.buffer)
        runtime.copy_d2h(out, src.base.ptr)
        runtime.synchronize()
        return out

    # ------------------------------------------------------------------ helpers
    def _copy_if_overlapping(self, other: "DeferredArray") -> "DeferredArray":
        """deferred.py:291-303: a source that aliases the destination (but is not the very same
        window) is materialised first, so `a[1:] += a[:-1]` behaves like NumPy."""
        if not self.base.overlaps(other.base) or self.base.same_window(other.base):
            return other
        copy = DeferredArray(Store.empty(other.shape, other.dtype))
        copy.copy(other, deep=True)
        return copy

    def _broadcast(self, shape) -> Store:
        base = self.base
        return base if base.shape == shape else base.broadcast_to(shape)

    # ------------------------------------------------------------------ UNARY_OP
    def unary_op(self, op: UnaryOpCode, src: "DeferredArray", where: Any = True,
                 args: Sequence[Any] = (), multiout: Optional[Sequence["DeferredArray"]] = None
                 ) -> None:
        lhs = self.base
        src = self._copy_if_overlapping(_rep(src))
        rhs = src._broadcast(lhs.shape)
        out2 = None
        if multiout:
            for extra_out in multiout:
                src = extra_out._copy_if_overlapping(src)
            out2 = multiout[0].base.descriptor()
        extra = _host_scalars(args, src.dtype) if op == UnaryOpCode.CLIP else None
        if out2 is None and extra is None and fusion.capture("U", int(op), lhs, (rhs,)):
            return
        d_out, d_in = lhs.descriptor(), rhs.descriptor()
        _lib.check(runtime.lib.cnb_unary_op(int(op), ctypes.byref(d_out),
                                            None if out2 is None else ctypes.byref(out2),
                                            ctypes.byref(d_in), _vp(extra), runtime.stream))

    # ------------------------------------------------------------------ BINARY_OP
    def binary_op(self, op_code: BinaryOpCode, src1: "DeferredArray", src2: "DeferredArray",
                  where: Any = True, args: Sequence[Any] = ()) -> None:
        lhs = self.base
        src1 = self._copy_if_overlapping(_rep(src1))
        src2 = self._copy_if_overlapping(_rep(src2))
        rhs1 = src1._broadcast(lhs.shape)
        rhs2 = src2._broadcast(lhs.shape)
        extra = _host_scalars(args, np.float64) if op_code == BinaryOpCode.ISCLOSE else None
        if extra is None and fusion.capture("B", int(op_code), lhs, (rhs1, rhs2)):
            return
        d_out, d1, d2 = lhs.descriptor(), rhs1.descriptor(), rhs2.descriptor()
        _lib.check(runtime.lib.cnb_binary_op(int(op_code), ctypes.byref(d_out), ctypes.byref(d1),
                                             ctypes.byref(d2), _vp(extra), runtime.stream))

    def unary_op_prepared(self, op_code: int, rhs: Store) -> None:
        """unary_op for a freshly allocated output of the operand's shape (the ufunc fast path)."""
        lhs = self.base
        if fusion.capture("U", op_code, lhs, (rhs,), 0, True):
            return
        d_out, d_in = lhs.descriptor(), rhs.descriptor()
        _lib.check(runtime.lib.cnb_unary_op(int(op_code), ctypes.byref(d_out), None,
                                            ctypes.byref(d_in), None, runtime.stream))

    def binary_op_prepared(self, op_code: int, rhs1: Store, rhs2: Store) -> None:
        """binary_op for a freshly allocated output and operands already broadcast to its shape
        (the ufunc fast path): nothing to replicate, no aliasing to resolve, no extra scalars."""
        lhs = self.base
        if fusion.capture("B", op_code, lhs, (rhs1, rhs2), 0, True):
            return
        d_out, d1, d2 = lhs.descriptor(), rhs1.descriptor(), rhs2.descriptor()
        _lib.check(runtime.lib.cnb_binary_op(int(op_code), ctypes.byref(d_out), ctypes.byref(d1),
                                             ctypes.byref(d2), None, runtime.stream))

    def isclose(self, rhs1, rhs2, rtol: float, atol: float, equal_nan: bool) -> None:
        assert not equal_nan
        self.binary_op(BinaryOpCode.ISCLOSE, rhs1, rhs2, True, (rtol, atol))

    # ------------------------------------------------------------------ BINARY_RED
    def binary_reduction(self, op: BinaryOpCode, src1: "DeferredArray", src2: "DeferredArray",
                         broadcast: Any, args: Sequence[Any]) -> None:
        """deferred.py:3330-3364: `self` (a bool scalar) <- all(op(src1, src2))."""
        if hasattr(src1, "gather") or hasattr(src2, "gather"):
            from .distributed import partitioned_binary_reduction

            if partitioned_binary_reduction(self, op, src1, src2, broadcast, args):
                return
        src1, src2 = _rep(src1), _rep(src2)
        rhs1, rhs2 = src1.base, src2.base
        if broadcast is not None:
            rhs1, rhs2 = rhs1.broadcast_to(broadcast), rhs2.broadcast_to(broadcast)
        self.fill(np.array(True))
        lhs = self.base
        while lhs.ndim > 1:
            lhs = lhs.project(0, 0)
        if lhs.ndim == 0:
            lhs = lhs.promote(0, 1)
        extra = _host_scalars(args, np.float64) if op == BinaryOpCode.ISCLOSE else None
        d_out, d1, d2 = lhs.descriptor(), rhs1.descriptor(), rhs2.descriptor()
        _lib.check(runtime.lib.cnb_binary_red(int(op), ctypes.byref(d_out), ctypes.byref(d1),
                                              ctypes.byref(d2), _vp(extra), runtime.stream))

    # ------------------------------------------------------------------ WHERE
    def where(self, mask: "DeferredArray", one: "DeferredArray", two: "DeferredArray") -> None:
        lhs = self.base
        m = self._copy_if_overlapping(_rep(mask))._broadcast(lhs.shape)
        a = self._copy_if_overlapping(_rep(one))._broadcast(lhs.shape)
        b = self._copy_if_overlapping(_rep(two))._broadcast(lhs.shape)
        if fusion.capture("W", 0, lhs, (m, a, b)):
            return
        d_out, dm, da, db = lhs.descriptor(), m.descriptor(), a.descriptor(), b.descriptor()
        _lib.check(runtime.lib.cnb_where(ctypes.byref(d_out), ctypes.byref(dm), ctypes.byref(da),
                                         ctypes.byref(db), runtime.stream))

    # ------------------------------------------------------------------ CONVERT
    def convert(self, rhs: "DeferredArray", warn: bool = True,
                nan_op: ConvertCode = ConvertCode.NOOP, temporary: bool = False) -> None:
        lhs = self.base
        rhs = _rep(rhs)
        if rhs.dtype == lhs.dtype:
            self.copy(rhs, deep=True)
            return
        rhs = self._copy_if_overlapping(rhs)
        src = rhs._broadcast(lhs.shape)
        if fusion.capture("C", 0, lhs, (src,), int(nan_op)):
            return
        d_out, d_in = lhs.descriptor(), src.descriptor()
        _lib.check(runtime.lib.cnb_convert(int(nan_op), ctypes.byref(d_out), ctypes.byref(d_in),
                                           runtime.stream))

    # ------------------------------------------------------------------ COPY / FILL
    def copy(self, rhs: "DeferredArray", deep: bool = False) -> None:
        """deferred.py:392-401: UNARY_OP(COPY) from rhs into this window."""
        rhs = _rep(rhs)
        if self.base.same_window(rhs.base):
            return
        self.unary_op(UnaryOpCode.COPY, rhs, True, ())

    def fill(self, value: Any) -> None:
        """deferred.py:1463-1496 (FILL task): value is a 0-d host array of this dtype."""
        val = np.ascontiguousarray(np.asarray(value, dtype=self.dtype).reshape(()))
        d_out = self.base.descriptor()
        _lib.check(runtime.lib.cnb_fill(ctypes.byref(d_out), _vp(val), runtime.stream))

    # ------------------------------------------------------------------ reductions
    def unary_reduction(self, op: UnaryRedCode, src: "DeferredArray",
                        where: Optional["DeferredArray"], orig_axis: Optional[int],
                        axes: Optional[Sequence[int]], keepdims: bool, args: Any,
                        initial: Any) -> None:
        """deferred.py:3170-3288.  `self` is the (pre-existing) result thunk."""
        if hasattr(src, "reduce_into"):  # row-partitioned source: local partial + NCCL combine
            src.reduce_into(self, op, where, orig_axis, axes, keepdims, args, initial)
            return
        where = None if where is None else _rep(where)
        lhs_array = self
        rhs_array = src
        argred = op in _ARG_REDS
        if argred:
            argred_dtype = runtime.get_argred_type(rhs_array.dtype)
            lhs_array = DeferredArray(Store.empty(self.shape, argred_dtype))
        if initial is not None:
            assert not argred
            fill_value = initial
        else:
            fill_value = _UNARY_RED_IDENTITIES[op](rhs_array.dtype)
        # map -> reduce fusion: a full reduction of a value the open fused chain produces joins the
        # chain (fusion.capture_reduce); the pre-fill travels with it as the value to fold into
        if lhs_array.size == 1 and where is None and not argred and not args and \
                type(rhs_array) is DeferredArray and \
                fusion.capture_reduce(int(op), lhs_array.base, rhs_array.base, fill_value):
            return
        lhs_array.fill(np.array(fill_value, dtype=lhs_array.dtype))

        d_in = rhs_array.base.descriptor()
        d_where = None
        if where is not None:
            d_where = where._broadcast(rhs_array.shape).descriptor()
        p_where = None if d_where is None else ctypes.byref(d_where)

        if lhs_array.size == 1:
            lhs = lhs_array.base
            while lhs.ndim > 1:
                lhs = lhs.project(0, 0)
            if lhs.ndim == 0:
                lhs = lhs.promote(0, 1)
            launch_scalar_red(op, lhs, rhs_array.base,
                              None if where is None else where._broadcast(rhs_array.shape),
                              None, None, args)
        else:
            assert axes is not None
            if len(axes) > 1:
                raise NotImplementedError("Need support for reducing multiple dimensions")
            axis = axes[0]
            result = lhs_array.base
            if keepdims:
                result = result.project(axis, 0)
            result = result.promote(axis, rhs_array.shape[axis])
            d_out = result.descriptor()
            _lib.check(runtime.lib.cnb_unary_red(int(op), int(axis), ctypes.byref(d_out),
                                                 ctypes.byref(d_in), p_where, 0, runtime.stream))
        if argred:
            self.unary_op(UnaryOpCode.GETARG, lhs_array, True, ())

    # ------------------------------------------------------------------ views
    def get_item(self, key: Any) -> "DeferredArray":
        return DeferredArray(_basic_index(self.base, key))

    def set_item(self, key: Any, rhs: "DeferredArray") -> None:
        view = DeferredArray(_basic_index(self.base, key))
        if view.base.size == 0:
            return
        if rhs.dtype != view.dtype:
            tmp = DeferredArray(Store.empty(rhs.shape, view.dtype))
            tmp.convert(rhs)
            rhs = tmp
        view.copy(rhs, deep=False)

    def transpose(self, axes: Sequence[int]) -> "DeferredArray":
        return DeferredArray(self.base.transpose(axes))

    def swapaxes(self, a: int, b: int) -> "DeferredArray":
        axes = list(range(self.ndim))
        axes[a], axes[b] = axes[b], axes[a]
        return self.transpose(axes)

    def squeeze(self, axis=None) -> "DeferredArray":
        if axis is None:
            drop = [d for d, n in enumerate(self.shape) if n == 1]
        else:
            drop = [axis % self.ndim] if isinstance(axis, int) else [a % self.ndim for a in axis]
            for d in drop:
                if self.shape[d] != 1:
                    raise ValueError("cannot select an axis to squeeze out which has size "
                                     "not equal to one")
        store = self.base
        for d in sorted(drop, reverse=True):
            store = store.project(d, 0)
        return DeferredArray(store)

    def reshape(self, newshape: Sequence[int]) -> "DeferredArray":
        if self.base.is_c_contiguous:
            return DeferredArray(self.base.reshape_contiguous(newshape))
        tmp = DeferredArray(Store.empty(self.shape, self.dtype))
        tmp.copy(self, deep=True)
        return DeferredArray(tmp.base.reshape_contiguous(newshape))

    def real_imag_view(self, part: int) -> "DeferredArray":
        """Zero-copy .real / .imag of a complex array."""
        assert self.dtype.kind == "c"
        rdt = np.dtype(np.float32 if self.dtype == np.complex64 else np.float64)
        s = self.base
        return DeferredArray(Store(s.buffer, rdt, s.shape, s.strides,
                                   s.offset + part * rdt.itemsize))


def _basic_index(store: Store, key: Any) -> Store:
    """Basic (view) indexing: ints, slices, None, Ellipsis (deferred.py:904-924 `_get_view`)."""
    if not isinstance(key, tuple):
        key = (key,)
    n_specified = sum(1 for k in key if k is not None and k is not Ellipsis)
    if n_specified > store.ndim:
        raise IndexError(f"too many indices for array: array is {store.ndim}-dimensional")
    if sum(1 for k in key if k is Ellipsis) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    expanded = []
    for k in key:
        if k is Ellipsis:
            expanded.extend([slice(None)] * (store.ndim - n_specified))
        else:
            expanded.append(k)
    if not any(k is Ellipsis for k in key):
        expanded.extend([slice(None)] * (store.ndim - n_specified))
    dim = 0
    for k in expanded:
        if k is None:
            store = store.promote(dim, 1)
            dim += 1
        elif isinstance(k, slice):
            store = store.slice(dim, k)
            dim += 1
        elif isinstance(k, (int, np.integer)):
            idx = int(k)
            if idx < 0:
                idx += store.shape[dim]
            if not (0 <= idx < store.shape[dim]):
                raise IndexError(f"index {int(k)} is out of bounds for axis {dim} "
                                 f"with size {store.shape[dim]}")
            store = store.project(dim, idx)
        else:
            raise NotImplementedError(
                "cunumeric_b200 implements basic (view) indexing only; advanced indexing is "
                "outside the hot-path scope (SURVEY §2.1 row 23)")
    return store
