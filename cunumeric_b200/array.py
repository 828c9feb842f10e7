"""ndarray: the NumPy-compatible array class in front of the hot path.

Covers the parts of the reference's cunumeric/array.py that route into BINARY_OP / UNARY_OP / WHERE /
CONVERT / UNARY_RED / SCALAR_UNARY_RED: operator dunders -> ufuncs (array.py:801-812), astype
(:1900-1989), sum/prod/max/min/argmax/argmin/all/any (+nan*), _perform_unary_reduction
(:4323-4418), _perform_where (:4455-4470), __getitem__/__setitem__ with basic (view) indexing
(:1031-1037, :1668-1680).  Everything else of the 4.5 kLoC class is out of scope."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from .config import UnaryOpCode, UnaryRedCode, is_supported_dtype
from .deferred import DeferredArray
from .distributed import create_empty_thunk, thunk_from_numpy
from .runtime import runtime
from .store import Store


# reductions that can run one axis at a time -> the op that continues over the partial results
_SEPARABLE_REDS = {
    UnaryRedCode.SUM: UnaryRedCode.SUM, UnaryRedCode.PROD: UnaryRedCode.PROD,
    UnaryRedCode.MAX: UnaryRedCode.MAX, UnaryRedCode.MIN: UnaryRedCode.MIN,
    UnaryRedCode.ALL: UnaryRedCode.ALL, UnaryRedCode.ANY: UnaryRedCode.ANY,
    UnaryRedCode.COUNT_NONZERO: UnaryRedCode.SUM,
}


def _normalize_axis_tuple(axis, ndim: int) -> tuple:
    if isinstance(axis, (int, np.integer)):
        axis = (int(axis),)
    out = []
    for a in axis:
        a = int(a)
        if not (-ndim <= a < ndim):
            raise np.exceptions.AxisError(a, ndim)
        out.append(a % ndim)
    if len(set(out)) != len(out):
        raise ValueError("repeated axis")
    return tuple(out)


def convert_to_cunumeric_ndarray(obj: Any, share: bool = False) -> "ndarray":
    """array.py `convert_to_cunumeric_ndarray`: anything array-like becomes a device ndarray."""
    if isinstance(obj, ndarray):
        return obj
    if isinstance(obj, (bool, int, float, complex)) and not isinstance(obj, np.generic):
        host = np.asarray(obj)
        # Python ints/floats are weak: they adopt int64/float64/complex128 only as a placeholder;
        # binary_ufunc._find_common_type resolves the real type
    else:
        host = np.asarray(obj)
    if host.dtype == object or not is_supported_dtype(host.dtype):
        raise TypeError(f"cunumeric_b200 does not support dtype={host.dtype}")
    if host.ndim == 0:
        if share:
            # internal, read-only operand (a scalar next to an array in a ufunc call): a window onto
            # the value-keyed constant cache (runtime.scalar_buffer), with its host value attached
            return ndarray(shape=(), dtype=host.dtype,
                           thunk=DeferredArray(Store.from_scalar(host), host_scalar=host))
        # a user-visible 0-d array owns its buffer: writing to it (`x += 1`, fill, out=x) must never
        # reach the shared constants, and it carries no host copy that could go stale
        store = Store.empty((), host.dtype)
        if not runtime.dry_run:
            runtime.copy_h2d(store.ptr, np.ascontiguousarray(host))
        return ndarray(shape=(), dtype=host.dtype, thunk=DeferredArray(store))
    return ndarray(shape=host.shape, dtype=host.dtype, thunk=thunk_from_numpy(host))


def _rebuild_from_host(host: np.ndarray) -> "ndarray":
    return convert_to_cunumeric_ndarray(host)


def broadcast_where(where, shape):
    if where is None or where is True:
        return None
    where = convert_to_cunumeric_ndarray(where)
    if where.dtype != np.bool_:
        where = where.astype(np.bool_)
    if where.shape != tuple(shape):
        np.broadcast_shapes(where.shape, tuple(shape))  # raises if incompatible
        where = ndarray(shape=shape, dtype=np.bool_,
                        thunk=DeferredArray(where._thunk._broadcast(shape)))
    return where


class ndarray:
    __array_priority__ = 100.0

    def __init__(self, shape, dtype=np.float64, buffer=None, offset=0, strides=None, order=None,
                 thunk: Optional[DeferredArray] = None, inputs=None) -> None:
        if thunk is None:
            if buffer is not None or strides is not None:
                raise NotImplementedError("ndarray(buffer=/strides=) is not supported")
            dtype = np.dtype(dtype)
            if not is_supported_dtype(dtype) and dtype.names is None:
                raise TypeError(f"cunumeric_b200 does not support dtype={dtype}")
            if isinstance(shape, (int, np.integer)):
                shape = (int(shape),)
            thunk = create_empty_thunk(tuple(shape), dtype, inputs)
        self._thunk = thunk
        self._writeback: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ properties
    @property
    def shape(self):
        return self._thunk.shape

    @property
    def dtype(self) -> np.dtype:
        return self._thunk.dtype

    @property
    def ndim(self) -> int:
        return self._thunk.ndim

    @property
    def size(self) -> int:
        return self._thunk.size

    @property
    def itemsize(self) -> int:
        return self.dtype.itemsize

    @property
    def nbytes(self) -> int:
        return self.size * self.itemsize

    @property
    def strides(self):
        return self._thunk.base.strides

    @property
    def T(self) -> "ndarray":
        return self.transpose()

    @property
    def real(self) -> "ndarray":
        if self.dtype.kind == "c":
            return ndarray(self.shape, thunk=self._thunk.real_imag_view(0))
        return self

    @property
    def imag(self) -> "ndarray":
        if self.dtype.kind == "c":
            return ndarray(self.shape, thunk=self._thunk.real_imag_view(1))
        out = ndarray(self.shape, self.dtype)
        out.fill(0)
        return out

    # ------------------------------------------------------------------ host interop
    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        host = self._thunk.__numpy_array__()
        if dtype is not None and np.dtype(dtype) != host.dtype:
            host = host.astype(dtype)
        return host

    def to_host(self, out: Optional[np.ndarray] = None, blocking: bool = True):
        """D2H copy, optionally into a caller-provided (e.g. pinned) buffer.  blocking=False starts
        the copy on the copy stream and returns a future (`.wait()` gives the host array)."""
        if blocking:
            return self._thunk.__numpy_array__(out)
        if out is None:
            from .runtime import runtime

            out = runtime.pinned_empty(self.shape, self.dtype)
        if not isinstance(self._thunk, DeferredArray):
            raise NotImplementedError("asynchronous to_host of a partitioned array")
        return self._thunk.to_host_async(out)

    def to_host_rows(self, out: Optional[np.ndarray] = None, blocking: bool = True):
        """Counterpart of from_host_rows: THIS rank's block of rows (the whole array on one GPU),
        copied into `out` (e.g. pinned) without any inter-GPU traffic.  blocking=False starts the
        copy on the D2H stream and returns a future (`.wait()`)."""
        local = getattr(self._thunk, "local_block_to_host", None)
        if local is not None:
            return local(out, blocking)
        return self.to_host(out, blocking)

    def item(self, *args):
        return self.__array__().item(*args)

    def tolist(self):
        return self.__array__().tolist()

    def __repr__(self) -> str:
        return repr(self.__array__())

    def __str__(self) -> str:
        return str(self.__array__())

    def __len__(self) -> int:
        if self.ndim == 0:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __bool__(self) -> bool:
        return bool(self.__array__())

    def __int__(self) -> int:
        return int(self.__array__())

    def __float__(self) -> float:
        return float(self.__array__())

    def __complex__(self) -> complex:
        return complex(self.__array__())

    def __index__(self) -> int:
        return self.__array__().__index__()

    # ------------------------------------------------------------------ views / shape
    def __getitem__(self, key: Any) -> "ndarray":
        thunk = self._thunk.get_item(key)
        return ndarray(thunk.shape, thunk=thunk)

    def __setitem__(self, key: Any, value: Any) -> None:
        """array.py:1668-1680 -> deferred.set_item: `view[:] = value` is a UNARY_OP(COPY)."""
        value = convert_to_cunumeric_ndarray(value, share=True)
        if value.dtype != self.dtype:
            value = value._astype(self.dtype, temporary=True)
        view = self._thunk.get_item(key)
        np.broadcast_shapes(value.shape, view.shape)
        if value.ndim > view.ndim:
            raise ValueError(f"could not broadcast input array from shape {value.shape} into "
                             f"shape {view.shape}")
        self._thunk.set_item(key, value._thunk)

    def transpose(self, *axes) -> "ndarray":
        if len(axes) == 0 or (len(axes) == 1 and axes[0] is None):
            perm = tuple(reversed(range(self.ndim)))
        elif len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            perm = _normalize_axis_tuple(axes[0], self.ndim)
        else:
            perm = _normalize_axis_tuple(axes, self.ndim)
        if len(perm) != self.ndim:
            raise ValueError("axes don't match array")
        return ndarray(None, thunk=self._thunk.transpose(perm))

    def swapaxes(self, a: int, b: int) -> "ndarray":
        return ndarray(None, thunk=self._thunk.swapaxes(a % self.ndim, b % self.ndim))

    def squeeze(self, axis=None) -> "ndarray":
        return ndarray(None, thunk=self._thunk.squeeze(axis))

    def reshape(self, *shape, order="C") -> "ndarray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape = tuple(self.size // max(known, 1) if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {shape}")
        return ndarray(None, thunk=self._thunk.reshape(shape))

    def ravel(self, order="C") -> "ndarray":
        return self.reshape(-1)

    def flatten(self, order="C") -> "ndarray":
        return self.reshape(-1).copy()

    def copy(self, order="C") -> "ndarray":
        out = ndarray(self.shape, self.dtype, inputs=(self,))
        out._thunk.copy(self._thunk, deep=True)
        return out

    __copy__ = copy

    def __deepcopy__(self, memo=None) -> "ndarray":
        """array.py:905 — a deep copy is a device copy (never a walk through Store / DeviceBuffer)."""
        return self.copy()

    def __reduce__(self):
        """array.py:1513 — pickling goes through the host array."""
        return (_rebuild_from_host, (self.__array__(),))

    def fill(self, value: Any) -> None:
        self._thunk.fill(np.asarray(value).astype(self.dtype))

    # ------------------------------------------------------------------ astype (CONVERT)
    def astype(self, dtype, order="C", casting="unsafe", subok=True, copy=True) -> "ndarray":
        """array.py:1900-1960."""
        dtype = np.dtype(dtype)
        if self.dtype == dtype:
            return self.copy() if copy else self
        casting_allowed = np.can_cast(self.dtype, dtype, casting)
        if not casting_allowed:
            raise TypeError(f"Cannot cast array data from '{self.dtype}' to '{dtype}' according "
                            f"to the rule '{casting}'")
        if self.dtype.kind == "c" and dtype.kind != "c":
            import warnings

            warnings.warn("Casting complex values to real discards the imaginary part",
                          np.exceptions.ComplexWarning, stacklevel=2)
        return self._astype(dtype, False)

    def _astype(self, dtype, temporary: bool = False) -> "ndarray":
        """array.py:1962-1989."""
        dtype = np.dtype(dtype)
        if self.dtype == dtype:
            return self
        if self._thunk.host_scalar is not None:
            with np.errstate(all="ignore"):
                return convert_to_cunumeric_ndarray(self._thunk.host_scalar.astype(dtype), share=True)
        result = ndarray(self.shape, dtype=dtype, inputs=(self,))
        result._thunk.convert(self._thunk, warn=False, temporary=temporary)
        return result

    def _maybe_convert(self, dtype, hints=None) -> "ndarray":
        if self.dtype == dtype:
            return self
        if self._thunk.host_scalar is not None:
            return self._astype(dtype)
        copy = ndarray(shape=self.shape, dtype=dtype, inputs=hints)
        copy._thunk.convert(self._thunk)
        return copy

    @staticmethod
    def find_common_type(*args) -> np.dtype:
        """array.py `find_common_type`: 0-d operands only break ties within a kind."""
        array_types = [a.dtype for a in args if a.ndim > 0]
        scalar_types = [a.dtype for a in args if a.ndim == 0]
        if not array_types:
            return np.result_type(*scalar_types)
        return np.result_type(*array_types, *scalar_types)

    # ------------------------------------------------------------------ operators -> ufuncs
    def _binop(name):  # noqa: N805
        cache = []

        def uf():
            if not cache:
                from . import _ufunc

                cache.append(getattr(_ufunc, name))
            return cache[0]

        def fwd(self, rhs):
            return uf()(self, rhs)

        def rev(self, lhs):
            return uf()(lhs, self)

        def inplace(self, rhs):
            return uf()(self, rhs, out=self)

        return fwd, rev, inplace

    __add__, __radd__, __iadd__ = _binop("add")
    __sub__, __rsub__, __isub__ = _binop("subtract")
    __mul__, __rmul__, __imul__ = _binop("multiply")
    __truediv__, __rtruediv__, __itruediv__ = _binop("true_divide")
    __floordiv__, __rfloordiv__, __ifloordiv__ = _binop("floor_divide")
    __mod__, __rmod__, __imod__ = _binop("remainder")
    __pow__, __rpow__, __ipow__ = _binop("power")
    __and__, __rand__, __iand__ = _binop("bitwise_and")
    __or__, __ror__, __ior__ = _binop("bitwise_or")
    __xor__, __rxor__, __ixor__ = _binop("bitwise_xor")
    __lshift__, __rlshift__, __ilshift__ = _binop("left_shift")
    __rshift__, __rrshift__, __irshift__ = _binop("right_shift")
    __eq__, _, _ = _binop("equal")
    __ne__, _, _ = _binop("not_equal")
    __lt__, _, _ = _binop("less")
    __le__, _, _ = _binop("less_equal")
    __gt__, _, _ = _binop("greater")
    __ge__, _, _ = _binop("greater_equal")
    del _binop, _
    __hash__ = None  # type: ignore

    def __neg__(self):
        from . import _ufunc

        if self.dtype == np.bool_ or self.dtype.kind not in "iufc":
            raise TypeError("The numpy boolean negative, the `-` operator, is not supported")
        return _ufunc.negative(self)

    def __pos__(self):
        from . import _ufunc

        return _ufunc.positive(self)

    def __abs__(self):
        from . import _ufunc

        return _ufunc.absolute(self)

    def __invert__(self):
        from . import _ufunc

        return _ufunc.invert(self)

    def conj(self):
        from . import _ufunc

        return _ufunc.conjugate(self) if self.dtype.kind == "c" else self

    conjugate = conj

    def clip(self, min=None, max=None, out=None) -> "ndarray":
        """array.py:2223-2290: UNARY_OP(CLIP) with min/max as scalar arguments."""
        if min is None and max is None:
            raise ValueError("One of max or min must be given")
        info = (np.iinfo(self.dtype) if self.dtype.kind in "iu" else
                np.finfo(self.dtype) if self.dtype.kind == "f" else None)
        lo = min if min is not None else (info.min if info else False)
        hi = max if max is not None else (info.max if info else True)
        for bound in (lo, hi):
            if np.ndim(bound) != 0:
                raise NotImplementedError("clip with array bounds is outside the hot-path scope")
        result = out if (out is not None and out.dtype == self.dtype) else ndarray(self.shape,
                                                                                  self.dtype)
        result._thunk.unary_op(UnaryOpCode.CLIP, self._thunk, True,
                               (np.array(lo, dtype=self.dtype), np.array(hi, dtype=self.dtype)))
        if out is not None and out is not result:
            out._thunk.convert(result._thunk)
            return out
        return result

    # ------------------------------------------------------------------ reductions
    def sum(self, axis=None, dtype=None, out=None, keepdims=False, initial=None, where=None):
        """array.py:3793-3838."""
        src = self
        if self.dtype == np.bool_:
            # temp bool->int conversion (array.py:3818-3826)
            src = self._astype(np.dtype(np.int32) if dtype is None else np.dtype(dtype), True)
        return ndarray._perform_unary_reduction(UnaryRedCode.SUM, src, axis=axis, dtype=dtype,
                                                out=out, keepdims=keepdims, initial=initial,
                                                where=where)

    def prod(self, axis=None, dtype=None, out=None, keepdims=False, initial=None, where=None):
        src = self
        if self.dtype == np.bool_:
            src = self._astype(np.dtype(np.int32) if dtype is None else np.dtype(dtype), True)
        return ndarray._perform_unary_reduction(UnaryRedCode.PROD, src, axis=axis, dtype=dtype,
                                                out=out, keepdims=keepdims, initial=initial,
                                                where=where)

    def max(self, axis=None, out=None, keepdims=False, initial=None, where=None):
        return ndarray._perform_unary_reduction(UnaryRedCode.MAX, self, axis=axis, out=out,
                                                keepdims=keepdims, initial=initial, where=where)

    def min(self, axis=None, out=None, keepdims=False, initial=None, where=None):
        return ndarray._perform_unary_reduction(UnaryRedCode.MIN, self, axis=axis, out=out,
                                                keepdims=keepdims, initial=initial, where=where)

    def _argred(self, op, axis, out, keepdims):
        if out is not None and out.dtype != np.int64:
            raise ValueError("output array must have int64 dtype")
        if axis is not None and not isinstance(axis, (int, np.integer)):
            raise ValueError("axis must be an integer")
        return ndarray._perform_unary_reduction(op, self, axis=axis, res_dtype=np.dtype(np.int64),
                                                out=out, keepdims=keepdims)

    def argmax(self, axis=None, out=None, keepdims=False):
        return self._argred(UnaryRedCode.ARGMAX, axis, out, keepdims)

    def argmin(self, axis=None, out=None, keepdims=False):
        return self._argred(UnaryRedCode.ARGMIN, axis, out, keepdims)

    def all(self, axis=None, out=None, keepdims=False, initial=None, where=None):
        return ndarray._perform_unary_reduction(UnaryRedCode.ALL, self, axis=axis,
                                                res_dtype=np.dtype(np.bool_), out=out,
                                                keepdims=keepdims, initial=initial, where=where)

    def any(self, axis=None, out=None, keepdims=False, initial=None, where=None):
        return ndarray._perform_unary_reduction(UnaryRedCode.ANY, self, axis=axis,
                                                res_dtype=np.dtype(np.bool_), out=out,
                                                keepdims=keepdims, initial=initial, where=where)

    def dot(self, rhs, out=None):
        from .module import dot

        return dot(self, rhs, out=out)

    def mean(self, axis=None, dtype=None, out=None, keepdims=False):
        """array.py:3146-3203: SUM followed by a true_divide."""
        if axis is not None and not isinstance(axis, (int, np.integer)):
            raise NotImplementedError("mean only supports int types for 'axis' currently")
        if dtype is None:
            dtype = np.dtype(np.float64) if self.dtype.kind in "biu" else self.dtype
        dtype = np.dtype(dtype)
        total = self.sum(axis=axis, dtype=dtype, keepdims=keepdims)
        divisor = self.size if axis is None else self.shape[int(axis) % max(self.ndim, 1)]
        from . import _ufunc

        result = _ufunc.true_divide(total, np.array(divisor, dtype=total.dtype))
        if out is not None:
            out._thunk.convert(result._thunk)
            return out
        return result

    def var(self, axis=None, dtype=None, out=None, ddof: int = 0, keepdims: bool = False):
        """array.py:3234-3323: two passes — the mean first, then <(x-mu)^2> directly (VARIANCE as one
        scalar reduction when the result is a scalar, else `delta = x - mu` + SUM_SQUARES along the
        axis) — and a division by (count - ddof)."""
        if axis is not None and not isinstance(axis, (int, np.integer)):
            raise NotImplementedError("cunumeric.var only supports int types for `axis` currently")
        if self.dtype.kind == "c":
            raise NotImplementedError("var of complex arrays is not supported")
        if dtype is None:
            dtype = np.dtype(np.float64) if self.dtype.kind in "biu" else self.dtype
        dtype = np.dtype(dtype)
        src = self if self.dtype == dtype else self._astype(dtype)
        mu = src.mean(axis=axis, dtype=dtype, keepdims=True)
        if axis is None:
            count, others = self.size, 1
        else:
            ax = int(axis) % max(self.ndim, 1)
            count = self.shape[ax]
            others = self.size // max(count, 1) if count else 0
        if axis is None or others == 1:
            mu_host = np.asarray(mu.__array__()).reshape(())  # the task takes mu as a scalar future
            result = ndarray._perform_unary_reduction(UnaryRedCode.VARIANCE, src, axis=axis,
                                                      dtype=dtype, keepdims=keepdims,
                                                      args=(mu_host,))
        else:
            delta = src - mu
            result = ndarray._perform_unary_reduction(UnaryRedCode.SUM_SQUARES, delta, axis=axis,
                                                      dtype=dtype, keepdims=keepdims)
        from . import _ufunc

        result = _ufunc.true_divide(result, np.array(count - ddof, dtype=dtype))
        if out is not None:
            out._thunk.convert(result._thunk)
            return out
        return result

    @classmethod
    def _perform_unary_reduction(cls, op: UnaryRedCode, src: "ndarray", axis: Any = None,
                                 dtype=None, res_dtype=None, out: Optional["ndarray"] = None,
                                 keepdims: bool = False, args: Any = None, initial: Any = None,
                                 where: Any = None) -> "ndarray":
        """array.py:4323-4418."""
        if res_dtype is not None:
            assert dtype is None
            dtype = src.dtype
        else:
            if dtype is not None:
                dtype = np.dtype(dtype)
                res_dtype = dtype
            elif out is not None:
                dtype = out.dtype
                res_dtype = out.dtype
            else:
                dtype = src.dtype
                res_dtype = src.dtype
        if op in (UnaryRedCode.ARGMAX, UnaryRedCode.ARGMIN, UnaryRedCode.MAX,
                  UnaryRedCode.MIN) and src.dtype.kind == "c":
            raise NotImplementedError("(arg)max/min not supported for complex-type arrays")
        if axis is None:
            axes = tuple(range(src.ndim))
        else:
            axes = _normalize_axis_tuple(axis, src.ndim)
        out_shape: tuple = ()
        for dim in range(src.ndim):
            if dim not in axes:
                out_shape += (src.shape[dim],)
            elif keepdims:
                out_shape += (1,)
        if 1 < len(axes) < src.ndim and op in _SEPARABLE_REDS and not args and \
                type(src._thunk) is DeferredArray:
            # Several (not all) axes.  The reference stops here (deferred.py:3259-3262 "Need support
            # for reducing multiple dimensions"); these reductions are separable, so they run as one
            # UNARY_RED per axis, innermost first: `where` applies to the first pass, `initial`
            # joins the last one, and COUNT_NONZERO continues as a SUM of the counts.
            cur = src
            order = sorted(axes, reverse=True)
            for i, ax in enumerate(order):
                first, last = i == 0, i == len(order) - 1
                if first:
                    cur = cls._perform_unary_reduction(
                        op, cur, axis=ax, keepdims=True, where=where,
                        initial=initial if last else None,
                        **({"res_dtype": res_dtype} if dtype == src.dtype and res_dtype != dtype
                           else {"dtype": dtype}))
                else:
                    cur = cls._perform_unary_reduction(
                        _SEPARABLE_REDS[op], cur, axis=ax, keepdims=True, res_dtype=cur.dtype,
                        initial=initial if last else None)
            cur = cur.reshape(out_shape)
            if out is None:
                return cur
            if out.shape != out_shape:
                raise ValueError(f"the output shapes do not match: expected {out_shape} but got "
                                 f"{out.shape}")
            out._thunk.convert(cur._thunk)
            return out
        if out is None:
            out = ndarray(shape=out_shape, dtype=res_dtype, inputs=(src, where))
        elif out.shape != out_shape:
            raise ValueError(f"the output shapes do not match: expected {out_shape} but got "
                             f"{out.shape}")
        if dtype != src.dtype:
            src = src.astype(dtype)
        if out.dtype == res_dtype:
            result = out
        else:
            result = ndarray(shape=out_shape, dtype=res_dtype, inputs=(src, where))
        where_array = broadcast_where(where, src.shape)
        result._thunk.unary_reduction(op, src._thunk,
                                      None if where_array is None else where_array._thunk,
                                      axis, axes, keepdims, args, initial)
        if result is not out:
            out._thunk.convert(result._thunk)
        return out

    @classmethod
    def _perform_binary_reduction(cls, op, one: "ndarray", two: "ndarray", dtype,
                                  extra_args=None) -> "ndarray":
        """array.py:4420-4452."""
        assert dtype is not None and np.dtype(dtype) == np.bool_
        broadcast = None
        if one.shape != two.shape:
            broadcast = np.broadcast_shapes(one.shape, two.shape)
        common_type = cls.find_common_type(one, two)
        one_thunk = one._maybe_convert(common_type)._thunk
        two_thunk = two._maybe_convert(common_type)._thunk
        dst = ndarray(shape=(), dtype=np.bool_)
        dst._thunk.binary_reduction(op, one_thunk, two_thunk, broadcast, extra_args or ())
        return dst

    @classmethod
    def _perform_where(cls, mask: "ndarray", one: "ndarray", two: "ndarray") -> "ndarray":
        """array.py:4455-4470."""
        args = (mask, one, two)
        mask = mask._maybe_convert(np.dtype(np.bool_), args)
        common_type = cls.find_common_type(one, two)
        one = one._maybe_convert(common_type, args)
        two = two._maybe_convert(common_type, args)
        out_shape = np.broadcast_shapes(mask.shape, one.shape, two.shape)
        out = ndarray(shape=out_shape, dtype=common_type, inputs=args)
        out._thunk.where(mask._thunk, one._thunk, two._thunk)
        return out
