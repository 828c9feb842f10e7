"""ufunc objects: dtype resolution, casting, `out=`, `.reduce` — the host-side dispatch of the hot
path.  Mirrors the behaviour of the reference's cunumeric/_ufunc/ufunc.py (unary_ufunc :349-437,
multiout_unary_ufunc :440-528, binary_ufunc :531-781): an ORDERED table of type signatures per
ufunc; resolution takes the exact match, else the first signature every operand can be safely cast
to, inserting CONVERT tasks.  Python scalars are weak (NEP 50), which is what the installed NumPy
does and what the reference's NumPy-1.x value-based rule gives for the cases its tests pin."""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple, Union

import math

import numpy as np

from ..config import BinaryOpCode, UnaryOpCode, UnaryRedCode
from ..deferred import DeferredArray
from ..runtime import runtime as _runtime
from ..distributed import _partitioning
from ..store import Store


def _single_gpu() -> bool:
    """New arrays are plain per-process device arrays: one GPU, or a `replicated()` block."""
    return _runtime.world_size == 1 or not _partitioning[0]


_ndarray_cls: list = []


def _ndarray_type():
    if not _ndarray_cls:
        from ..array import ndarray

        _ndarray_cls.append(ndarray)
    return _ndarray_cls[0]

float_dtypes = ["e", "f", "d"]
complex_dtypes = ["F", "D"]
float_and_complex = float_dtypes + complex_dtypes
integer_dtypes = ["b", "B", "h", "H", "i", "I", "l", "L", "q", "Q"]
all_but_boolean = integer_dtypes + float_and_complex
all_dtypes = ["?"] + all_but_boolean


def predicate_types_of(dtypes: Sequence[str]) -> list:
    return [ty + "?" for ty in dtypes]


def relation_types_of(dtypes: Sequence[str]) -> list:
    return [ty * 2 + "?" for ty in dtypes]


def _char(dtype: np.dtype) -> str:
    # 'q'/'Q' and 'l'/'L' are the same 64-bit types on this platform
    return dtype.char


class ufunc:
    _types: Dict[Any, str]
    _nin: int
    _nout: int

    def __init__(self, name: str, doc: str) -> None:
        self._name = name
        self.__doc__ = doc

    @property
    def __name__(self) -> str:
        return self._name

    @property
    def nin(self) -> int:
        return self._nin

    @property
    def nout(self) -> int:
        return self._nout

    @property
    def types(self) -> list:
        return [f"{''.join(i)}->{''.join(o)}" for i, o in self._types.items()]

    @property
    def ntypes(self) -> int:
        return len(self._types)

    def __repr__(self) -> str:
        return f"<ufunc {self._name}>"

    # ------------------------------------------------------------------ shared plumbing
    def _maybe_cast_input(self, arr, to_dtype, casting):
        to_dtype = np.dtype(to_dtype)
        if arr.dtype == to_dtype:
            return arr
        if not np.can_cast(arr.dtype, to_dtype, casting=casting):
            raise TypeError(f"Cannot cast ufunc '{self._name}' input from {arr.dtype} to "
                            f"{to_dtype} with casting rule '{casting}'")
        return arr._astype(to_dtype, temporary=True)

    def _maybe_create_result(self, out, out_shape, res_dtype, casting, inputs):
        from ..array import ndarray

        if out is None:
            return ndarray(shape=out_shape, dtype=res_dtype, inputs=inputs)
        if out.dtype != res_dtype:
            if not np.can_cast(res_dtype, out.dtype, casting=casting):
                raise TypeError(f"Cannot cast ufunc '{self._name}' output from {res_dtype} to "
                                f"{out.dtype} with casting rule '{casting}'")
            return ndarray(shape=out.shape, dtype=res_dtype, inputs=inputs)
        return out

    @staticmethod
    def _maybe_cast_output(out, result):
        if out is None or out is result:
            return result
        out._thunk.convert(result._thunk, warn=False)
        return out

    def _prepare_operands(self, *args: Any, out, where: Any = True):
        from ..array import convert_to_cunumeric_ndarray, ndarray

        max_nargs = self.nin + self.nout
        if len(args) < self.nin or len(args) > max_nargs:
            raise TypeError(f"{self._name}() takes from {self.nin} to {max_nargs} positional "
                            f"arguments but {len(args)} were given")
        inputs = tuple(convert_to_cunumeric_ndarray(arr, share=True) for arr in args[: self.nin])
        if len(args) > self.nin:
            if out is not None:
                raise TypeError("cannot specify 'out' as both a positional and keyword argument")
            computed_out = tuple(args[self.nin:])
            computed_out += (None,) * (self.nout - len(computed_out))
        elif out is None:
            computed_out = (None,) * self.nout
        elif not isinstance(out, tuple):
            computed_out = (out,)
        else:
            computed_out = out
        outputs = []
        self._numpy_outs = []
        for o in computed_out:
            if o is None or isinstance(o, ndarray):
                outputs.append(o)
            elif isinstance(o, np.ndarray):
                # NumPy arrays are accepted as `out` and written back (ufunc.py:262-272)
                dev = convert_to_cunumeric_ndarray(o)
                dev._writeback = o
                outputs.append(dev)
            else:
                raise TypeError("return arrays must be of ArrayType")
        if self.nout != len(outputs):
            raise ValueError("The 'out' tuple must have exactly one entry per ufunc output")
        shapes = [arr.shape for arr in inputs]
        shapes.extend(arr.shape for arr in outputs if arr is not None)
        out_shape = np.broadcast_shapes(*shapes)
        for o in outputs:
            if o is not None and o.shape != out_shape:
                raise ValueError(f"non-broadcastable output operand with shape {o.shape} doesn't "
                                 f"match the broadcast shape {out_shape}")
        if not isinstance(where, bool) or not where:
            raise NotImplementedError("the 'where' keyword is not yet supported")
        return inputs, tuple(outputs), out_shape, where

    @staticmethod
    def _finish(out):
        wb = getattr(out, "_writeback", None)
        if wb is not None:
            wb[...] = out.__array__()
            return wb
        return out


class unary_ufunc(ufunc):
    def __init__(self, name: str, doc: str, op_code: UnaryOpCode, types: Dict[str, str],
                 overrides: Dict[str, UnaryOpCode]) -> None:
        super().__init__(name, doc)
        self._types = types
        in_ty, out_ty = next(iter(types.items()))
        self._nin, self._nout = len(in_ty), len(out_ty)
        self._op_code = op_code
        self._resolution_cache: Dict[np.dtype, np.dtype] = {}
        self._fast_types: Dict[str, np.dtype] = {}
        self._overrides = overrides

    def _resolve_dtype(self, arr, precision_fixed: bool):
        c = _char(arr.dtype)
        if c in self._types:
            return arr, np.dtype(self._types[c])
        if not precision_fixed and arr.dtype in self._resolution_cache:
            to_dtype = self._resolution_cache[arr.dtype]
            return arr._astype(to_dtype, temporary=True), np.dtype(self._types[to_dtype.char])
        chosen = None
        if not precision_fixed:
            for in_ty in self._types.keys():
                if np.can_cast(arr.dtype, in_ty):
                    chosen = in_ty
                    break
        if chosen is None:
            raise TypeError(f"No matching signature of ufunc {self._name} is found for the "
                            "given casting")
        to_dtype = np.dtype(chosen)
        self._resolution_cache[arr.dtype] = to_dtype
        return arr._astype(to_dtype, temporary=True), np.dtype(self._types[to_dtype.char])

    def _fast_call(self, x):
        """`ufunc(array)` whose dtype is in the type table: same result as the general path."""
        ndarray = _ndarray_type()
        if type(x) is not ndarray or x.ndim == 0:
            return NotImplemented
        t = x._thunk
        if type(t) is not DeferredArray or not _single_gpu():
            return NotImplemented
        c = x.dtype.char
        res = self._fast_types.get(c)
        if res is None:
            if c not in self._types:
                return NotImplemented
            res = self._fast_types[c] = np.dtype(self._types[c])
        out = DeferredArray(Store.empty(x.shape, res))
        out.unary_op_prepared(self._overrides.get(c, self._op_code), t.base)
        result = object.__new__(ndarray)
        result._thunk, result._writeback = out, None
        return result

    def __call__(self, *args: Any, out=None, where: Any = True, casting: str = "same_kind",
                 order: str = "K", dtype=None, **kwargs: Any):
        if out is None and where is True and dtype is None and len(args) == 1 and not kwargs:
            fast = self._fast_call(args[0])
            if fast is not NotImplemented:
                return fast
        (x,), (out,), out_shape, where = self._prepare_operands(*args, out=out, where=where)
        precision_fixed = False
        if dtype is not None:
            precision_fixed = True
            x = self._maybe_cast_input(x, dtype, casting)
        x, res_dtype = self._resolve_dtype(x, precision_fixed)
        result = self._maybe_create_result(out, out_shape, res_dtype, casting, (x,))
        op_code = self._overrides.get(x.dtype.char, self._op_code)
        result._thunk.unary_op(op_code, x._thunk, where, ())
        return self._finish(self._maybe_cast_output(out, result))


class multiout_unary_ufunc(ufunc):
    def __init__(self, name: str, doc: str, op_code: UnaryOpCode, types: Dict[Any, Any]) -> None:
        super().__init__(name, doc)
        self._types = types
        in_ty, out_ty = next(iter(types.items()))
        self._nin, self._nout = len(in_ty), len(out_ty)
        self._op_code = op_code
        self._resolution_cache: Dict[np.dtype, np.dtype] = {}

    def _resolve_dtype(self, arr, precision_fixed: bool):
        c = _char(arr.dtype)
        if c in self._types:
            return arr, tuple(np.dtype(t) for t in self._types[c])
        chosen = None
        if not precision_fixed:
            for in_ty in self._types.keys():
                if np.can_cast(arr.dtype, in_ty):
                    chosen = in_ty
                    break
        if chosen is None:
            raise TypeError(f"No matching signature of ufunc {self._name} is found for the "
                            "given casting")
        to_dtype = np.dtype(chosen)
        return (arr._astype(to_dtype, temporary=True),
                tuple(np.dtype(t) for t in self._types[to_dtype.char]))

    def __call__(self, *args: Any, out=None, where: Any = True, casting: str = "same_kind",
                 order: str = "K", dtype=None, **kwargs: Any):
        (x,), outs, out_shape, where = self._prepare_operands(*args, out=out, where=where)
        precision_fixed = False
        if dtype is not None:
            precision_fixed = True
            x = self._maybe_cast_input(x, dtype, casting)
        x, res_dtypes = self._resolve_dtype(x, precision_fixed)
        results = tuple(self._maybe_create_result(o, out_shape, rd, casting, (x,))
                        for o, rd in zip(outs, res_dtypes))
        thunks = tuple(r._thunk for r in results)
        thunks[0].unary_op(self._op_code, x._thunk, where, (), multiout=thunks[1:])
        return tuple(self._finish(self._maybe_cast_output(o, r)) for o, r in zip(outs, results))


class binary_ufunc(ufunc):
    def __init__(self, name: str, doc: str, op_code: BinaryOpCode,
                 types: Dict[Tuple[str, str], str], red_code: Optional[UnaryRedCode] = None,
                 use_common_type: bool = True) -> None:
        super().__init__(name, doc)
        self._types = types
        in_ty, out_ty = next(iter(types.items()))
        self._nin, self._nout = len(in_ty), len(out_ty)
        self._op_code = op_code
        self._resolution_cache: Dict[Tuple[str, ...], Tuple[str, ...]] = {}
        self._fast_types = {}
        self._red_code = red_code
        self._use_common_type = use_common_type

    @staticmethod
    def _find_common_type(arrs, orig_args) -> np.dtype:
        from ..array import ndarray

        all_ndarray = all(isinstance(a, ndarray) for a in orig_args)
        if all_ndarray and len({a.dtype for a in arrs}) == 1:
            return arrs[0].dtype
        operands = []
        for arr, orig in zip(arrs, orig_args):
            if isinstance(orig, (bool, int, float, complex)) and not isinstance(orig, np.generic):
                operands.append(orig)  # weak Python scalar
            else:
                operands.append(arr.dtype)
        return np.result_type(*operands)

    def _resolve_dtype(self, arrs, orig_args, casting, precision_fixed: bool):
        if self._use_common_type:
            common = self._find_common_type(arrs, orig_args)
            to_dtypes = (common, common)
            key = (common.char, common.char)
        else:
            to_dtypes = tuple(a.dtype for a in arrs)
            key = tuple(a.dtype.char for a in arrs)
        if key in self._types:
            arrs = [a._astype(t, temporary=True) for a, t in zip(arrs, to_dtypes)]
            return arrs, np.dtype(self._types[key])
        if not precision_fixed and key in self._resolution_cache:
            chosen = self._resolution_cache[key]
            arrs = [a._astype(np.dtype(t), temporary=True) for a, t in zip(arrs, chosen)]
            return arrs, np.dtype(self._types[chosen])
        chosen = None
        if not precision_fixed:
            for in_dtypes in self._types.keys():
                if all(np.can_cast(t, to) for t, to in zip(to_dtypes, in_dtypes)):
                    chosen = in_dtypes
                    break
            if chosen is None and not self._use_common_type:
                for in_dtypes in self._types.keys():
                    if np.can_cast(arrs[0].dtype, in_dtypes[0]) and all(
                            np.can_cast(a.dtype, to, casting=casting)
                            for a, to in zip(arrs[1:], in_dtypes[1:])):
                        chosen = in_dtypes
                        break
        if chosen is None:
            raise TypeError(f"No matching signature of ufunc {self._name} is found for the "
                            "given casting")
        self._resolution_cache[key] = chosen
        arrs = [a._astype(np.dtype(t), temporary=True) for a, t in zip(arrs, chosen)]
        return arrs, np.dtype(self._types[chosen])

    # ---- fast path -----------------------------------------------------------------------------
    # `array (op) array` with equal dtype and shape, and `floating array (op) Python float/int`
    # (a weak scalar adopts the array's dtype) resolve to the same result as the general path
    # below, without its per-call type arithmetic.  With fused chains the GPU work of a whole
    # Black-Scholes step is ~0.6 ms, so the host cost per NumPy call is what bounds throughput.
    _scalar_cache: Dict[Tuple[str, type, Any], Any] = {}
    _fast_types: Dict[str, np.dtype]

    @classmethod
    def _weak_scalar(cls, value, dtype):
        # 0.0 == -0.0 (and they hash alike) but `x / -0.0`, `copysign(x, -0.0)`, `x * -0.0` differ:
        # the sign of a zero is part of the key
        key = (dtype.char, type(value), value,
               math.copysign(1.0, value) if value == 0 else 0.0)
        hit = cls._scalar_cache.get(key)
        if hit is None:
            ndarray = _ndarray_type()
            if len(cls._scalar_cache) > 512:
                cls._scalar_cache.clear()
            with np.errstate(all="ignore"):
                host = np.asarray(value).astype(dtype)
            hit = ndarray(shape=(), dtype=dtype,
                          thunk=DeferredArray(Store.from_scalar(host), host_scalar=host))
            cls._scalar_cache[key] = hit
        return hit

    _bcast_cache: Dict[Tuple[int, Tuple[int, ...]], Any] = {}

    @classmethod
    def _broadcast_scalar(cls, base, shape):
        """stride-0 view of a cached weak scalar, itself cached (the scalar ndarray stays alive in
        _scalar_cache, so id(base) is stable; both caches are cleared together)"""
        key = (id(base), shape)
        hit = cls._bcast_cache.get(key)
        if hit is None or hit[0] is not base:
            if len(cls._bcast_cache) > 2048:
                cls._bcast_cache.clear()
            hit = cls._bcast_cache[key] = (base, base.broadcast_to(shape))
        return hit[1]

    def _fast_call(self, a, b):
        ndarray = _ndarray_type()
        ta, tb = type(a), type(b)
        if ta is ndarray and tb is ndarray:
            dt = a.dtype
            if dt != b.dtype or a.shape != b.shape or a.ndim == 0:
                return NotImplemented
            x1, x2, shape = a, b, a.shape
        elif ta is ndarray and (tb is float or tb is int):
            dt = a.dtype
            if dt.kind != "f" or a.ndim == 0 or b != b:
                return NotImplemented
            x1, x2, shape = a, self._weak_scalar(b, dt), a.shape
        elif tb is ndarray and (ta is float or ta is int):
            dt = b.dtype
            if dt.kind != "f" or b.ndim == 0 or a != a:
                return NotImplemented
            x1, x2, shape = self._weak_scalar(a, dt), b, b.shape
        else:
            return NotImplemented
        res = self._fast_types.get(dt.char)
        if res is None:
            res = self._types.get((dt.char, dt.char))
            if res is None:
                return NotImplemented
            res = self._fast_types[dt.char] = np.dtype(res)
        t1, t2 = x1._thunk, x2._thunk
        if type(t1) is DeferredArray and type(t2) is DeferredArray and _single_gpu():
            # brand-new output: no aliasing to resolve, operands are same-shape or 0-d
            out = DeferredArray(Store.empty(shape, res))
            b1, b2 = t1.base, t2.base
            if b1.shape != shape:
                b1 = self._broadcast_scalar(b1, shape)
            if b2.shape != shape:
                b2 = self._broadcast_scalar(b2, shape)
            out.binary_op_prepared(self._op_code, b1, b2)
            result = object.__new__(ndarray)
            result._thunk, result._writeback = out, None
            return result
        result = ndarray(shape, res, inputs=(x1, x2))
        result._thunk.binary_op(self._op_code, t1, t2, True, ())
        return result

    def __call__(self, *args: Any, out=None, where: Any = True, casting: str = "same_kind",
                 order: str = "K", dtype=None, **kwargs: Any):
        if out is None and where is True and dtype is None and len(args) == 2 and \
                self._use_common_type and not kwargs:
            fast = self._fast_call(args[0], args[1])
            if fast is not NotImplemented:
                return fast
        arrs, (out,), out_shape, where = self._prepare_operands(*args, out=out, where=where)
        orig_args = args[: self.nin]
        precision_fixed = False
        if dtype is not None:
            precision_fixed = True
            arrs = [self._maybe_cast_input(a, dtype, casting) for a in arrs]
            orig_args = arrs
        arrs, res_dtype = self._resolve_dtype(arrs, orig_args, casting, precision_fixed)
        x1, x2 = arrs
        result = self._maybe_create_result(out, out_shape, res_dtype, casting, (x1, x2))
        result._thunk.binary_op(self._op_code, x1._thunk, x2._thunk, where, ())
        return self._finish(self._maybe_cast_output(out, result))

    def reduce(self, array, axis: Union[int, Tuple[int, ...], None] = 0, dtype=None, out=None,
               keepdims: bool = False, initial=None, where=None):
        """ufunc.reduce (ufunc.py:688-781): add->SUM, multiply->PROD, maximum->MAX, minimum->MIN,
        logical_and->ALL, logical_or->ANY."""
        from ..array import convert_to_cunumeric_ndarray

        array = convert_to_cunumeric_ndarray(array)
        if self._red_code is None:
            raise NotImplementedError(f"reduction for {self} is not yet implemented")
        if self._op_code in (BinaryOpCode.LOGICAL_AND, BinaryOpCode.LOGICAL_OR):
            res_dtype = np.dtype(np.bool_)
            if dtype is not None:
                raise TypeError("Cannot set dtype on a logical reduction")
        else:
            res_dtype = None
        if array.ndim == 0 and axis == 0:
            axis = None
        return array._perform_unary_reduction(self._red_code, array, axis=axis, dtype=dtype,
                                              out=out, keepdims=keepdims, initial=initial,
                                              where=where, res_dtype=res_dtype)


def _parse_unary_ufunc_type(ty: str) -> Tuple[str, str]:
    return (ty, ty) if len(ty) == 1 else (ty[0], ty[1:])


def create_unary_ufunc(summary: str, name: str, op_code: UnaryOpCode, types: Sequence[str],
                       overrides: Optional[Dict[str, UnaryOpCode]] = None) -> unary_ufunc:
    types_dict = dict(_parse_unary_ufunc_type(ty) for ty in types)
    return unary_ufunc(name, summary, op_code, types_dict, overrides or {})


def create_multiout_unary_ufunc(summary: str, name: str, op_code: UnaryOpCode,
                                types: Sequence[str]) -> multiout_unary_ufunc:
    types_dict = dict(_parse_unary_ufunc_type(ty) for ty in types)
    return multiout_unary_ufunc(name, summary, op_code, types_dict)


def _parse_binary_ufunc_type(ty: str):
    if len(ty) == 1:
        return ((ty, ty), ty)
    if len(ty) == 3:
        return ((ty[0], ty[1]), ty[2])
    raise NotImplementedError("Binary ufunc must have two inputs and one output")


def create_binary_ufunc(summary: str, name: str, op_code: BinaryOpCode, types: Sequence[str],
                        red_code: Optional[UnaryRedCode] = None,
                        use_common_type: bool = True) -> binary_ufunc:
    types_dict = dict(_parse_binary_ufunc_type(ty) for ty in types)
    return binary_ufunc(name, summary, op_code, types_dict, red_code=red_code,
                        use_common_type=use_common_type)
