"""Store: a typed, strided window onto a device allocation, with the zero-copy view algebra the
reference gets from legate.core stores: slice, project, promote, transpose, overlaps.
A Store never owns arithmetic; DeferredArray turns stores into cnb_store_t descriptors."""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from .config import MAX_DIM, dtype_code
from . import fusion
from .runtime import DeviceBuffer, runtime


def c_strides(shape: Sequence[int], itemsize: int) -> Tuple[int, ...]:
    strides = []
    acc = itemsize
    for n in reversed(shape):
        strides.append(acc)
        acc *= max(int(n), 1)
    return tuple(reversed(strides))


_STRIDES: dict = {}   # (shape, itemsize) -> C-order byte strides


class Store:
    __slots__ = ("buffer", "dtype", "shape", "strides", "offset", "_win", "_size")

    def __init__(self, buffer: DeviceBuffer, dtype, shape, strides=None, offset: int = 0) -> None:
        # `buffer` is bound LAST, together with the user count: a constructor that raises half-way
        # must not leave a Store whose __del__ un-counts a user it never counted (fusion.py treats
        # an output with zero users as unobservable)
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.strides = (c_strides(self.shape, self.dtype.itemsize) if strides is None
                        else tuple(int(s) for s in strides))
        self.offset = int(offset)
        self._win = None    # fusion._Window of this (immutable) store, built on first capture
        self._size = None
        self.buffer = buffer
        if buffer is not None:  # (shape-only probes carry no buffer)
            buffer.users += 1

    def __del__(self) -> None:
        try:
            buffer = self.buffer  # AttributeError if __init__ never got that far
        except AttributeError:
            return
        if buffer is not None:
            buffer.users -= 1

    # A Store counts as a user of its buffer (fusion.py drops stores into buffers nobody observes):
    # every duplicate has to go through the constructor, and a Store is never pickled (ndarray
    # pickles through the host array).
    def __copy__(self) -> "Store":
        return Store(self.buffer, self.dtype, self.shape, self.strides, self.offset)

    def __deepcopy__(self, memo=None) -> "Store":
        raise TypeError("a Store is a window onto device memory: copy the ndarray instead")

    def __reduce__(self):
        raise TypeError("a Store is a window onto device memory: pickle the ndarray instead")

    # ------------------------------------------------------------------ construction
    @staticmethod
    def empty(shape, dtype) -> "Store":
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in shape)
        size = math.prod(shape)
        # fast constructor: everything is already normalised
        st = object.__new__(Store)
        st.buffer = buffer = runtime.allocate(size * dtype.itemsize)
        st.dtype, st.shape, st.offset = dtype, shape, 0
        key = (shape, dtype.itemsize)
        strides = _STRIDES.get(key)
        if strides is None:
            if len(_STRIDES) > 4096:
                _STRIDES.clear()
            strides = _STRIDES[key] = c_strides(shape, dtype.itemsize)
        st.strides = strides
        st._win, st._size = None, size
        buffer.users += 1
        return st

    @staticmethod
    def from_scalar(value: np.ndarray) -> "Store":
        value = np.asarray(value)
        assert value.ndim == 0
        return Store(runtime.scalar_buffer(value), value.dtype, ())

    # ------------------------------------------------------------------ properties
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        if self._size is None:
            self._size = math.prod(self.shape)
        return self._size

    @property
    def ptr(self) -> int:
        # every consumer of device memory goes through here: pending fused chains run first
        if fusion.pending():
            fusion.flush()
        return self.buffer.ptr + self.offset

    @property
    def is_c_contiguous(self) -> bool:
        expect = self.dtype.itemsize
        for n, s in zip(reversed(self.shape), reversed(self.strides)):
            if n == 1:
                continue
            if s != expect:
                return False
            expect *= n
        return True

    # ------------------------------------------------------------------ view algebra
    def slice(self, dim: int, sl: slice) -> "Store":
        start, stop, step = sl.indices(self.shape[dim])
        n = len(range(start, stop, step))
        shape = self.shape[:dim] + (n,) + self.shape[dim + 1:]
        strides = self.strides[:dim] + (self.strides[dim] * step,) + self.strides[dim + 1:]
        offset = self.offset + (start * self.strides[dim] if n > 0 else 0)
        return Store(self.buffer, self.dtype, shape, strides, offset)

    def project(self, dim: int, index: int) -> "Store":
        """Drop `dim` at `index`."""
        if self.shape[dim] > 0 and not (0 <= index < self.shape[dim]):
            raise IndexError(f"index {index} is out of bounds for axis {dim}")
        shape = self.shape[:dim] + self.shape[dim + 1:]
        strides = self.strides[:dim] + self.strides[dim + 1:]
        return Store(self.buffer, self.dtype, shape, strides,
                     self.offset + index * self.strides[dim])

    def promote(self, dim: int, size: int = 1) -> "Store":
        """Insert a stride-0 (broadcast) dimension of `size` at `dim`."""
        shape = self.shape[:dim] + (int(size),) + self.shape[dim:]
        strides = self.strides[:dim] + (0,) + self.strides[dim:]
        return Store(self.buffer, self.dtype, shape, strides, self.offset)

    def transpose(self, axes: Sequence[int]) -> "Store":
        return Store(self.buffer, self.dtype, tuple(self.shape[a] for a in axes),
                     tuple(self.strides[a] for a in axes), self.offset)

    def broadcast_to(self, shape: Sequence[int]) -> "Store":
        """deferred.py:926-941 `_broadcast`: right-align, promote missing dims, stride-0 unit dims."""
        shape = tuple(int(s) for s in shape)
        if shape == self.shape:
            return self
        nd = len(shape)
        lead = nd - self.ndim
        if lead < 0:
            raise ValueError(f"cannot broadcast {self.shape} to {shape}")
        new_strides = []
        for d in range(nd):
            if d < lead:
                new_strides.append(0)
                continue
            n, s = self.shape[d - lead], self.strides[d - lead]
            if n == shape[d]:
                new_strides.append(s)
            elif n == 1:
                new_strides.append(0)
            else:
                raise ValueError(f"cannot broadcast {self.shape} to {shape}")
        return Store(self.buffer, self.dtype, shape, new_strides, self.offset)

    def reshape_contiguous(self, shape: Sequence[int]) -> "Store":
        assert self.is_c_contiguous
        return Store(self.buffer, self.dtype, shape, None, self.offset)

    def view_dtype(self, dtype, shape, strides) -> "Store":
        return Store(self.buffer, dtype, shape, strides, self.offset)

    def byte_bounds(self) -> Tuple[int, int]:
        lo = hi = self.offset
        if self.size == 0:
            return lo, lo
        for n, s in zip(self.shape, self.strides):
            if s >= 0:
                hi += (n - 1) * s
            else:
                lo += (n - 1) * s
        return lo, hi + self.dtype.itemsize

    def overlaps(self, other: "Store") -> bool:
        """Conservative alias test (deferred.py:291-303 uses store.overlaps)."""
        if self.buffer is not other.buffer:
            return False
        a0, a1 = self.byte_bounds()
        b0, b1 = other.byte_bounds()
        return a0 < b1 and b0 < a1

    def same_window(self, other: "Store") -> bool:
        return (self.buffer is other.buffer and self.offset == other.offset and
                self.shape == other.shape and self.strides == other.strides and
                self.dtype == other.dtype)

    # ------------------------------------------------------------------ C ABI descriptor
    def descriptor(self) -> _lib.cnb_store_t:
        if self.ndim > MAX_DIM:
            raise NotImplementedError(f"cunumeric_b200 supports at most {MAX_DIM} dimensions")
        if self.buffer.ready_event is not None:
            runtime.wait_ready(self.buffer)  # an asynchronous H2D copy is still filling it
        d = _lib.cnb_store_t()
        d.ptr = self.ptr
        d.dtype = dtype_code(self.dtype)
        d.ndim = self.ndim
        for i in range(self.ndim):
            d.shape[i] = self.shape[i]
            d.strides[i] = self.strides[i]
        return d
