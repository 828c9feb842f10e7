"""Process-level runtime: device selection, the compute stream, device allocations, host<->device
copies and (multi-GPU) the exchange communicator.

Plays the role the reference delegates to legate.core (cunumeric/runtime.py:545 `runtime`
singleton + Legion's allocator / StreamPool): one process drives one GPU; every task is launched
asynchronously on `runtime.stream`; nothing synchronises until a value is read on the host."""
from __future__ import annotations

import ctypes
import os
from typing import Any, Optional

import numpy as np

from . import _lib
from .config import argval_dtype, dtype_code


class DeviceBuffer:
    """An owned device allocation.  Blocks come from a size-keyed free list kept by the runtime
    (every task runs on the single compute stream, so a block released by Python's reference
    counting can be handed out again immediately: stream order already serialises the old reader
    and the new writer); misses go to the stream-ordered CUDA pool."""

    __slots__ = ("_ptr", "base", "nbytes", "_runtime", "ready_event", "busy_event", "users",
                 "readers", "shared", "__weakref__")

    def __init__(self, runtime: "Runtime", nbytes: int) -> None:
        self._runtime = runtime
        self.nbytes = max(int(nbytes), 1)
        # The block is taken on first use: a temporary that only ever lives inside a fused chain
        # (fusion.py) never touches memory and never allocates.
        self.base = self._ptr = None
        # set while an asynchronous H2D copy (on the copy stream) is still filling this buffer; the
        # compute stream waits for it the first time the buffer is used by a task
        self.ready_event = None
        # number of live Store windows onto this buffer (fusion.py: an output nobody can observe
        # any more is not written)
        self.users = 0
        # number of captured-but-not-yet-launched chains that read this buffer (fusion.py)
        self.readers = 0
        # a cached constant shared between arrays (runtime.scalar_buffer): never renamed
        self.shared = False
        # completion event of an asynchronous D2H copy that may still be reading the block
        self.busy_event = None

    @property
    def ptr(self) -> int:
        if self._ptr is None:
            self._runtime.ensure_initialized()
            self.base, self._ptr = self._runtime._take_block(self.nbytes)
        return self._ptr

    def release(self) -> None:
        """Give the block back now (a later `.ptr` would take a fresh one).  Only for buffers nobody
        can observe any more."""
        rt = self._runtime
        if rt is not None and rt.lib is not None and self._ptr:
            # a copy stream may still be writing (H2D) or reading (D2H) the block: its next owner
            # only ever sees the compute stream, so order that stream behind the copy first
            for ev in (self.ready_event, self.busy_event):
                if ev is not None:
                    rt.lib.cnb_stream_wait_event(rt.stream, ev)
            self.ready_event = self.busy_event = None
            rt._give_block(self.base, self._ptr, self.nbytes)
        self.base = self._ptr = None

    def __del__(self) -> None:
        try:
            self.release()
        except Exception:
            pass
        self._ptr = None


class PinnedBuffer:
    """Page-locked host memory exposed as a NumPy array (for asynchronous H2D / D2H)."""

    def __init__(self, runtime: "Runtime", nbytes: int) -> None:
        self._runtime = runtime
        self.nbytes = int(nbytes)
        self.ptr = _lib.check_ptr(runtime.lib.cnb_host_alloc(max(self.nbytes, 1)))

    def as_array(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        buf = (ctypes.c_byte * max(self.nbytes, 1)).from_address(self.ptr)
        buf._owner = self  # the ctypes buffer (hence every NumPy view of it) keeps the pages alive
        return np.frombuffer(buf, dtype=np.uint8, count=self.nbytes).view(dtype).reshape(shape)

    def __del__(self) -> None:
        try:
            if self.ptr and self._runtime.lib is not None:
                self._runtime.lib.cnb_host_free(self.ptr)
        except Exception:
            pass
        self.ptr = None


class Runtime:
    def __init__(self) -> None:
        self.lib: Any = None
        self.stream: Any = None
        self.device: int = -1
        self.rank: int = 0
        self.world_size: int = 1
        self.comm: Any = None
        self._scalar_cache: dict = {}
        self._argred_uids: dict = {}
        self._h2d_stream: Any = None
        self._d2h_stream: Any = None
        self._comm_stream: Any = None
        self._event_pool: list = []
        self._free_blocks: dict = {}
        self._cached_bytes = 0
        self._cache_limit = None
        self._colour = 0
        # dry run (fusion.trace_only): arrays are created and tasks captured, nothing touches a device
        self.dry_run = False

    # ------------------------------------------------------------------ lifecycle
    def ensure_initialized(self) -> None:
        if self.lib is not None:
            return
        lib = _lib.load()
        ndev = lib.cnb_device_count()
        if ndev <= 0:
            raise RuntimeError(
                "cunumeric_b200: no CUDA device visible. This library runs on B200 (sm_100a) "
                "only and has no CPU fallback.")
        device = int(os.environ.get("CUNUMERIC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        device %= ndev
        _lib.check(lib.cnb_init(device))
        self.lib = lib
        self.device = device
        # reference: cunumeric_perform_registration is the cffi registration callback
        lib.cunumeric_perform_registration()
        self.stream = _lib.check_ptr(lib.cnb_stream_create())

    @property
    def sm_count(self) -> int:
        self.ensure_initialized()
        return self.lib.cnb_sm_count()

    def launch_count(self) -> int:
        self.ensure_initialized()
        return int(self.lib.cnb_launch_count())

    def synchronize(self) -> None:
        from . import fusion

        fusion.flush()
        self.ensure_initialized()
        _lib.check(self.lib.cnb_stream_synchronize(self.stream))

    # ------------------------------------------------------------------ memory
    def allocate(self, nbytes: int) -> DeviceBuffer:
        if not self.dry_run:
            self.ensure_initialized()  # fail loudly without a device: there is no CPU fallback
        return DeviceBuffer(self, nbytes)  # the block itself is taken when `.ptr` is first read

    # Large blocks are "coloured": each gets a different sub-2-MiB start offset, so that the
    # operands of one task are never congruent modulo a large power of two.  Power-of-two sized
    # arrays laid out back to back otherwise walk the same L2 slices / HBM channels in lock step
    # (measured: WHERE on 2^30 fp16 elements dropped from 93 % to 58 % of the roofline).
    _COLOUR_MIN = 32 << 20
    _COLOUR_SPAN = 2 << 20

    def _take_block(self, nbytes: int):
        blocks = self._free_blocks.get(nbytes)
        if blocks:
            self._cached_bytes -= nbytes
            return blocks.pop()
        pad = self._COLOUR_SPAN if nbytes >= self._COLOUR_MIN else 0
        base = self.lib.cnb_malloc(nbytes + pad, self.stream)
        if not base and self._cached_bytes:
            self.release_cached_memory()  # out of memory: give the cached blocks back and retry
            base = self.lib.cnb_malloc(nbytes + pad, self.stream)
        base = _lib.check_ptr(base)
        offset = 0
        if pad:
            self._colour = (self._colour + 1) % 61
            offset = (self._colour * 33792) % pad & ~511  # 33 KiB steps, 512-byte aligned
        return base, base + offset

    def _give_block(self, base: int, ptr: int, nbytes: int) -> None:
        if self._cache_limit is None:
            free, total = ctypes.c_size_t(), ctypes.c_size_t()
            self.lib.cnb_mem_info(ctypes.byref(free), ctypes.byref(total))
            self._cache_limit = int(0.5 * total.value)
        if self._cached_bytes + nbytes > self._cache_limit:
            self.lib.cnb_free(base, self.stream)
            return
        self._free_blocks.setdefault(nbytes, []).append((base, ptr))
        self._cached_bytes += nbytes

    def adopt_block(self, buf: DeviceBuffer, new_base: int, new_ptr: int) -> None:
        """`buf` switches to a fresh block (fusion.py WAR renaming); the old one goes back to the
        free list.  Everything that will touch either block later is queued on the compute stream
        behind the kernel that filled the new one; an asynchronous D2H copy still reading the old
        block (copy stream) is waited for first."""
        if buf.busy_event is not None:
            _lib.check(self.lib.cnb_stream_wait_event(self.stream, buf.busy_event))
            buf.busy_event = None
        old_base, old_ptr = buf.base, buf._ptr
        buf.base, buf._ptr = new_base, new_ptr
        if old_ptr:
            self._give_block(old_base, old_ptr, buf.nbytes)

    def release_cached_memory(self) -> None:
        """Return every cached block to the CUDA pool."""
        for blocks in self._free_blocks.values():
            for base, _ in blocks:
                self.lib.cnb_free(base, self.stream)
        self._free_blocks.clear()
        self._cached_bytes = 0

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        self.ensure_initialized()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        return PinnedBuffer(self, n).as_array(shape, dtype)

    def copy_h2d(self, dst_ptr: int, src: np.ndarray) -> None:
        assert src.flags.c_contiguous
        if src.nbytes:
            _lib.check(self.lib.cnb_memcpy_h2d(dst_ptr, src.ctypes.data, src.nbytes, self.stream))

    def copy_d2h(self, dst: np.ndarray, src_ptr: int) -> None:
        assert dst.flags.c_contiguous
        if dst.nbytes:
            _lib.check(self.lib.cnb_memcpy_d2h(dst.ctypes.data, src_ptr, dst.nbytes, self.stream))

    # ------------------------------------------------------------------ asynchronous copies
    # Separate copy streams let the H2D of the next batch, the kernels of the current one and the
    # D2H of the previous one overlap (PCIe is full duplex).  Ordering is by events only.
    def _copy_streams(self):
        if self._h2d_stream is None:
            self._h2d_stream = _lib.check_ptr(self.lib.cnb_stream_create())
            self._d2h_stream = _lib.check_ptr(self.lib.cnb_stream_create())
        return self._h2d_stream, self._d2h_stream

    def comm_stream(self):
        """Stream of the halo exchanges that run next to the interior tiles of a fused chain
        (fusion._launch_tma); ordered with the compute stream by events only."""
        if self._comm_stream is None:
            self._comm_stream = _lib.check_ptr(self.lib.cnb_stream_create())
        return self._comm_stream

    def _event(self):
        return self._event_pool.pop() if self._event_pool else _lib.check_ptr(
            self.lib.cnb_event_create())

    def _recycle_event(self, ev) -> None:
        self._event_pool.append(ev)

    def copy_h2d_async(self, buf: DeviceBuffer, src: np.ndarray, offset: int = 0) -> None:
        """Fill `buf` (from byte `offset` on) from (pinned) host memory on the H2D stream."""
        assert src.flags.c_contiguous
        h2d, _ = self._copy_streams()
        # the pool handed out `buf` in compute-stream order: do not touch it before that point
        ev = self._event()
        _lib.check(self.lib.cnb_event_record(ev, self.stream))
        _lib.check(self.lib.cnb_stream_wait_event(h2d, ev))
        if src.nbytes:
            _lib.check(self.lib.cnb_memcpy_h2d(buf.ptr + offset, src.ctypes.data, src.nbytes, h2d))
        _lib.check(self.lib.cnb_event_record(ev, h2d))
        buf.ready_event = ev

    def wait_ready(self, buf: DeviceBuffer) -> None:
        ev = buf.ready_event
        if ev is not None:
            _lib.check(self.lib.cnb_stream_wait_event(self.stream, ev))
            buf.ready_event = None
            self._recycle_event(ev)

    def copy_d2h_async(self, dst: np.ndarray, src_ptr: int, buf: Optional[DeviceBuffer] = None):
        """Start a D2H copy on the D2H stream once the compute stream reaches this point; returns
        the event that marks its completion (`buf`: the allocation being read, told to wait for
        this event before it gives its block up)."""
        assert dst.flags.c_contiguous
        _, d2h = self._copy_streams()
        ev = self._event()
        _lib.check(self.lib.cnb_event_record(ev, self.stream))
        _lib.check(self.lib.cnb_stream_wait_event(d2h, ev))
        if dst.nbytes:
            _lib.check(self.lib.cnb_memcpy_d2h(dst.ctypes.data, src_ptr, dst.nbytes, d2h))
        _lib.check(self.lib.cnb_event_record(ev, d2h))
        if buf is not None:
            buf.busy_event = ev
        return ev

    def scalar_buffer(self, value: np.ndarray) -> DeviceBuffer:
        """Device copy of one scalar (the reference's Future-backed 0-d store); small LRU so that
        constants such as the `0.2` of the stencil are uploaded once."""
        key = (value.dtype.str, value.tobytes())
        buf = self._scalar_cache.get(key)
        if buf is None:
            if len(self._scalar_cache) > 1024:
                self._scalar_cache.clear()
            buf = self.allocate(value.nbytes)
            buf.shared = True
            staged = np.ascontiguousarray(value)
            if not self.dry_run:
                self.copy_h2d(buf.ptr, staged)
            # the H2D of a pageable source is staged by the driver before returning
            self._scalar_cache[key] = buf
        return buf

    # ------------------------------------------------------------------ arg-reduction struct types
    def get_argred_type(self, elem_dtype) -> np.dtype:
        """cunumeric/runtime.py:125-134: build {int64, T} and register its Argmax/Argmin redops."""
        self.ensure_initialized()
        elem_dtype = np.dtype(elem_dtype)
        dt = argval_dtype(elem_dtype)
        if elem_dtype not in self._argred_uids:
            uid = 1000 + dtype_code(elem_dtype)
            self.lib.cunumeric_register_reduction_op(uid, dtype_code(elem_dtype))
            self._argred_uids[elem_dtype] = uid
        return dt

    # ------------------------------------------------------------------ multi-GPU
    def init_distributed(self, rank: Optional[int] = None, world_size: Optional[int] = None,
                         exchange_id=None) -> None:
        """Create the NCCL clique (one process per GPU). `exchange_id(id_bytes|None) -> id_bytes`
        broadcasts rank 0's unique id; by default torch.distributed (any backend) is used as the
        bootstrap plumbing if it is already initialised."""
        self.ensure_initialized()
        if rank is None:
            rank = int(os.environ.get("RANK", "0"))
        if world_size is None:
            world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank, self.world_size = rank, world_size
        if world_size == 1 or self.comm is not None:
            return
        ident = ctypes.create_string_buffer(_lib.COMM_ID_BYTES)
        if rank == 0:
            _lib.check(self.lib.cnb_comm_unique_id(ident))
        if exchange_id is None:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("init_distributed needs torch.distributed initialised "
                                   "(or pass exchange_id=)")
            obj = [ident.raw if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            raw = obj[0]
        else:
            raw = exchange_id(ident.raw if rank == 0 else None)
        ident = ctypes.create_string_buffer(raw, _lib.COMM_ID_BYTES)
        self.comm = _lib.check_ptr(self.lib.cnb_comm_init(ident, world_size, rank))


runtime = Runtime()
