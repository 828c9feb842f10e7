"""PartitionedArray: the multi-GPU thunk (one process per GPU, SPMD).

In the reference every thunk is a DeferredArray and legate.core tiles its store over the GPUs
(SURVEY §2.2).  Here the tiling is explicit: a PartitionedArray owns a contiguous block of rows of
the global array (plus `halo` ghost rows on each side) in a local DeferredArray, and implements the
same thunk surface — unary_op / binary_op / where / convert / copy / fill / unary_reduction /
get_item / set_item — by running the single-GPU task on the local block.  The exchange steps Legion
performs implicitly are explicit, stream-ordered NCCL calls over NVLink:
  * shifted-slice operands (the stencil's north/south views): one grouped send/recv halo refresh per
    producer->consumer step (partition.plan_halo), re-done only after the base array was written;
  * scalar reductions: local partial -> ncclAllReduce (arg-reductions: ncclAllGather of the 16-byte
    {index, value} partials + one fold kernel with the lowest-index tie-break of the sequential fold);
  * axis-0 reductions (the partitioned axis): full-width local partial -> ncclAllReduce;
    reductions along other axes need no communication.
Everything the partition logic decides is a pure function of replicated metadata
(cunumeric_b200/partition.py), so all ranks issue the same collectives in the same order."""
from __future__ import annotations

import ctypes
import math
from typing import Any, Optional, Sequence, Tuple

import numpy as np

from . import _lib, fusion
from .config import BinaryOpCode, ConvertCode, UnaryRedCode, dtype_code
from .deferred import (_ARG_REDS, _UNARY_RED_IDENTITIES, DeferredArray, _basic_index,
                       launch_scalar_red)
from .partition import RowPartition, plan_fetch, plan_halo
from .runtime import runtime
from .store import Store

DEFAULT_HALO = 1

# how per-rank partials of each reduction are combined across ranks
_COMBINE = {
    UnaryRedCode.SUM: UnaryRedCode.SUM, UnaryRedCode.NANSUM: UnaryRedCode.SUM,
    UnaryRedCode.SUM_SQUARES: UnaryRedCode.SUM, UnaryRedCode.VARIANCE: UnaryRedCode.SUM,
    UnaryRedCode.COUNT_NONZERO: UnaryRedCode.SUM,
    UnaryRedCode.PROD: UnaryRedCode.PROD, UnaryRedCode.NANPROD: UnaryRedCode.PROD,
    UnaryRedCode.MAX: UnaryRedCode.MAX, UnaryRedCode.NANMAX: UnaryRedCode.MAX,
    UnaryRedCode.MIN: UnaryRedCode.MIN, UnaryRedCode.NANMIN: UnaryRedCode.MIN,
    UnaryRedCode.ALL: UnaryRedCode.ALL, UnaryRedCode.ANY: UnaryRedCode.ANY,
    UnaryRedCode.CONTAINS: UnaryRedCode.ANY,
}
# ... and the elementwise op that folds a combined partial into a pre-filled output
_FOLD_BINOP = {
    UnaryRedCode.SUM: BinaryOpCode.ADD, UnaryRedCode.PROD: BinaryOpCode.MULTIPLY,
    UnaryRedCode.MAX: BinaryOpCode.MAXIMUM, UnaryRedCode.MIN: BinaryOpCode.MINIMUM,
    UnaryRedCode.ALL: BinaryOpCode.LOGICAL_AND, UnaryRedCode.ANY: BinaryOpCode.LOGICAL_OR,
}


_ALIGN: dict = {}  # (owner tiling, halo, output tiling, first row of the view) -> 0 owned / 1 halo / 2 farther
_LEAD: dict = {}   # (inner shape, itemsize, halo) -> rows in front of the first owned row


def _lead_rows(inner: Tuple[int, ...], itemsize: int, halo: int) -> int:
    """Rows in front of the first owned row: the ghost rows, padded so that the OWNED block starts on
    a 128-byte (else 16-byte) boundary — with exactly `halo` rows in front, the local window of a
    1-D array would start one element into its buffer and every task on it would fall off the
    128-bit vector path (measured at 2 GPUs: add/bool at 0.17 of the roofline)."""
    key = (inner, itemsize, halo)
    lead = _LEAD.get(key)
    if lead is None:
        rb = math.prod(inner) * itemsize
        lead = halo
        if rb > 0:
            for unit in (128, 16):
                need = unit // math.gcd(rb, unit)          # rows per aligned step
                k = -(-halo // need) * need if halo else 0
                if k * rb <= max(4096, halo * rb * 8):
                    lead = int(k)
                    break
        if len(_LEAD) > 4096:
            _LEAD.clear()
        _LEAD[key] = lead
    return lead


class _Shared:
    """State shared by a base array and all of its views."""

    __slots__ = ("gshape", "dtype", "part", "halo", "lead", "local", "ghost_valid", "align_cache")

    def __init__(self, gshape, dtype, part: RowPartition, halo: int) -> None:
        self.gshape = gshape = tuple(int(s) for s in gshape)
        self.dtype = dtype = np.dtype(dtype)
        self.part = part
        self.halo = halo = int(halo)
        self.lead = lead = _lead_rows(gshape[1:], dtype.itemsize, halo)
        rows = part.count(runtime.rank) + lead + halo
        self.local = DeferredArray(Store.empty((rows,) + gshape[1:], dtype))
        self.ghost_valid = False
        self.align_cache = {}

    @property
    def row_bytes(self) -> int:
        return math.prod(self.gshape[1:]) * self.dtype.itemsize


def _comm_check(rc: int) -> None:
    _lib.check(rc)


class PartitionedArray:
    # a view is immutable: shape, induced partition, owned rows and the local windows are computed
    # once (the host cost per NumPy call bounds the iteration rate once the per-GPU work is < 1 ms)
    __slots__ = ("meta", "row0", "row1", "inner_key", "host_scalar", "_shape", "_part", "_owned",
                 "_local")

    def __init__(self, meta: _Shared, row0: int = 0, row1: Optional[int] = None,
                 inner_key: Tuple = ()) -> None:
        self.meta = meta
        self.row0 = row0
        self.row1 = meta.gshape[0] if row1 is None else row1
        self.inner_key = inner_key  # basic index applied to dims >= 1 of the local block
        self.host_scalar = None
        self._shape = self._part = self._owned = None
        self._local = {}

    # ------------------------------------------------------------------ construction
    @staticmethod
    def empty(shape, dtype, like: Optional["PartitionedArray"] = None,
              halo: int = DEFAULT_HALO) -> "PartitionedArray":
        shape = tuple(int(s) for s in shape)
        if like is not None and like.shape[0] == shape[0]:
            part = like.part  # aligned with the operand it will be computed from
        else:
            part = RowPartition.even(shape[0], runtime.world_size)
        return PartitionedArray(_Shared(shape, dtype, part, halo))

    @staticmethod
    def from_numpy(array: np.ndarray, halo: int = DEFAULT_HALO) -> "PartitionedArray":
        """Every rank holds the same host array (SPMD) and uploads only its own rows."""
        out = PartitionedArray.empty(array.shape, array.dtype, halo=halo)
        lo, hi = out.part.bounds(runtime.rank)
        if hi > lo:
            src = np.ascontiguousarray(array[lo:hi])
            runtime.copy_h2d(out.local_rows(lo, hi).base.ptr, src)
        return out

    @staticmethod
    def from_local_rows(block: np.ndarray, global_rows: int, halo: int = DEFAULT_HALO,
                        blocking: bool = True) -> "PartitionedArray":
        """SPMD upload: `block` holds THIS rank's rows of a (global_rows, ...) array split evenly
        over the ranks (RowPartition.even) — no rank ever materialises the whole array.
        blocking=False: the copy runs on the H2D stream (pinned source) and overlaps with kernels
        already queued; the first task that uses the array waits for it."""
        block = np.ascontiguousarray(block)
        out = PartitionedArray.empty((int(global_rows),) + block.shape[1:], block.dtype, halo=halo)
        lo, hi = out.part.bounds(runtime.rank)
        if block.shape[0] != hi - lo:
            raise ValueError(f"rank {runtime.rank} owns rows [{lo}, {hi}) of {global_rows}: expected "
                             f"a block of {hi - lo} rows, got {block.shape[0]}")
        if hi > lo:
            local = out.local_rows(lo, hi).base
            if blocking:
                runtime.copy_h2d(local.ptr, block)
            else:
                runtime.copy_h2d_async(local.buffer, block, local.offset)
        return out

    def local_block_to_host(self, out: Optional[np.ndarray] = None, blocking: bool = True):
        """This rank's rows of the view, copied to the host (no communication).  blocking=False
        returns a future (`.wait()`), the copy runs on the D2H stream."""
        vlo, vhi = self.owned
        if vhi > vlo:
            local = self.local_rows(vlo, vhi)
            if blocking:
                return local.__numpy_array__(out)
            if out is None:
                out = runtime.pinned_empty(local.shape, local.dtype)
            return local.to_host_async(out)
        out = np.empty((0,) + self.shape[1:], self.dtype) if out is None else out
        if blocking:
            return out
        from .deferred import HostFuture

        return HostFuture(None, None, out)

    # ------------------------------------------------------------------ properties
    @property
    def part(self) -> RowPartition:
        """Partition of THIS view's rows (view coordinates)."""
        p = self._part
        if p is None:
            m = self.meta
            p = self._part = m.part if (self.row0 == 0 and self.row1 == m.gshape[0]) else \
                m.part.window(self.row0, self.row1)
        return p

    @property
    def owned(self) -> Tuple[int, int]:
        o = self._owned
        if o is None:
            o = self._owned = self.part.bounds(runtime.rank)
        return o

    @property
    def shape(self) -> Tuple[int, ...]:
        sh = self._shape
        if sh is None:
            sh = (self.row1 - self.row0,) + self.meta.gshape[1:]
            if self.inner_key:
                sh = _basic_index(Store(None, self.meta.dtype, sh), (slice(None),) + self.inner_key).shape
            self._shape = sh
        return sh

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def dtype(self) -> np.dtype:
        return self.meta.dtype

    @property
    def scalar(self) -> bool:
        return False

    # ------------------------------------------------------------------ local access
    def local_rows(self, vlo: int, vhi: int) -> DeferredArray:
        """DeferredArray over view rows [vlo, vhi) (must lie inside owned +- halo)."""
        hit = self._local.get((vlo, vhi))
        if hit is not None:
            return hit
        m = self.meta
        lo, hi = m.part.bounds(runtime.rank)
        b0, b1 = self.row0 + vlo, self.row0 + vhi
        assert lo - m.halo <= b0 and b1 <= hi + m.halo, (b0, b1, lo, hi, m.halo)
        base = m.local.base.slice(0, slice(b0 - (lo - m.lead), b1 - (lo - m.lead)))
        if self.inner_key:
            base = _basic_index(base, (slice(None),) + self.inner_key)
        out = self._local[(vlo, vhi)] = DeferredArray(base)
        return out

    def _invalidate(self) -> None:
        self.meta.ghost_valid = False

    # ------------------------------------------------------------------ exchange
    def _run_transfers(self, transfers, dst_base_row0: int, dst: DeferredArray, stream=None) -> None:
        """Execute a transfer plan in one NCCL group: sends read this rank's block, receives land
        in `dst` (a row-contiguous local buffer whose row 0 is base row `dst_base_row0`)."""
        m = self.meta
        lib, comm = runtime.lib, runtime.comm
        if stream is None:
            stream = runtime.stream
        lo, _ = m.part.bounds(runtime.rank)
        rb = m.row_bytes
        mine = [t for t in transfers if t.src == runtime.rank or t.dst == runtime.rank]
        if not mine:
            return
        # Take the pointers BEFORE opening the group: Store.ptr first runs whatever is still deferred
        # (fused chains and queued halo exchanges, i.e. other NCCL groups), which must neither be
        # nested inside this group nor be launched after it.
        src_base, dst_base = m.local.base.ptr, dst.base.ptr
        _comm_check(lib.cnb_comm_group_start())
        for t in mine:
            if t.src == runtime.rank:
                ptr = src_base + (t.row_lo - (lo - m.lead)) * rb
                _comm_check(lib.cnb_comm_send(comm, ptr, t.nrows * rb, t.dst, stream))
            else:
                ptr = dst_base + (t.row_lo - dst_base_row0) * rb
                _comm_check(lib.cnb_comm_recv(comm, ptr, t.nrows * rb, t.src, stream))
        _comm_check(lib.cnb_comm_group_end())

    def exchange_halo(self) -> None:
        """Refresh the ghost rows of the base array from the neighbouring ranks."""
        m = self.meta
        if m.ghost_valid or runtime.world_size == 1 or m.halo == 0:
            m.ghost_valid = True
            return
        lo, _ = m.part.bounds(runtime.rank)
        plan = m.align_cache.get("halo_plan")
        if plan is None:
            plan = m.align_cache["halo_plan"] = [t for t in plan_halo(m.part, m.halo)
                                                 if t.src == runtime.rank or t.dst == runtime.rank]
        # The exchange keeps its place in program order between the deferred fused chains
        # (fusion.enqueue): it reads the rows the previous chain writes and fills the ghosts the
        # next chain reads, and takes its pointers when it runs — the base buffer may have been
        # renamed by then.  Flushing here instead would launch the previous iteration's chain
        # before its temporaries have died.
        # The chain in front of it may run AROUND the exchange (fusion.Overlap): the byte ranges tell
        # it which of its tile rows the exchange depends on.
        ranges = m.align_cache.get("halo_ranges")
        if ranges is None:
            rb, row0 = m.row_bytes, lo - m.lead
            ranges = m.align_cache["halo_ranges"] = (
                [((t.row_lo - row0) * rb, (t.row_hi - row0) * rb) for t in plan if t.src == runtime.rank],
                [((t.row_lo - row0) * rb, (t.row_hi - row0) * rb) for t in plan if t.dst == runtime.rank])
        overlap = fusion.Overlap(
            m.local.base.buffer, ranges[0], ranges[1],
            lambda stream: self._run_transfers(plan, lo - m.lead, m.local, stream))
        fusion.enqueue(lambda: self._run_transfers(plan, lo - m.lead, m.local), overlap)
        m.ghost_valid = True

    def _ensure_aligned_with(self, out_part: RowPartition) -> int:
        """This view is about to be read row for row by a task whose output rows are tiled by
        `out_part`: make the rows every rank needs available (collective) and return where they are
        — 0: owned, 1: owned + ghost rows (refreshed here), 2: farther away (the caller fetches them,
        `_fetch_rows`).  The classification is a
        pure function of the two tilings and is cached (temporaries of a loop body are new arrays with
        the same tiling every iteration)."""
        m = self.meta
        key = (m.part, m.halo, out_part, self.row0)
        kind = _ALIGN.get(key)
        if kind is None:
            kind = 0   # 0: owned, 1: within the halo, 2: farther
            for r in range(out_part.world):
                a, b = out_part.bounds(r)
                if b <= a:
                    continue
                nlo, nhi = self.row0 + a, self.row0 + b
                lo, hi = m.part.bounds(r)
                if nlo < lo - m.halo or nhi > hi + m.halo:
                    kind = 2
                    break
                if nlo < lo or nhi > hi:
                    kind = 1
            if len(_ALIGN) > 4096:
                _ALIGN.clear()
            _ALIGN[key] = kind
        if kind == 1:
            self.exchange_halo()
        return kind

    def _fetch_rows(self, out_part: RowPartition) -> DeferredArray:
        """Rows of this view that are farther than the halo depth from their owners, for a task whose
        output rows are tiled by `out_part`: every rank receives the rows it is about to read into a
        temporary block (one grouped send/recv, `plan_fetch`; rows it owns itself are copied locally)
        — the general redistribution Legion performs when an operand's tiling is not aligned with the
        output's.  Collective."""
        m = self.meta
        rank = runtime.rank
        needs = []
        for r in range(out_part.world):
            a, b = out_part.bounds(r)
            needs.append((self.row0 + a, self.row0 + b) if b > a else (0, 0))
        nlo, nhi = needs[rank]
        tmp = DeferredArray(Store.empty((max(0, nhi - nlo),) + m.gshape[1:], m.dtype))
        lo, hi = m.part.bounds(rank)
        a, b = max(nlo, lo), min(nhi, hi)
        if b > a:
            DeferredArray(tmp.base.slice(0, slice(a - nlo, b - nlo))).copy(
                PartitionedArray(m).local_rows(a, b), deep=True)
        self._run_transfers(plan_fetch(m.part, needs), nlo, tmp)
        base = tmp.base
        if self.inner_key:
            base = _basic_index(base, (slice(None),) + self.inner_key)
        return DeferredArray(base)

    def gather(self) -> DeferredArray:
        """Replicated copy of the whole view on every rank (collective)."""
        m = self.meta
        n = m.gshape[0]
        full = DeferredArray(Store.empty(m.gshape, m.dtype))
        lo, hi = m.part.bounds(runtime.rank)
        if hi > lo:
            DeferredArray(full.base.slice(0, slice(lo, hi))).copy(
                PartitionedArray(m).local_rows(lo, hi), deep=True)
        if runtime.world_size > 1:
            self._run_transfers(plan_fetch(m.part, [(0, n)] * runtime.world_size), 0, full)
        view = full.base.slice(0, slice(self.row0, self.row1))
        if self.inner_key:
            view = _basic_index(view, (slice(None),) + self.inner_key)
        return DeferredArray(view)

    def __numpy_array__(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        return self.gather().__numpy_array__(out)

    # ------------------------------------------------------------------ operand alignment
    def _operand(self, src: Any, vlo: int, vhi: int) -> DeferredArray:
        """Local piece of `src` aligned with this (output) view's rows [vlo, vhi)."""
        if type(src) is PartitionedArray:
            sshape, myshape = src.shape, self.shape
            if len(sshape) == len(myshape) and sshape[0] == myshape[0]:
                if src._ensure_aligned_with(self.part) == 2:
                    return src._fetch_rows(self.part)
                # (a rank that owns none of the output rows takes part in the collectives above and
                # then has nothing to read)
                return src.local_rows(vlo, vhi) if vhi > vlo else None
            # cannot be aligned row-for-row with the output (lower rank, or broadcast along the
            # partitioned axis): replicate it (collective) and fall through to the replicated rules
            src = src.gather()
        # replicated operand: scalars and lower-rank / unit-row arrays broadcast, full-height ones
        # are cut to the local rows
        if src.ndim == self.ndim and self.ndim > 0 and src.shape[0] == self.shape[0] \
                and self.shape[0] != 1:
            return DeferredArray(src.base.slice(0, slice(vlo, vhi)))
        return src

    def _local_task(self, fn, srcs: Sequence[Any]) -> None:
        vlo, vhi = self.owned
        local_srcs = [self._operand(s, vlo, vhi) for s in srcs]  # collective part first
        if vhi > vlo:
            fn(self.local_rows(vlo, vhi), *local_srcs)
        self._invalidate()

    # ------------------------------------------------------------------ thunk surface
    def unary_op(self, op, src, where=True, args=(), multiout=None) -> None:
        if multiout:
            vlo, vhi = self.owned
            outs = [o.local_rows(*o.owned) for o in multiout]
            self._local_task(lambda out, s: out.unary_op(op, s, where, args, multiout=outs), [src])
            for o in multiout:
                o._invalidate()
            return
        self._local_task(lambda out, s: out.unary_op(op, s, where, args), [src])

    def binary_op(self, op_code, src1, src2, where=True, args=()) -> None:
        self._local_task(lambda out, a, b: out.binary_op(op_code, a, b, where, args), [src1, src2])

    def isclose(self, rhs1, rhs2, rtol, atol, equal_nan) -> None:
        assert not equal_nan
        self.binary_op(BinaryOpCode.ISCLOSE, rhs1, rhs2, True, (rtol, atol))

    def where(self, mask, one, two) -> None:
        self._local_task(lambda out, m, a, b: out.where(m, a, b), [mask, one, two])

    def convert(self, rhs, warn=True, nan_op=ConvertCode.NOOP, temporary=False) -> None:
        self._local_task(lambda out, s: out.convert(s, warn, nan_op, temporary), [rhs])

    def copy(self, rhs, deep=False) -> None:
        self._local_task(lambda out, s: out.copy(s, deep), [rhs])

    def fill(self, value) -> None:
        vlo, vhi = self.owned
        if vhi > vlo:
            self.local_rows(vlo, vhi).fill(value)
        self._invalidate()

    def _broadcast(self, shape):
        raise NotImplementedError("a partitioned array cannot be the source of a replicated task; "
                                  "gather() it first")

    # ------------------------------------------------------------------ views
    def get_item(self, key: Any) -> Any:
        if not isinstance(key, tuple):
            key = (key,)
        k0 = key[0] if key else slice(None)
        rest = tuple(key[1:])
        if k0 is Ellipsis:
            k0, rest = slice(None), (Ellipsis,) + rest
        if self.inner_key:
            # a view of a view: only further ROW slicing is supported
            if any(not (k is Ellipsis or k == slice(None)) for k in rest):
                raise NotImplementedError("column indexing of an already column-sliced "
                                          "partitioned view")
            rest = self.inner_key
        n = self.row1 - self.row0
        if isinstance(k0, slice):
            start, stop, step = k0.indices(n)
            if step != 1:
                raise NotImplementedError("stepped slices along the partitioned axis")
            stop = max(stop, start)
            return PartitionedArray(self.meta, self.row0 + start, self.row0 + stop, rest)
        if isinstance(k0, (int, np.integer)):
            # a single row: a view that only its owner holds
            idx = int(k0) + (n if int(k0) < 0 else 0)
            if not (0 <= idx < n):
                raise IndexError(f"index {int(k0)} is out of bounds for axis 0 with size {n}")
            return _RowView(self.meta, self.row0 + idx, rest)
        raise NotImplementedError("only basic indexing is supported on partitioned arrays")

    def set_item(self, key: Any, rhs: Any) -> None:
        view = self.get_item(key)
        if isinstance(view, _RowView):
            view.assign(rhs)
            return
        if isinstance(rhs, DeferredArray) and rhs.dtype != view.dtype:
            tmp = DeferredArray(Store.empty(rhs.shape, view.dtype))
            tmp.convert(rhs)
            rhs = tmp
        elif isinstance(rhs, PartitionedArray) and rhs.dtype != view.dtype:
            tmp = PartitionedArray.empty(rhs.shape, view.dtype, like=rhs)
            tmp.convert(rhs)
            rhs = tmp
        view.copy(rhs, deep=False)

    def transpose(self, axes):
        if tuple(axes) == tuple(range(self.ndim)):
            return self
        raise NotImplementedError("transposing a partitioned array needs an all-to-all; gather() "
                                  "it first")

    def squeeze(self, axis=None):
        raise NotImplementedError("squeeze on partitioned arrays")

    def reshape(self, newshape):
        if tuple(newshape) == tuple(self.shape):
            return self
        raise NotImplementedError("reshape on partitioned arrays")

    # ------------------------------------------------------------------ reductions
    def unary_reduction(self, op, src, where, orig_axis, axes, keepdims, args, initial) -> None:
        """This (partitioned) array is the RESULT of a reduction."""
        if isinstance(src, PartitionedArray):
            src.reduce_into(self, op, where, orig_axis, axes, keepdims, args, initial)
            return
        tmp = DeferredArray(Store.empty(self.shape, self.dtype))
        tmp.unary_reduction(op, src, where, orig_axis, axes, keepdims, args, initial)
        self.copy(tmp)

    def reduce_into(self, lhs: Any, op: UnaryRedCode, where: Any, orig_axis, axes, keepdims: bool,
                    args: Any, initial: Any) -> None:
        """unary_reduction with this array as the source and `lhs` (pre-existing, replicated or
        partitioned) as the result."""
        argred = op in _ARG_REDS
        rank, world = runtime.rank, runtime.world_size
        vlo, vhi = self.owned
        src_local = self.local_rows(vlo, vhi) if vhi > vlo else None
        where_local = None
        if where is not None:
            if not isinstance(where, PartitionedArray):
                wl = DeferredArray(where._broadcast(self.shape).slice(0, slice(vlo, vhi)))
            else:
                wl = where.local_rows(vlo, vhi)
            where_local = wl if vhi > vlo else None
        elem = self.dtype
        val_dtype = runtime.get_argred_type(elem) if argred else _val_dtype(op, elem)
        ident = np.array(_UNARY_RED_IDENTITIES[op](elem), dtype=val_dtype)
        scalar_out = all(d in axes for d in range(self.ndim))

        if scalar_out:
            partial = DeferredArray(Store.empty((1,), val_dtype))
            partial.fill(ident)
            if src_local is not None:
                origin = (vlo,) + (0,) * (self.ndim - 1)
                launch_scalar_red(op, partial.base, src_local.base,
                                  None if where_local is None else where_local.base,
                                  origin, self.shape, args)
            total = _allreduce(partial, op, argred, elem)
            _fold_into(lhs, total, op, argred, initial, elem)
            return

        if len(axes) > 1:
            raise NotImplementedError("Need support for reducing multiple dimensions")
        axis = axes[0]
        out_shape = tuple(n for d, n in enumerate(self.shape) if d != axis)
        if axis != 0:
            # independent per row block: the result is partitioned like the source
            if not isinstance(lhs, PartitionedArray) or not lhs.part.same_as(self.part):
                tmp = PartitionedArray.empty(lhs.shape, lhs.dtype, like=self)
                self.reduce_into(tmp, op, where, orig_axis, axes, keepdims, args, initial)
                _assign_any(lhs, tmp)
                return
            if src_local is not None:
                lhs.local_rows(vlo, vhi).unary_reduction(op, src_local, where_local, orig_axis, axes,
                                                         keepdims, args, initial)
            lhs._invalidate()
            return
        # axis 0 is the partitioned axis: full-width partial per rank, then allreduce
        partial = DeferredArray(Store.empty(out_shape, val_dtype))
        partial.fill(ident)
        if src_local is not None:
            promoted = partial.base.promote(0, src_local.shape[0])
            d_out, d_in = promoted.descriptor(), src_local.base.descriptor()
            d_w = None if where_local is None else where_local.base.descriptor()
            _lib.check(runtime.lib.cnb_unary_red(
                int(op), 0, ctypes.byref(d_out), ctypes.byref(d_in),
                None if d_w is None else ctypes.byref(d_w), vlo, runtime.stream))
        total = _allreduce(partial, op, argred, elem)
        if keepdims:
            total = DeferredArray(total.base.promote(0, 1))
        _fold_into(lhs, total, op, argred, initial, elem)


class _RowView:
    """`a[i]` / `a[i, ...] = v` on a partitioned array: only the owner of row i touches it."""

    def __init__(self, meta: _Shared, base_row: int, rest: Tuple) -> None:
        self.meta, self.base_row, self.rest = meta, base_row, rest

    @property
    def shape(self) -> Tuple[int, ...]:
        probe = Store(None, self.meta.dtype, self.meta.gshape[1:])
        return _basic_index(probe, self.rest).shape if self.rest else probe.shape

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def dtype(self) -> np.dtype:
        return self.meta.dtype

    def get_item(self, key):
        raise NotImplementedError("reading a single row of a partitioned array (a[i]) is a "
                                  "one-element-task path outside the hot-path scope")

    def _local(self) -> Optional[DeferredArray]:
        lo, hi = self.meta.part.bounds(runtime.rank)
        if not (lo <= self.base_row < hi):
            return None
        row = self.meta.local.base.project(0, self.base_row - (lo - self.meta.lead))
        if self.rest:
            row = _basic_index(row, self.rest)
        return DeferredArray(row)

    def assign(self, rhs: Any) -> None:
        if isinstance(rhs, PartitionedArray):
            raise NotImplementedError("assigning a partitioned array to a single row")
        local = self._local()
        if local is not None:
            if rhs.dtype != local.dtype:
                tmp = DeferredArray(Store.empty(rhs.shape, local.dtype))
                tmp.convert(rhs)
                rhs = tmp
            local.copy(rhs, deep=False)
        self.meta.ghost_valid = False


# ---------------------------------------------------------------------- helpers
def _val_dtype(op: UnaryRedCode, elem: np.dtype) -> np.dtype:
    if op in (UnaryRedCode.ALL, UnaryRedCode.ANY, UnaryRedCode.CONTAINS):
        return np.dtype(np.bool_)
    if op == UnaryRedCode.COUNT_NONZERO:
        return np.dtype(np.uint64)
    return elem


def _field_view(av: DeferredArray, field: str) -> DeferredArray:
    s = av.base
    dt, off = s.dtype.fields[field][0], s.dtype.fields[field][1]
    return DeferredArray(Store(s.buffer, dt, s.shape, s.strides, s.offset + off))


def _nccl_allreduce(buf: DeferredArray, red: UnaryRedCode) -> None:
    """In-place NCCL allreduce of a dense local buffer."""
    assert buf.base.is_c_contiguous
    _lib.check(runtime.lib.cnb_comm_allreduce(runtime.comm, buf.base.ptr, buf.base.ptr, buf.size,
                                              dtype_code(buf.dtype), int(red), runtime.stream))


def _needs_gather_fold(dtype: np.dtype, red: UnaryRedCode) -> bool:
    """Partials NCCL cannot combine the way the local kernels do: 16-bit integers (no NCCL type),
    complex PROD / MAX / MIN (NCCL only sums pairs of floats), and floating MAX / MIN — ncclMax /
    ncclMin propagate NaN differently from the reference's fold `if (b > a) a = b`, which would make
    the result depend on the number of ranks."""
    if dtype in (np.dtype(np.int16), np.dtype(np.uint16)):
        return True
    if dtype.kind == "c" and red != UnaryRedCode.SUM:
        return True
    return dtype.kind == "f" and red in (UnaryRedCode.MAX, UnaryRedCode.MIN)


def _gather_fold(partial: DeferredArray, red: UnaryRedCode) -> DeferredArray:
    """ncclAllGather of the per-rank partials, then the library's own reduction kernel along the
    rank axis (in rank order): the same fold, hence the same NaN behaviour, as on one GPU."""
    assert partial.base.is_c_contiguous
    world = runtime.world_size
    gathered = DeferredArray(Store.empty((world,) + tuple(partial.shape), partial.dtype))
    _lib.check(runtime.lib.cnb_comm_allgather(runtime.comm, partial.base.ptr, gathered.base.ptr,
                                              partial.size * partial.dtype.itemsize, runtime.stream))
    out = DeferredArray(Store.empty(partial.shape, partial.dtype))
    if partial.size == 1:
        flat = DeferredArray(gathered.base.reshape_contiguous((world,)))
        out.unary_reduction(red, flat, None, None, (0,), False, (), None)
    else:
        out.unary_reduction(red, gathered, None, 0, (0,), False, (), None)
    return out


def _allreduce(partial: DeferredArray, op: UnaryRedCode, argred: bool, elem: np.dtype
               ) -> DeferredArray:
    """Combine the per-rank partials (dense VAL array, same shape on every rank)."""
    if runtime.world_size == 1:
        return partial
    if not argred:
        red = _COMBINE[op]
        if _needs_gather_fold(partial.dtype, red):
            return _gather_fold(partial, red)
        _nccl_allreduce(partial, red)
        return partial
    # arg-reductions: ncclAllGather of the 16-byte {index, value} partials, then ONE kernel folds them in
    # rank order with the reduction's own tie-breaking (lowest global index; cnb_argval_fold)
    assert partial.base.is_c_contiguous
    world = runtime.world_size
    gathered = DeferredArray(Store.empty((world,) + tuple(partial.shape), partial.dtype))
    _lib.check(runtime.lib.cnb_comm_allgather(runtime.comm, partial.base.ptr, gathered.base.ptr,
                                              partial.size * partial.dtype.itemsize, runtime.stream))
    out = DeferredArray(Store.empty(partial.shape, partial.dtype))
    _lib.check(runtime.lib.cnb_argval_fold(int(op), dtype_code(elem), out.base.ptr, gathered.base.ptr,
                                           world, partial.size, runtime.stream))
    return out


def partitioned_binary_reduction(lhs: DeferredArray, op: BinaryOpCode, src1: Any, src2: Any,
                                 broadcast: Any, args: Any) -> bool:
    """BINARY_RED over row-partitioned operands: every rank folds its own rows into a local bool
    (the point tasks of deferred.py:3330-3364), the bools are ANDed across ranks (the
    ProdReduction<bool> the reference's runtime applies to the scalar result).  Returns False when
    the operands cannot be aligned row for row (the caller then gathers)."""
    lead = src1 if isinstance(src1, PartitionedArray) else src2
    shape = tuple(broadcast) if broadcast is not None else lead.shape
    if lead.shape != shape or lead.ndim == 0:
        return False
    vlo, vhi = lead.owned
    ops = []
    for s in (src1, src2):
        if s is lead:
            ops.append(None)
        elif isinstance(s, PartitionedArray):
            if s.shape != shape:
                return False
            ops.append(lead._operand(s, vlo, vhi))  # collective (row fetch) if misaligned
        else:
            cut = s.base.broadcast_to(shape) if s.shape != shape else s.base
            ops.append(DeferredArray(cut.slice(0, slice(vlo, vhi))))
    partial = DeferredArray(Store.empty((1,), np.bool_))
    partial.fill(np.array(True))
    if vhi > vlo:
        mine = lead.local_rows(vlo, vhi)
        a, b = [mine if o is None else o for o in ops]
        partial.binary_reduction(op, a, b, None, args)
    _nccl_allreduce(partial, UnaryRedCode.ALL)
    lhs.copy(DeferredArray(partial.base.reshape_contiguous(lhs.shape)), deep=True)
    return True


def _assign_any(lhs: Any, src: Any) -> None:
    if isinstance(lhs, PartitionedArray):
        lhs.copy(src)
    elif isinstance(src, PartitionedArray):
        lhs.copy(src.gather())
    else:
        lhs.copy(src)


def _fold_into(lhs: Any, total: DeferredArray, op: UnaryRedCode, argred: bool, initial: Any,
               elem: np.dtype) -> None:
    """lhs <- fold(prefill, total), prefill = `initial` or the Python-side identity
    (deferred.py:3207-3213); arg-reductions then extract the index (GETARG)."""
    if argred:
        arg = _field_view(total, "arg")
        if isinstance(lhs, PartitionedArray):
            lhs.copy(DeferredArray(arg.base.broadcast_to(lhs.shape)))
        else:
            lhs.copy(DeferredArray(arg.base.reshape_contiguous(lhs.shape))
                     if arg.base.is_c_contiguous else _reshape_to(arg, lhs.shape))
        return
    total_v = total if total.shape == tuple(lhs.shape) else _reshape_to(total, lhs.shape)
    prefill = initial if initial is not None else _UNARY_RED_IDENTITIES[op](elem)
    init = DeferredArray(Store.from_scalar(np.array(prefill, dtype=lhs.dtype)))
    lhs.binary_op(_FOLD_BINOP[_COMBINE[op]], init, total_v)


def _reshape_to(src: DeferredArray, shape) -> DeferredArray:
    dense = src
    if not src.base.is_c_contiguous:
        dense = DeferredArray(Store.empty(src.shape, src.dtype))
        dense.copy(src, deep=True)
    return DeferredArray(dense.base.reshape_contiguous(tuple(shape)))


def replicate(thunk: Any) -> DeferredArray:
    """Operand of a replicated (DeferredArray) task: partitioned inputs are gathered."""
    if isinstance(thunk, PartitionedArray):
        return thunk.gather()
    return thunk


# ---------------------------------------------------------------------- thunk factories
def _min_partition_volume() -> int:
    import os

    return int(os.environ.get("CUNUMERIC_B200_MIN_PARTITION", "65536"))


_partitioning = [True]


class replicated:
    """`with cunumeric_b200.replicated(): ...` — inside the block new arrays are NOT row-partitioned
    even in a multi-GPU job: every rank works on its own full copy (independent replicas of a
    workload that has no exchange step)."""

    def __enter__(self):
        self._old = _partitioning[0]
        _partitioning[0] = False
        return self

    def __exit__(self, *exc) -> None:
        _partitioning[0] = self._old


def create_empty_thunk(shape, dtype, inputs=None) -> Any:
    """runtime.create_empty_thunk (cunumeric/runtime.py:448-460): pick the thunk type for a new
    array.  Single-GPU: always a DeferredArray.  Multi-GPU: row-partitioned when it can be aligned
    with a partitioned input (the reference's alignment constraint) or when it is big enough to be
    worth tiling (the reference's MIN_GPU_CHUNK policy); small arrays are replicated."""
    shape = tuple(int(s) for s in shape)
    if runtime.world_size == 1 or len(shape) == 0:
        return DeferredArray(Store.empty(shape, dtype))
    like = None
    if not _partitioning[0]:
        inputs = [i for i in inputs or () if isinstance(getattr(i, "_thunk", None), PartitionedArray)]
        if not inputs:
            return DeferredArray(Store.empty(shape, dtype))
    any_partitioned = False
    for inp in inputs or ():
        thunk = getattr(inp, "_thunk", None)
        if isinstance(thunk, PartitionedArray):
            any_partitioned = True
            if like is None and thunk.shape and thunk.shape[0] == shape[0]:
                like = thunk
    if like is not None:
        return PartitionedArray.empty(shape, dtype, like=like)
    volume = int(np.prod(shape, dtype=np.int64))
    if not any_partitioned and shape[0] >= 2 * runtime.world_size \
            and volume >= _min_partition_volume():
        return PartitionedArray.empty(shape, dtype)
    return DeferredArray(Store.empty(shape, dtype))


def thunk_from_numpy(array: np.ndarray) -> Any:
    array = np.asarray(array)
    if runtime.world_size > 1 and _partitioning[0] and array.ndim >= 1 and array.shape[0] >= 2 * runtime.world_size \
            and array.size >= _min_partition_volume():
        return PartitionedArray.from_numpy(array)
    return DeferredArray.from_numpy(array)
