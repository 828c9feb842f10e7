"""Lazy fusion of elementwise task chains.

The reference issues one task per NumPy operation (deferred.py:3139 `unary_op`, :3302 `binary_op`,
:3366 `where`, :1348 `convert`), so Black-Scholes moves 604 B per option and the stencil 128 B per
point although 20 B / ~40 B would do.  Here the thunk layer *captures* elementwise tasks instead of
launching them; a run of tasks over one iteration space becomes ONE generated kernel in which
every intermediate lives in registers and only the outputs somebody can still observe are stored.

Semantics are exactly those of eager, in-order execution:
* a captured chain is flushed before anything else can touch device memory — every consumer
  obtains pointers through `Store.ptr`, which flushes first (other tasks, reductions, copies to
  and from the host, NCCL, synchronize);
* a task joins the open chain only if it has the same shape and is free of cross-element hazards
  with it: an operand that overlaps something the chain writes must be the *same window* (then the
  value is taken from registers, element for element); otherwise the chain is flushed first;
* an output is dropped only if no `Store` window onto its buffer is alive any more (a Python
  temporary that was freed) — nobody could read it;
* the generated kernel composes the very same device functors as the per-task kernels
  (ops_binary.cuh / ops_unary.cuh / ops_convert.cuh, compiled with the same flags, -fmad=false), so
  every intermediate is rounded exactly as if it had been stored and re-loaded: results are
  bit-identical to op-by-op execution (tests/test_fusion.py).

Kernels are compiled with nvcc for sm_100a the second time a chain signature is seen (the first
time the chain simply runs op-by-op) and cached in memory and on disk
(cunumeric_b200/_fused_cache/); `__graft_entry__.build()` pre-compiles the chains of the benchmark
programs by tracing them without a device.  A chain whose kernel is not (yet) available, or whose
layout the generated kernel does not cover, is replayed op-by-op through the C ABI — always valid.

CUNUMERIC_B200_FUSION = 0 (off) | 1 (default) | always (compile on first sight).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import subprocess
import tempfile
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import MAX_DIM, dtype_code

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE_DIR = os.environ.get("CUNUMERIC_B200_FUSED_CACHE", os.path.join(_HERE, "_fused_cache"))
_CSRC = os.path.join(_HERE, "csrc")
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

MAX_TASKS = 192
MAX_INPUTS = 24
MAX_OUTPUTS = 8
THREADS = 256

_mode = os.environ.get("CUNUMERIC_B200_FUSION", "1").lower()


def set_mode(mode: str) -> str:
    """'0' off, '1' compile a chain the second time it is seen, 'always' compile at first sight."""
    global _mode
    flush()
    _plan_memo.clear()
    old, _mode = _mode, str(mode).lower()
    return old


def enabled() -> bool:
    return _mode not in ("0", "off", "false")


class _Window:
    """A (buffer, offset, shape, strides, dtype) window, held without a Store so that it does not
    count as a user of the buffer."""

    __slots__ = ("buffer", "offset", "shape", "strides", "dtype", "key", "lo", "hi")

    def __init__(self, store) -> None:
        self.buffer = store.buffer
        self.offset = store.offset
        self.shape = store.shape
        self.strides = store.strides
        self.dtype = store.dtype
        self.key = (id(store.buffer), store.offset, store.shape, store.strides, store.dtype.num)
        lo = hi = store.offset
        for n, s in zip(store.shape, store.strides):
            if s >= 0:
                hi += (n - 1) * s
            else:
                lo += (n - 1) * s
        self.lo, self.hi = lo, hi + store.dtype.itemsize

    @staticmethod
    def fresh(store) -> "_Window":
        """Window of a freshly allocated dense store (Store.empty): offset 0, spans its buffer."""
        w = object.__new__(_Window)
        w.buffer = buf = store.buffer
        w.offset, w.shape, w.strides, w.dtype = 0, store.shape, store.strides, store.dtype
        w.key = (id(buf), 0, store.shape, store.strides, store.dtype.num)
        w.lo, w.hi = 0, buf.nbytes
        return w

    def overlaps(self, other: "_Window") -> bool:
        return self.buffer is other.buffer and self.lo < other.hi and other.lo < self.hi

    def store(self):
        from .store import Store

        return Store(self.buffer, self.dtype, self.shape, self.strides, self.offset)


class _Task:
    __slots__ = ("kind", "op", "nan_op", "ins", "out", "window")

    def __init__(self, kind, op, nan_op, ins, out, window) -> None:
        self.kind, self.op, self.nan_op, self.ins, self.out, self.window = (
            kind, op, nan_op, ins, out, window)


class _Chain:
    def __init__(self) -> None:
        self.shape: Optional[Tuple[int, ...]] = None
        self.tasks: List[_Task] = []
        self.dtypes: List[np.dtype] = []      # per value id
        self.ext: List[Optional[_Window]] = []  # per value id: the window of an external input
        self.ext_index: dict = {}             # key -> value id
        self.written: dict = {}               # key -> (value id, window), latest write
        # hazard index: id(buffer) -> windows written / read through, so that an overlap test only
        # looks at windows of the same allocation (usually none)
        self.w_by_buf: dict = {}
        self.r_by_buf: dict = {}
        # id(buffer) -> (window, offset, rows, row_bytes, pitch): writes to this buffer are redirected
        # to a fresh block (WAR renaming, see capture())
        self.renamed: dict = {}


class _Opaque:
    """A non-elementwise step (e.g. the NCCL halo exchange) that must keep its place in program
    order between deferred chains.  `fn()` obtains its pointers when it RUNS.  `overlap` (optional)
    describes the step to the chain in front of it, which may then issue it itself — on another
    stream, between its boundary tiles and its interior tiles (see Overlap)."""

    __slots__ = ("fn", "overlap")

    def __init__(self, fn, overlap=None) -> None:
        self.fn = fn
        self.overlap = overlap


class Overlap:
    """What a queued exchange touches, for the chain that runs right before it: `buffer` is the
    allocation it works on, `send` / `recv` the byte ranges (relative to the start of the buffer) it
    reads / writes, `run(stream)` issues it on `stream`.  A chain that writes `buffer` through a
    renamed, TMA-tiled window launches the tile rows that produce the `send` bytes first, hands the
    exchange to the communication stream, and computes the interior while the bytes travel; it sets
    `done` so that the queue does not run the step again."""

    __slots__ = ("buffer", "send", "recv", "run", "done")

    def __init__(self, buffer, send, recv, run) -> None:
        self.buffer, self.send, self.recv, self.run = buffer, send, recv, run
        self.done = False


_PLAIN_DTYPES = frozenset(np.dtype(t) for t in (
    np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
    np.float16, np.float32, np.float64, np.complex64, np.complex128))
_chain = _Chain()
# Sealed chains / opaque steps that have not run yet, oldest first.  A chain is not launched the
# moment a hazard closes it: it waits here until `_MAX_SEALED` younger chains have been sealed (or
# anything needs device memory).  By then the Python temporaries it produced have usually been
# dropped (`average`, `work` of the previous Jacobi iteration), so their stores are elided — the
# liveness test runs at launch time, not at capture time.
_queue: list = []
_MAX_SEALED = max(0, int(os.environ.get("CUNUMERIC_B200_FUSION_DEPTH", "1")))
_RENAME = os.environ.get("CUNUMERIC_B200_RENAME", "1").lower() not in ("0", "off", "false")
# Halo exchange next to the interior tiles of the chain in front of it (Overlap): OFF by default.
# Measured at 8 GPUs (profiles/r02_scaling.md): 0.529 ms / iteration with it, 0.519 without — the
# persistent interior kernel fills every SM before the NCCL kernel of the other stream is placed, so
# the exchange still runs after it and the extra boundary launch is pure cost.
_OVERLAP = os.environ.get("CUNUMERIC_B200_HALO_OVERLAP", "0").lower() in ("1", "on", "true")
_flushing = False
_seen: dict = {}
_kernels: dict = {}   # signature hash -> (vec kernel, strided kernel, plan class) | None (= unusable)
stats = {"captured": 0, "fused_launches": 0, "fused_tasks": 0, "replayed_tasks": 0,
         "elided_tasks": 0, "compiled": 0, "renamed": 0, "deferred": 0, "tma_launches": 0,
         "fused_reductions": 0, "tma_refused": 0, "overlapped_exchanges": 0}


_rt: list = []


def _get_runtime():
    from .runtime import runtime

    _rt.append(runtime)
    return runtime


def pending() -> bool:
    return (bool(_chain.tasks) or bool(_queue)) and not _flushing


# ---------------------------------------------------------------------------------------------
# capture
# ---------------------------------------------------------------------------------------------
def capture(kind: str, op: int, lhs, rhs: Sequence[Any], nan_op: int = 0, fresh: bool = False) -> bool:
    """Try to defer an elementwise task `lhs = kind/op(rhs...)` (stores already broadcast to
    lhs.shape).  Returns False if the caller must launch it eagerly.  `fresh`: lhs was allocated for
    this task by the ufunc fast path (dense, numeric, the operands already have its shape): nothing
    can alias it yet, so the checks on the output side are skipped."""
    global _chain
    if _flushing or _mode in ("0", "off", "false"):
        return False
    runtime = _rt[0] if _rt else _get_runtime()
    if runtime.lib is None and not runtime.dry_run:
        runtime.ensure_initialized()  # no device -> fail loudly right here
    shape = lhs.shape
    if fresh:
        if lhs._size == 0 or len(shape) > MAX_DIM:
            return False
    else:
        if len(shape) > MAX_DIM or lhs.size == 0:
            return False
        for n, st in zip(shape, lhs.strides):
            if st == 0 and n > 1:
                return False
        if lhs.dtype.kind == "V":  # Argval structs
            return False
        for r in rhs:
            if r.shape != shape or r.dtype.kind == "V":
                return False
    c = _chain
    if c.tasks and (c.shape != shape or len(c.tasks) >= MAX_TASKS or
                    len(c.ext_index) + len(rhs) > MAX_INPUTS):
        seal()
        c = _chain
    out_w = lhs._win
    if out_w is None:
        out_w = lhs._win = _Window.fresh(lhs) if fresh else _Window(lhs)
    in_w = []
    for r in rhs:
        w = r._win
        if w is None:
            w = r._win = _Window(r)
        in_w.append(w)
    rename = None
    if c.tasks:
        hazard = war = False
        written = c.written
        for w in in_w:  # read of something the chain writes through a different window
            if w.key not in written:
                for ww in c.w_by_buf.get(id(w.buffer), ()):
                    if w.lo < ww.hi and ww.lo < w.hi:
                        hazard = True
        bid = id(out_w.buffer)
        if fresh:
            pass  # a buffer nobody has seen yet: no write-after-write / write-after-read possible
        elif not hazard and out_w.key not in written:
            if bid in c.renamed:  # a second write window on a renamed buffer: keep it simple
                hazard = True
            for ww in c.w_by_buf.get(bid, ()):  # write over a chain output, different window
                if out_w.lo < ww.hi and ww.lo < out_w.hi:
                    hazard = True
        if not hazard and not fresh:
            for rw in c.r_by_buf.get(bid, ()):  # write over something read through another window
                if rw.key != out_w.key and out_w.lo < rw.hi and rw.lo < out_w.hi:
                    war = True
            if war and bid not in c.renamed:
                # Write-after-read against the chain's own (shifted) operands — the stencil's
                # `center[:] = work` after reading north/east/west/south of the same grid.  Instead
                # of closing the chain, redirect the write to a FRESH block for the whole buffer
                # (register renaming at buffer granularity): the chain keeps reading the old block,
                # the bytes outside the written window are copied over (cnb_copy_complement), and
                # the buffer object switches to the new block once the kernel is queued.
                rename = _rename_geometry(c, out_w) if _RENAME else None
                if rename is None:
                    hazard = True
        if hazard:
            seal()
            c = _chain
            rename = None
    c.shape = shape
    ins = []
    for w in in_w:
        hit = c.written.get(w.key)
        if hit is not None:
            ins.append(hit[0])
            continue
        vid = c.ext_index.get(w.key)
        if vid is None:
            vid = len(c.dtypes)
            c.dtypes.append(w.dtype)
            c.ext.append(w)
            c.ext_index[w.key] = vid
            c.r_by_buf.setdefault(id(w.buffer), []).append(w)
            # a chain that has not run yet reads this buffer: an OLDER pending chain must not
            # elide the store that produces it, even if every Store onto it has died meanwhile
            w.buffer.readers += 1
        ins.append(vid)
    # an input window of THIS task that overlaps its own output through a different window is
    # resolved by the caller (DeferredArray._copy_if_overlapping) before we get here
    out = len(c.dtypes)
    c.dtypes.append(out_w.dtype)
    c.ext.append(None)
    if out_w.key not in c.written:
        c.w_by_buf.setdefault(id(out_w.buffer), []).append(out_w)
    if rename is not None:
        c.renamed[id(out_w.buffer)] = rename
    c.written[out_w.key] = (out, out_w)
    c.tasks.append(_Task(kind, int(op), int(nan_op), tuple(ins), out, out_w))
    stats["captured"] += 1
    return True


# reductions that may join a chain as its last step (map -> reduce fusion): value reductions without
# extra arguments whose accumulator the block-reduce helpers of cnb_reduce.cuh can carry
MAX_REDUCTIONS = 4
_FUSABLE_REDS: dict = {}   # UnaryRedCode value -> True, filled on first use


def _fusable_red(op: int) -> bool:
    if not _FUSABLE_REDS:
        from .config import UnaryRedCode as R

        for code in (R.SUM, R.PROD, R.MAX, R.MIN, R.ALL, R.ANY, R.COUNT_NONZERO, R.NANSUM, R.NANPROD,
                     R.NANMAX, R.NANMIN):
            _FUSABLE_REDS[int(code)] = True
    return int(op) in _FUSABLE_REDS


def capture_reduce(op: int, lhs, src, fill_value) -> bool:
    """Map -> reduce fusion: SCALAR_UNARY_RED of a value the OPEN chain produces (`sum(abs(a - b))`,
    `(x * y).sum()`, test_map_reduce.py:22-31, the convergence test of test_jacobi.py) joins the
    chain as a reduce task: the fused kernel folds the value straight out of registers (thread ->
    warp shuffle -> shared memory -> one partial per CTA -> the last CTA folds the partials in CTA
    order into the 1-element store), the mapped array itself is only stored if somebody can still
    observe it.  `lhs`: the 1-element result store; `fill_value`: what the reference pre-fills it
    with (identity or `initial`, deferred.py:3207-3213).  Returns False if the caller must launch
    the reduction task eagerly."""
    if _flushing or _mode in ("0", "off", "false") or not _fusable_red(op):
        return False
    c = _chain
    if not c.tasks or src.shape != c.shape or len(c.tasks) >= MAX_TASKS or \
            len(c.ext_index) + 1 > MAX_INPUTS or lhs.dtype.kind == "V" or src.dtype.kind == "V":
        return False
    if src.dtype not in _PLAIN_DTYPES or src.dtype == np.complex128:
        return False
    w = src._win
    if w is None:
        w = src._win = _Window(src)
    hit = c.written.get(w.key)
    if hit is None:
        return False   # reducing an array that already lives in memory: the plain kernel does that
    if sum(1 for t in c.tasks if t.kind == "R") >= MAX_REDUCTIONS:
        return False
    out_w = lhs._win
    if out_w is None:
        out_w = lhs._win = _Window(lhs)
    bid = id(out_w.buffer)
    if bid in c.w_by_buf or bid in c.r_by_buf or out_w.buffer.shared:
        return False   # the result store is brand-new in every use the API makes of this path
    from .store import Store

    init = Store.from_scalar(np.array(fill_value, dtype=lhs.dtype)).broadcast_to(c.shape)
    iw = _Window(init)
    vid = c.ext_index.get(iw.key)
    if vid is None:
        vid = len(c.dtypes)
        c.dtypes.append(iw.dtype)
        c.ext.append(iw)
        c.ext_index[iw.key] = vid
        c.r_by_buf.setdefault(id(iw.buffer), []).append(iw)
        iw.buffer.readers += 1
    out = len(c.dtypes)
    c.dtypes.append(out_w.dtype)
    c.ext.append(None)
    c.w_by_buf.setdefault(bid, []).append(out_w)
    c.written[out_w.key] = (out, out_w)
    c.tasks.append(_Task("R", int(op), 0, (hit[0], vid), out, out_w))
    stats["captured"] += 1
    stats["fused_reductions"] += 1
    return True


# ---------------------------------------------------------------------------------------------
# flush
# ---------------------------------------------------------------------------------------------
def seal() -> None:
    """Close the open chain.  It runs once `_MAX_SEALED` younger chains are sealed behind it, or at
    the next flush()."""
    global _chain
    if _flushing or not _chain.tasks:
        return
    _queue.append(_chain)
    _chain = _Chain()
    stats["deferred"] += 1
    _drain(_MAX_SEALED)


def enqueue(fn, overlap: Optional[Overlap] = None) -> None:
    """Run `fn()` in program order with respect to the deferred chains: right away if nothing is
    pending, else after everything captured so far (and before everything captured later)."""
    if _flushing or not pending():
        fn()
        return
    seal()
    if not _queue:
        fn()
        return
    _queue.append(_Opaque(fn, overlap if _OVERLAP else None))


def flush() -> None:
    """Run everything that is pending (fused if possible, else op-by-op), in program order."""
    global _chain
    if _flushing:
        return
    if _chain.tasks:
        _queue.append(_chain)
        _chain = _Chain()
    if _queue:
        _drain(0)


def _drain(keep: int) -> None:
    """Run the oldest pending steps until at most `keep` sealed chains remain."""
    global _flushing
    if _flushing:
        return
    _flushing = True
    try:
        while _queue:
            if sum(1 for q in _queue if isinstance(q, _Chain)) <= keep and \
                    not isinstance(_queue[0], _Opaque):
                break
            step = _queue.pop(0)
            if isinstance(step, _Opaque):
                if step.overlap is None or not step.overlap.done:
                    step.fn()
            else:
                nxt = _queue[0] if _queue else None
                _run_chain(step, nxt.overlap if isinstance(nxt, _Opaque) else None)
    finally:
        _flushing = False


def _rename_geometry(c: _Chain, w: _Window):
    """(window, offset, rows, row_bytes, pitch) if `w` is a pitched box (inner-contiguous rows, positive
    strides) that covers at least half of its buffer and is the chain's first write to it — the
    cases where copying the complement costs less than the write itself.  Else None."""
    buf = w.buffer
    if id(buf) in c.w_by_buf or buf.shared:
        return None
    item = w.dtype.itemsize
    dims = sorted(((st, n) for n, st in zip(w.shape, w.strides) if n != 1), reverse=True)
    if any(st <= 0 for st, _ in dims):
        return None
    # merge jointly contiguous dims, innermost first
    row_bytes, rest = item, []
    for st, n in reversed(dims):
        if not rest and st == row_bytes:
            row_bytes *= n
        else:
            rest.append((st, n))
    if len(rest) > 1:
        return None
    pitch, rows = rest[0] if rest else (row_bytes, 1)
    if rows > 1 and pitch < row_bytes:
        return None
    if 2 * rows * row_bytes < buf.nbytes:
        return None
    if w.offset + (rows - 1) * pitch + row_bytes > buf.nbytes:
        return None
    return (w, w.offset, rows, row_bytes, pitch)


_plan_memo: dict = {}   # structural key of a chain -> (entry, n_keep, output value ids, input value ids)


def _run_chain(c: _Chain, overlap: Optional[Overlap] = None) -> None:
    runtime = _rt[0] if _rt else _get_runtime()
    # this chain's own reads are resolved by this launch: what remains in `readers` are the reads
    # of YOUNGER pending chains
    for w in c.ext:
        if w is not None:
            w.buffer.readers -= 1
    # outputs somebody can still observe: latest write per window; the buffer still has a live
    # Store, or a pending chain reads it
    live = [(vid, w) for vid, w in c.written.values()
            if w.buffer.users > 0 or w.buffer.readers > 0]
    # A program that repeats (a time-step loop) produces structurally identical chains: value ids
    # are assigned in program order, so (tasks, live outputs, dtypes, scalar flags) identifies the
    # dead-code elimination result, the signature and the kernel without redoing that work.
    memo_key = (tuple((t.kind, t.op, t.nan_op, t.ins, t.out) for t in c.tasks),
                tuple(sorted(v for v, _ in live)),
                tuple(dt.num for dt in c.dtypes),
                tuple(w is not None and not any(w.strides) for w in c.ext))
    hit = _plan_memo.get(memo_key)
    if hit is not None and not runtime.dry_run:
        entry, n_keep, out_vids, in_vids = hit
        by_vid = {vid: w for vid, w in live}
        stats["elided_tasks"] += len(c.tasks) - n_keep
        if _launch(entry, c.shape, [by_vid[v] for v in out_vids], [c.ext[v] for v in in_vids],
                   n_keep, c.renamed, overlap=overlap):
            return
    needed = set(v for v, _ in live)
    keep: List[_Task] = []
    for t in reversed(c.tasks):
        if t.out in needed:
            keep.append(t)
            needed.update(t.ins)
    keep.reverse()
    if hit is None:
        stats["elided_tasks"] += len(c.tasks) - len(keep)
    if not keep:
        return
    if len(keep) == 1 or len(live) > MAX_OUTPUTS or hit is not None:
        _replay(c, keep)
        return
    ext_ids = sorted(v for v in needed if c.ext[v] is not None)
    # canonical numbering: external inputs in order of first use, then tasks in order
    order: dict = {}
    for t in keep:
        for v in t.ins:
            if c.ext[v] is not None and v not in order:
                order[v] = len(order)
    n_in = len(order)
    for t in keep:
        order[t.out] = len(order)
    assert len(ext_ids) == n_in
    # (dtype code, is-scalar): a stride-0 operand (Python scalar / 0-d array) is read once per
    # thread and does not count towards the bytes in flight
    in_codes = tuple((dtype_code(c.dtypes[v]), not any(c.ext[v].strides))
                     for v in order if c.ext[v] is not None)
    tasks_sig = tuple((t.kind, t.op, t.nan_op, tuple(order[v] for v in t.ins), order[t.out],
                       dtype_code(c.dtypes[t.out])) for t in keep)
    outs = sorted(((order[v], dtype_code(w.dtype)) for v, w in live))
    sig = (in_codes, tasks_sig, tuple(outs))
    entry = _lookup(sig)
    if entry is None:
        _replay(c, keep)
        return
    # operands: stored outputs first (in signature order), then inputs (in signature order)
    out_pairs = sorted(((order[v], v, w) for v, w in live), key=lambda p: p[0])
    out_windows = [w for _, _, w in out_pairs]
    in_vids = [v for v in order if c.ext[v] is not None]
    in_windows = [c.ext[v] for v in in_vids]
    if runtime.dry_run:
        _launch(entry, c.shape, out_windows, in_windows, len(keep), c.renamed, dry=True)
        return
    if len(_plan_memo) > 4096:
        _plan_memo.clear()
    _plan_memo[memo_key] = (entry, len(keep), [v for _, v, _ in out_pairs], in_vids)
    if not _launch(entry, c.shape, out_windows, in_windows, len(keep), c.renamed, overlap=overlap):
        _replay(c, keep)


def _replay(c: _Chain, tasks: List[_Task]) -> None:
    """Op-by-op execution of (the needed part of) a chain through the per-task C ABI."""
    from .runtime import runtime

    if runtime.dry_run:
        return
    from . import deferred

    last_use = {}
    for i, t in enumerate(tasks):
        for v in t.ins:
            last_use[v] = i
    produced = {}
    for i, t in enumerate(tasks):
        ins = []
        for v in t.ins:
            w = c.ext[v] if c.ext[v] is not None else produced[v]
            ins.append(w.store())
        if t.kind == "R":
            # reduce task replayed as the reference issues it: pre-fill the result, then the
            # SCALAR_UNARY_RED task folds into it
            from .config import UnaryOpCode

            out_store = t.window.store()
            first = ins[1]
            while first.ndim > 0:
                first = first.project(0, 0)
            deferred.launch_elementwise("U", int(UnaryOpCode.COPY), 0, out_store,
                                        [first.broadcast_to(out_store.shape)])
            deferred.launch_scalar_red(t.op, out_store, ins[0], None, None, None, ())
        else:
            deferred.launch_elementwise(t.kind, t.op, t.nan_op, t.window.store(), ins)
        del ins
        produced[t.out] = t.window
        # a dead temporary (no live Store, no later reader in the chain) gives its block back right
        # away, as eager execution would: replaying a long chain must not hold every intermediate
        for v in set(t.ins):
            if last_use.get(v) == i and c.ext[v] is None:
                w = produced[v]
                bid = id(w.buffer)
                # ... and only if this window is the chain's sole handle on the allocation
                if w.buffer.users == 0 and len(c.w_by_buf.get(bid, ())) == 1 and \
                        bid not in c.r_by_buf and c.written[w.key][0] == v:
                    w.buffer.release()
    stats["replayed_tasks"] += len(tasks)


# ---------------------------------------------------------------------------------------------
# kernel lookup / generation / compilation
# ---------------------------------------------------------------------------------------------
_GENERATOR_VERSION = 12
_src_tag: List[str] = []


def _source_tag() -> str:
    """Digest of everything a generated kernel is built from besides its signature: the functor /
    engine headers and this generator.  Part of every cache key, so a cubin compiled against older
    headers is never reused."""
    if not _src_tag:
        h = hashlib.sha1(str(_GENERATOR_VERSION).encode())
        for name in ("cnb_common.cuh", "cnb_elementwise.cuh", "cnb_tma.cuh", "ops_math.cuh",
                     "ops_binary.cuh", "ops_unary.cuh", "ops_convert.cuh"):
            try:
                with open(os.path.join(_CSRC, name), "rb") as f:
                    h.update(f.read())
            except OSError:
                h.update(name.encode())
        try:
            with open(os.path.join(_INCLUDE, "cunumeric_b200.h"), "rb") as f:
                h.update(f.read())
        except OSError:
            pass
        _src_tag.append(h.hexdigest())
    return _src_tag[0]


def _hash(sig) -> str:
    knobs = "".join(f"{k}={os.environ.get(k, '')};" for k in
                    ("CNB_FUSED_B", "CNB_FUSED_U", "CNB_FUSED_MINBLOCKS", "CNB_FUSED_VEC_MINBLOCKS"))
    return hashlib.sha1((_source_tag() + knobs + repr(sig)).encode()).hexdigest()[:20]


def _lookup(sig):
    h = _hash(sig)
    if h in _kernels:
        return _kernels[h]
    path = os.path.join(_CACHE_DIR, f"fused_{h}.cubin")
    if not os.path.exists(path):
        n = _seen.get(h, 0) + 1
        _seen[h] = n
        if _mode != "always" and n < 2:
            return None  # first sighting: run op-by-op, compile when the chain comes back
        try:
            ok = _compile(sig, h, path)
        except OSError:  # read-only cache directory, no temp space, ...: stay op-by-op
            ok = False
        if not ok:
            _kernels[h] = None
            return None
    from .runtime import runtime

    if runtime.dry_run:
        return ("dry", "dry", None, _geometry(sig), None, sig)
    entry = _load(sig, h, path)
    _kernels[h] = entry
    return entry


def _nvcc() -> Optional[str]:
    exe = shutil.which("nvcc")
    if exe is None and os.path.exists("/usr/local/cuda/bin/nvcc"):
        exe = "/usr/local/cuda/bin/nvcc"
    return exe


def _compile(sig, h: str, path: str, source: Optional[str] = None) -> bool:
    exe = _nvcc()
    if exe is None:
        return False
    os.makedirs(_CACHE_DIR, exist_ok=True)
    src = generate_source(sig, h) if source is None else source
    # private temporaries, atomic renames: the ranks of one job may compile the same chain at once
    fd, tmp_src = tempfile.mkstemp(suffix=".cu", dir=_CACHE_DIR)
    with os.fdopen(fd, "w") as f:
        f.write(src)
    fd, tmp = tempfile.mkstemp(suffix=".cubin", dir=_CACHE_DIR)
    os.close(fd)
    cmd = [exe, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-fmad=false", "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177",
           "-I", _CSRC, "-I", _INCLUDE, "-cubin", "-o", tmp, tmp_src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    os.replace(tmp_src, os.path.join(_CACHE_DIR, f"fused_{h}.cu"))
    if res.returncode != 0:
        os.unlink(tmp)
        with open(os.path.join(_CACHE_DIR, f"fused_{h}.err"), "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        return False
    os.replace(tmp, path)  # atomic: ranks of one job may compile the same chain concurrently
    stats["compiled"] += 1
    return True


def _plan_type(nops: int):
    class Operand(ctypes.Structure):
        _fields_ = [("ptr", ctypes.c_void_p), ("inner_stride", ctypes.c_int64),
                    ("row_stride", ctypes.c_int64)]

    class Plan(ctypes.Structure):
        _fields_ = [("inner", ctypes.c_int64), ("rows", ctypes.c_int64),
                    ("tiles_per_row", ctypes.c_int64), ("num_tiles", ctypes.c_int64),
                    ("vec", ctypes.c_int32), ("out_pad", ctypes.c_int32),
                    ("op", Operand * nops),
                    ("red_partials", ctypes.c_void_p), ("red_ticket", ctypes.c_void_p)]

    return Plan


def _load(sig, h: str, path: str):
    from .runtime import runtime

    runtime.ensure_initialized()
    with open(path, "rb") as f:
        image = f.read()
    module = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(image, len(image))
    if runtime.lib.cnb_module_load(buf, len(image), ctypes.byref(module)) != 0:
        return None
    kernels = []
    for suffix in ("vec", "str"):
        k = ctypes.c_void_p()
        if runtime.lib.cnb_module_get_kernel(module, f"fused_{h}_{suffix}".encode(),
                                             ctypes.byref(k)) != 0:
            return None
        kernels.append(k)
    in_codes, tasks_sig, outs = sig
    geo = _geometry(sig)
    return (kernels[0], kernels[1], _plan_type(len(outs) + len(in_codes)), geo, buf, sig)


_SIZES = [1, 1, 2, 4, 8, 1, 2, 4, 8, 2, 4, 8, 8, 16]  # bytes per dtype code (bool ... complex128)


def _geometry(sig):
    """Elements per 128-bit chunk (E), chunks in flight per thread (U) and tile size — the rules of
    cnb_elementwise.cuh:EwShape applied to the operands the fused kernel actually loads / stores."""
    in_codes, tasks_sig, outs = sig
    in_sizes = [_SIZES[c] for c, _ in in_codes]
    arr_sizes = [_SIZES[c] for c, scalar in in_codes if not scalar]
    out_sizes = [_SIZES[c] for _, c in outs]
    reds = _reductions(sig)
    stored = [_SIZES[c] for v, c in outs if v not in reds]   # a fused reduction stores one element
    max_in = max(arr_sizes) if arr_sizes else 1
    if stored:
        e = max(1, min(16 // max(stored), 64 // max_in))
    else:
        e = max(1, 16 // (min(arr_sizes) if arr_sizes else 4))  # store-less: 16 bytes of the narrowest input
    in_bytes = sum(arr_sizes)
    u = max(1, min(8, 128 // max(1, e * in_bytes))) if in_bytes else 4
    # long chains are issue-bound, not memory-bound (Black-Scholes: ~170 instructions per element):
    # one chunk per thread keeps the register count low enough for 5-6 resident CTAs per SM, which
    # is what fills the issue slots (measured: 0.616 -> 0.586 ms per 1e8 options, profiles/r02_*)
    heavy = len(tasks_sig) > 16
    if heavy:
        u = 1
    # strided kernel: batches of B elements per thread.  Operands of a fused chain are often shifted
    # views of one array (the stencil's five neighbours), whose loads mostly hit L1/L2 but still
    # occupy the thread's load slots, so keep >= 4 elements (~160 B of requests) in flight
    b = max(4, min(16, 64 // max(1, in_bytes)))
    b = int(os.environ.get("CNB_FUSED_B") or b)           # tuning knobs (part of the cache key)
    u = int(os.environ.get("CNB_FUSED_U") or u)
    while e * u < b:
        u += 1
    return {"E": e, "U": u, "B": b, "TILE": THREADS * e * u, "in_sizes": in_sizes,
            "out_sizes": out_sizes, "minblocks": int(os.environ.get("CNB_FUSED_MINBLOCKS", "0")),
            "ctas_per_sm": 6 if heavy else 0,
            "red_slots": [j for j, (v, _) in enumerate(outs) if v in reds]}


def _gen_body(sig, L: List[str]) -> None:
    """Value types T<v> and the straight-line `body(inputs..., outputs...)` composed from the
    per-task functors (shared by every kernel flavour of a chain)."""
    in_codes, tasks_sig, outs = sig
    n_in = len(in_codes)
    vtype = {}
    for i, (code, _) in enumerate(in_codes):
        vtype[i] = code
    for kind, op, nan_op, ins, out, code in tasks_sig:
        vtype[out] = code
    for v, code in sorted(vtype.items()):
        L.append(f"using T{v} = type_of<{code}>;")
    reds = _reductions(sig)      # out value id -> (op, source value id, init value id)
    in_params = ", ".join(f"const T{i}& v{i}" for i in range(n_in))
    # a reduce task's "output" of the per-element body is the element to fold (type of its source)
    out_params = ", ".join(f"T{reds[v][1] if v in reds else v}& y{j}" for j, (v, _) in enumerate(outs))
    L.append(f"__device__ __forceinline__ void body({in_params}{', ' if in_params else ''}{out_params})\n{{")
    for kind, op, nan_op, ins, out, code in tasks_sig:
        if kind == "R":
            continue
        if kind == "B":
            a, b = ins
            L.append(f"  using F{out} = typename BinaryFn<{op}>::template fn<T{a}>;")
            L.append(f"  static_assert(F{out}::valid && std::is_same<typename F{out}::Out, T{out}>::value && "
                     f"std::is_same<typename F{out}::Rhs2, T{b}>::value, \"task {out}\");")
            L.append(f"  const T{out} v{out} = F{out}()(v{a}, v{b});")
        elif kind == "U":
            (a,) = ins
            L.append(f"  using F{out} = typename UnaryFn<{op}>::template fn<T{a}>;")
            L.append(f"  static_assert(F{out}::valid && std::is_same<typename F{out}::Out, T{out}>::value, "
                     f"\"task {out}\");")
            L.append(f"  const T{out} v{out} = F{out}()(v{a});")
        elif kind == "C":
            (a,) = ins
            L.append(f"  static_assert(ConvertFn<{nan_op}, T{out}, T{a}>::valid, \"task {out}\");")
            L.append(f"  T{out} v{out};\n  {{ Unused u_; ConvertFn<{nan_op}, T{out}, T{a}>()(v{out}, u_, v{a}, u_, u_); }}")
        elif kind == "W":
            m, a, b = ins
            L.append(f"  static_assert(std::is_same<T{m}, bool>::value && std::is_same<T{a}, T{out}>::value && "
                     f"std::is_same<T{b}, T{out}>::value, \"task {out}\");")
            L.append(f"  const T{out} v{out} = v{m} ? v{a} : v{b};")
        else:
            raise ValueError(kind)
    for j, (v, _) in enumerate(outs):
        L.append(f"  y{j} = v{reds[v][1] if v in reds else v};")
    L.append("}")


def _reductions(sig) -> dict:
    """Reduce tasks of a chain signature: {output value id: (redop, source value id, init value id)}."""
    return {t[4]: (t[1], t[3][0], t[3][1]) for t in sig[1] if t[0] == "R"}


def generate_source(sig, h: str) -> str:
    in_codes, tasks_sig, outs = sig
    geo = _geometry(sig)
    n_in, n_out = len(in_codes), len(outs)
    nops = n_in + n_out
    E, U, B = geo["E"], geo["U"], geo["B"]
    reds = _reductions(sig)
    red_j = [j for j, (v, _) in enumerate(outs) if v in reds]      # output slots that are reductions
    L: List[str] = []
    L.append(f"// generated by cunumeric_b200/fusion.py — fused chain {h}: {len(tasks_sig)} tasks, "
             f"{n_in} inputs, {n_out - len(red_j)} stored outputs"
             + (f", {len(red_j)} fused reduction(s)" if red_j else ""))
    L.append('#include "cnb_elementwise.cuh"\n#include "ops_binary.cuh"\n#include "ops_unary.cuh"\n'
             '#include "ops_convert.cuh"' + ('\n#include "cnb_reduce.cuh"\n#include "ops_reduce.cuh"' if red_j else "")
             + '\nusing namespace cnb;\nnamespace {')
    L.append(f"constexpr int NOPS = {nops}, E = {E}, U = {U}, B = {B}, TILE = {THREADS} * E * U;")
    L.append("struct FOperand { char* ptr; long long inner_stride; long long row_stride; };")
    # red_partials / red_ticket: scratch of the map -> reduce epilogue, filled in by cnb_launch_fused
    L.append("struct FPlan { long long inner, rows, tiles_per_row, num_tiles; int vec, out_pad; "
             "FOperand op[NOPS]; char* red_partials; unsigned int* red_ticket; };")
    L.append("""template <typename T, int N>
__device__ __forceinline__ void fload_vec(Pack<T, N>& r, const FOperand& o, long long off, long long e)
{
  if (o.inner_stride == 0) {
    Pack<T, 1> s;
    ld_bytes<sizeof(T)>(s.raw, o.ptr + off);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = s[0];
  } else {
    ld_bytes<sizeof(T) * N>(r.raw, o.ptr + off + e * (long long)sizeof(T));
  }
}
template <typename T>
__device__ __forceinline__ void fload_one(Pack<T, 1>& r, const FOperand& o, long long off, long long e)
{
  ld_bytes<sizeof(T)>(r.raw, o.ptr + off + e * o.inner_stride);
}""")
    scalar_in = [scalar for _, scalar in in_codes]
    _gen_body(sig, L)
    # element type the body hands out for output slot j (a reduction hands out its SOURCE element)
    ytype = [f"T{reds[v][1]}" if v in reds else f"T{v}" for v, _ in outs]
    for j in red_j:
        v = outs[j][0]
        op, src, init = reds[v]
        L.append(f"using R{j} = typename RedFn<{op}>::template fn<T{src}>;")
        L.append(f"static_assert(R{j}::valid && !R{j}::needs_index && sizeof(typename R{j}::Acc) <= 16 && "
                 f"std::is_same<typename R{j}::Val, T{v}>::value, \"reduce task {v}\");")
    L.append("}  // namespace")

    # operand k: outputs 0..n_out-1, inputs n_out..nops-1
    def rowoffs():
        return "\n".join(f"    const long long off{k} = row * plan.op[{k}].row_stride;" for k in range(nops))

    hoist = "".join(f"  Pack<T{i}, 1> s{i};\n  ld_bytes<sizeof(T{i})>(s{i}.raw, plan.op[{n_out + i}].ptr);\n"
                    for i in range(n_in) if scalar_in[i])
    hoist += "".join(f"  const R{j} red{j}(nullptr);\n  typename R{j}::Acc acc{j} = R{j}::identity();\n"
                     for j in red_j)
    tile_loop = """  for (long long tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
    long long row = 0, ct = tile;
    if (plan.rows > 1) {
      row = tile / plan.tiles_per_row;
      ct  = tile - row * plan.tiles_per_row;
    }
    const long long col0 = ct * TILE;
""" + rowoffs() + "\n"

    def consume(ind: str, idx: str, addr) -> List[str]:
        """After body(): fold the reduction slots, store the others.  `idx`: element index into the
        r{j} packs ("" = whole pack, for vector stores); `addr(j)`: store address of slot j."""
        s = []
        for j in red_j:
            s.append(f"{ind}acc{j} = R{j}::fold(acc{j}, red{j}.convert(r{j}[{idx or 0}], 0));")
        return s

    def scalar_elem(ind: str, e: str) -> str:
        s = []
        for i in range(n_in):
            if not scalar_in[i]:
                s.append(f"{ind}Pack<T{i}, 1> a{i};\n{ind}fload_one<T{i}>(a{i}, plan.op[{n_out + i}], off{n_out + i}, {e});")
        for j in range(n_out):
            s.append(f"{ind}Pack<{ytype[j]}, 1> r{j};")
        args = ", ".join([f"s{i}[0]" if scalar_in[i] else f"a{i}[0]" for i in range(n_in)] +
                         [f"r{j}[0]" for j in range(n_out)])
        s.append(f"{ind}body({args});")
        s += consume(ind, "0", None)
        for j, (v, _) in enumerate(outs):
            if j not in red_j:
                s.append(f"{ind}st_bytes<sizeof(T{v})>(plan.op[{j}].ptr + off{j} + {e} * plan.op[{j}].inner_stride, r{j}.raw);")
        return "\n".join(s)

    def vector_chunk(ind: str, load, addr_expr) -> List[str]:
        """U chunks of E elements per thread: all loads first, then compute + fold / store."""
        s = []
        for i in range(n_in):
            if not scalar_in[i]:
                s.append(f"{ind}Pack<T{i}, E> a{i}[U];")
        s.append(f"#pragma unroll\n{ind}for (int u = 0; u < U; ++u) {{")
        s.append(f"{ind}  const long long e = {addr_expr};")
        for i in range(n_in):
            if not scalar_in[i]:
                s.append(f"{ind}  " + load(i))
        s.append(f"{ind}}}\n#pragma unroll\n{ind}for (int u = 0; u < U; ++u) {{")
        s.append(f"{ind}  const long long e = {addr_expr};")
        for j in range(n_out):
            s.append(f"{ind}  Pack<{ytype[j]}, E> r{j};")
        args = ", ".join([f"s{i}[0]" if scalar_in[i] else f"a{i}[u][i_]" for i in range(n_in)] +
                         [f"r{j}[i_]" for j in range(n_out)])
        s.append(f"#pragma unroll\n{ind}  for (int i_ = 0; i_ < E; ++i_) {{\n{ind}    body({args});")
        s += consume(ind + "    ", "i_", None)
        s.append(f"{ind}  }}")
        for j, (v, _) in enumerate(outs):
            if j not in red_j:
                s.append(f"{ind}  st_bytes<sizeof(T{v}) * E>(plan.op[{j}].ptr + off{j} + e * (long long)sizeof(T{v}), r{j}.raw);")
        s.append(f"{ind}}}")
        return s

    def epilogue() -> str:
        """map -> reduce: block fold, one partial per CTA, the last CTA (ticket) folds the partials in
        CTA order into the pre-filled 1-element result (reduce-accessor semantics) — the finishing
        step of scalar_red.inl, once per fused reduction."""
        if not red_j:
            return ""
        s = ["  {", "    __shared__ bool is_last;"]
        for n, j in enumerate(red_j):
            s.append(f"    __shared__ RawSmem<typename R{j}::Acc, RED_WARPS> rs{j};")
            s.append(f"    acc{j} = block_reduce<R{j}>(acc{j}, rs{j}.ptr());")
        s.append("    if (threadIdx.x == 0) {")
        for n, j in enumerate(red_j):
            s.append(f"      *reinterpret_cast<typename R{j}::Acc*>(plan.red_partials + "
                     f"((size_t){n} * gridDim.x + blockIdx.x) * 16) = acc{j};")
        s.append("      __threadfence();\n      is_last = atomicAdd(plan.red_ticket, 1u) == gridDim.x - 1;\n    }")
        s.append("    __syncthreads();\n    if (is_last) {\n      __threadfence();")
        s.append(f"      const int per = ((int)gridDim.x + {THREADS} - 1) / {THREADS};\n"
                 "      const int lo = (int)threadIdx.x * per, hi = min(lo + per, (int)gridDim.x);")
        for n, j in enumerate(red_j):
            v = outs[j][0]
            init = reds[v][2]
            s.append(f"      {{\n        typename R{j}::Acc total = R{j}::identity();\n"
                     f"        for (int i = lo; i < hi; ++i) {{\n          typename R{j}::Acc p;\n"
                     f"          ld_bytes<sizeof(p)>(&p, plan.red_partials + ((size_t){n} * gridDim.x + i) * 16);\n"
                     f"          total = R{j}::fold(total, p);\n        }}\n"
                     f"        total = block_reduce<R{j}>(total, rs{j}.ptr());\n"
                     f"        if (threadIdx.x == 0)\n"
                     f"          *reinterpret_cast<T{v}*>(plan.op[{j}].ptr) = "
                     f"R{j}::finish(R{j}::fold(R{j}::lift(s{init}[0]), total));\n      }}")
        s.append("      if (threadIdx.x == 0) *plan.red_ticket = 0;\n    }\n  }")
        return "\n".join(s)

    # ---- vector kernel
    # issue-bound chains: 5 CTAs / SM (measured best for Black-Scholes: 4 -> 0.54 ms, 5 -> 0.53, 6 spills)
    vec_min = int(os.environ.get("CNB_FUSED_VEC_MINBLOCKS") or 5)
    vec_bounds = f"{THREADS}, {vec_min}" if geo["ctas_per_sm"] else f"{THREADS}"
    L.append(f'extern "C" __global__ void __launch_bounds__({vec_bounds}) fused_{h}_vec(const __grid_constant__ FPlan plan)\n{{')
    L.append("  const int tid = threadIdx.x;\n" + hoist)
    # dense 1-D fast path (plan.vec == 2: one row, every array operand contiguous): no row / tile
    # arithmetic and no broadcast tests per tile — on an issue-bound chain (Black-Scholes) the generic
    # per-tile prologue is ~10 % of all instructions
    L.append("  if (plan.vec == 2) {\n    const long long full = plan.inner / TILE;")
    for k in range(nops):
        L.append(f"    const long long off{k} = 0;")
    L.append("    for (long long tile = blockIdx.x; tile < full; tile += gridDim.x) {")
    L += vector_chunk(
        "      ",
        lambda i: f"ld_bytes<sizeof(T{i}) * E>(a{i}[u].raw, plan.op[{n_out + i}].ptr + e * (long long)sizeof(T{i}));",
        f"tile * TILE + ((long long)u * {THREADS} + tid) * E")
    L.append("    }")
    L.append("    if (full * TILE < plan.inner && full %% gridDim.x == blockIdx.x) {\n"
             "      for (long long e = full * TILE + tid; e < plan.inner; e += %d) {" % THREADS)
    L.append(scalar_elem("        ", "e"))
    L.append("      }\n    }\n  } else {")
    L.append(tile_loop)
    L.append("    if (col0 + TILE > plan.inner) {\n      for (long long e = col0 + tid; e < plan.inner; e += %d) {" % THREADS)
    L.append(scalar_elem("        ", "e"))
    L.append("      }\n      continue;\n    }")
    L += vector_chunk(
        "    ",
        lambda i: f"fload_vec<T{i}, E>(a{i}[u], plan.op[{n_out + i}], off{n_out + i}, e);",
        f"col0 + (long long)(u * {THREADS} + tid) * E")
    L.append("  }\n  }")
    L.append(epilogue())
    L.append("}")

    # ---- strided kernel (coalesced element accesses, batches of B, store-aligned rows)
    v0 = outs[0][0]
    lb = f"{THREADS}, {geo['minblocks']}" if geo["minblocks"] else f"{THREADS}"
    L.append(f'extern "C" __global__ void __launch_bounds__({lb}) fused_{h}_str(const __grid_constant__ FPlan plan)\n{{')
    L.append("  const int tid = threadIdx.x;\n" + hoist)
    L.append(tile_loop)
    L.append(f"""    constexpr int N = E * U;
    long long shift = 0;
    if (plan.out_pad != 0)
      shift = static_cast<long long>((reinterpret_cast<unsigned long long>(plan.op[0].ptr + off0) & 127ull) / sizeof(T{v0}));
    const long long tbase = col0 - shift;
#pragma unroll 1
    for (int j0 = 0; j0 < N; j0 += B) {{""")
    for i in range(n_in):
        if not scalar_in[i]:
            L.append(f"      Pack<T{i}, 1> a{i}[B];")
    L.append("#pragma unroll\n      for (int j = 0; j < B; ++j) {\n        const long long e = tbase + (long long)(j0 + j) * %d + tid;\n        if (j0 + j < N && e >= 0 && e < plan.inner) {" % THREADS)
    for i in range(n_in):
        if not scalar_in[i]:
            L.append(f"          fload_one<T{i}>(a{i}[j], plan.op[{n_out + i}], off{n_out + i}, e);")
    L.append("        }\n      }\n#pragma unroll\n      for (int j = 0; j < B; ++j) {\n        const long long e = tbase + (long long)(j0 + j) * %d + tid;\n        if (j0 + j < N && e >= 0 && e < plan.inner) {" % THREADS)
    for j in range(n_out):
        L.append(f"          Pack<{ytype[j]}, 1> r{j};")
    args = ", ".join([f"s{i}[0]" if scalar_in[i] else f"a{i}[j][0]" for i in range(n_in)] +
                     [f"r{j}[0]" for j in range(n_out)])
    L.append(f"          body({args});")
    L += consume("          ", "0", None)
    for j, (v, _) in enumerate(outs):
        if j not in red_j:
            L.append(f"          st_bytes<sizeof(T{v})>(plan.op[{j}].ptr + off{j} + e * plan.op[{j}].inner_stride, r{j}.raw);")
    L.append("        }\n      }\n    }\n  }")
    L.append(epilogue())
    L.append("}")
    return "\n".join(L) + "\n"


# ---------------------------------------------------------------------------------------------
# TMA-staged flavour: pitched 2-D operands through cp.async.bulk.tensor.2d
# ---------------------------------------------------------------------------------------------
# The vector kernel needs every operand 16-byte aligned and inner-contiguous; views such as the
# stencil's center / north / east / west / south (one element off a 16-byte boundary, row pitch
# N + 2) used to fall to the element-wise strided kernel (0.63 of the HBM roofline).  The TMA
# flavour serves any chain whose array operands are row-pitched windows (inner-contiguous, pitch a
# multiple of 16 bytes): per operand BUFFER one tensor map; per tile one box per "shift group"
# (windows of one buffer that are small shifts of each other share a box with a halo, so one fetch
# serves all five stencil operands); TR x TC tiles, S boxes in flight per CTA issued by one thread;
# operands are gathered from shared memory into registers, the stage is handed back, results go
# straight to global memory.  Box origins are rounded down to 16 bytes (the engine rejects others)
# and the remainder becomes a run-time column shift into the tile.
TMA_TC = 128                                          # tile columns
TMA_TR = int(os.environ.get("CNB_TMA_TR", "16"))      # tile rows
TMA_RPT = TMA_TR * TMA_TC // 256                      # rows per thread (256 threads per CTA)
TMA_MAX_STAGES = int(os.environ.get("CNB_TMA_STAGES", "4"))
TMA_SMEM_BUDGET = 88 * 1024                  # per CTA: two CTAs per SM
TMA_MAX_SHIFT_ROWS, TMA_MAX_SHIFT_COLS = 8, 32
_TMA = os.environ.get("CUNUMERIC_B200_TMA", "1").lower() not in ("0", "off", "false")
_TMA_CSHIFT = os.environ.get("CUNUMERIC_B200_TMA_CSHIFT", "1") != "0"
_tma_kernels: dict = {}    # hash -> (kernel, meta) | None
_tma_recipes: dict = {}    # window keys -> layout | None


def _tma_layout(sig, shape, out_windows, in_windows, dims):
    """Group structure of a launch, or None if the TMA flavour does not apply.
    Returns (lay, groups): lay[i] = None (scalar input) | (g, dr, dc); groups[g] = dict(buffer,
    row_stride, itemsize, r0, c0, hr, hc) with (r0, c0) the buffer coordinates of the group's
    top-left member."""
    in_codes, tasks_sig, outs = sig
    if len(dims) != 2:
        return None
    rows, row_st = dims[0]
    inner, inner_st = dims[1]
    n_out = len(out_windows)
    if rows < 16 or inner < TMA_TC // 2:
        return None
    for k, w in enumerate(out_windows):
        if inner_st[k] != w.dtype.itemsize or row_st[k] <= 0:
            return None
    lay: list = []
    groups: list = []
    for i, w in enumerate(in_windows):
        k = n_out + i
        item = w.dtype.itemsize
        if in_codes[i][1]:          # scalar (all strides 0)
            lay.append(None)
            continue
        rs = row_st[k]
        if inner_st[k] != item or rs <= 0 or rs % 16 or item not in (1, 2, 4, 8):
            return None
        r, rem = divmod(w.lo, rs)
        if rem % item:
            return None
        c = rem // item
        width, height = rs // item, w.buffer.nbytes // rs
        if c + inner > width or r + rows > height:
            return None
        for g, grp in enumerate(groups):
            if grp["buffer"] is w.buffer and grp["row_stride"] == rs and grp["itemsize"] == item and \
                    abs(r - grp["members"][0][1]) <= TMA_MAX_SHIFT_ROWS and \
                    abs(c - grp["members"][0][2]) <= TMA_MAX_SHIFT_COLS:
                grp["members"].append((i, r, c))
                break
        else:
            groups.append({"buffer": w.buffer, "row_stride": rs, "itemsize": item,
                           "width": width, "height": height, "members": [(i, r, c)]})
        lay.append(None)  # placeholder, filled below
    if not groups or len(groups) > 8:
        return None
    for g, grp in enumerate(groups):
        r0 = min(m[1] for m in grp["members"])
        c0 = min(m[2] for m in grp["members"])
        grp["r0"], grp["c0"] = r0, c0
        grp["hr"] = max(m[1] for m in grp["members"]) - r0
        grp["hc"] = max(m[2] for m in grp["members"]) - c0
        for i, r, c in grp["members"]:
            lay[i] = (g, r - r0, c - c0)
    return tuple(lay), groups


def _tma_geometry(sig, lay):
    """Compile-time constants of the TMA kernel for a (signature, group structure)."""
    in_codes, tasks_sig, outs = sig
    ng = 1 + max(e[0] for e in lay if e is not None)
    gs = []
    for g in range(ng):
        members = [(i, e) for i, e in enumerate(lay) if e is not None and e[0] == g]
        item = _SIZES[in_codes[members[0][0]][0]]
        hr = max(e[1] for _, e in members)
        hc = max(e[2] for _, e in members)
        a = max(1, 16 // item)                      # box origin alignment in elements
        w = -(-(TMA_TC + hc + a - 1) // a) * a      # box width incl. halo and alignment slack
        hgt = TMA_TR + hr
        if w > 256 or hgt > 256 or (w * item) % 16:
            return None
        gs.append({"item": item, "hr": hr, "hc": hc, "align": a, "w": w, "h": hgt,
                   "bytes": w * hgt * item, "type": members[0][0]})
    off = 0
    for g in gs:
        g["off"] = off
        off += -(-g["bytes"] // 128) * 128
    stage = off
    if stage > TMA_SMEM_BUDGET // 2:
        return None
    stages = max(2, min(TMA_MAX_STAGES, TMA_SMEM_BUDGET // stage))
    return {"groups": gs, "stage": stage, "stages": stages, "smem": stage * stages,
            "tx_bytes": sum(g["bytes"] for g in gs)}


def generate_tma_source(sig, lay, h: str) -> str:
    in_codes, tasks_sig, outs = sig
    geo = _tma_geometry(sig, lay)
    gs = geo["groups"]
    n_in, n_out, ng = len(in_codes), len(outs), len(gs)
    scalars = [i for i in range(n_in) if lay[i] is None]
    L: List[str] = []
    L.append(f"// generated by cunumeric_b200/fusion.py — fused chain {h} (TMA flavour): "
             f"{len(tasks_sig)} tasks, {n_in} inputs in {ng} tensor-map group(s), {n_out} stored outputs")
    L.append('#include "cnb_elementwise.cuh"\n#include "cnb_tma.cuh"\n#include "ops_binary.cuh"\n'
             '#include "ops_unary.cuh"\n#include "ops_convert.cuh"\nusing namespace cnb;\nnamespace {')
    L.append(f"constexpr int TC = {TMA_TC}, TR = {TMA_TR}, RPT = {TMA_RPT}, S = {geo['stages']}, "
             f"STAGE = {geo['stage']}, TX_BYTES = {geo['tx_bytes']}, NG = {ng};")
    L.append("struct alignas(64) TMap { unsigned char bytes[128]; };")
    L.append("struct TOut { char* ptr; long long row_stride; };")
    L.append("struct TParams {\n  TMap maps[NG];\n  long long inner, rows;\n"
             "  int tiles_x, num_tiles, cshift, ty_split, ty_skip, pad_;\n"
             "  int gx[NG], gy[NG], gshift[NG];\n"
             f"  TOut out[{n_out}];\n  const char* scalar[{max(1, len(scalars))}];\n"
             "  unsigned int* sched;   // {next tile, finished CTAs}: filled in by cnb_launch_fused_tma\n};")
    _gen_body(sig, L)
    L.append("}  // namespace")
    L.append(f'extern "C" __global__ void __launch_bounds__({THREADS}) fused_{h}_tma('
             "const __grid_constant__ TParams P)\n{")
    L.append("""  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[S];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&full[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();""")
    for n, i in enumerate(scalars):
        L.append(f"  Pack<T{i}, 1> s{i};\n  ld_bytes<sizeof(T{i})>(s{i}.raw, P.scalar[{n}]);")
    L.append("""  const unsigned long long policy = l2_evict_last();
  // Tiles are numbered row-major and handed out DYNAMICALLY: the producer thread draws the next tile
  // number from a global counter when it issues that tile's loads (S tiles ahead of its use) and
  // leaves it in a small shared ring for the consumers.  However unevenly the CTAs progress, the
  // tiles in flight on the chip stay a compact window of consecutive numbers — a band a few tile
  // rows high — so the halo rows / columns a tile shares with its neighbours are requested within
  // microseconds of each other and are served by L2.  (With a static round-robin assignment the
  // CTAs drift apart and every halo row is fetched from DRAM twice: measured 16.6 GB instead of
  // 12.8 GB read per sweep of a 12.8 GB grid, 5.4 ms instead of 3.7 ms.)
  __shared__ int tile_ring[16];
  auto issue = [&](int k) {
    const int t = (int)atomicAdd(P.sched, 1u);
    tile_ring[k & 15] = t;
    if (t < P.num_tiles) {
      int ty = t / P.tiles_x;
      const int tx = t - ty * P.tiles_x;
      if (ty >= P.ty_split) ty += P.ty_skip;   // a launch over a subset of the tile rows (see _launch_tma)
      const int s = k % S;
      const uint32_t bar = smem_u32(&full[s]);
      mbar_arrive_expect_tx(bar, TX_BYTES);""")
    for g, grp in enumerate(gs):
        L.append(f"      tma_load_2d(smem_u32(smem + s * STAGE + {grp['off']}), &P.maps[{g}], "
                 f"P.gx[{g}] + tx * TC, P.gy[{g}] + ty * TR, bar, policy);")
    L.append("""    }
  };
  if (tid == 0)
    for (int k = 0; k < S; ++k) issue(k);
  __syncthreads();
  const int cx = tid % TC, r0 = (tid / TC) * RPT;""")
    for g in range(ng):
        L.append(f"  const int sh{g} = P.gshift[{g}] + cx;")
    L.append("""  for (int k = 0;; ++k) {
    const int t = tile_ring[k & 15];
    if (t >= P.num_tiles) break;
    int ty = t / P.tiles_x;
    const int tx = t - ty * P.tiles_x;
    if (ty >= P.ty_split) ty += P.ty_skip;
    const int s = k % S;
    mbar_wait(smem_u32(&full[s]), (k / S) & 1);""")
    # gather: every (row, column) offset of every group some row of this thread needs
    for g, grp in enumerate(gs):
        ty_ = f"T{grp['type']}"
        L.append(f"    const {ty_}* tile{g} = reinterpret_cast<const {ty_}*>(smem + s * STAGE + {grp['off']});")
        need = sorted({(j + e[1], e[2]) for e in lay if e is not None and e[0] == g
                       for j in range(TMA_RPT)})
        for rr, cc in need:
            L.append(f"    const {ty_} g{g}_{rr}_{cc} = tile{g}[(r0 + {rr}) * {grp['w']} + sh{g} + {cc}];")
    L.append("""    __syncthreads();              // every thread holds its operands: the stage can be refilled
    if (tid == 0) issue(k + S);
    // tile columns are shifted left by `cshift` so that the stores of a warp start on a 128-byte
    // line of the first output (partial sectors at both ends of every warp store otherwise)
    const long long col = (long long)tx * TC + cx - P.cshift;
    const long long row = (long long)ty * TR + r0;
    if (col >= 0 && col < P.inner) {""")
    for j in range(TMA_RPT):
        L.append(f"      if (row + {j} < P.rows) {{")
        for o, (v, _) in enumerate(outs):
            L.append(f"        T{v} y{o};")
        args = []
        for i in range(n_in):
            if lay[i] is None:
                args.append(f"s{i}[0]")
            else:
                g, dr, dc = lay[i]
                args.append(f"g{g}_{j + dr}_{dc}")
        args += [f"y{o}" for o in range(n_out)]
        L.append(f"        body({', '.join(args)});")
        for o, (v, _) in enumerate(outs):
            L.append(f"        *reinterpret_cast<T{v}*>(P.out[{o}].ptr + (row + {j}) * P.out[{o}].row_stride + "
                     f"col * (long long)sizeof(T{v})) = y{o};")
        L.append("      }")
    L.append("    }\n  }")
    L.append("""  // the last CTA to finish re-arms the scheduler words for the next launch
  if (tid == 0 && atomicAdd(P.sched + 1, 1u) == gridDim.x - 1) {
    P.sched[0] = 0;
    P.sched[1] = 0;
  }
}""")
    return "\n".join(L) + "\n"


def _tma_params_type(ng: int, n_out: int, n_scalar: int):
    class TOut(ctypes.Structure):
        _fields_ = [("ptr", ctypes.c_void_p), ("row_stride", ctypes.c_int64)]

    class Tail(ctypes.Structure):
        _fields_ = [("inner", ctypes.c_int64), ("rows", ctypes.c_int64),
                    ("tiles_x", ctypes.c_int32), ("num_tiles", ctypes.c_int32),
                    ("cshift", ctypes.c_int32), ("ty_split", ctypes.c_int32),
                    ("ty_skip", ctypes.c_int32), ("pad_", ctypes.c_int32),
                    ("gx", ctypes.c_int32 * ng), ("gy", ctypes.c_int32 * ng),
                    ("gshift", ctypes.c_int32 * ng),
                    ("out", TOut * n_out), ("scalar", ctypes.c_void_p * max(1, n_scalar)),
                    ("sched", ctypes.c_void_p)]

    return Tail


def _lookup_tma(sig, lay):
    """Kernel of the TMA flavour for (signature, group structure): memory, disk cache, nvcc."""
    h = hashlib.sha1((_source_tag() + "tma" + repr((sig, lay, TMA_TC, TMA_TR, TMA_RPT,
                                                     TMA_MAX_STAGES, TMA_SMEM_BUDGET))
                      ).encode()).hexdigest()[:20]
    if h in _tma_kernels:
        return _tma_kernels[h]
    from .runtime import runtime

    geo = _tma_geometry(sig, lay)
    entry = None
    if geo is not None:
        path = os.path.join(_CACHE_DIR, f"fused_{h}.cubin")
        ok = os.path.exists(path)
        if not ok:
            try:
                ok = _compile(sig, h, path, source=generate_tma_source(sig, lay, h))
            except OSError:
                ok = False
        if ok and runtime.dry_run:
            return ("dry", h)
        if ok:
            with open(path, "rb") as f:
                image = f.read()
            module, kern = ctypes.c_void_p(), ctypes.c_void_p()
            buf = ctypes.create_string_buffer(image, len(image))
            if runtime.lib.cnb_module_load(buf, len(image), ctypes.byref(module)) == 0 and \
                    runtime.lib.cnb_module_get_kernel(module, f"fused_{h}_tma".encode(),
                                                      ctypes.byref(kern)) == 0:
                in_codes, tasks_sig, outs = sig
                n_scalar = sum(1 for e in lay if e is None)
                entry = (kern, geo, _tma_params_type(len(geo["groups"]), len(outs), n_scalar), buf)
    _tma_kernels[h] = entry
    return entry


def _overlap_split(ov: Overlap, out_windows, renamed, inner: int, rows: int, row_st, tiles_y: int):
    """(top, bottom) tile rows that must run before the exchange `ov` may start, or None if the chain
    cannot run around it.  The chain has to write ov.buffer through its renamed window only: then its
    reads of that buffer go to the OLD block, which the exchange does not touch, and the tile rows
    that write none of the exchanged bytes are independent of the exchange."""
    buf = ov.buffer
    geo = renamed.get(id(buf)) if renamed else None
    if geo is None:
        return None
    top = bot = 0
    found = False
    for k, w in enumerate(out_windows):
        if w.buffer is not buf:
            continue
        rs = row_st[k]
        if geo[0].key != w.key or rs <= 0 or found:
            return None
        found = True
        row_bytes = inner * w.dtype.itemsize
        for lo, hi in list(ov.send) + list(ov.recv):
            # window rows r with [offset + r rs, offset + r rs + row_bytes) intersecting [lo, hi)
            first = max(0, -(-(lo - w.offset - row_bytes + 1) // rs))
            last = min(rows - 1, (hi - 1 - w.offset) // rs)
            if first > last:
                continue
            if last + 1 <= rows - first:
                top = max(top, last + 1)
            else:
                bot = max(bot, rows - first)
    if not found:
        return None
    nt, nb = -(-top // TMA_TR), -(-bot // TMA_TR)
    if nt + nb == 0 or (nt + nb) * 4 > tiles_y:
        return None
    return nt, nb


def _launch_tma(sig, lay, groups, inner, rows, row_st, out_windows, in_windows, ptrs, algo,
                ntasks: int, commit, renamed=None, overlap: Optional[Overlap] = None) -> bool:
    """Launch the TMA flavour and commit the renamed blocks.  With an exchange to overlap: the tile
    rows that produce the bytes the exchange sends run first; the exchange is then issued on the
    communication stream, and the interior tile rows on the compute stream next to it; the compute
    stream waits for the exchange at the end."""
    from .runtime import runtime

    entry = _lookup_tma(sig, lay)
    if entry is None or entry[0] == "dry":
        return False
    kern, geo, tail_cls, _ = entry
    n_out = len(out_windows)
    ng = len(groups)
    ops = (_lib.cnb_tma_operand_t * ng)()
    tail = tail_cls()
    item0 = out_windows[0].dtype.itemsize
    cshift = (ptrs[0] % 128) // item0 if (ptrs[0] % item0 == 0 and _TMA_CSHIFT) else 0
    tail.cshift = cshift
    for g, (grp, gg) in enumerate(zip(groups, geo["groups"])):
        i0 = grp["members"][0][0]
        w0 = in_windows[i0]
        # base of the buffer block this launch reads (inputs are never renamed)
        base = ptrs[n_out + i0] - w0.lo
        ops[g].base = base
        ops[g].width, ops[g].height = grp["width"], grp["height"]
        ops[g].pitch_bytes, ops[g].elem_bytes = grp["row_stride"], grp["itemsize"]
        ops[g].box_width, ops[g].box_height = gg["w"], gg["h"]
        a = gg["align"]
        gx = ((grp["c0"] - cshift) // a) * a        # may be negative: zero-filled, never stored
        tail.gx[g] = gx
        tail.gshift[g] = grp["c0"] - cshift - gx
        tail.gy[g] = grp["r0"]
    tail.inner, tail.rows = inner, rows
    tiles_x = -(-(inner + cshift) // TMA_TC)
    tiles_y = -(-rows // TMA_TR)
    if tiles_x * tiles_y >= 2 ** 31 - 2 ** 20:
        return False
    for k in range(n_out):
        tail.out[k].ptr = ptrs[k]
        tail.out[k].row_stride = row_st[k]
    n = 0
    for i, e in enumerate(lay):
        if e is None:
            tail.scalar[n] = ptrs[n_out + i]
            n += 1
    tail.tiles_x = tiles_x
    lib = runtime.lib

    def launch(count: int, skip_at: int, skip: int) -> int:
        """`count` tile rows: launch row y is tile row y, or y + skip from launch row `skip_at` on."""
        tail.ty_split, tail.ty_skip = skip_at, skip
        tail.num_tiles = tiles_x * count
        share = count / tiles_y
        return lib.cnb_launch_fused_tma(kern, ops, ng, ctypes.byref(tail), ctypes.sizeof(tail),
                                        geo["smem"], tail.num_tiles, int(inner * rows * share),
                                        int(algo * share), ntasks, 2, runtime.stream)

    split = None
    if overlap is not None and not overlap.done:
        split = _overlap_split(overlap, out_windows, renamed, inner, rows, row_st, tiles_y)
    if split is None:
        rc = launch(tiles_y, tiles_y, 0)
    else:
        nt, nb = split
        # boundary: tile rows [0, nt) and [tiles_y - nb, tiles_y)
        rc = launch(nt + nb, nt, tiles_y - nt - nb)
    if rc == -4:     # CNB_ERR_UNSUPPORTED: the driver refused a tensor map — the strided flavour serves
        stats["tma_refused"] += 1
        return False
    _lib.check(rc)
    commit()         # the exchange below takes the pointers of the adopted blocks
    if split is not None:
        comm_stream = runtime.comm_stream()
        ev = runtime._event()
        _lib.check(lib.cnb_event_record(ev, runtime.stream))
        _lib.check(lib.cnb_stream_wait_event(comm_stream, ev))
        overlap.run(comm_stream)
        overlap.done = True
        _lib.check(lib.cnb_event_record(ev, comm_stream))
        # interior: tile rows [nt, tiles_y - nb)
        _lib.check(launch(tiles_y - nt - nb, 0, nt))
        _lib.check(lib.cnb_stream_wait_event(runtime.stream, ev))
        runtime._recycle_event(ev)
        stats["overlapped_exchanges"] += 1
    return True


# ---------------------------------------------------------------------------------------------
# launch
# ---------------------------------------------------------------------------------------------
def _canonical(shape, strides_list):
    """cnb ew_make_plan in Python: drop unit dims, order by the first operand's |stride|, merge
    jointly contiguous dims.  Returns [(extent, [stride per operand])] slowest first."""
    dims = [d for d in range(len(shape)) if shape[d] != 1]
    if not dims:
        return [(1, [0] * len(strides_list))]
    dims.sort(key=lambda d: -abs(strides_list[0][d]))
    merged = [[shape[dims[0]], [s[dims[0]] for s in strides_list]]]
    for d in dims[1:]:
        n = shape[d]
        st = [s[d] for s in strides_list]
        last = merged[-1]
        if all(last[1][k] == n * st[k] for k in range(len(st))):
            last[0] *= n
            last[1] = st
        else:
            merged.append([n, st])
    return [(n, st) for n, st in merged]


def _resolve_pointers(out_windows, in_windows, renamed):
    """Device pointers of the windows of one fused launch, and a `commit()` to call once the
    kernel is queued.  Inputs (and outputs of buffers that are not renamed) point into the
    buffers' current blocks.  An output window of a RENAMED buffer points into a fresh block that
    already holds a copy of everything outside the window; commit() makes the buffer object adopt
    that block and returns the old one to the allocator (safe in stream order: every later user
    is queued behind this kernel)."""
    from .runtime import runtime

    for w in out_windows:
        if w.buffer.ready_event is not None:
            runtime.wait_ready(w.buffer)
    for w in in_windows:
        if w.buffer.ready_event is not None:
            runtime.wait_ready(w.buffer)
    in_ptrs = [w.buffer.ptr + w.offset for w in in_windows]
    out_ptrs = []
    swaps = []
    fresh: dict = {}
    for w in out_windows:
        buf = w.buffer
        geo = renamed.get(id(buf)) if renamed else None
        if geo is None or geo[0].key != w.key:
            out_ptrs.append(buf.ptr + w.offset)
            continue
        if id(buf) not in fresh:
            old_ptr = buf.ptr
            new_base, new_ptr = runtime._take_block(buf.nbytes)
            _, offset, rows, row_bytes, pitch = geo
            _lib.check(runtime.lib.cnb_copy_complement(new_ptr, old_ptr, buf.nbytes, offset, rows,
                                                       row_bytes, pitch, runtime.stream))
            fresh[id(buf)] = new_ptr
            swaps.append((buf, new_base, new_ptr))
            stats["renamed"] += 1
        out_ptrs.append(fresh[id(buf)] + w.offset)

    def commit() -> None:
        for buf, new_base, new_ptr in swaps:
            runtime.adopt_block(buf, new_base, new_ptr)

    return out_ptrs + in_ptrs, commit


def _distinct_bytes(windows, dims_of, renamed=None, n_out: int = 0) -> int:
    """Algorithmic bytes of one fused launch: distinct bytes touched per buffer.  Windows of one
    buffer whose byte ranges overlap (the five shifted views of the stencil grid) are counted once:
    min(sum of their sizes, length of the union of their ranges)."""
    spans: dict = {}
    total = 0
    for k, w in enumerate(windows):
        # the output window of a renamed buffer lives in a different block than the inputs
        key = (id(w.buffer), bool(renamed) and k < n_out and id(w.buffer) in renamed)
        spans.setdefault(key, []).append((w.lo, w.hi, dims_of(k) * w.dtype.itemsize))
    for lst in spans.values():
        lst.sort()
        cur_lo, cur_hi, cur_n = lst[0]
        for lo, hi, n in lst[1:]:
            if lo < cur_hi:
                cur_hi = max(cur_hi, hi)
                cur_n = min(cur_n + n, cur_hi - cur_lo)
            else:
                total += cur_n
                cur_lo, cur_hi, cur_n = lo, hi, n
        total += cur_n
    return total


def _launch(entry, shape, out_windows, in_windows, ntasks: int, renamed=None, dry: bool = False,
            overlap: Optional[Overlap] = None) -> bool:
    """Pick the kernel flavour for this layout (128-bit vector / TMA-staged tiles / strided) and
    launch it.  `dry`: only make sure the kernels the layout needs are compiled (trace_only)."""
    from .runtime import runtime

    k_vec, k_str, plan_cls, geo, _, sig = entry
    windows = list(out_windows) + list(in_windows)
    red_slots = geo["red_slots"]
    zero = (0,) * len(shape)
    dims = _canonical(shape, [zero if k in red_slots else w.strides for k, w in enumerate(windows)])
    if len(dims) > 2:
        return False
    inner, inner_st = dims[-1]
    rows, row_st = (dims[0] if len(dims) == 2 else (1, [0] * len(windows)))
    sizes = geo["out_sizes"] + geo["in_sizes"]
    E = geo["E"]
    n_out = len(out_windows)
    vec = True
    dense1d = rows == 1
    scalar_flags = [False] * n_out + [sc for _, sc in sig[0]]
    for k, w in enumerate(windows):
        if k in red_slots:
            continue            # 1-element result of a fused reduction: written once by one thread
        size = sizes[k]
        bcast = inner_st[k] == 0 and k >= n_out
        align = size if bcast else min(16, size * E)
        if bcast and not scalar_flags[k]:
            dense1d = False     # an array operand broadcast along the row: generic vector path
        if not bcast and inner_st[k] != size and inner != 1:
            vec = False
        # blocks are at least 256-byte aligned: the window offset decides
        if w.offset % align or row_st[k] % align:
            vec = False
    tma = None
    if not vec and _TMA and not red_slots:
        tma = _tma_layout(sig, shape, out_windows, in_windows, dims)
        if tma is not None and _lookup_tma(sig, tma[0]) is None:
            tma = None
    if dry:
        return True
    ptrs, commit = _resolve_pointers(out_windows, in_windows, renamed)
    algo = _distinct_bytes(
        windows, lambda k: (inner if inner_st[k] != 0 else 1) * (rows if row_st[k] != 0 else 1),
        renamed, n_out)
    if tma is not None and _launch_tma(sig, tma[0], tma[1], inner, rows, row_st, out_windows,
                                       in_windows, ptrs, algo, ntasks, commit, renamed, overlap):
        stats["fused_launches"] += 1
        stats["fused_tasks"] += ntasks
        stats["tma_launches"] += 1
        return True
    plan = plan_cls()
    for k, w in enumerate(windows):
        plan.op[k].ptr = ptrs[k]
        plan.op[k].inner_stride = inner_st[k]
        plan.op[k].row_stride = row_st[k]
    tile = geo["TILE"]
    out_pad = 0
    if not vec and 0 not in red_slots and inner_st[0] == sizes[0] and sizes[0] < 128:
        out_pad = 128 // sizes[0] - 1
    plan.inner, plan.rows = inner, rows
    plan.tiles_per_row = (inner + out_pad + tile - 1) // tile
    plan.num_tiles = plan.tiles_per_row * rows
    plan.vec, plan.out_pad = (2 if (vec and dense1d and inner != 1) else int(vec)), out_pad
    _lib.check(runtime.lib.cnb_launch_fused(k_vec if vec else k_str, ctypes.byref(plan),
                                            ctypes.sizeof(plan), plan.num_tiles, inner * rows, algo,
                                            ntasks, geo["ctas_per_sm"], len(red_slots), runtime.stream))
    commit()
    stats["fused_launches"] += 1
    stats["fused_tasks"] += ntasks
    return True


# ---------------------------------------------------------------------------------------------
# tracing without a device (used by __graft_entry__.build to pre-compile benchmark chains)
# ---------------------------------------------------------------------------------------------
def trace_only(fn) -> int:
    """Run `fn()` with the runtime in dry-run mode: arrays are created lazily, elementwise tasks
    are captured, every chain that forms is generated + compiled into the disk cache, nothing is
    launched.  Returns the number of kernels compiled."""
    global _mode
    from .runtime import runtime

    flush()
    before = stats["compiled"]
    old_mode, old_dry = _mode, runtime.dry_run
    _mode, runtime.dry_run = "always", True
    try:
        fn()
        flush()
    finally:
        flush()
        _mode, runtime.dry_run = old_mode, old_dry
        drop_scalar_caches()
    return stats["compiled"] - before


def drop_scalar_caches() -> None:
    """Forget cached 0-d operands.  Scalars created during a dry run were never uploaded, so they
    must not survive into a run on a device."""
    from ._ufunc.ufunc import binary_ufunc
    from .runtime import runtime

    binary_ufunc._scalar_cache.clear()
    binary_ufunc._bcast_cache.clear()
    runtime._scalar_cache.clear()
