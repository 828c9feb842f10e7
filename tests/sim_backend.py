"""TEST INFRASTRUCTURE ONLY: a host-memory stand-in for libcunumeric_b200.so, used by
tests/test_fusion_sim.py to exercise the capture / hazard / liveness logic of
cunumeric_b200/fusion.py on machines without a GPU.  It implements a handful of opcodes with
NumPy, both as per-task "launches" and as an interpreter for fused chains, so that a random NumPy
program run through the lazy thunk layer can be compared with plain NumPy.  Nothing under
cunumeric_b200/ imports this; the product has no CPU path."""
import ctypes

import numpy as np

from cunumeric_b200.config import BinaryOpCode, UnaryOpCode

DTYPES = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32,
          np.uint64, np.float16, np.float32, np.float64, np.complex64, np.complex128]

BINARY = {
    int(BinaryOpCode.ADD): np.add, int(BinaryOpCode.SUBTRACT): np.subtract,
    int(BinaryOpCode.MULTIPLY): np.multiply, int(BinaryOpCode.MAXIMUM): np.maximum,
    int(BinaryOpCode.MINIMUM): np.minimum, int(BinaryOpCode.GREATER): np.greater,
    int(BinaryOpCode.LESS): np.less, int(BinaryOpCode.LOGICAL_AND): np.logical_and,
    int(BinaryOpCode.LOGICAL_OR): np.logical_or,
}
UNARY = {
    int(UnaryOpCode.COPY): lambda x: x.copy(), int(UnaryOpCode.NEGATIVE): np.negative,
    int(UnaryOpCode.ABSOLUTE): np.absolute, int(UnaryOpCode.SQUARE): np.square,
}


from cunumeric_b200.config import UnaryRedCode  # noqa: E402

# fused reduce tasks: fold(pre-fill, reduce(source))
REDUCE = {
    int(UnaryRedCode.SUM): lambda init, x: init + x.sum(dtype=x.dtype),
    int(UnaryRedCode.PROD): lambda init, x: init * x.prod(dtype=x.dtype),
    int(UnaryRedCode.MAX): lambda init, x: np.maximum(init, x.max()),
    int(UnaryRedCode.MIN): lambda init, x: np.minimum(init, x.min()),
    int(UnaryRedCode.ALL): lambda init, x: np.logical_and(init, x.all()),
    int(UnaryRedCode.ANY): lambda init, x: np.logical_or(init, x.any()),
}


def _window(ptr, dtype, shape, strides):
    """NumPy view of device (= host) memory described by a base pointer and byte strides."""
    dtype = np.dtype(dtype)
    lo = hi = 0
    for n, s in zip(shape, strides):
        if n == 0:
            return np.empty(shape, dtype)
        if s >= 0:
            hi += (n - 1) * s
        else:
            lo += (n - 1) * s
    span = hi - lo + dtype.itemsize
    raw = (ctypes.c_uint8 * span).from_address(ptr + lo)
    return np.ndarray(shape=tuple(shape), dtype=dtype, buffer=raw, offset=-lo, strides=tuple(strides))


def _desc_view(ref):
    d = ref._obj
    shape = [d.shape[i] for i in range(d.ndim)]
    strides = [d.strides[i] for i in range(d.ndim)]
    return _window(d.ptr, DTYPES[d.dtype], shape, strides)


class SimLib:
    def __init__(self) -> None:
        self.blocks = {}
        self.launches = 0
        self.fused_launches = 0

    # ---- runtime
    def cnb_malloc(self, nbytes, stream):
        buf = np.full(int(nbytes) + 64, 0xCD, dtype=np.uint8)  # poison: an elided store shows up
        addr = buf.ctypes.data
        addr += (-addr) % 16
        self.blocks[addr] = buf
        return addr

    def cnb_free(self, base, stream):
        self.blocks.pop(base, None)
        return 0

    def cnb_mem_info(self, free_ref, total_ref):
        free_ref._obj.value = 1 << 34
        total_ref._obj.value = 1 << 34
        return 0

    def cnb_memcpy_h2d(self, dst, src, n, stream):
        ctypes.memmove(dst, src, n)
        return 0

    def cnb_memcpy_d2h(self, dst, src, n, stream):
        ctypes.memmove(dst, src, n)
        return 0

    def cnb_stream_synchronize(self, stream):
        return 0

    def cnb_stream_wait_event(self, stream, event):
        return 0

    def cnb_copy_complement(self, dst, src, nbytes, offset, rows, row_bytes, pitch, stream):
        """dst <- src outside the pitched window (include/cunumeric_b200.h)."""
        d = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(dst), dtype=np.uint8)
        s = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(src), dtype=np.uint8)
        keep = np.ones(nbytes, dtype=bool)
        for r in range(rows):
            keep[offset + r * pitch: offset + r * pitch + row_bytes] = False
        d[keep] = s[keep]
        self.launches += 1
        self.complement_copies = getattr(self, "complement_copies", 0) + 1
        return 0

    def cnb_launch_count(self):
        return self.launches

    # ---- per-task launches
    def cnb_fill(self, out, value, stream):
        v = _desc_view(out)
        raw = (ctypes.c_uint8 * v.dtype.itemsize).from_address(value.value if hasattr(value, "value") else value)
        v[...] = np.frombuffer(raw, dtype=v.dtype)[0]
        self.launches += 1
        return 0

    def cnb_unary_op(self, op, out, out2, inp, extra, stream):
        o, a = _desc_view(out), _desc_view(inp)
        o[...] = UNARY[op](a.copy())
        self.launches += 1
        return 0

    def cnb_binary_op(self, op, out, in1, in2, extra, stream):
        o, a, b = _desc_view(out), _desc_view(in1), _desc_view(in2)
        with np.errstate(all="ignore"):
            o[...] = BINARY[op](a.copy(), b.copy())
        self.launches += 1
        return 0

    def cnb_where(self, out, m, a, b, stream):
        o = _desc_view(out)
        o[...] = np.where(_desc_view(m).copy(), _desc_view(a).copy(), _desc_view(b).copy())
        self.launches += 1
        return 0

    # ---- reductions (reduce-accessor semantics: fold into the pre-filled output)
    _AXIS_RED = {
        int(UnaryRedCode.SUM): (np.add, lambda x, ax: x.sum(axis=ax, dtype=x.dtype)),
        int(UnaryRedCode.PROD): (np.multiply, lambda x, ax: x.prod(axis=ax, dtype=x.dtype)),
        int(UnaryRedCode.MAX): (np.maximum, lambda x, ax: x.max(axis=ax)),
        int(UnaryRedCode.MIN): (np.minimum, lambda x, ax: x.min(axis=ax)),
        int(UnaryRedCode.ALL): (np.logical_and, lambda x, ax: x.all(axis=ax)),
        int(UnaryRedCode.ANY): (np.logical_or, lambda x, ax: x.any(axis=ax)),
        int(UnaryRedCode.COUNT_NONZERO): (np.add, lambda x, ax: np.count_nonzero(x, axis=ax)),
    }

    def cnb_unary_red(self, op, axis, out, inp, where, flags, stream):
        assert where is None, "the stand-in has no masked reductions"
        fold, red = self._AXIS_RED[op]
        o, a = _desc_view(out), _desc_view(inp)
        first = [slice(None)] * o.ndim
        first[axis] = slice(0, 1)                 # the output is promoted (stride 0) along the axis
        target = o[tuple(first)]
        part = np.expand_dims(red(a.copy(), axis), axis)
        target[...] = fold(target.copy(), part).astype(o.dtype)
        self.launches += 1
        return 0

    def cnb_scalar_unary_red(self, op, out, inp, where, origin, gshape, extra, stream):
        assert where is None and extra is None
        fold, red = self._AXIS_RED[op]
        o, a = _desc_view(out), _desc_view(inp)
        o[...] = fold(o.copy(), red(a.copy(), None)).astype(o.dtype)
        self.launches += 1
        return 0

    def cnb_binary_red(self, op, out, in1, in2, extra, stream):
        assert op == int(BinaryOpCode.EQUAL), "the stand-in folds EQUAL only"
        o, a, b = _desc_view(out), _desc_view(in1), _desc_view(in2)
        o[...] = np.logical_and(o.copy(), np.array_equal(a, b))
        self.launches += 1
        return 0

    def cnb_convert(self, nan_op, out, inp, stream):
        o = _desc_view(out)
        o[...] = _desc_view(inp).copy().astype(o.dtype)
        self.launches += 1
        return 0


def window_view(w, ptr=None):
    return _window(w.buffer.ptr + w.offset if ptr is None else ptr, w.dtype, w.shape, w.strides)


def make_fused_launcher(lib: SimLib, schedule_rng=None):
    """Interpreter for a fused chain.  A real fused kernel may process the elements in ANY order,
    each element reading its inputs and writing its outputs independently of the others, so every
    launch picks one of three legal schedules at random:
      0  every external input is read before anything is written, outputs stored in program order
      1  row by row, ascending, the outputs of a row stored in REVERSE order
      2  row by row, descending
    If the capture rules let a cross-element hazard into a chain, at least one schedule disagrees
    with NumPy."""
    rng = schedule_rng or np.random.default_rng(0)

    def evaluate(sig, vals):
        in_codes, tasks, outs = sig
        with np.errstate(all="ignore"):
            for kind, op, nan_op, ins, out, code in tasks:
                args = [vals[v] for v in ins]
                if kind == "R":
                    vals[out] = np.asarray(REDUCE[op](np.asarray(args[1]).flat[0], args[0])).astype(DTYPES[code]).reshape(())
                    continue
                if kind == "B":
                    r = BINARY[op](*args)
                elif kind == "U":
                    r = UNARY[op](args[0])
                elif kind == "W":
                    r = np.where(*args)
                else:
                    r = args[0].astype(DTYPES[code])
                vals[out] = np.asarray(r, dtype=DTYPES[code])
        return vals

    def launch(entry, shape, out_windows, in_windows, ntasks, renamed=None, dry=False, overlap=None):
        from cunumeric_b200 import fusion

        _, sig = entry
        outs = sig[2]
        # the product's own pointer resolution (incl. WAR renaming: fresh block + complement copy)
        ptrs, commit = fusion._resolve_pointers(out_windows, in_windows, renamed)
        n_out = len(out_windows)
        outs_v = [window_view(w, p) for w, p in zip(out_windows, ptrs[:n_out])]
        ins_v = [window_view(w, p) for w, p in zip(in_windows, ptrs[n_out:])]

        def run_rows(rows):
            for i in rows:
                vals = evaluate(sig, {k: v[i:i + 1].copy() for k, v in enumerate(ins_v)})
                pairs = list(zip(outs, outs_v))
                for (v, code), o in (reversed(pairs) if i % 2 else pairs):
                    o[i:i + 1] = vals[v]

        # a queued exchange the chain may run around (fusion.Overlap), decided by the product's own
        # dependence analysis: boundary tile rows, then the exchange, then the interior
        split = None
        if overlap is not None and not overlap.done and len(shape) == 2 and \
                not any(t[0] == "R" for t in sig[1]):
            windows = list(out_windows) + list(in_windows)
            dims = fusion._canonical(shape, [w.strides for w in windows])
            if len(dims) == 2 and dims[0][0] == shape[0]:
                rows, row_st = dims[0]
                tiles_y = -(-rows // fusion.TMA_TR)
                split = fusion._overlap_split(overlap, out_windows, renamed, dims[1][0], rows, row_st,
                                              tiles_y)
        if split is not None:
            nt, nb = split
            top, bot = nt * fusion.TMA_TR, shape[0] - nb * fusion.TMA_TR
            assert 0 < top <= bot < shape[0] or nt == 0 or nb == 0
            run_rows(list(range(0, top)) + list(range(bot, shape[0])))
            commit()
            overlap.run(None)
            overlap.done = True
            fusion.stats["overlapped_exchanges"] += 1
            run_rows(range(bot - 1, top - 1, -1))
            lib.stored_outputs = getattr(lib, "stored_outputs", []) + [len(out_windows)]
            lib.launches += 1
            lib.fused_launches += 1
            return True
        mode = int(rng.integers(3)) if len(shape) >= 1 and shape[0] > 1 else 0
        if any(t[0] == "R" for t in sig[1]):
            mode = 0   # a reduction spans the rows: evaluate the chain in one piece
        if mode == 0:
            vals = evaluate(sig, {i: v.copy() for i, v in enumerate(ins_v)})
            for (v, code), o in zip(outs, outs_v):
                o[...] = vals[v]
        else:
            rows = range(shape[0]) if mode == 1 else range(shape[0] - 1, -1, -1)
            for i in rows:
                vals = evaluate(sig, {k: v[i:i + 1].copy() for k, v in enumerate(ins_v)})
                pairs = list(zip(outs, outs_v))
                for (v, code), o in (reversed(pairs) if mode == 1 else pairs):
                    o[i:i + 1] = vals[v]
        commit()
        lib.stored_outputs = getattr(lib, "stored_outputs", []) + [len(out_windows)]
        lib.launches += 1
        lib.fused_launches += 1
        return True

    return launch
