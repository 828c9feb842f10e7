"""sum(axis=1) over a few very long rows: one CTA per row (CNB_AXIS_ROW_SPLIT=0) vs the two-stage split."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn  # noqa: E402

cn.runtime.ensure_initialized()
lib = cn.runtime.lib
for shape in ((4, 1 << 28), (1, 1 << 29), (16, 1 << 26)):
    x = cn.full(shape, 0.5, dtype=np.float32)
    r = x.sum(axis=1)
    cn.synchronize()
    e0, e1 = lib.cnb_event_create(), lib.cnb_event_create()
    lib.cnb_event_record(e0, cn.runtime.stream)
    for _ in range(10):
        r = x.sum(axis=1)
    lib.cnb_event_record(e1, cn.runtime.stream)
    cn.synchronize()
    ms = ctypes.c_float()
    lib.cnb_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    n = int(np.prod(shape))
    print(f"split={os.environ.get('CNB_AXIS_ROW_SPLIT', '1')} {shape} sum axis=1: {ms.value / 10:.3f} ms "
          f"{n * 4 / (ms.value / 10) / 1e6:.0f} GB/s  result[0]={float(np.array(r)[0])}")
    del x, r

# per-launch breakdown of one split reduction (live trace: CUDA events around every launch)
from cunumeric_b200 import _lib  # noqa: E402

x = cn.full((4, 1 << 28), 0.5, dtype=np.float32)
r = x.sum(axis=1)
cn.synchronize()
_lib.check(lib.cnb_trace_start(64))
r = x.sum(axis=1)
cn.synchronize()
n = lib.cnb_trace_stop()
rec = _lib.cnb_trace_record_t()
for i in range(n):
    _lib.check(lib.cnb_trace_get(i, ctypes.byref(rec)))
    print(f"  launch {i}: task {rec.task} op {rec.op} dtype {rec.dtype} kind {rec.kernel_kind} "
          f"elems {rec.elems} ms {rec.ms:.4f}")
