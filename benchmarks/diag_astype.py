"""Diagnostic: kernel time vs wall time of astype float64->float32 at 2^30 elements."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn
from cunumeric_b200 import _lib

n = 1 << 30
lib = None
def kernel_ms(fn, reps=5):
    global lib
    fn(); cn.synchronize()
    lib = cn.runtime.lib
    _lib.check(lib.cnb_trace_start(64))
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    cn.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    k = lib.cnb_trace_stop()
    rec = _lib.cnb_trace_record_t(); tot = 0.0
    for i in range(k):
        lib.cnb_trace_get(i, ctypes.byref(rec)); tot += rec.ms
    return wall, tot / reps, k // reps

for src, dst, val in ((np.float64, np.float32, 1.25), (np.float64, np.float32, 0.0), (np.float32, np.float64, 1.25), (np.int64, np.float64, 3), (np.complex128, np.complex64, 1.25)):
    a = cn.empty((n,), dtype=src); a.fill(val)
    out = cn.empty((n,), dtype=dst)
    w1 = kernel_ms(lambda: out._thunk.convert(a._thunk))
    w2 = kernel_ms(lambda: a.astype(dst))
    print(np.dtype(src).name, "->", np.dtype(dst).name, "fill", val, "| prealloc: wall %.3f ms kernel %.3f ms (%d launches) | astype: wall %.3f kernel %.3f (%d)" % (w1 + w2), flush=True)
    del a, out
