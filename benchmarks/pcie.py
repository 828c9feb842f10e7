"""Host<->device copy ceilings on this box (pinned memory, CUDA events): H2D alone, D2H alone, both at
once on separate streams — the bound of bench.py's e2e leg."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn
from cunumeric_b200 import _lib
cn.runtime.ensure_initialized()
rt, lib = cn.runtime, cn.runtime.lib
n = 1 << 30
h1, h2 = rt.pinned_empty((n,), np.uint8), rt.pinned_empty((n,), np.uint8)
h1[:] = 1
d1, d2 = rt.allocate(n), rt.allocate(n)
s1, s2 = lib.cnb_stream_create(), lib.cnb_stream_create()
def run(label, fn, nbytes, reps=5):
    fn(); lib.cnb_stream_synchronize(s1); lib.cnb_stream_synchronize(s2)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    lib.cnb_stream_synchronize(s1); lib.cnb_stream_synchronize(s2)
    dt = (time.perf_counter() - t0) / reps
    print(f"{label}: {nbytes / dt / 1e9:.1f} GB/s", flush=True)
run("H2D 1 GiB", lambda: lib.cnb_memcpy_h2d(d1.ptr, h1.ctypes.data, n, s1), n)
run("D2H 1 GiB", lambda: lib.cnb_memcpy_d2h(h2.ctypes.data, d2.ptr, n, s2), n)
def both():
    lib.cnb_memcpy_h2d(d1.ptr, h1.ctypes.data, n, s1); lib.cnb_memcpy_d2h(h2.ctypes.data, d2.ptr, n, s2)
run("H2D + D2H concurrently (sum)", both, 2 * n)
for chunk in (1 << 26, 1 << 24):
    def chunks():
        for o in range(0, n, chunk):
            lib.cnb_memcpy_h2d(d1.ptr + o, h1.ctypes.data + o, chunk, s1)
    run(f"H2D in {chunk >> 20} MiB chunks", chunks, n)
