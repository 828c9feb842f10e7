"""Microbenchmark: GB/s of representative elementwise kernel variants for the current
CNB_EW_CTAS_PER_SM setting (read once per process).  Used to pick the grid-size policy."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn
from cunumeric_b200 import _lib

cn.runtime.ensure_initialized()
lib = cn.runtime.lib
e0, e1 = lib.cnb_event_create(), lib.cnb_event_create()

def timeit(fn0, reps=10):
    def fn():
        fn0()
        cn.flush()

    for _ in range(3): fn()
    cn.synchronize()
    lib.cnb_event_record(e0, cn.runtime.stream)
    for _ in range(reps): fn()
    lib.cnb_event_record(e1, cn.runtime.stream)
    cn.synchronize()
    ms = ctypes.c_float(); lib.cnb_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    return ms.value * 1e-3 / reps

def filled(shape, dt, v):
    a = cn.empty(shape, dtype=dt); a.fill(v); return a

res = {}
n = 1 << 28
for name, dt in (("f32", np.float32), ("f64", np.float64), ("f16", np.float16)):
    s = np.dtype(dt).itemsize
    a, b, out = filled((n,), dt, 1.5), filled((n,), dt, 2.5), cn.empty((n,), dtype=dt)
    res[f"aa_mul_{name}"] = 3 * s * n / timeit(lambda: cn.multiply(a, b, out=out)) / 1e9
    res[f"scalar_mul_{name}"] = 2 * s * n / timeit(lambda: cn.multiply(a, 0.5, out=out)) / 1e9
    res[f"unary_abs_{name}"] = 2 * s * n / timeit(lambda: cn.absolute(a, out=out)) / 1e9
    m = filled((n,), np.bool_, True)
    res[f"where_{name}"] = (1 + 3 * s) * n / timeit(lambda: out._thunk.where(m._thunk, a._thunk, b._thunk)) / 1e9
    ob = cn.empty((n,), dtype=np.bool_)
    res[f"greater_{name}"] = (2 * s + 1) * n / timeit(lambda: cn.greater(a, b, out=ob)) / 1e9
    del a, b, out, m, ob
N = 16384
g = filled((N + 2, N + 2), np.float64, 1.0)
w = cn.empty((N, N), dtype=np.float64)
c, no = g[1:-1, 1:-1], g[0:-2, 1:-1]
res["strided_add_f64"] = 24 * N * N / timeit(lambda: cn.add(c, no, out=w)) / 1e9
res["strided_copy_f64"] = 16 * N * N / timeit(lambda: c._thunk.copy(w._thunk)) / 1e9
print(json.dumps({"ctas": os.environ.get("CNB_EW_CTAS_PER_SM", "default"), **{k: round(v) for k, v in res.items()}}))
