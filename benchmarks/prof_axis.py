"""Profiling driver (for ncu): a few axis / scalar reductions and a fill on 32768x32768 fp32."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn  # noqa: E402

r = int(os.environ.get("ROWS", "32768"))
x = cn.empty((r, r), dtype=np.float32)
x.fill(0.5)
x[r // 3, :] = 2.0
which = os.environ.get("WHICH", "axis0")
for _ in range(3):
    if which == "axis0":
        x.sum(axis=0)
        x.argmax(axis=0)
    elif which == "axis1":
        x.sum(axis=1)
    else:
        x.sum()
cn.synchronize()
