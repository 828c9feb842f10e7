"""Profiling driver (for ncu): a few Jacobi iterations of examples/stencil.py at N=40000 fp64."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunumeric_b200 as cn  # noqa: E402
from cunumeric_b200.workloads import stencil_init, stencil_run  # noqa: E402

grid = stencil_init(int(os.environ.get("N", "40000")), np.float64)
stencil_run(grid, 3)
cn.synchronize()
