"""Diagnostic (not a benchmark): where does the time of a distributed stencil iteration go?"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

import cunumeric_b200 as cn  # noqa: E402
from cunumeric_b200 import _lib  # noqa: E402
from cunumeric_b200.distributed import PartitionedArray  # noqa: E402
from cunumeric_b200.workloads import stencil_init, stencil_run  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cn.runtime.init_distributed(rank, world)
lib = cn.runtime.lib
N = int(os.environ.get("N", "40000"))
ITERS = 10


def timed(label, fn, reps=3):
    fn()
    cn.synchronize()
    dist.barrier()
    _lib.check(lib.cnb_trace_start(4096))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t_host = time.perf_counter() - t0
    cn.synchronize()
    t_all = time.perf_counter() - t0
    n = lib.cnb_trace_stop()
    rec = _lib.cnb_trace_record_t()
    kms = 0.0
    for i in range(n):
        lib.cnb_trace_get(i, ctypes.byref(rec))
        kms += rec.ms
    if True:
        print(f"[rank {rank}] {label}: wall {1e3 * t_all / reps / ITERS:.2f} ms/iter, host-issue "
              f"{1e3 * t_host / reps / ITERS:.2f} ms/iter, kernels {kms / reps / ITERS:.2f} ms/iter "
              f"({n // reps // ITERS} launches/iter)", flush=True)


grid = stencil_init(N, np.float64)
timed("with halo exchange", lambda: stencil_run(grid, ITERS))
orig = PartitionedArray.exchange_halo
PartitionedArray.exchange_halo = lambda self: setattr(self.meta, "ghost_valid", True)
timed("halo exchange disabled", lambda: stencil_run(grid, ITERS))
PartitionedArray.exchange_halo = orig


def only_halo():
    for _ in range(ITERS):
        grid._thunk.meta.ghost_valid = False
        grid._thunk.exchange_halo()


timed("halo exchange only", only_halo)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler  # noqa: E402
for period in (0.02, 0.1):
    sampler = ClockSampler(cn.runtime.device, period)
    sampler.start()
    timed(f"with halo exchange + NVML sampler every {period}s", lambda: stencil_run(grid, ITERS))
    print(rank, sampler.stop(), flush=True)
dist.barrier()
dist.destroy_process_group()
