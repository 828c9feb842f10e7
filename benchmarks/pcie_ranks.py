"""Host<->device copy bandwidth with N ranks copying AT THE SAME TIME (one rank per GPU, torchrun):
the ceiling of bench.py's e2e legs at N > 1.  Every rank binds itself to its GPU's CPUs like bench.py
does, allocates pinned buffers, and all ranks copy 2 GiB each way between barriers.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/pcie_ranks.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import cunumeric_b200 as cn  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cn.runtime.ensure_initialized()
binding = bench.pin_to_gpu_numa(cn.runtime.device)
rt, lib = cn.runtime, cn.runtime.lib
n = 1 << 31
h_in, h_out = rt.pinned_empty((n,), np.uint8), rt.pinned_empty((n,), np.uint8)
h_in[:] = 1
d = rt.allocate(n)
s1, s2 = lib.cnb_stream_create(), lib.cnb_stream_create()


def timed(fn, reps=3):
    fn()
    lib.cnb_stream_synchronize(s1)
    lib.cnb_stream_synchronize(s2)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    lib.cnb_stream_synchronize(s1)
    lib.cnb_stream_synchronize(s2)
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


t_h2d = timed(lambda: lib.cnb_memcpy_h2d(d.ptr, h_in.ctypes.data, n, s1))
t_d2h = timed(lambda: lib.cnb_memcpy_d2h(h_out.ctypes.data, d.ptr, n, s2))
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_rank": n, "binding_rank0": binding,
                      "h2d_gbs_per_rank": n / t_h2d / 1e9, "h2d_gbs_aggregate": world * n / t_h2d / 1e9,
                      "d2h_gbs_per_rank": n / t_d2h / 1e9, "d2h_gbs_aggregate": world * n / t_d2h / 1e9}))
dist.barrier()
dist.destroy_process_group()
