"""Stencil leg of bench.py (BASELINE.json configs[3]): examples/stencil.py fp64, N x N interior,
row-partitioned over the GPUs, halo rows exchanged over NVLink (NCCL send/recv, stream-ordered).
Strong scaling: the global grid is fixed, each rank owns (N+2)/world rows."""
from __future__ import annotations

import ctypes
import json
import time

import numpy as np


def run_stencil(args, rank, world, dist, ClockSampler, load_peaks, max_over_ranks, barrier,
                trace_summary=None) -> None:
    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib
    from cunumeric_b200.partition import RowPartition, halo_bytes
    from cunumeric_b200.workloads import (STENCIL_BYTES_PER_POINT_F64, STENCIL_TASKS_PER_ITER,
                                          stencil_init, stencil_run)

    n, iters = args.stencil_n, args.stencil_iters
    cn.runtime.ensure_initialized()
    if world > 1:
        cn.runtime.init_distributed(rank, world)
    lib = cn.runtime.lib
    peak_gbs, peak_src = load_peaks()
    from cunumeric_b200 import fusion

    fused_on = getattr(args, "fusion", "on") == "on" and fusion.enabled()
    fusion.set_mode("1" if fused_on else "0")
    grid = stencil_init(n, np.float64)
    sampler = ClockSampler(cn.runtime.device)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        stencil_run(grid, iters)
        cn.flush()
    cn.synchronize()

    events = [lib.cnb_event_create() for _ in range(args.steps + 1)]
    _lib.check(lib.cnb_trace_start(args.steps * iters * (STENCIL_TASKS_PER_ITER + 2)))
    barrier(dist)
    cn.synchronize()
    sampler.mark_begin()
    launches0 = cn.runtime.launch_count()
    lib.cnb_event_record(events[0], cn.runtime.stream)
    for i in range(args.steps):
        stencil_run(grid, iters)
        cn.flush()  # nothing stays pending: the step's last COPY is issued inside the timed region
        lib.cnb_event_record(events[i + 1], cn.runtime.stream)
    cn.synchronize()
    sampler.mark_end()
    barrier(dist)
    launches = cn.runtime.launch_count() - launches0
    n_rec = lib.cnb_trace_stop()
    ms = ctypes.c_float()
    _lib.check(lib.cnb_event_elapsed_ms(events[0], events[-1], ctypes.byref(ms)))
    step_ms = []
    for i in range(args.steps):
        one = ctypes.c_float()
        _lib.check(lib.cnb_event_elapsed_ms(events[i], events[i + 1], ctypes.byref(one)))
        step_ms.append(round(one.value, 3))
    clocks = sampler.stop()
    elapsed = max_over_ranks(dist, ms.value * 1e-3)
    points = float(n) * n * iters * args.steps
    value = points / elapsed
    # fused: one kernel reads the five shifted views (the re-reads hit L1/L2: 8 B of DRAM reads per
    # point) and writes `average` and `work` (both stay observable in the program text), then the COPY
    # back into `center` reads and writes 8 B each: 8 + 16 + 16 = 40 B per point
    bytes_per_point = 40 if fused_on else STENCIL_BYTES_PER_POINT_F64
    gbs_per_gpu = value * bytes_per_point / 1e9 / world
    kernel_roofline = None
    if trace_summary is not None:
        kernel_roofline, whole = trace_summary(cn, n_rec, peak_gbs)
        if kernel_roofline is not None:
            kernel_roofline["peak_source"] = peak_src
            kernel_roofline.update(whole)
    part = RowPartition.even(n + 2, world)
    sent, recv = halo_bytes(part, 1, (n + 2) * 8, min(1, world - 1))
    if rank == 0:
        print(json.dumps({
            "metric": "stencil_points_per_second", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"stencil fp64 N={n} (examples/stencil.py), {iters} Jacobi "
                                   f"iterations per step, {STENCIL_TASKS_PER_ITER} tasks per "
                                   "iteration, " + ("4 ADD + MULTIPLY run as one fused kernel, then "
                                                    "the COPY" if fused_on else "issued op-by-op") +
                                   ", rows partitioned over the GPUs",
                       "execution": "fused" if fused_on else "op-by-op",
                       "N": n, "iters_per_step": iters,
                       "algorithmic_bytes_per_point": bytes_per_point,
                       "halo_bytes_per_iter_per_gpu": {"sent": sent, "received": recv},
                       "l2_policy": f"grid {(n + 2) ** 2 * 8 / 1e9:.1f} GB and temporaries exceed "
                                    "the 126 MB L2" if n >= 8000 else "working set fits L2"},
            "gpu_launches": int(launches), "clocks": clocks, "step_ms_rank0": step_ms,
            "roofline": kernel_roofline or {"bound": "hbm", "achieved": gbs_per_gpu,
                                            "peak": peak_gbs, "unit": "GB/s",
                                            "frac": gbs_per_gpu / peak_gbs, "traffic": None},
            "whole_iteration": {"algorithmic_gbs_per_gpu": gbs_per_gpu,
                                "frac_of_hbm_peak": gbs_per_gpu / peak_gbs,
                                "note": "4 ADD on pitched views + scalar MULTIPLY + COPY: 128 B per "
                                        "point op-by-op, 40 B per point fused"},
            "cpu_baseline": None, "e2e": None,
        }))
