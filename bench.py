#!/usr/bin/env python
"""bench.py — the headline benchmark of the hot path.

    python bench.py --gpus N --steps K --warmup W [--workload stencil|black_scholes] [--impl reference]

Own arm (default workload `stencil`, BASELINE.json configs[3] — the config the ">= 7x from 1 to 8
B200" target is quoted on; it fits one GPU, so N = 1 runs the very same job):
  stencil        examples/stencil.py fp64, N x N = 40000^2 interior points, ROW-PARTITIONED over the
                 N GPUs (one process per GPU, NCCL clique created by runtime.init_distributed), halo
                 rows exchanged over NVLink inside the timed region.  STRONG scaling: the global
                 grid is fixed.  One step = ITERS Jacobi iterations (6 tasks each, issued through
                 the cunumeric NumPy API exactly as the reference issues them).
                 metric = interior points updated per second, whole job.
  The same JSON line carries, as extra keys,
    `c1`             configs[0]: N = 1000 x 100 iterations on one B200 and on ONE host core (N = 1);
    `black_scholes`  configs[1]: examples/black_scholes.py fp32, 1e8 options per GPU (no exchange
                     step: N independent replicas, weak), fused and op-by-op;
    `sweeps`         configs[2] and [4]: the reduction sweep (32768^2 fp32, sum/max/argmax x axis
                     0 / axis 1 / full) and the elementwise dtype sweep (add / multiply / where /
                     astype over 2^30 elements x 6 dtypes), arrays row-partitioned over the N GPUs.
  black_scholes  makes configs[1] the primary line instead (round-1 behaviour).
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU arithmetic
(oracle/_ref: its functor headers compiled with g++, OpenMP over all host cores) on a bounded
sample of the same workload, honouring --steps / --warmup."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BS_METRIC, BS_UNIT = "black_scholes_elements_per_second", "options/s"
ST_METRIC, ST_UNIT = "stencil_points_per_second", "points/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="own", choices=["own", "reference"])
    p.add_argument("--workload", default="stencil", choices=["stencil", "black_scholes"])
    p.add_argument("--n", type=int, default=100_000_000, help="options per GPU (black_scholes)")
    p.add_argument("--stencil-n", type=int, default=40000)
    p.add_argument("--stencil-iters", type=int, default=100, help="Jacobi iterations per step")
    p.add_argument("--cpu-sample", type=int, default=10_000_000,
                   help="options in the bounded CPU sample (black_scholes)")
    p.add_argument("--cpu-stencil-n", type=int, default=4000,
                   help="grid edge of the bounded CPU sample (stencil)")
    p.add_argument("--cpu-stencil-iters", type=int, default=5,
                   help="Jacobi iterations per step of the bounded CPU sample (stencil)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-chunks", type=int, default=8,
                   help="chunks per step of the Black-Scholes e2e leg (upload / compute / download)")
    p.add_argument("--fusion", choices=["on", "off"], default="on",
                   help="on: elementwise chains run as fused kernels (default product path); "
                        "off: one kernel per task")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extras", action="store_true",
                   help="primary workload only (no black_scholes / sweeps keys)")
    p.add_argument("--sweep-log2n", type=int, default=30)
    p.add_argument("--sweep-rows", type=int, default=32768)
    p.add_argument("--sweep-reps", type=int, default=10)
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel: str):
    """DRAM bytes per launch of a kernel from the committed ncu --set full captures
    (profiles/r02_traffic.json, else r01): launch-weighted mean over the kernel's variants.  None
    when no capture covers this kernel."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                data = json.load(f)
            variants = data["kernels"][kernel]
            n = sum(v["launches_per_step"] for v in variants.values())
            total = sum((v["dram_read_bytes"] + v["dram_write_bytes"]) * v["launches_per_step"]
                        for v in variants.values())
            return total / n, f"profiles/{name} (ncu --set full, per launch)"
        except Exception:
            continue
    return None, None


class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md recipe).
    Uses NVML in a background thread (same counters as the recipe's nvidia-smi line, without
    spawning a process that contends for the driver lock while kernels are being launched)."""

    def __init__(self, device: int, period_s: float = 0.01) -> None:
        self.device = device
        self.period = period_s
        self.samples = []
        self._stop = None
        self._thread = None
        self._err = None
        self._nv = None
        self._handle = None
        self._max = None
        # NVML initialisation takes ~100 ms and briefly stalls the driver: do it here, outside the
        # timed region; start() only spawns the polling thread.  Only rank 0 samples (its own GPU):
        # NVML queries take a driver-wide lock, and eight ranks polling at once delay each other's
        # kernel launches.
        if int(os.environ.get("RANK", "0")) != 0:
            self._err = "clocks are sampled on rank 0 only"
            return
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._handle = nv.nvmlDeviceGetHandleByIndex(self._uuid_index(nv))
            self._max = nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM)
            self._nv = nv
        except Exception as exc:
            self._err = f"nvml unavailable: {exc}"

    def _uuid_index(self, nv):
        # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ent = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ent):
                e = ent[self.device]
                if e.isdigit():
                    return int(e)
                for i in range(nv.nvmlDeviceGetCount()):
                    h = nv.nvmlDeviceGetHandleByIndex(i)
                    if nv.nvmlDeviceGetUUID(h).startswith(e):
                        return i
        return self.device

    def start(self) -> None:
        import threading

        if self._nv is None:
            return
        nv, h = self._nv, self._handle
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.samples.append((sm, reasons, power, time.perf_counter()))
                except Exception as exc:  # keep sampling errors out of the measurement
                    self._err = str(exc)
                self._stop.wait(self.period)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    # The polling thread is started BEFORE the warm-up (the first NVML queries of a process are slow
    # and contend with kernel launches); only samples taken inside [mark_begin, mark_end] count.
    def mark_begin(self) -> None:
        self._t0 = time.perf_counter()

    def mark_end(self) -> None:
        self._t1 = time.perf_counter()

    def stop(self) -> dict:
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self._err or "not started"]}
        self._stop.set()
        self._thread.join(timeout=2)
        import pynvml as nv

        t0, t1 = getattr(self, "_t0", None), getattr(self, "_t1", None)
        if t0 is not None and t1 is not None:
            inside = [s for s in self.samples if t0 <= s[3] <= t1]
            self.samples = inside or self.samples[-1:]
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self._max, "reasons": [self._err or "no samples"]}
        names = {
            "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
            "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
            "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
            "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
            "hw_power_brake": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown,
        }
        reasons = sorted(n for n, bit in names.items() if any(s[1] & bit for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self._max, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(s[2] for s in self.samples)}


def dist_setup(world: int):
    """Plumbing only: torch.distributed (gloo) for barrier + max-over-ranks of the device timings,
    and to hand rank 0's NCCL unique id to the others.  The data path uses the library's own NCCL
    clique (runtime.init_distributed -> cnb_comm_init)."""
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return rank, None
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    return rank, dist


def barrier(dist) -> None:
    if dist is not None:
        dist.barrier()


def max_over_ranks(dist, value: float) -> float:
    if dist is None:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pin_to_gpu_numa(device: int) -> dict:
    """Bind this rank (and the pinned host buffers it allocates from now on) to the CPUs of the NUMA
    node its GPU hangs off, if the platform says which one that is; else spread the ranks over the
    visible CPUs so that eight ranks do not all fault their staging buffers on node 0."""
    info = {"bound": False}
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(device)
        cpus = None
        try:
            words = (os.cpu_count() + 63) // 64
            mask = nv.nvmlDeviceGetCpuAffinity(h, words)
            cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        except Exception:
            cpus = None
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        allowed = sorted(os.sched_getaffinity(0))
        if cpus:
            cpus = [c for c in cpus if c in allowed]
        if not cpus:
            cpus = allowed
        # inside the GPU's CPU set, give every rank its own slice (no two ranks on one core)
        per = max(1, len(cpus) // max(1, world))
        mine = cpus[(rank * per) % len(cpus): (rank * per) % len(cpus) + per] or cpus
        os.sched_setaffinity(0, mine)
        info = {"bound": True, "cpus": f"{mine[0]}-{mine[-1]}", "ncpus": len(mine)}
    except Exception as exc:
        info["error"] = str(exc)[:120]
    return info


# ------------------------------------------------------------------------------------------------
# CPU legs (the reference's functors, oracle/_ref) — reported baselines, never the target
# ------------------------------------------------------------------------------------------------
def _cpu_black_scholes_step(n_sample: int):
    from oracle import refnp
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs

    S, X, T = black_scholes_inputs(n_sample, np.float32, seed=0)
    S, X, T = refnp.array(S), refnp.array(X), refnp.array(T)
    return lambda: black_scholes(S, X, T, 0.02, 0.3, xp=refnp)


def _cpu_stencil_step(n_sample: int, iters: int):
    from oracle import refnp
    from cunumeric_b200.workloads import stencil_init, stencil_run

    grid = stencil_init(n_sample, np.float64, xp=refnp)
    return lambda: stencil_run(grid, iters)


def cpu_baseline(workload: str, args, reps: int = 2) -> dict:
    from oracle import ref, refnp

    cores = os.cpu_count() or 1
    refnp.set_threads(cores)
    if workload == "black_scholes":
        step, units, unit = _cpu_black_scholes_step(args.cpu_sample), args.cpu_sample, BS_UNIT
        sample = (f"{args.cpu_sample} options, best of {reps}; reference functors (oracle/_ref, "
                  "g++ -O2 -fopenmp, schedule(static)) op-by-op, 63 tasks")
    else:
        ns, it = args.cpu_stencil_n, args.cpu_stencil_iters
        step, units, unit = _cpu_stencil_step(ns, it), float(ns) * ns * it, ST_UNIT
        sample = (f"N={ns} grid, {it} Jacobi iterations, best of {reps}; reference functors "
                  "(oracle/_ref, g++ -O2 -fopenmp, schedule(static)) op-by-op, 6 tasks per iteration")
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
    return {"value": units / best, "unit": unit, "cores": cores, "kind": "reference",
            "sample": sample, "available": ref.available()}


def run_reference(args) -> None:
    """`--impl reference`: the reference's CPU implementation of the path on the host cores, on a
    bounded sample of the own arm's workload, `--steps` steps after `--warmup` warm-ups."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refnp

    cores = os.cpu_count() or 1
    refnp.set_threads(cores)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload == "black_scholes":
        n = args.cpu_sample
        step, units = _cpu_black_scholes_step(n), float(n)
        metric, unit, dtype, scaling = BS_METRIC, BS_UNIT, "f32", "weak"
        config = {"workload": WORKLOAD_BS + f"the reference's CPU functors issued op-by-op (OpenMP, "
                              f"all host cores); each step is a bounded sample of {n} options",
                  "execution": "reference-cpu", "options_per_gpu": args.n,
                  "options_per_step_sampled": n, "tasks_per_step": 63}
        sample = f"{n} options per step (bounded sample of the 1e8-option workload)"
    else:
        ns, it = args.cpu_stencil_n, args.cpu_stencil_iters
        step, units = _cpu_stencil_step(ns, it), float(ns) * ns * it
        metric, unit, dtype, scaling = ST_METRIC, ST_UNIT, "f64", "strong"
        config = {"workload": workload_stencil(args.stencil_n, args.stencil_iters) +
                              "; the reference's CPU functors issued op-by-op (OpenMP, all host "
                              f"cores); each step is a bounded sample: N={ns}, {it} iterations",
                  "execution": "reference-cpu", "N": args.stencil_n,
                  "iters_per_step": args.stencil_iters, "N_sampled": ns, "iters_per_step_sampled": it}
        sample = f"N={ns} grid, {it} Jacobi iterations per step (bounded sample of N={args.stencil_n})"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = units * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
        "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference",
                         "sample": sample + f", {steps} steps; reference functors compiled from "
                                            "/root/reference/src (oracle/_ref), OpenMP"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# the workloads both arms (own / --impl reference) are measured on
WORKLOAD_BS = ("black_scholes fp32 1e8 options per GPU (examples/black_scholes.py, BASELINE.json "
               "configs[1]), 63 elementwise tasks per step through the NumPy API; ")


def workload_stencil(n: int, iters: int) -> str:
    return (f"stencil fp64 N={n} (examples/stencil.py, BASELINE.json configs[3]), {iters} Jacobi "
            "iterations per step, 6 tasks per iteration through the NumPy API, grid rows "
            "partitioned over the GPUs with a halo exchange per iteration")


# ------------------------------------------------------------------------------------------------
def trace_summary(cn, n_records: int, peak_gbs: float):
    """Group the live per-launch records by kernel (task, op, dtype) and pick the dominant one."""
    from cunumeric_b200 import _lib
    from cunumeric_b200.config import BinaryOpCode, UnaryOpCode

    lib = cn.runtime.lib
    groups = {}
    rec = _lib.cnb_trace_record_t()
    total_ms = 0.0
    for i in range(n_records):
        _lib.check(lib.cnb_trace_get(i, ctypes.byref(rec)))
        key = (rec.task, rec.op, rec.dtype)
        g = groups.setdefault(key, {"launches": 0, "ms": 0.0, "bytes": 0})
        g["launches"] += 1
        g["ms"] += rec.ms
        g["bytes"] += rec.bytes
        total_ms += rec.ms
    if not groups:
        return None, {}
    names = {5: "BINARY_OP", 43: "UNARY_OP", 49: "WHERE", 11: "CONVERT", 33: "SCALAR_UNARY_RED",
             44: "UNARY_RED", 19: "FILL"}

    def _name(key, long=False):
        if key[0] == 1000:
            if key[1] == 0:
                return "gap_copy_kernel (complement of a renamed window)"
            return (f"fused_kernel<{key[1]} elementwise tasks>" if long else f"FUSED:{key[1]} tasks")
        nm = str(key[1])
        if key[0] == 5:
            nm = BinaryOpCode(key[1]).name
        elif key[0] == 43:
            nm = UnaryOpCode(key[1]).name
        base = f"{names.get(key[0], key[0])}:{nm}:dtype{key[2]}"
        return f"ew_kernel<{base}>" if long else base

    top_key, top = max(groups.items(), key=lambda kv: kv[1]["ms"])
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": None, "kernel": _name(top_key, long=True),
        "launches": top["launches"], "avg_launch_ms": top["ms"] / top["launches"],
        "algorithmic_bytes_per_launch": top["bytes"] / top["launches"],
        "share_of_step": top["ms"] / total_ms,
    }
    roofline["per_kernel"] = {
        _name(k): {"launches": g["launches"], "ms": round(g["ms"], 3),
                   "gbs": round(g["bytes"] / (g["ms"] * 1e-3) / 1e9, 1)}
        for k, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
    all_bytes = sum(g["bytes"] for g in groups.values())
    whole = {"kernel_ms_total": total_ms,
             "algorithmic_gbs_all_kernels": all_bytes / (total_ms * 1e-3) / 1e9,
             "frac_all_kernels": all_bytes / (total_ms * 1e-3) / 1e9 / peak_gbs}
    return roofline, whole


class DeviceTimer:
    """CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, cn, dist) -> None:
        self.cn, self.dist, self.lib = cn, dist, cn.runtime.lib
        self.e0, self.e1 = self.lib.cnb_event_create(), self.lib.cnb_event_create()

    def begin(self) -> None:
        barrier(self.dist)
        self.cn.synchronize()
        self.lib.cnb_event_record(self.e0, self.cn.runtime.stream)

    def end(self) -> float:
        from cunumeric_b200 import _lib

        self.lib.cnb_event_record(self.e1, self.cn.runtime.stream)
        self.cn.synchronize()
        barrier(self.dist)
        ms = ctypes.c_float()
        _lib.check(self.lib.cnb_event_elapsed_ms(self.e0, self.e1, ctypes.byref(ms)))
        return max_over_ranks(self.dist, ms.value * 1e-3)


# ------------------------------------------------------------------------------------------------
# stencil (configs[3]) — the primary line
# ------------------------------------------------------------------------------------------------
def stencil_leg(args, rank: int, world: int, dist, sampler) -> dict:
    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib, fusion
    from cunumeric_b200.partition import RowPartition, halo_bytes
    from cunumeric_b200.workloads import (STENCIL_BYTES_PER_POINT_F64, STENCIL_TASKS_PER_ITER,
                                          stencil_init, stencil_run)

    n, iters, steps, warmup = args.stencil_n, args.stencil_iters, args.steps, max(args.warmup, 3)
    lib = cn.runtime.lib
    peak_gbs, peak_src = load_peaks()
    fused_on = args.fusion == "on" and fusion.enabled()
    fusion.set_mode("1" if fused_on else "0")
    # parity guard on the very path that is about to be timed (partitioned, fused, renamed, TMA):
    # a small even-N grid, a few iterations, bit-exact against NumPy on every rank
    ns = 1022
    small = stencil_init(ns, np.float64)
    stencil_run(small, 7)
    ref_small = stencil_init(ns, np.float64, xp=np)
    stencil_run(ref_small, 7)
    selfcheck = bool(np.array_equal(np.asarray(small), ref_small))
    del small
    grid = stencil_init(n, np.float64)
    for _ in range(warmup):
        stencil_run(grid, iters)
        cn.flush()
    cn.synchronize()
    # host cost of issuing one iteration with an EMPTY launch queue (20 iterations are ~60 launches, far
    # below the depth at which the driver blocks the issuing thread); the figure taken inside the timed
    # region below includes the time the host spends blocked behind a full queue
    t_free = time.perf_counter()
    stencil_run(grid, 20)
    cn.flush()
    t_free = (time.perf_counter() - t_free) / 20
    cn.synchronize()

    _lib.check(lib.cnb_trace_start(steps * iters * (STENCIL_TASKS_PER_ITER + 3) + 16))
    timer = DeviceTimer(cn, dist)
    stats0 = dict(fusion.stats)
    timer.begin()
    sampler.mark_begin()
    launches0 = cn.runtime.launch_count()
    t_issue = time.perf_counter()
    for _ in range(steps):
        stencil_run(grid, iters)
        cn.flush()  # nothing stays pending: the step's last iteration is issued inside the timed region
    t_issue = time.perf_counter() - t_issue   # host time to ISSUE the work (launches are asynchronous)
    elapsed = timer.end()
    sampler.mark_end()
    launches = cn.runtime.launch_count() - launches0
    n_rec = lib.cnb_trace_stop()
    fstats = {k: fusion.stats[k] - stats0[k] for k in stats0}
    points = float(n) * n * iters * steps
    value = points / elapsed
    roofline, whole = trace_summary(cn, n_rec, peak_gbs)
    # bytes one iteration has to move per interior point (fp64): op-by-op 128 (4 ADD x 24 + scalar
    # MULTIPLY 16 + COPY 16).  Fused: the chain [4 ADD, MULTIPLY, COPY] is ONE kernel
    # that reads the grid once (the five shifted views are one buffer) and writes the new interior
    # into the renamed block: 16 B per point, the lower bound SURVEY §8d states.  The temporaries
    # `average` / `work` of an iteration are dead by the time its kernel is launched.
    part = RowPartition.even(n + 2, world)
    sent, recv = halo_bytes(part, 1, (n + 2) * 8, min(1, world - 1))
    if roofline is not None:
        roofline["peak_source"] = peak_src
        roofline.update(whole)
        bytes_per_point = (roofline["algorithmic_gbs_all_kernels"] * 1e9 *
                           roofline["kernel_ms_total"] * 1e-3) / (points / world)
        roofline["traffic"], roofline["traffic_source"] = load_traffic(
            "fused_stencil" if fused_on else roofline["kernel"])
    else:
        bytes_per_point = 16 if fused_on else STENCIL_BYTES_PER_POINT_F64
    gbs_per_gpu = value * bytes_per_point / 1e9 / world
    out = {
        "metric": ST_METRIC, "value": value, "unit": ST_UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * elapsed / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_stencil(n, iters) +
                               ("; each iteration's 4 ADD + MULTIPLY + COPY run as ONE fused kernel "
                                "(write-after-read renaming of the grid buffer, bit-identical to "
                                "op-by-op)" if fused_on else "; issued op-by-op, one kernel per task"),
                   "execution": "fused" if fused_on else "op-by-op",
                   "N": n, "iters_per_step": iters, "ms_per_iteration": 1e3 * elapsed / steps / iters,
                   "host_issue_ms_per_iteration": 1e3 * t_issue / steps / iters,
                   "host_issue_ms_per_iteration_empty_queue": 1e3 * t_free,
                   "host_issue_note": "the first figure is taken inside the timed region and includes "
                                      "the time the issuing thread is blocked behind a full launch queue "
                                      "(it tracks the GPU time); the second is the host's own cost",
                   "algorithmic_bytes_per_point": round(bytes_per_point, 3),
                   "collective": ("ncclSend/ncclRecv of one ghost row per neighbour per iteration "
                                  "(grouped, stream-ordered)" if world > 1 else "none (1 GPU)"),
                   "halo_bytes_per_iter_per_gpu": {"sent": sent, "received": recv},
                   "fusion_stats": fstats,
                   "selfcheck": {"N": ns, "iters": 7, "bit_exact_vs_numpy": selfcheck},
                   "l2_policy": f"grid {(n + 2) ** 2 * 8 / 1e9:.1f} GB (/{world} per GPU) exceeds the "
                                "126 MB L2; no flush needed" if (n + 2) ** 2 * 8 / world > 2.5e8 else
                                "per-GPU block near L2 size: numbers include L2 hits"},
        "gpu_launches": int(launches),
        "roofline": roofline or {"bound": "hbm", "achieved": gbs_per_gpu, "peak": peak_gbs,
                                 "unit": "GB/s", "frac": gbs_per_gpu / peak_gbs, "traffic": None},
        "whole_iteration": {"algorithmic_gbs_per_gpu": gbs_per_gpu,
                            "frac_of_hbm_peak": gbs_per_gpu / peak_gbs,
                            "note": "wall time of the step (kernels + halo exchange + launch gaps) "
                                    "against the bytes the kernels have to move"},
    }
    del grid
    if not selfcheck:
        raise RuntimeError("stencil self-check failed: the partitioned / fused path disagrees with NumPy")
    return out


def stencil_e2e(args, rank: int, world: int, dist) -> dict:
    """The same step through the public API from HOST buffers: every rank uploads its block of rows
    from pinned host memory, runs ITERS iterations and reads its block back."""
    import cunumeric_b200 as cn
    from cunumeric_b200.partition import RowPartition
    from cunumeric_b200.workloads import stencil_run

    n, iters = args.stencil_n, args.stencil_iters
    part = RowPartition.even(n + 2, world)
    lo, hi = part.bounds(rank)
    host = cn.pinned_empty((hi - lo, n + 2), np.float64)
    host[...] = 0.0
    host[:, 0] = host[:, -1] = -273.15
    if hi == n + 2:
        host[-1, :] = -273.15
    if lo == 0:
        host[0, :] = 40.0
    steps = max(2, min(args.steps, 3))

    def step():
        grid = cn.from_host_rows(host, n + 2)
        stencil_run(grid, iters)
        grid.to_host_rows(host)

    step()
    barrier(dist)
    cn.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    cn.synchronize()
    dt = max_over_ranks(dist, time.perf_counter() - t0)
    nbytes = (n + 2) * (n + 2) * 8
    return {"value": float(n) * n * iters * steps / dt, "unit": ST_UNIT,
            "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": steps,
            "ms_per_step": 1e3 * dt / steps,
            "api": "cunumeric_b200.from_host_rows(pinned block of rows) -> stencil_run(grid, "
                   f"{iters}) -> grid.to_host_rows(pinned block): every rank moves its own "
                   f"{(hi - lo) * (n + 2) * 8 / 1e9:.2f} GB each way per step, copies serialised with "
                   "the iterations (the stencil needs the whole block before it can start; "
                   "software-pipelining neighbouring steps over the copy streams was measured and is "
                   "slower — 0.96 vs 0.85 s per step at N=1: DMA traffic starves next to a kernel that "
                   "saturates HBM)"}


def c1_leg(args) -> dict:
    """BASELINE configs[0]: examples/stencil.py fp64 N = 1000, 100 iterations (+5 warm-up, as
    stencil.py:68-73 does) — the reference's own CPU-runnable case (`legate --cpus 1`), timed on both
    arms: the product on one B200 (the working set fits L2, the run is bound by the host issuing
    6 NumPy calls per iteration) and the reference's functors on ONE host core."""
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import stencil_init, stencil_run

    n, iters = 1000, 100
    grid = stencil_init(n, np.float64)
    stencil_run(grid, 5)
    cn.synchronize()
    t0 = time.perf_counter()
    stencil_run(grid, iters)
    cn.synchronize()
    own = time.perf_counter() - t0
    out = {"config": "examples/stencil.py fp64 N=1000, 100 iterations (BASELINE configs[0])",
           "own": {"seconds": own, "points_per_s": float(n) * n * iters / own, "timing": "wall clock "
                   "around stencil_run + synchronize (host-bound: 0.1 ms of GPU work per iteration)"}}
    try:
        from oracle import refnp

        refnp.set_threads(1)
        g = stencil_init(n, np.float64, xp=refnp)
        stencil_run(g, 5)
        t0 = time.perf_counter()
        stencil_run(g, iters)
        ref_s = time.perf_counter() - t0
        out["reference_cpu"] = {"seconds": ref_s, "points_per_s": float(n) * n * iters / ref_s, "cores": 1,
                                "kind": "reference", "note": "the reference's functors (oracle/_ref) "
                                "issued op-by-op on one core = `legate --cpus 1`"}
    except Exception as exc:
        out["reference_cpu"] = {"unavailable": str(exc)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
# Black-Scholes (configs[1])
# ------------------------------------------------------------------------------------------------
def black_scholes_leg(args, rank: int, world: int, dist, sampler, primary: bool) -> dict:
    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib, fusion
    from cunumeric_b200.workloads import (BLACK_SCHOLES_BYTES_PER_OPTION_F32, BLACK_SCHOLES_TASKS,
                                          black_scholes, black_scholes_inputs)

    n = args.n
    lib = cn.runtime.lib
    peak_gbs, peak_src = load_peaks()
    R, V = 0.02, 0.3
    steps = args.steps
    # synthetic inputs, generated on the host once, staged in PINNED memory (for the e2e leg) and
    # made resident in HBM before the timed region.  Replicas: every rank works on its own 1e8
    # options (the path has no exchange step), so the arrays are plain per-rank device arrays.
    Sh, Xh, Th = (cn.pinned_empty(n, np.float32) for _ in range(3))
    for dst, src in zip((Sh, Xh, Th), black_scholes_inputs(n, np.float32, seed=rank)):
        dst[...] = src
    S, X, T = cn.array(Sh), cn.array(Xh), cn.array(Th)
    cn.synchronize()

    def timed(fused: bool, nsteps: int, sample_clocks: bool):
        """W warm-up steps, then `nsteps` timed steps.  Every step ends with cn.flush(), so the
        step's results (call, put) are materialised in HBM inside the timed region."""
        fusion.set_mode("1" if fused else "0")
        for _ in range(max(args.warmup, 3)):
            call, put = black_scholes(S, X, T, R, V)
            cn.flush()
        cn.synchronize()
        _lib.check(lib.cnb_trace_start(nsteps * (BLACK_SCHOLES_TASKS + 8)))
        timer = DeviceTimer(cn, dist)
        timer.begin()
        if sample_clocks:
            sampler.mark_begin()
        launches0 = cn.runtime.launch_count()
        for _ in range(nsteps):
            call, put = black_scholes(S, X, T, R, V)
            cn.flush()
        elapsed = timer.end()
        if sample_clocks:
            sampler.mark_end()
        launches = cn.runtime.launch_count() - launches0
        n_rec = lib.cnb_trace_stop()
        roofline, whole = trace_summary(cn, n_rec, peak_gbs)
        if roofline is not None:
            roofline["peak_source"] = peak_src
            roofline.update(whole)
            roofline["traffic"], roofline["traffic_source"] = load_traffic(roofline["kernel"])
        return elapsed, launches, roofline

    fused_on = args.fusion == "on" and fusion.enabled()
    obo_steps = steps if not fused_on else max(3, min(steps, 5))
    obo_elapsed, obo_launches, obo_roofline = timed(False, obo_steps, primary and not fused_on)
    op_by_op = {"value": n * world * obo_steps / obo_elapsed, "unit": BS_UNIT, "steps": obo_steps,
                "ms_per_step": 1e3 * obo_elapsed / obo_steps, "gpu_launches": int(obo_launches),
                "algorithmic_bytes_per_option": BLACK_SCHOLES_BYTES_PER_OPTION_F32,
                "whole_step_algorithmic_gbs":
                    BLACK_SCHOLES_BYTES_PER_OPTION_F32 * n * obo_steps / obo_elapsed / 1e9,
                "roofline": obo_roofline}
    if fused_on:
        elapsed, launches, roofline = timed(True, steps, primary)
        bytes_per_option = 20  # 3 fp32 inputs + 2 fp32 outputs; 61 intermediates stay in registers
        if roofline is not None:
            roofline["limiter"] = ("instruction issue, not HBM: ~170 SASS instructions per option "
                                   "(4 IEEE divisions, 3 exp, log, sqrt, no FMA contraction — the "
                                   "bit-parity contract with op-by-op execution); DRAM traffic = "
                                   "algorithmic bytes (profiles/r02_fused_black_scholes.md)")
    else:
        elapsed, launches, roofline = obo_elapsed, obo_launches, obo_roofline
        bytes_per_option = BLACK_SCHOLES_BYTES_PER_OPTION_F32
    fusion.set_mode("1" if fused_on else "0")
    value = n * world * steps / elapsed

    e2e = None
    if not args.no_e2e:
        call_h, put_h = cn.pinned_empty(n, np.float32), cn.pinned_empty(n, np.float32)
        e2e_steps = max(2, min(steps, 5))
        chunk = max(1, n // args.e2e_chunks)

        def e2e_step():
            # host buffers in, host buffers out: upload of chunk i+1, kernels of chunk i and
            # download of chunk i-1 overlap on separate streams (cunumeric_b200.map_chunks)
            cn.map_chunks(lambda s, x, t: black_scholes(s, x, t, R, V), (Sh, Xh, Th),
                          (call_h, put_h), chunk)

        e2e_step()
        barrier(dist)
        cn.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        cn.synchronize()
        dt = max_over_ranks(dist, time.perf_counter() - t0)
        e2e = {"value": n * world * e2e_steps / dt, "unit": BS_UNIT,
               "h2d_bytes_per_step": 3 * 4 * n * world, "d2h_bytes_per_step": 2 * 4 * n * world,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "cunumeric_b200.map_chunks(black_scholes, pinned inputs, pinned outputs, "
                      f"chunk={chunk}): from_host(blocking=False) -> black_scholes() -> "
                      "to_host(blocking=False), 3 streams"}
        del call_h, put_h
    out = {
        "metric": BS_METRIC, "value": value, "unit": BS_UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_BS +
                               ("the thunk layer captures the chain and runs it as ONE fused kernel "
                                "(bit-identical to op-by-op; `op_by_op` holds the one-kernel-per-task "
                                "measurement)" if fused_on else "issued op-by-op, one kernel per task"),
                   "execution": "fused" if fused_on else "op-by-op",
                   "options_per_gpu": n, "tasks_per_step": BLACK_SCHOLES_TASKS,
                   "algorithmic_bytes_per_option": bytes_per_option,
                   "collective": "none: the path has no exchange step, every GPU runs its own "
                                 "1e8-option replica",
                   "l2_policy": "inputs, outputs (and op-by-op temporaries) are 400 MB each, "
                                "larger than the 126 MB L2; no flush needed",
                   "whole_step_algorithmic_gbs": bytes_per_option * n * steps / elapsed / 1e9},
        "gpu_launches": int(launches), "roofline": roofline, "op_by_op": op_by_op, "e2e": e2e,
    }
    del S, X, T, Sh, Xh, Th
    return out


# ------------------------------------------------------------------------------------------------
# configs[2] and [4]: reduction sweep and elementwise dtype sweep, row-partitioned over the GPUs
# ------------------------------------------------------------------------------------------------
def sweeps_leg(args, rank: int, world: int, dist) -> dict:
    import cunumeric_b200 as cn

    peak, _ = load_peaks()
    timer = DeviceTimer(cn, dist)
    reps = args.sweep_reps
    cases = []

    def run(fn):
        def once():
            r = fn()
            cn.flush()
            return r

        for _ in range(3):
            once()
        best = None
        for _ in range(2 if world > 1 else 1):   # collectives jitter: the better of two passes at N > 1
            timer.begin()
            for _ in range(reps):
                once()
            sec = timer.end() / reps
            best = sec if best is None else min(best, sec)
        return best

    def report(name, elems, bytes_per_elem, sec, **extra):
        gbs = elems * bytes_per_elem / sec / 1e9
        rec = {"case": name, "ms": round(sec * 1e3, 4), "elements_per_s": elems / sec,
               "gbs_per_gpu": round(gbs / world, 1), "frac": round(gbs / world / peak, 4)}
        rec.update(extra)
        cases.append(rec)

    def filled(shape, dtype, value):
        a = cn.empty(shape, dtype=dtype)
        a.fill(value)
        return a

    # ---- C3
    r = args.sweep_rows
    x = filled((r, r), np.float32, 0.5)
    x[r // 3, :] = 2.0
    x[:, r // 5] = 3.0
    n = r * r
    for opname, fn in (("sum", lambda ax: x.sum(axis=ax)), ("max", lambda ax: x.max(axis=ax)),
                       ("argmax", lambda ax: x.argmax(axis=ax))):
        for ax, label in ((0, "axis0"), (1, "axis1"), (None, "full")):
            sec = run(lambda: fn(ax))
            coll = "none"
            if world > 1 and ax != 1:
                what = f"the {r * 4 // 1024} KiB partial" if ax == 0 else "one partial per GPU"
                if opname == "sum":
                    coll = f"ncclAllReduce of {what}"
                elif opname == "max":   # NaN-exact: ncclMax does not fold like `if (b > a) a = b`
                    coll = f"ncclAllGather of {what} + the library's MAX kernel over the ranks"
                else:
                    coll = (f"ncclAllGather of {what.replace('partial', '{index, value} partial')} + one "
                            "fold kernel (cnb_argval_fold, lowest index wins ties)")
            report(f"C3 {opname} {label} {r}x{r} f32", n, 4, sec, collective=coll)
    del x
    # ---- C5
    n = 1 << args.sweep_log2n
    dtypes = [("f16", np.float16, 1.25), ("f32", np.float32, 1.25), ("f64", np.float64, 1.25),
              ("i64", np.int64, 3), ("bool", np.bool_, True), ("c128", np.complex128, 1.25 + 0.5j)]
    astype_to = {"f16": np.float32, "f32": np.float64, "f64": np.float32, "i64": np.float64,
                 "bool": np.float32, "c128": np.complex64}
    for name, dt, val in dtypes:
        s = np.dtype(dt).itemsize
        a, b = filled((n,), dt, val), filled((n,), dt, val)
        out = cn.empty((n,), dtype=dt)
        report(f"C5 add {name}", n, 3 * s, run(lambda: cn.add(a, b, out=out)))
        report(f"C5 multiply {name}", n, 3 * s, run(lambda: cn.multiply(a, b, out=out)))
        mask = filled((n,), np.bool_, True)
        mask[n // 2:] = False
        report(f"C5 where {name}", n, 1 + 3 * s, run(lambda: cn.where(mask, a, b)))
        del mask, out, b
        dst = np.dtype(astype_to[name])
        report(f"C5 astype {name}->{dst.name}", n, s + dst.itemsize, run(lambda: a.astype(dst)))
        del a
    cn.runtime.release_cached_memory()
    fr = [c["frac"] for c in cases]
    return {"n_gpus": world, "reps": reps, "peak_gbs": peak,
            "frac_min": min(fr), "frac_median": statistics.median(fr),
            "passes": 2 if world > 1 else 1,
            "note": "per case: ms per call (CUDA events, max over ranks; at N > 1 the better of two passes), "
                    "algorithmic GB/s per GPU "
                    "(SURVEY §8d bytes per element) and its fraction of the measured HBM peak; "
                    "arrays are row-partitioned over the GPUs, every per-GPU array exceeds L2",
            "cases": cases}


# ------------------------------------------------------------------------------------------------
def main() -> None:
    args = parse_args()
    # NCCL's own log lines (NCCL_DEBUG=INFO/VERSION set by the caller) stay on: they go to stderr
    # unless the caller chose a file, so that stdout carries exactly one JSON line
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched without torchrun: re-exec under it (one rank per GPU)
        port = os.environ.get("MASTER_PORT", "29511")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", port,
               os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, dist = dist_setup(world)

    import cunumeric_b200 as cn

    cn.runtime.ensure_initialized()
    numa = pin_to_gpu_numa(cn.runtime.device) if world > 1 else {"bound": False}
    if world > 1:
        # one NCCL clique for the job: halo exchange, reduction combines
        cn.runtime.init_distributed(rank, world)
        print(f"[bench rank {rank}] NCCL communicator up: rank {rank} nranks {world} "
              f"cudaDev {cn.runtime.device}", file=sys.stderr, flush=True)
    sampler = ClockSampler(cn.runtime.device)
    sampler.start()

    if args.workload == "stencil":
        line = stencil_leg(args, rank, world, dist, sampler)
        line["clocks"] = sampler.stop()
        line["e2e"] = None if args.no_e2e else stencil_e2e(args, rank, world, dist)
        cn.runtime.release_cached_memory()
        if not args.no_extras:
            with cn.replicated():  # N independent replicas: the path has no exchange step
                bs = black_scholes_leg(args, rank, world, dist, sampler, primary=False)
            line["black_scholes"] = {k: bs[k] for k in ("metric", "value", "unit", "ms_per_step",
                                                        "scaling", "dtype", "config", "gpu_launches",
                                                        "roofline", "op_by_op", "e2e")}
            cn.runtime.release_cached_memory()
            line["sweeps"] = sweeps_leg(args, rank, world, dist)
            if world == 1 and not args.no_cpu_baseline:
                line["c1"] = c1_leg(args)
    else:
        with cn.replicated():
            line = black_scholes_leg(args, rank, world, dist, sampler, primary=True)
        line["clocks"] = sampler.stop()
    line["host_binding"] = numa
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args.workload, args)
            if args.workload == "stencil" and not args.no_extras:
                line["black_scholes"]["cpu_baseline"] = cpu_baseline("black_scholes", args)
        except Exception as exc:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": line["unit"], "cores": os.cpu_count(),
                   "kind": "reference", "sample": f"unavailable: {exc}"}
    line["cpu_baseline"] = cpu
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
