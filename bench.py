#!/usr/bin/env python
"""bench.py — the headline benchmark of the hot path.

    python bench.py --gpus N --steps K --warmup W [--workload black_scholes|stencil] [--impl reference]

Own arm: one "step" is one pass of the workload over one batch of synthetic input resident in HBM:
  black_scholes (default, BASELINE.json configs[1]): examples/black_scholes.py fp32, 1e8 options per
      GPU — 63 elementwise tasks per step, issued op-by-op through the cunumeric NumPy API exactly
      as the reference issues them.  metric = options (elements) per second.
  stencil (configs[3]): examples/stencil.py fp64 N=40000 row-partitioned over the GPUs, ITERS
      Jacobi iterations per step (6 tasks each), halo rows exchanged over NVLink.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU arithmetic
(oracle/_ref: its functor headers compiled with g++, OpenMP over all host cores) on a bounded
sample of the same workload."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "black_scholes_elements_per_second"
UNIT = "options/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="own", choices=["own", "reference"])
    p.add_argument("--workload", default="black_scholes", choices=["black_scholes", "stencil"])
    p.add_argument("--n", type=int, default=100_000_000, help="options per GPU (black_scholes)")
    p.add_argument("--stencil-n", type=int, default=40000)
    p.add_argument("--stencil-iters", type=int, default=10, help="Jacobi iterations per step")
    p.add_argument("--cpu-sample", type=int, default=10_000_000,
                   help="options in the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-chunks", type=int, default=8,
                   help="chunks per step of the e2e leg (upload / compute / download pipeline)")
    p.add_argument("--fusion", choices=["on", "off"], default="on",
                   help="on: elementwise chains run as fused kernels (default product path); "
                        "off: one kernel per task")
    p.add_argument("--no-e2e", action="store_true")
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
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r01_traffic.json): launch-weighted mean over its array-array and scalar-operand
    launches.  None when no capture covers this kernel."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            data = json.load(f)
        variants = data["kernels"][kernel]
        n = sum(v["launches_per_step"] for v in variants.values())
        total = sum((v["dram_read_bytes"] + v["dram_write_bytes"]) * v["launches_per_step"]
                    for v in variants.values())
        return total / n, "profiles/r01_traffic.json (ncu --set full, per launch)"
    except Exception:
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
    and to hand rank 0's NCCL unique id to the others."""
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


# ------------------------------------------------------------------------------------------------
def cpu_baseline_black_scholes(n_sample: int, reps: int = 2) -> dict:
    from oracle import ref, refnp
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs

    cores = os.cpu_count() or 1
    refnp.set_threads(cores)
    S, X, T = black_scholes_inputs(n_sample, np.float32, seed=0)
    S, X, T = refnp.array(S), refnp.array(X), refnp.array(T)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        black_scholes(S, X, T, 0.02, 0.3, xp=refnp)
        best = min(best, time.perf_counter() - t0)
    return {"value": n_sample / best, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{n_sample} options, best of {reps}; reference functors (oracle/_ref, "
                      f"g++ -O2 -fopenmp, schedule(static)) op-by-op, 63 tasks",
            "available": ref.available()}


def run_reference(args) -> None:
    """`--impl reference`: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refnp
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs

    cores = os.cpu_count() or 1
    refnp.set_threads(cores)
    n = args.cpu_sample
    S, X, T = black_scholes_inputs(n, np.float32, seed=0)
    S, X, T = refnp.array(S), refnp.array(X), refnp.array(T)
    for _ in range(min(args.warmup, 2)):
        black_scholes(S, X, T, 0.02, 0.3, xp=refnp)
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        black_scholes(S, X, T, 0.02, 0.3, xp=refnp)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    sample = (f"{n} options per step (bounded sample of the 1e8-option workload), {steps} steps; "
              "reference functors compiled from /root/reference/src (oracle/_ref), OpenMP")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_BS + "the reference's CPU functors issued op-by-op "
                                             "(OpenMP, all host cores) on a bounded sample per step",
                   "execution": "reference-cpu", "options_per_gpu": 100_000_000,
                   "options_per_step_sampled": n, "tasks_per_step": 63},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# the workload both arms (own / --impl reference) are measured on
WORKLOAD_BS = ("black_scholes fp32 1e8 options per GPU (examples/black_scholes.py, BASELINE.json "
               "configs[1]), 63 elementwise tasks per step through the NumPy API; ")


# ------------------------------------------------------------------------------------------------
def trace_summary(cn, n_records: int, peak_gbs: float):
    """Group the live per-launch records by kernel (task, op, dtype) and pick the dominant one."""
    import ctypes

    from cunumeric_b200 import _lib

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
    top_key, top = max(groups.items(), key=lambda kv: kv[1]["ms"])
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    names = {5: "BINARY_OP", 43: "UNARY_OP", 49: "WHERE", 11: "CONVERT", 33: "SCALAR_UNARY_RED",
             44: "UNARY_RED", 19: "FILL"}
    from cunumeric_b200.config import BinaryOpCode, UnaryOpCode

    opname = str(top_key[1])
    if top_key[0] == 5:
        opname = BinaryOpCode(top_key[1]).name
    elif top_key[0] == 43:
        opname = UnaryOpCode(top_key[1]).name
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": None,
        "kernel": (f"fused_kernel<{top_key[1]} elementwise tasks>" if top_key[0] == 1000 else
                   f"ew_kernel<{names.get(top_key[0], top_key[0])}:{opname}:dtype{top_key[2]}>"),
        "launches": top["launches"],
        "avg_launch_ms": top["ms"] / top["launches"],
        "algorithmic_bytes_per_launch": top["bytes"] / top["launches"],
        "share_of_step": top["ms"] / total_ms,
    }
    def _name(key):
        nm = str(key[1])
        if key[0] == 5:
            nm = BinaryOpCode(key[1]).name
        elif key[0] == 43:
            nm = UnaryOpCode(key[1]).name
        if key[0] == 1000:
            return f"FUSED:{key[1]} tasks"
        return f"{names.get(key[0], key[0])}:{nm}:dtype{key[2]}"

    roofline["per_kernel"] = {
        _name(k): {"launches": g["launches"], "ms": round(g["ms"], 3),
                   "gbs": round(g["bytes"] / (g["ms"] * 1e-3) / 1e9, 1)}
        for k, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
    all_bytes = sum(g["bytes"] for g in groups.values())
    whole = {"kernel_ms_total": total_ms, "algorithmic_gbs_all_kernels": all_bytes / (total_ms * 1e-3) / 1e9,
             "frac_all_kernels": all_bytes / (total_ms * 1e-3) / 1e9 / peak_gbs}
    return roofline, whole


def run_black_scholes(args, rank: int, world: int, dist) -> None:
    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib
    from cunumeric_b200.workloads import (BLACK_SCHOLES_BYTES_PER_OPTION_F32, BLACK_SCHOLES_TASKS,
                                          black_scholes, black_scholes_inputs)

    n = args.n
    cn.runtime.ensure_initialized()
    lib = cn.runtime.lib
    peak_gbs, peak_src = load_peaks()
    R, V = 0.02, 0.3
    # synthetic inputs, generated on the host once, staged in PINNED memory (for the e2e leg) and
    # made resident in HBM before the timed region
    Sh, Xh, Th = (cn.pinned_empty(n, np.float32) for _ in range(3))
    for dst, src in zip((Sh, Xh, Th), black_scholes_inputs(n, np.float32, seed=rank)):
        dst[...] = src
    S, X, T = cn.array(Sh), cn.array(Xh), cn.array(Th)
    cn.synchronize()

    from cunumeric_b200 import fusion
    import ctypes

    sampler = ClockSampler(cn.runtime.device)
    sampler.start()

    def timed(fused: bool, steps: int, sample_clocks: bool):
        """W warm-up steps, then `steps` timed steps: CUDA events on the launching stream, barrier +
        synchronize on both sides.  Every step ends with cn.flush(), so the step's results (call,
        put) are materialised in HBM inside the timed region — nothing is left pending."""
        fusion.set_mode("1" if fused else "0")
        for _ in range(max(args.warmup, 3)):
            call, put = black_scholes(S, X, T, R, V)
            cn.flush()
        cn.synchronize()
        ev0, ev1 = lib.cnb_event_create(), lib.cnb_event_create()
        _lib.check(lib.cnb_trace_start(steps * (BLACK_SCHOLES_TASKS + 8)))
        barrier(dist)
        cn.synchronize()
        if sample_clocks:
            sampler.mark_begin()
        launches0 = cn.runtime.launch_count()
        lib.cnb_event_record(ev0, cn.runtime.stream)
        for _ in range(steps):
            call, put = black_scholes(S, X, T, R, V)
            cn.flush()
        lib.cnb_event_record(ev1, cn.runtime.stream)
        cn.synchronize()
        if sample_clocks:
            sampler.mark_end()
        barrier(dist)
        launches = cn.runtime.launch_count() - launches0
        n_rec = lib.cnb_trace_stop()
        ms = ctypes.c_float()
        _lib.check(lib.cnb_event_elapsed_ms(ev0, ev1, ctypes.byref(ms)))
        elapsed = max_over_ranks(dist, ms.value * 1e-3)
        roofline, whole = trace_summary(cn, n_rec, peak_gbs)
        if roofline is not None:
            roofline["peak_source"] = peak_src
            roofline.update(whole)
            roofline["traffic"], roofline["traffic_source"] = load_traffic(roofline["kernel"])
        return elapsed, launches, roofline

    fused_on = args.fusion == "on" and fusion.enabled()
    # the op-by-op leg (one kernel per task, as the reference issues them) is always measured too
    obo_steps = args.steps if not fused_on else max(3, min(args.steps, 5))
    obo_elapsed, obo_launches, obo_roofline = timed(False, obo_steps, not fused_on)
    op_by_op = {"value": n * world * obo_steps / obo_elapsed, "unit": UNIT, "steps": obo_steps,
                "ms_per_step": 1e3 * obo_elapsed / obo_steps, "gpu_launches": int(obo_launches),
                "algorithmic_bytes_per_option": BLACK_SCHOLES_BYTES_PER_OPTION_F32,
                "whole_step_algorithmic_gbs":
                    BLACK_SCHOLES_BYTES_PER_OPTION_F32 * n * obo_steps / obo_elapsed / 1e9,
                "roofline": obo_roofline}
    if fused_on:
        elapsed, launches, roofline = timed(True, args.steps, True)
        bytes_per_option = 20  # 3 fp32 inputs + 2 fp32 outputs; 61 intermediates stay in registers
        if roofline is not None:
            roofline["limiter"] = ("instruction issue, not HBM: ~173 SASS instructions per option "
                                   "(4 IEEE divisions, 3 exp, log, sqrt, no FMA contraction — the "
                                   "bit-parity contract with op-by-op execution); ncu: 80 % of issue "
                                   "slots busy, DRAM traffic = algorithmic bytes "
                                   "(profiles/r01_fusion_ncu_summary.md)")
    else:
        elapsed, launches, roofline = obo_elapsed, obo_launches, obo_roofline
        bytes_per_option = BLACK_SCHOLES_BYTES_PER_OPTION_F32
    fusion.set_mode("1" if fused_on else "0")
    clocks = sampler.stop()
    value = n * world * args.steps / elapsed

    # ---- e2e: the same step through the public API from HOST buffers (pinned), copies inside
    e2e = None
    if not args.no_e2e:
        call_h, put_h = cn.pinned_empty(n, np.float32), cn.pinned_empty(n, np.float32)
        e2e_steps = max(2, min(args.steps, 5))

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
        e2e = {"value": n * world * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": 3 * 4 * n, "d2h_bytes_per_step": 2 * 4 * n,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "cunumeric_b200.map_chunks(black_scholes, pinned inputs, pinned outputs, "
                      f"chunk={chunk}): from_host(blocking=False) -> black_scholes() -> "
                      "to_host(blocking=False), 3 streams"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_black_scholes(args.cpu_sample)
        except Exception as exc:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {exc}"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_BS +
                                   ("the thunk layer captures the chain and runs it as ONE fused "
                                    "kernel (bit-identical to op-by-op; `op_by_op` holds the "
                                    "one-kernel-per-task measurement)" if fused_on else
                                    "issued op-by-op, one kernel per task"),
                       "execution": "fused" if fused_on else "op-by-op",
                       "options_per_gpu": n, "tasks_per_step": BLACK_SCHOLES_TASKS,
                       "algorithmic_bytes_per_option": bytes_per_option,
                       "l2_policy": "inputs, outputs (and op-by-op temporaries) are 400 MB each, "
                                    "larger than the 126 MB L2; no flush needed",
                       "whole_step_algorithmic_gbs":
                           bytes_per_option * n * args.steps / elapsed / 1e9},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "op_by_op": op_by_op, "cpu_baseline": cpu, "e2e": e2e,
        }))


def main() -> None:
    # rank 0 must print exactly one JSON line on stdout: keep NCCL's "NCCL version ..." banner
    # (NCCL_DEBUG=VERSION in some images) off it
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse_args()
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
    if args.workload == "black_scholes":
        run_black_scholes(args, rank, world, dist)
    else:
        from cunumeric_b200.bench_stencil import run_stencil

        run_stencil(args, rank, world, dist, ClockSampler, load_peaks, max_over_ranks, barrier,
                    trace_summary)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
