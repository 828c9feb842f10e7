"""Per-kernel throughput sweeps for BASELINE.json configs[2] (C3: sum/max/argmax over 32768x32768
fp32 along axis 0, axis 1 and full) and configs[4] (C5: add / multiply / where / astype over 2^30
elements for fp16, fp32, fp64, int64, bool, complex128).

    python benchmarks/sweep.py [--c3] [--c5] [--log2n 30] [--reps 10]
    torchrun --nproc-per-node N benchmarks/sweep.py ...      # arrays row-partitioned over N GPUs

Prints one JSON line per case: elements/s, algorithmic GB/s (SURVEY §8d bytes per element) and the
fraction of the measured HBM peak, per GPU.  Inputs are device-generated (fills); every array is
far larger than the 126 MB L2.  Timing: CUDA events on the compute stream, max over ranks."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cunumeric_b200 as cn  # noqa: E402
from cunumeric_b200 import _lib  # noqa: E402


def peak_gbs() -> float:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"])
    return 6650.0


class Timer:
    def __init__(self, dist) -> None:
        self.lib = cn.runtime.lib
        self.dist = dist
        self.e0, self.e1 = self.lib.cnb_event_create(), self.lib.cnb_event_create()

    def run(self, fn, reps: int, warmup: int = 3) -> float:
        def fn(fn=fn):  # issue the task now: a pending chain would be deduplicated by fusion
            r = fn()
            cn.flush()
            return r

        for _ in range(warmup):
            fn()
        cn.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.lib.cnb_event_record(self.e0, cn.runtime.stream)
        for _ in range(reps):
            fn()
        self.lib.cnb_event_record(self.e1, cn.runtime.stream)
        cn.synchronize()
        ms = ctypes.c_float()
        _lib.check(self.lib.cnb_event_elapsed_ms(self.e0, self.e1, ctypes.byref(ms)))
        sec = ms.value * 1e-3 / reps
        if self.dist is not None:
            import torch

            t = torch.tensor([sec], dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec


def report(rank, world, name, elems, bytes_per_elem, sec, peak, extra=None):
    if rank != 0:
        return
    gbs = elems * bytes_per_elem / sec / 1e9
    rec = {"case": name, "n_gpus": world, "elements": elems, "ms": sec * 1e3,
           "elements_per_s": elems / sec, "bytes_per_element": bytes_per_elem,
           "algorithmic_gbs_total": gbs, "algorithmic_gbs_per_gpu": gbs / world,
           "frac_of_hbm_peak_per_gpu": gbs / world / peak, "peak_gbs": peak}
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)


def filled(shape, dtype, value):
    a = cn.empty(shape, dtype=dtype)
    a.fill(value)
    return a


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3", action="store_true")
    ap.add_argument("--c5", action="store_true")
    ap.add_argument("--cvt", action="store_true")
    ap.add_argument("--fill", action="store_true", help="write-only / read-only ceilings")
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    if not (args.c3 or args.c5 or args.cvt or args.fill):
        args.c3 = args.c5 = True
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    cn.runtime.ensure_initialized()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")
        cn.runtime.init_distributed(rank, world)
    peak = peak_gbs()
    timer = Timer(dist)

    if args.c5:
        n = 1 << args.log2n
        cases = [("float16", np.float16, 1.25), ("float32", np.float32, 1.25),
                 ("float64", np.float64, 1.25), ("int64", np.int64, 3), ("bool", np.bool_, True),
                 ("complex128", np.complex128, 1.25 + 0.5j)]
        astype_to = {"float16": np.float32, "float32": np.float64, "float64": np.float32,
                     "int64": np.float64, "bool": np.float32, "complex128": np.complex64}
        for name, dt, val in cases:
            s = np.dtype(dt).itemsize
            a = filled((n,), dt, val)
            b = filled((n,), dt, val)
            out = cn.empty((n,), dtype=dt)
            sec = timer.run(lambda: cn.add(a, b, out=out), args.reps)
            report(rank, world, f"C5 add {name}", n, 3 * s, sec, peak)
            sec = timer.run(lambda: cn.multiply(a, b, out=out), args.reps)
            report(rank, world, f"C5 multiply {name}", n, 3 * s, sec, peak)
            mask = filled((n,), np.bool_, True)
            mask[n // 2:] = False
            sec = timer.run(lambda: cn.where(mask, a, b), args.reps)
            report(rank, world, f"C5 where {name}", n, 1 + 3 * s, sec, peak)
            del mask, out, b
            dst = np.dtype(astype_to[name])
            sec = timer.run(lambda: a.astype(dst), args.reps)
            report(rank, world, f"C5 astype {name}->{dst.name}", n, s + dst.itemsize, sec, peak)
            del a

    if args.fill:
        # context for write-heavy kernels: a pure store stream (FILL), a 1:1 copy and a pure load
        # stream (scalar SUM) — the measured "peak" in MEASURED_PEAKS.json is a 1:1 copy
        n = 1 << args.log2n
        for name, dt in (("float32", np.float32), ("float64", np.float64), ("int8", np.int8)):
            s = np.dtype(dt).itemsize
            a = cn.empty((n,), dtype=dt)
            sec = timer.run(lambda: a.fill(1), args.reps)
            report(rank, world, f"FILL {name} (write only)", n, s, sec, peak)
            b = cn.empty((n,), dtype=dt)
            sec = timer.run(lambda: b._thunk.copy(a._thunk, deep=True), args.reps)
            report(rank, world, f"COPY {name} (1:1)", n, 2 * s, sec, peak)
            sec = timer.run(lambda: a.sum(), args.reps)
            report(rank, world, f"SUM {name} (read only)", n, s, sec, peak)
            del a, b

    if args.cvt:
        # every CONVERT pair (14 x 13) — hunts for conversion instructions that issue slowly
        n = 1 << 27
        dts = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32,
               np.uint64, np.float16, np.float32, np.float64, np.complex64, np.complex128]
        for src in dts:
            a = filled((n,), src, 1)
            for dst in dts:
                if dst == src:
                    continue
                out = cn.empty((n,), dtype=dst)
                sec = timer.run(lambda: out._thunk.convert(a._thunk), 5, warmup=2)
                report(rank, world, f"CVT {np.dtype(src).name}->{np.dtype(dst).name}", n,
                       np.dtype(src).itemsize + np.dtype(dst).itemsize, sec, peak)
                del out
            del a

    if args.c3:
        r = args.rows
        x = filled((r, r), np.float32, 0.5)
        x[r // 3, :] = 2.0
        x[:, r // 5] = 3.0
        n = r * r
        for opname, fn in (("sum", lambda ax: x.sum(axis=ax)), ("max", lambda ax: x.max(axis=ax)),
                           ("argmax", lambda ax: x.argmax(axis=ax))):
            for ax, label in ((0, "axis0"), (1, "axis1"), (None, "full")):
                sec = timer.run(lambda: fn(ax), args.reps)
                report(rank, world, f"C3 {opname} {label} 32768x32768 float32"
                       if r == 32768 else f"C3 {opname} {label} {r}x{r} float32", n, 4, sec, peak)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
