"""Generates tests/golden/oracle_vectors.npz from oracle/_ref (the reference's own functors compiled
from /root/reference/src by oracle/Makefile).  Run where /root/reference is mounted:

    make -C oracle && python tests/golden/make_golden.py

The fixtures pin (a) the oracle build that travels to the GPU box and (b) the CUDA path when the
oracle shared object is not available there.  Inputs are seeded; every (opcode, dtype) pair the
reference marks valid is covered with 96 elements."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref  # noqa: E402

import parity_utils as pu  # noqa: E402
import test_parity_elementwise as tpe  # noqa: E402
import test_parity_reductions as tpr  # noqa: E402

N = 96


def main() -> None:
    out = {}
    tpe.N = N
    for op in ref.BINARY_OPS:
        for dt in pu.DTYPES:
            if ref.binary_out_dtype(op, dt) is None:
                continue
            rng = pu.rng_for("golden-binary", op, dt.name)
            a, b = tpe.binary_inputs(op, dt, rng)
            with np.errstate(all="ignore"):
                r = ref.binary_op(op, a, b, 1e-3, 1e-5)
            k = f"binary/{op}/{dt.name}"
            out[k + "/a"], out[k + "/b"], out[k + "/out"] = a, b, r
    for op in tpe.UNARY_SINGLE:
        for dt in pu.DTYPES:
            if ref.unary_out_dtype(op, dt) is None:
                continue
            rng = pu.rng_for("golden-unary", op, dt.name)
            a = tpe.unary_inputs(op, dt, rng)
            extra = None
            if op == "CLIP":
                extra = tuple(np.array(v).astype(dt) for v in (
                    (False, True) if dt == np.bool_ else (-3, 5) if dt.kind != "u" else (2, 9)))
            with np.errstate(all="ignore"):
                r = ref.unary_op(op, a, extra=extra)
            k = f"unary/{op}/{dt.name}"
            out[k + "/a"], out[k + "/out"] = a, r
    for op in ("FREXP", "MODF"):
        for dt in pu.FLOAT_DTYPES:
            rng = pu.rng_for("golden-multi", op, dt.name)
            a = pu.make_input(dt, N, rng)
            r1, r2 = ref.unary_multiout(op, a)
            k = f"multiout/{op}/{dt.name}"
            out[k + "/a"], out[k + "/out1"], out[k + "/out2"] = a, r1, r2
    for nan_op in ref.CONVERT_OPS:
        for src in pu.DTYPES:
            if nan_op != "NOOP" and src.kind not in "fc":
                continue
            for dst in pu.DTYPES:
                if src == dst:
                    continue
                rng = pu.rng_for("golden-convert", nan_op, src.name, dst.name)
                a = pu.make_input(src, N, rng, "small")
                if dst.kind in "ub" and src.kind in "fc":
                    a = (np.abs(a.real).astype(src) if src.kind == "f"
                         else (np.abs(a.real) + 1j * a.imag).astype(src))
                if dst.kind == "u" and src.kind == "i":
                    a = np.abs(a)
                if nan_op != "NOOP":
                    a[::7] = np.nan
                with np.errstate(all="ignore"):
                    r = ref.convert(a, dst, nan_op)
                k = f"convert/{nan_op}/{src.name}/{dst.name}"
                out[k + "/a"], out[k + "/out"] = a, r
    for op in ref.RED_OPS:
        for dt in pu.DTYPES:
            try:
                ref.red_identity(op, dt)
            except ref.InvalidOp:
                continue
            rng = pu.rng_for("golden-red", op, dt.name)
            a = tpr.red_input(op, dt, (12, 8), rng)
            extra = None
            if op == "CONTAINS":
                extra = a[3, 3]
            if op == "VARIANCE":
                extra = np.array(1).astype(dt)
            pre = tpr.python_prefill(op, dt)
            k = f"red/{op}/{dt.name}"
            out[k + "/a"] = a
            out[k + "/scalar"] = ref.scalar_unary_red(op, a, extra=extra, initial=pre)
            if extra is not None:
                out[k + "/extra"] = np.asarray(extra)
            if op != "CONTAINS":
                out[k + "/axis0"] = ref.unary_red(op, a, 0, initial=pre)
                out[k + "/axis1"] = ref.unary_red(op, a, 1, initial=pre)
    # BINARY_RED (array_equal / allclose): equal pair, one mismatch, one near-miss per dtype
    for op in ("EQUAL", "ISCLOSE"):
        for dt in pu.DTYPES:
            rng = pu.rng_for("golden-binred", op, dt.name)
            a = pu.make_input(dt, N, rng, "small")
            if dt.kind in "fc":
                a = np.where(np.isnan(a) | np.isinf(a), np.ones_like(a), a)
            b_same = a.copy()
            b_diff = a.copy()
            b_diff[N // 3] = (not b_diff[N // 3]) if dt.kind == "b" else b_diff[N // 3] + dt.type(1)
            b_near = a.copy()
            if dt.kind in "fc":
                b_near = (a * (1 + 3e-4)).astype(dt)
            k = f"binred/{op}/{dt.name}"
            out[k + "/a"], out[k + "/b_same"], out[k + "/b_diff"], out[k + "/b_near"] = a, b_same, b_diff, b_near
            out[k + "/out"] = np.array([ref.binary_red(op, a, b, 1e-3, 1e-5)
                                        for b in (b_same, b_diff, b_near)])
    path = os.path.join(HERE, "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays to {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
