"""Measures the worst observed error of the ops whose parity bar is wider than 2 ulp (complex
transcendentals / divide / power, and CBRT fp64), in the units tests/test_parity_elementwise.py judges
them in, so that each bound can be set at 2x the observed worst case.

    python benchmarks/measure_tolerances.py   # on a B200; prints a dict literal"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_utils as pu  # noqa: E402
import test_parity_elementwise as t  # noqa: E402
from oracle import ref  # noqa: E402


def worst_scaled(got, exp, scale):
    part = got.real.dtype
    eps = np.finfo(part).eps
    ok = np.isfinite(exp) & np.isfinite(got)
    err = np.abs(got[ok].astype(np.complex128) - exp[ok].astype(np.complex128))
    sc = np.broadcast_to(np.asarray(scale, dtype=np.float64), exp.shape)[ok]
    return float((err / (eps * np.maximum(sc, float(np.finfo(part).tiny)))).max()) if err.size else 0.0


out = {}
for dt in pu.COMPLEX_DTYPES:
    for op in ref.BINARY_OPS:
        odt = ref.binary_out_dtype(op, dt)
        if odt is None or odt.kind != "c" or t.binary_tolerance(op, dt, odt) == 0:
            continue
        worst = 0.0
        for rep in range(4):
            rng = pu.rng_for("binary", op, dt.name) if rep == 0 else pu.rng_for("tol", op, dt.name, rep)
            a, b = t.binary_inputs(op, dt, rng)
            with np.errstate(all="ignore"):
                exp = ref.binary_op(op, a, b)
                got = pu.gpu_binary(op, a, b, odt, ())
                worst = max(worst, worst_scaled(got, exp, np.abs(exp)))
        out[("B", op, dt.name)] = worst
    for op in t.UNARY_SINGLE:
        odt = ref.unary_out_dtype(op, dt)
        if odt is None or t.unary_tolerance(op, dt, odt) == 0:
            continue
        worst = 0.0
        for rep in range(4):
            rng = pu.rng_for("unary", op, dt.name) if rep == 0 else pu.rng_for("tol", op, dt.name, rep)
            a = t.unary_inputs(op, dt, rng)
            with np.errstate(all="ignore"):
                exp = ref.unary_op(op, a)
                got = pu.gpu_unary(op, a, odt, ())
                scale = np.abs(exp) if op != "EXPM1" else np.maximum(np.abs(exp), 1.0)
                worst = max(worst, worst_scaled(got, exp, scale))
        out[("U", op, dt.name)] = worst
# CBRT fp64 in ulp
rng = pu.rng_for("unary", "CBRT", "float64")
a = t.unary_inputs("CBRT", np.dtype(np.float64), rng)
exp = ref.unary_op("CBRT", a)
got = pu.gpu_unary("CBRT", a, np.dtype(np.float64), ())
out[("U", "CBRT", "float64", "ulp")] = float(np.max(np.abs(got - exp) / np.spacing(np.abs(exp))))
for k, v in sorted(out.items()):
    print(f"    {k!r}: {v:.2f},")
