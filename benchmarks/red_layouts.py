"""UNARY_RED / SCALAR_UNARY_RED on transposed, strided and sliced views (north_star bullet 2 asks about
"strided and transposed" reductions): algorithmic GB/s and fraction of the measured HBM peak per case.

    python benchmarks/red_layouts.py [--n 16384] [--reps 10]

The plan canonicalises dimensions by stride, so a transposed operand runs the kernel of the axis it
really reduces over (COLUMN <-> ROW swap); stepped views are bounded by the sectors they touch — the
`sector_frac` column is the fraction of the peak counting every 32-byte sector the view touches."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import cunumeric_b200 as cn  # noqa: E402
from sweep import Timer, peak_gbs  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    n = args.n
    cn.runtime.ensure_initialized()
    peak = peak_gbs()
    timer = Timer(None)
    x = cn.full((n, n), 1.0, dtype=np.float32)
    x[n // 3, n // 5] = 7.0
    cases = [
        ("x.sum(axis=0)  [contiguous, COLUMN]", lambda: x.sum(axis=0), n * n, 1.0),
        ("x.sum(axis=1)  [contiguous, ROW]", lambda: x.sum(axis=1), n * n, 1.0),
        ("x.T.sum(axis=0)  [transposed -> ROW kernel]", lambda: x.T.sum(axis=0), n * n, 1.0),
        ("x.T.sum(axis=1)  [transposed -> COLUMN kernel]", lambda: x.T.sum(axis=1), n * n, 1.0),
        ("x.T.max(axis=1)", lambda: x.T.max(axis=1), n * n, 1.0),
        ("x.T.argmax(axis=0)", lambda: x.T.argmax(axis=0), n * n, 1.0),
        ("x.T.sum()  [transposed, full]", lambda: x.T.sum(), n * n, 1.0),
        ("x[1:-1, 1:-1].sum(axis=0)  [pitched, misaligned]", lambda: x[1:-1, 1:-1].sum(axis=0),
         (n - 2) * (n - 2), 1.0),
        ("x[1:-1, 1:-1].sum(axis=1)  [pitched, misaligned]", lambda: x[1:-1, 1:-1].sum(axis=1),
         (n - 2) * (n - 2), 1.0),
        ("x[1:-1, 1:-1].sum()  [pitched, misaligned, full]", lambda: x[1:-1, 1:-1].sum(),
         (n - 2) * (n - 2), 1.0),
        ("x[::2].sum(axis=0)  [every other row]", lambda: x[::2].sum(axis=0), n * n // 2, 1.0),
        ("x[::2].sum(axis=1)  [every other row]", lambda: x[::2].sum(axis=1), n * n // 2, 1.0),
        ("x[:, ::2].sum(axis=0)  [every other column]", lambda: x[:, ::2].sum(axis=0), n * n // 2, 2.0),
        ("x[:, ::2].sum(axis=1)  [every other column]", lambda: x[:, ::2].sum(axis=1), n * n // 2, 2.0),
        ("x[:, ::8].sum(axis=1)  [every 8th column = one per sector]", lambda: x[:, ::8].sum(axis=1),
         n * n // 8, 8.0),
    ]
    for name, fn, elems, sector_ratio in cases:
        sec = timer.run(fn, args.reps)
        gbs = elems * 4 / sec / 1e9
        print(json.dumps({"case": name, "elements": elems, "ms": round(sec * 1e3, 4),
                          "algorithmic_gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                          "sector_frac": round(gbs * sector_ratio / peak, 4)}), flush=True)
    # correctness of the views timed above, against NumPy on a small instance
    rng = np.random.default_rng(0)
    a = rng.standard_normal((257, 130)).astype(np.float32)
    A = cn.array(a)
    ok = (np.allclose(A.T.sum(axis=0).__array__(), a.T.sum(axis=0), rtol=1e-4, atol=1e-4)
          and np.array_equal(A.T.argmax(axis=0).__array__(), a.T.argmax(axis=0))
          and np.array_equal(A.T.max(axis=1).__array__(), a.T.max(axis=1))
          and np.allclose(A[:, ::2].sum(axis=1).__array__(), a[:, ::2].sum(axis=1), rtol=1e-4, atol=1e-4)
          and np.allclose(A[::2].sum(axis=0).__array__(), a[::2].sum(axis=0), rtol=1e-4, atol=1e-4))
    print(json.dumps({"views_match_numpy": bool(ok)}))


if __name__ == "__main__":
    main()
