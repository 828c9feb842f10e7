"""GPU parity against the COMMITTED golden fixtures (tests/golden/oracle_vectors.npz, generated from
the reference's own functors by tests/golden/make_golden.py).  Unlike the seeded parity tests this
needs no oracle build at run time: stored inputs go through the CUDA path and are compared with the
stored reference outputs under the same bars (bit-exact for integer / boolean / index results and
IEEE arithmetic, <= 2 ulp for transcendentals, norm-wise for complex functions, n*eps for sums)."""
import os

import numpy as np
import pytest

import parity_utils as pu
import test_parity_elementwise as tpe
import test_parity_reductions as tpr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz")


@pytest.fixture(scope="module")
def golden():
    g = np.load(GOLDEN)
    groups = {}
    for k in g.files:
        head, leaf = k.rsplit("/", 1)
        groups.setdefault(head, {})[leaf] = g[k]
    return groups


def _keys(golden, kind):
    return sorted(k for k in golden if k.split("/")[0] == kind)


def test_golden_binary(golden):
    n = 0
    for k in _keys(golden, "binary"):
        _, op, dtn = k.split("/")
        v = golden[k]
        a, b, exp = v["a"], v["b"], v["out"]
        dt, odt = a.dtype, exp.dtype
        got = pu.gpu_binary(op, a, b, odt, (1e-3, 1e-5) if op == "ISCLOSE" else ())
        tol = tpe.binary_tolerance(op, dt, odt)
        if dt.kind == "c" and odt.kind == "c" and tol > 0:
            with np.errstate(all="ignore"):
                pu.assert_close_scaled(got, exp, np.abs(exp), tpe.COMPLEX_EPS.get(op, tol), k)
        elif op in ("LOGADDEXP", "LOGADDEXP2") and dt.kind == "f":
            scale = np.maximum(np.maximum(np.abs(a.astype(np.float64)), np.abs(b.astype(np.float64))), 1.0)
            pu.assert_close_scaled(got, exp, scale, 2 * tpe.TRANSCENDENTAL_ULP, k)
        else:
            pu.assert_close_ulp(got, exp, tol, k)
        n += 1
    assert n == 358


def test_golden_unary(golden):
    n = 0
    for k in _keys(golden, "unary"):
        _, op, dtn = k.split("/")
        v = golden[k]
        a, exp = v["a"], v["out"]
        dt, odt = a.dtype, exp.dtype
        extra = ()
        if op == "CLIP":
            extra = tuple(np.array(x).astype(dt) for x in (
                (False, True) if dt == np.bool_ else (-3, 5) if dt.kind != "u" else (2, 9)))
        got = pu.gpu_unary(op, a, odt, extra)
        tol = tpe.unary_tolerance(op, dt, odt)
        if dt.kind == "c" and tol > 0:
            with np.errstate(all="ignore"):
                scale = np.abs(exp) if op != "EXPM1" else np.maximum(np.abs(exp), 1.0)
                pu.assert_close_scaled(got, exp, scale, tol, k)
        else:
            pu.assert_close_ulp(got, exp, tol, k)
        n += 1
    assert n >= 333


def test_golden_multiout_and_convert(golden):
    from cunumeric_b200.config import ConvertCode, UnaryOpCode

    for k in _keys(golden, "multiout"):
        _, op, dtn = k.split("/")
        v = golden[k]
        a = v["a"]
        keep = np.isfinite(a)  # frexp's exponent for inf/nan is unspecified
        o1 = pu.new_thunk(a.shape, v["out1"].dtype)
        o2 = pu.new_thunk(a.shape, v["out2"].dtype)
        o1.unary_op(UnaryOpCode[op], pu.to_device(a), True, (), multiout=(o2,))
        pu.assert_close_ulp(o1.__numpy_array__()[keep], v["out1"][keep], 0, k + " out1")
        pu.assert_close_ulp(o2.__numpy_array__()[keep], v["out2"][keep], 0, k + " out2")
    n = 0
    for k in _keys(golden, "convert"):
        _, nan_op, srcn, dstn = k.split("/")
        v = golden[k]
        out = pu.new_thunk(v["a"].shape, np.dtype(dstn))
        out.convert(pu.to_device(v["a"]), nan_op=ConvertCode[nan_op])
        pu.assert_close_ulp(out.__numpy_array__(), v["out"], 0, k)
        n += 1
    assert n == 312


def test_golden_reductions(golden):
    n = 0
    for k in _keys(golden, "red"):
        _, op, dtn = k.split("/")
        v = golden[k]
        a = v["a"]
        args = (v["extra"][()],) if "extra" in v else None
        got = tpr.thunk_reduce(op, a, args=args)
        tpr.check_reduction(op, a, got, v["scalar"], a.size, k + " scalar")
        for axis in (0, 1):
            if f"axis{axis}" in v:
                got = tpr.thunk_reduce(op, a, axis=axis)
                tpr.check_reduction(op, a, got, v[f"axis{axis}"], a.shape[axis], f"{k} axis{axis}")
        n += 1
    assert n > 150


def test_golden_binary_red(golden):
    import cunumeric_b200 as cn

    n = 0
    for k in _keys(golden, "binred"):
        _, op, dtn = k.split("/")
        v = golden[k]
        A = cn.array(v["a"])
        for name, exp in zip(("b_same", "b_diff", "b_near"), v["out"]):
            B = cn.array(v[name])
            got = cn.array_equal(A, B) if op == "EQUAL" else cn.allclose(A, B, rtol=1e-3, atol=1e-5)
            assert bool(got) is bool(exp), (k, name)
        n += 1
    assert n == 28
