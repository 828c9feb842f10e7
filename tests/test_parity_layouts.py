"""GPU parity for the layouts the task contract allows beyond dense: row-pitched
slices (the stencil operands), stride-0 broadcast operands and scalars, transposed / stepped views,
unaligned starts, 3-D/4-D rects, empty rects, in-place and overlapping operands."""
import numpy as np
import pytest

from oracle import ref

import parity_utils as pu

pytestmark = pytest.mark.gpu


def dev(a):
    import cunumeric_b200 as cn

    return cn.array(a)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.float16, np.int64, np.bool_, np.complex128],
                         ids=lambda d: np.dtype(d).name)
def test_stencil_views(dt):
    """examples/stencil.py:44-49 operand shapes: shifted (N,N) views of an (N+2)^2 grid."""
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    n = 301
    rng = pu.rng_for("stencil", dt.name)
    g = pu.make_input(dt, (n + 2) * (n + 2), rng, "small").reshape(n + 2, n + 2)
    G = cn.array(g)
    views = lambda x: (x[1:-1, 1:-1], x[0:-2, 1:-1], x[1:-1, 2:], x[1:-1, 0:-2], x[2:, 1:-1])
    c, no, e, w, s = views(G)
    hc, hn, he, hw, hs = views(g)
    got = (c + no + e + w + s).__array__()
    with np.errstate(all="ignore"):
        exp = ref.binary_op("ADD", ref.binary_op("ADD", ref.binary_op("ADD", ref.binary_op(
            "ADD", hc, hn), he), hw), hs)
    pu.assert_close_ulp(got, exp, 0, "5-point sum")
    # write back into the interior view (UNARY_OP COPY into a pitched destination)
    c[:] = cn.array(exp)
    g2 = g.copy()
    g2[1:-1, 1:-1] = exp
    assert np.array_equal(G.__array__(), g2, equal_nan=dt.kind in "fc")


@pytest.mark.parametrize("shape_a,shape_b", [((7, 1), (1, 9)), ((4, 5, 6), (6,)), ((3, 1, 5, 1), (1, 4, 1, 7)),
                                             ((1,), (1000,)), ((), (37,)), ((2, 3), ())])
def test_broadcast_operands(shape_a, shape_b):
    rng = pu.rng_for("bcast", shape_a, shape_b)
    a = rng.normal(size=shape_a).astype(np.float32)
    b = rng.normal(size=shape_b).astype(np.float32)
    exp = ref.binary_op("MULTIPLY", *np.broadcast_arrays(a, b))
    got = pu.gpu_binary("MULTIPLY", a, b, np.float32)
    pu.assert_close_ulp(got, exp.reshape(got.shape), 0, "broadcast multiply")


def test_python_scalar_operand_keeps_dtype():
    import cunumeric_b200 as cn

    a = np.linspace(0, 1, 1000, dtype=np.float32)
    A = cn.array(a)
    out = 0.2 * A
    assert out.dtype == np.float32  # float32_array * 0.2 must stay float32
    exp = ref.binary_op("MULTIPLY", np.full_like(a, np.float32(0.2)), a)
    pu.assert_close_ulp(out.__array__(), exp, 0, "0.2*a")


def test_transposed_and_stepped_views():
    import cunumeric_b200 as cn

    rng = pu.rng_for("transposed")
    a = rng.normal(size=(130, 70))
    b = rng.normal(size=(70, 130))
    A, B = cn.array(a), cn.array(b)
    pu.assert_close_ulp((A.T + B).__array__(), ref.binary_op("ADD", a.T, b), 0, "A.T + B")
    pu.assert_close_ulp((A.T * B.T.T).__array__(), ref.binary_op("MULTIPLY", a.T, b), 0, "A.T*B")
    pu.assert_close_ulp((A[::2, 1::3] - A[1::2, 2::3]).__array__(),
                        ref.binary_op("SUBTRACT", a[::2, 1::3], a[1::2, 2::3]), 0, "stepped")
    pu.assert_close_ulp(cn.exp(A[::-1, ::-2]).__array__(), ref.unary_op("EXP", a[::-1, ::-2]), 2,
                        "negative strides")
    # uniformly transposed task: output and inputs share the permuted layout
    out = cn.empty((130, 70)).T
    cn.add(A.T, A.T, out=out)
    pu.assert_close_ulp(out.__array__(), ref.binary_op("ADD", a.T, a.T), 0, "all transposed")


@pytest.mark.parametrize("shape", [(3, 4, 5), (2, 3, 4, 5), (6, 1, 7), (1, 1, 1, 1)])
def test_nd_rects(shape):
    import cunumeric_b200 as cn

    rng = pu.rng_for("nd", shape)
    a = rng.normal(size=shape).astype(np.float32)
    A = cn.array(a)
    sl = tuple(slice(0, None) if n < 3 else slice(1, n - 1) for n in shape)
    exp = ref.binary_op("ADD", a[sl], a[sl])
    pu.assert_close_ulp((A[sl] + A[sl]).__array__(), exp, 0, "nd views")
    pu.assert_close_ulp(cn.negative(A).__array__(), -a, 0, "nd dense")


def test_unaligned_starts_all_dtypes():
    import cunumeric_b200 as cn

    for dt in pu.DTYPES:
        rng = pu.rng_for("unaligned", dt.name)
        a = pu.make_input(dt, 9000, rng, "small")
        A = cn.array(a)
        for off in (1, 3):
            got = cn.add(A[off:], A[:-off]).__array__()
            with np.errstate(all="ignore"):
                exp = ref.binary_op("ADD", a[off:], a[:-off])
            pu.assert_close_ulp(got, exp, 0, f"unaligned {dt.name} +{off}")


def test_empty_and_scalar_rects():
    import cunumeric_b200 as cn

    for shape in [(0,), (0, 5), (3, 0, 2)]:
        a = cn.zeros(shape, dtype=np.float32)
        assert (a + a).shape == shape
        assert (a + a).__array__().size == 0
    s = cn.array(np.float64(3.0))
    assert float(s * s) == 9.0


def test_inplace_and_overlap():
    """tests/integration/test_overlap.py:24-78 semantics (deferred.py:291-303)."""
    import cunumeric_b200 as cn

    a = np.arange(1, 2001, dtype=np.float64)
    A = cn.array(a)
    A += A
    a += a
    assert np.array_equal(A.__array__(), a)
    A[1:] += A[:-1]
    a[1:] += a[:-1]
    assert np.array_equal(A.__array__(), a)
    cn.sin(A[:-1], out=A[1:])
    np.sin(a[:-1], out=a[1:])
    pu.assert_close_ulp(A.__array__(), a, 2, "sin overlap")
    A[1:] = A[:-1]
    a[1:] = a[:-1].copy()
    pu.assert_close_ulp(A.__array__(), a, 2, "shift copy")
