"""GPU parity at the FULL sizes of BASELINE.json's configs (VERDICT r1 task 2): byte offsets beyond
2^31 and 2^32, grids of ~10^6 CTAs' worth of tiles, 16 GiB operands.

  C5  add / where / astype over 2^30 elements, fp32 and complex128: the arrays are filled on the device
      and seeded with random data in windows — head, tail, +-4 KiB around every multiple of 2^31 and 2^32
      BYTES, and 64 random tiles — and every window of the result is compared with the oracle.
  C3  sum / max / argmax over 32768 x 32768 fp32 along axis 0, axis 1 and over everything, against the
      oracle on the whole 4 GiB array (OpenMP loops of the reference's functors).
  C2  Black-Scholes on 1e8 options, fused and op-by-op: bit-identical to each other everywhere, and
      within the whole-chain tolerance of the oracle on windows.

Bars as in the small-size suites: bit-exact for add / where / astype / max / arg-indices, n * eps for
sums."""
import os

import numpy as np
import pytest

from oracle import ref

import parity_utils as pu

pytestmark = pytest.mark.gpu

WINDOW = 4096  # elements per checked tile


def _free_gib() -> float:
    import ctypes

    import cunumeric_b200 as cn

    cn.runtime.ensure_initialized()
    free, total = ctypes.c_size_t(), ctypes.c_size_t()
    cn.runtime.lib.cnb_mem_info(ctypes.byref(free), ctypes.byref(total))
    return free.value / 2 ** 30


def windows(n: int, itemsize: int, rng) -> list:
    """Element ranges to check in an array of `n` elements of `itemsize` bytes."""
    out = {(0, min(n, WINDOW)), (max(0, n - WINDOW), n)}
    half = 4096 // itemsize
    for k in range(1, (n * itemsize) // 2 ** 31 + 1):
        e = k * 2 ** 31 // itemsize           # every multiple of 2^31 bytes (covers 2^32 too)
        if 0 < e < n:
            out.add((max(0, e - half), min(n, e + half)))
    for _ in range(64):
        s = int(rng.integers(0, max(1, n - WINDOW)))
        out.add((s, min(n, s + WINDOW)))
    return sorted(out)


def _seed(arr, host_chunks):
    """Overwrite the windows of the device array with the given host data."""
    import cunumeric_b200 as cn

    for (s, e), data in host_chunks.items():
        arr[s:e] = cn.array(data)


@pytest.mark.parametrize("dt", [np.float32, np.complex128], ids=lambda d: np.dtype(d).name)
def test_c5_elementwise_2_pow_30(dt):
    import cunumeric_b200 as cn

    dt = np.dtype(dt)
    n = 1 << 30
    need = 3.3 * n * dt.itemsize / 2 ** 30
    if _free_gib() < need + 4:
        pytest.skip(f"needs {need:.0f} GiB of device memory")
    rng = pu.rng_for("c5-scale", dt.name)
    wins = windows(n, dt.itemsize, rng)
    assert any(s * dt.itemsize >= 2 ** 31 for s, _ in wins)
    assert n * dt.itemsize <= 2 ** 32 or any(s * dt.itemsize >= 2 ** 32 for s, _ in wins)

    def rand(m):
        if dt.kind == "c":
            return (rng.uniform(-1, 1, m) + 1j * rng.uniform(-1, 1, m)).astype(dt)
        return rng.uniform(-1, 1, m).astype(dt)

    ha = {w: rand(w[1] - w[0]) for w in wins}
    hb = {w: rand(w[1] - w[0]) for w in wins}
    hm = {w: rng.random(w[1] - w[0]) < 0.5 for w in wins}
    a = cn.full((n,), 1.25, dtype=dt)
    b = cn.full((n,), -0.5, dtype=dt)
    _seed(a, ha)
    _seed(b, hb)
    # ---- add (bit-exact)
    out = cn.add(a, b)
    for w in wins:
        got = np.array(out[w[0]:w[1]])
        exp = ref.binary_op("ADD", ha[w], hb[w])
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), ("add", w)
    assert np.array_equal(np.array(out[12345678:12345678 + 8]), np.full(8, 0.75, dtype=dt))
    del out
    # ---- astype (bit-exact): fp32 -> fp64, complex128 -> complex64
    to = np.dtype(np.float64 if dt == np.float32 else np.complex64)
    conv = a.astype(to)
    for w in wins:
        got = np.array(conv[w[0]:w[1]])
        exp = ref.convert(ha[w], to)
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), ("astype", w)
    del conv
    # ---- where (byte-exact)
    mask = cn.full((n,), True, dtype=np.bool_)
    mask[n // 2:] = False
    _seed(mask, hm)
    sel = cn.where(mask, a, b)
    for w in wins:
        got = np.array(sel[w[0]:w[1]])
        exp = ref.where(hm[w], ha[w], hb[w])
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), ("where", w)
    unseeded = n // 2 + 777777
    assert np.array_equal(np.array(sel[unseeded:unseeded + 4]), np.full(4, -0.5, dtype=dt))
    del sel, mask, a, b
    cn.runtime.release_cached_memory()


def test_c3_reductions_32768_squared():
    import cunumeric_b200 as cn

    r = 32768
    if _free_gib() < 10:
        pytest.skip("needs 10 GiB of device memory")
    rng = pu.rng_for("c3-scale")
    # 4 GiB of fp32: a random 1024-row block repeated 32 times, each copy shifted by a different
    # offset so that maxima / arg-maxima are unique per column and per row
    block = rng.standard_normal((1024, r), dtype=np.float32)
    x = np.empty((r, r), dtype=np.float32)
    for k in range(32):
        np.add(block, np.float32(1e-3 * ((k * 7) % 32)), out=x[k * 1024:(k + 1) * 1024])
    del block
    X = cn.array(x)
    cores = os.cpu_count() or 1
    eps = np.finfo(np.float32).eps
    # ---- axis 0 / axis 1
    for axis in (0, 1):
        exp = ref.unary_red("SUM", x, axis, nthreads=cores)
        got = np.array(X.sum(axis=axis))
        bound = r * eps * np.abs(x).sum(axis=axis, dtype=np.float64)
        assert np.all(np.abs(got.astype(np.float64) - exp.astype(np.float64)) <= bound), ("sum", axis)
        assert np.array_equal(np.array(X.max(axis=axis)), ref.unary_red("MAX", x, axis, nthreads=cores))
        assert np.array_equal(np.array(X.argmax(axis=axis)),
                              ref.unary_red("ARGMAX", x, axis, nthreads=cores)["arg"]), ("argmax", axis)
    # ---- full: the oracle's sequential fp32 accumulation over 2^30 elements is itself n * eps away from
    # the exact sum; both are judged against the float64 sum with the n * eps bar
    exact = x.sum(dtype=np.float64)
    mag = np.abs(x).sum(dtype=np.float64)
    assert abs(float(X.sum()) - exact) <= x.size * eps * mag
    assert float(X.max()) == float(ref.scalar_unary_red("MAX", x, nthreads=cores))
    am = ref.scalar_unary_red("ARGMAX", x, nthreads=cores)
    assert int(X.argmax()) == int(am["arg"]) and int(am["arg"]) == int(x.argmax())
    del X
    cn.runtime.release_cached_memory()


def test_c2_black_scholes_1e8_fused_and_op_by_op():
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs
    from oracle import refnp

    n = 100_000_000
    if _free_gib() < 40:
        pytest.skip("needs 40 GiB of device memory (op-by-op temporaries)")
    S, X, T = black_scholes_inputs(n, np.float32, seed=0)
    dS, dX, dT = cn.array(S), cn.array(X), cn.array(T)
    old = fusion.set_mode("always")
    try:
        call_f, put_f = black_scholes(dS, dX, dT, 0.02, 0.3)
        cn.flush()
        fusion.set_mode("0")
        call_e, put_e = black_scholes(dS, dX, dT, 0.02, 0.3)
    finally:
        fusion.set_mode(old)
    # fused == op-by-op, bit for bit, over all 1e8 options (compared on the device: BINARY_RED EQUAL)
    assert bool(cn.array_equal(call_f, call_e)) and bool(cn.array_equal(put_f, put_e))
    rng = pu.rng_for("c2-scale")
    for s, e in windows(n, 4, rng):
        co, po = (r.a for r in black_scholes(refnp.array(S[s:e]), refnp.array(X[s:e]),
                                             refnp.array(T[s:e]), 0.02, 0.3, xp=refnp))
        for got, exp, name in ((np.array(call_f[s:e]), co, "call"), (np.array(put_f[s:e]), po, "put")):
            # 63 tasks, each within its own bar of the reference (bit-exact + - * /, <= 2 ulp exp / log /
            # sqrt, tests/test_parity_elementwise.py); the composition is judged at 32 eps of the
            # magnitudes that enter the final subtraction (S * cnd(d1) and X * exp(-rT) * cnd(d2))
            scale = np.abs(S[s:e]) + np.abs(X[s:e])
            assert np.all(np.abs(got - exp) <= 32 * np.finfo(np.float32).eps * scale), (name, s)
    del call_f, put_f, call_e, put_e, dS, dX, dT
    cn.runtime.release_cached_memory()
