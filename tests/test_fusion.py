"""Lazy fusion of elementwise chains (cunumeric_b200/fusion.py).

CPU part: chains are captured and their kernels generated + compiled with nvcc WITHOUT a device
(the dry-run tracer `__graft_entry__.build()` uses).  GPU part: every program below is run twice —
fused (`always`: compile at first sight) and op-by-op (`0`) — and the results must be bit-identical,
because the fused kernel composes the same device functors with the same rounding points."""
import os
import shutil

import numpy as np
import pytest

import parity_utils as pu


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def run_both(program):
    """program() -> list of cunumeric arrays; returns (fused results, eager results, fused stats)."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    old = fusion.set_mode("always")
    try:
        before = dict(fusion.stats)
        fused = [np.array(x.__array__()) for x in program()]
        delta = {k: fusion.stats[k] - before[k] for k in before}
        fusion.set_mode("0")
        eager = [np.array(x.__array__()) for x in program()]
    finally:
        fusion.set_mode(old)
    return fused, eager, delta


def assert_identical(fused, eager):
    assert len(fused) == len(eager)
    for f, e in zip(fused, eager):
        assert f.dtype == e.dtype and f.shape == e.shape
        assert np.array_equal(_bits(f), _bits(e)), (f.dtype, np.flatnonzero(_bits(f) != _bits(e))[:5])


# ------------------------------------------------------------------------------------ CPU
def test_trace_and_compile_without_a_device(tmp_path, monkeypatch):
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from cunumeric_b200.workloads import black_scholes, stencil_init, stencil_run

    if cn.runtime.lib is not None:
        pytest.skip("runtime already initialised on a device")
    monkeypatch.setattr(fusion, "_CACHE_DIR", str(tmp_path))
    fusion._kernels.clear()

    def bs():
        S, X, T = (cn.empty((1000,), dtype=np.float32) for _ in range(3))
        out = black_scholes(S, X, T, 0.02, 0.3)
        cn.flush()
        return out

    assert fusion.trace_only(bs) == 1
    (src,) = [f for f in os.listdir(tmp_path) if f.endswith(".cu")]
    text = open(tmp_path / src).read()
    assert "63 tasks, 16 inputs, 2 stored outputs" in text  # 61 intermediates stay in registers
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))

    held = []

    def stencil():
        # lazily allocated grid: the boundary fills are skipped in a dry run
        g = cn.empty((66, 66), dtype=np.float64)
        held.append((g, stencil_run(g, 2)))

    # per iteration ONE chain: [4 ADD + MULTIPLY + COPY] — the COPY back into `center` overwrites the
    # shifted views the chain reads, which write-after-read renaming of the grid buffer makes legal.
    # Chains are launched one iteration late: iteration 1 stores only the new interior (`average`
    # and `work` of that iteration are dead by then), the last one also the returned `work`.
    # Each of the two chains is compiled in two flavours: vector / strided, and TMA-staged tiles (the
    # five shifted views are ONE tensor-map group: one box with a halo serves them all).
    n = fusion.trace_only(stencil)
    assert n == 4
    texts = [open(tmp_path / f).read() for f in os.listdir(tmp_path) if f.endswith(".cu")]
    assert any("6 tasks, 6 inputs, 1 stored outputs" in t for t in texts), [t[:120] for t in texts]
    assert any("6 tasks, 6 inputs, 2 stored outputs" in t for t in texts), [t[:120] for t in texts]
    tma = [t for t in texts if "(TMA flavour)" in t]
    assert len(tma) == 2 and all("6 inputs in 1 tensor-map group(s)" in t for t in tma)
    assert all("tma_load_2d(" in t and "g0_1_1" in t for t in tma)


def test_hazard_rules_dry():
    """Windows that overlap a chain output through a DIFFERENT window force a flush."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    if cn.runtime.lib is not None:
        pytest.skip("runtime already initialised on a device")
    seen = []

    def prog():
        a = cn.empty((100,), dtype=np.float32)
        b = a + 1.0          # chain: [t0]
        c = b * 2.0          # same window of b -> joins: [t0, t1]
        seen.append(len(fusion._chain.tasks))
        d = b[1:] + 1.0      # different shape -> flush, new chain [t2]
        seen.append(len(fusion._chain.tasks))
        a[10:20] = 3.0       # fill is not elementwise-captured: flushes through Store.ptr in real runs
        e = cn.empty((99,), dtype=np.float32)
        f = e + 1.0
        e[:] = f             # writes the window the chain read through the same window: joins
        seen.append(len(fusion._chain.tasks))
        return c, d

    cn.runtime.dry_run = True
    old = fusion.set_mode("always")
    try:
        prog()
    except RuntimeError:
        pass  # the fill needs a device; everything before it is what we check
    finally:
        fusion._chain = fusion._Chain()
        fusion._queue.clear()
        fusion.set_mode(old)
        cn.runtime.dry_run = False
        fusion.drop_scalar_caches()
    assert seen[:2] == [2, 1]


def _dry(prog):
    """Run prog() with capture on and no device; returns the chains (task kinds, live outputs) that
    were flushed, in order."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    if cn.runtime.lib is not None:
        pytest.skip("runtime already initialised on a device")
    flushed = []
    _dry.renames = []
    orig = fusion._run_chain

    def spy(c, overlap=None):
        live = sum(1 for _, w in c.written.values() if w.buffer.users > 0)
        flushed.append(([t.kind + str(t.op) for t in c.tasks], live))
        _dry.renames.append((len(c.tasks), live, len(c.renamed)))
        # no compilation in these tests: the chain bookkeeping is what is checked

    cn.runtime.dry_run = True
    old = fusion.set_mode("always")
    fusion._run_chain = spy
    try:
        keep = prog()
        fusion.flush()
    finally:
        fusion._run_chain = orig
        fusion._chain = fusion._Chain()
        fusion._queue.clear()
        fusion.set_mode(old)
        cn.runtime.dry_run = False
        fusion.drop_scalar_caches()
    del keep
    return flushed


def test_dead_temporaries_are_not_outputs_dry():
    import cunumeric_b200 as cn

    def prog():
        a = cn.empty((1000,), dtype=np.float32)
        b = cn.empty((1000,), dtype=np.float32)
        r = cn.sqrt(a * a + b * b) / (a + 1.0)   # 6 tasks, 5 temporaries die immediately
        return r

    (chain,) = _dry(prog)
    tasks, live = chain
    assert len(tasks) == 6 and live == 1


def test_inplace_update_joins_but_shifted_write_flushes_dry():
    import cunumeric_b200 as cn

    def prog():
        a = cn.empty((64, 64), dtype=np.float64)
        b = cn.empty((64, 64), dtype=np.float64)
        a += b                 # reads and writes the same window: joins
        a *= 2.0               # joins, value of `a` comes from registers
        c = a[1:, :] + 1.0     # a different window of something the chain wrote: flush first
        a[:-1, :] = c          # write-after-read on a[1:, :]: joins, `a`'s buffer is RENAMED
        d = a[:20, :] + 1.0
        a[10:30, :] = d        # write-after-read again, but the window is < half the buffer: flush
        return a, c, d

    chains = _dry(prog)
    assert [len(t) for t, _ in chains] == [2, 2, 1, 1]
    assert [r for _, _, r in _dry.renames] == [0, 1, 0, 0]


def test_chain_length_and_shape_changes_flush_dry():
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    def prog():
        x = cn.empty((128,), dtype=np.float32)
        y = x
        for _ in range(fusion.MAX_TASKS + 5):
            y = y + 1.0
        z = cn.empty((64,), dtype=np.float32) * 2.0   # another shape
        return y, z

    chains = _dry(prog)
    assert [len(t) for t, _ in chains] == [fusion.MAX_TASKS, 5, 1]
    assert all(live == 1 for _, live in chains[1:])


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.float16], ids=lambda d: np.dtype(d).name)
def test_black_scholes_fused_is_bit_identical(dt):
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import black_scholes, black_scholes_inputs

    S, X, T = black_scholes_inputs(100003, np.float32)
    S, X, T = (v.astype(dt) for v in (S, X, T))

    def prog():
        return black_scholes(cn.array(S), cn.array(X), cn.array(T), 0.02, 0.3)

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    assert delta["fused_launches"] == 1 and delta["fused_tasks"] == 63


@pytest.mark.gpu
@pytest.mark.parametrize("n,dt", [(30, np.float64), (301, np.float64), (64, np.float32)])
def test_stencil_fused_is_bit_identical(n, dt):
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import stencil_init, stencil_run

    def prog():
        g = stencil_init(n, dt, xp=cn)
        w = stencil_run(g, 4)
        return g, w

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    g_np = stencil_init(n, dt, xp=np)
    w_np = stencil_run(g_np, 4)
    assert np.array_equal(fused[0], g_np) and np.array_equal(fused[1], w_np)
    # 4 iterations x [4 ADD + MULTIPLY]; the boundary writes of stencil_init may add one chain
    assert delta["fused_launches"] in (4, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dt", [(254, np.float64), (1002, np.float64), (254, np.float32),
                                  (510, np.float32), (126, np.float16)])
def test_stencil_tma_flavour_is_bit_identical(n, dt):
    """Even N: the row pitch is a multiple of 16 bytes, so the chain runs as TMA-staged tiles (one box
    with a halo per tile for the five shifted views, results stored straight from registers)."""
    import cunumeric_b200 as cn
    from cunumeric_b200.workloads import stencil_init, stencil_run

    def prog():
        g = stencil_init(n, dt, xp=cn)
        w = stencil_run(g, 5)
        return g, w

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    g_np = stencil_init(n, dt, xp=np)
    w_np = stencil_run(g_np, 5)
    if dt != np.float16:
        assert np.array_equal(fused[0], g_np) and np.array_equal(fused[1], w_np)
    assert delta["tma_launches"] == 5 and delta["renamed"] == 5


@pytest.mark.gpu
def test_tma_flavour_mixed_groups_and_outputs():
    """Two buffers (two tensor-map groups, one of them with shifted members), a dense third operand,
    scalars, a compare + where + convert in the chain, two stored outputs of different dtypes."""
    import cunumeric_b200 as cn

    rng = pu.rng_for("fusion-tma")
    a0 = rng.normal(size=(300, 520))
    b0 = rng.normal(size=(302, 524)).astype(np.float32)
    c0 = rng.normal(size=(298, 516))

    def prog():
        a, b, c = cn.array(a0), cn.array(b0), cn.array(c0)
        up, down = a[0:-2, 3:-1], a[2:, 1:-3]           # one group: shifts (0,3) and (2,1)
        bb = b[3:-1, 5:-3]                              # fp32 window, 4-byte misaligned
        t = (up + down) * 0.5 - c
        m = t > bb.astype(np.float64)
        r = cn.where(m, t, c * 2.0)
        q = r.astype(np.float32) + bb
        return r, q

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    assert delta["tma_launches"] >= 1
    exp_t = (a0[0:-2, 3:-1] + a0[2:, 1:-3]) * 0.5 - c0
    exp_r = np.where(exp_t > b0[3:-1, 5:-3].astype(np.float64), exp_t, c0 * 2.0)
    assert np.array_equal(fused[0], exp_r)


@pytest.mark.gpu
def test_inplace_views_and_hazards():
    import cunumeric_b200 as cn

    rng = pu.rng_for("fusion-hazards")
    a0 = rng.normal(size=(257, 130))
    b0 = rng.normal(size=(257, 130))

    def prog():
        a, b = cn.array(a0), cn.array(b0)
        a += b                      # in place: external read and final store of the same window
        a *= 2.0
        c = a - b
        a[1:] = a[:-1] + c[1:]      # shifted self-overlap: needs the copy + a flush
        d = a[:, ::2] * b[:, 1::2]  # strided inner dim -> strided kernel
        e = cn.where(d > 0.5, d, -d) + a[:, :65]
        a[0, :] = 1.0               # scalar write into a row between chains
        f = (a.T + 1.0) * b.T       # uniformly transposed operands
        row = a[3] * 2.0
        g = b + row                 # (130,) broadcast over rows
        return a, c, d, e, f, g

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    assert delta["fused_launches"] >= 3


@pytest.mark.gpu
def test_mixed_dtypes_conversions_and_dead_temporaries():
    import cunumeric_b200 as cn

    rng = pu.rng_for("fusion-mixed")
    x0 = rng.normal(size=70001).astype(np.float32)
    i0 = rng.integers(-50, 50, size=70001).astype(np.int32)

    def prog():
        x, i = cn.array(x0), cn.array(i0)
        y = x.astype(np.float64) * 3.0 + i           # CONVERT f32->f64, CONVERT i32->f64 inside
        m = (y > 1.0) & (i != 0)                      # compares -> bool, logical/bitwise on bool
        z = cn.where(m, y, 0.0)
        h = (x * x).astype(np.float16)
        k = i // 7 + i % 5 - (i << 1)
        q = cn.sqrt(cn.absolute(y)) + cn.exp(-cn.absolute(x))
        c = (x + 1j * x).astype(np.complex64) * (2 - 1j)
        s = float((z * 2.0).sum())                    # a reduction consumes a chain result
        t = z + s
        return y, m, z, h, k, q, c, t

    fused, eager, delta = run_both(prog)
    assert_identical(fused[:-1], eager[:-1])
    # `s` is folded INSIDE the fused kernel (map -> reduce fusion): another association order than the
    # standalone SCALAR_UNARY_RED kernel, so the sum — and what is computed from it — agrees within the
    # n * eps contract of floating-point reductions, not bit for bit
    n_eps = 70001 * np.finfo(np.float64).eps
    assert np.all(np.abs(fused[-1] - eager[-1]) <= n_eps * np.abs(fused[2] * 2.0).sum())
    assert delta["fused_launches"] >= 1 and delta["fused_reductions"] == 1


@pytest.mark.gpu
def test_rewriting_the_same_output_and_3d_fallback():
    import cunumeric_b200 as cn

    rng = pu.rng_for("fusion-rewrite")
    a0 = rng.normal(size=(5, 7, 9, 11)).astype(np.float32)

    def prog():
        a = cn.array(a0)
        out = cn.empty(a.shape, dtype=np.float32)
        cn.add(a, 1.0, out=out)
        cn.multiply(out, out, out=out)          # reads and rewrites the same window
        cn.subtract(out, a, out=out)
        v = a[:, 1:, :, 2:] * 2.0 + a[:, :-1, :, :-2]   # 4-D windows that do not merge to 2-D
        w = v - 1.0
        return out, v, w

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)


@pytest.mark.gpu
def test_default_mode_compiles_on_second_sighting(tmp_path, monkeypatch):
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    monkeypatch.setattr(fusion, "_CACHE_DIR", str(tmp_path))
    x0 = np.linspace(0, 1, 50000, dtype=np.float32)
    old = fusion.set_mode("1")
    try:
        results = []
        for rep in range(3):
            before = dict(fusion.stats)
            x = cn.array(x0)
            y = cn.tanh(x * 3.0 + 0.25) - x / 7.0
            results.append(np.array(y.__array__()))
            d = {k: fusion.stats[k] - before[k] for k in before}
            if rep == 0:
                assert d["fused_launches"] == 0 and d["replayed_tasks"] == 5
            else:
                assert d["fused_launches"] == 1, d
        assert np.array_equal(results[0], results[1]) and np.array_equal(results[1], results[2])
    finally:
        fusion.set_mode(old)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    # (nbytes, offset, rows, row_bytes, pitch) — every word width of the complement kernel
    (4096, 64, 10, 256, 400),        # 16-byte words
    (4096 + 8, 72, 10, 248, 392),    # 8-byte words
    (4000, 36, 9, 100, 404),         # 4-byte words
    (3001, 7, 11, 113, 257),         # bytes
    (1 << 20, 4104, 254, 4000, 4104),  # stencil-like: thin gaps, row-sized head and tail
    (5000, 0, 1, 5000, 5000),        # the window is the whole buffer: nothing to copy
    (5000, 100, 1, 300, 300),        # single row: head and tail only
])
def test_copy_complement_matches_numpy(case):
    """cnb_copy_complement (write-after-read renaming): dst <- src outside the pitched window, dst
    untouched inside it."""
    import ctypes

    import cunumeric_b200 as cn
    from cunumeric_b200 import _lib

    nbytes, offset, rows, row_bytes, pitch = case
    rng = pu.rng_for("complement")
    src_h = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    dst_h = np.full(nbytes, 0xEE, dtype=np.uint8)
    src, dst = cn.array(src_h), cn.array(dst_h)
    rt = cn.runtime
    _lib.check(rt.lib.cnb_copy_complement(dst._thunk.base.ptr, src._thunk.base.ptr, nbytes, offset, rows,
                                          row_bytes, pitch, rt.stream))
    got = np.array(dst)
    exp = src_h.copy()
    for r in range(rows):
        exp[offset + r * pitch: offset + r * pitch + row_bytes] = 0xEE
    assert np.array_equal(got, exp), np.flatnonzero(got != exp)[:8]


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float64, np.float32, np.float16, np.int32, np.int64],
                         ids=lambda d: np.dtype(d).name)
def test_map_reduce_fusion_against_the_oracle(dt):
    """A full reduction of a value the chain produces runs inside the fused kernel
    (fusion.capture_reduce).  Bars of the reduction suites: MAX / MIN / ALL / integer sums exact, floating
    sums within n * eps of the oracle (the reference's functors + its sequential fold)."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from oracle import ref

    dt = np.dtype(dt)
    rng = pu.rng_for("map-reduce", dt.name)
    n = 70001 if dt != np.float16 else 3001   # (an fp16 sum of 70 001 terms leaves the fp16 range)
    a0 = pu.make_input(dt, n, rng, "small")
    b0 = pu.make_input(dt, n, rng, "small")
    a, b = cn.array(a0), cn.array(b0)
    old = fusion.set_mode("always")
    try:
        before = dict(fusion.stats)
        l0 = cn.runtime.launch_count()
        delta = cn.sum(cn.absolute(a - b))
        cn.flush()
        assert cn.runtime.launch_count() - l0 == 1          # one kernel: no fill, no reduction launch
        t = a * b
        mx, mn = t.max(), t.min()
        pos = (t > 0).any()
        d = cn.dot(a, b)
        got = [np.array(v) for v in (delta, mx, mn, pos, d, t)]
        stats = {k: fusion.stats[k] - before[k] for k in before}
    finally:
        fusion.set_mode(old)
    assert stats["fused_reductions"] == 5 and stats["replayed_tasks"] == 0
    diff = ref.unary_op("ABSOLUTE", ref.binary_op("SUBTRACT", a0, b0))
    prod = ref.binary_op("MULTIPLY", a0, b0)
    assert np.array_equal(got[5], prod)
    exp_sum = ref.scalar_unary_red("SUM", diff)
    exp_dot = ref.scalar_unary_red("SUM", prod)
    if dt.kind in "iu":
        assert got[0] == exp_sum and got[4] == exp_dot
    else:
        eps = float(np.finfo(dt).eps)
        for g, e, terms in ((got[0], exp_sum, diff), (got[4], exp_dot, prod)):
            assert abs(float(g) - float(e)) <= n * eps * np.abs(terms.astype(np.float64)).sum()
    assert got[1] == ref.scalar_unary_red("MAX", prod) and got[2] == ref.scalar_unary_red("MIN", prod)
    assert bool(got[3]) == bool((prod > 0).any())


@pytest.mark.gpu
def test_jacobi_with_convergence_test_in_the_loop():
    """tests/integration/test_jacobi.py:23-57 of the reference: `delta = sum(abs(work - center))`
    inside the iteration.  The whole iteration — 4 ADD, MULTIPLY, SUBTRACT, ABSOLUTE, the SUM and the
    COPY back into the (renamed) grid — is one kernel."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    def program(xp, n, iters):
        grid = xp.zeros((n + 2, n + 2), np.float32)
        grid[:, 0] = -273.15
        grid[:, -1] = -273.15
        grid[-1, :] = -273.15
        grid[0, :] = 40.0
        center, north, east = grid[1:-1, 1:-1], grid[0:-2, 1:-1], grid[1:-1, 2:]
        west, south = grid[1:-1, 0:-2], grid[2:, 1:-1]
        deltas = []
        for _ in range(iters):
            average = center + north + east + west + south
            work = 0.2 * average
            deltas.append(xp.sum(xp.absolute(work - center)))
            center[:] = work
        return grid, [float(d) for d in deltas]

    old = fusion.set_mode("always")
    try:
        before = dict(fusion.stats)
        g, deltas = program(cn, 10, 2)           # the reference's own vector: 10 x 10, 2 iterations
        g2, deltas2 = program(cn, 300, 5)
        stats = {k: fusion.stats[k] - before[k] for k in before}
    finally:
        fusion.set_mode(old)
    g_np, deltas_np = program(np, 10, 2)
    assert np.array_equal(np.array(g), g_np) and np.allclose(deltas, deltas_np, rtol=1e-5)
    g2_np, deltas2_np = program(np, 300, 5)
    assert np.array_equal(np.array(g2), g2_np) and np.allclose(deltas2, deltas2_np, rtol=1e-5)
    # (the boundary fills of the two grids are single-task chains: those replay op-by-op)
    assert stats["fused_reductions"] == 7 and stats["renamed"] == 7 and stats["fused_launches"] >= 7


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_tma_flavour_random_shift_groups(seed):
    """Randomised layouts for the TMA-staged flavour: windows of one or two pitched buffers at random
    (row, column) shifts, random dtypes (1 / 2 / 4 / 8-byte elements), ragged extents, optional dense
    third operand and scalar, one or two outputs (dense or a pitched, misaligned window).  Fused ==
    op-by-op bit for bit, whichever flavour the launcher picks."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(1000 + seed)
    dt = np.dtype([np.float64, np.float32, np.float16, np.int32, np.int8, np.uint16][seed % 6])
    align = max(1, 16 // dt.itemsize)
    rows = int(rng.integers(16, 90))
    inner = int(rng.integers(64, 400))
    pad_r, pad_c = 4, int(rng.integers(1, 3)) * align        # room for shifts; pitch stays 16-byte multiple
    width = -(-(inner + 2 * pad_c) // align) * align
    shape = (rows + 2 * pad_r, width)

    def data(shp):
        if dt.kind == "f":
            return rng.normal(size=shp).astype(dt)
        return rng.integers(-20, 20, size=shp).astype(dt) if dt.kind == "i" else \
            rng.integers(0, 40, size=shp).astype(dt)

    a0, b0, c0 = data(shape), data(shape), data((rows, inner))
    shifts = [(int(rng.integers(0, 2 * pad_r + 1)), int(rng.integers(0, width - inner + 1)))
              for _ in range(int(rng.integers(2, 6)))]
    out_shift = (int(rng.integers(0, 2 * pad_r + 1)), int(rng.integers(0, width - inner + 1)))
    use_b, use_c, pitched_out = rng.random() < 0.5, rng.random() < 0.5, rng.random() < 0.5

    def prog():
        a, b, c = cn.array(a0), cn.array(b0), cn.array(c0)
        acc = None
        for k, (r, col) in enumerate(shifts):
            src = b if (use_b and k % 2) else a
            v = src[r:r + rows, col:col + inner]
            acc = v if acc is None else (acc + v if k % 3 else acc - v)
        if use_c:
            acc = acc * c
        res = acc + acc if dt.kind != "f" else acc * 0.5
        outs = [res]
        if pitched_out:
            dst = cn.array(b0)
            dst[out_shift[0]:out_shift[0] + rows, out_shift[1]:out_shift[1] + inner] = res
            outs.append(dst)
        if dt.kind == "f":
            outs.append(res > 0)
        return outs

    fused, eager, delta = run_both(prog)
    assert_identical(fused, eager)
    assert delta["fused_launches"] >= 1 and delta["tma_refused"] == 0
    # a window that does not start on a 16-byte boundary rules the vector flavour out: TMA tiles then
    misaligned = any((c * dt.itemsize) % 16 for _, c in shifts) or \
        (pitched_out and (out_shift[1] * dt.itemsize) % 16 != 0)
    # ... provided every pitched operand's row pitch is a multiple of 16 bytes (the dense `c` has
    # pitch = inner * itemsize)
    tma_able = not use_c or (inner * dt.itemsize) % 16 == 0
    if misaligned and tma_able and inner >= 64 and rows >= 16:
        assert delta["tma_launches"] >= 1, (delta, shifts)
