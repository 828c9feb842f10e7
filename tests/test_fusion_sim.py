"""Randomised differential test of the lazy-fusion bookkeeping (capture, hazard rules, liveness /
dead-store elision, replay) on CPU: random NumPy programs over overlapping views are run through
cunumeric_b200 with a host-memory stand-in for the CUDA library (tests/sim_backend.py) and compared
with plain NumPy.  The stand-in executes a fused chain by reading every external input first and
writing the live outputs last, i.e. with the most aggressive cross-element reordering a real fused
kernel could exhibit, and poisons fresh allocations so that a wrongly elided store is caught."""
import numpy as np
import pytest

import sim_backend


@pytest.fixture(params=[0, 1, 2], ids=lambda s: f"sched{s}")
def sim(request):
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    rt = cn.runtime
    if rt.lib is not None:
        pytest.skip("a real device runtime is live in this process")
    fusion.drop_scalar_caches()
    fusion._plan_memo.clear()
    lib = sim_backend.SimLib()
    saved = (fusion._lookup, fusion._launch, fusion._mode)
    rt.lib, rt.stream, rt.device = lib, None, 0
    fusion._lookup = lambda sig: ("sim", sig)
    fusion._launch = sim_backend.make_fused_launcher(lib, np.random.default_rng(request.param))
    fusion._mode = "always"
    try:
        yield lib
    finally:
        fusion._chain = fusion._Chain()
        fusion._queue.clear()
        fusion._plan_memo.clear()
        fusion._lookup, fusion._launch, fusion._mode = saved
        fusion.drop_scalar_caches()
        rt._free_blocks.clear()
        rt._cached_bytes = 0
        rt._cache_limit = None
        rt.lib, rt.stream, rt.device = None, None, -1


R, C = 9, 11
SUB = [np.s_[1:-1, 1:-1], np.s_[0:-2, 1:-1], np.s_[2:, 1:-1], np.s_[1:-1, 0:-2], np.s_[1:-1, 2:]]
FULL = [np.s_[:, :], np.s_[::-1, :], np.s_[:, ::-1]]
BINOPS = ["add", "subtract", "multiply", "maximum", "minimum"]


def run_program(seed: int, xp, steps: int = 60):
    """The same random program text for xp = numpy and xp = cunumeric_b200."""
    rng = np.random.default_rng(seed)
    data = np.random.default_rng(1000 + seed)
    base = {k: xp.array(data.normal(size=(R, C))) for k in "abc"}
    temps = {}
    checks = []

    def pick_operand(shape_kind):
        names = [k for k, v in temps.items() if v[0] == shape_kind]
        if names and rng.random() < 0.6:
            return temps[names[rng.integers(len(names))]][1]
        b = base["abc"[rng.integers(3)]]
        views = SUB if shape_kind == "sub" else FULL
        return b[views[rng.integers(len(views))]]

    for step in range(steps):
        kind = rng.integers(8)
        shape_kind = "sub" if rng.random() < 0.6 else "full"
        op = getattr(xp, BINOPS[rng.integers(len(BINOPS))])
        if kind <= 2:      # new temporary from two operands (arrays, views or a scalar)
            x = pick_operand(shape_kind)
            y = pick_operand(shape_kind) if rng.random() < 0.7 else float(rng.integers(1, 5))
            temps[f"t{step}"] = (shape_kind, op(x, y))
        elif kind == 3:    # assignment of a temporary / view into a view of a base array
            dst = base["abc"[rng.integers(3)]]
            views = SUB if shape_kind == "sub" else FULL
            dst[views[rng.integers(len(views))]] = pick_operand(shape_kind)
        elif kind == 4:    # in-place update of a view (may overlap its operand)
            dst = base["abc"[rng.integers(3)]]
            views = SUB if shape_kind == "sub" else FULL
            v = dst[views[rng.integers(len(views))]]
            v += pick_operand(shape_kind)
        elif kind == 5:    # where / compare / convert chain
            x, y = pick_operand(shape_kind), pick_operand(shape_kind)
            t = xp.where(x > y, x, -y)
            temps[f"t{step}"] = (shape_kind, t.astype(np.float32).astype(np.float64))
        elif kind == 6 and temps:   # a temporary dies
            temps.pop(list(temps)[rng.integers(len(temps))])
        elif kind == 7 and temps:   # the host looks at a value
            name = list(temps)[rng.integers(len(temps))]
            checks.append(np.array(temps[name][1]))
    final = [np.array(v) for v in base.values()] + [np.array(v[1]) for v in temps.values()]
    return checks + final


@pytest.mark.parametrize("seed", range(40))
def test_random_programs_match_numpy(sim, seed):
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    before = dict(fusion.stats)
    got = run_program(seed, cn)
    exp = run_program(seed, np)
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.shape == e.shape and g.dtype == e.dtype, (seed, i)
        assert np.array_equal(g, e, equal_nan=True), (seed, i, np.argwhere(g != e)[:3])
    assert fusion.stats["captured"] > before["captured"]


def test_sim_really_fuses_and_elides(sim):
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    a = cn.array(np.arange(12.0).reshape(3, 4))
    before = dict(fusion.stats)
    r = (a * 2.0 + 1.0) * (a - 3.0)      # 4 tasks, 3 dead temporaries
    out = np.array(r)
    assert np.array_equal(out, (np.arange(12.0).reshape(3, 4) * 2 + 1) * (np.arange(12.0).reshape(3, 4) - 3))
    d = {k: fusion.stats[k] - before[k] for k in before}
    assert d["captured"] == 4 and sim.fused_launches == 1


def _np_stencil(n, iters):
    grid = np.zeros((n + 2, n + 2))
    grid[:, 0] = grid[:, -1] = grid[-1, :] = -273.15
    grid[0, :] = 40.0
    work = None
    for _ in range(iters):
        c, no, e = grid[1:-1, 1:-1], grid[0:-2, 1:-1], grid[1:-1, 2:]
        w, so = grid[1:-1, 0:-2], grid[2:, 1:-1]
        work = 0.2 * (c + no + e + w + so)
        c[:] = work
    return grid, work


def test_stencil_runs_as_one_renamed_kernel_per_iteration(sim):
    """examples/stencil.py through the lazy layer: `center[:] = work` joins the chain of the four
    ADDs and the MULTIPLY (write-after-read renaming), chains are launched one iteration late, by
    which time `average` and `work` of that iteration are dead and never stored."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from cunumeric_b200.workloads import stencil_init, stencil_run

    n, iters = 12, 7
    grid = stencil_init(n)
    cn.flush()
    before = dict(fusion.stats)
    fused0, comp0 = sim.fused_launches, getattr(sim, "complement_copies", 0)
    sim.stored_outputs = []
    work = stencil_run(grid, iters)
    got_grid, got_work = np.array(grid), np.array(work)
    exp_grid, exp_work = _np_stencil(n, iters)
    assert np.array_equal(got_grid, exp_grid) and np.array_equal(got_work, exp_work)
    d = {k: fusion.stats[k] - before[k] for k in before}
    assert sim.fused_launches - fused0 == iters          # one kernel per iteration
    assert d["renamed"] == iters and sim.complement_copies - comp0 == iters
    assert d["captured"] == 6 * iters and d["replayed_tasks"] == 0
    # stores: the renamed interior every iteration, `work` only for the last one (still observable)
    # -> every other intermediate (4 per iteration + `work` x (iters - 1)) is elided or kept in registers
    assert sim.stored_outputs[-iters:] == [1] * (iters - 1) + [2]


def test_deferred_chain_keeps_a_store_a_younger_chain_reads(sim):
    """A temporary that dies after a YOUNGER pending chain captured it as an input must still be
    stored by the older chain (DeviceBuffer.readers)."""
    import cunumeric_b200 as cn

    a = cn.array(np.arange(20.0).reshape(4, 5))
    t = a * 2.0                   # chain 1 (open)
    v = a[1:3, :]
    v[:] = t[0:2, :] + 1.0        # different shape -> chain 1 sealed, chain 2 reads t through a view
    del t                         # no Store onto t's buffer is left
    exp = np.arange(20.0).reshape(4, 5)
    exp[1:3, :] = (exp * 2.0)[0:2, :] + 1.0
    assert np.array_equal(np.array(a), exp)


def test_opaque_steps_keep_program_order(sim):
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    a = cn.array(np.ones((3, 4)))
    log = []
    b = a + 1.0
    fusion.enqueue(lambda: log.append(("step", float(np.array(b).sum()))))   # reads b when it RUNS
    c = b * 3.0
    assert log == [] or log == [("step", 24.0)]
    assert np.array_equal(np.array(c), np.full((3, 4), 6.0))
    assert log == [("step", 24.0)]


def test_map_reduce_joins_the_chain(sim):
    """sum(abs(a - b)) (the convergence test of test_jacobi.py) and dot(a, b): the reduction is the
    chain's last task, the mapped temporaries never reach memory."""
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion

    rng = np.random.default_rng(0)
    a0, b0 = rng.normal(size=(13, 7)), rng.normal(size=(13, 7))
    a, b = cn.array(a0), cn.array(b0)
    before = dict(fusion.stats)
    fused0, launches0 = sim.fused_launches, sim.launches
    delta = cn.sum(cn.absolute(a - b))
    got = float(delta)
    assert np.isclose(got, np.abs(a0 - b0).sum(), rtol=1e-12)
    d = {k: fusion.stats[k] - before[k] for k in before}
    assert d["fused_reductions"] == 1 and sim.fused_launches - fused0 == 1
    assert sim.launches - launches0 == 1                 # no fill, no separate reduction launch
    assert sim.stored_outputs[-1] == 1                   # only the 1-element result
    # max with `initial`, and a mapped array that stays observable next to its reduction
    t = a * b
    m = t.max(initial=0.5)
    assert float(m) == max(0.5, (a0 * b0).max()) and np.array_equal(np.array(t), a0 * b0)
    x0, y0 = rng.normal(size=257), rng.normal(size=257)
    assert np.isclose(float(cn.dot(cn.array(x0), cn.array(y0))), np.dot(x0, y0), rtol=1e-12)


def test_user_zero_d_arrays_own_their_buffer(sim):
    """`cn.array(2.0)` must not be a window onto the shared constant cache — an
    in-place write to it would change what every later `arr * 2.0` reads."""
    import cunumeric_b200 as cn

    x = cn.array(2.0)
    assert not x._thunk.base.buffer.shared and x._thunk.host_scalar is None
    x += 1.0
    assert float(np.array(x)) == 3.0
    assert np.array_equal(np.array(cn.array(np.ones(4)) * 2.0), np.full(4, 2.0))
    y = cn.asarray(np.float64(5.0))
    y.fill(7.0)
    assert float(np.array(y)) == 7.0
    assert np.array_equal(np.array(cn.array(np.ones(3)) * 5.0), np.full(3, 5.0))
    # +0.0 and -0.0 are different constants: the cache key carries the sign
    from cunumeric_b200._ufunc.ufunc import binary_ufunc

    a = binary_ufunc._weak_scalar(0.0, np.dtype(np.float32))
    b = binary_ufunc._weak_scalar(-0.0, np.dtype(np.float32))
    assert a is not b
    assert not np.signbit(a._thunk.host_scalar) and np.signbit(b._thunk.host_scalar)


def test_overlap_split_dependence_analysis(sim):
    """fusion._overlap_split: which tile rows of a chain a queued halo exchange depends on."""
    from cunumeric_b200 import fusion
    from cunumeric_b200.store import Store

    tr = fusion.TMA_TR
    rows, cols, item = 40 * tr + 2, 64, 8          # local block: 1 ghost row + owned rows + 1 ghost row
    pitch = cols * item
    buf = Store.empty((rows, cols), np.float64)
    centre = buf.slice(0, slice(1, rows - 1)).slice(1, slice(1, cols - 1))
    w = fusion._Window(centre)
    geo = {id(buf.buffer): (w, w.offset, rows - 2, (cols - 2) * item, pitch)}
    send = [(1 * pitch, 2 * pitch), ((rows - 2) * pitch, (rows - 1) * pitch)]
    recv = [(0, pitch), ((rows - 1) * pitch, rows * pitch)]
    ov = fusion.Overlap(buf.buffer, send, recv, lambda stream: None)
    tiles_y = -(-(rows - 2) // tr)
    args = (cols - 2, rows - 2, [pitch], tiles_y)
    assert fusion._overlap_split(ov, [w], geo, *args) == (1, 1)
    # a halo of tr + 1 rows reaches into the second tile row
    deep = fusion.Overlap(buf.buffer, [(pitch, (tr + 2) * pitch)], [], lambda stream: None)
    assert fusion._overlap_split(deep, [w], geo, *args) == (2, 0)
    # not renamed: the chain's reads and the exchange's writes share a block
    assert fusion._overlap_split(ov, [w], None, *args) is None
    assert fusion._overlap_split(ov, [w], {}, *args) is None
    # the chain does not write the exchanged buffer / writes it through another window as well
    other = fusion._Window(Store.empty((rows - 2, cols - 2), np.float64))
    assert fusion._overlap_split(ov, [other], geo, *args) is None
    w2 = fusion._Window(buf.slice(0, slice(0, rows - 2)).slice(1, slice(0, cols - 2)))
    assert fusion._overlap_split(ov, [w, w2], geo, cols - 2, rows - 2, [pitch, pitch], tiles_y) is None
    # too few tile rows for an interior worth the extra launch
    assert fusion._overlap_split(ov, [w], geo, cols - 2, rows - 2, [pitch], 4) is None
    # an exchange that touches none of the written rows does not depend on the chain
    far = fusion.Overlap(buf.buffer, [], [(0, pitch)], lambda stream: None)
    assert fusion._overlap_split(far, [w], geo, *args) is None


def test_multi_axis_reductions_run_axis_by_axis(sim):
    """sum / prod / max / min / all / any / count_nonzero over SEVERAL axes (the reference raises
    NotImplementedError, deferred.py:3259-3262): one UNARY_RED per axis, innermost first."""
    import cunumeric_b200 as cn

    rng = np.random.default_rng(11)
    a = rng.integers(-4, 5, size=(3, 4, 5, 6)).astype(np.int64)
    A = cn.array(a)
    for axes in ((0, 1), (1, 3), (0, 2, 3), (-1, 0), (2, 1)):
        for keepdims in (False, True):
            assert np.array_equal(A.sum(axis=axes, keepdims=keepdims).__array__(),
                                  a.sum(axis=axes, keepdims=keepdims))
            assert np.array_equal(A.max(axis=axes, keepdims=keepdims).__array__(),
                                  a.max(axis=axes, keepdims=keepdims))
            assert np.array_equal(A.min(axis=axes, keepdims=keepdims).__array__(),
                                  a.min(axis=axes, keepdims=keepdims))
        assert np.array_equal(A.prod(axis=axes).__array__(), a.prod(axis=axes))
        assert np.array_equal((A > 0).all(axis=axes).__array__(), (a > 0).all(axis=axes))
        assert np.array_equal((A > 3).any(axis=axes).__array__(), (a > 3).any(axis=axes))
        assert np.array_equal(cn.count_nonzero(A, axis=axes).__array__(), np.count_nonzero(a, axis=axes))
    assert np.array_equal(A.sum(axis=(0, 2), initial=7).__array__(), a.sum(axis=(0, 2), initial=7))
    assert np.array_equal(A.sum(axis=(1, 2), dtype=np.float64).__array__(),
                          a.sum(axis=(1, 2), dtype=np.float64))
    out = cn.empty((3, 6), dtype=np.int64)
    assert A.sum(axis=(1, 2), out=out) is out
    assert np.array_equal(out.__array__(), a.sum(axis=(1, 2)))
    with pytest.raises(ValueError):
        A.sum(axis=(1, 2), out=cn.empty((3, 5), dtype=np.int64))
    with pytest.raises(ValueError):
        A.sum(axis=(1, 1))
    # all axes named explicitly: the scalar reduction, as before
    assert int(A.sum(axis=(0, 1, 2, 3))) == int(a.sum())


def test_tma_launch_runs_a_queued_exchange_between_boundary_and_interior_tiles(sim, monkeypatch):
    """Host side of fusion._launch_tma with an Overlap (opt-in, CUNUMERIC_B200_HALO_OVERLAP=1): the
    sequence of calls on the two streams, and the tile rows each launch covers.  The device side (the
    generated kernel's ty_split / ty_skip mapping) is covered by tests/dist_worker.py on GPUs."""
    import ctypes

    from cunumeric_b200 import fusion
    from cunumeric_b200.config import BinaryOpCode
    from cunumeric_b200.runtime import runtime
    from cunumeric_b200.store import Store

    tr, tc = fusion.TMA_TR, fusion.TMA_TC
    rows, cols, item = 24 * tr + 2, 4 * tc + 2, 8
    pitch = cols * item
    grid = Store.empty((rows, cols), np.float64)
    centre = grid.slice(0, slice(1, rows - 1)).slice(1, slice(1, cols - 1))
    north = grid.slice(0, slice(0, rows - 2)).slice(1, slice(1, cols - 1))
    out_w, in_w = [fusion._Window(centre)], [fusion._Window(centre), fusion._Window(north)]
    shape = centre.shape
    sig = (((11, False), (11, False)), (("B", int(BinaryOpCode.ADD), 0, (0, 1), 2, 11),), ((2, 11),))
    dims = fusion._canonical(shape, [w.strides for w in out_w + in_w])
    lay, groups = fusion._tma_layout(sig, shape, out_w, in_w, dims)
    geo = fusion._tma_geometry(sig, lay)
    tail_cls = fusion._tma_params_type(len(geo["groups"]), 1, 0)
    monkeypatch.setattr(fusion, "_lookup_tma", lambda s, l: (1, geo, tail_cls, None))

    log = []

    def launch(kern, ops, ng, tail_ref, tail_bytes, smem, num_tiles, elements, algo, ntasks, cps, stream):
        t = ctypes.cast(tail_ref, ctypes.POINTER(tail_cls)).contents
        assert t.num_tiles == num_tiles
        log.append(("launch", stream, t.num_tiles // t.tiles_x, t.ty_split, t.ty_skip))
        return 0

    lib = sim
    lib.cnb_launch_fused_tma = launch
    lib.cnb_stream_create = lambda: 777
    lib.cnb_event_create = lambda: 55
    lib.cnb_event_record = lambda ev, stream: log.append(("record", stream)) or 0
    lib.cnb_stream_wait_event = lambda stream, ev: log.append(("wait", stream)) or 0
    monkeypatch.setattr(runtime, "_comm_stream", None)
    monkeypatch.setattr(runtime, "_event_pool", [])

    renamed = {id(grid.buffer): (out_w[0], out_w[0].offset, rows - 2, (cols - 2) * item, pitch)}
    ptrs = [grid.buffer.ptr + w.offset for w in out_w + in_w]
    tiles_y = -(-(rows - 2) // tr)
    inner, n_rows = dims[1][0], dims[0][0]
    row_st = dims[0][1]
    committed = []

    def run(overlap):
        del log[:], committed[:]
        ok = fusion._launch_tma(sig, lay, groups, inner, n_rows, row_st, out_w, in_w, ptrs, 0, 1,
                                lambda: committed.append(len(log)), renamed, overlap)
        assert ok and committed, "the launch must commit the renamed blocks itself"
        return list(log)

    # no exchange queued behind the chain: one launch over every tile row
    assert run(None) == [("launch", runtime.stream, tiles_y, tiles_y, 0)]
    # an exchange of the first / last owned row: boundary tile rows, exchange on the communication
    # stream behind an event, interior next to it, compute stream waits for the exchange at the end
    send = [(1 * pitch, 2 * pitch), ((rows - 2) * pitch, (rows - 1) * pitch)]
    recv = [(0, pitch), ((rows - 1) * pitch, rows * pitch)]
    ov = fusion.Overlap(grid.buffer, send, recv, lambda stream: log.append(("exchange", stream)))
    seq = run(ov)
    assert seq == [("launch", runtime.stream, 2, 1, tiles_y - 2),      # tile rows 0 and tiles_y - 1
                   ("record", runtime.stream), ("wait", 777), ("exchange", 777), ("record", 777),
                   ("launch", runtime.stream, tiles_y - 2, 0, 1),      # tile rows 1 .. tiles_y - 2
                   ("wait", runtime.stream)]
    assert ov.done and committed == [1], "blocks are adopted right after the boundary launch"
    # first rank: nothing to exchange above -> only the last tile row is boundary
    ov = fusion.Overlap(grid.buffer, send[1:], recv[1:], lambda stream: log.append(("exchange", stream)))
    seq = run(ov)
    assert seq[0] == ("launch", runtime.stream, 1, 0, tiles_y - 1)
    assert seq[5] == ("launch", runtime.stream, tiles_y - 1, 0, 0)
    # an exchange that is already done (or a chain that cannot run around it) changes nothing
    ov.done = True
    assert run(ov) == [("launch", runtime.stream, tiles_y, tiles_y, 0)]


def run_reduction_program(seed: int, xp, steps: int = 50):
    """Random int64 programs (wrap-around arithmetic is exact, so any association order gives the
    same bits) that mix elementwise tasks on views, in-place updates and reductions — full ones
    (which may join the open chain as its last task) and axis ones — whose results feed later tasks."""
    rng = np.random.default_rng(seed)
    data = np.random.default_rng(2000 + seed)
    base = {k: xp.array(data.integers(-9, 10, size=(R, C)).astype(np.int64)) for k in "abc"}
    temps, scalars, checks = {}, [], []
    ops = ["add", "subtract", "multiply", "maximum", "minimum"]

    def pick(shape_kind):
        names = [k for k, v in temps.items() if v[0] == shape_kind]
        if names and rng.random() < 0.6:
            return temps[names[rng.integers(len(names))]][1]
        b = base["abc"[rng.integers(3)]]
        views = SUB if shape_kind == "sub" else FULL
        return b[views[rng.integers(len(views))]]

    for step in range(steps):
        kind = rng.integers(10)
        shape_kind = "sub" if rng.random() < 0.6 else "full"
        op = getattr(xp, ops[rng.integers(len(ops))])
        if kind <= 2:
            x = pick(shape_kind)
            r = rng.random()
            if r < 0.5:
                y = pick(shape_kind)
            elif r < 0.75 or not scalars:
                y = int(rng.integers(1, 5))
            else:
                y = scalars[rng.integers(len(scalars))]        # a reduction result as 0-d operand
            temps[f"t{step}"] = (shape_kind, op(x, y))
        elif kind == 3:
            dst = base["abc"[rng.integers(3)]]
            views = SUB if shape_kind == "sub" else FULL
            dst[views[rng.integers(len(views))]] = pick(shape_kind)
        elif kind == 4:
            dst = base["abc"[rng.integers(3)]]
            views = SUB if shape_kind == "sub" else FULL
            v = dst[views[rng.integers(len(views))]]
            v += pick(shape_kind)
        elif kind == 5:       # full reduction, often of a value the open chain has just produced
            x = pick(shape_kind)
            if rng.random() < 0.6:
                x = op(x, pick(shape_kind))
            red = ["sum", "max", "min"][rng.integers(3)]
            scalars.append(getattr(x, red)())
        elif kind == 6:       # axis reduction, broadcast back along the reduced axis
            x = pick(shape_kind)
            axis = int(rng.integers(2))
            red = ["sum", "max", "min"][rng.integers(3)]
            r_ = getattr(x, red)(axis=axis, keepdims=True)
            temps[f"t{step}"] = (shape_kind, xp.add(x, r_))
        elif kind == 7 and temps:
            temps.pop(list(temps)[rng.integers(len(temps))])
        elif kind == 8 and temps:
            name = list(temps)[rng.integers(len(temps))]
            checks.append(np.array(temps[name][1]))
        elif kind == 9 and scalars:
            checks.append(np.array(scalars[rng.integers(len(scalars))]))
    final = [np.array(v) for v in base.values()] + [np.array(v[1]) for v in temps.values()] + \
            [np.array(s) for s in scalars]
    return checks + final


@pytest.mark.parametrize("seed", range(30))
def test_random_programs_with_reductions_match_numpy(sim, seed):
    import cunumeric_b200 as cn

    with np.errstate(over="ignore"):
        got = run_reduction_program(seed, cn)
        exp = run_reduction_program(seed, np)
    assert len(got) == len(exp)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.shape == e.shape and g.dtype == e.dtype, (seed, i, g.shape, e.shape, g.dtype, e.dtype)
        assert np.array_equal(g, e), (seed, i, np.argwhere(g != e)[:3])
