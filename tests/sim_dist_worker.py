"""TEST INFRASTRUCTURE ONLY: the multi-GPU host logic (PartitionedArray, halo exchange queued between
deferred fused chains, buffer renaming) on CPU — one process per "GPU" under torchrun, the CUDA
library replaced by tests/sim_backend.SimLib and NCCL send/recv by gloo.  Run by
tests/test_partitioned_sim.py."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["CUNUMERIC_B200_MIN_PARTITION"] = "1"
os.environ.setdefault("CUNUMERIC_B200_HALO_OVERLAP", "1")   # opt-in path (fusion.Overlap): covered here
os.environ.setdefault("CNB_TMA_TR", "4")   # tile rows of 4: small grids still have interior tile rows

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sim_backend  # noqa: E402


class SimCommLib(sim_backend.SimLib):
    """SimLib + the cnb_comm_* entry points over gloo (grouped send/recv run at group end)."""

    def __init__(self) -> None:
        super().__init__()
        self._group = None

    def cnb_comm_unique_id(self, ident):
        return 0

    def cnb_comm_init(self, ident, world, rank):
        return 1

    def cnb_comm_group_start(self):
        self._group = []
        return 0

    def cnb_comm_send(self, comm, ptr, nbytes, peer, stream):
        raw = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(ptr), dtype=np.uint8).copy()
        self._group.append(("send", torch.from_numpy(raw), peer, None))
        return 0

    def cnb_comm_recv(self, comm, ptr, nbytes, peer, stream):
        self._group.append(("recv", torch.empty(nbytes, dtype=torch.uint8), peer, ptr))
        return 0

    def cnb_comm_group_end(self):
        reqs = []
        for kind, t, peer, ptr in self._group:
            reqs.append((dist.isend(t, peer) if kind == "send" else dist.irecv(t, peer), kind, t, ptr))
        for r, kind, t, ptr in reqs:
            r.wait()
            if kind == "recv":
                ctypes.memmove(ptr, t.numpy().ctypes.data, t.numel())
        self._group = None
        return 0


    # ---- collectives on dense buffers: gather the raw bytes over gloo, fold locally in rank order
    def _gather(self, src, nbytes):
        mine = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(src), dtype=np.uint8).copy()
        outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, torch.from_numpy(mine))
        return [t.numpy() for t in outs]

    def cnb_comm_allgather(self, comm, src, dst, nbytes, stream):
        for r, raw in enumerate(self._gather(src, nbytes)):
            ctypes.memmove(dst + r * nbytes, raw.ctypes.data, nbytes)
        return 0

    def cnb_comm_allreduce(self, comm, src, dst, count, dtype, red, stream):
        from cunumeric_b200.config import UnaryRedCode

        dt = np.dtype(sim_backend.DTYPES[dtype])
        parts = np.stack([raw.view(dt) for raw in self._gather(src, count * dt.itemsize)])
        fold = {int(UnaryRedCode.SUM): lambda p: p.sum(axis=0, dtype=dt),
                int(UnaryRedCode.PROD): lambda p: p.prod(axis=0, dtype=dt),
                int(UnaryRedCode.MAX): lambda p: p.max(axis=0), int(UnaryRedCode.MIN): lambda p: p.min(axis=0),
                int(UnaryRedCode.ALL): lambda p: p.all(axis=0), int(UnaryRedCode.ANY): lambda p: p.any(axis=0)}
        res = np.ascontiguousarray(fold[red](parts).astype(dt))
        ctypes.memmove(dst, res.ctypes.data, res.nbytes)
        return 0


PR, PC = 23, 9
# (view, first base row): the row offset decides which neighbours' rows a task needs
P_SUB = [(np.s_[1:-1, 1:-1], 1), (np.s_[0:-2, 1:-1], 0), (np.s_[2:, 1:-1], 2), (np.s_[1:-1, 0:-2], 1),
         (np.s_[1:-1, 2:], 1)]


def partitioned_program(seed: int, xp, steps: int = 40, free: bool = False):
    """free=False: outputs are always tiled like the interior rows (offset 1), so every operand is at
    most one row away from its owner — the halo depth, served by the ghost rows.  free=True: any view
    may come first, so operands can be two rows away from their owners and are fetched into temporary
    blocks (PartitionedArray._fetch_rows)."""
    rng = np.random.default_rng(seed)
    data = np.random.default_rng(3000 + seed)
    base = {k: xp.array(data.integers(-9, 10, size=(PR, PC)).astype(np.int64)) for k in "ab"}
    temps, scalars, checks = [], [], []
    ops = ["add", "subtract", "multiply", "maximum", "minimum"]

    def view(only_centre_rows=False):
        cands = [v for v, r0 in P_SUB if r0 == 1 or not only_centre_rows or free]
        return base["ab"[rng.integers(2)]][cands[rng.integers(len(cands))]]

    def operand(first=False):
        if temps and rng.random() < 0.5:
            return temps[rng.integers(len(temps))]
        return view(only_centre_rows=first)

    for step in range(steps):
        kind = rng.integers(9)
        op = getattr(xp, ops[rng.integers(len(ops))])
        if kind <= 2:
            x = operand(first=True)
            r = rng.random()
            y = operand() if r < 0.6 else (int(rng.integers(1, 5)) if r < 0.8 or not scalars
                                           else scalars[rng.integers(len(scalars))])
            temps.append(op(x, y))
        elif kind == 3:      # assignment into a view (any row offset) from an interior-tiled operand
            v, r0 = P_SUB[rng.integers(len(P_SUB))]
            base["ab"[rng.integers(2)]][v] = operand(first=True)
        elif kind == 4:      # in-place update of an interior-tiled view
            dst = view(only_centre_rows=True)
            dst += operand()
        elif kind == 5:      # full reduction (local partial + combine across the ranks)
            x = operand(first=True)
            if rng.random() < 0.5:
                x = op(x, operand())
            scalars.append(getattr(x, ["sum", "max", "min"][rng.integers(3)])())
        elif kind == 6:      # axis reduction broadcast back: axis 1 is local, axis 0 needs the combine
            x = operand(first=True)
            axis = int(rng.integers(2))
            r_ = getattr(x, ["sum", "max", "min"][rng.integers(3)])(axis=axis, keepdims=True)
            temps.append(xp.add(x, r_))
        elif kind == 7 and temps:
            temps.pop(rng.integers(len(temps)))
        elif kind == 8 and temps:
            checks.append(np.array(temps[rng.integers(len(temps))]))
    return checks + [np.array(v) for v in base.values()] + [np.array(t) for t in temps] + \
        [np.array(s_) for s_ in scalars]


def main() -> None:
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import cunumeric_b200 as cn
    from cunumeric_b200 import fusion
    from cunumeric_b200.distributed import PartitionedArray
    from cunumeric_b200.workloads import stencil_init, stencil_run

    rt = cn.runtime
    lib = SimCommLib()
    rt.lib, rt.stream, rt.device = lib, None, 0
    rt.rank, rt.world_size, rt.comm = rank, world, 1
    fusion._lookup = lambda sig: ("sim", sig)
    fusion._launch = sim_backend.make_fused_launcher(lib, np.random.default_rng(rank))
    mode = os.environ.get("SIM_FUSION", "always")
    fusion._mode = mode

    for n, iters in ((30, 3), (13, 5), (64, 4), (130, 6)):
        g = stencil_init(n, np.float64, xp=cn)
        assert isinstance(g._thunk, PartitionedArray)
        w = stencil_run(g, iters)
        g_np = stencil_init(n, np.float64, xp=np)
        w_np = stencil_run(g_np, iters)
        got_w, got_g = w.__array__(), g.__array__()
        assert np.array_equal(got_g, g_np), f"rank {rank}: grid mismatch n={n} mode={mode}: " \
            f"{np.argwhere(got_g != g_np)[:4].tolist()}"
        assert np.array_equal(got_w, w_np), f"rank {rank}: work mismatch n={n} mode={mode}: " \
            f"{np.argwhere(got_w != w_np)[:4].tolist()}"
    if mode == "always" and fusion._OVERLAP and fusion._RENAME and fusion._MAX_SEALED >= 1:
        # the n=130 grid has enough tile rows per rank: its halo exchanges ran BETWEEN the boundary
        # and the interior rows of the chain in front of them (fusion.Overlap)
        assert fusion.stats["overlapped_exchanges"] >= 4, fusion.stats
    # a second program: elementwise on shifted row views of a partitioned array, in-place update
    rng = np.random.default_rng(3)
    a0 = rng.normal(size=(41, 7))
    A = cn.array(a0)
    for _ in range(3):
        up = A[1:-1] + A[0:-2] + A[2:]      # output aligned with the middle rows: neighbours at +-1
        A[1:-1] = up * 0.25
        a0[1:-1] = (a0[1:-1] + a0[0:-2] + a0[2:]) * 0.25
    assert np.array_equal(A.__array__(), a0), f"rank {rank}: shifted-row update mismatch"
    # ---- reductions of partitioned arrays: local partial, then the combine across the ranks
    # (ncclAllReduce, or ncclAllGather + the library's own fold for floating MAX / MIN), then the fold
    # into the pre-filled result
    x = rng.integers(-9, 10, size=(37, 6)).astype(np.int64)
    X = cn.array(x)
    assert isinstance(X._thunk, PartitionedArray)
    assert int(X.sum()) == int(x.sum()) and int(X.max()) == int(x.max()) and int(X.min()) == int(x.min())
    assert int(X.sum(initial=100)) == int(x.sum()) + 100
    assert np.array_equal(X.sum(axis=0).__array__(), x.sum(axis=0))           # partitioned axis: combine
    assert np.array_equal(X.max(axis=0).__array__(), x.max(axis=0))
    assert np.array_equal(X.sum(axis=1).__array__(), x.sum(axis=1))           # local
    assert np.array_equal(X.min(axis=1).__array__(), x.min(axis=1))
    assert np.array_equal(X.sum(axis=0, keepdims=True).__array__(), x.sum(axis=0, keepdims=True))
    assert bool((X > -100).all()) and not bool((X > 100).any())
    assert int(cn.count_nonzero(X)) == int(np.count_nonzero(x))
    f = rng.normal(size=(29, 5))
    F = cn.array(f)
    assert float(F.max()) == f.max() and float(F.min()) == f.min()             # gather + fold path
    assert np.array_equal(F.max(axis=0).__array__(), f.max(axis=0))
    assert np.allclose(float(F.sum()), f.sum(), rtol=1e-13)
    assert np.allclose(F.sum(axis=0).__array__(), f.sum(axis=0), rtol=1e-13)
    # ---- operands whose rows are FARTHER than the halo depth from their owners: fetched into a
    # temporary block by one grouped send/recv (PartitionedArray._fetch_rows)
    h0 = rng.integers(-9, 10, size=(31, 5)).astype(np.int64)
    H = cn.array(h0)
    assert isinstance(H._thunk, PartitionedArray)
    assert np.array_equal((H[2:] + H[:-2]).__array__(), h0[2:] + h0[:-2])          # two rows apart
    assert np.array_equal((H[7:] * H[:-7]).__array__(), h0[7:] * h0[:-7])          # crosses whole blocks
    assert np.array_equal((H[0:-2, 1:-1] - H[2:, 1:-1]).__array__(), h0[0:-2, 1:-1] - h0[2:, 1:-1])
    assert np.array_equal(cn.maximum(H[:5], H[-5:]).__array__(), np.maximum(h0[:5], h0[-5:]))
    assert bool(cn.array_equal(H[3:], H[3:])) and not bool(cn.array_equal(H[3:], H[:-3]))
    T = H[5:] + 1                       # a temporary tiled like rows 5.. of H
    assert np.array_equal((T + H[:-5]).__array__(), h0[5:] + 1 + h0[:-5])
    H[:-4] = H[4:] + 0
    h0[:-4] = h0[4:] + 0
    assert np.array_equal(H.__array__(), h0)
    # ---- random programs on partitioned arrays (the same text on every rank and for NumPy): shifted
    # views, temporaries, assignments into views, in-place updates, reductions fed back as operands.
    # int64 wrap-around arithmetic is exact, so every association order gives the same bits.
    for seed in range(int(os.environ.get("SIM_PROGRAMS", "6"))):
        for free in (False, True):
            with np.errstate(over="ignore"):
                got = partitioned_program(seed, cn, free=free)
                exp = partitioned_program(seed, np, free=free)
            assert len(got) == len(exp)
            for i, (g, e) in enumerate(zip(got, exp)):
                assert g.shape == e.shape and np.array_equal(g, e), \
                    f"rank {rank}: program {seed} (free={free}) value {i} differs (mode {mode}): " \
                    f"{np.argwhere(g != e)[:3].tolist()}"
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
