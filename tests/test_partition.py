"""CPU tests of the multi-GPU host logic: partitioning, induced (aligned) partitions of views, the
halo / fetch plans — and, over a real 2-process gloo group, that executing a plan puts the right
rows into every rank's ghost region (the same plans drive NCCL send/recv on the GPUs)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from cunumeric_b200.partition import RowPartition, Transfer, halo_bytes, plan_fetch, plan_halo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_even_partition_covers_all_rows():
    for n in (0, 1, 7, 8, 9, 40002, 32768):
        for w in (1, 2, 3, 4, 8):
            p = RowPartition.even(n, w)
            assert p.starts[0] == 0 and p.starts[-1] == n and p.world == w
            assert all(a <= b for a, b in zip(p.starts, p.starts[1:]))
            assert sum(p.count(r) for r in range(w)) == n
            assert max(p.count(r) for r in range(w)) - min(p.count(r) for r in range(w)) <= 1
            for row in range(0, n, max(1, n // 17)):
                lo, hi = p.bounds(p.owner(row))
                assert lo <= row < hi


def test_window_is_the_aligned_partition_of_a_view():
    """grid[1:-1] keeps every row with its owner: the stencil's `center` and the arrays computed
    from it share one tiling (the reference's add_alignment)."""
    p = RowPartition.even(40002, 8)
    c = p.window(1, 40001)
    assert c.nrows == 40000
    for r in range(8):
        lo, hi = p.bounds(r)
        clo, chi = c.bounds(r)
        assert (clo + 1, chi + 1) == (max(lo, 1), min(hi, 40001))
    assert p.window(0, 40002).same_as(p)
    assert p.window(5, 5).nrows == 0


def test_halo_plan_stencil_config():
    """C4: N=40000 fp64 on 8 GPUs -> each interior rank sends and receives 2 x 320 016 bytes."""
    p = RowPartition.even(40002, 8)
    plan = plan_halo(p, 1)
    assert len(plan) == 14  # 7 boundaries x 2 directions
    for t in plan:
        assert abs(t.src - t.dst) == 1 and t.nrows == 1
        lo, hi = p.bounds(t.src)
        assert lo <= t.row_lo < hi
    row_bytes = 40002 * 8
    assert halo_bytes(p, 1, row_bytes, 3) == (2 * row_bytes, 2 * row_bytes)
    assert halo_bytes(p, 1, row_bytes, 0) == (row_bytes, row_bytes)
    assert plan_halo(RowPartition.even(100, 1), 1) == []


def test_fetch_plan_is_consistent_and_minimal():
    p = RowPartition.even(10, 3)  # (0,3,6,10)
    need = [(0, 5), (2, 7), (9, 10)]
    plan = plan_fetch(p, need)
    got = {r: set() for r in range(3)}
    for t in plan:
        assert t.src != t.dst
        slo, shi = p.bounds(t.src)
        assert slo <= t.row_lo and t.row_hi <= shi
        got[t.dst] |= set(range(t.row_lo, t.row_hi))
    for r, (lo, hi) in enumerate(need):
        olo, ohi = p.bounds(r)
        assert got[r] == set(range(lo, hi)) - set(range(olo, ohi))


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    from cunumeric_b200.partition import RowPartition, plan_halo, plan_fetch

    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()

    def run(transfers, part, local, halo, dst, dst_row0):
        # the executor of cunumeric_b200.distributed._run_transfers, with gloo instead of NCCL
        lo, hi = part.bounds(rank)
        reqs = []
        for t in transfers:
            if t.src == rank:
                rows = torch.from_numpy(local[t.row_lo - (lo - halo): t.row_hi - (lo - halo)].copy())
                reqs.append(dist.isend(rows, t.dst))
            elif t.dst == rank:
                buf = torch.empty((t.nrows,) + local.shape[1:], dtype=torch.float64)
                reqs.append((dist.irecv(buf, t.src), buf, t))
        for r in reqs:
            if isinstance(r, tuple):
                r[0].wait()
                dst[r[2].row_lo - dst_row0: r[2].row_hi - dst_row0] = r[1].numpy()
            else:
                r.wait()

    n, cols, halo = 11, 5, 1
    full = np.arange(n * cols, dtype=np.float64).reshape(n, cols)
    part = RowPartition.even(n, world)
    lo, hi = part.bounds(rank)
    local = np.full((hi - lo + 2 * halo, cols), -1.0)
    local[halo: halo + hi - lo] = full[lo:hi]
    run(plan_halo(part, halo), part, local, halo, local, lo - halo)
    if lo > 0:
        assert np.array_equal(local[0], full[lo - 1]), (rank, local[0])
    if hi < n:
        assert np.array_equal(local[-1], full[hi]), (rank, local[-1])
    # one Jacobi-style step on the interior using the ghosts == the global computation
    glob = full.copy()
    glob[1:-1] = (full[:-2] + full[2:]) * 0.5
    a, b = max(lo, 1), min(hi, n - 1)
    mine = (local[a - 1 - (lo - halo): b - 1 - (lo - halo)] + local[a + 1 - (lo - halo): b + 1 - (lo - halo)]) * 0.5
    assert np.array_equal(mine, glob[a:b])
    # gather-by-plan: everyone fetches everything
    whole = np.full((n, cols), -1.0)
    whole[lo:hi] = full[lo:hi]
    run(plan_fetch(part, [(0, n)] * world), part, local, halo, whole, 0)
    assert np.array_equal(whole, full)
    # scalar-reduction combine: local partial + allreduce == global
    t = torch.tensor([full[lo:hi].sum()], dtype=torch.float64)
    dist.all_reduce(t)
    assert t.item() == full.sum()
    # arg-reduction combine: max of values, then min of candidate indices (lowest index wins ties)
    vals = np.array([3., 9., 1., 9., 2., 9., 0., 4., 9., 1., 5.])
    part1 = RowPartition.even(len(vals), world)
    l1, h1 = part1.bounds(rank)
    li = l1 + int(np.argmax(vals[l1:h1]))
    best = torch.tensor([vals[li]]); dist.all_reduce(best, op=dist.ReduceOp.MAX)
    cand = torch.tensor([li if vals[li] == best.item() else 2**62]); dist.all_reduce(cand, op=dist.ReduceOp.MIN)
    assert cand.item() == int(np.argmax(vals)) == 1
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


@pytest.mark.parametrize("world", [2, 3])
def test_plans_execute_correctly_over_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = 29600 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("ok") == world
