"""Row-block partitioning and exchange planning (pure host logic, no device access).

The reference leaves partitioning to legate.core: the output store is tiled over NUM_PROCS
processors and every input is *aligned* to that tiling (deferred.py:3152-3165 `add_alignment`);
shifted-slice operands then reach one row into the neighbour's tile and Legion issues the copies
(SURVEY §2.2, §3.5).  Here the same decisions are explicit and replicated on every rank (SPMD, one
process per GPU): arrays are split into contiguous blocks of rows, a view's tiling is the one its
base array induces, and every rank can compute — without communication — which rows each rank must
send to whom.  The executor (cunumeric_b200/distributed.py) turns a plan into grouped NCCL
send/recv calls; tests/test_partition.py runs the same plans over gloo on the CPU."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class RowPartition:
    """Contiguous row blocks: rank r owns rows [starts[r], starts[r+1]) of `nrows`."""

    nrows: int
    starts: Tuple[int, ...]  # world+1 non-decreasing entries, starts[0] == 0, starts[-1] == nrows

    @staticmethod
    def even(nrows: int, world: int) -> "RowPartition":
        """Equal split, like legate's default tiling of a 1-D launch space."""
        return RowPartition(nrows, tuple((r * nrows) // world for r in range(world + 1)))

    @property
    def world(self) -> int:
        return len(self.starts) - 1

    def bounds(self, rank: int) -> Tuple[int, int]:
        return self.starts[rank], self.starts[rank + 1]

    def count(self, rank: int) -> int:
        lo, hi = self.bounds(rank)
        return hi - lo

    def owner(self, row: int) -> int:
        if not (0 <= row < self.nrows):
            raise IndexError(row)
        for r in range(self.world):
            if self.starts[r] <= row < self.starts[r + 1]:
                return r
        raise AssertionError

    def window(self, start: int, stop: int) -> "RowPartition":
        """Partition induced on the view rows [start, stop) of this array: view row i is base row
        start+i and stays with its owner (the alignment rule)."""
        n = max(0, stop - start)
        starts = tuple(min(max(s - start, 0), n) for s in self.starts)
        return RowPartition(n, starts)

    def same_as(self, other: "RowPartition") -> bool:
        return self.nrows == other.nrows and self.starts == other.starts


@dataclass(frozen=True)
class Transfer:
    src: int
    dst: int
    row_lo: int  # global rows [row_lo, row_hi) of the SOURCE array
    row_hi: int

    @property
    def nrows(self) -> int:
        return self.row_hi - self.row_lo


def plan_fetch(owner: RowPartition, need: Sequence[Tuple[int, int]]) -> List[Transfer]:
    """All transfers so that every rank r holds base rows need[r] = [lo, hi) of an array whose rows
    are distributed by `owner`.  Rows a rank already owns are not transferred.  The list is the same
    on every rank and ordered (dst, src), so executing it in order is deadlock-free inside one
    grouped send/recv."""
    out: List[Transfer] = []
    for dst, (lo, hi) in enumerate(need):
        lo, hi = max(lo, 0), min(hi, owner.nrows)
        if hi <= lo:
            continue
        for src in range(owner.world):
            if src == dst:
                continue
            slo, shi = owner.bounds(src)
            a, b = max(lo, slo), min(hi, shi)
            if b > a:
                out.append(Transfer(src, dst, a, b))
    return out


def plan_halo(owner: RowPartition, depth: int) -> List[Transfer]:
    """Refresh `depth` ghost rows above and below every rank's block (the stencil halo)."""
    need = []
    for r in range(owner.world):
        lo, hi = owner.bounds(r)
        need.append((lo - depth, hi + depth) if hi > lo else (0, 0))
    return plan_fetch(owner, need)


def halo_bytes(owner: RowPartition, depth: int, row_bytes: int, rank: int) -> Tuple[int, int]:
    """(bytes sent, bytes received) by `rank` in one halo refresh."""
    sent = recv = 0
    for t in plan_halo(owner, depth):
        if t.src == rank:
            sent += t.nrows * row_bytes
        if t.dst == rank:
            recv += t.nrows * row_bytes
    return sent, recv
