"""X-slab sharding of the grid across ranks: host-side placement logic only.

GPU g of G owns grid columns [floor(g*cols/G), floor((g+1)*cols/G)).  Entities live on the owner of their column,
subscribers on the owner of their centre's column; a query touches at most `halo` columns beyond its slab, so
one all-gather of border-column entity records per tick makes every rank's cell lists complete for its
subscribers.  The per-tick driver (`ShardedTick`) is engine-agnostic so the N>1 control flow is testable on
CPU with the gloo backend.
"""
import math

import numpy as np


def slab_columns(cols: int, world: int, rank: int):
    return (rank * cols) // world, ((rank + 1) * cols) // world


def owner_of_column(col, cols: int, world: int):
    """Inverse of slab_columns for arrays of columns (col < 0 -> rank 0)."""
    col = np.asarray(col, np.int64)
    # the owner is the largest g with floor(g*cols/world) <= col
    g = np.minimum(((col + 1) * world - 1) // cols, world - 1)
    g = np.where(col < 0, 0, g)
    return g.astype(np.int64)


def halo_columns(radius: float, grid_width: float) -> int:
    return max(1, int(math.ceil(radius / grid_width)))


def column_of(x, offx: float, w: float, cols: int):
    """Grid column of each x (placement only; -1 outside the world).  The engine recomputes cells on the GPU."""
    c = np.floor((np.asarray(x, np.float64) - offx) / w)
    return np.where((c >= 0) & (c < cols), c, -1).astype(np.int64)


class SlotTable:
    """Host-managed subscriber slots of every rank (the engine's slots are the host's to assign: chd_add_subscribers /
    chd_migrate_in name them).  Every process of a multi-rank host keeps the tables of ALL ranks (routing decisions are a pure
    function of the subscribers' positions, so they are computed identically everywhere: no host communication)."""

    def __init__(self, world: int):
        self.world = world
        self.owner = {}                              # subscriber -> rank
        self.slot = {}                               # subscriber -> local slot on its rank
        self.at = [dict() for _ in range(world)]     # per rank: slot -> subscriber
        self._free = [list() for _ in range(world)]
        self._used = [0] * world

    def take(self, rank: int) -> int:
        if self._free[rank]:
            return self._free[rank].pop()
        self._used[rank] += 1
        return self._used[rank] - 1

    def release(self, rank: int, slots):
        """Slots vacated by a tick's emigrants are reusable from the NEXT tick on (the engine frees them in that tick's update)."""
        self._free[rank] += list(slots)

    def place(self, j, rank: int) -> int:
        s = self.take(rank)
        self.owner[j], self.slot[j] = rank, s
        self.at[rank][s] = j
        return s

    def remove(self, j):
        rank, s = self.owner.pop(j), self.slot.pop(j)
        del self.at[rank][s]
        return rank, s


def plan_migrations(table: SlotTable, new_owner):
    """One tick's routing changes.  new_owner: dict / sequence subscriber -> rank that owns its centre's column NOW.
    Returns (arrivals, out_lists, in_calls, vacated):
      arrivals[rank]   = [(subscriber, slot)] first placements                      -> chd_add_subscribers
      out_lists[rank]  = [subscriber] emigrants in blob-record order                -> chd_migrate_out(vacated[rank])
      in_calls         = [(dst, src, first_index, [(subscriber, slot)])]            -> chd_migrate_in(dst: src, first_index, slots, conns)
      vacated[rank]    = the emigrants' old slots (release them after the tick)
    and applies the changes to `table`.  Records of one source rank are grouped by destination, so each (src, dst) pair is one
    contiguous run of src's blob."""
    world = table.world
    items = new_owner.items() if hasattr(new_owner, "items") else enumerate(new_owner)
    arrivals = [[] for _ in range(world)]
    moves = {}
    for j, r in items:
        r = int(r)
        if j not in table.owner:
            arrivals[r].append(j)
        elif table.owner[j] != r:
            moves.setdefault((table.owner[j], r), []).append(j)
    out_lists = [[] for _ in range(world)]
    in_calls = []
    for (a, b) in sorted(moves):
        in_calls.append([b, a, len(out_lists[a]), list(moves[(a, b)])])
        out_lists[a] += moves[(a, b)]
    vacated = [[table.slot[j] for j in out_lists[a]] for a in range(world)]
    for a in range(world):
        for j in out_lists[a]:
            table.remove(j)
    for call in in_calls:
        call[3] = [(j, table.place(j, call[0])) for j in call[3]]
    arrivals = [[(j, table.place(j, r)) for j in arrivals[r]] for r in range(world)]
    return arrivals, out_lists, [tuple(c) for c in in_calls], vacated
