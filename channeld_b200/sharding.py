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


class ShardedTick:
    """One tick of the sharded pipeline on this rank.

    engine   : object with set_entities/export_border/import_halo/build (channeld_b200.engine.Engine, or a CPU
               stand-in in the gloo tests)
    gather   : callable(local_records [cap,2] -> all_records [world*cap,2]) — torch.distributed all_gather
    """

    def __init__(self, engine, rank, world, border_cap, gather):
        self.e, self.rank, self.world, self.cap, self.gather = engine, rank, world, border_cap, gather

    def step(self, x, z, records_local):
        self.e.set_entities(x, z)
        if self.world > 1:
            n = self.e.export_border(records_local, self.cap)
            allrec = self.gather(records_local)
            self.e.import_halo(allrec, self.cap * self.world, self.rank * self.cap, self.cap)
        else:
            n = 0
        self.e.build()
        return n
