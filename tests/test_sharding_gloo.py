"""N>1 control flow on CPU: X-slab placement + border all-gather + halo import, world_size 2 over gloo.

The engine is replaced by a CPU stand-in that implements the same four calls with the oracle (the checker);
the thing under test is the host logic in channeld_b200/sharding.py — slab ownership, the exported border set,
the skip-own-segment convention and that the union over ranks of per-subscriber visible sets equals the
single-rank answer."""
import os
import socket

import numpy as np
import pytest

from channeld_b200 import sharding, synth


def test_slab_columns_partition():
    for cols in (1, 2, 15, 64, 256):
        for world in (1, 2, 4, 8):
            if world > cols:
                continue
            ranges = [sharding.slab_columns(cols, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == cols
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0] and a[0] < a[1]
            owner = sharding.owner_of_column(np.arange(cols), cols, world)
            for r, (lo, hi) in enumerate(ranges):
                assert (owner[lo:hi] == r).all()
    assert sharding.slab_columns(15, 8, 7) == (13, 15) and sharding.slab_columns(15, 8, 0) == (0, 1)
    assert sharding.halo_columns(50, 2000) == 1 and sharding.halo_columns(2500, 2000) == 2


class CpuStandInEngine:
    """Same call surface as channeld_b200.engine.Engine for the sharded tick; cell ids come from the oracle."""

    def __init__(self, orc, og, wc, col_lo, col_hi, halo, gid):
        self.orc, self.og, self.wc = orc, og, wc
        self.col_lo, self.col_hi, self.halo, self.gid = col_lo, col_hi, halo, np.asarray(gid, np.uint32)

    def set_entities(self, x, z):
        ids = self.orc.cell_of(self.og, x, z)
        self.cell = np.where(ids == 0, 0xFFFFFFFF, ids - 65536).astype(np.uint32)
        self.halo_rec = np.zeros((0, 2), np.uint32)

    def export_border(self, records, cap):
        col = (self.cell % self.wc.cols).astype(np.int64)
        valid = self.cell != 0xFFFFFFFF
        interior = (col >= self.col_lo + self.halo) & (col + self.halo < self.col_hi)
        near_l, near_r = col < self.col_lo + self.halo, col + self.halo >= self.col_hi
        flag = valid & ~interior & ((near_l & (self.col_lo > 0)) | (near_r & (self.col_hi < self.wc.cols)) | (col < self.col_lo) | (col >= self.col_hi))
        sel = np.nonzero(flag)[0]
        assert len(sel) <= cap
        records[:] = 0xFFFFFFFF
        records[: len(sel), 0] = self.gid[sel]
        records[: len(sel), 1] = self.cell[sel]
        return len(sel)

    def import_halo(self, allrec, n, skip_first, skip_count):
        rec = np.asarray(allrec).reshape(-1, 2)[:n]
        keep = rec[:, 1] != 0xFFFFFFFF
        keep[skip_first:skip_first + skip_count] = False
        col = (rec[:, 1] % self.wc.cols).astype(np.int64)
        keep &= (col + self.halo >= self.col_lo) & (col < self.col_hi + self.halo)
        self.halo_rec = rec[keep]

    def build(self):
        own = self.cell != 0xFFFFFFFF
        cells = np.concatenate([self.cell[own], self.halo_rec[:, 1]])
        ids = np.concatenate([self.gid[own], self.halo_rec[:, 0]])
        order = np.lexsort((ids, cells))
        self.sorted_cell, self.sorted_id = cells[order], ids[order]

    def visible(self, cells):
        out = []
        for c in cells:
            lo, hi = np.searchsorted(self.sorted_cell, [c, c + 1])
            out.append(self.sorted_id[lo:hi])
        return np.concatenate(out) if out else np.zeros(0, np.uint32)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from tests import _oracle

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.load()
        wc = synth.scaled(synth.CONFIGS["benchmark"], 6000, 300)
        og = _oracle.make_grid(wc.offx, wc.offz, wc.w, wc.h, wc.cols, wc.rows, wc.server_cols, wc.server_rows)
        ex, ez = synth.entities(wc)
        ex2, ez2 = synth.move_entities(wc, ex, ez, 1, 900.0)  # big moves: some entities leave their owner's slab
        radius = 2500.0  # halo of 2 columns
        halo = sharding.halo_columns(radius, wc.w)
        lo, hi = sharding.slab_columns(wc.cols, world, rank)
        ent_col = sharding.column_of(ex, wc.offx, wc.w, wc.cols)
        mine = np.nonzero((ent_col >= lo) & (ent_col < hi))[0]
        cap = 4000
        eng = CpuStandInEngine(orc, og, wc, lo, hi, halo, mine)

        def gather(local):
            t = torch.from_numpy(local.view(np.int32).reshape(-1).copy())
            out = torch.empty(world * t.numel(), dtype=torch.int32)
            dist.all_gather_into_tensor(out, t)
            return out.numpy().view(np.uint32).reshape(-1, 2)

        tick = sharding.ShardedTick(eng, rank, world, cap, gather)
        rec = np.zeros((cap, 2), np.uint32)
        n_exported = tick.step(ex2[mine], ez2[mine], rec)
        # subscribers whose centre column (after the move) is in this slab are answered here
        conn, cx, cz, r = synth.subscribers(wc, ex2, ez2, radius)
        sub_col = sharding.column_of(cx, wc.offx, wc.w, wc.cols)
        mine_s = np.nonzero((sub_col >= lo) & (sub_col < hi))[0]
        want = orc.sphere_tick(og, ex2, ez2, cx[mine_s], cz[mine_s], r[mine_s])
        ok = True
        for k, j in enumerate(mine_s):
            cells = want["pair_cell"][want["pair_off"][k]:want["pair_off"][k + 1]] - 65536
            got = eng.visible(cells)
            ok &= np.array_equal(got, want["vis_entity"][want["vis_off"][k]:want["vis_off"][k + 1]])
        q.put((rank, bool(ok), int(n_exported), len(mine_s), int(len(eng.halo_rec))))
    finally:
        dist.destroy_process_group()


def test_sharded_tick_world2_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] > 0 and r[4] > 0 for r in res), res  # both ranks exported and adopted border entities
    assert sum(r[3] for r in res) > 250  # (nearly) every subscriber was answered by exactly one rank
