"""N>1 host logic on CPU, world_size 2 over gloo (channeld_b200/sharding.py: the code tests/run_multigpu_parity.py and a sharded
host run around chd_tick_sharded): slab ownership, the exported border set and the skip-own-segment convention (the engine is
replaced by a CPU stand-in that implements export / import / build with the oracle, the checker), that the union over ranks of
per-subscriber visible sets equals the single-rank answer, and the subscriber routing plan (SlotTable / plan_migrations):
identical on every rank without communication, every subscriber on the owner of its column, state carried through blobs."""
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

        rec = np.zeros((cap, 2), np.uint32)
        # the sequence chd_tick_sharded runs inside the library: assign -> export border -> ONE all-gather -> import halo -> build
        eng.set_entities(ex2[mine], ez2[mine])
        n_exported = eng.export_border(rec, cap)
        eng.import_halo(gather(rec), cap * world, rank * cap, cap)
        eng.build()
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


def _migration_worker(rank, world, port, q):
    import pickle

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cols, n_sub = 15, 400
        rng = np.random.default_rng(11)  # same stream on every rank: routing needs no communication
        col = rng.integers(0, cols, n_sub)
        table = sharding.SlotTable(world)
        state = {}          # this rank's "engine": slot -> (subscriber, fan-out state token)
        plans = []
        for tick in range(8):
            if tick:
                col = np.clip(col + rng.integers(-2, 3, n_sub), 0, cols - 1)
            new_owner = sharding.owner_of_column(col, cols, world)
            arrivals, out_lists, in_calls, vacated = sharding.plan_migrations(table, new_owner)
            plans.append((arrivals, out_lists, in_calls, vacated))
            # what chd_migrate_out / the all-gather / chd_migrate_in do, with python objects: blob = emigrants' state in record order
            blob = [state.pop(s_) for s_ in vacated[rank]]
            assert [b[0] for b in blob] == out_lists[rank]
            blobs = [None] * world
            dist.all_gather_object(blobs, blob)
            for (dst, src, first, pairs_) in in_calls:
                if dst == rank:
                    for k, (j, s_) in enumerate(pairs_):
                        rec = blobs[src][first + k]
                        assert rec[0] == j and s_ not in state
                        state[s_] = rec
            for (j, s_) in arrivals[rank]:
                assert s_ not in state
                state[s_] = (j, "born@%d" % tick)
            for a in range(world):
                table.release(a, vacated[a])
            # invariants: every subscriber on the owner of its column, tables and engine agree, nobody lost or duplicated
            mine = {j for j in range(n_sub) if new_owner[j] == rank}
            assert {v[0] for v in state.values()} == mine == set(table.at[rank].values())
            assert all(table.at[rank][s_] == v[0] for s_, v in state.items())
            assert all(v[1].startswith("born@") for v in state.values())  # state travelled intact
        # the plan is the same on every rank
        digests = [None] * world
        dist.all_gather_object(digests, pickle.dumps(plans))
        q.put((rank, len(set(digests)) == 1, sum(len(p[1][rank]) for p in plans)))
    finally:
        dist.destroy_process_group()


def test_migration_plan_world2_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_migration_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert all(r[2] > 10 for r in res), res  # both ranks sent subscribers away


def test_plan_migrations_groups_records_by_destination():
    t = sharding.SlotTable(3)
    arrivals, out_lists, in_calls, vacated = sharding.plan_migrations(t, {10: 0, 11: 0, 12: 1, 13: 2})
    assert [len(a) for a in arrivals] == [2, 1, 1] and not any(out_lists) and not in_calls
    arrivals, out_lists, in_calls, vacated = sharding.plan_migrations(t, {10: 2, 11: 1, 12: 1, 13: 0})
    assert out_lists == [[11, 10], [], [13]]  # rank 0's records: destination 1 first, then destination 2
    assert [(d, s_, f, [j for j, _ in p]) for d, s_, f, p in in_calls] == [(1, 0, 0, [11]), (2, 0, 1, [10]), (0, 2, 0, [13])]
    assert vacated == [[1, 0], [], [0]] and t.owner == {10: 2, 11: 1, 12: 1, 13: 0}
    # a slot vacated this tick is not reused in the same tick
    assert t.slot[13] == 2 and t.slot[10] == 1 and t.slot[11] == 1
