"""Multi-GPU parity run (one rank per GPU, launched by tests/test_multigpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_multigpu_parity.py

Every rank owns an X-slab and drives the SHARDED engine through the library's own exchange (chd_comm_init /
chd_tick_sharded: one ncclAllGather per tick issued inside libchd_b200.so).  Entities and subscribers move for several ticks
and are RE-HOMED as they cross slab borders:
  * entities: the owner keeps exporting an entity that left its slab (visibility stays exact); chd_get_rehome tells the hosts,
    which hand it over for the next tick (sender drops, receiver adopts);
  * subscribers: the host routes every subscriber to the owner of its centre's column; chd_migrate_out / chd_migrate_in move the
    subscriptions + fan-out state inside the same all-gather.
Oracle: every rank also runs a plain single-GPU engine over the WHOLE world (itself parity-tested against the CPU oracle by
tests/test_gpu_parity.py) and compares, for EVERY local subscriber on EVERY tick: (channel, dist, interval, flags, lastFanOutTime,
lastMessageIndex) of every pair bit-identical, visible lists identical as sets of global ids, fan-out decisions identical.
The first tick is also checked against the CPU oracle directly."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from channeld_b200 import capi, engine, sharding, synth  # noqa: E402
from tests import _oracle  # noqa: E402

TICK = 33_000_000


def due_set(due, conn_of_slot):
    return {(int(conn_of_slot[d["sub"]]), int(d["channel_id"]), int(d["kind"]), int(d["n_selected"]), int(d["first_sel"]), int(d["last_sel"]),
             int(d["sel_hash"]), int(d["last_message_index"]), int(d["window_hi"])) for d in due}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    orc = _oracle.load()
    failures = 0
    n_ticks = int(os.environ.get("CHD_PARITY_TICKS", "6"))
    # the 2x2 grid gives every rank of a 2-GPU run a ONE-column slab (empty interior, as at 8 GPUs on 15 columns)
    # last field: exchange through the peer windows (stores over NVLink + flags, the default) or through one ncclAllGather per tick
    for name, n_ent, n_sub, radius, max_move, collective in (("benchmark", 120_000, 6_000, 50.0, 700.0, False), ("benchmark", 60_000, 3_000, 2500.0, 900.0, True),
                                                             ("benchmark", 60_000, 3_000, 2500.0, 900.0, False), ("handover", 200_000, 4_000, 50.0, 120.0, False),
                                                             ("2x2", 50_000, 2_000, 50.0, 400.0, False), ("2x2", 50_000, 2_000, 500.0, 300.0, True),
                                                             ("2x2", 50_000, 2_000, 500.0, 300.0, False)):
        wc = synth.scaled(synth.CONFIGS[name], n_ent, n_sub)
        if world > wc.cols:
            continue
        og = _oracle.make_grid(wc.offx, wc.offz, wc.w, wc.h, wc.cols, wc.rows, wc.server_cols, wc.server_rows)
        x, z = synth.entities(wc)
        halo = sharding.halo_columns(radius, wc.w)
        lo, hi = sharding.slab_columns(wc.cols, world, rank)
        ent_owner = sharding.owner_of_column(sharding.column_of(x, wc.offx, wc.w, wc.cols), wc.cols, world)
        mine = np.nonzero(ent_owner == rank)[0]
        cap = n_ent  # generous border capacity for the test
        stream = torch.cuda.Stream(device=dev)
        # ---- the sharded engine
        e = engine.Engine(wc.cfg(), n_ent + cap * world + 16, n_sub, device=local, max_visible=1 << 27, max_pairs=n_sub * 64 + 1024)
        e.set_stream(stream.cuda_stream)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(engine.Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        e.comm_init(uid.cpu().numpy().tobytes(), rank, world, halo, cap, migrate_subscribers=n_sub, migrate_pairs=n_sub * 16)
        info = e.comm_info()
        assert (info["col_lo"], info["col_hi"], info["world"]) == (lo, hi, world)
        if collective:
            e.use_collective(True)
        mode = e.exchange_mode()
        assert mode == (1 if collective else mode) and mode in (1, 2)
        if rank == 0:
            print("%s r=%g: exchange mode %d (%s)" % (name, radius, mode, "peer windows" if mode == 2 else "ncclAllGather"), flush=True)
        # ---- the single-GPU reference engine over the whole world (slot j = subscriber j)
        e1 = engine.Engine(wc.cfg(), n_ent, n_sub, device=local, max_visible=1 << 28, max_pairs=n_sub * 64 + 1024)
        e1.set_stream(stream.cuda_stream)
        conn, cx, cz, r = synth.subscribers(wc, x, z, radius)
        e1.set_subscribers(conn)
        # host-side routing (channeld_b200/sharding.py): subscriber j -> (owner rank, local slot); every process computes all ranks' tables
        table = sharding.SlotTable(world)
        last_owner = np.zeros(n_sub, np.int64)
        ring_state = None
        n_mig_total = n_rehome_total = 0
        with torch.cuda.stream(stream):
            for tick in range(n_ticks):
                t_ns = (tick + 1) * TICK
                if tick:
                    x, z = synth.move_entities(wc, x, z, tick, max_move)  # entities (and the subscribers standing on them) drift
                conn, cx, cz, r = synth.subscribers(wc, x, z, radius)
                col = sharding.column_of(cx, wc.offx, wc.w, wc.cols)
                new_owner = np.where(col >= 0, sharding.owner_of_column(col, wc.cols, world), last_owner)  # out of the world: stays where it was
                last_owner = new_owner
                # ---- subscriber routing / migration (identical decisions on every rank)
                arrivals, out_lists, in_calls, vacated = sharding.plan_migrations(table, new_owner)
                if out_lists[rank]:
                    e.migrate_out(np.array(vacated[rank], np.uint32))
                for (dst, src, first, pairs_) in in_calls:
                    if dst == rank:
                        e.migrate_in(src, first, np.array([s_ for _, s_ in pairs_], np.uint32), conn[[j for j, _ in pairs_]])
                if arrivals[rank]:
                    e.add_subscribers(np.array([s_ for _, s_ in arrivals[rank]], np.uint32), conn[[j for j, _ in arrivals[rank]]])
                n_mig_total += sum(len(v) for v in out_lists)
                # ---- update rings (global cells, same on every rank)
                ring_state, roff, rarr, rsnd, ridx, rcmi = synth.update_rings(wc, tick, t_ns, TICK, 6, n_sub, ring_len=32, state=ring_state)
                e.set_rings(roff, rarr, rsnd, ridx, rcmi)
                e1.set_rings(roff, rarr, rsnd, ridx, rcmi)
                # ---- the sharded tick
                e.set_entities(x[mine], z[mine])
                e.set_entity_ids(mine.astype(np.uint32))
                my_slots = np.array(sorted(table.at[rank]), np.uint32)
                my_subs = np.array([table.at[rank][int(s_)] for s_ in my_slots], np.int64)
                batch, keep = engine.make_batch(len(my_slots), sub=my_slots, sphere=(cx[my_subs], cz[my_subs], r[my_subs]))
                s = e.tick_sharded(batch, t_ns, capi.TICK_ALL)
                for a in range(world):
                    table.release(a, vacated[a])
                # ---- the single-GPU reference tick
                e1.set_entities(x, z)
                b1, k1 = engine.make_batch(n_sub, sphere=(cx, cz, r))
                s1 = e1.tick(b1, t_ns, capi.TICK_ALL)
                # ---- compare every local subscriber
                p, p1 = e.get_pairs(s.n_pairs), e1.get_pairs(s1.n_pairs)
                voff, vis = e.get_visible()
                voff1, vis1 = e1.get_visible()
                bad = 0
                for s_, j in zip(my_slots.tolist(), my_subs.tolist()):
                    a, b = slice(p["off"][s_], p["off"][s_ + 1]), slice(p1["off"][j], p1["off"][j + 1])
                    ok = all(np.array_equal(p[k][a], p1[k][b]) for k in ("channel", "dist", "interval", "flags", "last", "last_index"))
                    ok = ok and np.array_equal(np.sort(vis[int(voff[s_]):int(voff[s_ + 1])]), np.sort(vis1[int(voff1[j]):int(voff1[j + 1])]))
                    bad += 0 if ok else 1
                # slots that are not in use hold no pairs
                in_use = np.zeros(e.n_slots + 1, bool)
                in_use[my_slots] = True
                stray = int(((p["off"][1:] - p["off"][:-1])[~in_use[:-1][: len(p["off"]) - 1]]).sum()) if e.n_slots else 0
                bad += stray
                conn_of_slot = np.zeros(max(e.n_slots, 1), np.uint32)
                conn_of_slot[my_slots] = conn[my_subs]
                got_due = due_set(e.get_due(s.n_due), conn_of_slot)
                mine_conn = set(conn[my_subs].tolist())
                want_due = {d for d in due_set(e1.get_due(s1.n_due), conn) if d[0] in mine_conn}
                if got_due != want_due:
                    bad += 1 + len(got_due ^ want_due)
                if tick == 0:  # the reference engine itself against the CPU oracle (whole world)
                    want = orc.sphere_tick(og, x, z, cx, cz, r)
                    if not (np.array_equal(p1["channel"], want["pair_cell"]) and np.array_equal(vis1, want["vis_entity"])):
                        bad += 1
                failures += bad
                # ---- entity re-homing for the next tick: sender drops, receiver adopts
                g_id, g_dst, g_n = e.get_rehome()
                allre = [None] * world
                dist.all_gather_object(allre, (g_id.tolist(), g_dst.tolist()))
                leaving = set(allre[rank][0])
                adopt = [g for (ids, dsts) in allre for g, d_ in zip(ids, dsts) if d_ == rank]
                if leaving or adopt:
                    mine = np.array(sorted((set(mine.tolist()) - leaving) | set(adopt)), np.int64)
                n_rehome_total += sum(len(v[0]) for v in allre)
                print("rank %d %s r=%g tick %d: own=%d subs=%d/%d checked=%d mismatches=%d migrated=%d rehomed=%d" %
                      (rank, name, radius, tick, len(mine), len(my_slots), n_sub, len(my_slots), bad,
                       sum(len(v) for v in out_lists), sum(len(v[0]) for v in allre)), flush=True)
                # ownership invariant after re-homing: every in-world entity sits on the owner of its column
                ent_owner = sharding.owner_of_column(sharding.column_of(x, wc.offx, wc.w, wc.cols), wc.cols, world)
                in_world = sharding.column_of(x, wc.offx, wc.w, wc.cols) >= 0
                z_ok = (sharding.column_of(z, wc.offz, wc.h, wc.rows) >= 0)
                ok_mask = in_world & z_ok
                if not np.array_equal(np.nonzero((ent_owner == rank) & ok_mask)[0], mine[ok_mask[mine]]):
                    failures += 1
                    print("rank %d: ownership invariant violated" % rank, flush=True)
        n_coll = e.collective_count()
        if (mode == 2 and n_coll != 0) or (mode == 1 and n_coll != n_ticks):
            failures += 1
            print("rank %d: %d collectives in mode %d" % (rank, n_coll, mode), flush=True)
        if n_mig_total == 0 and name != "2x2":
            print("rank %d %s: WARNING no subscriber migrated" % (rank, name), flush=True)
        e.close()
        e1.close()
    t = torch.tensor([failures], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t[0]) != 0:
        raise SystemExit("multi-GPU parity FAILED: %d mismatches" % int(t[0]))
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
