"""Multi-GPU parity check (run under torchrun on a GPU box, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_multigpu_parity.py

Every rank owns an X-slab, exchanges border entity records with ONE all-gather per tick, rebuilds its cell CSR
over own + halo entities and answers the subscribers whose centre lies in its slab.  Each rank checks its
subscribers' (cell, dist) pairs and visible-entity lists against the single-process oracle: the union over ranks
equals the single-GPU / oracle answer (SURVEY.md §8e parity row)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from channeld_b200 import capi, engine, sharding, synth  # noqa: E402
from tests import _oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    orc = _oracle.load()
    failures = 0
    # the 2x2 grid gives every rank of a 2-GPU run a ONE-column slab (empty interior, as at 8 GPUs on 15 columns)
    for name, n_ent, n_sub, radius, max_move in (("benchmark", 120_000, 6_000, 50.0, 60.0), ("benchmark", 60_000, 3_000, 2500.0, 900.0),
                                                 ("handover", 200_000, 4_000, 50.0, 120.0), ("2x2", 50_000, 2_000, 50.0, 60.0),
                                                 ("2x2", 50_000, 2_000, 500.0, 300.0)):
        wc = synth.scaled(synth.CONFIGS[name], n_ent, n_sub)
        if world > wc.cols:
            continue
        og = _oracle.make_grid(wc.offx, wc.offz, wc.w, wc.h, wc.cols, wc.rows, wc.server_cols, wc.server_rows)
        ex, ez = synth.entities(wc)
        halo = sharding.halo_columns(radius, wc.w)
        lo, hi = sharding.slab_columns(wc.cols, world, rank)
        ent_col = sharding.column_of(ex, wc.offx, wc.w, wc.cols)
        mine = np.nonzero(((ent_col >= lo) & (ent_col < hi)) | ((ent_col < 0) & (rank == 0)))[0]
        cap = n_ent  # generous border capacity for the test
        e = engine.Engine(wc.cfg(), len(mine) + cap * world + 16, n_sub, device=local, max_visible=1 << 27)
        stream = torch.cuda.Stream(device=dev)
        e.set_stream(stream.cuda_stream)
        e.set_slab(lo, hi, halo)
        e.set_entity_ids(mine.astype(np.uint32))
        rec_local = torch.full((cap * 2,), -1, dtype=torch.int32, device=dev)
        rec_all = torch.empty((cap * 2 * world,), dtype=torch.int32, device=dev)
        x, z = ex.copy(), ez.copy()
        with torch.cuda.stream(stream):
            for tick in range(3):
                x, z = synth.move_entities(wc, x, z, tick, max_move)  # entities drift across slab borders
                e.set_entities(x[mine], z[mine])
                n_exp = e.export_border(rec_local, cap)
                dist.all_gather_into_tensor(rec_all, rec_local)
                e.import_halo(rec_all, cap * world, rank * cap, cap)
                conn, cx, cz, r = synth.subscribers(wc, x, z, radius)
                sub_col = sharding.column_of(cx, wc.offx, wc.w, wc.cols)
                smine = np.nonzero((sub_col >= lo) & (sub_col < hi))[0]
                e.set_subscribers(conn[smine]) if tick == 0 else None
                if tick == 0:
                    s0 = smine  # subscriber placement is static; follow the same subscribers afterwards
                q = s0
                batch, keep = engine.make_batch(len(q), sub=np.arange(len(q), dtype=np.uint32), sphere=(cx[q], cz[q], r[q]))
                s = e.tick(batch, (tick + 1) * 33_000_000, capi.TICK_BUILD | capi.TICK_EMIT)
                want = orc.sphere_tick(og, x, z, cx[q], cz[q], r[q])
                pairs = e.get_pairs(s.n_pairs)
                voff, vis = e.get_visible()
                # a subscriber that drifted more than `halo` columns out of this slab would need re-homing: skip those
                col_now = sharding.column_of(cx[q], wc.offx, wc.w, wc.cols)
                ok_sub = (col_now >= lo - 0) & (col_now < hi + 0) if halo < 2 else (col_now >= lo - 1) & (col_now < hi + 1)
                bad = 0
                for k in np.nonzero(ok_sub)[0]:
                    a = slice(pairs["off"][k], pairs["off"][k + 1])
                    b = slice(int(want["pair_off"][k]), int(want["pair_off"][k + 1]))
                    if not (np.array_equal(pairs["channel"][a], want["pair_cell"][b]) and np.array_equal(pairs["dist"][a], want["pair_dist"][b])):
                        bad += 1
                        continue
                    got = vis[int(voff[k]):int(voff[k + 1])]
                    exp = want["vis_entity"][int(want["vis_off"][k]):int(want["vis_off"][k + 1])]
                    # halo entities of one cell come from one owner in its id order; own and halo never mix inside a
                    # cell unless an entity left its owner's slab: compare as sets per cell => sort within the list
                    if not np.array_equal(np.sort(got), np.sort(exp)):
                        bad += 1
                failures += bad
                print("rank %d %s r=%g tick %d: own=%d exported=%d subs=%d checked=%d mismatches=%d" %
                      (rank, name, radius, tick, len(mine), n_exp, len(q), int(ok_sub.sum()), bad), flush=True)
        e.close()
    t = torch.tensor([failures], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t[0]) != 0:
        raise SystemExit("multi-GPU parity FAILED: %d mismatching subscribers" % int(t[0]))
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
