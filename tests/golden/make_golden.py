"""Generates the golden vectors under tests/golden/ with the CPU oracle (oracle/, pinned by the reference's own
known-answer tests in tests/test_oracle_kat.py; the Go reference itself cannot run in this image).

    python -m tests.golden.make_golden        # from the repository root

  random_queries.json   per grid of tests/test_gpu_parity.GRIDS: 200 seeded random SpatialInterestQuery of every AOI kind /
                        combination (incl. failing ones) with the expected status and {channel id: dist} result of
                        QueryChannelIds (spatial.go:182-317)
  config1_r{50,500,2500}.npz   BASELINE config #1 (spatial_static_2x2, 1 000 entities / 256 subscribers): statuses, (cell, dist)
                        pairs, visible-list offsets and the visible entity lists of one tick

tests/test_golden.py checks that the oracle still reproduces these files bit for bit (CPU), tests/test_z_gpu_golden.py
checks the CUDA path against them through the C ABI (GPU) without running the oracle.
"""
import json

import numpy as np

from tests import _oracle
from tests.golden import golden_io as gio


def make_queries(oracle):
    from tests.test_gpu_parity import GRIDS, _random_queries

    out = []
    for gi, g in enumerate(GRIDS):
        rng = np.random.default_rng(gio.SEED + gi)
        og = _oracle.make_grid(*g)
        cases = []
        for q in _random_queries(rng, g, gio.QUERIES_PER_GRID):
            d = gio.query_to_dict(q)
            res, st = oracle.query(og, **gio.oracle_kwargs(d))
            cases.append({"query": d, "status": int(st), "result": sorted([int(k), int(v)] for k, v in res.items()) if st == 0 else None})
        out.append({"grid": list(g), "cases": cases})
    return out


def make_config1(oracle, radius):
    from channeld_b200 import synth
    from tests.test_gpu_parity import _oracle_grid

    wc = synth.CONFIGS["2x2"]
    ex, ez = synth.entities(wc)
    conn, cx, cz, r = synth.subscribers(wc, ex, ez, float(radius))
    want = oracle.sphere_tick(_oracle_grid(wc), ex, ez, cx, cz, r)
    return {k: np.ascontiguousarray(v) for k, v in want.items()}


def main():
    oracle = _oracle.load()
    with open(gio.QUERIES, "w") as f:
        json.dump(make_queries(oracle), f, separators=(",", ":"))
    for radius in gio.CONFIG1_RADII:
        np.savez_compressed(gio.CONFIG1 % radius, **make_config1(oracle, radius))


if __name__ == "__main__":
    main()
