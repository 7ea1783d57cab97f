"""GPU parity against the COMMITTED golden vectors (tests/golden/), through the C ABI, without running the oracle:
QueryChannelIds for 8 grids x 200 random queries of every AOI kind, and one full tick of BASELINE config #1."""
import numpy as np
import pytest

from tests.golden import golden_io as gio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def chd():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from channeld_b200 import capi, controller, engine, synth

    capi.lib()

    class NS:
        pass

    ns = NS()
    ns.capi, ns.controller, ns.engine, ns.synth = capi, controller, engine, synth
    return ns


def test_golden_random_queries(chd):
    for block in gio.load_queries():
        g = block["grid"]
        c = chd.controller.GpuStaticGrid2DSpatialController(max_entities=16, max_subscribers=2048, max_queries=2048, max_spots=1 << 15,
                                                            max_window_cells=1 << 22, max_pairs=1 << 20)
        c.LoadConfig(dict(WorldOffsetX=g[0], WorldOffsetZ=g[1], GridWidth=g[2], GridHeight=g[3], GridCols=g[4], GridRows=g[5],
                          ServerCols=1, ServerRows=1))
        got = c.QueryChannelIdsBatch([gio.dict_to_query(case["query"]) for case in block["cases"]])
        for case, res in zip(block["cases"], got):
            if case["status"] != 0:
                assert isinstance(res, chd.controller.SpatialError), case
            else:
                assert res == {int(k): int(v) for k, v in case["result"]}, case


@pytest.mark.parametrize("radius", gio.CONFIG1_RADII)
def test_golden_config1_tick(chd, radius):
    want = gio.load_config1(radius)
    wc = chd.synth.CONFIGS["2x2"]
    ex, ez = chd.synth.entities(wc)
    conn, cx, cz, r = chd.synth.subscribers(wc, ex, ez, float(radius))
    e = chd.engine.Engine(wc.cfg(), wc.n_entities, wc.n_subscribers, max_visible=1 << 20)
    e.set_entities(ex, ez)
    e.set_subscribers(conn)
    batch, keep = chd.engine.make_batch(len(cx), sub=np.arange(len(cx), dtype=np.uint32), sphere=(cx, cz, r))
    s = e.tick(batch, 0, chd.capi.TICK_BUILD | chd.capi.TICK_EMIT)
    assert s.n_pairs == len(want["pair_cell"]) and s.n_visible == len(want["vis_entity"])
    pairs = e.get_pairs()
    np.testing.assert_array_equal(pairs["off"].astype(np.uint64), want["pair_off"])
    np.testing.assert_array_equal(pairs["channel"], want["pair_cell"])
    np.testing.assert_array_equal(pairs["dist"], want["pair_dist"])
    voff, vis = e.get_visible()
    np.testing.assert_array_equal(voff, want["vis_off"])
    np.testing.assert_array_equal(vis, want["vis_entity"])
    np.testing.assert_array_equal(e.get_query_status(len(cx)), want["status"])
