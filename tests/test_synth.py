"""Host logic: the deterministic synthetic world generator."""
import numpy as np

from channeld_b200 import synth


def test_splitmix64_known_answers():
    # splitmix64 reference sequence for seed 0 (public test vector of the algorithm): first outputs of
    # state += golden; mix(state)
    states = np.arange(3, dtype=np.uint64) * synth._GOLDEN  # wraps mod 2^64 (array arithmetic: no overflow warning)
    got = [int(v) for v in synth.splitmix64(states)]
    assert got == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]


def test_entities_deterministic_and_in_range():
    wc = synth.CONFIGS["benchmark"]
    x1, z1 = synth.entities(wc, 10000)
    x2, z2 = synth.entities(wc, 5000, first=5000)
    np.testing.assert_array_equal(x1[5000:], x2)
    np.testing.assert_array_equal(z1[5000:], z2)
    assert x1.min() >= wc.offx and x1.max() <= wc.offx + wc.w * wc.cols
    conn, cx, cz, r = synth.subscribers(synth.scaled(wc, 10000, 1000), x1, z1)
    assert len(conn) == 1000 and conn[0] == 1 and cx[1] == x1[10] and (r == 50.0).all()


def test_rings_shape():
    wc = synth.CONFIGS["2x2"]
    st, off, arr, snd, idx, cmi = synth.update_rings(wc, 0, 33_000_000, 33_000_000, 8, 100, ring_len=16)
    assert list(off) == [0, 8, 16, 24, 32] and len(arr) == 32 and (cmi == 8).all()
    st, off, arr, snd, idx, cmi = synth.update_rings(wc, 1, 66_000_000, 33_000_000, 12, 100, ring_len=16, state=st)
    assert list(off) == [0, 16, 32, 48, 64] and (cmi == 20).all()
    a = arr.reshape(4, 16)
    assert (np.diff(a, axis=1) >= 0).all() and idx.reshape(4, 16)[0, -1] == 20
