"""BroadcastType_ADJACENT_CHANNELS recipient sets (message.go:188-239): the oracle restatement against cases worked out
by hand from the reference code.  The reference has NO test for this branch (no *_test.go mentions ADJACENT_CHANNELS),
so beyond these hand cases the restatement itself is the ground truth for the GPU parity test — parity unpinned."""
import numpy as np

from tests._oracle import make_grid

S0 = 65536
ALL, BUT_SENDER, BUT_OWNER, BUT_CLIENT, BUT_SERVER, ADJ = 2, 4, 8, 16, 32, 64
SERVER, CLIENT = 1, 2


def _csr(cells, lists):
    off, conn, typ = [0], [], []
    for c in range(cells):
        for cid, t in lists.get(c, []):
            conn.append(cid)
            typ.append(t)
        off.append(len(conn))
    return np.array(off, np.uint32), np.array(conn, np.uint32), np.array(typ, np.uint8)


def test_adjacent_broadcast_hand_cases(oracle):
    g = make_grid(0, 0, 100, 100, 3, 3)
    # 3 x 3 grid, cells 0..8 row-major; connection 7 is subscribed to three cells (must be sent to once), 1 is a server
    lists = {0: [(10, CLIENT), (7, CLIENT)], 1: [(1, SERVER)], 4: [(20, CLIENT), (7, CLIENT), (1, SERVER)], 8: [(30, CLIENT), (7, CLIENT)],
             5: []}
    off, conn, typ = _csr(9, lists)
    bc = lambda ch, flags, snd=0, cli=0: oracle.adjacent_broadcast(g, S0 + ch, flags, snd, cli, off, conn, typ).tolist()
    # centre cell 4: all eight neighbours + itself
    assert bc(4, ADJ) == [1, 7, 10, 20, 30]
    # ALL_BUT_OWNER drops the centre channel's own list, not connections that are also subscribed to a neighbour
    assert bc(4, ADJ | BUT_OWNER) == [1, 7, 10, 30]
    assert bc(4, ADJ | BUT_SENDER, snd=7) == [1, 10, 20, 30]
    assert bc(4, ADJ | BUT_SENDER, snd=99) == [1, 7, 10, 20, 30]
    assert bc(4, ADJ, snd=7) == [1, 7, 10, 20, 30]  # the sender is only skipped when ALL_BUT_SENDER is set
    assert bc(4, ADJ | BUT_CLIENT) == [1]
    assert bc(4, ADJ | BUT_SERVER) == [7, 10, 20, 30]
    assert bc(4, ADJ, cli=20) == [1, 7, 10, 30]  # ServerForwardMessage.ClientConnId is always skipped
    assert bc(4, ADJ | BUT_CLIENT | BUT_SERVER) == []
    # corner cell 0: neighbours 1, 3, 4 + itself; cell 8 is out of reach
    assert bc(0, ADJ) == [1, 7, 10, 20]
    assert bc(0, ADJ | BUT_OWNER) == [1, 7, 20]
    # corner cell 8: neighbours 4, 5, 7 + itself
    assert bc(8, ADJ) == [1, 7, 20, 30]
    # edge cell 2: neighbours 1, 4, 5 + itself (empty)
    assert bc(2, ADJ | BUT_OWNER) == [1, 7, 20]


def test_adjacent_broadcast_dedup_random(oracle):
    """Against a plain-Python set construction on random subscriber lists."""
    rng = np.random.default_rng(5)
    g = make_grid(-450, -200, 100, 50, 9, 8)
    cells = 72
    types = {cid: int(rng.integers(1, 3)) for cid in range(1, 200)}
    lists = {c: [(int(cid), types[int(cid)]) for cid in rng.choice(np.arange(1, 200), size=int(rng.integers(0, 25)), replace=False)]
             for c in range(cells)}
    off, conn, typ = _csr(cells, lists)
    for _ in range(300):
        ch = int(rng.integers(0, cells))
        flags = ADJ | int(rng.choice([0, BUT_SENDER, BUT_OWNER, BUT_CLIENT, BUT_SERVER, BUT_SENDER | BUT_OWNER, BUT_OWNER | BUT_SERVER]))
        snd, cli = int(rng.integers(0, 200)), int(rng.choice([0, int(rng.integers(1, 200))]))
        gx, gy = ch % 9, ch // 9
        want = set()
        for y in range(gy - 1, gy + 2):
            for x in range(gx - 1, gx + 2):
                if 0 <= x < 9 and 0 <= y < 8 and not ((x, y) == (gx, gy) and flags & BUT_OWNER):
                    want |= {cid for cid, _ in lists[x + 9 * y]}
        want = {c for c in want if not (flags & BUT_SENDER and c == snd) and not (flags & BUT_CLIENT and types[c] == CLIENT)
                and not (flags & BUT_SERVER and types[c] == SERVER) and c != cli}
        got = oracle.adjacent_broadcast(g, S0 + ch, flags, snd, cli, off, conn, typ).tolist()
        assert got == sorted(want)
