"""GPU test of the byte half of the fan-out (chd_set_payload_bytes / chd_assemble_payloads / chd_frame_packets, SURVEY.md §8f
rank 1 and 4) at MESSAGE level (byte parity with Go's marshal is undefined): the bytes the device assembles are
decoded with the protobuf runtime and compared with proto.Merge (MergeFrom) over the ring entries the literal tickData
emulation (oracle channel) actually merged; the framed packets are parsed like connection.go's reader does (5-byte tag, snappy
via an independent decoder: pyarrow's)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
S0 = 65536
MS = 1_000_000


@pytest.fixture(scope="module")
def chd():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from channeld_b200 import capi, engine

    capi.lib()

    class NS:
        pass

    ns = NS()
    ns.capi, ns.engine = capi, engine
    return ns


def _snappy_decode(body):
    import pyarrow as pa

    n, shift, i = 0, 0, 0  # uncompressed length varint
    while True:
        b = body[i]
        n |= (b & 0x7F) << shift
        i += 1
        shift += 7
        if not b & 0x80:
            break
    return pa.decompress(bytes(body), decompressed_size=n, codec="snappy", asbytes=True)


def test_payload_assembly_and_framing(chd, oracle):
    from tests import _wire as W
    from tests._oracle import make_grid

    capi = chd.capi
    g = (-300.0, -250.0, 100.0, 100.0, 6, 5)
    og = make_grid(*g)
    rng = np.random.default_rng(5)
    S, N = 40, 200
    cells = g[4] * g[5]
    e = chd.engine.Engine(chd.engine.grid_cfg(*g), N, S, max_visible=1 << 20)
    conn = np.arange(101, 101 + S, dtype=np.uint32)
    e.set_subscribers(conn)
    e.set_entities(rng.uniform(-300, 300, N), rng.uniform(-250, 250, N))
    e.build()
    chans = [oracle.channel() for _ in range(cells)]
    rings = [[] for _ in range(cells)]      # (arrival, sender, index, message)
    full = [W.TestChannelDataMessage() for _ in range(cells)]  # ch.data.msg: every update merged in (data.go:149-160)
    msg_index = np.zeros(cells, np.uint64)
    cx, cz = rng.uniform(-280, 280, S), rng.uniform(-230, 230, S)
    rad = rng.choice([30.0, 60.0, 120.0], S)
    compression = (np.arange(S) % 2).astype(np.uint8)  # every other connection asked for snappy
    t = 0
    n_checked = n_update_msgs = n_shared = n_multi_frame = 0
    for tick in range(10):
        t += 33 * MS if tick % 3 else 90 * MS
        cx += rng.uniform(-40, 40, S)
        cz += rng.uniform(-40, 40, S)
        batch, keep = chd.engine.make_batch(S, sphere=(cx, cz, rad))
        e.update_interest(batch, t)
        s = e.summary()
        pairs = e.get_pairs()
        for j in range(S):
            res, st = oracle.query(og, sphere=(cx[j], cz[j], rad[j]))
            if st != 0:
                continue
            have = set(int(c) for c in pairs["channel"][pairs["off"][j]:pairs["off"][j + 1]])
            for c in range(cells):
                ch = S0 + c
                if ch in res:
                    chans[c].subscribe(int(conn[j]), t, oracle.damping(res[ch], 20), 0, True, False)
                else:
                    chans[c].unsubscribe(int(conn[j]))
            assert have == set(res.keys())
        for _ in range(int(rng.integers(10, 40))):
            c = int(rng.integers(0, cells))
            arrival = t - int(rng.integers(0, 60)) * MS
            sender = int(rng.choice(conn)) if rng.random() < 0.4 else 7
            m = W.TestChannelDataMessage()
            if rng.random() < 0.7:
                m.text = "t%d-%d" % (tick, int(rng.integers(0, 1000)))
            if rng.random() < 0.5:
                m.num = int(rng.integers(1, 1 << 20))
            for _k in range(int(rng.integers(0, 3))):
                m.list.append("e%d" % int(rng.integers(0, 100)))
            for _k in range(int(rng.integers(0, 3))):
                m.kv[int(rng.integers(0, 6))] = "v%d" % int(rng.integers(0, 100))
            msg_index[c] += 1
            rings[c].append((arrival, sender, int(msg_index[c]), m))
            full[c].MergeFrom(m)
            chans[c].on_update(arrival, sender)
        ring_off = np.concatenate([[0], np.cumsum([len(r) for r in rings])]).astype(np.uint32)
        flat = [x for r in rings for x in r]
        e.set_rings(ring_off, np.array([f[0] for f in flat], np.int64), np.array([f[1] for f in flat], np.uint32),
                    np.array([f[2] for f in flat], np.uint64), msg_index)
        ebytes = [f[3].SerializeToString() for f in flat]
        eoff = np.concatenate([[0], np.cumsum([len(b) for b in ebytes])]).astype(np.uint64)
        fbytes = [m.SerializeToString() for m in full]
        foff = np.concatenate([[0], np.cumsum([len(b) for b in fbytes])]).astype(np.uint64)
        eb = np.frombuffer(b"".join(ebytes) or b"\0", np.uint8).copy()
        fb = np.frombuffer(b"".join(fbytes) or b"\0", np.uint8).copy()
        st_ = e.L.chd_set_payload_bytes(e.h, capi.ptr(eoff), len(flat), capi.ptr(eb), capi.ptr(foff), capi.ptr(fb), W.TYPE_URL.encode(),
                                         W.MSG_CHANNEL_DATA_UPDATE)
        assert st_ == capi.OK, e.L.chd_last_error(e.h)
        e.fanout_tick(t)
        s = e.summary()
        due = e.get_due(s.n_due)
        # ---- what the reference would send: per (conn, channel, step) the merge of the entries the emulation picked
        want = {}
        for c in range(cells):
            for d in chans[c].tick_data_ex(t):
                if d["kind"] == 0:
                    m = W.TestChannelDataMessage()
                    m.CopyFrom(full[c])
                else:
                    m = W.TestChannelDataMessage()
                    for pos in d["selected"]:
                        m.MergeFrom(rings[c][pos][3])  # proto.Merge of every selected entry, in order (data.go:246-256)
                    n_update_msgs += 1
                want.setdefault((d["conn"], S0 + c), []).append(m)
        # ---- per-class blobs
        ncls = C.c_uint32()
        blob_len = C.c_uint64()
        cap_cls = max(int(s.n_due), 1)
        cls_off = np.zeros(cap_cls + 1, np.uint64)
        blob = np.zeros(1 << 22, np.uint8)
        st_ = e.L.chd_assemble_payloads(e.h, C.byref(ncls), capi.ptr(cls_off), cap_cls, capi.ptr(blob), blob.size, C.byref(blob_len))
        assert st_ == capi.OK, e.L.chd_last_error(e.h)
        cls_of, cls_rep, cls_cnt = e.due_classes(int(s.n_due))
        assert len(cls_rep) == ncls.value
        n_shared += int(s.n_due) - ncls.value
        got = {}
        for i, d in enumerate(due):
            k = int(cls_of[i])
            entry = bytes(blob[int(cls_off[k]):int(cls_off[k + 1])])
            pk = W.Packet()
            pk.ParseFromString(entry)  # one Packet.messages entry parses as a Packet with one message
            assert len(pk.messages) == 1
            mp = pk.messages[0]
            assert (mp.channelId, mp.msgType, mp.broadcast, mp.stubId) == (int(d["channel_id"]), 8, 0, 0)
            cdu = W.ChannelDataUpdateMessage()
            cdu.ParseFromString(mp.msgBody)
            assert cdu.contextConnId == 0 and cdu.data.type_url == W.TYPE_URL
            m = W.TestChannelDataMessage()
            m.ParseFromString(cdu.data.value)
            got.setdefault((int(conn[d["sub"]]), int(d["channel_id"])), []).append(m)
        assert set(got) == set(want)
        for key in want:
            assert sorted(x.SerializeToString(deterministic=True) for x in got[key]) == sorted(x.SerializeToString(deterministic=True) for x in want[key]), key
            n_checked += len(want[key])
        # ---- framed packets per connection
        conn_off = np.zeros(S + 1, np.uint64)
        conn_len = np.zeros(S, np.uint32)
        conn_frames = np.zeros(S, np.uint32)
        out = np.zeros(1 << 23, np.uint8)
        out_len, dropped = C.c_uint64(), C.c_uint32()
        st_ = e.L.chd_frame_packets(e.h, capi.ptr(compression), capi.ptr(conn_off), capi.ptr(conn_len), capi.ptr(conn_frames), capi.ptr(out), out.size,
                                     C.byref(out_len), C.byref(dropped))
        assert st_ == capi.OK, e.L.chd_last_error(e.h)
        assert dropped.value == 0
        for j in range(S):
            buf = bytes(out[int(conn_off[j]):int(conn_off[j]) + int(conn_len[j])])
            pos, msgs, frames = 0, [], 0
            while pos < len(buf):  # connection.go:456-520: tag, size, compression type, body
                assert buf[pos] == 67 and buf[pos + 1] == 72
                size, ct = (buf[pos + 2] << 8) | buf[pos + 3], buf[pos + 4]
                assert ct == compression[j] and size <= 0xFFFF
                body = buf[pos + 5:pos + 5 + size]
                assert len(body) == size
                if ct == 1:
                    body = _snappy_decode(body)
                    assert len(body) <= 0xFFFF
                pk = W.Packet()
                pk.ParseFromString(body)
                msgs += list(pk.messages)
                pos += 5 + size
                frames += 1
            assert frames == conn_frames[j]
            n_multi_frame += frames > 1
            mine = [(int(d["channel_id"]), i) for i, d in enumerate(due) if d["sub"] == j]
            assert len(msgs) == len(mine)
            for mp, (ch, i) in zip(msgs, sorted(mine, key=lambda x: x[1])):  # packet order = ascending due index
                k = int(cls_of[i])
                assert mp.SerializeToString() == W.Packet.FromString(bytes(blob[int(cls_off[k]):int(cls_off[k + 1])])).messages[0].SerializeToString()
    assert n_checked > 300 and n_update_msgs > 100 and n_shared > 20
    e.close()


def test_framing_splits_and_drops(chd):
    """connection.go:626-714 / :73-77 with sizes under control: 3 cells in a row, 4 connections subscribed to all of them.
    FULL sends of 50 KB per cell -> 150 KB per connection -> three packets; then UPDATE windows of 60 KB (cell 0), 30 KB (cell 1)
    and 90 KB (cell 2: >= MaxPacketSize - 5, dropped like queuedMessagePackSender.Send does)."""
    from tests import _wire as W

    capi = chd.capi
    e = chd.engine.Engine(chd.engine.grid_cfg(0, 0, 100, 100, 3, 1), 8, 4)
    conn = np.array([11, 12, 13, 14], np.uint32)
    e.set_subscribers(conn)
    e.set_entities(np.array([50.0]), np.array([50.0]))
    e.build()
    batch, keep = chd.engine.make_batch(4, sphere=(np.full(4, 150.0), np.full(4, 50.0), np.full(4, 140.0)))
    e.update_interest(batch, 0)
    assert e.summary().n_pairs == 12

    def msg(n):
        m = W.TestChannelDataMessage()
        m.text = "x" * n
        return m.SerializeToString()

    def run(t, ring_entries, fulls, compression):
        """ring_entries: per cell list of (arrival, sender, bytes)"""
        roff = np.concatenate([[0], np.cumsum([len(r) for r in ring_entries])]).astype(np.uint32)
        flat = [x for r in ring_entries for x in r]
        e.set_rings(roff, np.array([f[0] for f in flat], np.int64), np.array([f[1] for f in flat], np.uint32),
                    np.arange(1, len(flat) + 1, dtype=np.uint64), np.array([len(r) for r in ring_entries], np.uint64))
        eoff = np.concatenate([[0], np.cumsum([len(f[2]) for f in flat])]).astype(np.uint64)
        eb = np.frombuffer(b"".join(f[2] for f in flat) or b"\0", np.uint8).copy()
        foff = np.concatenate([[0], np.cumsum([len(b) for b in fulls])]).astype(np.uint64)
        fb = np.frombuffer(b"".join(fulls) or b"\0", np.uint8).copy()
        assert e.L.chd_set_payload_bytes(e.h, capi.ptr(eoff), len(flat), capi.ptr(eb), capi.ptr(foff), capi.ptr(fb), W.TYPE_URL.encode(), 8) == capi.OK
        e.fanout_tick(t)
        s = e.summary()
        due = e.get_due(s.n_due)
        ncls, blen = C.c_uint32(), C.c_uint64()
        cls_off = np.zeros(int(s.n_due) + 2, np.uint64)
        blob = np.zeros(1 << 20, np.uint8)
        assert e.L.chd_assemble_payloads(e.h, C.byref(ncls), capi.ptr(cls_off), int(s.n_due) + 1, capi.ptr(blob), blob.size, C.byref(blen)) == capi.OK
        cls_of, _, _ = e.due_classes(int(s.n_due))
        conn_off, conn_len, conn_frames = np.zeros(5, np.uint64), np.zeros(4, np.uint32), np.zeros(4, np.uint32)
        out = np.zeros(1 << 21, np.uint8)
        olen, dropped = C.c_uint64(), C.c_uint32()
        comp = np.array(compression, np.uint8)
        assert e.L.chd_frame_packets(e.h, capi.ptr(comp), capi.ptr(conn_off), capi.ptr(conn_len), capi.ptr(conn_frames), capi.ptr(out), out.size,
                                     C.byref(olen), C.byref(dropped)) == capi.OK
        per_conn = []
        for j in range(4):
            buf = bytes(out[int(conn_off[j]):int(conn_off[j]) + int(conn_len[j])])
            pos, frames = 0, []
            while pos < len(buf):
                assert buf[pos:pos + 2] == b"CH" and buf[pos + 4] == comp[j]
                size = (buf[pos + 2] << 8) | buf[pos + 3]
                body = buf[pos + 5:pos + 5 + size]
                if comp[j]:
                    body = _snappy_decode(body)
                pk = W.Packet.FromString(body)
                frames.append([(m.channelId, len(m.msgBody)) for m in pk.messages])
                assert len(body) <= 0xFFFF
                pos += 5 + size
            assert len(frames) == conn_frames[j]
            per_conn.append(frames)
        sizes = {}
        for i, d in enumerate(due):
            k = int(cls_of[i])
            sizes.setdefault(int(d["sub"]), []).append((int(d["channel_id"]), int(cls_off[k + 1] - cls_off[k])))
        return per_conn, sizes, dropped.value, due

    # tick 1: everybody's first fan-out = FULL channel data, 50 KB per cell
    fulls = [msg(50000), msg(50000), msg(50000)]
    per_conn, sizes, dropped, due = run(100 * MS, [[], [], []], fulls, [0, 1, 0, 1])
    assert dropped == 0 and len(due) == 12 and (due["kind"] == 0).all()
    for j in range(4):
        assert [len(f) for f in per_conn[j]] == [1, 1, 1]  # 3 x 50 KB cannot share a 64 KB packet
        assert sorted(ch for f in per_conn[j] for ch, _ in f) == [S0, S0 + 1, S0 + 2]
    # tick 2: updates at 150 ms fall into the catch-up windows; 60 KB + 30 KB are sent (two packets), 90 KB is dropped
    ring = [[(150 * MS, 7, msg(30000)), (151 * MS, 7, msg(30000))], [(150 * MS, 7, msg(30000))],
            [(150 * MS, 7, msg(30000)), (151 * MS, 7, msg(30000)), (152 * MS, 7, msg(30000))]]
    per_conn, sizes, dropped, due = run(200 * MS, ring, fulls, [1, 0, 1, 0])
    assert (due["kind"] == 1).all()
    n_big = sum(1 for j in sizes for ch, L in sizes[j] if L >= 0xFFFF - 5)
    assert dropped == n_big > 0
    for j in range(4):
        sent = [(ch, L) for ch, L in sizes[j] if L < 0xFFFF - 5]
        # greedy packets in due order (connection.go:643-659)
        want, cur = [], []
        for ch, L in sent:
            if sum(x[1] for x in cur) + L > 0xFFFF:
                want.append(cur); cur = []
            cur.append((ch, L))
        if cur:
            want.append(cur)
        assert [[ch for ch, _ in f] for f in per_conn[j]] == [[ch for ch, _ in f] for f in want], (j, per_conn[j], want)
        assert len(want) >= 2
    e.close()
