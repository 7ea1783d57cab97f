"""Deterministic synthetic worlds: the same bits from numpy here, from C++ or from Go.

RNG = splitmix64 in counter mode: u(i,k) = (splitmix64(seed*0x9E3779B97F4A7C15 + 2*i + k) >> 11) * 2^-53.
Entity i:   x = offX + u(i,0)*worldW ; z = offZ + u(i,1)*worldH   (multiply, then add: two roundings).
Subscriber j stands on entity j*(N//S): sphere AOI of radius r around it, connection id j+1.
"""
from dataclasses import dataclass

import numpy as np

from .engine import grid_cfg

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        z = x + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed, i, k):
    i = np.asarray(i, np.uint64)
    with np.errstate(over="ignore"):
        ctr = np.uint64(seed) * _GOLDEN + np.uint64(2) * i + np.uint64(k)
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


@dataclass
class WorldConfig:
    name: str
    offx: float
    offz: float
    w: float
    h: float
    cols: int
    rows: int
    server_cols: int
    server_rows: int
    n_entities: int
    n_subscribers: int
    radius: float
    seed: int

    def cfg(self):
        return grid_cfg(self.offx, self.offz, self.w, self.h, self.cols, self.rows, self.server_cols, self.server_rows, 0)

    @property
    def cells(self):
        return self.cols * self.rows


# BASELINE.json configs (grid constants of #1/#2 from config/spatial_static_2x2.json / spatial_static_benchmark.json)
CONFIGS = {
    "2x2": WorldConfig("spatial_static_2x2 1K/256", -2000, -2000, 2000, 2000, 2, 2, 1, 2, 1000, 256, 50.0, 1),
    "benchmark": WorldConfig("spatial_static_benchmark 1M/100K r=50", -15000, -15000, 2000, 2000, 15, 15, 3, 3,
                             1_000_000, 100_000, 50.0, 2),
    "10m": WorldConfig("64x64 grid 10M/1M r=50", -64000, -64000, 2000, 2000, 64, 64, 1, 1, 10_000_000, 1_000_000, 50.0, 3),
    "handover": WorldConfig("256x256 grid 10M/1M r=50", -12800, -12800, 100, 100, 256, 256, 1, 1, 10_000_000, 1_000_000,
                            50.0, 5),
}


def scaled(cfg: WorldConfig, n_entities, n_subscribers):
    import copy

    c = copy.copy(cfg)
    c.n_entities, c.n_subscribers = int(n_entities), int(n_subscribers)
    return c


def entities(wc: WorldConfig, n=None, first=0):
    n = wc.n_entities if n is None else n
    i = np.arange(first, first + n, dtype=np.uint64)
    x = wc.offx + uniform(wc.seed, i, 0) * (wc.w * wc.cols)
    z = wc.offz + uniform(wc.seed, i, 1) * (wc.h * wc.rows)
    return x, z


def subscribers(wc: WorldConfig, ex, ez, radius=None):
    """-> (conn_id[S], cx[S], cz[S], r[S])"""
    S, N = wc.n_subscribers, len(ex)
    stride = max(N // S, 1)
    idx = (np.arange(S, dtype=np.int64) * stride) % N
    r = np.full(S, wc.radius if radius is None else radius, np.float64)
    return np.arange(1, S + 1, dtype=np.uint32), ex[idx].copy(), ez[idx].copy(), r


def move_entities(wc: WorldConfig, x, z, tick, max_step):
    """Deterministic per-tick displacement, uniform in [-max_step, max_step] per axis, reflected at the walls."""
    n = len(x)
    i = np.arange(n, dtype=np.uint64)
    dx = (uniform(wc.seed + 1000 + tick, i, 0) * 2.0 - 1.0) * max_step
    dz = (uniform(wc.seed + 1000 + tick, i, 1) * 2.0 - 1.0) * max_step
    x_lo, x_hi = wc.offx, wc.offx + wc.w * wc.cols
    z_lo, z_hi = wc.offz, wc.offz + wc.h * wc.rows
    nx, nz = x + dx, z + dz
    nx = np.where(nx < x_lo, 2 * x_lo - nx, nx)
    nx = np.where(nx >= x_hi, 2 * x_hi - nx - 1e-9, nx)
    nz = np.where(nz < z_lo, 2 * z_lo - nz, nz)
    nz = np.where(nz >= z_hi, 2 * z_hi - nz - 1e-9, nz)
    return nx, nz


def update_rings(wc: WorldConfig, tick, t_ns, tick_ns, updates_per_cell, n_conn, ring_len=64, state=None):
    """Synthetic per-cell update rings: every tick each cell receives `updates_per_cell` updates with arrival
    times spread over the tick and pseudo-random senders; the ring keeps the newest `ring_len`.
    Returns (state, ring_off, arrival, sender, index, channel_msg_index)."""
    C = wc.cells
    if state is None:
        state = dict(arrival=np.zeros((C, 0), np.int64), sender=np.zeros((C, 0), np.uint32), index=np.zeros((C, 0), np.uint64),
                     msg_index=np.zeros(C, np.uint64))
    u = updates_per_cell
    c = np.arange(C, dtype=np.uint64)[:, None]
    k = np.arange(u, dtype=np.uint64)[None, :]
    frac = uniform(wc.seed + 7, c * np.uint64(1_000_003) + k + np.uint64(tick) * np.uint64(97), 0)
    arr = (t_ns - tick_ns + np.sort((frac * tick_ns).astype(np.int64), axis=1)).astype(np.int64)
    snd = (splitmix64(c * np.uint64(7919) + k + np.uint64(tick) * np.uint64(104729)) % np.uint64(max(n_conn, 1))).astype(np.uint32) + 1
    idx = state["msg_index"][:, None] + k + np.uint64(1)
    state["msg_index"] = state["msg_index"] + np.uint64(u)
    state["arrival"] = np.concatenate([state["arrival"], arr], axis=1)[:, -ring_len:]
    state["sender"] = np.concatenate([state["sender"], snd], axis=1)[:, -ring_len:]
    state["index"] = np.concatenate([state["index"], idx.astype(np.uint64)], axis=1)[:, -ring_len:]
    L = state["arrival"].shape[1]
    ring_off = (np.arange(C + 1, dtype=np.uint64) * L).astype(np.uint32)
    return (state, ring_off, np.ascontiguousarray(state["arrival"].reshape(-1)), np.ascontiguousarray(state["sender"].reshape(-1)),
            np.ascontiguousarray(state["index"].reshape(-1)), state["msg_index"].copy())
