"""Helpers for the overload-refresh tests: a host (numpy) engine with the interface of HaccSR's refresh calls, built on
the oracle (test infrastructure, so the host logic of hacc_coral_b200/refresh.py can be exercised on CPU with gloo),
and generators of per-rank particle sets for a periodic global box."""
import ctypes as C

import numpy as np

from oracle import refresh_oracle as RO

F32 = RO.KEYS_F32


def message_bytes(n):
    b = n * 8 + 8 * n * 4 + n * 2
    return (b + 15) & ~15


class HostEngine:
    """numpy stand-in for the device part (csrc/refresh.cu) with the same message byte layout:
    id[n] | x y z vx vy vz mass phi [n] | mask[n], padded to 16 B."""
    torch_device = "cpu"

    def __init__(self, p):
        self.p = {k: np.ascontiguousarray(v).copy() for k, v in p.items()}
        self._msgs = None

    def refresh_message_bytes(self, n):
        return message_bytes(int(n))

    def refresh_begin(self, alo, ahi, ol, slot_of_dir):
        m = RO.alive_mask(self.p, alo, ahi)
        self.p = {k: v[m] for k, v in self.p.items()}
        idx = RO.classify(self.p, alo, ahi, ol)
        counts = np.zeros(27, dtype=np.int64)
        self._msgs = {}
        for d, ii in idx.items():
            s = int(slot_of_dir[d])
            counts[s] = ii.size
            self._msgs[s] = RO.message(self.p, ii, d, alo, ahi)
        return counts, int(m.sum())

    def refresh_pack(self, off, ptr):
        for s, q in self._msgs.items():
            n = q["x"].size
            if n == 0:
                continue
            raw = (C.c_ubyte * message_bytes(n)).from_address(ptr + int(off[s]))
            buf = np.frombuffer(raw, dtype=np.uint8)
            buf[:8 * n] = q["id"].astype(np.int64).view(np.uint8)
            for j, k in enumerate(F32):
                buf[8 * n + 4 * n * j: 8 * n + 4 * n * (j + 1)] = q[k].astype(np.float32).view(np.uint8)
            buf[40 * n: 42 * n] = q["mask"].astype(np.uint16).view(np.uint8)

    def refresh_append(self, ptr, n):
        n = int(n)
        if n == 0:
            return
        raw = (C.c_ubyte * message_bytes(n)).from_address(ptr)
        buf = np.frombuffer(raw, dtype=np.uint8)
        q = {"id": buf[:8 * n].view(np.int64).copy(), "mask": buf[40 * n:42 * n].view(np.uint16).copy()}
        for j, k in enumerate(F32):
            q[k] = buf[8 * n + 4 * n * j: 8 * n + 4 * n * (j + 1)].view(np.float32).copy()
        self.p = {k: np.concatenate([self.p[k], q[k]]) for k in self.p}


def global_particles(dims, ext, n_per_rank, seed):
    """Random particles in the periodic global box dims*ext, with global ids, as float64 positions."""
    rng = np.random.default_rng(seed)
    n = n_per_rank * dims[0] * dims[1] * dims[2]
    pos = rng.random((n, 3)) * (np.asarray(dims) * np.asarray(ext, dtype=np.float64))
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    return pos, vel


def rank_particles(pos, vel, dims, pos_of_rank, ext, ol, junk_ghosts=True, seed=0):
    """The alive particles of one rank in its local frame (alive region [ol, ol+ext)), followed by stale ghosts
    that the refresh must drop."""
    lo = np.asarray(pos_of_rank) * np.asarray(ext, dtype=np.float64)
    m = np.all((pos >= lo) & (pos < lo + ext), axis=1)
    ids = np.nonzero(m)[0]
    loc = (pos[m] - lo + ol).astype(np.float32)
    # float32 rounding may push a coordinate onto the upper face: keep it strictly inside the alive region
    hi = (np.asarray(ext, dtype=np.float64) + ol).astype(np.float32)
    loc = np.minimum(loc, np.nextafter(hi, np.float32(0)))
    n = ids.size
    p = {"x": loc[:, 0].copy(), "y": loc[:, 1].copy(), "z": loc[:, 2].copy(), "vx": vel[m, 0].copy(), "vy": vel[m, 1].copy(),
         "vz": vel[m, 2].copy(), "mass": np.ones(n, np.float32), "phi": (ids % 97).astype(np.float32),
         "id": ids.astype(np.int64), "mask": (ids % 5).astype(np.uint16)}
    if junk_ghosts:
        rng = np.random.default_rng(seed + 1)
        k = max(n // 10, 1)
        junk = {a: (rng.random(k) * ol * 0.999).astype(np.float32) for a in ("x", "y", "z")}
        for a in ("vx", "vy", "vz", "phi"):
            junk[a] = np.zeros(k, np.float32)
        junk["mass"] = np.ones(k, np.float32)
        junk["id"] = -np.arange(1, k + 1, dtype=np.int64)
        junk["mask"] = np.zeros(k, np.uint16)
        p = {a: np.concatenate([p[a], junk[a]]) for a in p}
    return p


def brute_force_ghost_ids(pos, dims, pos_of_rank, ext, ol):
    """Global ids (with multiplicity of periodic images) of the particles a rank must hold as ghosts: images inside
    its alive region grown by ol but outside the alive region.  float64; valid away from float32 boundary ties."""
    box = np.asarray(dims) * np.asarray(ext, dtype=np.float64)
    lo = np.asarray(pos_of_rank) * np.asarray(ext, dtype=np.float64)
    out = []
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                q = pos + np.array([sx, sy, sz]) * box - lo
                grown = np.all((q >= -ol) & (q <= np.asarray(ext) + ol), axis=1)
                alive = np.all((q >= 0) & (q < np.asarray(ext)), axis=1)
                out.append(np.nonzero(grown & ~alive)[0])
    return np.sort(np.concatenate(out))
