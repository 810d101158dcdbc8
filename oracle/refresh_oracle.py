"""numpy restatement of the overload refresh -- TEST INFRASTRUCTURE ONLY (never imported by hacc_coral_b200/).

Follows, in the local grid units of the tree, Particles::copyAliveIntoVectors (reference src/cpu/Particles.cxx:975:
alive = lo <= x < hi), ParticleExchange::identifyExchangeParticles (src/halo_finder/ParticleExchange.cxx:542-574:
a particle not strictly inside [lo+ol, hi-ol] goes to every neighbour whose inclusive slab holds it, slabs from
calculateExchangeRegions :280-450) and ParticleExchange::exchange (:650-762: position shifted into the receiver's
frame, everything else copied, received particles appended).  PARITY UNPINNED against the running reference: the
exchange needs MPI, which is absent here; it is pinned instead by an independent brute-force
property (tests/test_refresh_cpu.py): after a refresh every rank holds exactly the periodic images of the global
particle set that fall inside its alive region grown by the overload width.
"""
import numpy as np

KEYS_F32 = ("x", "y", "z", "vx", "vy", "vz", "mass", "phi")


def dir_vector(d):
    return (d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1)


def alive_mask(p, alo, ahi):
    m = np.ones(p["x"].size, dtype=bool)
    for k, a in enumerate(("x", "y", "z")):
        m &= (p[a] >= np.float32(alo[k])) & (p[a] < np.float32(ahi[k]))
    return m


def classify(p, alo, ahi, ol):
    """dict direction -> indices (ascending) of the alive particles in p that the neighbour in that direction needs."""
    n = p["x"].size
    low, high, full, inner = [], [], [], np.ones(n, dtype=bool)
    for k, a in enumerate(("x", "y", "z")):
        lo, hi = np.float32(alo[k]), np.float32(ahi[k])
        mlo, mhi = np.float32(lo + np.float32(ol)), np.float32(hi - np.float32(ol))
        v = p[a]
        low.append((v >= lo) & (v <= mlo)); high.append((v >= mhi) & (v <= hi)); full.append((v >= lo) & (v <= hi))
        inner &= (v > mlo) & (v < mhi)
    out = {}
    for d in range(27):
        if d == 13:
            continue
        s = dir_vector(d)
        m = ~inner
        for k in range(3):
            m = m & (low[k] if s[k] < 0 else (high[k] if s[k] > 0 else full[k]))
        out[d] = np.nonzero(m)[0]
    return out


def message(p, idx, d, alo, ahi):
    """The particles `idx` of p as the neighbour in direction d receives them."""
    s = dir_vector(d)
    q = {k: np.ascontiguousarray(v[idx]) for k, v in p.items()}
    for k, a in enumerate(("x", "y", "z")):
        ext = np.float32(np.float32(ahi[k]) - np.float32(alo[k]))
        q[a] = (q[a] - np.float32(s[k]) * ext).astype(np.float32)
    return q


def refresh_all(parts, decomps, alo, ahi, ol):
    """Single-process refresh of every rank: parts[r] = particle dict of rank r, decomps[r] = Decomposition of r.
    Ghosts are appended in (source rank, direction) order, like hacc_coral_b200.refresh.overload_refresh."""
    alive = [{k: v[alive_mask(p, alo, ahi)] for k, v in p.items()} for p in parts]
    sends = [classify(a, alo, ahi, ol) for a in alive]
    out = []
    for me, dec in enumerate(decomps):
        pieces = [alive[me]]
        for r, dr in enumerate(decomps):
            for d in range(27):
                if d != 13 and dr.neighbor(d) == me:
                    pieces.append(message(alive[r], sends[r][d], d, alo, ahi))
        out.append({k: np.concatenate([q[k] for q in pieces]) for k in parts[me]})
    return out
