"""Harness for the decomposed (multi-GPU) short-range step: one global Zel'dovich snapshot cut into HACC's 3-D domain
decomposition (reference src/halo_finder/Partition.cxx:121-137, src/simulation/Domain.cxx:65-79), one rank per GPU.

Used by bench.py (side block `refresh`) and tools/step_check.py.  Every rank generates the same global snapshot (same seed),
keeps the alive particles of its sub-volume in local grid units, rebuilds its overload zone with haccsr_refresh (C ABI:
device classify + pack, one grouped ncclSend/ncclRecv over NVLink, device append) and compares the result with the overload
zone extracted directly from the global snapshot (periodic images).  Test infrastructure used: none (the comparison is against
a direct extraction, not the oracle)."""
import numpy as np

CART = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}      # the layouts of the weak-scaling curve


def cart_position(dims, rank):
    return (rank // (dims[1] * dims[2]), (rank // dims[2]) % dims[1], rank % dims[2])


def image_key(p, ol, ext):
    """id * 27 + code of the periodic image the particle is (which side of the alive region, per dimension)."""
    code = np.zeros(p["x"].size, dtype=np.int64)
    for k, a in enumerate(("x", "y", "z")):
        s = np.where(p[a] < np.float32(ol), 0, np.where(p[a] >= np.float32(ol + ext[k]), 2, 1))
        code = code * 3 + s
    return p["id"].astype(np.int64) * 27 + code


def near_boundary(p, ol, ext, eps=1e-3):
    """Particles within eps of a plane where float32 rounding decides membership (outer ghost faces, alive faces): the two
    constructions may legitimately disagree on those (about one particle per face at 256^2 cells), so they are left out
    of the set comparison."""
    m = np.zeros(p["x"].size, dtype=bool)
    for k, a in enumerate(("x", "y", "z")):
        for plane in (0.0, ol, ol + ext[k], 2 * ol + ext[k]):
            m |= np.abs(p[a].astype(np.float64) - plane) < eps
    return m


def make_comm(local, world, rank, dist):
    """An NCCL communicator owned by libhaccsr; the 128-byte id travels over torch.distributed (a HACC rank would MPI_Bcast it)."""
    import torch
    from hacc_coral_b200.capi import NcclComm
    if world == 1:
        return None

    def bcast(b):
        t = torch.tensor(list(b), dtype=torch.uint8, device="cuda:%d" % local)
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tolist())
    return NcclComm(local, world, rank, bcast)


def extract(global_shape, dims, rank, ol, device, z=50.0, seed=5009888, growth_boost=1.0, with_direct=True):
    """(alive particles of `rank` in local grid units, directly extracted alive + ghosts or None, sub-volume extent).
    global_shape = particles per dimension of the periodic global box (int or triple)."""
    import torch
    from hacc_coral_b200 import synth
    gs = (int(global_shape),) * 3 if np.isscalar(global_shape) else tuple(int(t) for t in global_shape)
    ext = [g // d for g, d in zip(gs, dims)]
    assert all(e * d == g for e, d, g in zip(ext, dims, gs))
    pr = cart_position(dims, rank)
    origin = np.array([pr[k] * ext[k] for k in range(3)], dtype=np.float64)
    dev = torch.device(device)
    pos = synth.zeldovich_torch(gs, z=z, seed=seed, ghost=0, growth_boost=growth_boost, device=dev, return_tensor=True)
    box = torch.as_tensor(np.asarray(gs, dtype=np.float64), device=dev)
    t_origin = torch.as_tensor(origin, device=dev)
    exta = torch.as_tensor(np.asarray(ext, dtype=np.float64), device=dev)
    ahi32 = (exta + ol).to(torch.float32)
    ahi32m = torch.nextafter(ahi32, torch.zeros_like(ahi32))
    top32 = torch.nextafter((exta + 2 * ol).to(torch.float32), torch.zeros_like(ahi32))

    def pack(loc32, ids):
        loc = loc32.cpu().numpy()
        p = synth._pack(loc[:, 0], loc[:, 1], loc[:, 2])
        p["id"] = ids.cpu().numpy().astype(np.int64)
        return p

    m = ((pos >= t_origin) & (pos < t_origin + exta)).all(dim=1)
    alive = pack(torch.minimum((pos[m] - t_origin + ol).to(torch.float32), ahi32m), torch.nonzero(m).reshape(-1))
    direct = None
    if with_direct:
        # every periodic image inside the alive region grown by ol, with the float32 arithmetic of the exchange (the owner's
        # local float32 coordinate shifted by a whole number of sub-volume extents, ParticleExchange.cxx:672-673,702-708),
        # so that both constructions hold bit-identical positions
        owner_origin = torch.floor(pos / exta) * exta
        owner_local = torch.minimum((pos - owner_origin + ol).to(torch.float32), ahi32m)
        pieces, ids = [], []
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    shift = torch.tensor([sx, sy, sz], device=dev, dtype=torch.float64) * box
                    q = pos + shift - t_origin + ol
                    mm = ((q >= 0) & (q < exta + 2 * ol)).all(dim=1)
                    delta = (owner_origin[mm] + shift - t_origin).to(torch.float32)        # -ext, 0 or +ext per dimension
                    pieces.append(torch.minimum(owner_local[mm] + delta, top32)); ids.append(torch.nonzero(mm).reshape(-1))
                    del q, mm, delta
        direct = pack(torch.cat(pieces), torch.cat(ids))
        del pieces, ids, owner_origin, owner_local
    del pos, m
    torch.cuda.empty_cache()
    return alive, direct, ext


def compare_sets(got, direct, ol, ext):
    """(same multiset of (id, image) away from float32 boundary ties, largest position difference on the common set)."""
    ka, kb = image_key(got, ol, ext), image_key(direct, ol, ext)
    na, nb_ = near_boundary(got, ol, ext), near_boundary(direct, ol, ext)
    same = np.setxor1d(ka[~na], kb[~nb_]).size == 0 and abs(int(ka.size) - int(kb.size)) <= 64
    _, ia, ib = np.intersect1d(ka, kb, return_indices=True)
    dpos = max(float(np.abs(got[a][ia].astype(np.float64) - direct[a][ib]).max()) for a in ("x", "y", "z")) if ia.size else -1.0
    return bool(same), dpos, int(np.setxor1d(ka, kb).size)


def refresh_block(np_side, ol, rank, world, local, dist, reps=3, peer_copy_gbs=770.0):
    """bench.py side block: the overload refresh of one global (np_side * dims)^3 snapshot through haccsr_refresh."""
    import torch
    import hacc_coral_b200 as H
    dims = CART[world]
    G = [np_side * d for d in dims]        # every rank holds np_side^3 alive particles: the global box grows with the layout
    alive, direct, ext = extract(G, dims, rank, float(ol), "cuda:%d" % local)
    alo, ahi = (float(ol),) * 3, tuple(float(ol) + e for e in ext)
    comm = make_comm(local, world, rank, dist)
    g = H.HaccSR(int(direct["x"].size * 1.05) + 4096, device=local)
    g.upload(alive)
    g.refresh(comm, dims, rank, alo, ahi, float(ol))                # warm-up: NCCL channels, buffers (the refresh is idempotent)
    ms, info = [], None
    for _ in range(reps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        info = g.refresh(comm, dims, rank, alo, ahi, float(ol))
        ms.append(info["ms_total"])
    got = g.download()
    g.close()
    if comm is not None:
        comm.close()
    same, dpos, nties = compare_sets(got, direct, float(ol), ext)
    t = torch.tensor([float(np.median(ms)), 0.0 if same else 1.0, dpos, float(nties)], device="cuda:%d" % local, dtype=torch.float64)
    s = torch.tensor([float(info["bytes_sent_remote"]), float(info["bytes_sent"]), float(info["ghosts"]), float(info["alive"])],
                     device="cuda:%d" % local, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    t, s = t.cpu().numpy(), s.cpu().numpy()
    remote_per_gpu = s[0] / world
    return {"dims": list(dims), "global_side": G, "alive_per_gpu": int(s[3] / world), "ghosts_per_gpu": int(s[2] / world),
            "ms": float(t[0]), "bytes_sent_per_gpu": s[1] / world, "bytes_over_nvlink_per_gpu": remote_per_gpu,
            "nvlink_GBps_per_gpu": remote_per_gpu / (t[0] * 1e-3) / 1e9 if t[0] > 0 else None,
            "peer_copy_reference_GBps": peer_copy_gbs,
            "equals_direct_extraction_on_all_ranks": bool(t[1] == 0.0), "max_position_difference": float(t[2]),
            "boundary_ties_max_per_rank": int(t[3]),
            "path": "haccsr_refresh (C ABI): device classify + pack, one grouped ncclSend/ncclRecv, device append; ms = whole call, device events, max over ranks"}
