#!/usr/bin/env python
"""Multi-GPU check of one decomposed short-range step (BASELINE config C4 scaled by --global-side), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/step_check.py --global-side 512          # N = 8: 2x2x2 sub-volumes of 256^3 alive particles each

Every rank generates the same global Zel'dovich snapshot (same seed), keeps the alive particles of its sub-volume of
HACC's 3-D decomposition (reference src/halo_finder/Partition.cxx:121-137, src/simulation/Domain.cxx:65-79) in local grid
units, and then
  1. rebuilds its overload (ghost) zone with haccsr_refresh (C ABI) -- device classify + pack, ONE grouped
     ncclSend/ncclRecv, device append (replaces ParticleExchange, src/halo_finder/ParticleExchange.cxx:488-762);
  2. checks the result against the ghost zone extracted directly from the global snapshot (periodic images): same
     multiset of (id, image), positions equal to float32 rounding;
  3. runs the short-range kick (tree build + lists + force kernel) on the refreshed particles and on the directly
     extracted ones and compares the kicks by (id, image): same tree (the centre-of-mass splits are exact sums, so they
     do not depend on particle order), FP32 sums in a different in-leaf order.
Rank 0 prints one JSON line: refresh time and bytes, kick time, G interactions/s over all ranks (max-over-ranks time),
and the parity figures.  Test infrastructure used: none (the comparison is against a direct extraction, not the oracle)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


from tools.decomposed import CART, compare_sets, extract, image_key, make_comm   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-side", type=int, default=256)
    ap.add_argument("--ppn", type=int, default=512)
    ap.add_argument("--ol", type=int, default=11)
    ap.add_argument("--z", type=float, default=50.0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import hacc_coral_b200 as H
    from hacc_coral_b200 import synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = CART[world]
    G, ol = args.global_side, float(args.ol)
    alive, direct, ext = extract(G, dims, rank, ol, "cuda:%d" % local, z=args.z)
    comm = make_comm(local, world, rank, dist if world > 1 else None)

    alo, ahi = (ol,) * 3, tuple(ol + e for e in ext)
    cap = int(direct["x"].size * 1.05) + 4096
    g = H.HaccSR(cap, device=local)
    g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
    g.upload(alive)
    g.refresh(comm, dims, rank, alo, ahi, ol)              # warm-up: NCCL channels, buffers (the refresh is idempotent)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = g.refresh(comm, dims, rank, alo, ahi, ol)
    e1.record()
    torch.cuda.synchronize()
    ms_refresh = e0.elapsed_time(e1)
    got = g.download()
    same_set, dpos, _ = compare_sets(got, direct, ol, ext)

    top = float(max(ext) + 2 * ol)                         # tree box [0, max(nglt)]^3 (Particles.cxx:1213-1216)
    lo, hi = [0.0] * 3, [top] * 3
    flo, fhi = [3.2] * 3, [e + 2 * ol - 3.2 for e in ext]
    g.kick(lo, hi, flo, fhi, 0.5, args.ppn)                # warm
    g.upload(got)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    st = g.kick(lo, hi, flo, fhi, 0.5, args.ppn)
    e1.record()
    torch.cuda.synchronize()
    ms_kick = e0.elapsed_time(e1)
    out_a = g.download()
    g.upload(direct)
    st_b = g.kick(lo, hi, flo, fhi, 0.5, args.ppn)
    out_b = g.download()
    g.close()
    err_med = err_max = -1.0
    n_ties = 0
    if same_set:
        ka, kb = image_key(out_a, ol, ext), image_key(out_b, ol, ext)
        _, oa, ob = np.intersect1d(ka, kb, return_indices=True)
        # boundary ties: a particle whose float32 coordinate lands exactly on a slab face is a ghost in one construction
        # and not in the other (about one per face); particles that can see such a particle (within rmax) are left out
        ties = np.concatenate([np.stack([out_a[a][~np.isin(ka, kb)] for a in ("x", "y", "z")], axis=1),
                               np.stack([out_b[a][~np.isin(kb, ka)] for a in ("x", "y", "z")], axis=1)]).astype(np.float64)
        n_ties = int(ties.shape[0])
        pa = np.stack([out_a[a][oa] for a in ("x", "y", "z")], axis=1).astype(np.float64)
        keep = np.ones(pa.shape[0], dtype=bool)
        for tpos in ties[:256]:
            keep &= ((pa - tpos) ** 2).sum(axis=1) > 3.2 ** 2
        va = np.stack([out_a[k][oa] for k in ("vx", "vy", "vz")], axis=1).astype(np.float64)[keep]
        vb = np.stack([out_b[k][ob] for k in ("vx", "vy", "vz")], axis=1).astype(np.float64)[keep]
        nb = np.sqrt((vb * vb).sum(axis=1))
        rms = np.sqrt((nb * nb).mean())
        d = np.sqrt(((va - vb) ** 2).sum(axis=1))
        err_med, err_max = float(np.median(d) / rms), float(d.max() / rms)
    exact = n_ties == 0 and ka.size == kb.size       # then the two kicks see the same particle set
    ok = bool(same_set and n_ties <= 64 and err_max < 1e-4 and
              (not exact or (st["nodes"] == st_b["nodes"] and st["pairs_evaluated"] == st_b["pairs_evaluated"])))
    t = torch.tensor([ms_refresh, ms_kick, 0.0 if ok else 1.0, dpos, err_med, err_max, float(n_ties)], device="cuda", dtype=torch.float64)
    s = torch.tensor([float(st["pairs_evaluated"]), float(info["bytes_sent"]), float(info["ghosts"]), float(info["alive"])],
                     device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if rank == 0:
        t, s = t.cpu().numpy(), s.cpu().numpy()
        print(json.dumps({
            "check": "decomposed short-range step: overload refresh over NCCL + kick", "n_gpus": world, "dims": list(dims),
            "global_side": G, "alive_per_gpu": int(s[3] / world), "ghosts_per_gpu": int(s[2] / world),
            "refresh_equals_direct_extraction_on_all_ranks": bool(t[2] == 0.0), "max_position_difference": t[3], "boundary_ties_max_per_rank": int(t[6]),
            "kick_refresh_vs_direct_median_rel": t[4], "kick_refresh_vs_direct_max_rel": t[5],
            "ms_refresh_max": t[0], "refresh_GB_sent_total": s[1] / 1e9, "ms_kick_max": t[1],
            "Ginteractions_per_s": s[0] / (t[1] * 1e-3) / 1e9, "ppn": args.ppn}))
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
