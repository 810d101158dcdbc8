#!/usr/bin/env python
"""Multi-GPU check of the overload refresh over NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/refresh_check.py [--per-rank 200000]

Every rank builds its part of a seeded periodic global particle set, refreshes its ghosts with
hacc_coral_b200.refresh.overload_refresh (device classify + pack, ONE all_to_all_single over NCCL, device append) and
compares all ten arrays bit for bit with the numpy oracle (test infrastructure) evaluated for all ranks in-process;
rank 0 prints one JSON line with the timing of the refresh (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-rank", type=int, default=200000)
    ap.add_argument("--ext", type=float, default=64.0)
    ap.add_argument("--ol", type=float, default=8.0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import hacc_coral_b200 as H
    from hacc_coral_b200.refresh import Decomposition, overload_refresh
    from oracle import refresh_oracle as RO
    from tests import refresh_util as U
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = Decomposition.for_world(world)
    ext, ol = (args.ext,) * 3, args.ol
    alo, ahi = (ol,) * 3, (ol + args.ext,) * 3
    pos, vel = U.global_particles(dims, ext, args.per_rank, seed=2024)
    decs = [Decomposition(dims, r) for r in range(world)]
    parts = [U.rank_particles(pos, vel, dims, d.pos, ext, ol, seed=r) for r, d in enumerate(decs)]
    want = RO.refresh_all(parts, decs, alo, ahi, ol)[rank]
    g = H.HaccSR(int(want["x"].size) + 1024, device=local)
    g.upload(parts[rank])
    overload_refresh(g, decs[rank], alo, ahi, ol)          # warm-up (NCCL channels, allocations); idempotent
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = overload_refresh(g, decs[rank], alo, ahi, ol)
    e1.record()
    torch.cuda.synchronize()
    out = g.download()
    g.close()
    ok = all(np.array_equal(out[k], want[k]) for k in want)
    t = torch.tensor([e0.elapsed_time(e1), 0.0 if ok else 1.0, float(info["bytes_sent"]), float(info["ghosts"])],
                     device="cuda", dtype=torch.float64)
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"check": "overload refresh over NCCL", "n_gpus": world, "dims": dims, "alive_per_rank": args.per_rank,
                          "bit_identical_to_oracle_on_all_ranks": bool(tm[1].item() == 0.0), "ms_refresh_max": tm[0].item(),
                          "ghosts_total": int(ts[3].item()), "bytes_sent_total": int(ts[2].item()),
                          "GB_per_s_per_gpu": ts[2].item() / world / (tm[0].item() * 1e-3) / 1e9}))
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
