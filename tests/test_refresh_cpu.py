"""CPU tests of the overload refresh: the host plan of hacc_coral_b200/refresh.py, the numpy oracle, and the N > 1
path over torch.distributed with the gloo backend (world_size 2 and 4) using the host engine of tests/refresh_util.py
in place of the device kernels."""
import os
import socket

import numpy as np
import pytest

from hacc_coral_b200.refresh import Decomposition, RefreshPlan, dir_index, dir_vector, opposite, overload_refresh
from oracle import refresh_oracle as RO
from tests import refresh_util as U

EXT, OL = (6.0, 5.0, 7.0), 1.5
ALO = (OL, OL, OL)
AHI = tuple(OL + e for e in EXT)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 1), (4, 3, 2)])
def test_plan_is_consistent(dims):
    size = dims[0] * dims[1] * dims[2]
    plans = [RefreshPlan(Decomposition(dims, r)) for r in range(size)]
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 50, size=(size, 26))
    for me, plan in enumerate(plans):
        assert sorted(plan.order) == [d for d in range(27) if d != 13]
        assert plan.dest == sorted(plan.dest)                      # send buffer is ordered by destination rank
        for d in plan.order:
            assert Decomposition(dims, plan.decomp.neighbor(d)).neighbor(opposite(d)) == me
            assert dir_index(dir_vector(d)) == d
        off, per_rank, total = plan.send_layout(counts[me], U.message_bytes)
        assert total == per_rank.sum() == sum(U.message_bytes(int(c)) for c in counts[me])
        per_src, bytes_from = plan.recv_layout(counts, U.message_bytes)
        # what I expect from r is what r's send layout puts into my chunk
        for r in range(size):
            _, pr, _ = plans[r].send_layout(counts[r], U.message_bytes)
            assert pr[me] == bytes_from[r]
        assert sum(len(m) for m in per_src) == 26                  # every rank receives 26 messages in total


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 2), (3, 1, 2)])
def test_oracle_matches_brute_force(dims):
    pos, vel = U.global_particles(dims, EXT, 400, seed=3)
    size = dims[0] * dims[1] * dims[2]
    decs = [Decomposition(dims, r) for r in range(size)]
    parts = [U.rank_particles(pos, vel, dims, d.pos, EXT, OL, seed=r) for r, d in enumerate(decs)]
    out = RO.refresh_all(parts, decs, ALO, AHI, OL)
    for r, d in enumerate(decs):
        alive = RO.alive_mask(out[r], ALO, AHI)
        n_alive = int(RO.alive_mask(parts[r], ALO, AHI).sum())
        assert np.array_equal(out[r]["id"][:n_alive], parts[r]["id"][:n_alive]) and alive[:n_alive].all()
        assert not alive[n_alive:].any() and (out[r]["id"] >= 0).all()          # stale ghosts dropped
        assert np.array_equal(np.sort(out[r]["id"][n_alive:]), U.brute_force_ghost_ids(pos, dims, d.pos, EXT, OL))
        # ghosts sit where the periodic image of their owner is, in this rank's frame
        box = np.asarray(dims) * np.asarray(EXT)
        lo = np.asarray(d.pos) * np.asarray(EXT)
        g = {k: v[n_alive:] for k, v in out[r].items()}
        q = np.stack([g["x"], g["y"], g["z"]], axis=1).astype(np.float64) - OL + lo
        delta = (q - pos[g["id"]] + box / 2) % box - box / 2
        assert np.abs(delta).max() < 1e-5
        assert np.array_equal(g["vx"], vel[g["id"], 0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, dims, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, vel = U.global_particles(dims, EXT, 300, seed=11)
        decs = [Decomposition(dims, r) for r in range(world)]
        parts = [U.rank_particles(pos, vel, dims, d.pos, EXT, OL, seed=r) for r, d in enumerate(decs)]
        want = RO.refresh_all(parts, decs, ALO, AHI, OL)[rank]
        eng = U.HostEngine(parts[rank])
        info = overload_refresh(eng, decs[rank], ALO, AHI, OL)
        ok = all(np.array_equal(eng.p[k], want[k]) for k in want)
        q.put((rank, ok, info["ghosts"], int(want["x"].size - info["alive"])))
    except Exception as ex:                       # report instead of leaving the parent waiting for the queue
        q.put((rank, False, repr(ex), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(2, 1, 1), (2, 2, 1)])
def test_refresh_over_gloo(dims):
    import torch.multiprocessing as mp
    world = dims[0] * dims[1] * dims[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, dims, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    for rank, ok, ghosts, want_ghosts in res:
        assert ok, "rank %d: refreshed arrays differ from the oracle" % rank
        assert ghosts == want_ghosts and ghosts > 0


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 4)])
def test_c_plan_equals_python_plan(dims):
    """The message plan of haccsr_refresh (C++, csrc/exchange.cu) is the plan of hacc_coral_b200/refresh.py that the gloo tests
    above exercise: same slot order, same destinations, for every rank of the decomposition."""
    from hacc_coral_b200 import capi
    from hacc_coral_b200.refresh import Decomposition, RefreshPlan
    size = dims[0] * dims[1] * dims[2]
    for rank in range(size):
        order, dest = capi.refresh_plan(dims, rank)
        plan = RefreshPlan(Decomposition(dims, rank))
        assert order == list(plan.order) and dest == list(plan.dest), rank
    with pytest.raises(capi.HaccSRError):
        capi.refresh_plan(dims, size)
