"""GPU tests of the overload refresh (-m gpu): the device kernels of csrc/refresh.cu through the C ABI against the
numpy oracle (bit-exact: the refresh only compares, copies and shifts), on one GPU:
  * single rank, periodic self-exchange through all 26 directions (what `mpirun -np 1` does in the reference,
    ParticleExchange.cxx:676-695) driven by hacc_coral_b200.refresh.overload_refresh;
  * several virtual ranks (one context each) on the same GPU with the messages routed by hand from the same plan,
    which exercises the multi-rank byte layout without NCCL.
The NCCL path proper is tools/refresh_check.py under torchrun (gpurun --gpus N)."""
import numpy as np
import pytest

import hacc_coral_b200 as H
from hacc_coral_b200.refresh import Decomposition, RefreshPlan, overload_refresh
from oracle import refresh_oracle as RO
from tests import refresh_util as U

pytestmark = pytest.mark.gpu

EXT, OL = (20.0, 17.0, 23.0), 3.0
ALO = (OL, OL, OL)
AHI = tuple(OL + e for e in EXT)


def _same(out, want):
    for k in want:
        assert np.array_equal(out[k], want[k]), k


def test_single_rank_periodic_refresh():
    dims = (1, 1, 1)
    pos, vel = U.global_particles(dims, EXT, 60000, seed=5)
    dec = Decomposition(dims, 0)
    p = U.rank_particles(pos, vel, dims, dec.pos, EXT, OL, seed=0)
    want = RO.refresh_all([p], [dec], ALO, AHI, OL)[0]
    g = H.HaccSR(int(want["x"].size) + 1000)
    g.upload(p)
    info = overload_refresh(g, dec, ALO, AHI, OL)
    out = g.download()
    assert info["alive"] == 60000 and info["ghosts"] == want["x"].size - 60000 and g.resident() == want["x"].size
    _same(out, want)
    # ghosts are where the brute-force periodic images are
    assert np.array_equal(np.sort(out["id"][60000:]), U.brute_force_ghost_ids(pos, dims, dec.pos, EXT, OL))
    # a second refresh of the refreshed set is idempotent (ghosts dropped, same ghosts rebuilt)
    info2 = overload_refresh(g, dec, ALO, AHI, OL)
    out2 = g.download()
    g.close()
    assert info2 == info
    _same(out2, want)


@pytest.mark.parametrize("dims", [(2, 1, 1), (2, 2, 2)])
def test_virtual_ranks_on_one_gpu(dims):
    import torch
    size = dims[0] * dims[1] * dims[2]
    pos, vel = U.global_particles(dims, EXT, 8000, seed=6)
    decs = [Decomposition(dims, r) for r in range(size)]
    parts = [U.rank_particles(pos, vel, dims, d.pos, EXT, OL, seed=r) for r, d in enumerate(decs)]
    want = RO.refresh_all(parts, decs, ALO, AHI, OL)
    plans = [RefreshPlan(d) for d in decs]
    ctxs = [H.HaccSR(int(w["x"].size) + 64) for w in want]
    try:
        counts, sendbufs, offs = [], [], []
        for r in range(size):
            ctxs[r].upload(parts[r])
            c, _ = ctxs[r].refresh_begin(ALO, AHI, OL, plans[r].slot_of_dir)
            off, _, total = plans[r].send_layout(c, ctxs[r].refresh_message_bytes)
            buf = torch.empty(max(int(total), 16), dtype=torch.uint8, device="cuda")
            ctxs[r].refresh_pack(off, buf.data_ptr())
            counts.append(c); sendbufs.append(buf); offs.append(off)
        torch.cuda.synchronize()
        for me in range(size):
            for r in range(size):                      # (source rank, direction) order, as overload_refresh appends
                for s, d in enumerate(plans[r].order):
                    if plans[r].dest[s] == me:
                        ctxs[me].refresh_append(sendbufs[r].data_ptr() + int(offs[r][s]), int(counts[r][s]))
            _same(ctxs[me].download(), want[me])
    finally:
        for c in ctxs:
            c.close()


def test_refresh_then_kick_is_deterministic():
    """Messages keep the sender's particle order, so refresh + kick is reproducible bit for bit."""
    dims = (1, 1, 1)
    pos, vel = U.global_particles(dims, EXT, 30000, seed=7)
    dec = Decomposition(dims, 0)
    p = U.rank_particles(pos, vel, dims, dec.pos, EXT, OL, seed=0)
    outs = []
    for _ in range(2):
        g = H.HaccSR(80000)
        g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
        g.upload(p)
        overload_refresh(g, dec, ALO, AHI, OL)
        side = [e + 2 * OL for e in EXT]
        g.kick([0.0] * 3, [max(side)] * 3, [OL] * 3, [s - OL for s in side], 0.5, 128)
        outs.append(g.download())
        g.close()
    _same(outs[0], outs[1])


def test_c_abi_refresh_single_rank():
    """haccsr_refresh (plan, pack, exchange, append in one C call) on a 1 x 1 x 1 decomposition: every neighbour is the rank
    itself, so all 26 messages take the device-local path (ParticleExchange.cxx:676-695); no communicator needed."""
    dims = (1, 1, 1)
    pos, vel = U.global_particles(dims, EXT, 60000, seed=11)
    dec = Decomposition(dims, 0)
    p = U.rank_particles(pos, vel, dims, dec.pos, EXT, OL, seed=0)
    want = RO.refresh_all([p], [dec], ALO, AHI, OL)[0]
    g = H.HaccSR(int(want["x"].size) + 1000)
    g.upload(p)
    info = g.refresh(None, dims, 0, ALO, AHI, OL)
    out = g.download()
    assert info["alive"] == 60000 and info["ghosts"] == want["x"].size - 60000 and g.resident() == want["x"].size
    assert info["bytes_sent_remote"] == 0 and info["bytes_sent"] == info["bytes_received"] and info["ms_total"] > 0
    _same(out, want)
    info2 = g.refresh(None, dims, 0, ALO, AHI, OL)      # idempotent
    out2 = g.download()
    g.close()
    assert {k: v for k, v in info2.items() if k != "ms_total"} == {k: v for k, v in info.items() if k != "ms_total"}
    _same(out2, want)


def test_c_abi_refresh_rejects_bad_arguments():
    g = H.HaccSR(1000)
    with pytest.raises(H.HaccSRError):
        g.refresh(None, (2, 1, 1), 0, ALO, AHI, OL)          # two ranks need a communicator
    with pytest.raises(H.HaccSRError):
        g.refresh(None, (1, 1, 1), 3, ALO, AHI, OL)          # rank outside the decomposition
    g.close()


def test_c_abi_refresh_capacity_is_checked_before_the_exchange():
    """A context too small for its ghosts: haccsr_refresh returns an error that names the rank before anything is exchanged
    (every rank reaches the same verdict from the gathered table), and the context keeps its alive particles."""
    dims = (1, 1, 1)
    pos, vel = U.global_particles(dims, EXT, 30000, seed=13)
    dec = Decomposition(dims, 0)
    p = U.rank_particles(pos, vel, dims, dec.pos, EXT, OL, seed=0)
    m = (p["x"] >= ALO[0]) & (p["x"] < AHI[0]) & (p["y"] >= ALO[1]) & (p["y"] < AHI[1]) & (p["z"] >= ALO[2]) & (p["z"] < AHI[2])
    q = {k: v[m] for k, v in p.items()}
    alive = int(m.sum())
    assert alive == 30000
    g = H.HaccSR(alive + 10)                   # the ghosts (thousands) do not fit
    g.upload(q)
    with pytest.raises(H.HaccSRError, match="on rank 0"):
        g.refresh(None, dims, 0, ALO, AHI, OL)
    assert g.resident() == alive
    g.close()


def test_kick_between_begin_and_pack_is_refused():
    """The candidate list of haccsr_refresh_begin indexes the particles as they lay at that moment (it lives in its own
    buffer, not in the build's scratch); a kick in between permutes them, and haccsr_refresh_pack refuses to pack stale
    indices instead of writing the wrong particles."""
    import torch
    dims = (1, 1, 1)
    pos, vel = U.global_particles(dims, EXT, 20000, seed=12)
    dec = Decomposition(dims, 0)
    p = U.rank_particles(pos, vel, dims, dec.pos, EXT, OL, seed=0)
    plan = RefreshPlan(dec)
    bufs = []
    for kick in (False, True):
        g = H.HaccSR(60000)
        g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
        g.upload(p)
        c, nal = g.refresh_begin(ALO, AHI, OL, plan.slot_of_dir)
        off, _, total = plan.send_layout(c, g.refresh_message_bytes)
        if kick:
            side = [e + 2 * OL for e in EXT]
            g.kick([0.0] * 3, [max(side)] * 3, [OL] * 3, [s - OL for s in side], 0.5, 64, count=nal, skip_force=True)
        buf = torch.zeros(max(int(total), 16), dtype=torch.uint8, device="cuda")
        if not kick:
            g.refresh_pack(off, buf.data_ptr())
            bufs.append(buf.cpu().numpy().copy())
        else:
            # after a kick the particles are in tree order: the candidate indices no longer describe them, and the library
            # must say so instead of packing stale indices
            with pytest.raises(H.HaccSRError):
                g.refresh_pack(off, buf.data_ptr())
        g.close()
