"""Overload (ghost-zone) refresh across the GPUs of one box: host-side plan + transport.

Replaces, for the short-range path, what MC3Extras::refreshParticles does at refresh steps
(reference src/simulation/MC3Extras.cxx:660-706) with ParticleExchange (src/halo_finder/ParticleExchange.cxx:488-762)
over the Cartesian topology of Partition (src/halo_finder/Partition.cxx:121-137): every rank keeps its alive
particles, sends the ones within the overload width of a face to the (up to 26) neighbours that need them as ghosts,
and appends what it receives.  The 13 paired MPI send/recv rounds with a barrier each (:630-632,729) become ONE
all-to-all-v over NCCL (NVSwitch makes every peer uniform): the device classifies and packs all 26 messages
into one buffer ordered by destination rank (csrc/refresh.cu), torch.distributed moves it, the device appends.
Messages whose destination is the rank itself (periodic wrap on an axis of extent 1, ParticleExchange.cxx:676-695)
never leave the GPU.

`torch.distributed` is plumbing only (rendezvous + NCCL); packing and unpacking are CUDA kernels of libhaccsr.
"""
import numpy as np

NDIR = 27


def dir_index(s):
    return (s[0] + 1) * 9 + (s[1] + 1) * 3 + (s[2] + 1)


def dir_vector(d):
    return (d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1)


def opposite(d):
    s = dir_vector(d)
    return dir_index((-s[0], -s[1], -s[2]))


class Decomposition:
    """Periodic Cartesian layout of ranks, x slowest (MPI_Cart_create order, Partition.cxx:121-137)."""

    def __init__(self, dims, rank):
        self.dims = tuple(int(d) for d in dims)
        self.size = self.dims[0] * self.dims[1] * self.dims[2]
        assert 0 <= rank < self.size
        self.rank = rank
        self.pos = self.position(rank)

    def position(self, rank):
        nx, ny, nz = self.dims
        return (rank // (ny * nz), (rank // nz) % ny, rank % nz)

    def rank_of(self, pos):
        nx, ny, nz = self.dims
        return ((pos[0] % nx) * ny + (pos[1] % ny)) * nz + (pos[2] % nz)

    def neighbor(self, d, rank=None):
        p = self.pos if rank is None else self.position(rank)
        s = dir_vector(d)
        return self.rank_of((p[0] + s[0], p[1] + s[1], p[2] + s[2]))

    @staticmethod
    def for_world(n):
        """The layouts of the weak-scaling curve: 1, 2x1x1, 2x2x1, 2x2x2."""
        table = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
        if n in table:
            return table[n]
        raise ValueError("no default decomposition for %d ranks" % n)


class RefreshPlan:
    """Message order: slot s = position of direction d when the 26 directions are sorted by (destination rank, d).
    Both sides derive the same order, so a receiver knows the layout of every chunk from the count table alone."""

    def __init__(self, decomp):
        self.decomp = decomp
        dirs = [d for d in range(NDIR) if d != 13]
        self.order = sorted(dirs, key=lambda d: (decomp.neighbor(d), d))       # slot -> direction
        self.slot_of_dir = np.full(NDIR, 26, dtype=np.int32)
        for s, d in enumerate(self.order):
            self.slot_of_dir[d] = s
        self.dest = [decomp.neighbor(d) for d in self.order]                   # slot -> destination rank

    def send_layout(self, counts_by_slot, message_bytes):
        """Byte offset of every slot in the send buffer and the bytes going to every rank."""
        off = np.zeros(NDIR, dtype=np.int64)
        per_rank = np.zeros(self.decomp.size, dtype=np.int64)
        pos = 0
        for s in range(26):
            off[s] = pos
            b = message_bytes(int(counts_by_slot[s]))
            pos += b
            per_rank[self.dest[s]] += b
        return off, per_rank, pos

    def recv_layout(self, all_counts, message_bytes):
        """all_counts[r][s] = particles in slot s of rank r.  Returns, per source rank, the list of
        (direction d, count) of the messages that rank sends to me, in the order they sit in its chunk, and the
        chunk size in bytes."""
        me = self.decomp.rank
        per_src, bytes_from = [], np.zeros(self.decomp.size, dtype=np.int64)
        for r in range(self.decomp.size):
            plan_r = self if r == me else RefreshPlan(Decomposition(self.decomp.dims, r))
            msgs = [(d, int(all_counts[r][s])) for s, d in enumerate(plan_r.order) if plan_r.dest[s] == me]
            per_src.append(msgs)
            bytes_from[r] = sum(message_bytes(n) for _, n in msgs)
        return per_src, bytes_from


def _exchange(recvbuf, sendbuf, recv_bytes, send_bytes, me, group):
    """All-to-all-v of byte chunks.  NCCL: one all_to_all_single.  Other backends (gloo in the CPU tests of the
    host logic): paired isend / irecv per peer; the chunk a rank sends to itself is copied."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recvbuf, sendbuf, output_split_sizes=[int(b) for b in recv_bytes],
                               input_split_sizes=[int(b) for b in send_bytes], group=group)
        return
    so = np.concatenate([[0], np.cumsum(send_bytes)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_bytes)]).astype(np.int64)
    ops = []
    for r in range(len(send_bytes)):
        if r == me:
            recvbuf[ro[r]:ro[r + 1]].copy_(sendbuf[so[r]:so[r + 1]])
            continue
        if send_bytes[r] > 0:
            ops.append(dist.P2POp(dist.isend, sendbuf[so[r]:so[r + 1]], r, group))
        if recv_bytes[r] > 0:
            ops.append(dist.P2POp(dist.irecv, recvbuf[ro[r]:ro[r + 1]], r, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def overload_refresh(g, decomp, alive_lo, alive_hi, ol, group=None):
    """Refresh the ghosts of context `g` (hacc_coral_b200.HaccSR).  Collective over `group` (default world).
    Returns a dict with the particle counts and the bytes moved."""
    import torch
    import torch.distributed as dist
    plan = RefreshPlan(decomp)
    counts, n_alive = g.refresh_begin(alive_lo, alive_hi, ol, plan.slot_of_dir)
    mb = g.refresh_message_bytes
    off, send_bytes, total_send = plan.send_layout(counts, mb)
    dev = torch.device(g.torch_device)
    multi = decomp.size > 1
    if multi:
        mine = torch.tensor(counts[:26], dtype=torch.int64, device=dev)
        table = torch.empty(decomp.size * 26, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table, mine, group=group)
        all_counts = table.cpu().numpy().reshape(decomp.size, 26)
    else:
        all_counts = np.asarray(counts[:26], dtype=np.int64)[None, :]
    per_src, recv_bytes = plan.recv_layout(all_counts, mb)
    sendbuf = torch.empty(max(int(total_send), 16), dtype=torch.uint8, device=dev)
    g.refresh_pack(off, sendbuf.data_ptr())
    total_recv = int(recv_bytes.sum())
    recvbuf = torch.empty(max(total_recv, 16), dtype=torch.uint8, device=dev)
    if multi:
        # one all-to-all-v; chunks are contiguous per destination because slots are sorted by destination rank
        _exchange(recvbuf[:total_recv], sendbuf[:int(total_send)], recv_bytes, send_bytes, decomp.rank, group)
    else:
        recvbuf[:total_recv].copy_(sendbuf[:int(total_send)])
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    # append in (source rank, direction) order: deterministic
    pos, n_ghost = 0, 0
    base = recvbuf.data_ptr()
    for r in range(decomp.size):
        for d, n in per_src[r]:
            g.refresh_append(base + pos, n)
            pos += mb(n)
            n_ghost += n
    return {"alive": int(n_alive), "ghosts": int(n_ghost), "sent": int(sum(counts[:26])), "bytes_sent": int(total_send),
            "bytes_received": total_recv}
