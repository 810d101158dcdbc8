"""ctypes binding of libhaccsr.so (include/haccsr.h).  No CPU fallback: a missing library or a missing
B200 raises HaccSRError."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

LAW_SR_POLY, LAW_SR_FIT, LAW_SR_INTERP, LAW_NEWTON = 0, 1, 2, 3
ARITH_FUSED, ARITH_X86, ARITH_FUSED_RS3 = 0, 1, 2       # include/haccsr.h HACCSR_ARITH_*
# reference src/halo_finder/ForceLaw.cxx:109-114 (== BGQStep16.c:167) and :98-104
POLY5 = np.array([0.269327, -0.0750978, 0.0114808, -0.00109313, 0.0000605491, -0.00000147177], dtype=np.float32)
POLY6 = np.array([0.271431, -0.0783394, 0.0133122, -0.00159485, 0.000132336, -0.00000663394, 0.000000147305],
                 dtype=np.float32)
RMAX = float(np.float32(3.116326355))   # reference ForceLaw.cxx:32


class HaccSRError(RuntimeError):
    pass


class KickStats(C.Structure):
    _fields_ = [("particles", C.c_int64), ("nodes", C.c_int64), ("leaves", C.c_int64),
                ("empty_leaves", C.c_int64), ("max_ppn", C.c_int64), ("mean_ppn", C.c_double),
                ("levels", C.c_int64), ("sink_leaves", C.c_int64), ("list_ranges", C.c_int64),
                ("pseudo_particles", C.c_int64), ("max_list", C.c_int64), ("pairs_evaluated", C.c_uint64),
                ("pairs_in_cutoff", C.c_uint64), ("ms_build", C.c_float), ("ms_walk", C.c_float),
                ("ms_force", C.c_float), ("ms_total", C.c_float), ("force_launches", C.c_int32),
                ("total_launches", C.c_int32), ("pairs_force_law", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class RefreshStats(C.Structure):
    _fields_ = [("alive", C.c_int64), ("ghosts", C.c_int64), ("sent", C.c_int64), ("bytes_sent", C.c_int64),
                ("bytes_received", C.c_int64), ("bytes_sent_remote", C.c_int64), ("messages_received", C.c_int32),
                ("reserved", C.c_int32), ("ms_total", C.c_float)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != "reserved"}


class Map2(C.Structure):
    _fields_ = [("tree_lo", C.c_float * 3), ("tree_hi", C.c_float * 3), ("force_lo", C.c_float * 3), ("force_hi", C.c_float * 3),
                ("fcoeff", C.c_float)]


class KickOpts(C.Structure):
    _fields_ = [("count_in_cutoff", C.c_int32), ("skip_force", C.c_int32), ("reserved", C.c_int32 * 6)]


def lib_path():
    # HACCSR_LIB: another build of the same library (tuning experiments); the default is the in-tree build
    return os.environ.get("HACCSR_LIB") or os.path.join(_HERE, "csrc", "libhaccsr.so")


_LIB = None


def load_library():
    """Load libhaccsr.so and declare every entry point of include/haccsr.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise HaccSRError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    fp, ip64, u16p = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_uint16)
    vp = C.c_void_p
    lib.haccsr_last_error.restype = C.c_char_p
    lib.haccsr_device_count.restype = C.c_int
    lib.haccsr_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int64]
    lib.haccsr_destroy.argtypes = [vp]
    lib.haccsr_set_stream.argtypes = [vp, vp]
    lib.haccsr_set_force_law.argtypes = [vp, C.c_int, fp, C.c_int, C.c_float, C.c_float]
    lib.haccsr_set_arithmetic.argtypes = [vp, C.c_int]
    lib.haccsr_set_culling.argtypes = [vp, C.c_int]
    lib.haccsr_upload.argtypes = [vp, C.c_int64] + [fp] * 8 + [ip64, u16p]
    lib.haccsr_download.argtypes = [vp, C.c_int64] + [fp] * 8 + [ip64, u16p]
    lib.haccsr_host_register.argtypes = [vp, C.c_size_t]
    lib.haccsr_host_unregister.argtypes = [vp]
    lib.haccsr_kick.argtypes = [vp, C.c_int64, fp, fp, fp, fp, C.c_float, C.c_int64, C.c_int, C.c_float,
                                C.POINTER(KickOpts), C.POINTER(KickStats)]
    lib.haccsr_kick_host.argtypes = [vp, C.c_int64] + [fp] * 8 + [ip64, u16p, fp, fp, fp, fp, C.c_float, C.c_int64, C.c_int,
                                                                  C.c_float, C.POINTER(KickOpts), C.POINTER(KickStats)]
    lib.haccsr_stream.argtypes = [vp, C.c_float]
    lib.haccsr_partition_in_box.argtypes = [vp, fp, ip64]
    lib.haccsr_fill_mass.argtypes = [vp, C.c_float]
    lib.haccsr_subcycle.argtypes = [vp, C.c_int, C.c_float, fp, fp, fp, fp, fp, C.c_float, C.c_int64, C.c_int, C.c_float,
                                    C.POINTER(KickStats)]
    i32p = C.POINTER(C.c_int32)
    lib.haccsr_map2_setup.argtypes = [i32p, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, C.POINTER(Map2)]
    lib.haccsr_map1_factor.restype = C.c_float
    lib.haccsr_map1_factor.argtypes = [C.c_float] * 4
    lib.haccsr_particles_subcycle.argtypes = [vp, C.c_int, i32p, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_float, C.c_int64, C.c_int, C.POINTER(KickStats)]
    lib.haccsr_cic.argtypes = [vp, i32p, C.c_float, vp, C.c_int]
    lib.haccsr_inverse_cic.argtypes = [vp, i32p, vp, C.c_int, C.c_float, C.c_float, C.c_int]
    lib.haccsr_refresh_message_bytes.restype = C.c_int64
    lib.haccsr_refresh_message_bytes.argtypes = [C.c_int64]
    lib.haccsr_refresh_begin.argtypes = [vp, fp, fp, C.c_float, i32p, ip64, ip64]
    lib.haccsr_refresh_pack.argtypes = [vp, ip64, vp]
    lib.haccsr_refresh_append.argtypes = [vp, vp, C.c_int64]
    lib.haccsr_refresh.argtypes = [vp, vp, i32p, C.c_int32, fp, fp, C.c_float, C.POINTER(RefreshStats)]
    lib.haccsr_refresh_plan.argtypes = [i32p, C.c_int32, i32p, i32p]
    lib.haccsr_nccl_unique_id.argtypes = [vp]
    lib.haccsr_nccl_comm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
    lib.haccsr_nccl_comm_destroy.argtypes = [vp]
    lib.haccsr_resident.restype = C.c_int64
    lib.haccsr_resident.argtypes = [vp]
    lib.haccsr_get_tree.argtypes = [vp, C.c_int64, ip64, i32p, i32p, i32p, i32p, fp]
    lib.haccsr_get_pseudo_particles.argtypes = [vp, C.c_int64, fp]
    u32p = C.POINTER(C.c_uint32)
    lib.haccsr_get_lists.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, ip64, ip64, ip64, u32p, u32p, fp]
    _LIB = lib
    return lib


EXPORTS = ["haccsr_last_error", "haccsr_device_count", "haccsr_create", "haccsr_destroy", "haccsr_set_stream",
           "haccsr_set_force_law", "haccsr_set_arithmetic", "haccsr_set_culling", "haccsr_upload", "haccsr_download", "haccsr_host_register",
           "haccsr_host_unregister", "haccsr_kick", "haccsr_kick_host", "haccsr_stream", "haccsr_partition_in_box",
           "haccsr_fill_mass", "haccsr_subcycle", "haccsr_map2_setup", "haccsr_map1_factor", "haccsr_particles_subcycle", "haccsr_cic", "haccsr_inverse_cic", "haccsr_refresh_message_bytes", "haccsr_refresh_begin",
           "haccsr_refresh_pack", "haccsr_refresh_append", "haccsr_refresh", "haccsr_refresh_plan", "haccsr_nccl_unique_id", "haccsr_nccl_comm_create",
           "haccsr_nccl_comm_destroy", "haccsr_resident", "haccsr_get_tree", "haccsr_get_pseudo_particles",
           "haccsr_get_lists"]

_F32 = ("x", "y", "z", "vx", "vy", "vz", "mass", "phi")


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


class HaccSR:
    """One context on one GPU: upload -> [stream, kick, stream]* -> download."""

    def __init__(self, max_particles, device=0, arith=None):
        self.lib = load_library()
        h = C.c_void_p()
        self._h = None
        self._check(self.lib.haccsr_create(C.byref(h), device, int(max_particles)))
        self._h = h
        if arith is not None:
            self.set_arithmetic(arith)
        self.n = 0
        self.device = device
        self.torch_device = "cuda:%d" % device

    def _check(self, rc):
        if rc != 0:
            raise HaccSRError("libhaccsr: %s (status %d)" % (self.lib.haccsr_last_error().decode(), rc))

    def close(self):
        if self._h is not None:
            self.lib.haccsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.haccsr_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_culling(self, on):
        self._check(self.lib.haccsr_set_culling(self._h, int(bool(on))))

    def set_arithmetic(self, mode):
        self._check(self.lib.haccsr_set_arithmetic(self._h, int(mode)))

    def set_force_law(self, kind=LAW_SR_POLY, coeffs=POLY5, rsm=0.007, rmax=RMAX):
        coeffs = np.ascontiguousarray(coeffs if coeffs is not None else [], dtype=np.float32)
        self._check(self.lib.haccsr_set_force_law(self._h, kind, _fp(coeffs), int(coeffs.size), rsm, rmax))

    def upload(self, p):
        n = int(np.asarray(p["x"]).size)
        a = {k: np.ascontiguousarray(p[k], dtype=np.float32) for k in _F32 if k in p}
        ids = np.ascontiguousarray(p["id"], dtype=np.int64) if "id" in p else None
        mask = np.ascontiguousarray(p["mask"], dtype=np.uint16) if "mask" in p else None
        self._check(self.lib.haccsr_upload(
            self._h, n, *[_fp(a.get(k)) for k in _F32],
            ids.ctypes.data_as(C.POINTER(C.c_int64)) if ids is not None else None,
            mask.ctypes.data_as(C.POINTER(C.c_uint16)) if mask is not None else None))
        self.n = n

    def download(self, n=None, out=None):
        n = self.n if n is None else n
        if out is None:
            out = {k: np.empty(n, dtype=np.float32) for k in _F32}
            out["id"] = np.empty(n, dtype=np.int64)
            out["mask"] = np.empty(n, dtype=np.uint16)
        self._check(self.lib.haccsr_download(
            self._h, n, *[_fp(out.get(k)) for k in _F32],
            out["id"].ctypes.data_as(C.POINTER(C.c_int64)) if out.get("id") is not None else None,
            out["mask"].ctypes.data_as(C.POINTER(C.c_uint16)) if out.get("mask") is not None else None))
        return out

    def kick(self, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, fcoeff=1.0, count=None,
             count_in_cutoff=False, skip_force=False, tdpts=1):
        st = KickStats()
        opts = KickOpts()
        opts.count_in_cutoff = int(count_in_cutoff)
        opts.skip_force = int(skip_force)
        n = self.n if count is None else count
        self._check(self.lib.haccsr_kick(self._h, n, _f3(tree_lo), _f3(tree_hi), _f3(force_lo), _f3(force_hi),
                                         theta, int(ppn), tdpts, fcoeff, C.byref(opts), C.byref(st)))
        return st.as_dict()

    def kick_host(self, p, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, fcoeff=1.0, count_in_cutoff=False, tdpts=1):
        """upload + kick + download on the numpy arrays of dict `p` IN PLACE (float32 / int64 / uint16, contiguous):
        the call a user of the reference's constructor makes."""
        for k in _F32:
            assert p[k].dtype == np.float32 and p[k].flags["C_CONTIGUOUS"], k
        n = int(p["x"].size)
        ids, mask = p.get("id"), p.get("mask")
        st, opts = KickStats(), KickOpts()
        opts.count_in_cutoff = int(count_in_cutoff)
        self._check(self.lib.haccsr_kick_host(
            self._h, n, *[_fp(p[k]) for k in _F32],
            ids.ctypes.data_as(C.POINTER(C.c_int64)) if ids is not None else None,
            mask.ctypes.data_as(C.POINTER(C.c_uint16)) if mask is not None else None,
            _f3(tree_lo), _f3(tree_hi), _f3(force_lo), _f3(force_hi), theta, int(ppn), tdpts, fcoeff,
            C.byref(opts), C.byref(st)))
        self.n = n
        return st.as_dict()

    def stream(self, prefactor_tau):
        self._check(self.lib.haccsr_stream(self._h, prefactor_tau))

    def fill_mass(self, value=1.0):
        self._check(self.lib.haccsr_fill_mass(self._h, value))

    def partition_in_box(self, hi):
        nin = C.c_int64()
        self._check(self.lib.haccsr_partition_in_box(self._h, _f3(hi), C.byref(nin)))
        return nin.value

    def subcycle(self, nsub, prefactor_tau, box_hi, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, fcoeff, tdpts=1):
        """Particles::subCycle on the resident particles; returns the summed stats."""
        st = KickStats()
        self._check(self.lib.haccsr_subcycle(self._h, int(nsub), prefactor_tau, _f3(box_hi), _f3(tree_lo), _f3(tree_hi),
                                             _f3(force_lo), _f3(force_hi), theta, int(ppn), tdpts, fcoeff, C.byref(st)))
        return st.as_dict()

    def particles_subcycle(self, nsub, nglt, edge, gpscal, alpha, pp, adot, tau, tau2, fscal, theta, ppn, tdpts=1):
        """Particles::subCycle from the TimeStepper's scalars (haccsr_particles_subcycle)."""
        st = KickStats()
        n3 = (C.c_int32 * 3)(*[int(t) for t in nglt])
        self._check(self.lib.haccsr_particles_subcycle(self._h, int(nsub), n3, edge, gpscal, alpha, pp, adot, tau, tau2, fscal,
                                                       theta, int(ppn), tdpts, C.byref(st)))
        return st.as_dict()

    # ---- PM coupling (csrc/cic.cu) ----
    def cic(self, ng, c, out=None):
        """Particles::cic on the resident particles; returns the (ng0, ng1, ng2) float32 density grid (written into `out`,
        e.g. a page-locked array, when given)."""
        ng3 = (C.c_int32 * 3)(*[int(t) for t in ng])
        rho = out if out is not None else np.empty(tuple(int(t) for t in ng), dtype=np.float32)
        assert rho.dtype == np.float32 and rho.flags["C_CONTIGUOUS"] and rho.size == int(np.prod([int(t) for t in ng]))
        self._check(self.lib.haccsr_cic(self._h, ng3, float(c), C.c_void_p(rho.ctypes.data), 0))
        return rho

    def inverse_cic(self, grid, tau, fscal, comp):
        """Particles::inverse_cic: v[comp] += interp(grid) * fscal * tau on the resident particles."""
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        ng3 = (C.c_int32 * 3)(*grid.shape)
        self._check(self.lib.haccsr_inverse_cic(self._h, ng3, C.c_void_p(grid.ctypes.data), 0, float(tau), float(fscal), int(comp)))

    # ---- overload refresh (device part; hacc_coral_b200/refresh.py drives it) ----
    def refresh_message_bytes(self, n):
        return int(self.lib.haccsr_refresh_message_bytes(int(n)))

    def refresh_begin(self, alive_lo, alive_hi, ol, slot_of_dir):
        sod = np.ascontiguousarray(slot_of_dir, dtype=np.int32)
        counts = np.zeros(27, dtype=np.int64)
        nal = C.c_int64()
        self._check(self.lib.haccsr_refresh_begin(self._h, _f3(alive_lo), _f3(alive_hi), ol,
                                                  sod.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  counts.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nal)))
        self.n = nal.value
        return counts, nal.value

    def refresh_pack(self, byte_off_by_slot, sendbuf_ptr):
        off = np.ascontiguousarray(byte_off_by_slot, dtype=np.int64)
        self._check(self.lib.haccsr_refresh_pack(self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(sendbuf_ptr)))

    def refresh_append(self, message_ptr, n):
        self._check(self.lib.haccsr_refresh_append(self._h, C.c_void_p(message_ptr), int(n)))
        self.n = self.resident()

    def refresh(self, comm, dims, rank, alive_lo, alive_hi, ol):
        """haccsr_refresh: the whole overload refresh of this rank (collective over `comm`, an NcclComm or None for 1x1x1)."""
        st = RefreshStats()
        d3 = (C.c_int32 * 3)(*[int(t) for t in dims])
        self._check(self.lib.haccsr_refresh(self._h, comm.handle if comm is not None else None, d3, int(rank), _f3(alive_lo),
                                            _f3(alive_hi), float(ol), C.byref(st)))
        self.n = self.resident()
        return st.as_dict()

    def resident(self):
        return int(self.lib.haccsr_resident(self._h))

    def tree(self):
        nn = C.c_int64()
        self._check(self.lib.haccsr_get_tree(self._h, 0, C.byref(nn), None, None, None, None, None))
        m = nn.value
        t = {k: np.empty(m, dtype=np.int32) for k in ("count", "offset", "cl", "cr")}
        box = np.empty((m, 10), dtype=np.float32)
        i32p = C.POINTER(C.c_int32)
        self._check(self.lib.haccsr_get_tree(self._h, m, C.byref(nn), t["count"].ctypes.data_as(i32p),
                                             t["offset"].ctypes.data_as(i32p), t["cl"].ctypes.data_as(i32p),
                                             t["cr"].ctypes.data_as(i32p), _fp(box)))
        t["xmin"], t["xmax"], t["xc"], t["ppm"] = box[:, 0:3], box[:, 3:6], box[:, 6:9], box[:, 9]
        return t

    def pseudo_particles(self):
        """(nodes, 12, 4) array of the quadrupole pseudo-particles (x, y, z, mass) of the last tdpts = 12 kick."""
        nn = C.c_int64()
        self._check(self.lib.haccsr_get_tree(self._h, 0, C.byref(nn), None, None, None, None, None))
        pp = np.empty((nn.value, 12, 4), dtype=np.float32)
        self._check(self.lib.haccsr_get_pseudo_particles(self._h, nn.value, _fp(pp)))
        return pp

    def lists(self):
        nn, nr, npool = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.haccsr_get_lists(self._h, 0, 0, 0, C.byref(nn), C.byref(nr), C.byref(npool),
                                              None, None, None))
        off = np.empty(nn.value + 1, dtype=np.uint32)
        ranges = np.empty((max(nr.value, 1), 2), dtype=np.uint32)
        pool = np.empty((max(npool.value, 1), 4), dtype=np.float32)
        u32p = C.POINTER(C.c_uint32)
        self._check(self.lib.haccsr_get_lists(self._h, nn.value + 1, max(nr.value, 1), max(npool.value, 1),
                                              C.byref(nn), C.byref(nr), C.byref(npool), off.ctypes.data_as(u32p),
                                              ranges.ctypes.data_as(u32p), _fp(pool)))
        return {"range_off": off, "ranges": ranges[:nr.value], "pool": pool[:npool.value]}


class NcclComm:
    """An ncclComm_t created by libhaccsr (haccsr_nccl_comm_create) for haccsr_refresh.  `broadcast` moves the 128-byte id
    from rank 0 to every rank: a callable bytes -> bytes (torch.distributed in bench.py, MPI_Bcast in a HACC rank)."""

    def __init__(self, device, nranks, rank, broadcast):
        self.lib = load_library()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            rc = self.lib.haccsr_nccl_unique_id(buf)
            if rc != 0:
                raise HaccSRError("libhaccsr: %s (status %d)" % (self.lib.haccsr_last_error().decode(), rc))
        ident = broadcast(bytes(buf))
        buf = (C.c_ubyte * 128).from_buffer_copy(ident)
        h = C.c_void_p()
        rc = self.lib.haccsr_nccl_comm_create(C.byref(h), int(device), int(nranks), int(rank), buf)
        if rc != 0:
            raise HaccSRError("libhaccsr: %s (status %d)" % (self.lib.haccsr_last_error().decode(), rc))
        self.handle = h

    def close(self):
        if self.handle:
            self.lib.haccsr_nccl_comm_destroy(self.handle)
            self.handle = None


def map2_setup(nglt, edge, gpscal, fscal, tau, step_fraction):
    """haccsr_map2_setup: the boxes and the kick coefficient Particles::map2 hands to the tree (host arithmetic only)."""
    lib = load_library()
    m = Map2()
    n3 = (C.c_int32 * 3)(*[int(t) for t in nglt])
    rc = lib.haccsr_map2_setup(n3, edge, gpscal, fscal, tau, step_fraction, C.byref(m))
    if rc != 0:
        raise HaccSRError(lib.haccsr_last_error().decode())
    return {"tree_lo": list(m.tree_lo), "tree_hi": list(m.tree_hi), "force_lo": list(m.force_lo), "force_hi": list(m.force_hi),
            "fcoeff": float(m.fcoeff)}


def map1_factor(pp, tau, adot, alpha):
    return float(load_library().haccsr_map1_factor(pp, tau, adot, alpha))


def refresh_plan(dims, rank):
    """haccsr_refresh_plan: (direction of every message slot, destination rank of every slot) as haccsr_refresh orders them."""
    lib = load_library()
    d3 = (C.c_int32 * 3)(*[int(t) for t in dims])
    order, dest = (C.c_int32 * 26)(), (C.c_int32 * 26)()
    rc = lib.haccsr_refresh_plan(d3, int(rank), order, dest)
    if rc != 0:
        raise HaccSRError(lib.haccsr_last_error().decode())
    return list(order), list(dest)
