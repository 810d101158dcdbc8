"""ctypes binding of oracle/_ref/libhaccref*.so -- TEST INFRASTRUCTURE ONLY.

The shared objects are the reference's own hot-path sources compiled by oracle/build_ref.sh
(reference src/halo_finder/RCBForceTree.cxx, ForceLaw.cxx, BGQCM.c, bigchunk.c) behind
oracle/ref_harness.cxx.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (hacc_coral_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

LAW_POLY5, LAW_POLY6, LAW_FIT, LAW_NEWTON, LAW_INTERP = 0, 1, 2, 3, 4
# 5th-order grid-force polynomial (reference ForceLaw.cxx:109-114 == BGQStep16.c:167)
POLY5 = np.array([0.269327, -0.0750978, 0.0114808, -0.00109313, 0.0000605491, -0.00000147177],
                 dtype=np.float32)


class RefStats(C.Structure):
    _fields_ = [("nodes", C.c_int64), ("leaves", C.c_int64), ("empty_leaves", C.c_int64),
                ("max_ppn", C.c_int64), ("mean_ppn", C.c_double), ("wall_s", C.c_double),
                ("pairs_eval", C.c_uint64), ("pairs_incut", C.c_uint64)]


_libs = {}


def available(vmax=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libhaccref_vmax.so" if vmax else "libhaccref.so"))


def _lib(vmax=False):
    key = bool(vmax)
    if key not in _libs:
        # the walk keeps 4*VMAX floats on each OpenMP worker's stack (RCBForceTree.cxx:940)
        os.environ.setdefault("OMP_STACKSIZE", "64M")
        path = os.path.join(_HERE, "_ref", "libhaccref_vmax.so" if vmax else "libhaccref.so")
        lib = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        lib.ref_rmax.restype = C.c_float
        lib.ref_force_law_eval.argtypes = [C.c_int, fp, C.c_int, C.c_float, C.c_int64, fp, fp]
        lib.ref_rcb_kick.argtypes = ([C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int64] + [fp] * 8 +
                                     [C.POINTER(C.c_int64), C.POINTER(C.c_uint16), fp, C.c_float,
                                      C.c_float, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_int,
                                      C.POINTER(RefStats)])
        lib.ref_fgrid_table.argtypes = [C.c_int, fp]
        lib.ref_fgrid_constants.argtypes = [fp]
        lib.ref_tree_size.restype = C.c_int64
        ip = C.POINTER(C.c_int64)
        lib.ref_tree_get.argtypes = [C.c_int64, ip, ip, ip, ip, fp]
        lib.ref_set_tdpts.argtypes = [C.c_int]
        lib.ref_tree_get_pp12.argtypes = [C.c_int64, fp]
        _libs[key] = lib
    return _libs[key]


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def rmax():
    return float(_lib().ref_rmax())


def fgrid_table(n):
    """The reference's grid-force interpolation table of n samples (FGrid::fgor_r2_interp)."""
    out = np.empty(n, dtype=np.float32)
    assert _lib().ref_fgrid_table(n, _fp(out)) == 0
    return out


def fgrid_constants():
    """b c d e f g h l of the reference's analytic grid-force fit (FGrid, ForceLaw.cxx:23-31)."""
    out = np.empty(8, dtype=np.float32)
    assert _lib().ref_fgrid_constants(_fp(out)) == 0
    return out


def force_law_eval(law, r2, rsm, coef=POLY5):
    r2 = np.ascontiguousarray(r2, dtype=np.float32)
    out = np.empty_like(r2)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    rc = _lib().ref_force_law_eval(law, _fp(coef), len(coef), rsm, r2.size, _fp(r2), _fp(out))
    assert rc == 0
    return out


def rcb_kick(p, tree_lo, tree_hi, force_lo, force_hi, rsm, theta, ppn, fcoeff=1.0, law=LAW_POLY5,
             coef=POLY5, ds=2, tmin=128, count_pairs=False, keep_tree=False, vmax=False, quiet=True, tdpts=1):
    """Run the reference RCBMonopoleForceTree constructor on a copy of particle dict `p`
    (keys x y z vx vy vz mass phi id mask).  Returns (particles in reference tree order, stats dict,
    tree dict or None)."""
    lib = _lib(vmax)
    q = {k: np.ascontiguousarray(p[k]).copy() for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "phi")}
    q["id"] = np.ascontiguousarray(p["id"], dtype=np.int64).copy()
    q["mask"] = np.ascontiguousarray(p["mask"], dtype=np.uint16).copy()
    for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "phi"):
        assert q[k].dtype == np.float32
    n = q["x"].size
    boxes = np.array(list(tree_lo) + list(tree_hi) + list(force_lo) + list(force_hi), dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    st = RefStats()
    lib.ref_set_tdpts(int(tdpts))       # 1: RCBMonopoleForceTree, 12: RCBQuadrupoleForceTree

    def call():
        return lib.ref_rcb_kick(law, _fp(coef), len(coef), int(count_pairs), int(quiet), n,
                                _fp(q["x"]), _fp(q["y"]), _fp(q["z"]), _fp(q["vx"]), _fp(q["vy"]), _fp(q["vz"]),
                                _fp(q["mass"]), _fp(q["phi"]), q["id"].ctypes.data_as(C.POINTER(C.c_int64)),
                                q["mask"].ctypes.data_as(C.POINTER(C.c_uint16)), _fp(boxes), rsm, theta,
                                ppn, ds, tmin, fcoeff, int(keep_tree), C.byref(st))
    if vmax:
        # the reference keeps 4 * VMAX floats of list on the stack of whichever thread walks a leaf (RCBForceTree.cxx:940),
        # 16 MB with VMAX raised to 2^20: the OpenMP workers get theirs from OMP_STACKSIZE, but the calling thread is the
        # OpenMP master and the process's main stack is 8 MB (a segmentation fault in the middle of the walk, round 1's
        # "--state clumpy" crash).  The call therefore runs on a thread with a stack of its own.
        import threading
        box = {}
        old = threading.stack_size(512 << 20)
        try:
            th = threading.Thread(target=lambda: box.setdefault("rc", call()))
            th.start()
        finally:
            threading.stack_size(old)
        th.join()
        rc = box.get("rc", -1)
    else:
        rc = call()
    assert rc == 0, rc
    stats = {f: getattr(st, f) for f, _ in RefStats._fields_}
    tree = None
    if keep_tree:
        m = lib.ref_tree_size()
        tree = {k: np.empty(m, dtype=np.int64) for k in ("count", "offset", "cl", "cr")}
        box = np.empty((m, 10), dtype=np.float32)
        ip = C.POINTER(C.c_int64)
        rc = lib.ref_tree_get(m, tree["count"].ctypes.data_as(ip), tree["offset"].ctypes.data_as(ip),
                              tree["cl"].ctypes.data_as(ip), tree["cr"].ctypes.data_as(ip), _fp(box))
        assert rc == 0
        tree["xmin"], tree["xmax"], tree["xc"], tree["ppm"] = box[:, 0:3], box[:, 3:6], box[:, 6:9], box[:, 9]
        if tdpts == 12:
            pp = np.empty((m, 13), dtype=np.float32)
            assert lib.ref_tree_get_pp12(m, _fp(pp)) == 0
            tree["tdr"], tree["ppm12"] = pp[:, 0], pp[:, 1:]
    lib.ref_set_tdpts(1)
    return q, stats, tree


# ---- the reference's cloud-in-cell loops (oracle/_ref/libhaccref_cic.so: src/cpu/Particles.cxx array_index, cic, inverse_cic) ----
_cic = None


def cic_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libhaccref_cic.so"))


def _cic_lib():
    global _cic
    if _cic is None:
        lib = C.CDLL(os.path.join(_HERE, "_ref", "libhaccref_cic.so"))
        fp, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        lib.ref_cic.argtypes = [C.c_int64, fp, fp, fp, i32p, C.c_float, fp]
        lib.ref_inverse_cic.argtypes = [C.c_int64] + [fp] * 7 + [i32p, fp, C.c_float, C.c_float, C.c_int]
        _cic = lib
    return _cic


def cic(p, ng, gpscal):
    """Particles::cic of the reference on particle dict p: returns the (ng0, ng1, ng2) float32 grid; the deposit weight is
    c = gpscal^3 formed in float as the reference does (Particles.cxx:605)."""
    x, y, z = (np.ascontiguousarray(p[k], dtype=np.float32).copy() for k in ("x", "y", "z"))
    ng3 = np.asarray(ng, dtype=np.int32)
    rho = np.zeros(int(np.prod(ng3)) + 1, dtype=np.float32)
    assert _cic_lib().ref_cic(x.size, _fp(x), _fp(y), _fp(z), ng3.ctypes.data_as(C.POINTER(C.c_int32)), float(gpscal), _fp(rho)) == 0
    return rho[:-1].reshape(tuple(int(t) for t in ng3))


def inverse_cic(p, grid, tau, fscal, comp):
    """Particles::inverse_cic(tau, fscal, comp) of the reference: returns the updated array (vx, vy, vz or phi for comp 0..3)."""
    a = {k: np.ascontiguousarray(p[k], dtype=np.float32).copy() for k in ("x", "y", "z", "vx", "vy", "vz", "phi")}
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    ng3 = np.asarray(grid.shape, dtype=np.int32)
    g = np.concatenate([grid.ravel(), np.zeros(1, np.float32)])
    assert _cic_lib().ref_inverse_cic(a["x"].size, *[_fp(a[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "phi")],
                                      ng3.ctypes.data_as(C.POINTER(C.c_int32)), _fp(g), float(tau), float(fscal), int(comp)) == 0
    return a[("vx", "vy", "vz", "phi")[comp]]
