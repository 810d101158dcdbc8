"""ctypes binding of oracle/liboracle.so (the plain-C restatement, haccsr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg.  The product (hacc_coral_b200/) never imports anything under oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LAW_POLY, LAW_NEWTON = 0, 3
FORM_GENERIC, FORM_BGQ_TAIL, FORM_FP64, FORM_GROSS, FORM_FUSED = 0, 1, 2, 3, 4
POLY5 = np.array([0.269327, -0.0750978, 0.0114808, -0.00109313, 0.0000605491, -0.00000147177],
                 dtype=np.float32)   # reference ForceLaw.cxx:109-114 == BGQStep16.c:167
POLY6 = np.array([0.271431, -0.0783394, 0.0133122, -0.00159485, 0.000132336, -0.00000663394,
                  0.000000147305], dtype=np.float32)   # reference ForceLaw.cxx:98-104
RMAX = np.float32(3.116326355)       # reference ForceLaw.cxx:32


class OrcStats(C.Structure):
    _fields_ = [("nodes", C.c_int64), ("leaves", C.c_int64), ("empty_leaves", C.c_int64),
                ("max_ppn", C.c_int64), ("sink_leaves", C.c_int64), ("max_list", C.c_int64),
                ("mean_ppn", C.c_double), ("pairs_eval", C.c_uint64), ("pairs_incut", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int64)
        lib.orc_run.restype = C.c_void_p
        lib.orc_run.argtypes = [C.c_int64, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int, fp, C.c_int,
                                C.c_float, C.c_float, C.c_float, C.c_int64, C.c_float, C.c_int, C.c_int,
                                C.c_int, C.POINTER(OrcStats)]
        for f in ("orc_nnode", "orc_nsink", "orc_nlist"):
            getattr(lib, f).restype = C.c_int64
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.orc_get_perm.argtypes = [C.c_void_p, ip]
        lib.orc_get_nodes.argtypes = [C.c_void_p, ip, ip, ip, ip, fp]
        lib.orc_get_lists.argtypes = [C.c_void_p, ip, ip, ip, C.POINTER(C.c_uint8)]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_set_tdpts.argtypes = [C.c_int]
        i32p = C.POINTER(C.c_int32)
        lib.orc_cic.argtypes = [C.c_int64, fp, fp, fp, i32p, C.c_float, fp]
        lib.orc_inverse_cic.argtypes = [C.c_int64, fp, fp, fp, fp, i32p, fp, C.c_float, C.c_float]
        lib.orc_get_pp12.argtypes = [C.c_void_p, fp]
        dp = C.POINTER(C.c_double)
        lib.orc_direct_sum.argtypes = [C.c_int64, fp, fp, fp, fp, C.c_int64, ip, fp, C.c_int, C.c_float,
                                       C.c_float, dp, dp, dp]
        lib.orc_force_law_eval.argtypes = [C.c_int, fp, C.c_int, C.c_float, C.c_float, C.c_int64, fp, fp]
        _LIB = lib
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def run(p, tree_lo, tree_hi, force_lo, force_hi, rsm, theta, ppn, fcoeff=1.0, rmax=RMAX, coef=POLY5,
        law=LAW_POLY, form=FORM_GENERIC, do_force=True, keep_lists=False, tdpts=1):
    """Build + walk + kick on particle dict `p` (keys x y z vx vy vz mass [phi id mask]).

    Returns dict with: particles permuted into tree order (all keys of p), 'perm', 'stats', 'tree'
    (count offset cl cr xmin xmax xc ppm) and, with keep_lists, 'lists' (sink_leaf, off, node, pseudo)."""
    lib = _lib()
    n = int(np.asarray(p["x"]).size)
    x, y, z, m = (np.ascontiguousarray(p[k], dtype=np.float32) for k in ("x", "y", "z", "mass"))
    vx, vy, vz = (np.ascontiguousarray(p[k], dtype=np.float32).copy() for k in ("vx", "vy", "vz"))
    boxes = np.array(list(tree_lo) + list(tree_hi) + list(force_lo) + list(force_hi), dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    st = OrcStats()
    lib.orc_set_tdpts(int(tdpts))       # RCBForceTree<TDPTS>: 1 = monopole (-R), 12 = quadrupole (-S)
    h = lib.orc_run(n, _fp(x), _fp(y), _fp(z), _fp(m), _fp(vx), _fp(vy), _fp(vz), _fp(boxes), law,
                    _fp(coef), len(coef), rsm, float(rmax), theta, ppn, fcoeff, form, int(do_force),
                    int(keep_lists), C.byref(st))
    try:
        perm = np.empty(n, dtype=np.int64)
        lib.orc_get_perm(h, _ip(perm))
        nn = lib.orc_nnode(h)
        tree = {k: np.empty(nn, dtype=np.int64) for k in ("count", "offset", "cl", "cr")}
        box = np.empty((nn, 10), dtype=np.float32)
        lib.orc_get_nodes(h, _ip(tree["count"]), _ip(tree["offset"]), _ip(tree["cl"]), _ip(tree["cr"]), _fp(box))
        tree["xmin"], tree["xmax"], tree["xc"], tree["ppm"] = box[:, 0:3], box[:, 3:6], box[:, 6:9], box[:, 9]
        if tdpts == 12:
            pp = np.empty((nn, 13), dtype=np.float32)
            lib.orc_get_pp12(h, _fp(pp))
            tree["tdr"], tree["ppm12"] = pp[:, 0], pp[:, 1:]
        out = {"perm": perm, "tree": tree, "stats": {f: getattr(st, f) for f, _ in OrcStats._fields_}}
        for k, v in p.items():
            if k in ("vx", "vy", "vz"):
                continue
            out[k] = np.asarray(v)[perm]
        out["vx"], out["vy"], out["vz"] = vx, vy, vz
        if keep_lists:
            ns, nl = lib.orc_nsink(h), lib.orc_nlist(h)
            sink = np.empty(ns, dtype=np.int64)
            off = np.empty(ns + 1, dtype=np.int64)
            node = np.empty(nl, dtype=np.int64)
            pseudo = np.empty(nl, dtype=np.uint8)
            lib.orc_get_lists(h, _ip(sink), _ip(off), _ip(node), pseudo.ctypes.data_as(C.POINTER(C.c_uint8)))
            out["lists"] = {"sink_leaf": sink, "off": off, "node": node, "pseudo": pseudo}
    finally:
        lib.orc_free(h)
        lib.orc_set_tdpts(1)
    return out


def direct_sum(p, sel, rsm, rmax=RMAX, coef=POLY5):
    """FP64 direct sum of the short-range force on particles `sel` from all particles within rmax."""
    lib = _lib()
    x, y, z, m = (np.ascontiguousarray(p[k], dtype=np.float32) for k in ("x", "y", "z", "mass"))
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    a = [np.empty(sel.size, dtype=np.float64) for _ in range(3)]
    dp = C.POINTER(C.c_double)
    lib.orc_direct_sum(x.size, _fp(x), _fp(y), _fp(z), _fp(m), sel.size, _ip(sel), _fp(coef), len(coef),
                       rsm, float(rmax), *(v.ctypes.data_as(dp) for v in a))
    return np.stack(a, axis=1)


def force_law_eval(r2, rsm, rmax=RMAX, coef=POLY5, law=LAW_POLY):
    r2 = np.ascontiguousarray(r2, dtype=np.float32)
    out = np.empty_like(r2)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    _lib().orc_force_law_eval(law, _fp(coef), len(coef), rsm, float(rmax), r2.size, _fp(r2), _fp(out))
    return out


def cic(p, ng, c):
    """Particles::cic restated (src/cpu/Particles.cxx:589-643): returns the (ng0, ng1, ng2) float32 grid."""
    x, y, z = (np.ascontiguousarray(p[k], dtype=np.float32) for k in ("x", "y", "z"))
    ng3 = np.asarray(ng, dtype=np.int32)
    rho = np.empty(int(np.prod(ng3.astype(np.int64))) + 1, dtype=np.float32)
    _lib().orc_cic(x.size, _fp(x), _fp(y), _fp(z), ng3.ctypes.data_as(C.POINTER(C.c_int32)), float(c), _fp(rho))
    return rho[:-1].reshape(tuple(int(t) for t in ng3))


def inverse_cic(p, grid, tau, fscal, comp):
    """Particles::inverse_cic restated (:647-714): returns the kicked copy of v[comp] (comp 3 = phi)."""
    x, y, z = (np.ascontiguousarray(p[k], dtype=np.float32) for k in ("x", "y", "z"))
    v = np.ascontiguousarray(p[("vx", "vy", "vz", "phi")[comp]], dtype=np.float32).copy()
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    ng3 = np.asarray(grid.shape, dtype=np.int32)
    g = np.concatenate([grid.ravel(), np.zeros(1, np.float32)])
    _lib().orc_inverse_cic(x.size, _fp(x), _fp(y), _fp(z), _fp(v), ng3.ctypes.data_as(C.POINTER(C.c_int32)), _fp(g),
                           float(tau), float(fscal))
    return v
