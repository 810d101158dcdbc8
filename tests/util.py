"""Shared helpers for the parity tests (oracle vs libhaccsr)."""
import numpy as np

RSM = 0.007           # reference indat:38
EDGE = 3.2            # reference indat:37
THETA = 0.5           # reference indat:41 (top-level); examples/*/indat use 0.1


def boxes(n, edge=EDGE):
    """Tree box [0,n]^3 and force box [edge, n-edge]^3 (reference src/cpu/Particles.cxx:1213-1228)."""
    return [0.0] * 3, [float(n)] * 3, [edge] * 3, [float(n) - edge] * 3


def by_id(p, keys=("vx", "vy", "vz")):
    o = np.argsort(p["id"], kind="stable")
    return {k: np.asarray(p[k])[o] for k in keys}


def tree_key_map(tree):
    """Map (offset, count) -> node index for non-empty nodes (independent of node numbering)."""
    return {(int(o), int(c)): i for i, (o, c) in enumerate(zip(tree["offset"], tree["count"])) if c > 0}


def compare_trees(ta, ida, tb, idb):
    """Compare two trees built over the same particles but with different node numbering and different
    (stable vs swap-sequence) order inside nodes.  Returns a dict of mismatch counts."""
    ka, kb = tree_key_map(ta), tree_key_map(tb)
    res = {"nodes_a": len(ka), "nodes_b": len(kb), "missing": 0, "box_mismatch": 0, "xc_mismatch": 0,
           "ppm_mismatch": 0, "leaf_flag_mismatch": 0, "leaf_members_mismatch": 0}
    for key, ia in ka.items():
        ib = kb.get(key)
        if ib is None:
            res["missing"] += 1
            continue
        if not (np.array_equal(ta["xmin"][ia], tb["xmin"][ib]) and np.array_equal(ta["xmax"][ia], tb["xmax"][ib])):
            res["box_mismatch"] += 1
        if not np.array_equal(ta["xc"][ia], tb["xc"][ib]):
            res["xc_mismatch"] += 1
        la = ta["cl"][ia] == 0 and ta["cr"][ia] == 0
        lb = tb["cl"][ib] == 0 and tb["cr"][ib] == 0
        if la != lb:
            res["leaf_flag_mismatch"] += 1
        if key[1] > 1 and ta["ppm"][ia] != tb["ppm"][ib]:
            res["ppm_mismatch"] += 1
        if la and lb:
            o, c = key
            if not np.array_equal(np.sort(ida[o:o + c]), np.sort(idb[o:o + c])):
                res["leaf_members_mismatch"] += 1
    return res


def accel_errors(a, b):
    """Relative error per particle |a-b|/|b| (b = oracle) with |b| floored at 1e-3 of the rms."""
    a = np.stack([a["vx"], a["vy"], a["vz"]], axis=1).astype(np.float64)
    b = np.stack([b["vx"], b["vy"], b["vz"]], axis=1).astype(np.float64)
    nb = np.sqrt((b * b).sum(axis=1))
    rms = np.sqrt((nb * nb).mean()) if nb.size else 0.0
    d = np.sqrt(((a - b) ** 2).sum(axis=1))
    rel = d / np.maximum(nb, 1e-3 * rms + 1e-30)
    return rel, d, nb, rms
