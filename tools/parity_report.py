#!/usr/bin/env python
"""Achieved parity of the GPU kick against the COMPILED reference, as numbers (profiles/parity_r2.json):

    python tools/parity_report.py [--out profiles/parity_r2.json] [--cases c1,c2,c3]

Cases (BASELINE.json configs): c1 = np = ng = 128 + overload shell in full (every leaf kicked); c2 / c3 = the benchmark's own
256^3 + shell snapshots (z = 50 near-uniform / shell-crossed clustered) with the FULL 21.5 M-particle tree and the force box
shrunk to a central sub-cube, so that the reference's constructor kicks a few ten thousand particles in seconds
(reference src/halo_finder/RCBForceTree.cxx:1166-1172: only leaves touching the force box are walked).
For every arithmetic mode of the pair kernel (include/haccsr.h) and with warp-level culling on, per case:
  tree        node / leaf census and the evaluated-pair count equal to the reference's (exact)
  in_cutoff   pairs inside the cutoff, GPU vs reference (exact for x86; the fused modes round r2 differently)
  literal     fraction of kicked particles with |da| <= 1e-5 |a_ref| (north_star's literal gate)
  rel         p50 / p99 / p99.9 / max of |da| / |a_ref|
  to_fp64     distance to the FP64 sum of the same pair set: GPU and CPU reference, p50 / p99.9, and their ratio
Uses oracle/ (the compiled reference and the plain-C restatement in its FP64 form): test infrastructure, not the product."""
import argparse
import json
import os
import sys
import time

# the reference keeps its interaction lists on the stacks of its OpenMP workers (4 * VMAX floats, RCBForceTree.cxx:940; 16 MB with
# VMAX raised for clustered snapshots); libgomp reads the variable once, when the first OpenMP runtime of the process starts
os.environ.setdefault("OMP_STACKSIZE", "64M")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RSM, THETA, EDGE = 0.007, 0.5, 3.2


def _by_id(p, keys=("vx", "vy", "vz")):
    o = np.argsort(p["id"], kind="stable")
    return np.stack([np.asarray(p[k])[o] for k in keys], axis=1).astype(np.float64)


def _q(v, qs=(0.5, 0.99, 0.999)):
    return [float(np.quantile(v, q)) for q in qs] + [float(v.max())]


def parity_case(p, boxes, ppn, modes, vmax=False, want_tree=False, tree_modes=None):
    """p: particle dict; boxes = (tree_lo, tree_hi, force_lo, force_hi).  Returns the report dict of one case."""
    import hacc_coral_b200 as H
    from oracle import oraclebind as O, refbind as R
    from tests.util import compare_trees
    t0 = time.time()
    ref, rst, rtree = R.rcb_kick(p, *boxes, RSM, THETA, ppn, fcoeff=1.0, law=R.LAW_POLY5, count_pairs=True, keep_tree=want_tree, vmax=vmax)
    t_ref = time.time() - t0
    o64 = O.run(p, *boxes, RSM, THETA, ppn, form=O.FORM_FP64)
    b, c = _by_id(ref), _by_id(o64)
    nb, nc = np.sqrt((b * b).sum(axis=1)), np.sqrt((c * c).sum(axis=1))
    kicked = nb > 0
    r_cpu = np.sqrt(((b - c) ** 2).sum(axis=1))[kicked] / nc[kicked]
    out = {"particles": int(p["x"].size), "kicked": int(kicked.sum()), "ppn": ppn, "reference_seconds": t_ref,
           "reference": {"nodes": int(rst["nodes"]), "leaves": int(rst["leaves"]), "pairs_evaluated": int(rst["pairs_eval"]),
                         "pairs_in_cutoff": int(rst["pairs_incut"])},
           "cpu_to_fp64": dict(zip(("p50", "p99", "p999", "max"), _q(r_cpu))), "modes": {}}
    for name, arith, cull in modes:
        g = H.HaccSR(p["x"].size, arith=arith)
        g.set_force_law(H.LAW_SR_POLY, H.POLY5, RSM, H.RMAX)
        g.set_culling(cull)
        g.upload(p)
        st = g.kick(*boxes, THETA, ppn, count_in_cutoff=True)
        got = g.download()
        tree = g.tree() if (want_tree and (tree_modes is None or name in tree_modes)) else None
        g.close()
        a = _by_id(got)
        d = np.sqrt(((a - b) ** 2).sum(axis=1))
        assert np.all(d[~kicked] == 0), "a particle the reference did not kick was kicked"
        rel = d[kicked] / nb[kicked]
        r_gpu = np.sqrt(((a - c) ** 2).sum(axis=1))[kicked] / nc[kicked]
        m = {"tree_census_equal": bool(st["nodes"] == rst["nodes"] and st["leaves"] == rst["leaves"] and st["max_ppn"] == rst["max_ppn"]),
             "pairs_evaluated_equal": bool(st["pairs_evaluated"] == rst["pairs_eval"]),
             "pairs_in_cutoff": int(st["pairs_in_cutoff"]), "pairs_in_cutoff_minus_reference": int(st["pairs_in_cutoff"]) - int(rst["pairs_incut"]),
             "literal_1e-5_fraction": float((rel <= 1e-5).mean()),
             "rel": dict(zip(("p50", "p99", "p999", "max"), _q(rel))),
             "gpu_to_fp64": dict(zip(("p50", "p99", "p999", "max"), _q(r_gpu))),
             "to_fp64_ratio_gpu_over_cpu": {"p50": float(np.median(r_gpu) / np.median(r_cpu)),
                                            "p999": float(np.quantile(r_gpu, 0.999) / np.quantile(r_cpu, 0.999))}}
        if want_tree and (tree_modes is None or name in tree_modes):
            ids_gpu = got["id"]
            cmp = compare_trees(rtree, ref["id"], tree, ids_gpu)
            m["tree_compare"] = {k: int(v) for k, v in cmp.items()}
        out["modes"][name] = m
    return out


def default_modes():
    import hacc_coral_b200 as H
    return [("fused", H.ARITH_FUSED, False), ("fused_rs3", H.ARITH_FUSED_RS3, False), ("x86", H.ARITH_X86, False),
            ("fused+cull", H.ARITH_FUSED, True), ("fused_rs3+cull", H.ARITH_FUSED_RS3, True)]


def make_case(name, sub=32):
    """(particles, boxes, vmax) of a named case."""
    from hacc_coral_b200 import synth
    if name == "c1":
        p = synth.zeldovich_torch(128, z=50.0, seed=5009888, ghost=11, device="cuda")
        side = 150.0
        return p, ([0.0] * 3, [side] * 3, [EDGE] * 3, [side - EDGE] * 3), False
    z, boost = (50.0, 1.0) if name == "c2" else (0.0, 0.35)
    p = synth.zeldovich_torch(256, z=z, seed=5009888, ghost=11, growth_boost=boost, device="cuda")
    side = 278.0
    c = side / 2
    return p, ([0.0] * 3, [side] * 3, [c - sub / 2] * 3, [c + sub / 2] * 3), name == "c3"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r2.json"))
    ap.add_argument("--cases", default="c1,c2,c3")
    ap.add_argument("--sub", type=int, default=32, help="side of the force sub-cube of c2 / c3 in cells")
    args = ap.parse_args()
    rep = {"what": "GPU kick vs the compiled reference (oracle/_ref, x86-64 build) and vs the FP64 sum of the same pairs; "
                   "poly5, theta 0.5, ppn 512; c2/c3: full 21.5 M-particle tree, force box = central %d^3-cell sub-cube" % args.sub,
           "cases": {}}
    for name in args.cases.split(","):
        p, boxes, vmax = make_case(name, args.sub)
        rep["cases"][name] = parity_case(p, boxes, 512, default_modes(), vmax=vmax, want_tree=(name != "c1"))
        print(name, json.dumps(rep["cases"][name])[:600], file=sys.stderr)
    with open(args.out, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
