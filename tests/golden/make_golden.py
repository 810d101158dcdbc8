"""Generate tests/golden/ref_*.npz by running the COMPILED REFERENCE (oracle/_ref/libhaccref.so, built
from /root/reference by oracle/build_ref.sh) on small seeded snapshots.

Run in the build container:  python tests/golden/make_golden.py
Each fixture stores the input particles, the call parameters, and the reference's outputs: kicked
velocities ordered by particle id, the reference's tree census and its evaluated / in-cutoff pair counts
(the latter from the counting ForceLaw wrapper in oracle/ref_harness.cxx).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from hacc_coral_b200 import synth          # noqa: E402
from oracle import refbind as R            # noqa: E402

CASES = {
    # name: (particles, box side, ppn, theta, law, edge)
    "lattice16_ppn64": (synth.jitter_lattice(16, seed=11), 16, 64, 0.5, R.LAW_POLY5, 3.2),
    "lattice8_rootleaf": (synth.jitter_lattice(8, seed=12), 8, 512, 0.5, R.LAW_POLY5, 0.0),
    "clustered6k_ppn32": (synth.clustered(6000, 24.0, seed=13, n_clumps=6), 24, 32, 0.5, R.LAW_POLY5, 3.2),
    "clustered6k_ppn32_theta01": (synth.clustered(6000, 24.0, seed=13, n_clumps=6), 24, 32, 0.1, R.LAW_POLY5, 3.2),
    "zeld24_ppn100_poly6": (synth.zeldovich(24, z=50.0, seed=14, ghost=0), 24, 100, 0.5, R.LAW_POLY6, 3.2),
    # the other two grid-force evaluators of the reference: analytic fit (run_hacc.sh default) and -i 1024
    "lattice16_ppn64_fit": (synth.jitter_lattice(16, seed=11), 16, 64, 0.5, R.LAW_FIT, 3.2),
    "clustered6k_ppn32_fit": (synth.clustered(6000, 24.0, seed=13, n_clumps=6), 24, 32, 0.5, R.LAW_FIT, 3.2),
    "lattice16_ppn64_interp1024": (synth.jitter_lattice(16, seed=11), 16, 64, 0.5, R.LAW_INTERP, 3.2),
    # RCBQuadrupoleForceTree (-S): 12 pseudo-particles per accepted node (7th tuple entry = TDPTS)
    "quad_clustered6k_ppn32": (synth.clustered(6000, 24.0, seed=13, n_clumps=6), 24, 32, 0.5, R.LAW_POLY5, 3.2, 12),
    "quad_clustered20k_ppn64_theta03": (synth.clustered(20000, 24.0, seed=15), 24, 64, 0.3, R.LAW_POLY5, 3.2, 12),
    "quad_lattice16_ppn64": (synth.jitter_lattice(16, seed=11), 16, 64, 0.5, R.LAW_POLY5, 3.2, 12),
}
NINTERP = 1024

ONLY = sys.argv[1:]      # optional: fixture names to (re)generate; default all

for name, case in CASES.items():
    if ONLY and name not in ONLY:
        continue
    p, n, ppn, theta, law, edge = case[:6]
    tdpts = case[6] if len(case) > 6 else 1
    lo, hi, flo, fhi = [0.0] * 3, [float(n)] * 3, [edge] * 3, [float(n) - edge] * 3
    extra = {}
    coef = R.POLY5
    if law == R.LAW_INTERP:
        coef = np.zeros(NINTERP, dtype=np.float32)      # the harness only reads its length (= nInterp)
        extra["table"] = R.fgrid_table(NINTERP)
    if law == R.LAW_FIT:
        extra["fit"] = R.fgrid_constants()
    q, st, tree = R.rcb_kick(p, lo, hi, flo, fhi, 0.007, theta, ppn, fcoeff=1.0, law=law, coef=coef, count_pairs=True,
                             keep_tree=True, tdpts=tdpts, vmax=(tdpts == 12))
    if tdpts == 12:   # pseudo-particle radius and masses of every node with more than 12 particles, keyed by (offset, count)
        big = tree["count"] > 12
        extra.update(tdpts=12, pp_offset=tree["offset"][big], pp_count=tree["count"][big], pp_tdr=tree["tdr"][big],
                     pp_ppm12=tree["ppm12"][big], pp_leaf=((tree["cl"] == 0) & (tree["cr"] == 0))[big])
    o = np.argsort(q["id"], kind="stable")
    out = os.path.join(HERE, "ref_%s.npz" % name)
    np.savez_compressed(out, x=p["x"], y=p["y"], z=p["z"], n=n, ppn=ppn, theta=np.float32(theta), law=law,
                        edge=np.float32(edge), rsm=np.float32(0.007),
                        vx=q["vx"][o], vy=q["vy"][o], vz=q["vz"][o],
                        nodes=st["nodes"], leaves=st["leaves"], empty_leaves=st["empty_leaves"],
                        max_ppn=st["max_ppn"], mean_ppn=st["mean_ppn"], pairs_eval=st["pairs_eval"],
                        pairs_incut=st["pairs_incut"], **extra,
                        leaf_offset=tree["offset"][(tree["cl"] == 0) & (tree["cr"] == 0) & (tree["count"] > 0)],
                        leaf_count=tree["count"][(tree["cl"] == 0) & (tree["cr"] == 0) & (tree["count"] > 0)])
    print(name, p["x"].size, {k: st[k] for k in ("nodes", "leaves", "pairs_eval", "pairs_incut")}, os.path.getsize(out))
