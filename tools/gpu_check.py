"""Quick GPU-vs-oracle check (development tool; the real gates are tests/ -m gpu)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hacc_coral_b200 as H
from hacc_coral_b200 import synth
from oracle import oraclebind as O
from tests.util import RSM, boxes, by_id, compare_trees, accel_errors


def run(name, p, n, ppn, theta, law_coef=H.POLY5):
    lo, hi, flo, fhi = boxes(n)
    t0 = time.time()
    o = O.run(p, lo, hi, flo, fhi, RSM, theta, ppn, coef=law_coef)
    t1 = time.time()
    g = H.HaccSR(p["x"].size)
    g.set_force_law(H.LAW_SR_POLY, law_coef, RSM, H.RMAX)
    g.upload(p)
    st = g.kick(lo, hi, flo, fhi, theta, ppn, count_in_cutoff=True)
    out = g.download()
    tr = g.tree()
    t2 = time.time()
    cmp_ = compare_trees(tr, out["id"], o["tree"], o["id"])
    rel, d, nb, rms = accel_errors(by_id(out), by_id(o))
    if p["x"].size <= 400000:
        o64 = O.run(p, lo, hi, flo, fhi, RSM, theta, ppn, coef=law_coef, form=O.FORM_FP64)
        r1, _, _, _ = accel_errors(by_id(out), by_id(o64))
        r2, _, _, _ = accel_errors(by_id(o), by_id(o64))
        print("   vs FP64 same-pair-set sum: gpu max %.3e p99.9 %.3e median %.3e | cpu-ref max %.3e p99.9 %.3e median %.3e" % (
            r1.max(), np.quantile(r1, 0.999), np.median(r1), r2.max(), np.quantile(r2, 0.999), np.median(r2)))
    os_ = o["stats"]
    print("== %s n=%d N=%d ppn=%d theta=%.2f  oracle %.1fs gpu-call %.2fs" % (name, n, p["x"].size, ppn, theta, t1 - t0, t2 - t1))
    print("   tree:", cmp_)
    print("   nodes gpu/orc %d/%d leaves %d/%d sink %d/%d maxlist %d/%d" % (st["nodes"], os_["nodes"], st["leaves"], os_["leaves"], st["sink_leaves"], os_["sink_leaves"], st["max_list"], os_["max_list"]))
    print("   pairs eval gpu/orc %d/%d  incut %d/%d" % (st["pairs_evaluated"], os_["pairs_eval"], st["pairs_in_cutoff"], os_["pairs_incut"]))
    print("   accel rel err: max %.3e  p99.9 %.3e  median %.3e   rms|a| %.4f  frac>1e-5: %.2e" % (rel.max(), np.quantile(rel, 0.999), np.median(rel), rms, (rel > 1e-5).mean()))
    print("   ms build %.3f walk %.3f force %.3f total %.3f  launches %d  Gpairs/s(force) %.1f" % (st["ms_build"], st["ms_walk"], st["ms_force"], st["ms_total"], st["total_launches"], st["pairs_evaluated"] / max(st["ms_force"], 1e-6) / 1e6))
    g.close()
    return rel.max()


if __name__ == "__main__":
    O.build()
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run("lattice", synth.jitter_lattice(16), 16, 64, 0.5)
        run("lattice", synth.jitter_lattice(24), 24, 512, 0.5)
        run("lattice", synth.jitter_lattice(32), 32, 100, 0.5)
        run("clustered", synth.clustered(40000, 32.0), 32, 128, 0.5)
        run("clustered", synth.clustered(40000, 32.0), 32, 16, 0.3)
    else:
        run("lattice", synth.jitter_lattice(48), 48, 512, 0.5)
        run("zeld", synth.zeldovich(64, ghost=0), 64, 512, 0.5)
        run("clustered", synth.clustered(300000, 64.0), 64, 512, 0.5)
