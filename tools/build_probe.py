#!/usr/bin/env python
"""Build-phase timing probe: ms_build / ms_walk / ms_force of one kick on the bench snapshot (no parity checks).
    python tools/build_probe.py [uniform|clustered] [np_side] [ppn]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hacc_coral_b200 as H  # noqa: E402
from hacc_coral_b200 import synth  # noqa: E402

state = sys.argv[1] if len(sys.argv) > 1 else "uniform"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ppn = int(sys.argv[3]) if len(sys.argv) > 3 else 512
z, boost = (50.0, 1.0) if state == "uniform" else (0.0, 0.35)
p = synth.zeldovich_torch(side, z=z, seed=5009888, ghost=11, growth_boost=boost, device="cuda")
n = p["x"].size
nglt = side + 22
g = H.HaccSR(n)
g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
g.upload(p)
b = ([0.0] * 3, [float(nglt)] * 3, [3.2] * 3, [nglt - 3.2] * 3)
for _ in range(2):
    g.kick(*b, 0.5, ppn, skip_force=True)
ms = [g.kick(*b, 0.5, ppn, skip_force=True) for _ in range(5)]
print("%s side=%d ppn=%d n=%d dbg=%s: build %.3f ms (min %.3f), walk %.3f, levels %d, nodes %d" % (
    state, side, ppn, n, os.environ.get("HACCSR_BUILD_DEBUG", "0"), np.mean([m["ms_build"] for m in ms]),
    min(m["ms_build"] for m in ms), np.mean([m["ms_walk"] for m in ms]), ms[-1]["levels"], ms[-1]["nodes"]))
g.close()
