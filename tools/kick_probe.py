#!/usr/bin/env python
"""Three full kicks (build + walk + force) on the bench snapshot, for profiler captures of the second one.
    python tools/kick_probe.py [uniform|clustered] [np_side] [ppn] [fused|x86|fused_rs3] [cull]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hacc_coral_b200 as H  # noqa: E402
from hacc_coral_b200 import synth  # noqa: E402

state = sys.argv[1] if len(sys.argv) > 1 else "uniform"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ppn = int(sys.argv[3]) if len(sys.argv) > 3 else 512
arith = {"fused": H.ARITH_FUSED, "x86": H.ARITH_X86, "fused_rs3": H.ARITH_FUSED_RS3}[sys.argv[4] if len(sys.argv) > 4 else "fused"]
z, boost = (50.0, 1.0) if state == "uniform" else (0.0, 0.35)
p = synth.zeldovich_torch(side, z=z, seed=5009888, ghost=11, growth_boost=boost, device="cuda")
nglt = side + 22
g = H.HaccSR(p["x"].size, arith=arith)
g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
g.set_culling(len(sys.argv) > 5 and sys.argv[5] == "cull")
g.upload(p)
b = ([0.0] * 3, [float(nglt)] * 3, [3.2] * 3, [nglt - 3.2] * 3)
for _ in range(3):          # a profiler capture takes the second kick: warm, and not the last work of the process
    st = g.kick(*b, 0.5, ppn)
print("%s side=%d ppn=%d: build %.3f walk %.3f force %.3f ms, %d levels, %d launches" % (
    state, side, ppn, st["ms_build"], st["ms_walk"], st["ms_force"], st["levels"], st["total_launches"]))
g.close()
