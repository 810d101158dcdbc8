"""Development probe: haccsr_subcycle against the same loop made of the individual C-ABI calls, per sub-step."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hacc_coral_b200 as H
from hacc_coral_b200 import synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 3
GHOST = 11
nglt = side + 2 * GHOST
p = synth.zeldovich_torch(side, z=50.0, seed=5009888, ghost=GHOST, device="cuda")
n = p["x"].size
vmax = max(float(np.abs(p[k]).max()) for k in ("vx", "vy", "vz"))
pt = 0.02 / vmax if vmax > 0 else 0.01
print("n", n, "vmax", vmax, "pt", pt, "pos range", [(float(p[k].min()), float(p[k].max())) for k in ("x", "y", "z")])
lo, hi = [0.0] * 3, [float(nglt)] * 3
flo, fhi = [3.2] * 3, [nglt - 3.2] * 3
g = H.HaccSR(n)
g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
g.upload(p)
for s in range(nsub):
    t0 = time.time()
    g.stream(pt)
    nin = g.partition_in_box(hi)
    g.fill_mass(1.0)
    st = g.kick(lo, hi, flo, fhi, 0.5, 512, fcoeff=1e-3, count=nin)
    g.stream(pt)
    print("manual step", s, "nin", nin, "pairs", st["pairs_evaluated"], "ms build/walk/force", st["ms_build"], st["ms_walk"], st["ms_force"],
          "wall", time.time() - t0)
g.upload(p)
t0 = time.time()
st = g.subcycle(nsub, pt, hi, lo, hi, flo, fhi, 0.5, 512, 1e-3)
print("subcycle", nsub, "pairs", st["pairs_evaluated"], "ms build/walk/force", st["ms_build"], st["ms_walk"], st["ms_force"], "particles", st["particles"],
      "wall", time.time() - t0)
g.close()
