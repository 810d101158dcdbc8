"""Where the time of haccsr_kick_host goes (development probe): phase times reported by the library next to the
wall time of the call, on the bench workload."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hacc_coral_b200 as H
from hacc_coral_b200 import synth
side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = synth.zeldovich_torch(side, z=50.0, seed=5009888, ghost=11, device="cuda")
n = p["x"].size; nglt = side + 22
pin = {k: torch.from_numpy(v.copy()).pin_memory().numpy() for k, v in p.items()}
g = H.HaccSR(n); g.set_force_law(H.LAW_SR_POLY, H.POLY5, 0.007, H.RMAX)
lo, hi, flo, fhi = [0.0]*3, [float(nglt)]*3, [3.2]*3, [nglt-3.2]*3
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = g.kick_host(pin, lo, hi, flo, fhi, 0.5, 512)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("kick_host wall %.1f ms | build %.1f walk %.1f force %.1f total(dev, after xyzm upload) %.1f | launches %d" % (
        1e3*(t1-t0), st["ms_build"], st["ms_walk"], st["ms_force"], st["ms_total"], st["force_launches"]))
g.upload(pin)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = g.kick(lo, hi, flo, fhi, 0.5, 512)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("kick (resident) wall %.1f ms | build %.1f walk %.1f force %.1f total %.1f" % (1e3*(t1-t0), st["ms_build"], st["ms_walk"], st["ms_force"], st["ms_total"]))
t0 = time.perf_counter(); g.upload(pin); t1 = time.perf_counter(); out = g.download(); t2 = time.perf_counter()
print("upload %.1f ms, download (pageable numpy) %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1)))
