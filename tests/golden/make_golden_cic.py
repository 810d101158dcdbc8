"""Generate tests/golden/ref_cic_*.npz by running the COMPILED REFERENCE's own cloud-in-cell loops
(oracle/_ref/libhaccref_cic.so = Particles::array_index / cic / inverse_cic cut out of /root/reference/src/cpu/Particles.cxx
by oracle/build_ref.sh) on a small seeded particle set.

Run in the build container:  python tests/golden/make_golden_cic.py
The fixture stores the inputs (positions, the four arrays inverse_cic can update, the input grid, the scalars) and the
reference's outputs: the deposited density grid and the four updated arrays."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from hacc_coral_b200 import synth          # noqa: E402
from oracle import refbind as R            # noqa: E402

n, side, ng = 12000, 20.0, (20, 22, 21)
p = synth.clustered(n, side, seed=21, n_clumps=5)
rng = np.random.default_rng(22)
for k in ("x", "y", "z"):
    p[k][:60] = (rng.random(60) * (side + 10) - 5).astype(np.float32)      # some outside the grid: they only touch the overflow slot
    p[k][60:100] = np.round(p[k][60:100])                                  # some exactly on cell faces
for k in ("vx", "vy", "vz", "phi"):
    p[k] = rng.standard_normal(n).astype(np.float32)
gpscal = np.float32(0.93)
c = np.float32(np.float32(gpscal * gpscal) * gpscal)                         # Particles.cxx:605, float arithmetic
rho = R.cic(p, ng, gpscal)
grid = rng.standard_normal(ng).astype(np.float32)
tau, fscal = np.float32(0.37), np.float32(1.9)
out = {("out_" + k): R.inverse_cic(p, grid, tau, fscal, comp) for comp, k in enumerate(("vx", "vy", "vz", "phi"))}
path = os.path.join(HERE, "cic_ref_clustered12k.npz")
np.savez_compressed(path, x=p["x"], y=p["y"], z=p["z"], vx=p["vx"], vy=p["vy"], vz=p["vz"], phi=p["phi"], ng=np.asarray(ng, np.int32),
                    gpscal=gpscal, c=c, rho=rho, grid=grid, tau=tau, fscal=fscal, **out)
print(path, os.path.getsize(path), float(rho.sum()), float(rho.max()))
