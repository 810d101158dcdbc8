"""Generate tests/golden/transfer_cmbM000.npz from the reference's shipped transfer function.

Run in the build container (needs /root/reference):  python tests/golden/make_transfer_table.py
Combination rule: T = T_b*Omega_b/Omega_m + T_cdm*(1 - Omega_b/Omega_m), normalised to its first row
(reference src/initializer/Cosmology.cpp:66-86; columns of cmbM000.tf: k, T_cdm, T_bar, ...).
Only the derived (k, T) table is stored -- 2 float64 columns -- not the reference file.
"""
import os
import sys

import numpy as np

REF = os.environ.get("HACC_REFERENCE", "/root/reference")
src = os.path.join(REF, "cmbM000.tf")
if not os.path.exists(src):
    sys.exit("reference transfer function not found at %s" % src)
t = np.loadtxt(src)
h, omega_dm, omega_b = 0.7, 0.23387755, 0.0226 / 0.7 ** 2      # reference indat:23-27
fb = omega_b / (omega_dm + omega_b)
T = t[:, 2] * fb + t[:, 1] * (1.0 - fb)
T = T / T[0]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transfer_cmbM000.npz")
np.savez(out, k=t[:, 0], T=T)
print("wrote", out, t.shape, "k range", t[0, 0], t[-1, 0])
