"""Opt-in kernels that have been measured stand-alone but not yet run through the parity suite on hardware.  Runs LAST (file
name) and is a non-strict xfail, so whatever happens here cannot colour the gates in the other files."""
import numpy as np
import pytest

import hacc_coral_b200 as H
from hacc_coral_b200 import synth
from tests.util import RSM, boxes

pytestmark = pytest.mark.gpu


def _kick(p, b, ppn, tdpts):
    g = H.HaccSR(p["x"].size)
    try:
        g.set_force_law(H.LAW_SR_POLY, H.POLY5, RSM, H.RMAX)
        g.upload(p)
        st = g.kick(*b, 0.5, ppn, count_in_cutoff=True, tdpts=tdpts)
        return g.download(), st, g.tree()
    finally:
        g.close()


@pytest.mark.timeout(180)
@pytest.mark.xfail(strict=False, reason="k_cm_warp (HACCSR_CM_KERNEL=warp): validated stand-alone only (tools/microbench_cm.cu)")
@pytest.mark.parametrize("kind,n,ppn", [("clustered", 32, 100), ("zeld", 48, 512), ("lattice", 20, 64)])
def test_cm_warp_kernel_builds_the_identical_tree(kind, n, ppn, monkeypatch):
    """The persistent-warp centroid pass must leave every bit where k_cm_tile leaves it: node table, permutation, kicks."""
    if kind == "clustered":
        p = synth.clustered(60000, float(n), seed=5)
    elif kind == "zeld":
        p = synth.zeldovich(n, z=50.0, seed=6, ghost=0)
    else:
        p = synth.jitter_lattice(n, seed=7)
    b = boxes(n)
    monkeypatch.delenv("HACCSR_CM_KERNEL", raising=False)
    out0, st0, tr0 = _kick(p, b, ppn, 1)
    monkeypatch.setenv("HACCSR_CM_KERNEL", "warp")
    out1, st1, tr1 = _kick(p, b, ppn, 1)
    assert st0["nodes"] == st1["nodes"] and st0["pairs_evaluated"] == st1["pairs_evaluated"]
    assert st0["pairs_in_cutoff"] == st1["pairs_in_cutoff"]
    for k in ("count", "offset", "cl", "cr", "xmin", "xmax", "xc", "ppm"):
        assert np.array_equal(tr0[k], tr1[k]), k
    for k in out0:
        assert np.array_equal(out0[k], out1[k]), k
