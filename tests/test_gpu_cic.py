"""GPU tests (-m gpu) of the PM coupling kernels (csrc/cic.cu) through the C ABI against the reference's own loops
(Particles::cic / Particles::inverse_cic, reference src/cpu/Particles.cxx:589-714): the committed fixture generated from the
compiled reference (tests/golden/cic_ref_clustered12k.npz) and the oracle restatement that is bit-equal to it."""
import numpy as np
import pytest

import hacc_coral_b200 as H
from hacc_coral_b200 import synth

pytestmark = pytest.mark.gpu


def _snap(n=200000, side=40.0, seed=5):
    p = synth.clustered(n, side, seed=seed)
    rng = np.random.default_rng(seed + 1)
    for k in ("x", "y", "z"):
        p[k][:200] = (rng.random(200) * (side + 10) - 5).astype(np.float32)      # some outside the grid
        p[k][200:300] = np.round(p[k][200:300])                                   # some exactly on cell faces
    for k in ("vx", "vy", "vz", "phi"):
        p[k] = rng.standard_normal(n).astype(np.float32)
    return p


def test_inverse_cic_is_bit_identical_to_the_reference_loop(oracle):
    p = _snap()
    ng = (40, 41, 42)
    grid = np.random.default_rng(9).standard_normal(ng).astype(np.float32)
    g = H.HaccSR(p["x"].size)
    g.upload(p)
    for comp in range(4):
        g.inverse_cic(grid, tau=0.37, fscal=1.9, comp=comp)
    out = g.download()
    g.close()
    for comp, key in enumerate(("vx", "vy", "vz", "phi")):
        assert np.array_equal(out[key], oracle.inverse_cic(p, grid, 0.37, 1.9, comp)), key
    for key in ("x", "y", "z", "mass", "id"):
        assert np.array_equal(out[key], p[key])


def test_cic_matches_reference_loop_to_rounding_and_is_deterministic(oracle):
    p = _snap()
    ng, c = (40, 41, 42), 0.8
    g = H.HaccSR(p["x"].size)
    g.upload(p)
    rho1 = g.cic(ng, c)
    # the fixed-point sum does not depend on particle order: shuffle, deposit again, same bits
    perm = np.random.default_rng(1).permutation(p["x"].size)
    g.upload({k: v[perm] for k, v in p.items()})
    rho2 = g.cic(ng, c)
    g.close()
    assert np.array_equal(rho1, rho2)
    ref = oracle.cic(p, ng, c)
    # the reference rounds to float after every particle (order = particle order); the device rounds the exact sum once
    # (a clump cell collects ~1500 particles: the sequential float sum carries ~sqrt(n) * 6e-8 of the cell value)
    assert np.all(np.abs(rho1.astype(np.float64) - ref) <= 1e-5 * ref + 1e-6)
    ref64 = np.zeros(ng)                               # and the device is the closer of the two to the exact sum
    from tests.test_cic_cpu import _np_cic
    ref64 = _np_cic(p, ng, c)
    assert np.abs(rho1 - ref64).max() <= np.abs(ref - ref64).max() + 1e-6
    # mass: particles with all 8 cells inside deposit c each
    inside = np.ones(p["x"].size, bool)
    for k, n in zip(("x", "y", "z"), ng):
        inside &= (p[k] >= 0) & (p[k] < n - 1)
    g = H.HaccSR(int(inside.sum()))
    g.upload({k: v[inside] for k, v in p.items()})
    tot = float(g.cic(ng, c).astype(np.float64).sum())
    g.close()
    assert abs(tot - c * inside.sum()) <= 1e-6 * c * inside.sum()


def test_cic_empty_and_single_particle(oracle):
    ng = (4, 5, 6)
    g = H.HaccSR(8)
    g.upload(synth._pack(np.zeros(0), np.zeros(0), np.zeros(0)))
    assert not g.cic(ng, 1.0).any()
    one = synth._pack(np.array([1.25]), np.array([2.5]), np.array([3.75]))
    g.upload(one)
    rho = g.cic(ng, 2.0)
    g.close()
    assert np.array_equal(rho, oracle.cic(one, ng, 2.0))
    assert abs(float(rho.sum()) - 2.0) < 1e-6 and rho[1, 2, 3] == np.float32(2.0 * 0.75 * 0.5 * 0.25)


def test_against_the_compiled_reference_fixture():
    """haccsr_inverse_cic bit-identical, haccsr_cic within float rounding of the reference's sequential float sum (the device
    rounds the exact sum once: |difference| <= 1e-5 of the cell value + 1e-6)."""
    from tests.test_cic_cpu import load_cic_golden
    d, p, ng = load_cic_golden()
    g = H.HaccSR(p["x"].size)
    g.upload(p)
    rho = g.cic(ng, float(d["c"]))
    for comp in range(4):
        g.inverse_cic(d["grid"], tau=float(d["tau"]), fscal=float(d["fscal"]), comp=comp)
    out = g.download()
    g.close()
    for key in ("vx", "vy", "vz", "phi"):
        assert np.array_equal(out[key], d["out_" + key]), key
    ref = d["rho"].astype(np.float64)
    assert np.all(np.abs(rho.astype(np.float64) - ref) <= 1e-5 * ref + 1e-6)
    assert abs(float(rho.astype(np.float64).sum()) - float(ref.sum())) <= 1e-6 * float(ref.sum())
