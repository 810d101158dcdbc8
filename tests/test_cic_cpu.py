"""CPU tests of the CIC restatement (oracle/haccsr_oracle.c: orc_cic, orc_inverse_cic; reference src/cpu/Particles.cxx:589-714).
PINNED: bit-for-bit equal to the reference's own loops -- Particles::array_index / cic / inverse_cic cut out of Particles.cxx and
compiled by oracle/build_ref.sh into oracle/_ref/libhaccref_cic.so (the file as a whole needs MPI, the three functions do not) --
through the committed fixture tests/golden/cic_ref_clustered12k.npz (tests/golden/make_golden_cic.py) and, where the compiled
reference is present, live.  Plus properties any correct cloud-in-cell pair has and an independent float64 evaluation."""
import numpy as np
import pytest

from hacc_coral_b200 import synth


def _np_weights(p, ng):
    x, y, z = (p[k].astype(np.float64) for k in ("x", "y", "z"))
    i = [np.floor(v).astype(np.int64) for v in (x, y, z)]
    w0 = [1.0 + (ii - v) for ii, v in zip(i, (x, y, z))]
    return i, w0


def _np_cic(p, ng, c):
    i, w0 = _np_weights(p, ng)
    rho = np.zeros(tuple(ng), dtype=np.float64)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (w0[0] if dx == 0 else 1 - w0[0]) * (w0[1] if dy == 0 else 1 - w0[1]) * (w0[2] if dz == 0 else 1 - w0[2])
                ix, iy, iz = i[0] + dx, i[1] + dy, i[2] + dz
                ok = (ix >= 0) & (ix < ng[0]) & (iy >= 0) & (iy < ng[1]) & (iz >= 0) & (iz < ng[2])
                np.add.at(rho, (ix[ok], iy[ok], iz[ok]), c * w[ok])
    return rho


def _np_interp(p, grid):
    ng = grid.shape
    i, w0 = _np_weights(p, ng)
    f = np.zeros(p["x"].size)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (w0[0] if dx == 0 else 1 - w0[0]) * (w0[1] if dy == 0 else 1 - w0[1]) * (w0[2] if dz == 0 else 1 - w0[2])
                ix, iy, iz = i[0] + dx, i[1] + dy, i[2] + dz
                ok = (ix >= 0) & (ix < ng[0]) & (iy >= 0) & (iy < ng[1]) & (iz >= 0) & (iz < ng[2])
                f[ok] += grid[ix[ok], iy[ok], iz[ok]].astype(np.float64) * w[ok]
    return f


@pytest.fixture(scope="module")
def snap():
    p = synth.clustered(30000, 20.0, seed=3)
    rng = np.random.default_rng(4)
    # a few particles outside the grid (they only touch the overflow slot) and on cell faces
    for k in ("x", "y", "z"):
        p[k][:50] = (rng.random(50) * 30 - 5).astype(np.float32)
        p[k][50:80] = np.round(p[k][50:80])
        p[k + "v"[:0]] = p[k]
    for k in ("vx", "vy", "vz", "phi"):
        p[k] = rng.standard_normal(p["x"].size).astype(np.float32)
    return p


def test_cic_matches_float64_evaluation_and_conserves_mass(oracle, snap):
    ng, c = (20, 22, 21), 0.75
    rho = oracle.cic(snap, ng, c)
    ref = _np_cic(snap, ng, c)
    assert rho.shape == ng
    assert np.abs(rho - ref).max() <= 2e-6 * ref.max()
    # mass conservation: particles whose 8 cells are all inside deposit exactly c each (weights sum to 1)
    inside = np.ones(snap["x"].size, bool)
    for k, n in zip(("x", "y", "z"), ng):
        inside &= (snap[k] >= 0) & (snap[k] < n - 1)
    only = {k: v[inside] for k, v in snap.items()}
    assert abs(float(oracle.cic(only, ng, c).astype(np.float64).sum()) - c * inside.sum()) <= 1e-5 * c * inside.sum()


def test_inverse_cic_interpolates_linear_fields_exactly(oracle, snap):
    ng = (24, 24, 24)
    gx, gy, gz = np.meshgrid(np.arange(ng[0]), np.arange(ng[1]), np.arange(ng[2]), indexing="ij")
    grid = (0.5 + 0.25 * gx - 0.125 * gy + 2.0 * gz).astype(np.float32)        # exactly representable
    inside = np.ones(snap["x"].size, bool)
    for k, n in zip(("x", "y", "z"), ng):
        inside &= (snap[k] >= 0) & (snap[k] < n - 1)
    p = {k: v[inside] for k, v in snap.items()}
    p["vx"] = np.zeros(p["x"].size, np.float32)
    v = oracle.inverse_cic(p, grid, tau=1.0, fscal=1.0, comp=0)
    want = 0.5 + 0.25 * p["x"].astype(np.float64) - 0.125 * p["y"] + 2.0 * p["z"]
    assert np.abs(v - want).max() <= 2e-5


def test_inverse_cic_matches_float64_and_is_adjoint_of_cic(oracle, snap):
    ng = (20, 22, 21)
    rng = np.random.default_rng(7)
    grid = rng.standard_normal(ng).astype(np.float32)
    for comp, key in enumerate(("vx", "vy", "vz", "phi")):
        v = oracle.inverse_cic(snap, grid, tau=0.3, fscal=1.7, comp=comp)
        want = snap[key].astype(np.float64) + _np_interp(snap, grid) * 1.7 * 0.3
        assert np.abs(v - want).max() <= 3e-6 * max(1.0, np.abs(want).max()), key
    # <cic(particles), g> = c * sum_p interp(g)(x_p)
    c = 1.25
    lhs = float((oracle.cic(snap, ng, c).astype(np.float64) * grid).sum())
    rhs = c * float(_np_interp(snap, grid).sum())
    assert abs(lhs - rhs) <= 1e-4 * abs(rhs) + 1e-3


GOLD = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "cic_ref_clustered12k.npz")


def load_cic_golden():
    d = np.load(GOLD)
    p = {k: d[k] for k in ("x", "y", "z", "vx", "vy", "vz", "phi")}
    p["mass"] = np.ones(p["x"].size, np.float32)
    p["id"] = np.arange(p["x"].size, dtype=np.int64)
    p["mask"] = np.zeros(p["x"].size, np.uint16)
    return d, p, tuple(int(t) for t in d["ng"])


def test_oracle_cic_is_bit_equal_to_the_compiled_reference_fixture(oracle):
    d, p, ng = load_cic_golden()
    assert np.array_equal(oracle.cic(p, ng, float(d["c"])), d["rho"])
    for comp, key in enumerate(("vx", "vy", "vz", "phi")):
        assert np.array_equal(oracle.inverse_cic(p, d["grid"], float(d["tau"]), float(d["fscal"]), comp), d["out_" + key]), key


def test_oracle_cic_is_bit_equal_to_the_compiled_reference_live(oracle, snap):
    from oracle import refbind
    if not refbind.cic_available():
        pytest.skip("oracle/_ref/libhaccref_cic.so not built (no /root/reference on this box)")
    ng = (20, 22, 21)
    gpscal = np.float32(1.07)
    c = np.float32(np.float32(gpscal * gpscal) * gpscal)
    assert np.array_equal(oracle.cic(snap, ng, float(c)), refbind.cic(snap, ng, gpscal))
    grid = np.random.default_rng(17).standard_normal(ng).astype(np.float32)
    for comp in range(4):
        assert np.array_equal(oracle.inverse_cic(snap, grid, 0.21, 2.3, comp), refbind.inverse_cic(snap, grid, 0.21, 2.3, comp)), comp
