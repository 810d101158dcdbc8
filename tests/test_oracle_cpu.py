"""CPU tests: the plain-C restatement (oracle/haccsr_oracle.c) against (a) the golden fixtures generated
from the compiled reference and (b) the compiled reference itself when oracle/_ref is present."""
import glob
import os

import numpy as np
import pytest

from hacc_coral_b200 import synth
from tests.util import RSM, boxes, by_id

POLY = {0: "POLY5", 1: "POLY6"}


def _golden_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "ref_*.npz")))


def _load(path):
    d = np.load(path)
    n = d["x"].size
    p = {"x": d["x"], "y": d["y"], "z": d["z"], "vx": np.zeros(n, np.float32), "vy": np.zeros(n, np.float32),
         "vz": np.zeros(n, np.float32), "mass": np.ones(n, np.float32), "phi": np.zeros(n, np.float32),
         "id": np.arange(n, dtype=np.int64), "mask": np.zeros(n, np.uint16)}
    return d, p


def test_golden_fixtures_exist(golden_dir):
    assert len(_golden_files(golden_dir)) >= 5


@pytest.mark.parametrize("name", ["lattice16_ppn64", "lattice8_rootleaf", "clustered6k_ppn32",
                                  "clustered6k_ppn32_theta01", "zeld24_ppn100_poly6"])
def test_oracle_matches_golden_bitwise(oracle, golden_dir, name):
    d, p = _load(os.path.join(golden_dir, "ref_%s.npz" % name))
    n, edge = int(d["n"]), float(d["edge"])
    coef = getattr(oracle, POLY[int(d["law"])])
    o = oracle.run(p, *boxes(n, edge), float(d["rsm"]), float(d["theta"]), int(d["ppn"]), coef=coef)
    v = by_id(o)
    # bit-for-bit: same tree, same lists, same arithmetic order as the compiled reference
    assert np.array_equal(v["vx"], d["vx"]) and np.array_equal(v["vy"], d["vy"]) and np.array_equal(v["vz"], d["vz"])
    st = o["stats"]
    assert st["nodes"] == int(d["nodes"]) and st["leaves"] == int(d["leaves"])
    assert st["empty_leaves"] == int(d["empty_leaves"]) and st["max_ppn"] == int(d["max_ppn"])
    assert st["pairs_eval"] == int(d["pairs_eval"]) and st["pairs_incut"] == int(d["pairs_incut"])
    t = o["tree"]
    leaf = (t["cl"] == 0) & (t["cr"] == 0) & (t["count"] > 0)
    assert np.array_equal(t["offset"][leaf], d["leaf_offset"]) and np.array_equal(t["count"][leaf], d["leaf_count"])


QUAD = ["quad_clustered6k_ppn32", "quad_clustered20k_ppn64_theta03", "quad_lattice16_ppn64"]


@pytest.mark.parametrize("name", QUAD)
def test_quadrupole_oracle_matches_golden_bitwise(oracle, golden_dir, name):
    """RCBQuadrupoleForceTree (TDPTS = 12, -S; RCBForceTree.cxx:229-272,519-569,856-889): the restatement reproduces
    the compiled reference bit for bit -- kicks, pair counts, and the radius and 12 masses of every node's
    pseudo-particles."""
    d, p = _load(os.path.join(golden_dir, "ref_%s.npz" % name))
    assert int(d["tdpts"]) == 12
    n, edge = int(d["n"]), float(d["edge"])
    o = oracle.run(p, *boxes(n, edge), float(d["rsm"]), float(d["theta"]), int(d["ppn"]), tdpts=12)
    v = by_id(o)
    assert np.array_equal(v["vx"], d["vx"]) and np.array_equal(v["vy"], d["vy"]) and np.array_equal(v["vz"], d["vz"])
    st, t = o["stats"], o["tree"]
    assert st["nodes"] == int(d["nodes"]) and st["pairs_eval"] == int(d["pairs_eval"]) and st["pairs_incut"] == int(d["pairs_incut"])
    big = t["count"] > 12
    assert np.array_equal(t["offset"][big], d["pp_offset"]) and np.array_equal(t["count"][big], d["pp_count"])
    assert np.array_equal(t["tdr"][big], d["pp_tdr"]) and np.array_equal(t["ppm12"][big], d["pp_ppm12"])
    # the 12 masses carry the node's monopole: their sum is the node's particle count (unit masses) to rounding
    real = (t["count"] > 12) & ~((t["cl"] == 0) & (t["cr"] == 0) & (t["count"] > int(d["ppn"])))
    m = t["ppm12"][real].astype(np.float64)
    assert np.all(np.abs(m.sum(axis=1) - t["count"][real]) <= 1e-4 * np.abs(m).sum(axis=1))


def test_quadrupole_changes_the_far_field_only(oracle):
    """Same tree and same lists as the monopole run except that an accepted node contributes 12 entries instead of 1;
    on a clustered snapshot the two kicks differ by the (small) quadrupole correction."""
    p = synth.clustered(20000, 24.0, seed=15)
    b = boxes(24)
    o1 = oracle.run(p, *b, RSM, 0.5, 64)
    o12 = oracle.run(p, *b, RSM, 0.5, 64, tdpts=12)
    for k in ("count", "offset", "cl", "cr", "xmin", "xmax", "xc"):
        assert np.array_equal(o1["tree"][k], o12["tree"][k]), k
    a, c = by_id(o1), by_id(o12)
    d = np.sqrt(sum((a[k].astype(np.float64) - c[k]) ** 2 for k in ("vx", "vy", "vz")))
    nrm = np.sqrt(sum(a[k].astype(np.float64) ** 2 for k in ("vx", "vy", "vz")))
    rel = d / np.maximum(nrm, 1e-3)
    assert (d > 0).mean() > 0.01 and np.quantile(rel, 0.99) < 0.05     # only sinks that accepted a node change, and little
    assert o12["stats"]["pairs_eval"] > o1["stats"]["pairs_eval"]


def test_root_leaf_is_listed_twice(oracle, golden_dir):
    """N <= ppn: the reference walks the root leaf against itself and then appends it again
    (RCBForceTree.cxx:947 `tln < tl` is false for the root), so every pair is evaluated twice."""
    d, _ = _load(os.path.join(golden_dir, "ref_lattice8_rootleaf.npz"))
    assert int(d["pairs_eval"]) == 512 * 1024


def _ref():
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return refbind


@pytest.mark.parametrize("kind,n,ppn,theta", [("lattice", 20, 512, 0.5), ("lattice", 20, 37, 0.5),
                                              ("clustered", 24, 64, 0.5), ("clustered", 24, 64, 0.3),
                                              ("zeld", 24, 128, 0.5)])
def test_oracle_matches_compiled_reference(oracle, kind, n, ppn, theta):
    R = _ref()
    if kind == "lattice":
        p = synth.jitter_lattice(n, seed=101)
    elif kind == "clustered":
        p = synth.clustered(9000, float(n), seed=102, n_clumps=8)
    else:
        p = synth.zeldovich(n, z=50.0, seed=103, ghost=0)
    b = boxes(n)
    q, st, tree = R.rcb_kick(p, *b, RSM, theta, ppn, count_pairs=True, keep_tree=True)
    o = oracle.run(p, *b, RSM, theta, ppn)
    for k in ("x", "y", "z", "vx", "vy", "vz", "id"):
        assert np.array_equal(q[k], o[k]), k      # same permutation, same kicks, bit for bit
    for k in ("count", "offset", "cl", "cr", "xmin", "xmax", "xc", "ppm"):
        assert np.array_equal(tree[k], o["tree"][k]), k
    assert st["pairs_eval"] == o["stats"]["pairs_eval"] and st["pairs_incut"] == o["stats"]["pairs_incut"]


@pytest.mark.parametrize("kind,n,ppn,theta", [("clustered", 24, 64, 0.5), ("clustered", 24, 20, 0.3), ("lattice", 16, 64, 0.5)])
def test_quadrupole_oracle_matches_compiled_reference(oracle, kind, n, ppn, theta):
    R = _ref()
    p = synth.jitter_lattice(n, seed=111) if kind == "lattice" else synth.clustered(9000, float(n), seed=112, n_clumps=8)
    b = boxes(n)
    q, st, tree = R.rcb_kick(p, *b, RSM, theta, ppn, count_pairs=True, keep_tree=True, tdpts=12, vmax=True)
    o = oracle.run(p, *b, RSM, theta, ppn, tdpts=12)
    for k in ("x", "y", "z", "vx", "vy", "vz", "id"):
        assert np.array_equal(q[k], o[k]), k
    for k in ("count", "offset", "cl", "cr", "xmin", "xmax", "xc", "tdr", "ppm12"):
        assert np.array_equal(tree[k], o["tree"][k]), k
    assert st["pairs_eval"] == o["stats"]["pairs_eval"] and st["pairs_incut"] == o["stats"]["pairs_incut"]


def test_force_law_matches_compiled_reference(oracle):
    R = _ref()
    r2 = np.concatenate([np.linspace(0, 12, 4001), np.geomspace(1e-8, 1e3, 2000)]).astype(np.float32)
    assert np.array_equal(oracle.force_law_eval(r2, RSM), R.force_law_eval(R.LAW_POLY5, r2, RSM))
    assert np.array_equal(oracle.force_law_eval(r2, RSM, coef=oracle.POLY6), R.force_law_eval(R.LAW_POLY6, r2, RSM))


def test_force_law_known_values(oracle):
    """Closed-form anchors: the polynomial grid force at r2 = 0 is a0, f vanishes beyond rmax^2, and the
    poly5 law has the small step at the cutoff noted in SURVEY.md 8(c) (1/rmax^3 - g(rmax) = 1.26e-5)."""
    rmax2 = np.float32(oracle.RMAX) ** 2
    f = oracle.force_law_eval(np.array([rmax2 * 0.999999, rmax2 * 1.001, 100.0], np.float32), 0.0)
    assert f[1] == 0.0 and f[2] == 0.0
    assert abs(float(f[0]) - 1.26e-5) < 2e-6
    f0 = oracle.force_law_eval(np.array([0.0], np.float32), RSM)[0]
    assert abs(float(f0) - (RSM ** -3 - float(oracle.POLY5[0]))) / RSM ** -3 < 1e-6


def test_tree_matches_fp64_direct_sum(oracle):
    """The role of the reference's ForceTreeTest (src/halo_finder/ForceTreeTest.cxx:277-302): tree force vs
    direct O(N) sum with the same law.  Near-uniform lattice: no monopole is accepted, so
    the tree result is the direct sum up to FP32 rounding."""
    n = 16
    p = synth.jitter_lattice(n, seed=5)
    o = oracle.run(p, *boxes(n, 0.0), RSM, 0.5, 64)
    sel = np.arange(0, n ** 3, 37, dtype=np.int64)
    a64 = oracle.direct_sum(p, sel, RSM)
    v = by_id(o)
    a = np.stack([v["vx"][sel], v["vy"][sel], v["vz"][sel]], axis=1)
    gross = by_id(oracle.run(p, *boxes(n, 0.0), RSM, 0.5, 64, form=oracle.FORM_GROSS))["vx"][sel]
    err = np.sqrt(((a - a64) ** 2).sum(axis=1)) / gross
    assert err.max() < 1e-5


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65])
def test_ragged_sizes(oracle, n):
    rng = np.random.default_rng(n)
    q = synth._pack(rng.random(n) * 8, rng.random(n) * 8, rng.random(n) * 8)
    o = oracle.run(q, *boxes(8, 0.0), RSM, 0.5, 64)
    assert sorted(o["id"].tolist()) == list(range(n))
    if n == 65:
        assert o["stats"]["nodes"] == 3
    elif n > 0:
        assert o["stats"]["nodes"] == 1
    assert np.all(np.isfinite(o["vx"]))


def test_coincident_particles_degenerate_split(oracle):
    """All particles at one point: the split leaves one side empty, the node stays an oversized leaf with
    two orphan children (RCBForceTree.cxx:727-729) counted as empty leaves by printStats."""
    n = 40
    q = synth._pack(np.full(n, 3.0), np.full(n, 4.0), np.full(n, 5.0))
    o = oracle.run(q, *boxes(8, 0.0), RSM, 0.5, 16)
    assert o["stats"]["nodes"] == 3 and o["stats"]["empty_leaves"] == 2
    assert np.all(o["vx"] == 0) and np.all(o["vy"] == 0) and np.all(o["vz"] == 0)
