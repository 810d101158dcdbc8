"""CPU tests of the drop-in boundary: libhaccsr.so loads, exports every symbol include/haccsr.h declares,
and refuses loudly (no CPU fallback) when no B200 is visible."""
import ctypes
import os
import re

import pytest

import hacc_coral_b200 as H
from hacc_coral_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "haccsr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(haccsr_[a-z_0-9]+)\s*\(", src)))


def test_library_is_built_in_tree():
    H.build()
    assert os.path.exists(H.lib_path())
    assert os.path.dirname(H.lib_path()).endswith(os.path.join("hacc_coral_b200", "csrc"))


def test_exports_every_declared_symbol():
    H.build()
    lib = ctypes.CDLL(H.lib_path())
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libhaccsr.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names


def test_stats_struct_layout_matches_header():
    src = open(os.path.join(ROOT, "include", "haccsr.h")).read()
    body = re.search(r"typedef struct haccsr_stats \{(.*?)\} haccsr_stats;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int64_t|uint64_t|double|float|int32_t)\s+([a-z_]+);", body)
    assert fields == [f for f, _ in capi.KickStats._fields_]


def test_sass_is_blackwell_native():
    """The force kernel must contain TMA bulk copies (UBLKCP) and MUFU.RSQ; built for sm_100a only."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    H.build()
    elf = subprocess.run(["cuobjdump", "-lelf", H.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", H.lib_path()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "MUFU.RSQ" in sass and "SYNCS" in sass


def test_fails_loudly_without_gpu():
    lib = H.load_library()
    if lib.haccsr_device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(H.HaccSRError) as e:
        H.HaccSR(1000)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "hacc_coral_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cxx", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oraclebind" not in txt and "refbind" not in txt and "liboracle" not in txt, f


def test_integration_level2_snippet_compiles(tmp_path):
    """The reference-side change shown in INTEGRATION.md (Level 2: Particles::subCycle on the device) is compiled against
    include/haccsr.h and the facade's ForceLaw.h with stand-ins for the reference's Particles / TimeStepper / Domain
    members it touches (src/cpu/Particles.h:160-205, src/simulation/TimeStepper.h, Domain.h) -- the document cannot rot
    against the C ABI."""
    import subprocess
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```cpp\n(.*?)```", doc, flags=re.S)
    snippet = [b for b in blocks if "void Particles::subCycle" in b]
    assert len(snippet) == 1
    stub = r'''
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "ForceLaw.h"
struct TimeStepper { float pp() const; float adot() const; float fscal() const; float tau() const; float tau2() const; };
struct Domain { static void ng_local_total(int *); };
static int local_gpu = 0;
class Particles {
 public:
  void subCycle(TimeStepper *gts);
 private:
  float *m_xArr, *m_yArr, *m_zArr, *m_vxArr, *m_vyArr, *m_vzArr, *m_massArr, *m_phiArr;
  int64_t *m_idArr; uint16_t *m_maskArr;
  int64_t m_Np_local_total; int m_nsub, m_rcbTreePPN;
  float m_edge, m_alpha, m_gpscal, m_openAngle, m_fsrrmax;
  ForceLaw *m_fl;
};
'''
    src = tmp_path / "level2.cxx"
    src.write_text(stub + snippet[0])
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "hacc_coral_b200", "host"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_map2_glue_matches_the_reference_expressions():
    """haccsr_map2_setup / haccsr_map1_factor against the expressions of Particles::map2 (reference src/cpu/Particles.cxx:
    1213-1233) and Particles::map1 (:745-754) re-evaluated here with numpy under C's promotion rules, and against values
    worked out by hand for the shipped configuration (indat: ng = np, edge 3.2, nsub 5)."""
    import numpy as np
    f32, f64 = np.float32, np.float64
    # hand-computed: nglt = 151^3 (C1), edge = 3.2f, gpscal = 1, fscal = 1.5, tau = 0.1, nsub = 5
    m = capi.map2_setup((151, 151, 151), 3.2, 1.0, 1.5, 0.1, 1.0 / 5)
    assert m["tree_lo"] == [0.0] * 3 and m["tree_hi"] == [151.0] * 3
    assert m["force_lo"] == [float(f32(3.2))] * 3
    assert m["force_hi"] == [float(f32(151.0 - f64(f32(3.2))))] * 3 == [147.8000030517578] * 3
    pi32 = f32(4.0 * f64(np.arctan(f32(1.0), dtype=f32)))
    assert float(pi32) == 3.1415927410125732
    assert m["fcoeff"] == float(f32(1.0 / 4.0 / f64(pi32) * 1.5 * 0.1 * 0.2)) == 0.0023873241152614355
    # general case, non-cubic sub-volume and gpscal != 1: every intermediate in the reference's type
    rng = np.random.default_rng(3)
    for _ in range(20):
        nglt = [int(t) for t in rng.integers(20, 400, 3)]
        edge, gp = f32(rng.uniform(1, 5)), f32(rng.uniform(0.5, 2.0))
        fscal, tau, sf = rng.uniform(0.5, 3), rng.uniform(0.01, 0.2), 1.0 / int(rng.integers(1, 9))
        m = capi.map2_setup(nglt, float(edge), float(gp), fscal, tau, sf)
        assert m["tree_hi"] == [float(f32(1.0 * max(nglt)))] * 3
        assert m["force_hi"] == [float(f32(1.0 * n - f64(edge))) for n in nglt]
        divscal = f32(f32(gp * gp) * gp)
        assert m["fcoeff"] == float(f32(f64(divscal) / 4.0 / f64(pi32) * fscal * tau * sf))
        pp, t1, adot, alpha = f32(rng.uniform(0.1, 1)), f32(rng.uniform(0.001, 0.1)), f32(rng.uniform(0.5, 2)), f32(1.0)
        got = capi.map1_factor(float(pp), float(t1), float(adot), float(alpha))
        pf = f32(np.power(pp, f32(1.0 + 1.0 / f64(alpha)), dtype=f32))
        want = f32(f32(1.0 / f64(f32(f32(alpha * adot) * pf))) * t1)
        assert abs(got - float(want)) <= 2e-7 * abs(float(want))     # powf of libm vs numpy: last-bit differences only


def test_refresh_stats_struct_layout_matches_header():
    src = open(os.path.join(ROOT, "include", "haccsr.h")).read()
    body = re.search(r"typedef struct haccsr_refresh_stats \{(.*?)\} haccsr_refresh_stats;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"\b(?:int64_t|int32_t|float)\s+([a-z_, ]+);", body):
        fields += [t.strip() for t in decl.split(",")]
    assert fields == [f for f, _ in capi.RefreshStats._fields_]
