"""hacc_coral_b200 -- HACC's short-range RCB-tree force (reference BenWibking/hacc-coral) on B200.

The product is the C-ABI shared library built from csrc/ (include/haccsr.h) plus the C++ facade in
host/ that mirrors the reference's RCBForceTree constructor.  This Python package is a thin ctypes
binding used by the tests, the benchmark and tools; it never falls back to a CPU implementation.
"""
from .capi import (HaccSR, KickStats, LAW_SR_POLY, LAW_SR_FIT, LAW_SR_INTERP, LAW_NEWTON, ARITH_FUSED, ARITH_X86, ARITH_FUSED_RS3, POLY5, POLY6, RMAX, lib_path, load_library,
                   HaccSRError)
from .build import build

__all__ = ["HaccSR", "KickStats", "LAW_SR_POLY", "LAW_SR_FIT", "LAW_SR_INTERP", "LAW_NEWTON", "ARITH_FUSED", "ARITH_X86", "ARITH_FUSED_RS3", "POLY5", "POLY6", "RMAX", "lib_path",
           "load_library", "HaccSRError", "build"]
