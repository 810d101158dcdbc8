// host/RCBForceTree.cxx -- the reference's force-tree constructor as a thin C++ facade over libhaccsr.
// Replaces the link-level unit RCBForceTree.o + ForceLaw.o + BGQStep16.o + BGQCM.o of libBHForceTree.a
// (reference src/halo_finder/Makefile:211-221).  See INTEGRATION.md for the two-line Makefile change.
#include "RCBForceTree.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace {

struct Shared {
  std::mutex mu;
  haccsr_ctx *ctx = nullptr;
  int64_t cap = 0;
  int device = -1;
  int arith = -1;     // HACCSR_ARITH_*; -1 = $HACCSR_ARITH ("x86" / "fused") or the library default
  int cull = -1;      // warp-level culling; -1 = $HACCSR_CULL or off
  int rank = -1;      // -1 = $HACCSR_RANK or 0
};
Shared &shared() { static Shared s; return s; }

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "RCBForceTree (libhaccsr facade): %s: %s\n", what, haccsr_last_error());
  abort();
}

int pick_device() {
  Shared &s = shared();
  if (s.device >= 0) return s.device;
  const char *e = getenv("HACCSR_DEVICE");
  return e ? atoi(e) : 0;
}

// one context per process, grown when a larger particle count arrives (the reference rebuilds its tree
// object every sub-cycle but keeps the node memory in bigchunk; device allocations are kept the same way)
haccsr_ctx *context_for(int64_t count) {
  Shared &s = shared();
  if (s.ctx && count <= s.cap) return s.ctx;
  if (s.ctx) { haccsr_destroy(s.ctx); s.ctx = nullptr; s.cap = 0; }
  int64_t want = count + count / 8 + 1024;
  if (haccsr_create(&s.ctx, pick_device(), want) != 0) die("cannot create a device context (no CPU fallback)");
  s.cap = want;
  return s.ctx;
}

}  // namespace

void haccsr_facade_set_device(int device) { shared().device = device; }
void haccsr_facade_set_arithmetic(int mode) { shared().arith = mode; }
void haccsr_facade_set_culling(int on) { shared().cull = on ? 1 : 0; }
void haccsr_facade_set_rank(int rank) { shared().rank = rank; }

void haccsr_facade_release() {
  Shared &s = shared();
  std::lock_guard<std::mutex> lock(s.mu);
  if (s.ctx) haccsr_destroy(s.ctx);
  s.ctx = nullptr; s.cap = 0;
}

void haccsr_facade_pin_arrays(ID_T capacity, POSVEL_T *x, POSVEL_T *y, POSVEL_T *z, POSVEL_T *vx, POSVEL_T *vy,
                              POSVEL_T *vz, POSVEL_T *mass, POSVEL_T *phi, ID_T *id, MASK_T *mask) {
  POSVEL_T *f[8] = {x, y, z, vx, vy, vz, mass, phi};
  for (int i = 0; i < 8; ++i)
    if (f[i] && haccsr_host_register(f[i], (size_t)capacity * sizeof(POSVEL_T)) != 0) die("cudaHostRegister");
  if (id && haccsr_host_register(id, (size_t)capacity * sizeof(ID_T)) != 0) die("cudaHostRegister");
  if (mask && haccsr_host_register(mask, (size_t)capacity * sizeof(MASK_T)) != 0) die("cudaHostRegister");
}

template <int TDPTS>
RCBForceTree<TDPTS>::RCBForceTree(POSVEL_T *minLoc, POSVEL_T *maxLoc, POSVEL_T *minForceLoc, POSVEL_T *maxForceLoc,
                                  ID_T count, POSVEL_T *xLoc, POSVEL_T *yLoc, POSVEL_T *zLoc, POSVEL_T *xVel,
                                  POSVEL_T *yVel, POSVEL_T *zVel, POSVEL_T *mass, POSVEL_T *phiLoc, ID_T *idLoc,
                                  MASK_T *maskLoc, POSVEL_T /*avgMass*/, POSVEL_T fsm, POSVEL_T r, POSVEL_T oa, ID_T nd,
                                  ID_T /*ds*/, ID_T /*tmin*/, ForceLaw *fl, float fcoeff, POSVEL_T ppc)
    : particleCount(count) {
  static_assert(sizeof(POSVEL_T) == 4 && sizeof(ID_T) == 8 && sizeof(MASK_T) == 2,
                "libhaccsr is built for the reference's -DID_64 -DPOSVEL_32 types (include.mk:4)");
  memset(&m_stats, 0, sizeof(m_stats));
  Shared &s = shared();
  std::lock_guard<std::mutex> lock(s.mu);      // the reference constructor is not re-entrant either
  // the pseudo-particle contraction of the quadrupole tree is the reference's default 0.9 on the device (tree_build.cu pp_tdr;
  // no shipped call site passes anything else, Particles.cxx:1341-1366); another value would silently change tdr and the masses
  if (TDPTS != 1 && ppc != 0.9f) {
    fprintf(stderr, "RCBForceTree (libhaccsr facade): ppContract = %g is not supported (the device uses the reference's default 0.9)\n", ppc);
    abort();
  }
  haccsr_ctx *ctx = context_for(count);

  // the force law: fl == NULL means Newton with fcoeff = 1 (RCBForceTree.cxx:395-404)
  HaccsrLawDescription d;
  if (!fl) {
    d.kind = HACCSR_LAW_NEWTON; fcoeff = 1.0f;
  } else if (!fl->haccsr_describe(d)) {
    fprintf(stderr, "RCBForceTree (libhaccsr facade): this ForceLaw subclass does not implement haccsr_describe(); "
                    "a host functor cannot be evaluated per pair on the GPU and there is no CPU fallback\n");
    abort();
  }
  // the cutoff is the constructor's fsm (RCBForceTree.cxx:379,582), the softening its r for the BG/Q kernel
  // (:589) but the law's own rsm in the generic kernel (ForceLaw.cxx:187); the reference always passes the
  // same value for both (Particles.cxx:183,1329), and so must the caller here
  const float rsm = (d.kind == HACCSR_LAW_NEWTON) ? 0.0f : d.rsm;
  if (d.kind != HACCSR_LAW_NEWTON && rsm != r) {
    fprintf(stderr, "RCBForceTree (libhaccsr facade): rsm of the ForceLaw (%g) differs from the constructor's r (%g)\n", rsm, r);
    abort();
  }
  const float *coef = d.coef;
  int ncoef = d.ncoef;
  if (d.kind == HACCSR_LAW_SR_INTERP) { coef = d.table.data(); ncoef = (int)d.table.size(); }
  if (haccsr_set_force_law(ctx, d.kind, coef, ncoef, rsm, fsm) != 0) die("haccsr_set_force_law");
  int arith = s.arith;
  if (arith < 0) { const char *e = getenv("HACCSR_ARITH"); arith = (e && !strcmp(e, "x86")) ? HACCSR_ARITH_X86 : HACCSR_ARITH_FUSED; }
  if (haccsr_set_arithmetic(ctx, arith) != 0) die("haccsr_set_arithmetic");
  int cull = s.cull;
  if (cull < 0) { const char *e = getenv("HACCSR_CULL"); cull = (e && atoi(e) != 0) ? 1 : 0; }
  if (haccsr_set_culling(ctx, cull) != 0) die("haccsr_set_culling");

  // upload -> tree build, lists, force kernel, kick -> download, transfers overlapped with the kernels
  if (haccsr_kick_host(ctx, count, xLoc, yLoc, zLoc, xVel, yVel, zVel, mass, phiLoc, idLoc, maskLoc, minLoc, maxLoc,
                       minForceLoc, maxForceLoc, oa, nd, TDPTS, fcoeff, nullptr, &m_stats) != 0)
    die("haccsr_kick_host");
  int rank = s.rank;
  if (rank < 0) { const char *e = getenv("HACCSR_RANK"); rank = e ? atoi(e) : 0; }
  if (rank == 0 && !getenv("HACCSR_QUIET")) printStats(1e-3 * m_stats.ms_build);     // rank 0 only (RCBForceTree.cxx:503)
}

template <int TDPTS>
RCBForceTree<TDPTS>::~RCBForceTree() {}

template <int TDPTS>
void RCBForceTree<TDPTS>::printStats(double buildTime) {
  printf("\ttree post-build statistics (local for rank 0):\n");
  printf("\t\tparticles: %.2f\n", (double)particleCount);
  printf("\t\tnodes: %.2f (allocated:  %.2f)\n", (double)m_stats.nodes, (double)m_stats.nodes);
  printf("\t\tleaves: %.2f (empty: %.2f)\n", (double)m_stats.leaves, (double)m_stats.empty_leaves);
  printf("\t\tmean ppn: %.2f (max ppn: %lu)\n", m_stats.mean_ppn, (unsigned long)m_stats.max_ppn);
  printf("\t\tbuild time: %g s\n", buildTime);
}

template class RCBForceTree<QUADRUPOLE_TDPTS>;
template class RCBForceTree<MONOPOLE_TDPTS>;
