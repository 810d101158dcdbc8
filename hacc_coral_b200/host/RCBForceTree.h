// host/RCBForceTree.h -- header-compatible stand-in for the reference's src/halo_finder/RCBForceTree.h.
//
// Same class template, same 25-parameter constructor with the same defaults and the same typedefs
// (RCBForceTree.h:92-124,202-203), so the reference's call sites compile unchanged against it:
//   src/cpu/Particles.cxx:1313-1338    new RCBMonopoleForceTree(zero, ngltree, lo, hi, Np, x, ..., mask, 1.0,
//                                          m_fsrrmax, m_rsm, m_openAngle, ppn, levels, taskMin, m_fl, c); delete sft;
//   src/halo_finder/ForceTreeTest.cxx:188-237
// As in the reference the constructor does all the work (tree build, interaction lists, force kernel, kick of
// vx/vy/vz) and leaves the ten particle arrays permuted into tree order; here the work runs on the B200
// through the C ABI of libhaccsr (include/haccsr.h).  Errors abort, as the reference's assert/exit(1) do
// (RCBForceTree.cxx:819,1039).  No CPU fallback exists.
#ifndef RCBForceTree_h
#define RCBForceTree_h

#include <stdint.h>

#if defined(__has_include)
#if __has_include("BasicDefinition.h")
#include "BasicDefinition.h"   // building inside the reference tree: take its POSVEL_T / ID_T / MASK_T
#define HACCSR_HAVE_BASIC_DEFINITION 1
#endif
#endif
#ifndef HACCSR_HAVE_BASIC_DEFINITION
// stand-alone: the types of the reference's -DID_64 -DPOSVEL_32 -DGRID_32 build (include.mk:4)
typedef float POSVEL_T;
typedef int64_t ID_T;
typedef uint16_t MASK_T;
#ifndef DIMENSION
#define DIMENSION 3
#endif
#endif

#include "ForceLaw.h"
#include "haccsr.h"

#define QUADRUPOLE_TDPTS 12
#define MONOPOLE_TDPTS 1

template <int TDPTS>
class RCBForceTree {
 public:
  RCBForceTree(POSVEL_T *minLoc, POSVEL_T *maxLoc,            // tree box
               POSVEL_T *minForceLoc, POSVEL_T *maxForceLoc,  // only leaves touching this box are kicked
               ID_T count, POSVEL_T *xLoc, POSVEL_T *yLoc, POSVEL_T *zLoc, POSVEL_T *xVel, POSVEL_T *yVel,
               POSVEL_T *zVel, POSVEL_T *mass, POSVEL_T *phiLoc, ID_T *idLoc, MASK_T *maskLoc,
               POSVEL_T avgMass,  // unused, as in the reference
               POSVEL_T fsm,      // short-range cutoff radius rmax
               POSVEL_T r,        // Plummer softening rsm
               POSVEL_T oa,       // opening angle
               ID_T nd = 1,       // leaf size (-N)
               ID_T ds = 1,       // node-pool depth safety (-L); the device pool grows on demand instead
               ID_T tmin = 128,   // OpenMP task granularity of the reference build; no meaning on the GPU
               ForceLaw *fl = 0, float fcoeff = 0.0, POSVEL_T ppc = 0.9);
  ~RCBForceTree();

  // same text as the reference prints after its build (RCBForceTree.cxx:503-510)
  void printStats(double buildTime);

  // measurement fields of the kick this object performed; not in the reference
  const haccsr_stats &stats() const { return m_stats; }

 protected:
  ID_T particleCount;
  haccsr_stats m_stats;
};

typedef RCBForceTree<QUADRUPOLE_TDPTS> RCBQuadrupoleForceTree;
typedef RCBForceTree<MONOPOLE_TDPTS> RCBMonopoleForceTree;

// ---- optional extras of the facade (not in the reference) -----------------------------------------------
// Device used by the facade's shared context (default: $HACCSR_DEVICE or 0); call before the first tree.
void haccsr_facade_set_device(int device);
// Pair-kernel arithmetic of the following trees (HACCSR_ARITH_FUSED / HACCSR_ARITH_X86, include/haccsr.h);
// default: $HACCSR_ARITH ("fused" or "x86"), else the library default (fused).
void haccsr_facade_set_arithmetic(int mode);
// Warp-level culling in the pair kernel (haccsr_set_culling, include/haccsr.h): skips the force law where no lane of a warp holds
// a pair inside the cutoff; every output bit is unchanged, the kick is about 1.9x faster.  Default: $HACCSR_CULL (0 / 1), else off.
void haccsr_facade_set_culling(int on);
// Rank of this process (Partition::getMyProc()): printStats prints on rank 0 only, like the reference
// (RCBForceTree.cxx:503).  Default: $HACCSR_RANK, else 0.
void haccsr_facade_set_rank(int rank);
// Release the facade's shared context (device memory is otherwise kept between constructor calls, which is
// what bigchunk does for the reference's node pool, bigchunk.h:49-139).
void haccsr_facade_release();
// Page-lock the caller's particle arrays once so every constructor call copies at full PCIe rate.
void haccsr_facade_pin_arrays(ID_T capacity, POSVEL_T *x, POSVEL_T *y, POSVEL_T *z, POSVEL_T *vx, POSVEL_T *vy,
                              POSVEL_T *vz, POSVEL_T *mass, POSVEL_T *phi, ID_T *id, MASK_T *mask);

#endif  // RCBForceTree_h
