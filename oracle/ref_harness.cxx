// oracle/ref_harness.cxx -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" harness around the *unmodified, compiled* reference short-range force tree
// (reference src/halo_finder/RCBForceTree.{h,cxx}, ForceLaw.{h,cxx}, BGQCM.c, bigchunk.c).  It is
// built by oracle/build_ref.sh into oracle/_ref/libhaccref.so (git-ignored) and is used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the checker
// and as the timed CPU baseline.  Nothing under hacc_coral_b200/ links or loads it.
//
// What it does: construct RCBMonopoleForceTree exactly like the reference call site
// (reference src/cpu/Particles.cxx:1313-1338) on caller-provided arrays, with a selectable force law:
//   law 0  poly5  ForceLawSR over a 5th-order polynomial grid force (the north-star law; coefficients
//                 reference src/halo_finder/ForceLaw.cxx:107-116 == BGQStep16.c:167), evaluated through
//                 the reference's own ForceLawSR::f_over_r (ForceLaw.cxx:187-192) and nbody1 generic
//                 branch (RCBForceTree.cxx:601-618).  The polynomial evaluator is a local FGridEval
//                 subclass because the reference hard-wires POLY_ORDER 6 (ForceLaw.cxx:8).
//   law 1  poly6  the reference's FGridEvalPoly as shipped (-P).
//   law 2  fit    the reference's FGridEvalFit (run_hacc.sh default).
//   law 3  newton fl == NULL (RCBForceTree.cxx:395-404).
//   law 4  interp the reference's FGridEvalInterp with ncoef samples (-i n).
// A counting wrapper (optional) counts evaluated / in-cutoff pairs: the generic nbody1 calls
// f_over_r exactly once per evaluated pair (RCBForceTree.cxx:610).
// A derived probe class reads the protected node vector after construction so the tree topology of the
// real reference can be compared with ours.

#include "RCBForceTree.h"
#include "ForceLaw.h"

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <vector>
#include <unistd.h>
#include <fcntl.h>
#include <pthread.h>

namespace {

// g(r2) as a polynomial of caller-given order <= 6, with the same window as FGridEvalPoly::eval
// (reference ForceLaw.cxx:137-141).
class FGridEvalPolyN : public FGridEval {
public:
  FGridEvalPolyN(FGrid *fg, const float *a, int n) {
    for (int i = 0; i < 7; ++i) m_a[i] = (i < n) ? a[i] : 0.0f;
    m_r2min = 0.0f;
    m_r2max = fg->rmax() * fg->rmax();
  }
  float eval(float r2) {
    float ret = m_a[0] + r2*(m_a[1] + r2*(m_a[2] + r2*(m_a[3] + r2*(m_a[4] + r2*(m_a[5] + r2*m_a[6])))));
    return ret * (r2 >= m_r2min) * (r2 <= m_r2max);
  }
  float r2min() { return m_r2min; }
  float r2max() { return m_r2max; }
private:
  float m_a[7], m_r2min, m_r2max;
};

struct alignas(64) PairSlot { uint64_t eval, incut; };
static PairSlot g_slots[1024];
static std::atomic<int> g_next_slot(0);
static thread_local int t_slot = -1;

class CountingForceLaw : public ForceLaw {
public:
  CountingForceLaw(ForceLaw *inner, float r2max) : m_inner(inner), m_r2max(r2max) {}
  float f_over_r(float r2) {
    if (t_slot < 0) t_slot = g_next_slot.fetch_add(1) & 1023;
    PairSlot &s = g_slots[t_slot];
    // nested OpenMP teams create fresh threads, so slots can be shared after 1024 threads: count atomically
    __atomic_fetch_add(&s.eval, 1, __ATOMIC_RELAXED);
    if (r2 > 0.0f && r2 < m_r2max) __atomic_fetch_add(&s.incut, 1, __ATOMIC_RELAXED);
    return m_inner->f_over_r(r2);
  }
private:
  ForceLaw *m_inner;
  float m_r2max;
};

struct NodeDump {
  std::vector<int64_t> count, offset, cl, cr;
  std::vector<float> box;  // 10 floats per node: xmin[3] xmax[3] xc[3] ppm0
  std::vector<float> pp13; // TDPTS = 12 runs: tdr, ppm[12] per node
};
static NodeDump g_nodes;
static int g_tdpts = 1;      // 1: RCBMonopoleForceTree (-R), 12: RCBQuadrupoleForceTree (-S), RCBForceTree.h:202-203

struct ProbeBase {
  virtual ~ProbeBase() {}
  virtual void dump(NodeDump &d) const = 0;
};

template <int TD>
class ProbeTreeT : public RCBForceTree<TD>, public ProbeBase {
public:
  ProbeTreeT(float *lo, float *hi, float *flo, float *fhi, int64_t n, float *x, float *y, float *z,
             float *vx, float *vy, float *vz, float *m, float *phi, int64_t *id, uint16_t *mask,
             float rmax, float rsm, float theta, int64_t ppn, int64_t ds, int64_t tmin, ForceLaw *fl,
             float fcoeff)
      : RCBForceTree<TD>(lo, hi, flo, fhi, n, x, y, z, vx, vy, vz, m, phi, (ID_T *)id,
                         (MASK_T *)mask, 1.0f, rmax, rsm, theta, ppn, ds, tmin, fl, fcoeff) {}
  void dump(NodeDump &d) const {
    size_t n = this->tree.size();
    d.count.resize(n); d.offset.resize(n); d.cl.resize(n); d.cr.resize(n); d.box.resize(10 * n);
    d.pp13.assign(TD == 12 ? 13 * n : 0, 0.0f);
    for (size_t i = 0; i < n; ++i) {
      d.count[i] = this->tree[i].count; d.offset[i] = this->tree[i].offset;
      d.cl[i] = this->tree[i].cl; d.cr[i] = this->tree[i].cr;
      for (int k = 0; k < 3; ++k) {
        d.box[10*i + k] = this->tree[i].xmin[k];
        d.box[10*i + 3 + k] = this->tree[i].xmax[k];
        d.box[10*i + 6 + k] = this->tree[i].xc[k];
      }
      d.box[10*i + 9] = this->tree[i].ppm[0];
      if (TD == 12) {
        d.pp13[13*i] = this->tree[i].tdr;
        for (int q = 0; q < 12; ++q) d.pp13[13*i + 1 + q] = this->tree[i].ppm[q];
      }
    }
  }
};

struct CtorArgs {
  float *lo, *hi, *flo, *fhi; int64_t n;
  float *x, *y, *z, *vx, *vy, *vz, *mass, *phi; int64_t *id; uint16_t *mask;
  float rmax, rsm, theta; int64_t ppn, ds, tmin; ForceLaw *fl; float fcoeff;
  ProbeBase *out;
};
static void *ctor_thread(void *p) {
  CtorArgs *a = (CtorArgs *)p;
  if (g_tdpts == 12)
    a->out = new ProbeTreeT<12>(a->lo, a->hi, a->flo, a->fhi, a->n, a->x, a->y, a->z, a->vx, a->vy, a->vz,
                                a->mass, a->phi, a->id, a->mask, a->rmax, a->rsm, a->theta, a->ppn, a->ds,
                                a->tmin, a->fl, a->fcoeff);
  else
    a->out = new ProbeTreeT<1>(a->lo, a->hi, a->flo, a->fhi, a->n, a->x, a->y, a->z, a->vx, a->vy, a->vz,
                               a->mass, a->phi, a->id, a->mask, a->rmax, a->rsm, a->theta, a->ppn, a->ds,
                               a->tmin, a->fl, a->fcoeff);
  return 0;
}

static double now_s() {
  timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

}  // namespace

extern "C" {

struct ref_stats {
  int64_t nodes, leaves, empty_leaves, max_ppn;
  double mean_ppn;
  double wall_s;            // whole constructor (build + walk + force)
  uint64_t pairs_eval, pairs_incut;  // only when count_pairs != 0
};

// Select the instantiation the next ref_rcb_kick constructs: 1 = RCBMonopoleForceTree, 12 = RCBQuadrupoleForceTree.
void ref_set_tdpts(int t) { g_tdpts = (t == 12) ? 12 : 1; }
// tdr + ppm[12] per node of the last ref_rcb_kick(keep_tree=1) call made with TDPTS = 12.
int ref_tree_get_pp12(int64_t cap, float *pp13) {
  int64_t n = (int64_t)g_nodes.count.size();
  if (cap < n || g_nodes.pp13.size() != (size_t)(13 * n)) return 1;
  memcpy(pp13, g_nodes.pp13.data(), 13 * n * sizeof(float));
  return 0;
}

// rmax of the reference's FGrid (ForceLaw.cxx:32)
float ref_rmax(void) { FGrid fg; return fg.rmax(); }

// The reference's own interpolation table of the grid force (FGrid::fgor_r2_interp, ForceLaw.cxx:54-67):
// what a caller hands to haccsr_set_force_law(HACCSR_LAW_SR_INTERP, ...).
int ref_fgrid_table(int n, float *f_out) {
  FGrid fg;
  FGridEvalInterp ev(&fg, n);
  memcpy(f_out, ev.f(), (size_t)n * sizeof(float));
  return 0;
}
// the eight constants of the analytic fit are protected members of FGrid (ForceLaw.h:21); a probe subclass reads them
struct FGridProbe : public FGrid {
  void get(float *o) { o[0] = m_b; o[1] = m_c; o[2] = m_d; o[3] = m_e; o[4] = m_f; o[5] = m_g; o[6] = m_h; o[7] = m_l; }
};
int ref_fgrid_constants(float *out8) { FGridProbe p; p.get(out8); return 0; }

// f_over_r of the selected law, for force-law parity tests (ForceLaw.cxx:187-192).
int ref_force_law_eval(int law, const float *coef, int ncoef, float rsm, int64_t n, const float *r2, float *out) {
  FGrid fg;
  FGridEval *ev = 0;
  if (law == 0) ev = new FGridEvalPolyN(&fg, coef, ncoef);
  else if (law == 1) ev = new FGridEvalPoly(&fg);
  else if (law == 2) ev = new FGridEvalFit(&fg);
  else if (law == 4) ev = new FGridEvalInterp(&fg, ncoef);
  else return 1;
  ForceLawSR sr(ev, rsm);
  for (int64_t i = 0; i < n; ++i) out[i] = sr.f_over_r(r2[i]);
  delete ev;
  return 0;
}

// Run the reference constructor once.  boxes = treeLo[3] treeHi[3] forceLo[3] forceHi[3].
// Arrays are permuted in place into the reference's tree order and vx/vy/vz are kicked, exactly as
// Particles::map2 sees it.  quiet != 0 sends the reference's printStats text to /dev/null.
int ref_rcb_kick(int law, const float *coef, int ncoef, int count_pairs, int quiet, int64_t n,
                 float *x, float *y, float *z, float *vx, float *vy, float *vz, float *mass,
                 float *phi, int64_t *id, uint16_t *mask, const float *boxes, float rsm, float theta,
                 int64_t ppn, int64_t ds, int64_t tmin, float fcoeff, int keep_tree, ref_stats *st) {
  FGrid fg;
  FGridEval *ev = 0;
  ForceLaw *fl = 0, *cfl = 0;
  if (law == 0) ev = new FGridEvalPolyN(&fg, coef, ncoef);
  else if (law == 1) ev = new FGridEvalPoly(&fg);
  else if (law == 2) ev = new FGridEvalFit(&fg);
  else if (law == 4) ev = new FGridEvalInterp(&fg, ncoef);
  else if (law != 3) return 1;
  if (ev) fl = new ForceLawSR(ev, rsm);
  ForceLaw *use = fl;
  if (count_pairs && fl) {
    for (int i = 0; i < 1024; ++i) g_slots[i].eval = g_slots[i].incut = 0;
    cfl = new CountingForceLaw(fl, fg.rmax() * fg.rmax());
    use = cfl;
  }
  float lo[3], hi[3], flo[3], fhi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = boxes[k]; hi[k] = boxes[3+k]; flo[k] = boxes[6+k]; fhi[k] = boxes[9+k]; }

  int saved = -1;
  if (quiet) {
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); close(nul);
  }
  // The reference keeps `int idx[n]` (partition, RCBForceTree.cxx:635) and 4*VMAX floats per walk
  // (:940) on the stack, so the constructor runs on a thread with a large (lazily committed) stack.
  CtorArgs a = {lo, hi, flo, fhi, n, x, y, z, vx, vy, vz, mass, phi, id, mask, fg.rmax(), rsm, theta,
                ppn, ds, tmin, use, fcoeff, 0};
  double t0 = now_s();
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, (size_t)1 << 32);
  pthread_t th;
  if (pthread_create(&th, &attr, ctor_thread, &a) != 0) return 2;
  pthread_join(th, 0);
  pthread_attr_destroy(&attr);
  ProbeBase *t = a.out;
  double t1 = now_s();
  if (quiet) { fflush(stdout); dup2(saved, 1); close(saved); }

  t->dump(g_nodes);
  delete t;
  if (st) {
    memset(st, 0, sizeof(*st));
    st->nodes = (int64_t)g_nodes.count.size();
    int64_t leafParts = 0, nz = 0;
    // same census as RCBForceTree::printStats (RCBForceTree.cxx:468-478): starts at node 1
    for (int64_t i = 1; i < st->nodes; ++i) {
      if (g_nodes.cl[i] == 0 && g_nodes.cr[i] == 0) {
        if (g_nodes.count[i] > 0) { ++nz; leafParts += g_nodes.count[i]; if (g_nodes.count[i] > st->max_ppn) st->max_ppn = g_nodes.count[i]; }
        else st->empty_leaves++;
      }
    }
    st->leaves = nz + st->empty_leaves;
    st->mean_ppn = nz ? leafParts / (double)nz : 0.0;
    st->wall_s = t1 - t0;
    if (cfl) for (int i = 0; i < 1024; ++i) { st->pairs_eval += g_slots[i].eval; st->pairs_incut += g_slots[i].incut; }
  }
  if (!keep_tree) { NodeDump empty; g_nodes = empty; }
  delete cfl; delete fl; delete ev;
  return 0;
}

int64_t ref_tree_size(void) { return (int64_t)g_nodes.count.size(); }

// Copy out the node table of the last ref_rcb_kick(keep_tree=1) call.
int ref_tree_get(int64_t cap, int64_t *count, int64_t *offset, int64_t *cl, int64_t *cr, float *box10) {
  int64_t n = (int64_t)g_nodes.count.size();
  if (cap < n) return 1;
  memcpy(count, g_nodes.count.data(), n * sizeof(int64_t));
  memcpy(offset, g_nodes.offset.data(), n * sizeof(int64_t));
  memcpy(cl, g_nodes.cl.data(), n * sizeof(int64_t));
  memcpy(cr, g_nodes.cr.data(), n * sizeof(int64_t));
  memcpy(box10, g_nodes.box.data(), 10 * n * sizeof(float));
  return 0;
}

}  // extern "C"
