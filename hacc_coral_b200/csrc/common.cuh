// common.cuh -- shared declarations of libhaccsr (internal; the public ABI is include/haccsr.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/haccsr.h"

namespace haccsr {

// ---------------------------------------------------------------------------------------------
// Error plumbing: every CUDA call is checked; failures become a status + thread-local message.
void set_error(const char *fmt, ...);
#define HSR_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      haccsr::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)
#define HSR_TRY(call)                                                                          \
  do {                                                                                         \
    int rc__ = (call);                                                                         \
    if (rc__ != 0) return rc__;                                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Device data layout (all in HBM; sizes for N particles):
//   Particle SoA `Soa`: the reference's 10 arrays (src/cpu/Particles.h:195-205), 42 B/particle, two sets
//   (cur / alt) so the permutation into tree order is an out-of-place gather.
//   src4: packed (x,y,z,mass) float4 in TREE ORDER -- the force kernel's source/sink array; any range
//   of it is 16-B aligned so it can be moved by 1-D TMA bulk copies.
struct Soa {
  float *x, *y, *z, *vx, *vy, *vz, *mass, *phi;
  int64_t *id;
  uint16_t *mask;
};

// Tree node, 64 B, read by the walk as four 16-B loads.  Mirrors TreeNode
// (reference src/halo_finder/RCBForceTree.h:131-147) for TDPTS = 1.
struct __align__(16) Node {
  int count, offset, cl, cr;     // particles in node, first particle (tree order), children (0 = none)
  float xmin[3], xmax[3];        // tight bounding box
  float xc[3];                   // centre of mass
  float ppm;                     // monopole pseudo-particle mass
  int parent;
  int split;                     // build-time: split dimension 0..2, or -1 for a leaf
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

// Build-time accumulators of a node: order-independent (integer) so the build is deterministic.
struct NodeAcc {
  unsigned umin[3], umax[3];         // order-preserving uint encoding of float min / max
  unsigned long long lo[4], hi[4];   // exact fixed-point sums of w*x, w*y, w*z, w: total = hi * 2^32 + lo (tree_build.cu add_split)
};

struct ForceLawParams {
  int kind;          // HACCSR_LAW_*
  int ncoef;         // number of polynomial coefficients in use (<= 7)
  float a[8];        // SR_POLY: polynomial coefficients; SR_FIT: b c d e f g h l
  float b[8];        // SR_POLY: the polynomial re-expanded in s = r2 + rsm2 (HACCSR_ARITH_FUSED)
  float rsm2, rmax2, rmax, smax;   // smax = rmax2 + rsm2
  int ntab;          // SR_INTERP: table length and the evaluator's constants (ForceLaw.cxx:145-152)
  float tab_r2min, tab_r2max, tab_oodr2;
};

// Range-list entry flag: start index refers to the pseudo-particle pool.
static constexpr unsigned POOL_FLAG = 0x80000000u;

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return 2;
    }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Exact centroid sums of a finished node (128-bit two's complement, low and high word): a right child's sums are its
// parent's minus its left sibling's.
struct NodeTot {
  unsigned long long lo[4], hi[4];
};

// Build bookkeeping kept on the device so that the levels can be queued without a host round trip (tree_build.cu).
struct BuildState {
  int error;                            // 1 = node pool exhausted
  int unit_mass;                        // written with the root: 1 = every particle mass is exactly 1.0f
  unsigned tile_ticket, node_ticket;    // tile / node-block tickets of the running launch
  int nsplit[128];                      // nodes split at each level (-1 = level not reached, 0 = the tree ends here)
  int lvl_begin[129], lvl_end[129];     // node index range of each level
};

// no_pseudo bit 0: the node's list holds real particles only; bit 1: remainder item (sources over lanes, force.cu)
struct WorkItem { int node, sink_begin, sink_count, no_pseudo; };

}  // namespace haccsr

struct haccsr_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int sm_count = 0;
  int64_t cap = 0;       // particle capacity
  int64_t n_resident = 0;
  uint64_t order_epoch = 0;   // bumped by everything that moves particles to other indices (upload, tree build, compaction, append)
  uint64_t refresh_epoch = 0; // order_epoch when haccsr_refresh_begin made its candidate list
  haccsr::Soa cur{}, alt{};
  haccsr::ForceLawParams law{};
  bool law_set = false;
  int arith = HACCSR_ARITH_FUSED;
  int cull = 0;                                     // warp-level culling in the pair kernel (haccsr_set_culling)
  haccsr::DevBuf<float> law_table;                  // SR_INTERP: f[ntab] then r2[ntab]

  // build scratch
  haccsr::DevBuf<float4> recA, recB, src4;          // ping-pong (x,y,z,m) records, final tree-order array
  haccsr::DevBuf<unsigned> idxA, idxB, perm;        // original index travelling with the record; final perm
  haccsr::DevBuf<int> nidA, nidB;                   // node owning each position at the current level (-1 = final)
  haccsr::DevBuf<haccsr::Node> nodes;
  haccsr::DevBuf<haccsr::NodeAcc> acc;
  haccsr::DevBuf<float4> pp12;                      // TDPTS = 12: the 12 pseudo-particles (x,y,z,m) of every node
  int tdpts = 1;                                    // pseudo-particles per accepted node of the last build (1 or 12)
  haccsr::DevBuf<haccsr::NodeTot> tot;              // exact centroid sums of every finished node
  haccsr::DevBuf<unsigned long long> tile_desc, node_desc;   // look-back descriptors: per tile (split pass), per block (node kernel)
  haccsr::DevBuf<unsigned> tilecount, tilebase;     // per tile: counts and their exclusive scan (refresh.cu)
  haccsr::DevBuf<unsigned> scratch_u32;             // maxima for the fixed-point scales etc.
  haccsr::DevBuf<unsigned> scan_tmp;                // scan_exclusive on long inputs: per-block sums and their scan
  int pass_tpb = 0, pass_occ = 0;                   // split pass: worker threads per block, resident blocks per SM (set on first use)
  haccsr::BuildState *h_state = nullptr;            // pinned
  haccsr::BuildState *d_state = nullptr;
  int64_t *h_counters = nullptr;                    // pinned, misc read-backs
  unsigned long long *d_counters = nullptr;

  // tree of the last kick
  int64_t n_tree = 0;       // particles in the tree
  bool unit_mass = false;   // all masses of the last build are exactly 1.0f
  int n_nodes = 0, n_levels = 0;
  int level_begin[128], level_end[128];

  // walk output
  haccsr::DevBuf<unsigned> n_ranges, n_pseudo, range_off, pseudo_off, list_len;   // per node
  haccsr::DevBuf<uint2> ranges;
  haccsr::DevBuf<float4> pool;
  int64_t tot_ranges = 0, tot_pseudo = 0;

  // force work items
  haccsr::DevBuf<unsigned> item_cnt, item_off;
  haccsr::DevBuf<haccsr::WorkItem> items, items_sorted;   // in node order; sorted by decreasing work
  haccsr::DevBuf<unsigned> lpt_hist;
  int64_t n_items = 0;
  int item_policy = 1;   // force.cu leaf_cut: 1 = remainder items, 0 = padded groups only (env HACCSR_ITEM_POLICY, for A/B measurements)

  // PM coupling (cic.cu): fixed-point deposit accumulators and a staging copy of the grid
  haccsr::DevBuf<unsigned long long> cic_acc;
  haccsr::DevBuf<float> cic_grid;

  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // haccsr_kick_host: second stream + events so that the copies of the arrays the tree build does not read
  // (H2D) and of the arrays the force kernel does not write (D2H) hide behind the kernels
  // overload refresh (refresh.cu): state kept between haccsr_refresh_begin and haccsr_refresh_pack
  int64_t refresh_m = 0;                     // candidates (alive particles within the overload width of a face)
  int refresh_ntiles = 0;
  float refresh_alo[3] = {0, 0, 0}, refresh_ahi[3] = {0, 0, 0}, refresh_ol = 0.f;
  int refresh_slot_of_dir[27] = {0};
  int64_t refresh_count[27] = {0};
  haccsr::DevBuf<int> refresh_slots;
  haccsr::DevBuf<unsigned> refresh_cand;     // candidate indices (own buffer: a kick between begin and pack must not disturb it)
  long long *d_slotcount = nullptr;          // 32 message sizes (device)
  haccsr::DevBuf<unsigned char> xchg_send, xchg_recv;   // haccsr_refresh: packed messages out / in
  haccsr::DevBuf<long long> xchg_table;      // haccsr_refresh: gathered count table, append table
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_up2 = nullptr, ev_vready = nullptr, ev_built = nullptr, ev_main = nullptr;   // ev_up2: all uploads done; ev_vready: velocities in tree order
  bool defer_kick = false;            // haccsr_kick_host: the force kernel stores accelerations, apply_kick() kicks (force.cu)
  haccsr::DevBuf<float> kick_a[3];    // those accelerations
  bool wait_up2 = false;    // build_tree must wait for ev_up2 before it permutes the payload arrays
  // haccsr_kick_host: the force kernel runs as `force_groups` launches by particle range and the velocities of each
  // range go to the host arrays ho_v on the copy stream as soon as they are final (force.cu)
  const void *pending_ho = nullptr;   // host output arrays whose device->host copies are not queued yet (api.cu)
  int64_t pending_count = 0;
  int force_groups = 1;
  int64_t group_lo[9] = {0};          // particle ranges of the groups: [group_lo[g], group_lo[g+1])
  int host_groups = 2;                // haccsr_kick_host: launches by particle range; the last one covers host_last_frac of the particles
  float host_last_frac = 0.10f;       //   (0 = equal ranges).  Env HACCSR_HOST_GROUPS / HACCSR_HOST_LAST, for A/B measurements
  int64_t seg_off[17] = {0};          // item ranges of the launches: (group g, chunk items) = [2g, 2g+1), remainder items [2g+1, 2g+2)
  float *ho_v[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_grp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int launches = 0, force_launches = 0;
};

namespace haccsr {
// tree_build.cu
int build_tree(haccsr_ctx *c, int64_t n, const float lo[3], const float hi[3], int64_t ppn, int tdpts);
// walk.cu
int build_lists(haccsr_ctx *c, const float flo[3], const float fhi[3], float theta, haccsr_stats *st);
// force.cu
int run_force(haccsr_ctx *c, float fcoeff, bool count_in_cutoff, haccsr_stats *st);
// tree_build.cu: phi / id / mask into tree order by the completed permutation (second part of the split gather)
int gather_payload(haccsr_ctx *c, cudaStream_t st);
// force.cu: v = fma(fcoeff * m, a, v) for particles [lo, hi) from the deferred accelerations
int apply_kick(haccsr_ctx *c, float fcoeff, int64_t lo, int64_t hi);
// api.cu: queue the device->host copies of the seven arrays the force kernel does not write (haccsr_kick_host)
int issue_host_out(haccsr_ctx *c);
// api.cu: stable two-way partition of the ten arrays by a 0/1 flag
int compact_by_flags(haccsr_ctx *c, const unsigned *flag, unsigned *pref, int64_t n, int64_t *n_kept);
// refresh.cu: write the 26 messages into a device buffer, stream-ordered (no synchronisation)
int refresh_pack_async(haccsr_ctx *c, const int64_t byte_off_by_slot[27], void *sendbuf_device);
// scan utility (tree_build.cu): exclusive scan of n unsigned values; total written to *d_total (device).
int scan_exclusive(haccsr_ctx *c, const unsigned *in, unsigned *out, int64_t n, unsigned long long *d_total);
}  // namespace haccsr
