// tree_build.cu -- device build of the RCB force tree (level-synchronous), sm_100a.
//
// What it computes is the tree of the reference's createRCBForceTree
// (reference src/halo_finder/RCBForceTree.cxx:778-918): every node gets the TIGHT bounding box and the
// mass-weighted centroid of its particles (cm, src/halo_finder/BGQCM.c:181-212); a node with more than
// ppn particles is cut on the longest edge of that box (:838-852) at the centroid coordinate (:720),
// particles with key < pivot going left (:640); a split that leaves one side empty keeps the node as an
// oversized leaf with two orphan children (:727-729).
//
// How it is computed is not the reference's recursion.  All nodes of one level are processed together, one streaming
// pass over the particles per level (HBM-bound), with no host round trip between levels:
//   k_cm_tile      (root only) box / centroid sums
//   k_level_nodes  per node of the level: box, centroid, split decision; breadth-first child allocation (block scan +
//                  look-back over the blocks); the range of the next level goes to BuildState on the device
//   k_split_pass   per particle: left flag, rank among the node's left / right particles (segmented tile scan + decoupled
//                  look-back over the tiles), move to the other record buffer, and the box / centroid sums of the child it
//                  lands in; particles of finished leaves are written once to the final tree-order arrays
// The centroid sums are accumulated as wide fixed-point integers (two 64-bit words, add_split), so they are exact and independent
// of summation order: the build is deterministic, and the float centroid equals the reference's
// (double-accumulated) one except where the reference's own rounding error straddles a float boundary.
// Children are numbered breadth-first (parent index < child index, as the reference guarantees at :808-809).
// Left blocks keep input order like the reference; right blocks are kept in input order too (written back to front, un-mirrored
// through a per-node direction: k_split_pass) where the reference's swap loop scrambles them -- that order only affects the FP32
// summation order inside a leaf.  The level where nothing splits any more is finalised by k_gather.
#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace haccsr {

static constexpr int TILE = 1024;      // particles per tile / thread block
static constexpr int TPB = 256;        // threads per block in tile kernels
static constexpr int IPT = TILE / TPB; // items per thread in the split pass (striped: particle = tile base + j * threads + t)
static constexpr int SMAX = 32;        // node runs per tile accumulated in shared memory

// ---- helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_f(float f) {   // order-preserving float -> uint
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// Exact sums wider than 64 bits without a carry chain: a contribution v (a per-block int64 sum) is split as
// v = (v >> 32) * 2^32 + (v & 0xffffffff) and the two parts are added to two 64-bit words with result-less atomics (RED):
// the block does not wait for the returned value that a lo/hi carry would need.  total = hi * 2^32 + lo (hi signed).
// Headroom: |v| < 2^62 per block, so hi grows by < 2^30 and lo by < 2^32 per block: exact for > 2^31 blocks.
__device__ __forceinline__ void add_split(unsigned long long *lo, unsigned long long *hi, long long v) {
  if (v == 0) return;
  atomicAdd(lo, (unsigned long long)(v & 0xffffffffll));
  atomicAdd(hi, (unsigned long long)(v >> 32));
}
__device__ __forceinline__ double split_to_double(unsigned long long lo, unsigned long long hi) {
  // hi * 2^32 + lo as a 128-bit two's-complement integer, then -> double via the magnitude (avoids cancellation for
  // small negative sums)
  const __int128 t = ((__int128)(long long)hi << 32) + (__int128)lo;
  const bool neg = t < 0;
  const unsigned __int128 m = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  const double d = (double)(unsigned long long)(m >> 64) * 18446744073709551616.0 + (double)(unsigned long long)m;
  return neg ? -d : d;
}
__device__ __forceinline__ float comp(const float4 &r, int d) { return d == 0 ? r.x : (d == 1 ? r.y : r.z); }

// block-wide exclusive scan of one int per thread (TPB or 1024 threads); returns exclusive prefix,
// total in *total.  s_w must hold >= 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int *s_w, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();   // protect s_w from the previous use
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = (lane < nw) ? s_w[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += y;
    }
    s_w[lane] = xi - x;          // exclusive warp offsets
    if (lane == 31) s_w[32] = xi;
  }
  __syncthreads();
  *total = s_w[32];
  return s_w[w] + inc - v;
}

// ---- records + maxima ------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_init_records(const float *__restrict__ x, const float *__restrict__ y,
                                                      const float *__restrict__ z, const float *__restrict__ m,
                                                      int n, float4 *__restrict__ rec, unsigned *__restrict__ idx,
                                                      int *__restrict__ nid, unsigned *__restrict__ maxima) {
  float mc = 0.f, mm = 0.f;
  bool notunit = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 r = make_float4(x[i], y[i], z[i], m[i]);
    rec[i] = r; idx[i] = (unsigned)i; nid[i] = 0;
    notunit = notunit || (r.w != 1.0f);
    mc = fmaxf(mc, fmaxf(fabsf(r.x), fmaxf(fabsf(r.y), fabsf(r.z))));
    mm = fmaxf(mm, fabsf(r.w));
  }
  unsigned uc = __reduce_max_sync(0xffffffffu, __float_as_uint(mc));
  unsigned um = __reduce_max_sync(0xffffffffu, __float_as_uint(mm));
  if ((threadIdx.x & 31) == 0) { atomicMax(&maxima[0], uc); atomicMax(&maxima[1], um); }
  if (__any_sync(0xffffffffu, notunit) && (threadIdx.x & 31) == 0) atomicOr(&maxima[2], 1u);
}

// scales[0] = 2^kx applied to float products w*x, scales[1] = 2^km applied to w (both exact powers of two)
__global__ void k_root_init(Node *nodes, NodeAcc *acc, int n, float3 lo, float3 hi, const unsigned *maxima,
                            float *scales, BuildState *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Node r;
  r.count = n; r.offset = 0; r.cl = 0; r.cr = 0;
  r.xmin[0] = lo.x; r.xmin[1] = lo.y; r.xmin[2] = lo.z;
  r.xmax[0] = hi.x; r.xmax[1] = hi.y; r.xmax[2] = hi.z;
  r.xc[0] = r.xc[1] = r.xc[2] = 0.f; r.ppm = 0.f; r.parent = -1; r.split = -1;
  nodes[0] = r;
  NodeAcc a;
  for (int k = 0; k < 3; ++k) { a.umin[k] = 0xffffffffu; a.umax[k] = 0u; }
  for (int k = 0; k < 4; ++k) { a.lo[k] = 0; a.hi[k] = 0; }
  acc[0] = a;
  float mc = __uint_as_float(maxima[0]), mm = __uint_as_float(maxima[1]);
  int ec = 0, em = 0;
  if (mc > 0.f) frexpf(mc, &ec);
  if (mm > 0.f) frexpf(mm, &em);
  int kx = 50 - (ec + em), km = 50 - em;
  kx = max(-100, min(100, kx)); km = max(-100, min(100, km));
  scales[0] = ldexpf(1.0f, kx); scales[1] = ldexpf(1.0f, km);
  scales[2] = (float)(km - kx);   // exponent to undo: xc = (Sx / Sw) * 2^(km-kx)
  st->error = 0; st->tile_ticket = 0u; st->node_ticket = 0u;
  st->unit_mass = (n > 0 && maxima[2] == 0u) ? 1 : 0;
  for (int L = 0; L < 128; ++L) { st->nsplit[L] = -1; st->lvl_begin[L] = 0; st->lvl_end[L] = 0; }
  st->lvl_begin[0] = 0; st->lvl_end[0] = 1; st->lvl_begin[128] = 0; st->lvl_end[128] = 0;
}

// ---- per-node partial sums ---------------------------------------------------------------------------
struct Part {
  unsigned umin[3], umax[3];
  long long s[4];
};
__device__ __forceinline__ void part_reset(Part &p) {
  p.umin[0] = p.umin[1] = p.umin[2] = 0xffffffffu; p.umax[0] = p.umax[1] = p.umax[2] = 0u;
  p.s[0] = p.s[1] = p.s[2] = p.s[3] = 0;
}
// the product w*x is formed in float exactly as in BGQCM.c:203-206, then summed exactly (fixed point)
__device__ __forceinline__ void part_add_sums(long long (&s)[4], const float4 &r, float sx, float sm) {
  s[0] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.x), sx));
  s[1] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.y), sx));
  s[2] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.z), sx));
  s[3] += __float2ll_rn(__fmul_rn(r.w, sm));
}
__device__ __forceinline__ void part_add(Part &p, const float4 &r, float sx, float sm) {
  unsigned ex = enc_f(r.x), ey = enc_f(r.y), ez = enc_f(r.z);
  p.umin[0] = min(p.umin[0], ex); p.umax[0] = max(p.umax[0], ex);
  p.umin[1] = min(p.umin[1], ey); p.umax[1] = max(p.umax[1], ey);
  p.umin[2] = min(p.umin[2], ez); p.umax[2] = max(p.umax[2], ez);
  part_add_sums(p.s, r, sx, sm);
}

struct Slot {
  unsigned umin[3], umax[3];
  unsigned used, pad;
  unsigned long long s[4];
};
__device__ __forceinline__ void slot_reset(Slot &S) {
  S.umin[0] = S.umin[1] = S.umin[2] = 0xffffffffu; S.umax[0] = S.umax[1] = S.umax[2] = 0u;
  S.used = 0; S.s[0] = S.s[1] = S.s[2] = S.s[3] = 0;
}

__device__ __forceinline__ void flush_part(const Part &p, int nd, int n0, Slot *slots, NodeAcc *acc) {
  int sl = nd - n0;
  if (sl >= 0 && sl < SMAX) {
    Slot &S = slots[sl];
    for (int k = 0; k < 3; ++k) { atomicMin(&S.umin[k], p.umin[k]); atomicMax(&S.umax[k], p.umax[k]); }
    for (int k = 0; k < 4; ++k) atomicAdd(&S.s[k], (unsigned long long)p.s[k]);
    S.used = 1;
  } else {
    NodeAcc &A = acc[nd];
    for (int k = 0; k < 3; ++k) { atomicMin(&A.umin[k], p.umin[k]); atomicMax(&A.umax[k], p.umax[k]); }
    for (int k = 0; k < 4; ++k) add_split(&A.lo[k], &A.hi[k], p.s[k]);
  }
}

// Box and centroid sums of the ROOT (every deeper node gets them from the split pass that creates it).
// A warp owns CM_IPT rows of 32 consecutive particles (coalesced 128-byte node-id and 512-byte record loads, CM_IPT of
// each in flight per lane); every lane sums its particles, the 32 partials are combined by a plain warp reduction --
// REDUX for the six box bounds, a shuffle tree for the four 64-bit sums -- and lane 0 adds the result to the block's
// shared-memory slot, flushed to the node's global accumulator once per block.
static constexpr int CM_IPT = 8;
static constexpr int CM_TILE = TPB * CM_IPT;

__device__ __forceinline__ long long shfl_xor_ll(long long v, int o) {
  int lo = __shfl_xor_sync(0xffffffffu, (int)(unsigned)(v & 0xffffffffll), o);
  int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), o);
  return ((long long)hi << 32) | (long long)(unsigned)lo;
}

__global__ void __launch_bounds__(TPB) k_cm_tile(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                 NodeAcc *__restrict__ acc, const float *__restrict__ scales) {
  __shared__ int s_n0;
  __shared__ Slot slots[SMAX];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int wbase = blockIdx.x * CM_TILE + w * (32 * CM_IPT) + lane;
  int nd[CM_IPT];
  float4 r[CM_IPT];
#pragma unroll
  for (int k = 0; k < CM_IPT; ++k) {
    const int i = wbase + 32 * k;
    nd[k] = (i < n) ? __ldcs(nid + i) : -1;
  }
#pragma unroll
  for (int k = 0; k < CM_IPT; ++k) if (nd[k] >= 0) r[k] = __ldcs(rec + wbase + 32 * k);
  if (t == 0) s_n0 = INT_MAX;
  if (t < SMAX) slot_reset(slots[t]);
  __syncthreads();
  int mn = INT_MAX;
#pragma unroll
  for (int k = 0; k < CM_IPT; ++k) if (nd[k] >= 0) mn = min(mn, nd[k]);
  mn = __reduce_min_sync(0xffffffffu, mn);       // the warp's first node
  if (lane == 0 && mn != INT_MAX) atomicMin(&s_n0, mn);
  __syncthreads();
  const int n0 = s_n0;
  if (n0 == INT_MAX) return;   // no active particle in this tile
  const float sx = scales[0], sm = scales[1];
  int key = mn;
  while (key != INT_MAX) {     // warp-uniform: one pass per distinct node among the warp's particles
    Part p; part_reset(p);
    int next = INT_MAX;
#pragma unroll
    for (int k = 0; k < CM_IPT; ++k) {
      if (nd[k] == key) part_add(p, r[k], sx, sm);
      else if (nd[k] > key) next = min(next, nd[k]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) { p.umin[q] = __reduce_min_sync(0xffffffffu, p.umin[q]); p.umax[q] = __reduce_max_sync(0xffffffffu, p.umax[q]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) p.s[q] += shfl_xor_ll(p.s[q], o);
    }
    if (lane == 0) flush_part(p, key, n0, slots, acc);
    key = __reduce_min_sync(0xffffffffu, next);
  }
  __syncthreads();
  if (t < SMAX && slots[t].used) {
    NodeAcc &A = acc[n0 + t];
    const Slot &S = slots[t];
    for (int k = 0; k < 3; ++k) { atomicMin(&A.umin[k], S.umin[k]); atomicMax(&A.umax[k], S.umax[k]); }
    for (int k = 0; k < 4; ++k) add_split(&A.lo[k], &A.hi[k], (long long)S.s[k]);
  }
}

// ---- decoupled look-back (tile descriptors of the split pass, block descriptors of the node kernel) ---------------------
// One 64-bit word per tile: epoch << 34 | status << 32 | value.  The epoch is level + 1, so a word left by an earlier level
// (or the zero the build starts from) reads as "not published yet"; the value travels in the same word as the status, so a
// plain 64-bit store / load pair needs no fence.
static constexpr unsigned long long ST_AGG = 1ull, ST_PREFIX = 2ull;
__device__ __forceinline__ unsigned long long desc_make(unsigned epoch, unsigned long long status, unsigned v) {
  return ((unsigned long long)epoch << 34) | (status << 32) | (unsigned long long)v;
}
__device__ __forceinline__ void desc_store(unsigned long long *p, unsigned long long w) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long desc_load(const unsigned long long *p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// Warp-collective: the sum of the values of descriptors pos, pos-1, ... down to and including the nearest inclusive prefix.
// W windows of 32 descriptors are read at once (one L2 round trip for a look-back of 32*W tiles); a word that is not
// published yet makes the warp read again from there.  Terminates because every tile with a smaller ticket is held by a
// running block that publishes its aggregate before it waits for anything (tickets are handed out in increasing order).
template <int W>
__device__ __forceinline__ unsigned lookback(const unsigned long long *desc, int pos, unsigned epoch, int lane) {
  unsigned sum = 0;
  while (pos >= 0) {
    unsigned long long d[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
      const int j = pos - (q * 32 + lane);
      d[q] = j >= 0 ? desc_load(desc + j) : desc_make(epoch, ST_PREFIX, 0u);
    }
    bool done = false;
    int consumed = 0;
#pragma unroll
    for (int q = 0; q < W; ++q) {
      if (!done && consumed == q) {        // warp-uniform
        const bool valid = (unsigned)(d[q] >> 34) == epoch;
        const bool isp = valid && ((d[q] >> 32) & 3ull) == ST_PREFIX;
        const unsigned bv = __ballot_sync(0xffffffffu, valid), bp = __ballot_sync(0xffffffffu, isp);
        const unsigned lp = bp ? (unsigned)(__ffs(bp) - 1) : 31u;                 // nearest prefix of this window
        const unsigned need = lp == 31u ? 0xffffffffu : ((2u << lp) - 1u);        // lanes 0 .. lp
        if ((bv & need) == need) {
          sum += __reduce_add_sync(0xffffffffu, ((need >> lane) & 1u) ? (unsigned)(d[q] & 0xffffffffull) : 0u);
          consumed = q + 1;
          done = bp != 0u;
        }
      }
    }
    if (done) return sum;
    pos -= 32 * consumed;
  }
  return sum;
}

__device__ __forceinline__ __int128 split_to_i128(unsigned long long lo, unsigned long long hi) {
  return ((__int128)(long long)hi << 32) + (__int128)lo;
}
__device__ __forceinline__ double i128_to_double(__int128 t) {
  // via the magnitude (avoids cancellation for small negative sums)
  const bool neg = t < 0;
  const unsigned __int128 m = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  const double d = (double)(unsigned long long)(m >> 64) * 18446744073709551616.0 + (double)(unsigned long long)m;
  return neg ? -d : d;
}

// ---- node kernel: finalize the nodes of one level, decide splits, allocate children breadth-first --------------------------
// One launch per level, no host round trip: the node range of the level was written by the previous level's launch into
// BuildState, blocks take tickets (the host only knows an upper bound of the node count), and the rank of a split node among
// the level's split nodes comes from a block scan plus a look-back over the blocks' counts, so children sit at
// next_base + 2 * rank: the numbering is breadth-first and deterministic (parent index < child index, as the reference
// guarantees at :808-809).
// Sums: the split pass accumulates the centroid sums of LEFT children only; a right child's sums are its parent's minus its
// sibling's -- exact, because they are integers.
static constexpr int NTPB = 256;
__global__ void __launch_bounds__(NTPB) k_level_nodes(Node *__restrict__ nodes, NodeAcc *__restrict__ acc,
                                                      NodeTot *__restrict__ tot, BuildState *__restrict__ st,
                                                      unsigned long long *__restrict__ ndesc, int level, int ppn,
                                                      int max_nodes, const float *__restrict__ scales) {
  __shared__ int s_w[34];
  __shared__ int s_tk;
  __shared__ unsigned s_base;
  if (st->error || (level > 0 && st->nsplit[level - 1] <= 0)) return;     // the tree was finished by an earlier pass
  const int begin = st->lvl_begin[level], end = st->lvl_end[level];
  const int nblk = (end - begin + NTPB - 1) / NTPB;
  if (threadIdx.x == 0) s_tk = (int)atomicAdd(&st->node_ticket, 1u);
  __syncthreads();
  const int tk = s_tk;
  if (tk >= nblk) return;
  const int k = begin + tk * NTPB + (int)threadIdx.x;
  int split = 0;
  if (k < end) {
    Node nd = nodes[k];
    int d = -1;
    const int dir = nd.split == -2;        // the node's particles lie in reverse order (see k_split_pass)
    if (level > 0) {
      // a split that left one side empty (RCBForceTree.cxx:727-729): the parent stays an (oversized) leaf whose monopole is
      // the sum over its empty children, i.e. zero (:856-889), and both children stay empty orphans.  The pass has labelled
      // the parent's particles with one of the orphans; this level's pass writes them to their final place.
      const int kl = begin + ((k - begin) & ~1);
      const int c0 = nodes[kl].count, c1 = nodes[kl + 1].count;
      if (c0 == 0 || c1 == 0) {
        nd.count = 0;
        if (k == kl) { Node *p = nodes + nd.parent; p->cl = 0; p->cr = 0; p->ppm = 0.f; }
      }
    }
    if (nd.count > 0) {
      const NodeAcc a = acc[k];
      __int128 T[4];
      if (level == 0 || ((k - begin) & 1) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) T[q] = split_to_i128(a.lo[q], a.hi[q]);
      } else {
        const NodeTot pt = tot[nd.parent];
        const NodeAcc sib = acc[k - 1];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          T[q] = (((__int128)(long long)pt.hi[q] << 64) | (__int128)pt.lo[q]) - split_to_i128(sib.lo[q], sib.hi[q]);
      }
      NodeTot o;
#pragma unroll
      for (int q = 0; q < 4; ++q) { o.lo[q] = (unsigned long long)T[q]; o.hi[q] = (unsigned long long)(T[q] >> 64); }
      tot[k] = o;
      const double undo = ldexp(1.0, (int)scales[2]);
      for (int q = 0; q < 3; ++q) { nd.xmin[q] = dec_f(a.umin[q]); nd.xmax[q] = dec_f(a.umax[q]); }
      const double sw = i128_to_double(T[3]);
      for (int q = 0; q < 3; ++q) nd.xc[q] = (float)((i128_to_double(T[q]) / sw) * undo);     // BGQCM.c:209-211
      if (nd.count > ppn) {                                          // RCBForceTree.cxx:788
        const float l0 = __fsub_rn(nd.xmax[0], nd.xmin[0]), l1 = __fsub_rn(nd.xmax[1], nd.xmin[1]),
                    l2 = __fsub_rn(nd.xmax[2], nd.xmin[2]);
        d = (l0 > l1 && l0 > l2) ? 0 : ((l1 > l2) ? 1 : 2);          // :844-852
        split = 1;
      } else {
        // leaf monopole: sum of masses (pp<1>, :536-569); unused when count <= 1 (:788-797)
        nd.ppm = (nd.count > 1) ? (float)(sw / (double)scales[1]) : 0.f;
      }
    }
    nd.split = split ? (d | (dir << 2)) : (dir ? -2 : -1);
    nodes[k] = nd;
  }
  int total;
  const int excl = block_excl_scan(split, s_w, &total);
  const unsigned epoch = (unsigned)level + 1u;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane == 0) desc_store(ndesc + tk, desc_make(epoch, tk == 0 ? ST_PREFIX : ST_AGG, (unsigned)total));
    unsigned base = 0;
    if (tk > 0) {
      base = lookback<1>(ndesc, tk - 1, epoch, lane);
      if (lane == 0) desc_store(ndesc + tk, desc_make(epoch, ST_PREFIX, base + (unsigned)total));
    }
    if (lane == 0) s_base = base;
  }
  __syncthreads();
  const int next_base = end;
  if (split) {
    const int cl = next_base + 2 * ((int)s_base + excl);
    if (cl + 1 < max_nodes) {
      Node nd = nodes[k];
      const int d = nd.split & 3, dir = (nd.split >> 2) & 1;
      nodes[k].cl = cl; nodes[k].cr = cl + 1;     // withdrawn by the next level's launch on a degenerate split
      Node c;
      c.count = 0; c.offset = 0; c.cl = 0; c.cr = 0; c.ppm = 0.f; c.parent = k; c.split = -1;
      for (int q = 0; q < 3; ++q) { c.xmin[q] = nd.xmin[q]; c.xmax[q] = nd.xmax[q]; c.xc[q] = 0.f; }
      Node l = c, r = c;
      l.xmax[d] = nd.xc[d]; r.xmin[d] = nd.xc[d];                  // :747,763
      l.split = dir ? -2 : -1; r.split = dir ? -1 : -2;            // a left child keeps its parent's direction, a right child reverses it
      nodes[cl] = l; nodes[cl + 1] = r;
      NodeAcc z;
      for (int q = 0; q < 3; ++q) { z.umin[q] = 0xffffffffu; z.umax[q] = 0u; }
      for (int q = 0; q < 4; ++q) { z.lo[q] = 0; z.hi[q] = 0; }
      acc[cl] = z; acc[cl + 1] = z;
    }
  }
  if (tk == nblk - 1 && threadIdx.x == 0) {
    const int ns = (int)s_base + total;
    const bool overflow = next_base + 2 * ns >= max_nodes;      // the last child index must stay below max_nodes
    st->lvl_begin[level + 1] = next_base;
    st->lvl_end[level + 1] = next_base + (overflow ? 0 : 2 * ns);
    st->nsplit[level] = overflow ? 0 : ns;
    if (overflow) st->error = 1;
    st->tile_ticket = 0u;
  }
}

// ---- the split pass: one streaming pass over the particles per tree level -----------------------------------------------------
// Reads every record once and writes it once.  For a particle of a node that splits at this level: left flag from the pivot
// (key < centroid coordinate, :640), its rank among the node's left (right) particles before it, destination in the other
// record buffer, and its contribution to the box and centroid sums of the child it moves to, so the children need no pass
// of their own.  Particles of nodes that became leaves are written to their final place in src4 / perm.
// Left particles fill the node's range from the front, right particles from the back, so no particle needs the node's total
// left count (which only the node's last tile knows).  That writes the right block in reverse; a node therefore carries a
// direction (Node::split bit 2, or -2 for a leaf): scanning a reversed node front to back meets its particles last to
// first, so its left block comes out reversed and its right block in order -- a child's direction is its parent's, flipped
// for right children -- and a reversed leaf is mirrored when it is written to its final place.  The result is the STABLE
// partition on both sides: the order inside every leaf is a fixed function of the input order, and building the tree
// again from its own output leaves every particle where it is (like the reference's swap loop, :648-669).
//
// The rank needs the number of left particles of the node in all earlier tiles.  Tiles are taken in ticket order; each
// publishes the left count of its LAST node segment (the lefts behind its last node start) as an aggregate, or as an
// inclusive prefix when that segment starts inside the tile, and a tile whose first node started earlier sums its
// predecessors' words back to the nearest prefix (decoupled look-back, one L2 round trip for 32*LB_W tiles).  Inside the
// tile a segmented scan over the 32 per-(row, warp) ballots gives every particle the lefts since the last node start.
static constexpr int NSLOT = 32;       // children of one tile accumulated in shared memory (more: straight to global)
// Per-child accumulators of one tile in shared memory.  The four centroid sums are kept as three 21-bit limbs each in 32-bit
// words (native ATOMS.ADD; a 64-bit shared atomicAdd compiles to a compare-and-swap loop): a warp contributes less than 2^26 per
// limb, a tile's eight warps less than 2^29.
struct PassSlot {
  unsigned umin[3], umax[3];
  unsigned used, pad;
  int limb[4][3];
};
__device__ __forceinline__ void pslot_reset(PassSlot &S) {
  S.umin[0] = S.umin[1] = S.umin[2] = 0xffffffffu; S.umax[0] = S.umax[1] = S.umax[2] = 0u;
  S.used = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) { S.limb[q][0] = 0; S.limb[q][1] = 0; S.limb[q][2] = 0; }
}
#ifndef HSR_LBW
#define HSR_LBW 1
#endif
static constexpr int LB_W = HSR_LBW;   // look-back windows read at once
static constexpr int NST = 4;          // tiles of the shared-memory ring: B stage, A stage, two in flight

__device__ __forceinline__ unsigned tb_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tb_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tb_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tb_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tb_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// exact warp sum of per-lane 64-bit partials |v| < 2^57 as three REDUX.SUM over 21-bit limbs
__device__ __forceinline__ long long warp_sum_ll(long long v) {
  const int l0 = (int)(v & 0x1fffffll), l1 = (int)((v >> 21) & 0x1fffffll), l2 = (int)(v >> 42);
  const int s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1), s2 = __reduce_add_sync(0xffffffffu, l2);
  return (long long)s0 + ((long long)s1 << 21) + ((long long)s2 << 42);
}

// Structure of the kernel.  Tiles go round the blocks (block b owns tiles b, b+G, b+2G, ...: in iteration k the grid works
// on G consecutive tiles) and travel through a ring of NST shared-memory slots filled by 1-D TMA bulk copies (records, indices
// and node ids of a tile are three contiguous ranges; one thread posts them three tiles ahead, nobody stages anything in
// registers).  A tile is visited twice, one iteration apart:
//   A stage (tile k+1): left flags and node starts -> ballots, the segmented scan, and the tile's look-back word is PUBLISHED;
//   B stage (tile k)  : look-back (every word it meets was published an iteration ago, so it reads, it does not wait),
//                       destinations, stores, children's sums.
// Publishing a tile's count in the same iteration that consumes its predecessors' counts makes all blocks march in step
// and leaves the warps waiting at the barrier behind the look-back (measured: 54 % of all stall samples, 0.7 ms per
// level); with the count one iteration ahead a block can run an iteration ahead of its neighbours.
template <int PTPB>
__global__ void __launch_bounds__(PTPB + 32, PTPB == 256 ? 2 : 3) k_split_pass(const float4 *__restrict__ rec, const unsigned *__restrict__ idx,
                                                       const int *__restrict__ nid, Node *__restrict__ nodes,
                                                       NodeAcc *__restrict__ acc, const float *__restrict__ scales,
                                                       BuildState *__restrict__ st, unsigned long long *__restrict__ desc,
                                                       int level, int n, int ntiles, float4 *__restrict__ rec_out,
                                                       unsigned *__restrict__ idx_out, int *__restrict__ nid_out,
                                                       float4 *__restrict__ src4, unsigned *__restrict__ perm) {
  constexpr int PT = PTPB * IPT;                       // particles per tile
  constexpr int NE = IPT * (PTPB / 32);                // (row, warp) entries of a tile
  constexpr int SLOT_REC = 0, SLOT_IDX = PT * 16, SLOT_NID = PT * 20, SLOT_BYTES = PT * 24;
  static_assert(NE <= 32, "one warp scans the per-(row, warp) entries");
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned s_fb[NST][32], s_hb[NST][32];   // per slot and (row, warp) entry: ballots of the left flags and of the node starts
  __shared__ int s_cv[NST][32], s_cf[NST][32];        // per slot: exclusive segmented scan of the entries
  __shared__ int s_ev[32], s_ef[32];                  // A stage: lefts behind the entry's last node start (or all of them), has a node start
  __shared__ int s_misc[NST][4];   // per slot: [0] first node started in an earlier tile, [1] smallest child id, [2] carry, [3] own word
  __shared__ unsigned long long s_bar[NST];
  __shared__ PassSlot slots[NSLOT];
  if (st->error || (level > 0 && st->nsplit[level - 1] <= 0)) return;     // the tree was finished by an earlier pass
  if (st->nsplit[level] == 0) return;      // no node splits at this level: its leaves are finalised by k_gather, from these buffers
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned epoch = (unsigned)level + 1u;
  const unsigned below = (1u << lane) - 1u;
  const float sx = scales[0], sm = scales[1];
  const int G = gridDim.x, T00 = blockIdx.x;
  if (t == 0) {
    if (blockIdx.x == 0) st->node_ticket = 0u;
    for (int q = 0; q < NST; ++q) { tb_mbar_init(tb_smem_u32(&s_bar[q]), 1); s_misc[q][0] = 0; s_misc[q][1] = INT_MAX; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < NSLOT) pslot_reset(slots[t]);
  if (t < 32) { s_ev[t] = 0; s_ef[t] = 0; }           // entries beyond NE stay neutral in the scan
  __syncthreads();
  // one thread posts the three bulk copies of tile number k of this block
  auto issue = [&](int k) {
    const int T = T00 + k * G;
    if (T >= ntiles) return;
    const int q = k % NST, i0 = T * PT;
    const int cnt = min(PT, n - i0);
    const unsigned b4 = (unsigned)(((cnt * 4) + 15) & ~15);
    const unsigned bar = tb_smem_u32(&s_bar[q]), base = tb_smem_u32(ring + (size_t)q * SLOT_BYTES);
    s_misc[q][0] = 0; s_misc[q][1] = INT_MAX;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tb_mbar_expect_tx(bar, (unsigned)cnt * 16u + 2u * b4);
    tb_bulk_g2s(base + SLOT_REC, rec + i0, (unsigned)cnt * 16u, bar);
    tb_bulk_g2s(base + SLOT_IDX, idx + i0, b4, bar);
    tb_bulk_g2s(base + SLOT_NID, nid + i0, b4, bar);
  };
  // A stage of tile number k: flags, node starts, ballots, entries (consumed by warp 0 behind the next barrier)
  auto stage_a = [&](int k) {
    const int T = T00 + k * G, q = k % NST, tile0 = T * PT;
    tb_mbar_wait(tb_smem_u32(&s_bar[q]), (unsigned)((k / NST) & 1));
    const unsigned char *sl = ring + (size_t)q * SLOT_BYTES;
    const int *nidS = reinterpret_cast<const int *>(sl + SLOT_NID);
    const float *recS = reinterpret_cast<const float *>(sl + SLOT_REC);
    int cmin = INT_MAX;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const int p = j * PTPB + t, i = tile0 + p;
      const int ndv = i < n ? nidS[p] : -1;
      // unconditional loads (node 0 for a finished particle): the twelve node loads of a thread's four items go out together
      const float4 *np = reinterpret_cast<const float4 *>(nodes + max(ndv, 0));
      const float4 a = __ldg(np), c = __ldg(np + 2), d = __ldg(np + 3);
      const int cl = ndv >= 0 ? __float_as_int(a.z) : 0;
      const bool split = cl > 0;                                                 // the node splits at this level
      const int sp = __float_as_int(d.w) & 3, off = __float_as_int(a.y);
      const float pivot = sp == 0 ? c.z : (sp == 1 ? c.w : d.x);                 // xc[sp], RCBForceTree.cxx:720
      const int flag = split && recS[4 * p + sp] < pivot;                        // :640
      const int head = split && i == off;
      if (split) cmin = min(cmin, cl);
      if (split && p == 0 && off < tile0) s_misc[q][0] = 1;
      const unsigned fb = __ballot_sync(0xffffffffu, flag), hb = __ballot_sync(0xffffffffu, head);
      if (lane == 0) {
        const int e = j * (PTPB / 32) + w;
        s_fb[q][e] = fb; s_hb[q][e] = hb;
        s_ev[e] = hb ? __popc(fb >> (31 - __clz(hb))) : __popc(fb);
        s_ef[e] = hb != 0u;
      }
    }
    cmin = __reduce_min_sync(0xffffffffu, cmin);
    if (lane == 0 && cmin != INT_MAX) atomicMin(&s_misc[q][1], cmin);
  };
  // warp 0, behind the barrier that follows stage_a: segmented scan of the entries, publish the tile's word
  auto scan_publish = [&](int k) {
    const int T = T00 + k * G, q = k % NST;
    int iv = s_ev[lane], iff = s_ef[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int pv = __shfl_up_sync(0xffffffffu, iv, o), pf = __shfl_up_sync(0xffffffffu, iff, o);
      if (lane >= o) { if (!iff) iv += pv; iff |= pf; }
    }
    int ev = __shfl_up_sync(0xffffffffu, iv, 1), ef = __shfl_up_sync(0xffffffffu, iff, 1);
    if (lane == 0) { ev = 0; ef = 0; }
    s_cv[q][lane] = ev; s_cf[q][lane] = ef;
    if (lane == 31) {
      const int need = s_misc[q][0];
      desc_store(desc + T, desc_make(epoch, (!need || iff) ? ST_PREFIX : ST_AGG, (unsigned)iv));
      s_misc[q][3] = iv | (iff << 31);
    }
  };
  const int nk = T00 < ntiles ? (ntiles - T00 + G - 1) / G : 0;     // tiles of this block
  // The block is PTPB worker threads plus one helper warp.  The helper posts the bulk copies, sums the look-back words of
  // the tile the workers are about to finish while they run the A stage of the next one (the look-back is one or two L2
  // round trips: with warp 0 doing it between the barriers the other warps stood waiting, 30 % of all stall samples), and
  // scans / publishes the A stage's entries behind the barrier.
  const bool helper = t >= PTPB;
  // warp-collective, helper only: the number of left particles of tile k's first node in all earlier tiles, and the tile's
  // inclusive word if it has not published one yet
  auto resolve_carry = [&](int k) {
    const int T = T00 + k * G, q = k % NST;
    unsigned carry = 0;
    if (s_misc[q][0]) {
      carry = lookback<LB_W>(desc, T - 1, epoch, lane);
      const int own = s_misc[q][3];
      if (own >= 0 && lane == 0) desc_store(desc + T, desc_make(epoch, ST_PREFIX, carry + (unsigned)own));
    }
    if (lane == 0) s_misc[q][2] = (int)carry;
  };
  if (helper && lane == 0) for (int k = 0; k < NST - 1; ++k) issue(k);
  if (!helper && nk > 0) stage_a(0);
  __syncthreads();
  if (helper && nk > 0) scan_publish(0);
  __syncthreads();
  for (int k = 0; k < nk; ++k) {
    const int T0 = T00 + k * G, q = k % NST, tile0 = T0 * PT;
    // The helper sums the look-back words of tile k while the workers run the A stage of tile k+1.  The words it needs were
    // published an iteration ago, so it reads, it does not wait -- but it must not take longer than the A stage, or the
    // workers stand at the barrier: one window of 32 words a round (LB_W = 1) measured 6.1 ms per build, two 6.2, four 6.7,
    // eight 7.2 (21.5 M particles).  Doing the look-back a stage earlier, right behind the tile's own publication, is worse
    // (7.5 ms): the neighbours' words of the same iteration are then not out yet and the helper spins on them.
    if (helper) resolve_carry(k);
    else if (k + 1 < nk) stage_a(k + 1);
    __syncthreads();
    // no barrier behind the scan: its results (and the published word) belong to tile k+1, which the workers touch in the next
    // iteration; they go straight on with tile k, whose scan was made an iteration ago.  The bulk copies of tile k+NST-1 go
    // into tile k-1's slot: every thread left it before the last barrier.
    if (helper) {
      if (k + 1 < nk) scan_publish(k + 1);
      if (lane == 0) issue(k + NST - 1);
    }
    const int cbase = s_misc[q][1];          // in a register: the helper recycles the slot's words while the slots are flushed
    if (!helper) {
    // ---- B stage: destinations and stores ----------------------------------------------------------------------------------
    const int carry = s_misc[q][2];
    const unsigned char *sl = ring + (size_t)q * SLOT_BYTES;
    const int *nidS = reinterpret_cast<const int *>(sl + SLOT_NID);
    const float4 *recS = reinterpret_cast<const float4 *>(sl + SLOT_REC);
    const unsigned *idxS = reinterpret_cast<const unsigned *>(sl + SLOT_IDX);
    float4 r[IPT];
    int cl[IPT];             // > 0: child the particle's node sends its left particles to; else not a particle of a splitting node
    unsigned fl = 0;         // bit j: left flag of item j
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      // one straight-line body for the three kinds of particle (finished earlier / leaf now / moving to a child): the
      // stores go through selected pointers instead of three branches
      const int p = j * PTPB + t, i = tile0 + p;
      const int ndv = i < n ? nidS[p] : -1;
      const bool act = ndv >= 0;
      const float4 *np = reinterpret_cast<const float4 *>(nodes + (act ? ndv : 0));
      const float4 a = __ldg(np), d = __ldg(np + 3);
      int cnt = __float_as_int(a.x), off = __float_as_int(a.y);
      cl[j] = act ? __float_as_int(a.z) : 0;
      r[j] = recS[p];
      const unsigned ix = idxS[p];
      const bool split = cl[j] > 0;
      if (act && !split && __float_as_int(d.w) == -2 && cnt == 0) {     // reversed orphan (rare): its parent's range
        const float4 pa = __ldg(reinterpret_cast<const float4 *>(nodes + __float_as_int(d.z)));
        cnt = __float_as_int(pa.x); off = __float_as_int(pa.y);
      }
      // a leaf (or an orphan holding a degenerate node's particles) reaches its final place; a reversed one is mirrored
      const int fin = (__float_as_int(d.w) == -2) ? 2 * off + cnt - 1 - i : i;
      const int e = j * (PTPB / 32) + w;
      const unsigned fb = s_fb[q][e], hb = s_hb[q][e];
      const unsigned hm = hb & (below | (1u << lane));      // node starts at or before this lane in its row
      const int hl = 31 - __clz(hm | 1u);                      // last such start (0 if none: masked below)
      int lb = __popc(fb & below & ~((1u << hl) - 1u));
      if (!hm) lb = s_cv[q][e] + __popc(fb & below) + (s_cf[q][e] ? 0 : carry);     // the node started in an earlier row / tile
      const int flag = (fb >> lane) & 1u;
      fl |= (unsigned)flag << j;
      const int dest = flag ? off + lb : off + cnt - 1 - ((i - off) - lb);
      float4 *prec = split ? rec_out + dest : src4 + fin;
      unsigned *pidx = split ? idx_out + dest : perm + fin;
      if (act) { *prec = r[j]; *pidx = ix; }
      if (i < n) nid_out[split ? dest : i] = split ? cl[j] + 1 - flag : -1;
      if (split && i == off + cnt - 1) {                      // the node's last particle knows the split
        const int is = lb + flag;
        nodes[cl[j]].count = is;            nodes[cl[j]].offset = off;              // :731-746
        nodes[cl[j] + 1].count = cnt - is;  nodes[cl[j] + 1].offset = off + is;     // :732,762
      }
    }
    // ---- box and centroid sums of the children (warp-uniform loop over the distinct nodes of the warp's four rows) -----
    {
      int mk = INT_MAX;
#pragma unroll
      for (int j = 0; j < IPT; ++j) if (cl[j] > 0) mk = min(mk, cl[j]);
      int key = __reduce_min_sync(0xffffffffu, mk);
      while (key != INT_MAX) {
        unsigned bl[6], br[6];
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) { bl[qq] = 0xffffffffu; bl[3 + qq] = 0u; br[qq] = 0xffffffffu; br[3 + qq] = 0u; }
        long long s[4] = {0, 0, 0, 0};
        int next = INT_MAX;
        bool anyl = false, anyr = false;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
          if (cl[j] == key) {
            const unsigned ex = enc_f(r[j].x), ey = enc_f(r[j].y), ez = enc_f(r[j].z);
            if ((fl >> j) & 1u) {
              bl[0] = min(bl[0], ex); bl[1] = min(bl[1], ey); bl[2] = min(bl[2], ez);
              bl[3] = max(bl[3], ex); bl[4] = max(bl[4], ey); bl[5] = max(bl[5], ez);
              part_add_sums(s, r[j], sx, sm);
              anyl = true;
            } else {
              br[0] = min(br[0], ex); br[1] = min(br[1], ey); br[2] = min(br[2], ez);
              br[3] = max(br[3], ex); br[4] = max(br[4], ey); br[5] = max(br[5], ez);
              anyr = true;
            }
          } else if (cl[j] > key) next = min(next, cl[j]);
        }
        const bool wl = __any_sync(0xffffffffu, anyl), wr = __any_sync(0xffffffffu, anyr);
        int lm[4][3];        // warp sums of the 21-bit limbs of the four per-lane sums (|per-lane sum| < 2^52)
        if (wl) {
#pragma unroll
          for (int qq = 0; qq < 3; ++qq) { bl[qq] = __reduce_min_sync(0xffffffffu, bl[qq]); bl[3 + qq] = __reduce_max_sync(0xffffffffu, bl[3 + qq]); }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            lm[qq][0] = __reduce_add_sync(0xffffffffu, (int)(s[qq] & 0x1fffffll));
            lm[qq][1] = __reduce_add_sync(0xffffffffu, (int)((s[qq] >> 21) & 0x1fffffll));
            lm[qq][2] = __reduce_add_sync(0xffffffffu, (int)(s[qq] >> 42));
          }
        }
        if (wr) {
#pragma unroll
          for (int qq = 0; qq < 3; ++qq) { br[qq] = __reduce_min_sync(0xffffffffu, br[qq]); br[3 + qq] = __reduce_max_sync(0xffffffffu, br[3 + qq]); }
        }
        if (lane == 0) {
          const int sl2 = key - cbase;
          if (sl2 >= 0 && sl2 + 1 < NSLOT) {
            if (wl) {
              PassSlot &S = slots[sl2];
              for (int qq = 0; qq < 3; ++qq) { atomicMin(&S.umin[qq], bl[qq]); atomicMax(&S.umax[qq], bl[3 + qq]); }
              for (int qq = 0; qq < 4; ++qq) { atomicAdd(&S.limb[qq][0], lm[qq][0]); atomicAdd(&S.limb[qq][1], lm[qq][1]); atomicAdd(&S.limb[qq][2], lm[qq][2]); }
              S.used = 1;
            }
            if (wr) {
              PassSlot &S = slots[sl2 + 1];
              for (int qq = 0; qq < 3; ++qq) { atomicMin(&S.umin[qq], br[qq]); atomicMax(&S.umax[qq], br[3 + qq]); }
              S.used = 1;
            }
          } else {
            if (wl) {
              NodeAcc &A = acc[key];
              for (int qq = 0; qq < 3; ++qq) { atomicMin(&A.umin[qq], bl[qq]); atomicMax(&A.umax[qq], bl[3 + qq]); }
              for (int qq = 0; qq < 4; ++qq)
                add_split(&A.lo[qq], &A.hi[qq], (long long)lm[qq][0] + ((long long)lm[qq][1] << 21) + ((long long)lm[qq][2] << 42));
            }
            if (wr) {
              NodeAcc &A = acc[key + 1];
              for (int qq = 0; qq < 3; ++qq) { atomicMin(&A.umin[qq], br[qq]); atomicMax(&A.umax[qq], br[3 + qq]); }
            }
          }
        }
        key = __reduce_min_sync(0xffffffffu, next);
      }
    }
    }   // workers
    __syncthreads();
    if (t < NSLOT && slots[t].used) {
      NodeAcc &A = acc[cbase + t];
      PassSlot &S = slots[t];
      for (int qq = 0; qq < 3; ++qq) { atomicMin(&A.umin[qq], S.umin[qq]); atomicMax(&A.umax[qq], S.umax[qq]); }
      for (int qq = 0; qq < 4; ++qq)
        add_split(&A.lo[qq], &A.hi[qq], (long long)S.limb[qq][0] + ((long long)S.limb[qq][1] << 21) + ((long long)S.limb[qq][2] << 42));
      pslot_reset(S);
    }
  }
}

// single-block exclusive scan (per-tile and per-node counts: up to a few hundred thousand values); 8 consecutive values
// per thread and iteration, so 21 k tile counts take 3 rounds of the block scan instead of 21
__global__ void __launch_bounds__(1024) k_scan(const unsigned *__restrict__ in, unsigned *__restrict__ out,
                                               long long n, unsigned long long *__restrict__ d_total) {
  __shared__ int s_w[34];
  constexpr int PER = 8;
  unsigned long long running = 0;
  for (long long b = 0; b < n; b += (long long)blockDim.x * PER) {
    const long long i0 = b + (long long)threadIdx.x * PER;
    unsigned v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = (i0 + j < n) ? in[i0 + j] : 0u;
    unsigned t = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { const unsigned x = v[j]; v[j] = t; t += x; }
    int total;
    const unsigned long long base = running + (unsigned long long)(unsigned)block_excl_scan((int)t, s_w, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) if (i0 + j < n) out[i0 + j] = (unsigned)(base + v[j]);
    running += (unsigned long long)(unsigned)total;
  }
  if (threadIdx.x == 0 && d_total) *d_total = running;
}

// ---- monopole moments of internal nodes, bottom-up one level at a time (RCBForceTree.cxx:856-889) -------
__global__ void k_moments(Node *__restrict__ nodes, const float4 *__restrict__ src4, int begin, int end) {
  int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= end) return;
  int cl = nodes[k].cl, cr = nodes[k].cr;
  if (cl == 0 && cr == 0) return;
  float s = 0.f;
  int ch[2] = {cl, cr};
  for (int q = 0; q < 2; ++q) {
    int c = ch[q];
    if (c > 0 && nodes[c].count > 0) {
      float add = (nodes[c].count <= 1) ? src4[nodes[c].offset].w : nodes[c].ppm;
      s = __fadd_rn(s, add);
    }
  }
  nodes[k].ppm = s;
}

// ---- quadrupole pseudo-particles (RCBForceTree<12>, -S): reference RCBForceTree.cxx:229-272,519-569 ----------------
// Every node carries 12 pseudo-particles on an icosahedron of radius tdr = 0.9 * (smallest distance from the
// centroid to a face of the tight box) about the centroid; their masses reproduce the node's monopole, dipole and
// quadrupole (pp<12>, :536-569).  Leaves sum over their particles (:789-797), internal nodes over the pseudo-particles
// of their children, or over a child's particles when it has <= 12 of them (:856-889).  All arithmetic is float in the
// reference's order with explicit round-to-nearest intrinsics (the x86-64 build has no FMA); a leaf's sum over
// particles is a fixed-shape warp reduction instead of the reference's sequential loop, so masses agree to FP32
// rounding (~1e-6 relative), not bit for bit.  Stored as float4 (x,y,z,m): pp12[12*node + j].
__constant__ float c_d12[3][12];   // the 12-point spherical 4-design of Hardin & Sloane, point order of :229-272
static const float ICO_P = 0.525731112119134f, ICO_Q = 0.85065080835204f;
static const float h_d12[3][12] = {
    {0, 0, ICO_P, -ICO_P, ICO_Q, -ICO_Q, 0, 0, -ICO_P, ICO_P, -ICO_Q, ICO_Q},
    {ICO_Q, ICO_Q, 0, 0, ICO_P, ICO_P, -ICO_Q, -ICO_Q, 0, 0, -ICO_P, -ICO_P},
    {ICO_P, -ICO_P, ICO_Q, ICO_Q, 0, 0, -ICO_P, ICO_P, -ICO_Q, -ICO_Q, 0, 0}};

struct PPFrame {          // the target node's design points relative to its centroid
  float xj[12], yj[12], zj[12], rj[12];
  bool nz[12];            // rj2 != 0 (:553)
  float xc[3], tdr;
};
__device__ __forceinline__ float pp_tdr(const Node &nd) {                      // :519-523 times ppContract = 0.9
  float m = fminf(__fsub_rn(nd.xmax[0], nd.xc[0]), fminf(__fsub_rn(nd.xmax[1], nd.xc[1]), fminf(__fsub_rn(nd.xmax[2], nd.xc[2]),
            fminf(__fsub_rn(nd.xc[0], nd.xmin[0]), fminf(__fsub_rn(nd.xc[1], nd.xmin[1]), __fsub_rn(nd.xc[2], nd.xmin[2]))))));
  return __fmul_rn(0.9f, m);
}
__device__ __forceinline__ void pp_frame(const Node &nd, PPFrame &F, float4 *out /* 12 positions, or null */) {
  F.tdr = pp_tdr(nd);
  F.xc[0] = nd.xc[0]; F.xc[1] = nd.xc[1]; F.xc[2] = nd.xc[2];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const float px = __fadd_rn(__fmul_rn(F.tdr, c_d12[0][j]), nd.xc[0]);       // :525-534
    const float py = __fadd_rn(__fmul_rn(F.tdr, c_d12[1][j]), nd.xc[1]);
    const float pz = __fadd_rn(__fmul_rn(F.tdr, c_d12[2][j]), nd.xc[2]);
    if (out) out[j] = make_float4(px, py, pz, 0.f);
    F.xj[j] = __fsub_rn(px, nd.xc[0]); F.yj[j] = __fsub_rn(py, nd.xc[1]); F.zj[j] = __fsub_rn(pz, nd.xc[2]);   // :548-550
    const float rj2 = __fadd_rn(__fadd_rn(__fmul_rn(F.xj[j], F.xj[j]), __fmul_rn(F.yj[j], F.yj[j])), __fmul_rn(F.zj[j], F.zj[j]));
    F.nz[j] = rj2 != 0.0f;
    F.rj[j] = __fsqrt_rn(rj2);
  }
}
// ppm[j] += mass * (odr0 + odr1 + odr2) for one source (:541-566)
__device__ __forceinline__ void pp_add(const PPFrame &F, float x, float y, float z, float mass, float (&ppm)[12]) {
  const float K = 12.0f;
  const float odr0 = __fdiv_rn(1.0f, K), k3 = __fdiv_rn(3.0f, K), k5 = __fdiv_rn(5.0f, K);
  const float xi = __fsub_rn(x, F.xc[0]), yi = __fsub_rn(y, F.xc[1]), zi = __fsub_rn(z, F.xc[2]);
  const float ri = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xi, xi), __fmul_rn(yi, yi)), __fmul_rn(zi, zi)));
  const float q = __fdiv_rn(ri, F.tdr);
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float odr1 = 0.f, odr2 = 0.f;
    if (F.nz[j]) {
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(xi, F.xj[j]), __fmul_rn(yi, F.yj[j])), __fmul_rn(zi, F.zj[j]));
      const float aij = __fdiv_rn(dot, __fmul_rn(ri, F.rj[j]));
      odr1 = __fmul_rn(__fmul_rn(k3, q), aij);
      // (5/K)*q*q*0.5*(3*aij*aij - 1): the 0.5 is a double literal, so the last product is exact in double and
      // rounded to float once -- identical to rounding A*B in float and halving
      const float A = __fmul_rn(__fmul_rn(k5, q), q);
      const float B = __fsub_rn(__fmul_rn(__fmul_rn(3.0f, aij), aij), 1.0f);
      odr2 = __fmul_rn(__fmul_rn(A, B), 0.5f);
    }
    ppm[j] = __fadd_rn(ppm[j], __fmul_rn(mass, __fadd_rn(__fadd_rn(odr0, odr1), odr2)));
  }
}

// one warp per leaf: pseudo-particle masses from the leaf's particles (:789-797)
__global__ void __launch_bounds__(256) k_pp12_leaf(const Node *__restrict__ nodes, int n_nodes, int ppn,
                                                    const float4 *__restrict__ src4, float4 *__restrict__ pp12) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= n_nodes) return;
  const Node nd = nodes[k];
  if (nd.cl != 0 || nd.cr != 0 || nd.count <= 0) return;      // internal nodes: k_pp12_internal; empty orphans: unused
  PPFrame F;
  float4 pos[12];
  pp_frame(nd, F, pos);
  float ppm[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) ppm[j] = 0.f;
  // count <= 12: the pseudo-particles are never used (:791); count > ppn: a degenerate split, the reference leaves the
  // masses at zero because both children are empty (:727-729,856-889)
  if (nd.count > 12 && nd.count <= ppn) {
    for (int i = lane; i < nd.count; i += 32) {
      const float4 r = __ldg(src4 + nd.offset + i);
      pp_add(F, r.x, r.y, r.z, r.w, ppm);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j)
      for (int o = 16; o > 0; o >>= 1) ppm[j] = __fadd_rn(ppm[j], __shfl_xor_sync(0xffffffffu, ppm[j], o));
  }
  if (lane < 12) {
    float4 v = pos[0]; float m = ppm[0];
#pragma unroll
    for (int j = 1; j < 12; ++j) if (lane == j) { v = pos[j]; m = ppm[j]; }
    v.w = m;
    pp12[12 * (size_t)k + lane] = v;
  }
}

// one thread per internal node of one level, bottom-up: masses from the children (:856-889)
__global__ void __launch_bounds__(128) k_pp12_internal(const Node *__restrict__ nodes, int begin, int end,
                                                        const float4 *__restrict__ src4, float4 *__restrict__ pp12) {
  const int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= end) return;
  const Node nd = nodes[k];
  if (nd.cl == 0 && nd.cr == 0) return;
  PPFrame F;
  float4 pos[12];
  pp_frame(nd, F, pos);
  float ppm[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) ppm[j] = 0.f;
  const int ch[2] = {nd.cl, nd.cr};
  for (int s = 0; s < 2; ++s) {
    const int c = ch[s];
    if (c <= 0) continue;
    const int cnt = nodes[c].count, off = nodes[c].offset;
    if (cnt <= 0) continue;
    const float4 *srcp = (cnt <= 12) ? (src4 + off) : (pp12 + 12 * (size_t)c);
    const int m = (cnt <= 12) ? cnt : 12;
    for (int i = 0; i < m; ++i) {
      const float4 r = srcp[i];
      pp_add(F, r.x, r.y, r.z, r.w, ppm);
    }
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) { float4 v = pos[j]; v.w = ppm[j]; pp12[12 * (size_t)k + j] = v; }
}

// ---- permute the caller-visible arrays into tree order (the reference does this in place, :648-669) ------
// Also the last level's "pass": at the level where no node splits any more, every particle that is not final yet belongs to a
// leaf (or to an orphan holding a degenerate node's particles) and only has to be written to its final place -- mirrored if
// the leaf is a reversed one (k_split_pass) -- so the split pass skips that level and the gather takes those particles
// straight from the level's record buffers: one pass over the particles less.
// PART bit 0: positions and mass (what the walk and the force kernel read); also completes src4 / perm for the last level's
// leaves.  Bit 1: phi, id, mask; bit 2: velocities.  haccsr_kick (resident particles) does all three at once; haccsr_kick_host
// runs part 1 in the build and parts 2 + 4 later on its copy stream, from the completed permutation: they wait for the upload
// of arrays the build does not read, and must not hold it up (api.cu).
template <int PART>
__global__ void __launch_bounds__(256) k_gather(Soa in, Soa out, float4 *__restrict__ src4, unsigned *__restrict__ perm,
                                                const float4 *__restrict__ rec, const unsigned *__restrict__ idx,
                                                const int *__restrict__ nid, const Node *__restrict__ nodes, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (PART & 1) {
      const int nd = nid[i];
      unsigned p;
      float4 r;
      int fin = i;
      if (nd < 0) {
        p = perm[i]; r = src4[i];
      } else {
        p = idx[i]; r = rec[i];
        const float4 *np = reinterpret_cast<const float4 *>(nodes + nd);
        const float4 d = __ldg(np + 3);
        if (__float_as_int(d.w) == -2) {
          float4 a = __ldg(np);
          if (__float_as_int(a.x) == 0) a = __ldg(reinterpret_cast<const float4 *>(nodes + __float_as_int(d.z)));   // orphan: the parent's range
          fin = 2 * __float_as_int(a.y) + __float_as_int(a.x) - 1 - i;
        }
        src4[fin] = r;
        if (PART != 7) perm[fin] = p;       // the other parts run later, from the completed permutation
      }
      out.x[fin] = r.x; out.y[fin] = r.y; out.z[fin] = r.z; out.mass[fin] = r.w;
      if (PART & 4) { out.vx[fin] = in.vx[p]; out.vy[fin] = in.vy[p]; out.vz[fin] = in.vz[p]; }
      if (PART & 2) { out.phi[fin] = in.phi[p]; out.id[fin] = in.id[p]; out.mask[fin] = in.mask[p]; }
    } else {
      const unsigned p = perm[i];
      if (PART & 4) { out.vx[i] = in.vx[p]; out.vy[i] = in.vy[p]; out.vz[i] = in.vz[p]; }
      if (PART & 2) { out.phi[i] = in.phi[p]; out.id[i] = in.id[p]; out.mask[i] = in.mask[p]; }
    }
  }
}

// the later parts of the split gather (velocities, phi, id, mask), queued by haccsr_kick_host on its copy stream behind the uploads
int gather_payload(haccsr_ctx *c, cudaStream_t st) {
  const int n = (int)c->n_tree;
  if (n <= 0) return 0;
  int grid_lin = (n + TPB - 1) / TPB;
  if (grid_lin > c->sm_count * 16) grid_lin = c->sm_count * 16;
  // build_tree has swapped the two array sets: `alt` is the upload, `cur` the tree-ordered set
  k_gather<6><<<grid_lin, 256, 0, st>>>(c->alt, c->cur, c->src4.p, c->perm.p, nullptr, nullptr, nullptr, c->nodes.p, n);
  c->launches++;
  HSR_CUDA(cudaGetLastError());
  return 0;
}

// ---- host orchestration -------------------------------------------------------------------------------
// Long inputs (per-particle flags: the out-of-box compaction, the refresh classification) are scanned in three launches:
// per-block sums of 2048 elements, the single-block scan above over those sums, per-block scan + base.  The single-block
// kernel alone walks the array 1024 elements at a time: 60 ms for the 152 M flags of a 512^3 sub-volume.
static constexpr int SCAN_TPB = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_TPB * SCAN_IPT;
__global__ void __launch_bounds__(SCAN_TPB) k_scan_reduce(const unsigned *__restrict__ in, long long n,
                                                           unsigned *__restrict__ blocksum) {
  __shared__ unsigned s_w[SCAN_TPB / 32];
  const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
  unsigned v = 0;
  if (i0 + SCAN_IPT <= n) {
    const uint4 a = *reinterpret_cast<const uint4 *>(in + i0), b = *reinterpret_cast<const uint4 *>(in + i0 + 4);
    v = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  } else {
    for (int j = 0; j < SCAN_IPT; ++j) if (i0 + j < n) v += in[i0 + j];
  }
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < SCAN_TPB / 32; ++w) t += s_w[w];
    blocksum[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_apply(const unsigned *__restrict__ in, unsigned *__restrict__ out,
                                                          long long n, const unsigned *__restrict__ blockbase) {
  __shared__ int s_w[34];
  const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
  unsigned v[SCAN_IPT];
  if (i0 + SCAN_IPT <= n) {
    const uint4 a = *reinterpret_cast<const uint4 *>(in + i0), b = *reinterpret_cast<const uint4 *>(in + i0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) v[j] = (i0 + j < n) ? in[i0 + j] : 0u;
  }
  unsigned t = 0;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) { const unsigned x = v[j]; v[j] = t; t += x; }
  int total;
  const unsigned base = blockbase[blockIdx.x] + (unsigned)block_excl_scan((int)t, s_w, &total);
  if (i0 + SCAN_IPT <= n) {
    *reinterpret_cast<uint4 *>(out + i0) = make_uint4(base + v[0], base + v[1], base + v[2], base + v[3]);
    *reinterpret_cast<uint4 *>(out + i0 + 4) = make_uint4(base + v[4], base + v[5], base + v[6], base + v[7]);
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) if (i0 + j < n) out[i0 + j] = base + v[j];
  }
}

int scan_exclusive(haccsr_ctx *c, const unsigned *in, unsigned *out, int64_t n, unsigned long long *d_total) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (n < 16 * SCAN_TILE || !aligned) {
    k_scan<<<1, 1024, 0, c->stream>>>(in, out, (long long)n, d_total);
    c->launches++;
    HSR_CUDA(cudaGetLastError());
    return 0;
  }
  const int64_t nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
  HSR_TRY(c->scan_tmp.ensure(2 * (size_t)nblk + 2));
  unsigned *bsum = c->scan_tmp.p, *bbase = c->scan_tmp.p + nblk + 1;
  k_scan_reduce<<<(unsigned)nblk, SCAN_TPB, 0, c->stream>>>(in, (long long)n, bsum);
  k_scan<<<1, 1024, 0, c->stream>>>(bsum, bbase, (long long)nblk, d_total);
  k_scan_apply<<<(unsigned)nblk, SCAN_TPB, 0, c->stream>>>(in, out, (long long)n, bbase);
  c->launches += 3;
  HSR_CUDA(cudaGetLastError());
  return 0;
}

int build_tree(haccsr_ctx *c, int64_t n64, const float lo[3], const float hi[3], int64_t ppn64, int tdpts) {
  c->tdpts = tdpts;
  if (n64 >= (int64_t)INT_MAX - 2 * TILE) { set_error("too many particles for 32-bit indexing: %lld", (long long)n64); return 1; }
  const int n = (int)n64;
  const int ppn = (int)(ppn64 > INT_MAX ? INT_MAX : ppn64);
  cudaStream_t st = c->stream;
  // split pass geometry: 256 threads x 4 particles (1024-particle tiles, two blocks per SM); HACCSR_PASS_TPB=128 selects
  // 512-particle tiles, three blocks per SM (measured: 6.6 against 5.9 ms per build at 21.5 M particles)
  // (per context: the shared-memory attribute and the occupancy belong to the context's device)
  if (!c->pass_tpb) {
    const char *e = getenv("HACCSR_PASS_TPB");
    const int want = (e && atoi(e) == 128) ? 128 : 256;
    const void *fn = want == 256 ? (const void *)k_split_pass<256> : (const void *)k_split_pass<128>;
    const size_t dyn = (size_t)NST * want * IPT * 24;
    int occ = 0;
    HSR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    HSR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, want + 32, dyn));
    if (occ < 1) { set_error("k_split_pass does not fit on an SM"); return 2; }
    c->pass_tpb = want; c->pass_occ = occ;
  }
  const int ptpb = c->pass_tpb, occ_sp = c->pass_occ;
  const int PT = ptpb * IPT;
  const int ntiles = (n + PT - 1) / PT;
  const int grid_sp = ntiles < c->sm_count * occ_sp ? (ntiles > 0 ? ntiles : 1) : c->sm_count * occ_sp;
  // node pool: every split node has > ppn particles and two non-empty children, so nodes <= 2N-1; in
  // practice ~4N/ppn.  Start generously and grow (rebuild) if the pool runs out.
  int64_t want_nodes = 1024 + 8 * (n64 / (ppn > 0 ? ppn : 1));
  if (want_nodes > 2 * n64 + 8) want_nodes = 2 * n64 + 8;
  if ((int64_t)c->nodes.cap > want_nodes) want_nodes = (int64_t)c->nodes.cap;
  // nodes of one level: children of the previous level's split nodes, which hold > ppn particles each
  const int64_t level_cap = 2 * (n64 / ((int64_t)ppn + 1)) + 2;
  const int64_t nblk_cap = (level_cap + NTPB - 1) / NTPB + 1;

  // (+8: the bulk copies of the last tile are rounded up to 16 bytes)
  HSR_TRY(c->recA.ensure(n + 8)); HSR_TRY(c->recB.ensure(n + 8)); HSR_TRY(c->src4.ensure(n + 1));
  HSR_TRY(c->idxA.ensure(n + 8)); HSR_TRY(c->idxB.ensure(n + 8)); HSR_TRY(c->perm.ensure(n + 1));
  HSR_TRY(c->nidA.ensure(n + 8)); HSR_TRY(c->nidB.ensure(n + 8));
  HSR_TRY(c->tile_desc.ensure((size_t)ntiles + 1)); HSR_TRY(c->node_desc.ensure((size_t)nblk_cap + 1));
  HSR_TRY(c->scratch_u32.ensure(16));

  for (int attempt = 0; attempt < 6; ++attempt) {
    HSR_TRY(c->nodes.ensure(want_nodes)); HSR_TRY(c->acc.ensure(want_nodes)); HSR_TRY(c->tot.ensure(want_nodes));
    const int max_nodes = (int)(c->nodes.cap > (size_t)INT_MAX ? INT_MAX : c->nodes.cap);
    unsigned *maxima = c->scratch_u32.p;
    float *scales = reinterpret_cast<float *>(c->scratch_u32.p + 4);
    HSR_CUDA(cudaMemsetAsync(maxima, 0, 4 * sizeof(unsigned), st));
    c->n_tree = n; c->n_nodes = 1; c->n_levels = 0;
    if (n == 0) {
      k_root_init<<<1, 1, 0, st>>>(c->nodes.p, c->acc.p, 0, make_float3(lo[0], lo[1], lo[2]),
                                   make_float3(hi[0], hi[1], hi[2]), maxima, scales, c->d_state);
      c->launches++;
      HSR_CUDA(cudaGetLastError());
      c->level_begin[0] = 0; c->level_end[0] = 1; c->n_levels = 1; c->unit_mass = false;
      return 0;
    }
    HSR_CUDA(cudaMemsetAsync(c->tile_desc.p, 0, ((size_t)ntiles + 1) * sizeof(unsigned long long), st));
    HSR_CUDA(cudaMemsetAsync(c->node_desc.p, 0, ((size_t)nblk_cap + 1) * sizeof(unsigned long long), st));
    int grid_lin = (n + TPB - 1) / TPB;
    if (grid_lin > c->sm_count * 16) grid_lin = c->sm_count * 16;
    k_init_records<<<grid_lin, TPB, 0, st>>>(c->cur.x, c->cur.y, c->cur.z, c->cur.mass, n, c->recA.p, c->idxA.p,
                                             c->nidA.p, maxima);
    k_root_init<<<1, 1, 0, st>>>(c->nodes.p, c->acc.p, n, make_float3(lo[0], lo[1], lo[2]),
                                 make_float3(hi[0], hi[1], hi[2]), maxima, scales, c->d_state);
    k_cm_tile<<<(n + CM_TILE - 1) / CM_TILE, TPB, 0, st>>>(c->recA.p, c->nidA.p, n, c->acc.p, scales);
    c->launches += 3;
    HSR_CUDA(cudaGetLastError());

    float4 *rec = c->recA.p, *rec_o = c->recB.p;
    unsigned *idx = c->idxA.p, *idx_o = c->idxB.p;
    int *nid = c->nidA.p, *nid_o = c->nidB.p;
    // The levels are queued without waiting for each other's outcome: a launch for a level the tree does not have returns
    // at once (BuildState::nsplit).  The host looks at the state after the expected depth and then every two levels.
    int depth = 0;
    {
      int64_t q = n64 / (ppn > 0 ? ppn : 1);
      int lg = 0;
      while (((int64_t)1 << lg) < q) ++lg;
      depth = lg + 2;
    }
    int level = 0, n_levels = 0;
    bool overflow = false;
    while (!n_levels && !overflow) {
      if (depth > 127) depth = 127;
      for (; level < depth; ++level) {
        int64_t nl_cap = level < 40 ? ((int64_t)1 << level) : level_cap;
        if (nl_cap > level_cap) nl_cap = level_cap;
        const int gb = (int)((nl_cap + NTPB - 1) / NTPB);
        k_level_nodes<<<gb < 1 ? 1 : gb, NTPB, 0, st>>>(c->nodes.p, c->acc.p, c->tot.p, c->d_state, c->node_desc.p, level, ppn,
                                                        max_nodes, scales);
        {
          // cooperative launch: the look-back needs every block of the grid resident (the launch fails instead of hanging)
          const float4 *a0 = rec; const unsigned *a1 = idx; const int *a2 = nid; Node *a3 = c->nodes.p; NodeAcc *a4 = c->acc.p;
          const float *a5 = scales; BuildState *a6 = c->d_state; unsigned long long *a7 = c->tile_desc.p;
          int a8 = level, a9 = n, a10 = ntiles; float4 *a11 = rec_o; unsigned *a12 = idx_o; int *a13 = nid_o;
          float4 *a14 = c->src4.p; unsigned *a15 = c->perm.p;
          void *args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9, &a10, &a11, &a12, &a13, &a14, &a15};
          const void *fn = ptpb == 256 ? (const void *)k_split_pass<256> : (const void *)k_split_pass<128>;
          HSR_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid_sp), dim3(ptpb + 32), args, (size_t)NST * PT * 24, st));
        }
        c->launches += 2;
        float4 *tr = rec; rec = rec_o; rec_o = tr;
        unsigned *ti = idx; idx = idx_o; idx_o = ti;
        int *tn = nid; nid = nid_o; nid_o = tn;
      }
      HSR_CUDA(cudaGetLastError());
      HSR_CUDA(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(BuildState), cudaMemcpyDeviceToHost, st));
      HSR_CUDA(cudaStreamSynchronize(st));
      const BuildState &S = *c->h_state;
      if (S.error) { overflow = true; break; }
      for (int L = 0; L < level; ++L) if (S.nsplit[L] == 0) { n_levels = L + 1; break; }
      if (!n_levels) {
        if (level >= 127) { set_error("tree deeper than 127 levels"); return 1; }
        depth = level + 2;
      }
    }
    if (overflow) {
      if ((int64_t)c->nodes.cap >= 2 * n64 + 8) { set_error("node pool exhausted at its upper bound"); return 1; }
      want_nodes = (int64_t)c->nodes.cap * 2;
      if (want_nodes > 2 * n64 + 8) want_nodes = 2 * n64 + 8;
      // DevBuf::ensure reallocates (contents are rebuilt from scratch on the next attempt)
      continue;
    }
    {
      const BuildState &S = *c->h_state;
      c->n_levels = n_levels; c->n_nodes = S.lvl_end[n_levels - 1]; c->unit_mass = S.unit_mass != 0;
      for (int L = 0; L < n_levels; ++L) { c->level_begin[L] = S.lvl_begin[L]; c->level_end[L] = S.lvl_end[L]; }
    }
    const int nnodes = c->n_nodes;
    {
      // the record buffers the last level read from: level L reads A if L is even
      const bool even = ((c->n_levels - 1) & 1) == 0;
      const float4 *lrec = even ? c->recA.p : c->recB.p;
      const unsigned *lidx = even ? c->idxA.p : c->idxB.p;
      const int *lnid = even ? c->nidA.p : c->nidB.p;
      if (c->wait_up2) {
        // haccsr_kick_host: positions only, nothing to wait for; the six arrays still on their way are permuted on the copy
        // stream (gather_payload) and the kick itself is deferred (force.cu: apply_kick)
        c->wait_up2 = false;
        k_gather<1><<<grid_lin, 256, 0, st>>>(c->cur, c->alt, c->src4.p, c->perm.p, lrec, lidx, lnid, c->nodes.p, n);
      } else {
        k_gather<7><<<grid_lin, 256, 0, st>>>(c->cur, c->alt, c->src4.p, c->perm.p, lrec, lidx, lnid, c->nodes.p, n);
      }
      c->launches++;
      HSR_CUDA(cudaGetLastError());
    }
    for (int L = c->n_levels - 2; L >= 0; --L) {
      int nl = c->level_end[L] - c->level_begin[L];
      k_moments<<<(nl + 255) / 256, 256, 0, st>>>(c->nodes.p, c->src4.p, c->level_begin[L], c->level_end[L]);
      c->launches++;
    }
    if (tdpts == 12) {
      static bool design_loaded[64] = {false};
      if (!design_loaded[c->device & 63]) {
        HSR_CUDA(cudaMemcpyToSymbolAsync(c_d12, h_d12, sizeof(h_d12), 0, cudaMemcpyHostToDevice, st));
        design_loaded[c->device & 63] = true;
      }
      HSR_TRY(c->pp12.ensure(12 * (size_t)nnodes + 12));
      k_pp12_leaf<<<(int)(((int64_t)nnodes * 32 + 255) / 256), 256, 0, st>>>(c->nodes.p, nnodes, ppn, c->src4.p, c->pp12.p);
      c->launches++;
      for (int L = c->n_levels - 2; L >= 0; --L) {
        int nl = c->level_end[L] - c->level_begin[L];
        k_pp12_internal<<<(nl + 127) / 128, 128, 0, st>>>(c->nodes.p, c->level_begin[L], c->level_end[L], c->src4.p, c->pp12.p);
        c->launches++;
      }
      HSR_CUDA(cudaGetLastError());
    }
    // particles beyond n (out-of-box tail) keep their place: copy them across so cur/alt can be swapped
    if (c->n_resident > n) {
      size_t m = (size_t)(c->n_resident - n);
      Soa &a = c->cur, &b = c->alt;
      float *fa[8] = {a.x, a.y, a.z, a.vx, a.vy, a.vz, a.mass, a.phi};
      float *fb[8] = {b.x, b.y, b.z, b.vx, b.vy, b.vz, b.mass, b.phi};
      for (int q = 0; q < 8; ++q) HSR_CUDA(cudaMemcpyAsync(fb[q] + n, fa[q] + n, m * sizeof(float), cudaMemcpyDeviceToDevice, st));
      HSR_CUDA(cudaMemcpyAsync(b.id + n, a.id + n, m * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
      HSR_CUDA(cudaMemcpyAsync(b.mask + n, a.mask + n, m * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st));
    }
    Soa tmp = c->cur; c->cur = c->alt; c->alt = tmp;
    c->order_epoch++;
    return 0;
  }
  set_error("tree build did not converge on a node pool size");
  return 1;
}

}  // namespace haccsr
