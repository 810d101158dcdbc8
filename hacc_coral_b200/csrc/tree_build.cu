// tree_build.cu -- device build of the RCB force tree (level-synchronous), sm_100a.
//
// What it computes is the tree of the reference's createRCBForceTree
// (reference src/halo_finder/RCBForceTree.cxx:778-918): every node gets the TIGHT bounding box and the
// mass-weighted centroid of its particles (cm, src/halo_finder/BGQCM.c:181-212); a node with more than
// ppn particles is cut on the longest edge of that box (:838-852) at the centroid coordinate (:720),
// particles with key < pivot going left (:640); a split that leaves one side empty keeps the node as an
// oversized leaf with two orphan children (:727-729).
//
// How it is computed is not the reference's recursion.  All nodes of one level are processed together by
// passes over fixed 1024-particle tiles (HBM-bound streaming kernels):
//   A  k_cm_tile        per-warp box / centroid sums by node run -> per-block slots -> per-node accumulators
//   B  k_level_finalize  one block: box, centroid, split decision, BFS child allocation (prefix scan)
//   C1 k_left_count     per-tile count of "goes left" flags + local prefixes at node boundaries
//   C2 k_scan           exclusive scan of tile counts  => global prefix L(i) of left flags
//   B2 k_set_children   per node: is = L(end) - L(begin); child counts / offsets; degenerate splits
//   C3 k_scatter        stable two-way partition of (x,y,z,m | index) records into the other buffer;
//                       particles of finished leaves are written once to the final tree-order arrays.
// The centroid sums are accumulated as wide fixed-point integers (two 64-bit words, add_split), so they are exact and independent
// of summation order: the build is deterministic, and the float centroid equals the reference's
// (double-accumulated) one except where the reference's own rounding error straddles a float boundary.
// Children are numbered breadth-first (parent index < child index, as the reference guarantees at :808-809).
// Left blocks keep input order like the reference; right blocks are kept stable too (the reference's
// right-block order is an artefact of its swap loop and only affects FP32 summation order).
#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace haccsr {

static constexpr int TILE = 1024;      // particles per tile / thread block
static constexpr int TPB = 256;        // threads per block in tile kernels
static constexpr int IPT = TILE / TPB; // items per thread (k_cm_tile: consecutive; flag/scatter passes: striped i = base + j*TPB + t)
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
                            float *scales, LevelInfo *info) {
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
  info->begin = 0; info->end = 1; info->nsplit = 0; info->error = 0;
  info->unit_mass = (n > 0 && maxima[2] == 0u) ? 1 : 0;
}

// ---- pass A: per-tile partial sums ---------------------------------------------------------------
struct Part {
  unsigned umin[3], umax[3];
  long long s[4];
};
__device__ __forceinline__ void part_reset(Part &p) {
  p.umin[0] = p.umin[1] = p.umin[2] = 0xffffffffu; p.umax[0] = p.umax[1] = p.umax[2] = 0u;
  p.s[0] = p.s[1] = p.s[2] = p.s[3] = 0;
}
__device__ __forceinline__ void part_add(Part &p, const float4 &r, float sx, float sm) {
  unsigned ex = enc_f(r.x), ey = enc_f(r.y), ez = enc_f(r.z);
  p.umin[0] = min(p.umin[0], ex); p.umax[0] = max(p.umax[0], ex);
  p.umin[1] = min(p.umin[1], ey); p.umax[1] = max(p.umax[1], ey);
  p.umin[2] = min(p.umin[2], ez); p.umax[2] = max(p.umax[2], ez);
  // the product w*x is formed in float exactly as in BGQCM.c:203-206, then summed exactly
  p.s[0] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.x), sx));
  p.s[1] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.y), sx));
  p.s[2] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.z), sx));
  p.s[3] += __float2ll_rn(__fmul_rn(r.w, sm));
}

struct Slot {
  unsigned umin[3], umax[3];
  unsigned used, pad;
  unsigned long long s[4];
};

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

// A warp owns CM_IPT rows of 32 consecutive particles (coalesced 128-byte node-id and 512-byte record loads, CM_IPT of
// each in flight per lane).  Node ids are non-decreasing along the array, so those 32*CM_IPT particles are one to three
// runs; for each distinct node in turn (a warp-uniform loop) every lane sums its particles of that node, the 32 partials
// are combined by a plain warp reduction -- REDUX for the six box bounds, a shuffle tree for the four 64-bit sums -- and
// lane 0 adds the result to the block's shared-memory slot of the node.  (The first version striped items over the
// block and fell back to per-thread shared atomics; the second combined per-thread partials of 4 consecutive
// particles with a 14-word segmented warp scan -- 75 shuffles per 128 particles, issue-bound at 35 % of the HBM peak,
// profiles/r1j_build_ncu_summary.md.)
// (8 rows, no register cap: 4 rows or __launch_bounds__ minimum-block counts of 3-4 on any of the three tile kernels
// spill and were 4-18 % slower on the whole build)
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
  if (t < SMAX) {
    Slot &S = slots[t];
    S.umin[0] = S.umin[1] = S.umin[2] = 0xffffffffu; S.umax[0] = S.umax[1] = S.umax[2] = 0u;
    S.used = 0; S.s[0] = S.s[1] = S.s[2] = S.s[3] = 0;
  }
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

// Alternative k_cm_tile, selected with HACCSR_CM_KERNEL=warp (off by default: measured stand-alone only, see
// tools/microbench_cm.cu and profiles/r1o_microbench_cm.txt -- identical accumulators, 6-25 % faster -- but the parity suite
// has not been run on it inside the library yet).  Persistent warps over CONTIGUOUS particle ranges: the next rows' loads are
// in flight during the current rows' arithmetic, the node's per-lane partial stays in registers while the node id stays the
// same and is reduced and flushed with result-less global atomics only when the id changes; no shared memory, no barriers.
static constexpr int CMW_ROWS = 4;
__global__ void __launch_bounds__(TPB) k_cm_warp(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                 int per_warp, NodeAcc *__restrict__ acc, const float *__restrict__ scales) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long b64 = gw * (long long)per_warp;
  if (b64 >= n) return;
  const int begin = (int)b64, end = (int)min((long long)n, b64 + per_warp);
  const float sx = scales[0], sm = scales[1];
  constexpr int STEP = 32 * CMW_ROWS;
  int nd[CMW_ROWS], nd1[CMW_ROWS];
  float4 r[CMW_ROWS], r1[CMW_ROWS];
  auto load = [&](int pos, int (&d)[CMW_ROWS], float4 (&q)[CMW_ROWS]) {
#pragma unroll
    for (int k = 0; k < CMW_ROWS; ++k) { const int i = pos + 32 * k + lane; d[k] = (i < end) ? __ldcs(nid + i) : -1; }
#pragma unroll
    for (int k = 0; k < CMW_ROWS; ++k) if (d[k] >= 0) q[k] = __ldcs(rec + pos + 32 * k + lane);
  };
  auto reduce_flush = [&](Part &p, int node) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { p.umin[q] = __reduce_min_sync(0xffffffffu, p.umin[q]); p.umax[q] = __reduce_max_sync(0xffffffffu, p.umax[q]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) p.s[q] += shfl_xor_ll(p.s[q], o);
    }
    if (lane == 0) {
      NodeAcc &A = acc[node];
      for (int k = 0; k < 3; ++k) { atomicMin(&A.umin[k], p.umin[k]); atomicMax(&A.umax[k], p.umax[k]); }
      for (int k = 0; k < 4; ++k) add_split(&A.lo[k], &A.hi[k], p.s[k]);
    }
  };
  load(begin, nd, r);
  int K = -1;                 // node whose per-lane partial is carried in p
  Part p; part_reset(p);
  for (int pos = begin; pos < end; pos += STEP) {
    load(pos + STEP, nd1, r1);         // rows past `end` load nothing
    int mn = INT_MAX;
#pragma unroll
    for (int k = 0; k < CMW_ROWS; ++k) if (nd[k] >= 0) mn = min(mn, nd[k]);
    int cur = __reduce_min_sync(0xffffffffu, mn);
    while (cur != INT_MAX) {           // warp-uniform: the distinct nodes of this step, in increasing order
      if (cur != K) {
        if (K >= 0) reduce_flush(p, K);
        part_reset(p); K = cur;
      }
      int next = INT_MAX;
#pragma unroll
      for (int k = 0; k < CMW_ROWS; ++k) {
        if (nd[k] == cur) part_add(p, r[k], sx, sm);
        else if (nd[k] > cur) next = min(next, nd[k]);
      }
      cur = __reduce_min_sync(0xffffffffu, next);
    }
#pragma unroll
    for (int k = 0; k < CMW_ROWS; ++k) { nd[k] = nd1[k]; r[k] = r1[k]; }
  }
  if (K >= 0) reduce_flush(p, K);
}

// ---- pass B: finalize the nodes of one level, decide splits, allocate children breadth-first ----------
// Three launches so that a level of 65 k nodes is as parallel as a level of one: (a) per node: box, centroid, split
// decision; (b) k_scan over the split flags; (c) per split node: children at next_base + 2 * rank, so the numbering is
// breadth-first and deterministic (parent index < child index, as the reference guarantees at :808-809).
__global__ void __launch_bounds__(256) k_level_decide(Node *__restrict__ nodes, const NodeAcc *__restrict__ acc,
                                                       int begin, int end, int ppn, const float *__restrict__ scales,
                                                       unsigned *__restrict__ flags, LevelInfo *__restrict__ info) {
  const int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k == begin) info->error = 0;
  if (k >= end) return;
  const double undo = ldexp(1.0, (int)scales[2]);
  Node nd = nodes[k];
  int split = 0, d = -1;
  if (nd.count > 0) {
    const NodeAcc a = acc[k];
    for (int q = 0; q < 3; ++q) { nd.xmin[q] = dec_f(a.umin[q]); nd.xmax[q] = dec_f(a.umax[q]); }
    const double sw = split_to_double(a.lo[3], a.hi[3]);
    for (int q = 0; q < 3; ++q) {
      const double sxq = split_to_double(a.lo[q], a.hi[q]);
      nd.xc[q] = (float)((sxq / sw) * undo);                       // BGQCM.c:209-211
    }
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
  nd.split = split ? d : -1;
  nodes[k] = nd;
  flags[k - begin] = (unsigned)split;
}

__global__ void __launch_bounds__(256) k_level_children(Node *__restrict__ nodes, NodeAcc *__restrict__ acc, int begin,
                                                         int end, int next_base, int max_nodes,
                                                         const unsigned *__restrict__ flags,
                                                         const unsigned *__restrict__ ranks,
                                                         const unsigned long long *__restrict__ d_total,
                                                         LevelInfo *__restrict__ info) {
  const int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = (int)*d_total;
  const bool overflow = next_base + 2 * ns >= max_nodes;      // the last child index must stay below max_nodes
  if (k == begin) {
    info->begin = next_base; info->end = next_base + (overflow ? 0 : 2 * ns); info->nsplit = overflow ? 0 : ns;
    info->error = overflow ? 1 : 0;
  }
  if (k >= end || !flags[k - begin]) return;
  if (overflow) { nodes[k].split = -1; return; }
  Node nd = nodes[k];
  const int d = nd.split;
  const int cl = next_base + 2 * (int)ranks[k - begin];
  nodes[k].cl = cl; nodes[k].cr = cl + 1;     // provisional; cleared by k_set_children on a degenerate split
  Node c;
  c.count = 0; c.offset = 0; c.cl = 0; c.cr = 0; c.ppm = 0.f; c.parent = k; c.split = -1;
  for (int q = 0; q < 3; ++q) { c.xmin[q] = nd.xmin[q]; c.xmax[q] = nd.xmax[q]; c.xc[q] = 0.f; }
  Node l = c, r = c;
  l.xmax[d] = nd.xc[d]; r.xmin[d] = nd.xc[d];                  // :747,763
  nodes[cl] = l; nodes[cl + 1] = r;
  NodeAcc z;
  for (int q = 0; q < 3; ++q) { z.umin[q] = 0xffffffffu; z.umax[q] = 0u; }
  for (int q = 0; q < 4; ++q) { z.lo[q] = 0; z.hi[q] = 0; }
  acc[cl] = z; acc[cl + 1] = z;
}

// ---- shared by C1 and C3: left flags of a tile and their exclusive prefix in particle order --------------
struct ItemInfo { int sp, flag, excl, offset, count, cl, cr; };

// striped tile layout of the flag passes: particle i = tile*TILE + j*TPB + t
__device__ __forceinline__ void st_load_nid(const int *__restrict__ nid, int n, int tile, int ntiles, int nd[IPT]) {
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const int i = tile * TILE + j * TPB + (int)threadIdx.x;
    nd[j] = (tile < ntiles && i < n) ? __ldcs(nid + i) : -1;
  }
}
__device__ __forceinline__ void st_load_rec(const float4 *__restrict__ rec, int tile, const int nd[IPT], float4 r[IPT]) {
#pragma unroll
  for (int j = 0; j < IPT; ++j) if (nd[j] >= 0) r[j] = __ldcs(rec + tile * TILE + j * TPB + (int)threadIdx.x);
}
__device__ __forceinline__ void st_load_idx(const unsigned *__restrict__ idx, int tile, const int nd[IPT], unsigned x[IPT]) {
#pragma unroll
  for (int j = 0; j < IPT; ++j) if (nd[j] >= 0) x[j] = __ldcs(idx + tile * TILE + j * TPB + (int)threadIdx.x);
}

__device__ __forceinline__ int tile_flags_scan(const Node *__restrict__ nodes, const int nd[IPT], const float4 r[IPT],
                                               ItemInfo it[IPT], int *s_w) {
  // In particle order the tile is IPT rows of TPB/32 warps.  One ballot per (row, warp) gives the left count of 32
  // consecutive particles; the IPT*TPB/32 = 32 counts are scanned by one warp -- a single barrier pair instead of IPT
  // block-wide scans.  The node's metadata comes in three independent 16-byte loads (no load depends on the split
  // dimension read by another).
  static_assert(IPT * (TPB / 32) == 32, "one warp scans the per-(row, warp) counts");
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  unsigned bal[IPT];
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    int sp = -1, flag = 0;
    it[j].offset = 0; it[j].count = 0; it[j].cl = 0; it[j].cr = 0;
    if (nd[j] >= 0) {
      const float4 *np = reinterpret_cast<const float4 *>(nodes + nd[j]);
      const float4 a = __ldg(np), c = __ldg(np + 2), d = __ldg(np + 3);
      sp = __float_as_int(d.w);
      it[j].count = __float_as_int(a.x); it[j].offset = __float_as_int(a.y);
      it[j].cl = __float_as_int(a.z); it[j].cr = __float_as_int(a.w);
      if (sp >= 0) {
        const float pivot = sp == 0 ? c.z : (sp == 1 ? c.w : d.x);             // xc[sp], RCBForceTree.cxx:720
        flag = comp(r[j], sp) < pivot;                                         // :640
      }
    }
    bal[j] = __ballot_sync(0xffffffffu, flag);
    it[j].sp = sp; it[j].flag = flag;
    if (lane == 0) s_w[j * (TPB / 32) + w] = __popc(bal[j]);
  }
  __syncthreads();
  if (w == 0) {
    int x = s_w[lane], inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    s_w[lane] = inc - x;
    if (lane == 31) s_w[32] = inc;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < IPT; ++j) it[j].excl = s_w[j * (TPB / 32) + w] + __popc(bal[j] & ((1u << lane) - 1u));
  return s_w[32];
}

__global__ void __launch_bounds__(TPB) k_left_count(const float4 *__restrict__ rec, const int *__restrict__ nid,
                                                    const Node *__restrict__ nodes, int n, int ntiles,
                                                    unsigned *__restrict__ tilecount, int *__restrict__ lstart,
                                                    int *__restrict__ lend) {
  __shared__ int s_w[2][34];     // alternating per iteration: a slow warp may still read the previous tile's prefixes
  const int stride = gridDim.x;
  int nd[IPT], nd1[IPT], nd2[IPT];
  float4 r[IPT], r1[IPT];
  ItemInfo it[IPT];
  int tile = blockIdx.x;
  st_load_nid(nid, n, tile, ntiles, nd);
  st_load_nid(nid, n, tile + stride, ntiles, nd1);
  st_load_rec(rec, tile, nd, r);
  for (int k = 0; tile < ntiles; tile += stride, ++k) {
    st_load_nid(nid, n, tile + 2 * stride, ntiles, nd2);
    st_load_rec(rec, tile + stride, nd1, r1);
    const int total = tile_flags_scan(nodes, nd, r, it, s_w[k & 1]);
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if (it[j].sp >= 0) {
        const int i = tile * TILE + j * TPB + (int)threadIdx.x;
        if (i == it[j].offset) lstart[nd[j]] = it[j].excl;
        if (i == it[j].offset + it[j].count - 1) lend[nd[j]] = it[j].excl + it[j].flag;
      }
    }
    if (threadIdx.x == 0) tilecount[tile] = (unsigned)total;
#pragma unroll
    for (int j = 0; j < IPT; ++j) { nd[j] = nd1[j]; nd1[j] = nd2[j]; r[j] = r1[j]; }
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

__global__ void k_set_children(Node *__restrict__ nodes, int begin, int end, const unsigned *__restrict__ tilebase,
                               const int *__restrict__ lstart, const int *__restrict__ lend,
                               int *__restrict__ lbase, int *__restrict__ nleft) {
  int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= end) return;
  Node nd = nodes[k];
  if (nd.split < 0) return;
  int ts = nd.offset / TILE, te = (nd.offset + nd.count - 1) / TILE;
  int Ls = (int)tilebase[ts] + lstart[k], Le = (int)tilebase[te] + lend[k];
  int is = Le - Ls;
  if (is == 0 || is == nd.count) {
    // degenerate split (RCBForceTree.cxx:727-729): the node stays a leaf, its two children stay empty
    // orphans, and its monopole mass is the sum over those empty children, i.e. zero (:856-889).
    // nodes[k].split keeps the dimension for the rest of this level so that k_scatter recomputes the
    // same left flags k_left_count counted; nleft = -1 marks the node as finished.
    nodes[k].cl = 0; nodes[k].cr = 0; nodes[k].ppm = 0.f; nleft[k] = -1;
    return;
  }
  lbase[k] = Ls; nleft[k] = is;
  nodes[nd.cl].count = is;            nodes[nd.cl].offset = nd.offset;         // :731-746
  nodes[nd.cr].count = nd.count - is; nodes[nd.cr].offset = nd.offset + is;    // :732,762
}

__global__ void __launch_bounds__(TPB) k_scatter(const float4 *__restrict__ rec, const unsigned *__restrict__ idx,
                                                 const int *__restrict__ nid, const Node *__restrict__ nodes, int n,
                                                 int ntiles, const unsigned *__restrict__ tilebase,
                                                 const int *__restrict__ lbase, const int *__restrict__ nleft,
                                                 float4 *__restrict__ rec_out, unsigned *__restrict__ idx_out,
                                                 int *__restrict__ nid_out, float4 *__restrict__ src4,
                                                 unsigned *__restrict__ perm) {
  __shared__ int s_w[2][34];
  const int stride = gridDim.x;
  int nd[IPT], nd1[IPT], nd2[IPT];
  float4 r[IPT], r1[IPT];
  unsigned ix[IPT], ix1[IPT];
  ItemInfo it[IPT];
  int tile = blockIdx.x;
  st_load_nid(nid, n, tile, ntiles, nd);
  st_load_nid(nid, n, tile + stride, ntiles, nd1);
  st_load_rec(rec, tile, nd, r);
  st_load_idx(idx, tile, nd, ix);
  for (int k = 0; tile < ntiles; tile += stride, ++k) {
    st_load_nid(nid, n, tile + 2 * stride, ntiles, nd2);
    st_load_rec(rec, tile + stride, nd1, r1);
    st_load_idx(idx, tile + stride, nd1, ix1);
    // per-node split bookkeeping of this tile's particles: independent of the flag scan, issued before its barriers
    int nl[IPT], lb[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      nl[j] = 0; lb[j] = 0;
      if (nd[j] >= 0) { nl[j] = __ldg(nleft + nd[j]); lb[j] = __ldg(lbase + nd[j]); }
    }
    const int tb = (int)tilebase[tile];
    tile_flags_scan(nodes, nd, r, it, s_w[k & 1]);
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const int i = tile * TILE + j * TPB + (int)threadIdx.x;
      if (i >= n) continue;
      if (nd[j] < 0) { nid_out[i] = -1; continue; }
      if (it[j].sp < 0 || nl[j] < 0) {   // finished leaf (or degenerate split): final position reached
        src4[i] = r[j]; perm[i] = ix[j]; nid_out[i] = -1;
        continue;
      }
      const int off = it[j].offset;
      const int L = tb + it[j].excl - lb[j];       // left flags of this node before i
      int dest, child;
      if (it[j].flag) { dest = off + L; child = it[j].cl; }
      else { dest = off + nl[j] + ((i - off) - L); child = it[j].cr; }
      rec_out[dest] = r[j]; idx_out[dest] = ix[j]; nid_out[dest] = child;
    }
#pragma unroll
    for (int j = 0; j < IPT; ++j) { nd[j] = nd1[j]; nd1[j] = nd2[j]; r[j] = r1[j]; ix[j] = ix1[j]; }
  }
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
__global__ void __launch_bounds__(256) k_gather(Soa in, Soa out, const float4 *__restrict__ src4,
                                                const unsigned *__restrict__ perm, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned p = perm[i];
    float4 r = src4[i];
    out.x[i] = r.x; out.y[i] = r.y; out.z[i] = r.z; out.mass[i] = r.w;
    out.vx[i] = in.vx[p]; out.vy[i] = in.vy[p]; out.vz[i] = in.vz[p];
    out.phi[i] = in.phi[p]; out.id[i] = in.id[p]; out.mask[i] = in.mask[p];
  }
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
  const int ntiles = (n + TILE - 1) / TILE;
  // persistent grids of the tile kernels: as many blocks as are resident at once, each walking tiles with stride gridDim
  static int occ_lc = 0, occ_sc = 0;
  if (!occ_lc) {
    HSR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sc, k_scatter, TPB, 0));
    HSR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lc, k_left_count, TPB, 0));
    if (occ_lc < 1 || occ_sc < 1) { occ_lc = 0; set_error("tile kernels do not fit on an SM"); return 2; }
  }
  // experimental k_cm_warp (HACCSR_CM_KERNEL=warp): one contiguous range of particles per resident warp
  int cmw_per_warp = 0, cmw_blocks = 0;
  {
    const char *e = getenv("HACCSR_CM_KERNEL");
    if (e && !strcmp(e, "warp") && n > 0) {
      int occ = 0;
      HSR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cm_warp, TPB, 0));
      const long long warps = (long long)c->sm_count * (occ > 0 ? occ : 1) * (TPB / 32);
      const int step = 32 * CMW_ROWS;
      long long pw = ((long long)n + warps - 1) / warps;
      pw = (pw + step - 1) / step * step;
      cmw_per_warp = (int)pw;
      cmw_blocks = (int)(((long long)n + pw * (TPB / 32) - 1) / (pw * (TPB / 32)));
    }
  }
  const int grid_lc = ntiles < c->sm_count * occ_lc ? (ntiles > 0 ? ntiles : 1) : c->sm_count * occ_lc;
  const int grid_sc = ntiles < c->sm_count * occ_sc ? (ntiles > 0 ? ntiles : 1) : c->sm_count * occ_sc;
  // node pool: every split node has > ppn particles and two non-empty children, so nodes <= 2N-1; in
  // practice ~4N/ppn.  Start generously and grow (rebuild) if the pool runs out.
  int64_t want_nodes = 1024 + 8 * (n64 / (ppn > 0 ? ppn : 1));
  if (want_nodes > 2 * n64 + 8) want_nodes = 2 * n64 + 8;
  if ((int64_t)c->nodes.cap > want_nodes) want_nodes = (int64_t)c->nodes.cap;

  HSR_TRY(c->recA.ensure(n + 1)); HSR_TRY(c->recB.ensure(n + 1)); HSR_TRY(c->src4.ensure(n + 1));
  HSR_TRY(c->idxA.ensure(n + 1)); HSR_TRY(c->idxB.ensure(n + 1)); HSR_TRY(c->perm.ensure(n + 1));
  HSR_TRY(c->nidA.ensure(n + 1)); HSR_TRY(c->nidB.ensure(n + 1));
  HSR_TRY(c->tilecount.ensure(ntiles + 1)); HSR_TRY(c->tilebase.ensure(ntiles + 1));
  HSR_TRY(c->scratch_u32.ensure(16));

  for (int attempt = 0; attempt < 6; ++attempt) {
    HSR_TRY(c->nodes.ensure(want_nodes)); HSR_TRY(c->acc.ensure(want_nodes));
    HSR_TRY(c->lstart.ensure(want_nodes)); HSR_TRY(c->lend.ensure(want_nodes));
    HSR_TRY(c->lbase.ensure(want_nodes)); HSR_TRY(c->nleft.ensure(want_nodes));
    const int max_nodes = (int)(c->nodes.cap > (size_t)INT_MAX ? INT_MAX : c->nodes.cap);
    unsigned *maxima = c->scratch_u32.p;
    float *scales = reinterpret_cast<float *>(c->scratch_u32.p + 4);
    HSR_CUDA(cudaMemsetAsync(maxima, 0, 4 * sizeof(unsigned), st));
    c->n_tree = n; c->n_nodes = 1; c->n_levels = 0;
    if (n == 0) {
      k_root_init<<<1, 1, 0, st>>>(c->nodes.p, c->acc.p, 0, make_float3(lo[0], lo[1], lo[2]),
                                   make_float3(hi[0], hi[1], hi[2]), maxima, scales, c->d_level);
      c->launches++;
      HSR_CUDA(cudaGetLastError());
      c->level_begin[0] = 0; c->level_end[0] = 1; c->n_levels = 1; c->unit_mass = false;
      return 0;
    }
    int grid_lin = (n + TPB - 1) / TPB;
    if (grid_lin > c->sm_count * 16) grid_lin = c->sm_count * 16;
    k_init_records<<<grid_lin, TPB, 0, st>>>(c->cur.x, c->cur.y, c->cur.z, c->cur.mass, n, c->recA.p, c->idxA.p,
                                             c->nidA.p, maxima);
    k_root_init<<<1, 1, 0, st>>>(c->nodes.p, c->acc.p, n, make_float3(lo[0], lo[1], lo[2]),
                                 make_float3(hi[0], hi[1], hi[2]), maxima, scales, c->d_level);
    c->launches += 2;
    HSR_CUDA(cudaGetLastError());

    float4 *rec = c->recA.p, *rec_o = c->recB.p;
    unsigned *idx = c->idxA.p, *idx_o = c->idxB.p;
    int *nid = c->nidA.p, *nid_o = c->nidB.p;
    int begin = 0, end = 1, nnodes = 1, level = 0;
    bool overflow = false;
    for (;; ++level) {
      if (level >= 127) { set_error("tree deeper than 127 levels"); return 1; }
      c->level_begin[level] = begin; c->level_end[level] = end;
      if (cmw_per_warp) k_cm_warp<<<cmw_blocks, TPB, 0, st>>>(rec, nid, n, cmw_per_warp, c->acc.p, scales);
      else k_cm_tile<<<(n + CM_TILE - 1) / CM_TILE, TPB, 0, st>>>(rec, nid, n, c->acc.p, scales);
      {
        const int nl = end - begin, gb = (nl + 255) / 256;
        HSR_TRY(c->split_flag.ensure((size_t)nl + 1)); HSR_TRY(c->split_rank.ensure((size_t)nl + 1));
        k_level_decide<<<gb, 256, 0, st>>>(c->nodes.p, c->acc.p, begin, end, ppn, scales, c->split_flag.p, c->d_level);
        HSR_TRY(scan_exclusive(c, c->split_flag.p, c->split_rank.p, nl, c->d_counters + 13));
        k_level_children<<<gb, 256, 0, st>>>(c->nodes.p, c->acc.p, begin, end, nnodes, max_nodes, c->split_flag.p,
                                             c->split_rank.p, c->d_counters + 13, c->d_level);
      }
      c->launches += 3;
      HSR_CUDA(cudaMemcpyAsync(c->h_level, c->d_level, sizeof(LevelInfo), cudaMemcpyDeviceToHost, st));
      HSR_CUDA(cudaStreamSynchronize(st));
      LevelInfo li = *c->h_level;
      if (level == 0) c->unit_mass = li.unit_mass != 0;
      if (li.error) { overflow = true; break; }
      if (li.nsplit > 0) {
        k_left_count<<<grid_lc, TPB, 0, st>>>(rec, nid, c->nodes.p, n, ntiles, c->tilecount.p, c->lstart.p, c->lend.p);
        c->launches++;
        HSR_TRY(scan_exclusive(c, c->tilecount.p, c->tilebase.p, ntiles, nullptr));
        int nl = end - begin;
        k_set_children<<<(nl + 255) / 256, 256, 0, st>>>(c->nodes.p, begin, end, c->tilebase.p, c->lstart.p,
                                                          c->lend.p, c->lbase.p, c->nleft.p);
        c->launches++;
      }
      k_scatter<<<grid_sc, TPB, 0, st>>>(rec, idx, nid, c->nodes.p, n, ntiles, c->tilebase.p, c->lbase.p, c->nleft.p, rec_o,
                                        idx_o, nid_o, c->src4.p, c->perm.p);
      c->launches++;
      HSR_CUDA(cudaGetLastError());
      nnodes = li.end;
      if (li.nsplit == 0) break;
      begin = li.begin; end = li.end;
      float4 *tr = rec; rec = rec_o; rec_o = tr;
      unsigned *ti = idx; idx = idx_o; idx_o = ti;
      int *tn = nid; nid = nid_o; nid_o = tn;
    }
    if (overflow) {
      if ((int64_t)c->nodes.cap >= 2 * n64 + 8) { set_error("node pool exhausted at its upper bound"); return 1; }
      want_nodes = (int64_t)c->nodes.cap * 2;
      if (want_nodes > 2 * n64 + 8) want_nodes = 2 * n64 + 8;
      // DevBuf::ensure reallocates (contents are rebuilt from scratch on the next attempt)
      continue;
    }
    c->n_nodes = nnodes; c->n_levels = level + 1;
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
    if (c->wait_up2) { HSR_CUDA(cudaStreamWaitEvent(st, c->ev_up2, 0)); c->wait_up2 = false; }
    k_gather<<<grid_lin, 256, 0, st>>>(c->cur, c->alt, c->src4.p, c->perm.p, n);
    c->launches++;
    HSR_CUDA(cudaGetLastError());
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
    return 0;
  }
  set_error("tree build did not converge on a node pool size");
  return 1;
}

}  // namespace haccsr
