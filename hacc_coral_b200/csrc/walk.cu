// walk.cu -- interaction-list construction on the device, sm_100a.
//
// Rules (reference src/halo_finder/RCBForceTree.cxx): a leaf is a sink if its tight box touches the force
// box (:1166-1170); its list is built by a LIFO walk from the root (:934-936) in which ancestors of the
// sink are always opened (:947-962), any other node is accepted as a monopole if its box diagonal^2 does
// not exceed dist2*tan^2(theta) (:965-1021) -- then dropped if its centroid is farther than rmax (:1024),
// else appended as its own particles (count <= TDPTS, :1033-1049) or as TDPTS pseudo-particles (:1053-1062) --
// opened leaves are appended whole (:1063-1080), children of opened internal nodes are queued only if
// their box is within rmax of the sink's box in every dimension (:1088-1123), and the sink leaf itself
// comes last (:1126-1139).  All acceptance arithmetic is float, evaluated in the reference's order with
// explicit __f*_rn intrinsics so that no FMA contraction can flip a decision.
//
// What differs from the reference is the product: instead of copying coordinates into four stack arrays
// of VMAX = 16384 entries (:921,940), the walk emits (start,count) RANGES of the tree-ordered float4
// particle array (adjacent leaves are merged into one range) plus, per sink leaf, one contiguous block
// of accepted pseudo-particles in a pool -- 8 B per list node instead of 16 B per list particle, no
// length limit, and every range is directly a 1-D TMA bulk copy for the force kernel.
// One thread walks one sink leaf (the walk is latency-bound pointer chasing in L2; there are ~N/400
// leaves); two passes (count, scan, fill) size the output exactly.
#include "common.cuh"

#include <limits.h>
#include <math.h>

namespace haccsr {

static constexpr int STACK = 192;

struct WalkParams {
  float flo[3], fhi[3];
  float rmax, rmax2, tan_oa;
  int n_nodes;
  int tdpts;               // pseudo-particles per accepted node: 1 (monopole) or 12 (quadrupole)
  const float4 *pp12;      // tdpts == 12: the nodes' pseudo-particles (tree_build.cu)
};

__device__ __forceinline__ void load_node(const Node *__restrict__ nodes, int k, Node &nd) {
  const float4 *p = reinterpret_cast<const float4 *>(nodes + k);
  float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  nd.count = __float_as_int(a.x); nd.offset = __float_as_int(a.y); nd.cl = __float_as_int(a.z); nd.cr = __float_as_int(a.w);
  nd.xmin[0] = b.x; nd.xmin[1] = b.y; nd.xmin[2] = b.z; nd.xmax[0] = b.w;
  nd.xmax[1] = c.x; nd.xmax[2] = c.y; nd.xc[0] = c.z; nd.xc[1] = c.w;
  nd.xc[2] = d.x; nd.ppm = d.y; nd.parent = __float_as_int(d.z); nd.split = __float_as_int(d.w);
}

// pending-range merger: adjacent leaves (in either direction) become one range
struct Emit {
  unsigned ps, pc;     // pending range
  unsigned nr;         // ranges flushed
  unsigned np;         // pseudo-particles emitted
  unsigned long long len;  // sources so far
};

template <bool FILL>
__device__ __forceinline__ void emit_flush(Emit &e, uint2 *__restrict__ out) {
  if (e.pc) {
    if (FILL) out[e.nr] = make_uint2(e.ps, e.pc);
    e.nr++;
    e.pc = 0;
  }
}
template <bool FILL>
__device__ __forceinline__ void emit_range(Emit &e, unsigned s, unsigned c, uint2 *__restrict__ out) {
  if (c == 0) return;
  e.len += c;
  if (e.pc) {
    if (s + c == e.ps) { e.ps = s; e.pc += c; return; }
    if (e.ps + e.pc == s) { e.pc += c; return; }
    emit_flush<FILL>(e, out);
  }
  e.ps = s; e.pc = c;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_walk(const Node *__restrict__ nodes, WalkParams P,
                                              unsigned *__restrict__ n_ranges, unsigned *__restrict__ n_pseudo,
                                              unsigned *__restrict__ list_len, const unsigned *__restrict__ range_off,
                                              const unsigned *__restrict__ pseudo_off, uint2 *__restrict__ ranges,
                                              float4 *__restrict__ pool, int *__restrict__ err) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= P.n_nodes) return;
  Node T;
  load_node(nodes, tl, T);
  bool sink = (T.cl == 0 && T.cr == 0 && T.count > 0);
  if (sink) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      sink = sink && ((T.xmax[i] < P.fhi[i] && T.xmax[i] > P.flo[i]) || (T.xmin[i] < P.fhi[i] && T.xmin[i] > P.flo[i]));
  }
  if (!sink) {
    if (!FILL) { n_ranges[tl] = 0; n_pseudo[tl] = 0; list_len[tl] = 0; }
    return;
  }
  uint2 *out = FILL ? ranges + range_off[tl] : nullptr;
  float4 *pp = FILL ? pool + pseudo_off[tl] : nullptr;
  Emit e; e.ps = 0; e.pc = 0; e.nr = 0; e.np = 0; e.len = 0;

  int stack[STACK];
  int sp = 0;
  stack[sp++] = 0;
  const int tend = T.offset + T.count;
  while (sp > 0) {
    int v = stack[--sp];
    const bool need_close = v < 0;       // pushed by an opened non-ancestor: box-distance test pending
    const int tln = need_close ? ~v : v;
    Node N;
    load_node(nodes, tln, N);
    if (need_close) {                                                       // :1088-1123 (tested at pop)
      bool close = true;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float dist = 0.f;
        if (T.xmax[i] < N.xmin[i]) dist = __fsub_rn(N.xmin[i], T.xmax[i]);
        else if (T.xmin[i] > N.xmax[i]) dist = __fsub_rn(T.xmin[i], N.xmax[i]);
        if (dist > P.rmax) close = false;
      }
      if (!close) continue;
    }
    // ancestors of the sink leaf contain its particle range (ranges nest; splits are never degenerate).
    // tln == tl only happens for a root that is itself a leaf; the reference then runs the acceptance
    // test on the leaf against itself (its ancestor test is `tln < tl`, :947) -- mirrored here.
    if (tln != tl && N.offset <= T.offset && tend <= N.offset + N.count) {  // :947-962
      if (N.cl > 0 && N.cl != tl) { if (sp < STACK) stack[sp++] = N.cl; else *err = 1; }
      if (N.cr > 0 && N.cr != tl) { if (sp < STACK) stack[sp++] = N.cr; else *err = 1; }
      continue;
    }
    float dx = __fsub_rn(N.xc[0], T.xc[0]), dy = __fsub_rn(N.xc[1], T.xc[1]), dz = __fsub_rn(N.xc[2], T.xc[2]);
    float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));      // :965-968
    float sx = __fsub_rn(N.xmax[0], N.xmin[0]), sy = __fsub_rn(N.xmax[1], N.xmin[1]), sz = __fsub_rn(N.xmax[2], N.xmin[2]);
    float l2 = fminf(__fmul_rn(sx, sx), fminf(__fmul_rn(sy, sy), __fmul_rn(sz, sz)));                 // :970-973
    float dtt2 = __fmul_rn(__fmul_rn(dist2, P.tan_oa), P.tan_oa);                                     // :975
    bool big = l2 > dtt2;
    if (!big) {
      // :986-1018 (useRealOA = false): the four corner pairs all give the same squared diagonal
      float ddx = __fsub_rn(__fsub_rn(N.xmin[0], T.xc[0]), __fsub_rn(N.xmax[0], T.xc[0]));
      float ddy = __fsub_rn(__fsub_rn(N.xmin[1], T.xc[1]), __fsub_rn(N.xmax[1], T.xc[1]));
      float ddz = __fsub_rn(__fsub_rn(N.xmin[2], T.xc[2]), __fsub_rn(N.xmax[2], T.xc[2]));
      float dh2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
      big = dh2 > dtt2;
    }
    if (!big) {
      if (dist2 > P.rmax2) continue;                                        // :1024-1029
      if (N.count <= P.tdpts) emit_range<FILL>(e, (unsigned)N.offset, (unsigned)N.count, out);   // :1033-1049
      else if (P.tdpts == 1) {                                              // :1053-1062
        if (FILL) pp[e.np] = make_float4(N.xc[0], N.xc[1], N.xc[2], N.ppm);
        e.np++; e.len++;
      } else {
        if (FILL) {
          const float4 *q = P.pp12 + 12 * (size_t)tln;
#pragma unroll
          for (int j = 0; j < 12; ++j) pp[e.np + j] = __ldg(q + j);
        }
        e.np += 12; e.len += 12;
      }
      continue;
    }
    if (N.cl == 0 && N.cr == 0) {                                           // :1063-1080
      emit_range<FILL>(e, (unsigned)N.offset, (unsigned)N.count, out);
      continue;
    }
    if (N.cl > 0) { if (sp < STACK) stack[sp++] = ~N.cl; else *err = 1; }
    if (N.cr > 0) { if (sp < STACK) stack[sp++] = ~N.cr; else *err = 1; }
  }
  emit_flush<FILL>(e, out);
  if (e.np) {     // the block of accepted pseudo-particles, as one range into the pool
    if (FILL) out[e.nr] = make_uint2(POOL_FLAG | pseudo_off[tl], e.np);
    e.nr++;
  }
  if (FILL) out[e.nr] = make_uint2((unsigned)T.offset, (unsigned)T.count);  // :1126-1139 self last
  e.nr++; e.len += (unsigned)T.count;
  if (!FILL) {
    n_ranges[tl] = e.nr; n_pseudo[tl] = e.np;
    list_len[tl] = (unsigned)(e.len > 0xffffffffull ? 0xffffffffull : e.len);
    if (e.len > 0xffffffffull) *err = 2;
  }
}

// census + pair statistics: one thread per node, block-reduced then atomics on 64-bit counters.
// counters: 0 leaves 1 empty leaves 2 max ppn 3 leaf particles 4 sink leaves 5 max list 6 pairs evaluated
__global__ void __launch_bounds__(256) k_census(const Node *__restrict__ nodes, int n_nodes,
                                                const unsigned *__restrict__ n_ranges,
                                                const unsigned *__restrict__ list_len,
                                                unsigned long long *__restrict__ counters) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long leaf = 0, empty = 0, parts = 0, sinks = 0, pairs = 0;
  unsigned maxppn = 0, maxlist = 0;
  if (k < n_nodes) {
    int cnt = nodes[k].count;
    bool isleaf = nodes[k].cl == 0 && nodes[k].cr == 0;
    if (k >= 1 && isleaf) {               // printStats starts at node 1 (RCBForceTree.cxx:468)
      if (cnt > 0) { leaf = 1; parts = (unsigned long long)cnt; maxppn = (unsigned)cnt; }
      else empty = 1;
    }
    if (n_ranges[k] > 0) {
      sinks = 1; maxlist = list_len[k];
      pairs = (unsigned long long)list_len[k] * (unsigned long long)cnt;
    }
  }
  // warp reduce
  for (int o = 16; o > 0; o >>= 1) {
    leaf += __shfl_down_sync(0xffffffffu, leaf, o); empty += __shfl_down_sync(0xffffffffu, empty, o);
    parts += __shfl_down_sync(0xffffffffu, parts, o); sinks += __shfl_down_sync(0xffffffffu, sinks, o);
    pairs += __shfl_down_sync(0xffffffffu, pairs, o);
    maxppn = max(maxppn, __shfl_down_sync(0xffffffffu, maxppn, o));
    maxlist = max(maxlist, __shfl_down_sync(0xffffffffu, maxlist, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (leaf) atomicAdd(&counters[0], leaf);
    if (empty) atomicAdd(&counters[1], empty);
    if (maxppn) atomicMax(&counters[2], (unsigned long long)maxppn);
    if (parts) atomicAdd(&counters[3], parts);
    if (sinks) atomicAdd(&counters[4], sinks);
    if (maxlist) atomicMax(&counters[5], (unsigned long long)maxlist);
    if (pairs) atomicAdd(&counters[6], pairs);
  }
}

int build_lists(haccsr_ctx *c, const float flo[3], const float fhi[3], float theta, haccsr_stats *st) {
  cudaStream_t s = c->stream;
  const int nn = c->n_nodes;
  HSR_TRY(c->n_ranges.ensure(nn + 1)); HSR_TRY(c->n_pseudo.ensure(nn + 1)); HSR_TRY(c->list_len.ensure(nn + 1));
  HSR_TRY(c->range_off.ensure(nn + 2)); HSR_TRY(c->pseudo_off.ensure(nn + 2));
  WalkParams P;
  for (int i = 0; i < 3; ++i) { P.flo[i] = flo[i]; P.fhi[i] = fhi[i]; }
  P.rmax = c->law.rmax; P.rmax2 = c->law.rmax2;
  P.tan_oa = tanf(theta);                       // RCBForceTree.cxx:381
  P.n_nodes = nn;
  P.tdpts = c->tdpts; P.pp12 = c->pp12.p;
  int *d_err = reinterpret_cast<int *>(c->d_counters + 15);
  HSR_CUDA(cudaMemsetAsync(c->d_counters, 0, 16 * sizeof(unsigned long long), s));
  const int grid = (nn + 127) / 128;
  k_walk<false><<<grid, 128, 0, s>>>(c->nodes.p, P, c->n_ranges.p, c->n_pseudo.p, c->list_len.p, nullptr, nullptr,
                                     nullptr, nullptr, d_err);
  c->launches++;
  HSR_TRY(scan_exclusive(c, c->n_ranges.p, c->range_off.p, nn, c->d_counters + 8));
  HSR_TRY(scan_exclusive(c, c->n_pseudo.p, c->pseudo_off.p, nn, c->d_counters + 9));
  k_census<<<(nn + 255) / 256, 256, 0, s>>>(c->nodes.p, nn, c->n_ranges.p, c->list_len.p, c->d_counters);
  c->launches++;
  HSR_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  HSR_CUDA(cudaStreamSynchronize(s));
  const unsigned long long *h = reinterpret_cast<const unsigned long long *>(c->h_counters);
  int err = (int)(h[15] & 0xffffffffull);
  if (err == 1) { set_error("walk stack overflow (tree deeper than expected)"); return 1; }
  if (err == 2) { set_error("interaction list longer than 2^32 sources"); return 1; }
  c->tot_ranges = (int64_t)h[8]; c->tot_pseudo = (int64_t)h[9];
  if (c->tot_ranges >= (int64_t)0xffffffffll || c->tot_pseudo >= (int64_t)0x7fffffffll) {
    set_error("interaction lists too large for 32-bit offsets (%lld ranges, %lld pseudo-particles)",
              (long long)c->tot_ranges, (long long)c->tot_pseudo);
    return 1;
  }
  // one past-the-end entry so consumers can read range_off[k+1]
  HSR_CUDA(cudaMemcpyAsync(c->range_off.p + nn, c->d_counters + 8, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
  HSR_CUDA(cudaMemcpyAsync(c->pseudo_off.p + nn, c->d_counters + 9, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
  HSR_TRY(c->ranges.ensure((size_t)c->tot_ranges + 1)); HSR_TRY(c->pool.ensure((size_t)c->tot_pseudo + 1));
  k_walk<true><<<grid, 128, 0, s>>>(c->nodes.p, P, c->n_ranges.p, c->n_pseudo.p, c->list_len.p, c->range_off.p,
                                    c->pseudo_off.p, c->ranges.p, c->pool.p, d_err);
  c->launches++;
  HSR_CUDA(cudaGetLastError());
  if (st) {
    st->nodes = nn;
    st->leaves = (int64_t)(h[0] + h[1]); st->empty_leaves = (int64_t)h[1]; st->max_ppn = (int64_t)h[2];
    st->mean_ppn = h[0] ? (double)h[3] / (double)h[0] : 0.0;
    st->levels = c->n_levels;
    st->sink_leaves = (int64_t)h[4]; st->max_list = (int64_t)h[5]; st->pairs_evaluated = h[6];
    st->list_ranges = c->tot_ranges; st->pseudo_particles = c->tot_pseudo;
  }
  return 0;
}

}  // namespace haccsr
