// force.cu -- work items of the force phase and the dispatch of the pair kernels (force_kernels.cuh).
#include "force_kernels.cuh"

namespace haccsr {

// the launchers are instantiated in force_v0.cu ... force_v4.cu (one arithmetic variant per translation unit)
extern template int launch_force<1, 1, true, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<1, 2, true, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<1, 2, false, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<1, 3, true, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<1, 3, false, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, true, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, false, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, true, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, false, 0>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, true, 1>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, false, 1>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, true, 1>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, false, 1>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, true, 2>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, false, 2>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, true, 2>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, false, 2>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, true, 3>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, false, 3>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, true, 3>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, false, 3>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, true, 4>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<6, 0, false, 4>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, true, 4>(haccsr_ctx *, const ForceParams &, int, bool);
extern template int launch_force<7, 0, false, 4>(haccsr_ctx *, const ForceParams &, int, bool);

__global__ void k_item_count(const Node *__restrict__ nodes, const unsigned *__restrict__ n_ranges, int n_nodes,
                             int policy, unsigned *__restrict__ item_cnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_nodes) return;
  unsigned c = 0;
  if (n_ranges[k] > 0) {
    const LeafCut lc = leaf_cut(nodes[k].count, policy);
    c = (unsigned)(lc.chunks + (lc.rem ? 1 : 0));
  }
  item_cnt[k] = c;
}
__global__ void k_item_fill(const Node *__restrict__ nodes, const unsigned *__restrict__ item_cnt,
                            const unsigned *__restrict__ item_off, const unsigned *__restrict__ n_pseudo, int n_nodes,
                            int policy, WorkItem *__restrict__ items) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_nodes) return;
  if (item_cnt[k] == 0) return;
  const int cnt = nodes[k].count, off = nodes[k].offset;
  const LeafCut lc = leaf_cut(cnt, policy);
  const int nopseudo = n_pseudo[k] == 0 ? 1 : 0;
  unsigned o = item_off[k];
  int g0 = 0;
  for (int q = 0; q < lc.chunks; ++q) {
    const int ng = chunk_groups(lc, q);
    WorkItem w;
    w.node = k; w.sink_begin = off + 32 * g0;
    const int left = cnt - 32 * g0 - lc.rem;           // the last group may be a padded one
    w.sink_count = left < 32 * ng ? left : 32 * ng;
    w.no_pseudo = nopseudo;
    items[o++] = w;
    g0 += ng;
  }
  if (lc.rem) {
    WorkItem w;
    w.node = k; w.sink_begin = off + 32 * lc.groups; w.sink_count = lc.rem; w.no_pseudo = nopseudo | ITEM_REM;
    items[o++] = w;
  }
}

// ---- longest-processing-time-first order -----------------------------------------------------------------
// CTAs are dispatched in index order, so items are sorted by decreasing work (sink groups x list length) with a
// counting sort on 16 bins per octave: the last CTAs to start are then the cheapest ones and the tail of the
// kernel, where SMs run out of work, shrinks from about one average item to about one small item.  The order
// has no effect on results: every item owns its sinks.
static constexpr int LPT_BINS = 512;
static constexpr int MAX_GROUPS = 8;
// Items can additionally be split into `groups` launches by particle range (haccsr_ctx::group_lo): used by
// haccsr_kick_host, which copies the velocities of a range to the host as soon as the launches up to that range have
// finished, so only the last range's copy is exposed after the force kernel.  Inside each group the order is LPT.
struct GroupBounds { int groups; int lo[MAX_GROUPS + 1]; };
__device__ __forceinline__ int lpt_bin(const WorkItem &w, const unsigned *__restrict__ list_len, const GroupBounds &G) {
  const int rem = (w.no_pseudo & ITEM_REM) ? 1 : 0;
  const float groups_eq = rem ? 0.25f * (float)((w.sink_count + REM_SINKS - 1) / REM_SINKS) : (float)((w.sink_count + 31) / 32);
  float work = groups_eq * (float)list_len[w.node];
  int b = (int)(16.0f * __log2f(work + 1.0f));
  b = b < 0 ? 0 : (b > LPT_BINS - 1 ? LPT_BINS - 1 : b);
  int g = 0;
#pragma unroll
  for (int q = 1; q < MAX_GROUPS; ++q) g += (q < G.groups && w.sink_begin >= G.lo[q]) ? 1 : 0;
  return (2 * g + rem) * LPT_BINS + (LPT_BINS - 1 - b);       // segment (group, chunk | remainder items), heavy items first
}
__global__ void k_lpt_hist(const WorkItem *__restrict__ items, const unsigned *__restrict__ list_len, int n,
                           const GroupBounds G, unsigned *__restrict__ hist) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[lpt_bin(items[i], list_len, G)], 1u);
}
__global__ void k_lpt_scatter(const WorkItem *__restrict__ items, const unsigned *__restrict__ list_len, int n,
                              const GroupBounds G, unsigned *__restrict__ cursor, WorkItem *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WorkItem w = items[i];
  out[atomicAdd(&cursor[lpt_bin(w, list_len, G)], 1u)] = w;
}

__global__ void __launch_bounds__(256) k_apply_kick(float *__restrict__ vx, float *__restrict__ vy, float *__restrict__ vz,
                                                    const float *__restrict__ ax, const float *__restrict__ ay,
                                                    const float *__restrict__ az, const float4 *__restrict__ src4, float fcoeff,
                                                    long long lo, long long hi) {
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    const float a = ax[i];
    if (a != a) continue;                                    // not a sink
    const float c = fcoeff * __ldg(&src4[i].w);              // kick: v += fcoeff * m_i * a   (RCBForceTree.cxx:594-596 / :615-617)
    vx[i] = fmaf(c, a, vx[i]); vy[i] = fmaf(c, ay[i], vy[i]); vz[i] = fmaf(c, az[i], vz[i]);
  }
}

// deferred kick of particles [lo, hi): waits for the velocities to have been permuted into tree order on the copy stream
int apply_kick(haccsr_ctx *c, float fcoeff, int64_t lo, int64_t hi) {
  if (hi <= lo) return 0;
  HSR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_vready, 0));
  int64_t g = (hi - lo + 255) / 256;
  if (g > (int64_t)c->sm_count * 16) g = (int64_t)c->sm_count * 16;
  k_apply_kick<<<(int)g, 256, 0, c->stream>>>(c->cur.vx, c->cur.vy, c->cur.vz, c->kick_a[0].p, c->kick_a[1].p, c->kick_a[2].p,
                                              c->src4.p, fcoeff, (long long)lo, (long long)hi);
  c->launches++;
  HSR_CUDA(cudaGetLastError());
  return 0;
}

int run_force(haccsr_ctx *c, float fcoeff, bool count_in_cutoff, haccsr_stats *st) {
  cudaStream_t s = c->stream;
  const int nn = c->n_nodes;
  HSR_TRY(c->item_cnt.ensure(nn + 1)); HSR_TRY(c->item_off.ensure(nn + 1));
  // remainder items need a packed pair kernel: the polynomial and Newton laws
  int use_rem = (c->law.kind == HACCSR_LAW_SR_POLY || c->law.kind == HACCSR_LAW_NEWTON) ? c->item_policy : 0;
  k_item_count<<<(nn + 255) / 256, 256, 0, s>>>(c->nodes.p, c->n_ranges.p, nn, use_rem, c->item_cnt.p);
  c->launches++;
  HSR_TRY(scan_exclusive(c, c->item_cnt.p, c->item_off.p, nn, c->d_counters + 10));
  HSR_CUDA(cudaMemsetAsync(c->d_counters + 11, 0, 2 * sizeof(unsigned long long), s));   // in-cutoff pairs, force-law pairs
  HSR_CUDA(cudaMemcpyAsync(c->h_counters + 10, c->d_counters + 10, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  HSR_CUDA(cudaStreamSynchronize(s));
  c->n_items = c->h_counters[10];
  if (c->n_items > 0x7fffffffll) { set_error("too many force work items"); return 1; }
  if (c->n_items == 0) return 0;
  HSR_TRY(c->items.ensure((size_t)c->n_items));
  k_item_fill<<<(nn + 255) / 256, 256, 0, s>>>(c->nodes.p, c->item_cnt.p, c->item_off.p, c->n_pseudo.p, nn, use_rem, c->items.p);
  c->launches++;
  {
    const int ni = (int)c->n_items;
    const int groups = c->force_groups, nseg = 2 * groups, nb = nseg * LPT_BINS;
    GroupBounds G;
    G.groups = groups;
    for (int g = 0; g <= MAX_GROUPS; ++g) G.lo[g] = (int)(g <= groups ? (groups > 1 ? c->group_lo[g] : (g ? c->n_tree : 0)) : c->n_tree);
    HSR_TRY(c->items_sorted.ensure((size_t)c->n_items)); HSR_TRY(c->lpt_hist.ensure(4 * (size_t)MAX_GROUPS * LPT_BINS));
    unsigned *hist = c->lpt_hist.p, *cursor = c->lpt_hist.p + 2 * MAX_GROUPS * LPT_BINS;
    HSR_CUDA(cudaMemsetAsync(hist, 0, nb * sizeof(unsigned), s));
    k_lpt_hist<<<(ni + 255) / 256, 256, 0, s>>>(c->items.p, c->list_len.p, ni, G, hist);
    c->launches++;
    HSR_TRY(scan_exclusive(c, hist, cursor, nb, nullptr));
    // first item of every segment (group x {chunk items, remainder items}), for the launch configuration
    HSR_CUDA(cudaMemcpy2DAsync(c->h_counters + 16, sizeof(int64_t), cursor, LPT_BINS * sizeof(unsigned), sizeof(unsigned),
                               nseg, cudaMemcpyDeviceToHost, s));
    k_lpt_scatter<<<(ni + 255) / 256, 256, 0, s>>>(c->items.p, c->list_len.p, ni, G, cursor, c->items_sorted.p);
    c->launches++;
    HSR_CUDA(cudaGetLastError());
    HSR_CUDA(cudaStreamSynchronize(s));
    for (int g = 0; g < nseg; ++g) c->seg_off[g] = (int64_t)(uint32_t)(c->h_counters[16 + g] & 0xffffffffll);
    c->seg_off[nseg] = ni;
  }

  ForceParams P;
  P.items = c->items_sorted.p; P.range_off = c->range_off.p; P.ranges = c->ranges.p; P.list_len = c->list_len.p;
  P.src4 = c->src4.p; P.pool = c->pool.p;
  P.vx = c->cur.vx; P.vy = c->cur.vy; P.vz = c->cur.vz;
  P.defer = 0;
  if (c->defer_kick) {
    // haccsr_kick_host: the velocities arrive (and are permuted) while the force kernel runs, so it leaves the accelerations in
    // arrays of their own and apply_kick() does v = fma(fcoeff * m, a, v) range by range -- the same operation, the same bits.
    // NaN marks "not a sink": particles of leaves outside the force box keep their velocity untouched.
    const size_t nn = (size_t)c->n_tree + 1;
    for (int q = 0; q < 3; ++q) {
      HSR_TRY(c->kick_a[q].ensure(nn));
      HSR_CUDA(cudaMemsetAsync(c->kick_a[q].p, 0xff, nn * sizeof(float), s));
    }
    P.vx = c->kick_a[0].p; P.vy = c->kick_a[1].p; P.vz = c->kick_a[2].p;
    P.defer = 1;
  }
  P.incut = c->d_counters + 11;
  for (int i = 0; i < 8; ++i) { P.a[i] = c->law.a[i]; P.b[i] = -c->law.b[i]; }
  P.rsm2 = c->law.rsm2; P.rmax2 = c->law.rmax2; P.smax = c->law.smax; P.fcoeff = fcoeff;
  P.unit_mass = c->unit_mass ? 1 : 0;
  P.tab_f = c->law_table.p; P.tab_r2 = c->law_table.p ? c->law_table.p + c->law.ntab : nullptr;
  P.tab_r2min = c->law.tab_r2min; P.tab_r2max = c->law.tab_r2max; P.tab_oodr2 = c->law.tab_oodr2; P.ntab = c->law.ntab;
  const int ni = (int)c->n_items;
  int rc;
  HSR_TRY(issue_host_out(c));     // haccsr_kick_host: all read-backs are done, the large copies can go now
  const bool guard0 = !(c->law.rsm2 > 0.0f);
  if (c->law.kind == HACCSR_LAW_NEWTON) rc = launch_force<1, 1, true, 0>(c, P, ni, count_in_cutoff);
  else if (c->law.kind == HACCSR_LAW_SR_FIT) rc = guard0 ? launch_force<1, 2, true, 0>(c, P, ni, count_in_cutoff) : launch_force<1, 2, false, 0>(c, P, ni, count_in_cutoff);
  else if (c->law.kind == HACCSR_LAW_SR_INTERP) rc = guard0 ? launch_force<1, 3, true, 0>(c, P, ni, count_in_cutoff) : launch_force<1, 3, false, 0>(c, P, ni, count_in_cutoff);
  else {
    const bool guard = !(c->law.rsm2 > 0.0f);
    const bool fused = c->arith == HACCSR_ARITH_FUSED || c->arith == HACCSR_ARITH_FUSED_RS3;
    // template code FUSED: 1 fused, 2 fused + culling, 3 fused with rsqrt(s^3), 4 that + culling
    const int code = !fused ? 0 : ((c->arith == HACCSR_ARITH_FUSED_RS3 ? 3 : 1) + (c->cull ? 1 : 0));
#define HSR_LAUNCH(CODE)                                                                                                        \
  do {                                                                                                                          \
    if (c->law.ncoef <= 6) rc = guard ? launch_force<6, 0, true, CODE>(c, P, ni, count_in_cutoff) : launch_force<6, 0, false, CODE>(c, P, ni, count_in_cutoff); \
    else rc = guard ? launch_force<7, 0, true, CODE>(c, P, ni, count_in_cutoff) : launch_force<7, 0, false, CODE>(c, P, ni, count_in_cutoff); \
  } while (0)
    if (code == 1) HSR_LAUNCH(1);
    else if (code == 2) HSR_LAUNCH(2);
    else if (code == 3) HSR_LAUNCH(3);
    else if (code == 4) HSR_LAUNCH(4);
#undef HSR_LAUNCH
    else if (c->law.ncoef <= 6) rc = guard ? launch_force<6, 0, true, 0>(c, P, ni, count_in_cutoff) : launch_force<6, 0, false, 0>(c, P, ni, count_in_cutoff);
    else rc = guard ? launch_force<7, 0, true, 0>(c, P, ni, count_in_cutoff) : launch_force<7, 0, false, 0>(c, P, ni, count_in_cutoff);
  }
  if (rc) return rc;
  if (count_in_cutoff && st) {
    HSR_CUDA(cudaMemcpyAsync(c->h_counters + 11, c->d_counters + 11, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    HSR_CUDA(cudaStreamSynchronize(s));
    st->pairs_in_cutoff = (uint64_t)c->h_counters[11];
    st->pairs_force_law = ((c->arith == HACCSR_ARITH_FUSED || c->arith == HACCSR_ARITH_FUSED_RS3) && c->cull && c->law.kind == HACCSR_LAW_SR_POLY)
                              ? (uint64_t)c->h_counters[12] : st->pairs_evaluated;
  }
  return 0;
}

}  // namespace haccsr
