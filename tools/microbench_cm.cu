// tools/microbench_cm.cu -- development probe, not part of the library.
//
// k_cm_tile (tree_build.cu: tight box + exact centroid sums of every node of a level) stand-alone, in two designs, on
// synthetic levels (node ids non-decreasing along the array, as in the build):
//   A  the library's kernel: one 2048-particle tile per block, 8 rows per warp in registers, per-node warp reductions into
//      shared-memory slots behind block barriers, slots flushed to the global accumulators once per block;
//   B  persistent warps over CONTIGUOUS particle ranges: the next rows' loads are in flight during the current rows'
//      arithmetic, the node's per-lane partial stays in registers while the node id stays the same and is reduced and
//      flushed (result-less global atomics) only when the id changes -- no shared slots, no block barriers.
// Both fill the same NodeAcc array; the probe checks that they agree exactly (box bounds and the wide integer sums) and
// prints the time per level.  profiles/r1o_microbench_stream.txt says why B is the candidate: loads alone stream at the
// HBM peak, A's 130 us per level is the serial load -> arithmetic -> flush of its blocks at 16 warps per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_cm tools/microbench_cm.cu
//   tools/microbench_cm [N = 21485038]
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct NodeAcc {
  unsigned umin[3], umax[3];
  unsigned long long lo[4], hi[4];   // total = hi * 2^32 + lo
};
static constexpr int TPB = 256, SMAX = 32;

__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void add_split(unsigned long long *lo, unsigned long long *hi, long long v) {
  if (v == 0) return;
  atomicAdd(lo, (unsigned long long)(v & 0xffffffffll));
  atomicAdd(hi, (unsigned long long)(v >> 32));
}
struct Part { unsigned umin[3], umax[3]; long long s[4]; };
__device__ __forceinline__ void part_reset(Part &p) {
  p.umin[0] = p.umin[1] = p.umin[2] = 0xffffffffu; p.umax[0] = p.umax[1] = p.umax[2] = 0u;
  p.s[0] = p.s[1] = p.s[2] = p.s[3] = 0;
}
__device__ __forceinline__ void part_add(Part &p, const float4 &r, float sx, float sm) {
  unsigned ex = enc_f(r.x), ey = enc_f(r.y), ez = enc_f(r.z);
  p.umin[0] = min(p.umin[0], ex); p.umax[0] = max(p.umax[0], ex);
  p.umin[1] = min(p.umin[1], ey); p.umax[1] = max(p.umax[1], ey);
  p.umin[2] = min(p.umin[2], ez); p.umax[2] = max(p.umax[2], ez);
  p.s[0] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.x), sx));
  p.s[1] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.y), sx));
  p.s[2] += __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.z), sx));
  p.s[3] += __float2ll_rn(__fmul_rn(r.w, sm));
}
__device__ __forceinline__ long long shfl_xor_ll(long long v, int o) {
  int lo = __shfl_xor_sync(0xffffffffu, (int)(unsigned)(v & 0xffffffffll), o);
  int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), o);
  return ((long long)hi << 32) | (long long)(unsigned)lo;
}
__device__ __forceinline__ void warp_reduce(Part &p) {
#pragma unroll
  for (int q = 0; q < 3; ++q) { p.umin[q] = __reduce_min_sync(0xffffffffu, p.umin[q]); p.umax[q] = __reduce_max_sync(0xffffffffu, p.umax[q]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q) p.s[q] += shfl_xor_ll(p.s[q], o);
  }
}
__device__ __forceinline__ void flush_global(const Part &p, NodeAcc &A) {
  for (int k = 0; k < 3; ++k) { atomicMin(&A.umin[k], p.umin[k]); atomicMax(&A.umax[k], p.umax[k]); }
  for (int k = 0; k < 4; ++k) add_split(&A.lo[k], &A.hi[k], p.s[k]);
}

// ---- A: the library's k_cm_tile ----------------------------------------------------------------------------------
struct Slot { unsigned umin[3], umax[3]; unsigned used, pad; unsigned long long s[4]; };
__device__ __forceinline__ void flush_part(const Part &p, int nd, int n0, Slot *slots, NodeAcc *acc) {
  int sl = nd - n0;
  if (sl >= 0 && sl < SMAX) {
    Slot &S = slots[sl];
    for (int k = 0; k < 3; ++k) { atomicMin(&S.umin[k], p.umin[k]); atomicMax(&S.umax[k], p.umax[k]); }
    for (int k = 0; k < 4; ++k) atomicAdd(&S.s[k], (unsigned long long)p.s[k]);
    S.used = 1;
  } else flush_global(p, acc[nd]);
}
static constexpr int CM_IPT = 8, CM_TILE = TPB * CM_IPT;
__global__ void __launch_bounds__(TPB) k_cm_block(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                  NodeAcc *__restrict__ acc, float sx, float sm) {
  __shared__ int s_n0;
  __shared__ Slot slots[SMAX];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int wbase = blockIdx.x * CM_TILE + w * (32 * CM_IPT) + lane;
  int nd[CM_IPT]; float4 r[CM_IPT];
#pragma unroll
  for (int k = 0; k < CM_IPT; ++k) { const int i = wbase + 32 * k; nd[k] = (i < n) ? __ldcs(nid + i) : -1; }
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
  mn = __reduce_min_sync(0xffffffffu, mn);
  if (lane == 0 && mn != INT_MAX) atomicMin(&s_n0, mn);
  __syncthreads();
  const int n0 = s_n0;
  if (n0 == INT_MAX) return;
  int key = mn;
  while (key != INT_MAX) {
    Part p; part_reset(p);
    int next = INT_MAX;
#pragma unroll
    for (int k = 0; k < CM_IPT; ++k) {
      if (nd[k] == key) part_add(p, r[k], sx, sm);
      else if (nd[k] > key) next = min(next, nd[k]);
    }
    warp_reduce(p);
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

// ---- B: persistent warps over contiguous ranges ----------------------------------------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(TPB) k_cm_warp(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                 int per_warp, NodeAcc *__restrict__ acc, float sx, float sm) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long b64 = gw * (long long)per_warp;
  if (b64 >= n) return;
  const int begin = (int)b64, end = (int)min((long long)n, b64 + per_warp);
  constexpr int STEP = 32 * ROWS;
  int nd[ROWS], nd1[ROWS]; float4 r[ROWS], r1[ROWS];
  auto load = [&](int pos, int (&d)[ROWS], float4 (&q)[ROWS]) {
#pragma unroll
    for (int k = 0; k < ROWS; ++k) { const int i = pos + 32 * k + lane; d[k] = (i < end) ? __ldcs(nid + i) : -1; }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) if (d[k] >= 0) q[k] = __ldcs(rec + pos + 32 * k + lane);
  };
  load(begin, nd, r);
  int K = -1;                 // node whose per-lane partial is carried in p
  Part p; part_reset(p);
  for (int pos = begin; pos < end; pos += STEP) {
    load(pos + STEP, nd1, r1);         // rows past `end` load nothing
    int mn = INT_MAX;
#pragma unroll
    for (int k = 0; k < ROWS; ++k) if (nd[k] >= 0) mn = min(mn, nd[k]);
    int cur = __reduce_min_sync(0xffffffffu, mn);
    while (cur != INT_MAX) {           // warp-uniform: the distinct nodes of this step, increasing
      if (cur != K) {
        if (K >= 0) { warp_reduce(p); if (lane == 0) flush_global(p, acc[K]); }
        part_reset(p); K = cur;
      }
      int next = INT_MAX;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        if (nd[k] == cur) part_add(p, r[k], sx, sm);
        else if (nd[k] > cur) next = min(next, nd[k]);
      }
      cur = __reduce_min_sync(0xffffffffu, next);
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) { nd[k] = nd1[k]; r[k] = r1[k]; }
  }
  if (K >= 0) { warp_reduce(p); if (lane == 0) flush_global(p, acc[K]); }
}

static void make_level(int n, int run_lo, int run_hi, double finished_frac, int *hid, int *n_nodes) {
  int i = 0, node = 0;
  while (i < n) {
    int len = run_lo + (int)((run_hi - run_lo) * drand48());
    if (len > n - i) len = n - i;
    const bool fin = drand48() < finished_frac;      // a finished leaf: its particles carry -1
    for (int k = 0; k < len; ++k) hid[i + k] = fin ? -1 : node;
    if (!fin) node++;
    i += len;
  }
  *n_nodes = node > 0 ? node : 1;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 21485038;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs; N = %d particles\n", prop.name, prop.multiProcessorCount, n);
  float4 *h = (float4 *)malloc((size_t)n * sizeof(float4));
  int *hid = (int *)malloc((size_t)n * sizeof(int));
  srand48(11);
  for (int i = 0; i < n; ++i) h[i] = make_float4((float)(278.0 * drand48()), (float)(278.0 * drand48()), (float)(278.0 * drand48()), 1.0f);
  float4 *rec; int *nid; NodeAcc *accA, *accB;
  CK(cudaMalloc(&rec, (size_t)n * sizeof(float4))); CK(cudaMalloc(&nid, (size_t)n * sizeof(int)));
  CK(cudaMemcpy(rec, h, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
  const float sx = ldexpf(1.0f, 50 - 9 - 1), sm = ldexpf(1.0f, 50 - 1);     // as k_root_init picks them for |x| < 512, m = 1
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct { const char *name; int lo, hi; double fin; } levels[] = {
      {"root (one node)", n, n + 1, 0.0}, {"level ~6 (runs of ~330 k)", 300000, 360000, 0.0},
      {"level ~12 (runs of ~5 k)", 4000, 6500, 0.0}, {"deepest (runs of 256-400)", 256, 400, 0.0},
      {"deepest, 30 % finished leaves", 256, 400, 0.3}};
  for (auto &L : levels) {
    int nn = 0;
    make_level(n, L.lo, L.hi, L.fin, hid, &nn);
    CK(cudaMemcpy(nid, hid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    NodeAcc *init = (NodeAcc *)malloc((size_t)nn * sizeof(NodeAcc));
    for (int k = 0; k < nn; ++k) { for (int q = 0; q < 3; ++q) { init[k].umin[q] = 0xffffffffu; init[k].umax[q] = 0; } for (int q = 0; q < 4; ++q) init[k].lo[q] = init[k].hi[q] = 0; }
    CK(cudaMalloc(&accA, (size_t)nn * sizeof(NodeAcc))); CK(cudaMalloc(&accB, (size_t)nn * sizeof(NodeAcc)));
    auto time_it = [&](NodeAcc *acc, auto launch) {
      float best = 1e30f;
      for (int it = 0; it < 5; ++it) {
        CK(cudaMemcpy(acc, init, (size_t)nn * sizeof(NodeAcc), cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        launch(acc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it && ms < best) best = ms;
      }
      return best * 1e3f;
    };
    const float tA = time_it(accA, [&](NodeAcc *a) { k_cm_block<<<(n + CM_TILE - 1) / CM_TILE, TPB>>>(rec, nid, n, a, sx, sm); });
    NodeAcc *ra = (NodeAcc *)malloc((size_t)nn * sizeof(NodeAcc)), *rb = (NodeAcc *)malloc((size_t)nn * sizeof(NodeAcc));
    CK(cudaMemcpy(ra, accA, (size_t)nn * sizeof(NodeAcc), cudaMemcpyDeviceToHost));
    printf("%-34s %7d nodes   A block/tile %7.1f us |", L.name, nn, tA);
    auto run_b = [&](auto kern, int rows, const char *tag) {
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, 0));
      const long long warps = (long long)prop.multiProcessorCount * occ * (TPB / 32);
      const int step = 32 * rows;
      int per_warp = (int)(((long long)n + warps - 1) / warps);
      per_warp = (per_warp + step - 1) / step * step;
      const int blocks = (int)(((long long)n + (long long)per_warp * (TPB / 32) - 1) / ((long long)per_warp * (TPB / 32)));
      const float tB = time_it(accB, [&](NodeAcc *a) { kern<<<blocks, TPB>>>(rec, nid, n, per_warp, a, sx, sm); });
      CK(cudaMemcpy(rb, accB, (size_t)nn * sizeof(NodeAcc), cudaMemcpyDeviceToHost));
      long long bad = 0;
      for (int k = 0; k < nn; ++k) {
        for (int q = 0; q < 3; ++q) bad += (ra[k].umin[q] != rb[k].umin[q]) + (ra[k].umax[q] != rb[k].umax[q]);
        for (int q = 0; q < 4; ++q) {
          const __int128 ta = ((__int128)(long long)ra[k].hi[q] << 32) + (__int128)ra[k].lo[q];
          const __int128 tb = ((__int128)(long long)rb[k].hi[q] << 32) + (__int128)rb[k].lo[q];
          bad += ta != tb;
        }
      }
      printf("  B %s (%d/SM) %7.1f us %s |", tag, occ, tB, bad ? "MISMATCH" : "same");
    };
    run_b(k_cm_warp<2>, 2, "2 rows");
    run_b(k_cm_warp<4>, 4, "4 rows");
    run_b(k_cm_warp<8>, 8, "8 rows");
    printf("\n");
    CK(cudaFree(accA)); CK(cudaFree(accB)); free(init); free(ra); free(rb);
  }
  return 0;
}
