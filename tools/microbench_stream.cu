// tools/microbench_stream.cu -- development probe, not part of the library.
//
// Question ("candidates not tried"): the tree build's tile kernels read two streams per level, a
// float4 record and an int node id per particle (20 B), do a little arithmetic and leave; k_cm_tile reaches 35 % of the
// DRAM peak with its loads held in registers (128 registers, two blocks per SM).  How fast can the same streams be
// consumed when they arrive in a shared-memory ring by 1-D TMA bulk copies (cp.async.bulk + mbarrier, the pattern of
// force.cu's producer) from persistent blocks, against (a) one tile per block with direct loads and (b) persistent blocks
// with a register prefetch?  Every variant computes the same checksum (sum of x*w over particles with id >= 0, as exact
// int64 fixed point), so a broken pipeline shows up as a wrong number, and reports GB/s = 20 B * N / time.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_stream tools/microbench_stream.cu
//   tools/microbench_stream [N = 21485038]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static constexpr int TPB = 256;

__device__ __forceinline__ long long contrib(const float4 r, int id) {
  return id >= 0 ? __float2ll_rn(__fmul_rn(__fmul_rn(r.w, r.x), 1048576.0f)) : 0ll;
}
__device__ __forceinline__ void block_sum_to(long long v, unsigned long long *out) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ long long s_w[TPB / 32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < TPB / 32; ++w) t += s_w[w];
    atomicAdd(out, (unsigned long long)t);
  }
  __syncthreads();
}

// (a) one tile of TPB * IPT particles per block, all loads issued up front
template <int IPT>
__global__ void __launch_bounds__(TPB) k_direct(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                unsigned long long *out) {
  const int base = blockIdx.x * TPB * IPT + threadIdx.x;
  int id[IPT]; float4 r[IPT];
#pragma unroll
  for (int k = 0; k < IPT; ++k) { const int i = base + k * TPB; id[k] = i < n ? __ldcs(nid + i) : -1; }
#pragma unroll
  for (int k = 0; k < IPT; ++k) if (id[k] >= 0) r[k] = __ldcs(rec + base + k * TPB);
  long long v = 0;
#pragma unroll
  for (int k = 0; k < IPT; ++k) if (id[k] >= 0) v += contrib(r[k], id[k]);
  block_sum_to(v, out);
}

// (b) persistent blocks, tiles of TPB * IPT particles, ids two tiles ahead and records one tile ahead in registers
template <int IPT>
__global__ void __launch_bounds__(TPB) k_regpipe(const float4 *__restrict__ rec, const int *__restrict__ nid, int n,
                                                 int ntiles, unsigned long long *out) {
  auto load_id = [&](int tile, int (&id)[IPT]) {
#pragma unroll
    for (int k = 0; k < IPT; ++k) { const int i = tile * TPB * IPT + k * TPB + threadIdx.x; id[k] = (tile < ntiles && i < n) ? __ldcs(nid + i) : -1; }
  };
  auto load_rec = [&](int tile, const int (&id)[IPT], float4 (&r)[IPT]) {
#pragma unroll
    for (int k = 0; k < IPT; ++k) if (id[k] >= 0) r[k] = __ldcs(rec + tile * TPB * IPT + k * TPB + threadIdx.x);
  };
  int id0[IPT], id1[IPT], id2[IPT]; float4 r0[IPT], r1[IPT];
  int tile = blockIdx.x;
  const int stride = gridDim.x;
  load_id(tile, id0); load_id(tile + stride, id1); load_rec(tile, id0, r0);
  long long v = 0;
  for (; tile < ntiles; tile += stride) {
    load_id(tile + 2 * stride, id2);
    load_rec(tile + stride, id1, r1);
#pragma unroll
    for (int k = 0; k < IPT; ++k) if (id0[k] >= 0) v += contrib(r0[k], id0[k]);
#pragma unroll
    for (int k = 0; k < IPT; ++k) { id0[k] = id1[k]; id1[k] = id2[k]; r0[k] = r1[k]; }
  }
  block_sum_to(v, out);
}

// (c) persistent blocks, shared-memory ring of STAGES tiles of TILE particles filled by 1-D TMA bulk copies
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int TILE, int STAGES>
__global__ void __launch_bounds__(TPB) k_tma(const float4 *__restrict__ rec, const int *__restrict__ nid, int n, int ntiles,
                                             unsigned long long *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *s_rec = reinterpret_cast<float4 *>(smem);                                   // [STAGES][TILE]
  int *s_id = reinterpret_cast<int *>(smem + (size_t)STAGES * TILE * sizeof(float4));  // [STAGES][TILE]
  __shared__ __align__(8) unsigned long long bars[STAGES];
  const int stride = gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto produce = [&](int tile, int stage) {      // thread 0 only; whole tiles are 16-byte multiples, the last one is trimmed
    if (tile >= ntiles) return;
    const int cnt = min(TILE, n - tile * TILE);
    const int cnt4 = cnt & ~3;                  // bulk copies move multiples of 16 bytes: ids in groups of four
    const unsigned bar = smem_u32(&bars[stage]);
    mbar_expect_tx(bar, (unsigned)cnt * 16u + (unsigned)cnt4 * 4u);
    bulk_g2s(smem_u32(s_rec + (size_t)stage * TILE), rec + (size_t)tile * TILE, (unsigned)cnt * 16u, bar);
    if (cnt4) bulk_g2s(smem_u32(s_id + (size_t)stage * TILE), nid + (size_t)tile * TILE, (unsigned)cnt4 * 4u, bar);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) produce(blockIdx.x + s * stride, s);
  long long v = 0;
  int k = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += stride, ++k) {
    const int stage = k % STAGES;
    mbar_wait(smem_u32(&bars[stage]), (unsigned)(k / STAGES) & 1u);
    const int cnt = min(TILE, n - tile * TILE), cnt4 = cnt & ~3;
    const float4 *tr = s_rec + (size_t)stage * TILE;
    const int *ti = s_id + (size_t)stage * TILE;
#pragma unroll
    for (int j = 0; j < TILE / TPB; ++j) {
      const int q = j * TPB + threadIdx.x;
      if (q < cnt) {
        const int id = q < cnt4 ? ti[q] : __ldg(nid + (size_t)tile * TILE + q);   // the trimmed tail (< 4 ids) comes directly
        v += contrib(tr[q], id);
      }
    }
    __syncthreads();                             // every thread is done with this stage
    if (threadIdx.x == 0) produce(tile + STAGES * stride, stage);
  }
  block_sum_to(v, out);
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 21485038;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs; N = %d particles, 20 B each per pass\n", prop.name, prop.multiProcessorCount, n);
  float4 *h = (float4 *)malloc((size_t)n * sizeof(float4));
  int *hid = (int *)malloc((size_t)n * sizeof(int));
  srand48(3);
  long long expect = 0;
  for (int i = 0; i < n; ++i) {
    h[i] = make_float4((float)(278.0 * drand48()), (float)drand48(), (float)drand48(), 1.0f);
    hid[i] = (i % 97 == 0) ? -1 : i / 300;
    if (hid[i] >= 0) expect += llrintf(h[i].w * h[i].x * 1048576.0f);
  }
  float4 *rec; int *nid; unsigned long long *out;
  CK(cudaMalloc(&rec, (size_t)n * sizeof(float4))); CK(cudaMalloc(&nid, (size_t)n * sizeof(int))); CK(cudaMalloc(&out, 8));
  CK(cudaMemcpy(rec, h, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(nid, hid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char *name, auto launch) {
    float best = 1e30f; unsigned long long got = 0;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemset(out, 0, 8));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it && ms < best) best = ms;
      CK(cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost));
    }
    printf("%-44s %8.1f us  %7.0f GB/s  %s\n", name, best * 1e3, 20.0 * n / (best * 1e-3) / 1e9,
           (long long)got == expect ? "checksum ok" : "CHECKSUM WRONG");
  };
  const int sm = prop.multiProcessorCount;
  run("direct, 1024 per block", [&] { k_direct<4><<<(n + 1023) / 1024, TPB>>>(rec, nid, n, out); });
  run("direct, 2048 per block", [&] { k_direct<8><<<(n + 2047) / 2048, TPB>>>(rec, nid, n, out); });
  run("direct, 4096 per block", [&] { k_direct<16><<<(n + 4095) / 4096, TPB>>>(rec, nid, n, out); });
  {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_regpipe<4>, TPB, 0));
    const int nt = (n + 1023) / 1024;
    run("register pipeline, 1024-particle tiles", [&] { k_regpipe<4><<<sm * occ, TPB>>>(rec, nid, n, nt, out); });
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_regpipe<8>, TPB, 0));
    const int nt2 = (n + 2047) / 2048;
    run("register pipeline, 2048-particle tiles", [&] { k_regpipe<8><<<sm * occ, TPB>>>(rec, nid, n, nt2, out); });
  }
#define RUN_TMA(TILE, STAGES)                                                                                          \
  {                                                                                                                    \
    const size_t sh = (size_t)STAGES * TILE * 20;                                                                      \
    CK(cudaFuncSetAttribute(k_tma<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));               \
    int occ = 0;                                                                                                       \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma<TILE, STAGES>, TPB, sh));                             \
    const int nt = (n + TILE - 1) / TILE;                                                                              \
    char name[96];                                                                                                     \
    snprintf(name, sizeof name, "TMA ring, %d-particle tiles x %d stages, %d/SM", TILE, STAGES, occ);                  \
    run(name, [&] { k_tma<TILE, STAGES><<<sm * occ, TPB, sh>>>(rec, nid, n, nt, out); });                              \
  }
  RUN_TMA(1024, 2) RUN_TMA(1024, 3) RUN_TMA(1024, 4) RUN_TMA(2048, 2) RUN_TMA(2048, 3) RUN_TMA(2048, 4) RUN_TMA(4096, 2)
  return 0;
}
