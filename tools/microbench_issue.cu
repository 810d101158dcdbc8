// tools/microbench_issue.cu -- co-issue rules of the sm_100 FP32 paths, measured (not part of the product).
//
// Each kernel runs a register-only loop of independent chains: NF FMA-pipe instructions (scalar FFMA or
// packed FFMA2), NA ALU-pipe instructions (FSETP+FSEL pairs count as 2) and NM MUFU.RSQ per iteration, 8 warps
// per SMSP resident.  Reports cycles per iteration per warp-scheduler so one can read off whether FFMA2 costs
// one or two issue cycles and whether ALU / XU instructions hide behind it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_issue microbench_issue.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool PACKED, int NF, int NA, int NM>
__global__ void __launch_bounds__(256) k_mix(float *out, float a, float b, float thr, int iters, long long *cyc) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  float sel[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sel[i] = 0.f;
  float mu[4] = {1.5f, 2.5f, 3.5f, 4.5f};
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < NF; ++u) {
      if (PACKED) acc[u & 7] = __ffma2_rn(acc[u & 7], a2, b2);
      else { acc[u & 7].x = __fmaf_rn(acc[u & 7].x, a, b); }
    }
    if (!PACKED) {   // same number of lane-FMAs as the packed form
#pragma unroll
      for (int u = 0; u < NF; ++u) acc[u & 7].y = __fmaf_rn(acc[u & 7].y, a, b);
    }
#pragma unroll
    for (int u = 0; u < NA; ++u) sel[u & 7] = (acc[u & 7].x < thr) ? acc[(u + 1) & 7].y : sel[u & 7];   // FSETP + FSEL
#pragma unroll
    for (int u = 0; u < NM; ++u) mu[u & 3] = rsqrt_ftz(mu[u & 3]);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + sel[i];
  s += mu[0] + mu[1] + mu[2] + mu[3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

template <bool PACKED, int NF, int NA, int NM>
static void run(const char *name, int nsm, float *out, long long *d_cyc) {
  const int iters = 20000;
  const int grid = nsm * 4;     // 4 CTAs x 8 warps = 32 warps per SM = 8 per scheduler
  k_mix<PACKED, NF, NA, NM><<<grid, 256>>>(out, 0.999f, 0.001f, 0.5f, 100, d_cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_mix<PACKED, NF, NA, NM><<<grid, 256>>>(out, 0.999f, 0.001f, 0.5f, iters, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long cyc; CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  // whole-kernel view: every scheduler (4 per SM) serves grid*8/(4*nsm) warps, each running `iters` iterations
  const double warps_per_sched = (double)grid * 8.0 / (4.0 * nsm);
  const double sched_cycles = ms * 1e-3 * 1965e6;                  // at the 1965 MHz the sampler reports under load
  const double per_warp_iter = sched_cycles / (iters * warps_per_sched);
  int lane_fma = 2 * NF;                                           // FMA-pipe lane-cycles per warp-iteration
  int instr = (PACKED ? NF : 2 * NF) + 2 * NA + NM;
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mix<PACKED, NF, NA, NM>, 256, 0));
  printf("%-34s FMAlane=%2d instr=%2d : %.2f cycles/warp-iter (%.2f per lane-FMA-cycle, %.2f per instr) block0 clock %.0f MHz occ %d\n", name, lane_fma, instr,
         per_warp_iter, per_warp_iter / lane_fma, per_warp_iter / instr, (double)cyc / (ms * 1e3), occ);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  float *out; long long *d_cyc;
  CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 4 * 256 * sizeof(float)));
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  const int n = prop.multiProcessorCount;
#define R(P, NF, NA, NM) run<P, NF, NA, NM>(#P " NF=" #NF " NA=" #NA " NM=" #NM, n, out, d_cyc)
  R(false, 16, 0, 0); R(true, 16, 0, 0);
  R(false, 8, 0, 0); R(true, 8, 0, 0);
  R(false, 8, 1, 0); R(true, 8, 1, 0);
  R(false, 8, 2, 0); R(true, 8, 2, 0);
  R(false, 8, 4, 0); R(true, 8, 4, 0);
  R(false, 8, 0, 1); R(true, 8, 0, 1);
  R(false, 8, 0, 2); R(true, 8, 0, 2);
  R(false, 8, 2, 1); R(true, 8, 2, 1);
  R(false, 10, 2, 1); R(true, 10, 2, 1);
  R(true, 10, 4, 2);
  return 0;
}
