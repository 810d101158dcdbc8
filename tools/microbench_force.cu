// tools/microbench_force.cu -- inner-loop variants of the short-range pair kernel, timed in isolation.
//
// Not part of the product: a stand-alone probe used to choose the formulation of k_force (force.cu).  Every
// variant evaluates the same poly5 pair (30 flop) for S sinks per thread against a
// shared-memory tile of sources that is re-read NREP times, so the measurement is the issue / FMA-pipe
// limit of the formulation and nothing else.  Prints pairs/s and the fraction of the FP32 peak.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_force microbench_force.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Law { float a[6]; float rsm2, rmax2; float b[6]; float smax; };

__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

static constexpr int TILE = 128;

// ---- V0: scalar, r2 without contraction (bit-identical cutoff set to the CPU reference) ---------------
template <int S, bool EXACT>
__device__ __forceinline__ void pair_scalar(const float4 s, const float (&xi)[S], const float (&yi)[S], const float (&zi)[S],
                                            float (&ax)[S], float (&ay)[S], float (&az)[S], const Law &L) {
#pragma unroll
  for (int k = 0; k < S; ++k) {
    float dx = __fsub_rn(s.x, xi[k]), dy = __fsub_rn(s.y, yi[k]), dz = __fsub_rn(s.z, zi[k]);
    float r2;
    if (EXACT) r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    else r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float t = r2 + L.rsm2;
    float f = rsqrt_ftz(t * t * t);
    float p = L.a[5];
#pragma unroll
    for (int q = 4; q >= 0; --q) p = fmaf(p, r2, L.a[q]);
    f -= p;
    f *= s.w;
    f = (r2 < L.rmax2) ? f : 0.0f;
    ax[k] = fmaf(f, dx, ax[k]); ay[k] = fmaf(f, dy, ay[k]); az[k] = fmaf(f, dz, az[k]);
  }
}

template <int S, bool EXACT>
__global__ void __launch_bounds__(32) k_scalar(const float4 *__restrict__ src, float4 *__restrict__ out, Law L, int nrep) {
  __shared__ float4 tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += 32) tile[i] = src[(blockIdx.x * TILE + i) & 0xffff];
  __syncwarp();
  float xi[S], yi[S], zi[S], ax[S], ay[S], az[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    float4 s = src[(blockIdx.x * 32 * S + k * 32 + threadIdx.x) & 0xffff];
    xi[k] = s.x; yi[k] = s.y; zi[k] = s.z; ax[k] = ay[k] = az[k] = 0.f;
  }
  for (int r = 0; r < nrep; ++r) {
#pragma unroll 4
    for (int j = 0; j < TILE; ++j) pair_scalar<S, EXACT>(tile[j], xi, yi, zi, ax, ay, az, L);
  }
  float sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int k = 0; k < S; ++k) { sx += ax[k]; sy += ay[k]; sz += az[k]; }
  out[blockIdx.x * 32 + threadIdx.x] = make_float4(sx, sy, sz, 0.f);
}

// ---- V2: packed f32x2, two SINKS per instruction; the source is duplicated into register pairs --------
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not for the scalar forms), so the
// unfused sum of squares needs an operand ptxas cannot see through: XOR with a run-time zero (ALU pipe).
__device__ __forceinline__ float2 opaque2(float2 v, unsigned zero) {
  return make_float2(__uint_as_float(__float_as_uint(v.x) ^ zero), __uint_as_float(__float_as_uint(v.y) ^ zero));
}
template <int S2, int EXACT>   // S2 = sink pairs per thread
__device__ __forceinline__ void pair_sink2(const float4 s, const float2 (&xi)[S2], const float2 (&yi)[S2], const float2 (&zi)[S2],
                                           float2 (&ax)[S2], float2 (&ay)[S2], float2 (&az)[S2], const Law &L, unsigned zero) {
  const float2 sx = make_float2(s.x, s.x), sy = make_float2(s.y, s.y), sz = make_float2(s.z, s.z), sw = make_float2(s.w, s.w);
  const float2 rsm2 = make_float2(L.rsm2, L.rsm2);
#pragma unroll
  for (int k = 0; k < S2; ++k) {
    // d = s - xi  as  s + (-xi): the sinks are stored negated
    float2 dx = __fadd2_rn(sx, xi[k]), dy = __fadd2_rn(sy, yi[k]), dz = __fadd2_rn(sz, zi[k]);
    float2 r2;
    if (EXACT == 3) {   // squares as scalar mul.rn (never contracted by ptxas), sums packed
      float2 qx = make_float2(__fmul_rn(dx.x, dx.x), __fmul_rn(dx.y, dx.y));
      float2 qy = make_float2(__fmul_rn(dy.x, dy.x), __fmul_rn(dy.y, dy.y));
      float2 qz = make_float2(__fmul_rn(dz.x, dz.x), __fmul_rn(dz.y, dz.y));
      r2 = __fadd2_rn(__fadd2_rn(qx, qy), qz);
    } else if (EXACT == 2) r2 = __fadd2_rn(__fadd2_rn(opaque2(__fmul2_rn(dx, dx), zero), opaque2(__fmul2_rn(dy, dy), zero)), opaque2(__fmul2_rn(dz, dz), zero));
    else if (EXACT == 1) r2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
    else r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    float2 t = __fadd2_rn(r2, rsm2);
    float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
    float2 f = make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y));
    float2 p = make_float2(L.a[5], L.a[5]);
#pragma unroll
    for (int q = 4; q >= 0; --q) p = __ffma2_rn(p, r2, make_float2(L.a[q], L.a[q]));
    // f = (f - p) * w  ->  w*f - w*p : one FMUL2 + one FFMA2 either way; keep (f - p) * w
    f = __fmul2_rn(__fadd2_rn(f, make_float2(-p.x, -p.y)), sw);
    bool ina = r2.x < L.rmax2, inb = r2.y < L.rmax2;
    if (EXACT == 4) {
      // fused r2 is within 2 ulp of the unfused value: only pairs whose r2 lies within +-4 ulp of rmax2 can
      // differ in the cutoff decision; detect them with integer compares on the float bits and redo the
      // decision exactly (scalar mul.rn / add.rn) in a branch that is almost never taken
      const unsigned lo = __float_as_uint(L.rmax2) - 4u;
      bool amb = (__float_as_uint(r2.x) - lo) <= 8u;
      amb = amb || ((__float_as_uint(r2.y) - lo) <= 8u);
      if (__any_sync(0xffffffffu, amb)) {
        float ea = __fadd_rn(__fadd_rn(__fmul_rn(dx.x, dx.x), __fmul_rn(dy.x, dy.x)), __fmul_rn(dz.x, dz.x));
        float eb = __fadd_rn(__fadd_rn(__fmul_rn(dx.y, dx.y), __fmul_rn(dy.y, dy.y)), __fmul_rn(dz.y, dz.y));
        ina = ea < L.rmax2; inb = eb < L.rmax2;
      }
    }
    f.x = ina ? f.x : 0.0f;
    f.y = inb ? f.y : 0.0f;
    ax[k] = __ffma2_rn(f, dx, ax[k]); ay[k] = __ffma2_rn(f, dy, ay[k]); az[k] = __ffma2_rn(f, dz, az[k]);
  }
}

template <int S2, int EXACT, int UNROLL>
__global__ void __launch_bounds__(32) k_sink2(const float4 *__restrict__ src, float4 *__restrict__ out, Law L, int nrep, unsigned zero) {
  __shared__ float4 tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += 32) tile[i] = src[(blockIdx.x * TILE + i) & 0xffff];
  __syncwarp();
  float2 xi[S2], yi[S2], zi[S2], ax[S2], ay[S2], az[S2];
#pragma unroll
  for (int k = 0; k < S2; ++k) {
    float4 a = src[(blockIdx.x * 64 * S2 + (2 * k) * 32 + threadIdx.x) & 0xffff];
    float4 b = src[(blockIdx.x * 64 * S2 + (2 * k + 1) * 32 + threadIdx.x) & 0xffff];
    xi[k] = make_float2(-a.x, -b.x); yi[k] = make_float2(-a.y, -b.y); zi[k] = make_float2(-a.z, -b.z);
    ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
  }
  for (int r = 0; r < nrep; ++r) {
#pragma unroll UNROLL
    for (int j = 0; j < TILE; ++j) pair_sink2<S2, EXACT>(tile[j], xi, yi, zi, ax, ay, az, L, zero);
  }
  float sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int k = 0; k < S2; ++k) { sx += ax[k].x + ax[k].y; sy += ay[k].x + ay[k].y; sz += az[k].x + az[k].y; }
  out[blockIdx.x * 32 + threadIdx.x] = make_float4(sx, sy, sz, 0.f);
}

// ---- V3: packed f32x2, two SOURCES per instruction; the tile is stored pair-interleaved ------------------
// tile2[j] = (x0,x1,y0,y1), tile2[j + TILE/2] = (z0,z1,m0,m1) for sources 2j, 2j+1
template <int S, bool EXACT>
__device__ __forceinline__ void pair_src2(const float4 A, const float4 B, const float2 (&xi)[S], const float2 (&yi)[S], const float2 (&zi)[S],
                                          float2 (&ax)[S], float2 (&ay)[S], float2 (&az)[S], const Law &L) {
  const float2 sx = make_float2(A.x, A.y), sy = make_float2(A.z, A.w), sz = make_float2(B.x, B.y), sw = make_float2(B.z, B.w);
  const float2 rsm2 = make_float2(L.rsm2, L.rsm2);
#pragma unroll
  for (int k = 0; k < S; ++k) {
    float2 dx = __fadd2_rn(sx, xi[k]), dy = __fadd2_rn(sy, yi[k]), dz = __fadd2_rn(sz, zi[k]);
    float2 r2;
    if (EXACT) r2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
    else r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    float2 t = __fadd2_rn(r2, rsm2);
    float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
    float2 f = make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y));
    float2 p = make_float2(L.a[5], L.a[5]);
#pragma unroll
    for (int q = 4; q >= 0; --q) p = __ffma2_rn(p, r2, make_float2(L.a[q], L.a[q]));
    f = __fmul2_rn(__fadd2_rn(f, make_float2(-p.x, -p.y)), sw);
    f.x = (r2.x < L.rmax2) ? f.x : 0.0f;
    f.y = (r2.y < L.rmax2) ? f.y : 0.0f;
    ax[k] = __ffma2_rn(f, dx, ax[k]); ay[k] = __ffma2_rn(f, dy, ay[k]); az[k] = __ffma2_rn(f, dz, az[k]);
  }
}

template <int S, bool EXACT>
__global__ void __launch_bounds__(32) k_src2(const float4 *__restrict__ src, float4 *__restrict__ out, Law L, int nrep) {
  __shared__ float4 tile[TILE];
  for (int i = threadIdx.x; i < TILE / 2; i += 32) {
    float4 a = src[(blockIdx.x * TILE + 2 * i) & 0xffff], b = src[(blockIdx.x * TILE + 2 * i + 1) & 0xffff];
    tile[i] = make_float4(a.x, b.x, a.y, b.y);
    tile[i + TILE / 2] = make_float4(a.z, b.z, a.w, b.w);
  }
  __syncwarp();
  float2 xi[S], yi[S], zi[S], ax[S], ay[S], az[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    float4 a = src[(blockIdx.x * 32 * S + k * 32 + threadIdx.x) & 0xffff];
    xi[k] = make_float2(-a.x, -a.x); yi[k] = make_float2(-a.y, -a.y); zi[k] = make_float2(-a.z, -a.z);
    ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
  }
  for (int r = 0; r < nrep; ++r) {
#pragma unroll 2
    for (int j = 0; j < TILE / 2; ++j) pair_src2<S, EXACT>(tile[j], tile[j + TILE / 2], xi, yi, zi, ax, ay, az, L);
  }
  float sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int k = 0; k < S; ++k) { sx += ax[k].x + ax[k].y; sy += ay[k].x + ay[k].y; sz += az[k].x + az[k].y; }
  out[blockIdx.x * 32 + threadIdx.x] = make_float4(sx, sy, sz, 0.f);
}

// ---- V4: fused arithmetic (HACCSR_ARITH_FUSED): s-seeded FMA chain, polynomial and cutoff in s, unit masses ----
// MODE 5: f = rsqrt(s^3) - p(s)            (17 FMA-pipe operations per pair)
// MODE 6: f = fma(rs*rs, rs, -p(s)), rs = rsqrt(s), coefficients negated on the host   (16)
template <int S2, int MODE>
__device__ __forceinline__ void pair_fused(const float4 s, const float2 (&xi)[S2], const float2 (&yi)[S2], const float2 (&zi)[S2],
                                           float2 (&ax)[S2], float2 (&ay)[S2], float2 (&az)[S2], const Law &L) {
  const float2 sx = make_float2(s.x, s.x), sy = make_float2(s.y, s.y), sz = make_float2(s.z, s.z);
#pragma unroll
  for (int k = 0; k < S2; ++k) {
    float2 dx = __fadd2_rn(sx, xi[k]), dy = __fadd2_rn(sy, yi[k]), dz = __fadd2_rn(sz, zi[k]);
    float2 t = __ffma2_rn(dx, dx, make_float2(L.rsm2, L.rsm2));
    t = __ffma2_rn(dy, dy, t);
    t = __ffma2_rn(dz, dz, t);
    float2 p = make_float2(L.b[5], L.b[5]);
#pragma unroll
    for (int q = 4; q >= 0; --q) p = __ffma2_rn(p, t, make_float2(L.b[q], L.b[q]));
    float2 f;
    if (MODE == 5) {
      float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
      f = make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y));
      f = __fadd2_rn(f, make_float2(-p.x, -p.y));
    } else {
      float2 rs = make_float2(rsqrt_ftz(t.x), rsqrt_ftz(t.y));
      f = __ffma2_rn(__fmul2_rn(rs, rs), rs, p);     // p holds -poly (negated coefficients)
    }
    if (MODE == 7) {   // cutoff as predicated scalar accumulates: no FSEL
      if (t.x < L.smax) { ax[k].x = __fmaf_rn(f.x, dx.x, ax[k].x); ay[k].x = __fmaf_rn(f.x, dy.x, ay[k].x); az[k].x = __fmaf_rn(f.x, dz.x, az[k].x); }
      if (t.y < L.smax) { ax[k].y = __fmaf_rn(f.y, dx.y, ax[k].y); ay[k].y = __fmaf_rn(f.y, dy.y, ay[k].y); az[k].y = __fmaf_rn(f.y, dz.y, az[k].y); }
    } else {
    f.x = (t.x < L.smax) ? f.x : 0.0f;
    f.y = (t.y < L.smax) ? f.y : 0.0f;
    ax[k] = __ffma2_rn(f, dx, ax[k]); ay[k] = __ffma2_rn(f, dy, ay[k]); az[k] = __ffma2_rn(f, dz, az[k]);
    }
  }
}
template <int S2, int MODE, int UNROLL, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_fused(const float4 *__restrict__ src, float4 *__restrict__ out, Law L, int nrep) {
  __shared__ float4 tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += 32 * WARPS) tile[i] = src[(blockIdx.x * TILE + i) & 0xffff];
  __syncthreads();
  float2 xi[S2], yi[S2], zi[S2], ax[S2], ay[S2], az[S2];
  const int tid = blockIdx.x * 32 * WARPS + threadIdx.x;
#pragma unroll
  for (int k = 0; k < S2; ++k) {
    float4 a = src[(tid * 2 * S2 + 2 * k) & 0xffff];
    float4 b = src[(tid * 2 * S2 + 2 * k + 1) & 0xffff];
    xi[k] = make_float2(-a.x, -b.x); yi[k] = make_float2(-a.y, -b.y); zi[k] = make_float2(-a.z, -b.z);
    ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
  }
  for (int r = 0; r < nrep; ++r) {
#pragma unroll UNROLL
    for (int j = 0; j < TILE; ++j) pair_fused<S2, MODE>(tile[j], xi, yi, zi, ax, ay, az, L);
  }
  float sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int k = 0; k < S2; ++k) { sx += ax[k].x + ax[k].y; sy += ay[k].x + ay[k].y; sz += az[k].x + az[k].y; }
  out[tid] = make_float4(sx, sy, sz, 0.f);
}

template <typename F>
static void run(const char *name, F launch, int sinks_per_thread, int grid, int nrep, double peak_tflops) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(grid, nrep / 4 + 1);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0));
    launch(grid, nrep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double pairs = (double)grid * 32.0 * sinks_per_thread * (double)TILE * nrep;
  double tf = 30.0 * pairs / (best * 1e-3) / 1e12;
  printf("%-28s S=%d grid=%d  %.3f ms  %.1f Gpairs/s  %.2f TFLOP/s  %.1f%% of %.1f\n", name, sinks_per_thread, grid, best,
         pairs / (best * 1e-3) / 1e9, tf, 100.0 * tf / peak_tflops, peak_tflops);
}

int main(int argc, char **argv) {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  double peak = prop.multiProcessorCount * 128.0 * 2.0 * 1965e6 / 1e12;
  printf("%s, %d SMs, clockRate attr %.0f MHz; FP32 peak at 1965 MHz = %.2f TFLOP/s\n", prop.name, prop.multiProcessorCount, clk_khz / 1e3, peak);
  float4 *src, *out;
  const int NS = 65536;
  CK(cudaMalloc(&src, NS * sizeof(float4)));
  float4 *h = (float4 *)malloc(NS * sizeof(float4));
  srand48(7);
  for (int i = 0; i < NS; ++i) h[i] = make_float4(20.f * drand48(), 20.f * drand48(), 20.f * drand48(), 1.f);
  CK(cudaMemcpy(src, h, NS * sizeof(float4), cudaMemcpyHostToDevice));
  const int grid = prop.multiProcessorCount * 32 * 4;
  CK(cudaMalloc(&out, (size_t)grid * 32 * sizeof(float4)));
  Law L = {{0.269327f, -0.0750978f, 0.0114808f, -0.00109313f, 0.0000605491f, -0.00000147177f}, 0.007f * 0.007f, 3.116326355f * 3.116326355f};
  const int nrep = 200;
  const unsigned zero = (argc > 5) ? 1u : 0u;
#define RUN_SCALAR(S, E) run("scalar " #E, [&](int g, int n) { k_scalar<S, E><<<g, 32>>>(src, out, L, n); }, S, grid, nrep, peak)
#define RUN_SINK2(S2, E, U) run("sink-packed mode" #E " unroll" #U, [&](int g, int n) { k_sink2<S2, E, U><<<g, 32>>>(src, out, L, n, zero); }, 2 * S2, grid, nrep, peak)
#define RUN_SRC2(S, E) run("source-packed " #E, [&](int g, int n) { k_src2<S, E><<<g, 32>>>(src, out, L, n); }, S, grid, nrep, peak)
  for (int q = 0; q < 6; ++q) L.b[q] = L.a[q];
  L.smax = L.rmax2 + L.rsm2;
  Law Ln = L;
  for (int q = 0; q < 6; ++q) Ln.b[q] = -L.a[q];
#define RUN_FUSED(S2, M, U, W) run("fused mode" #M " unroll" #U " warps" #W, [&](int g, int n) { k_fused<S2, M, U, W><<<g / W, 32 * W>>>(src, out, (M >= 6) ? Ln : L, n); }, 2 * S2, grid, nrep, peak)
  if (argc > 1 && argv[1][0] == 'f') {
    RUN_FUSED(2, 7, 2, 1); RUN_FUSED(3, 7, 1, 1); RUN_FUSED(3, 7, 2, 1); RUN_FUSED(4, 7, 1, 1); RUN_FUSED(4, 7, 2, 1);
    RUN_FUSED(1, 5, 4, 1); RUN_FUSED(2, 5, 2, 1); RUN_FUSED(3, 5, 1, 1); RUN_FUSED(3, 5, 2, 1); RUN_FUSED(4, 5, 1, 1); RUN_FUSED(4, 5, 2, 1);
    RUN_FUSED(1, 6, 4, 1); RUN_FUSED(2, 6, 2, 1); RUN_FUSED(3, 6, 1, 1); RUN_FUSED(3, 6, 2, 1); RUN_FUSED(4, 6, 1, 1); RUN_FUSED(4, 6, 2, 1);
    RUN_FUSED(3, 6, 1, 4); RUN_FUSED(4, 6, 1, 4); RUN_FUSED(2, 6, 2, 4); RUN_FUSED(3, 5, 1, 4); RUN_FUSED(4, 5, 1, 4);
    RUN_SINK2(3, 3, 2); RUN_SINK2(4, 3, 1);
    return 0;
  }
  RUN_SCALAR(8, true); RUN_SCALAR(8, false);
  RUN_SINK2(1, 4, 4); RUN_SINK2(2, 4, 4); RUN_SINK2(3, 4, 2); RUN_SINK2(3, 4, 1); RUN_SINK2(4, 4, 1); RUN_SINK2(4, 4, 2);
  RUN_SINK2(1, 3, 4); RUN_SINK2(2, 3, 4); RUN_SINK2(3, 3, 2); RUN_SINK2(4, 3, 1); RUN_SINK2(4, 3, 2); RUN_SINK2(4, 3, 4); RUN_SINK2(6, 3, 1);
  RUN_SINK2(1, 0, 4); RUN_SINK2(2, 0, 4); RUN_SINK2(3, 0, 2); RUN_SINK2(4, 0, 1); RUN_SINK2(4, 0, 2); RUN_SINK2(4, 0, 4); RUN_SINK2(6, 0, 1);
  RUN_SRC2(4, true);
  return 0;
}
