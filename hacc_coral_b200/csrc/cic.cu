// cic.cu -- the particle side of the PM coupling on the device, sm_100a.
//
//   haccsr_cic          Particles::cic          (reference src/cpu/Particles.cxx:589-643): cloud-in-cell deposit of every
//                       resident particle onto the rank's local grid, rho[cell] += c * wx * wy * wz for the 8 cells
//                       around the particle; a cell outside [0, ng) is the reference's `safe` overflow slot (:391-394)
//                       and is dropped.
//   haccsr_inverse_cic  Particles::inverse_cic  (:647-714): interpolate one component of the PM gradient back to the
//                       particles with the same weights and kick  v[comp] += f * fscal * tau  (comp 3 = phi).
// The FFT Poisson solver between the two stays the reference's (north_star), so the grid crosses the boundary as a host
// array by default (grid_on_device = 0) or as a device pointer when the caller keeps it on the GPU.
//
// Arithmetic.  The weights are formed exactly as the reference's expressions evaluate under C's promotion rules:
// ab = float(1.0 + double(float(ix) - x)), the products c*ab*de*gh in float, and any factor (1.0 - de) promotes the rest
// of its product to double (:622-629, :696-703).  inverse_cic sums its 8 terms sequentially per particle like the
// reference, so it is bit-identical to the CPU loop.  The deposit is a scatter: the reference adds the particles'
// terms to a float cell one after the other (order = particle order); here every term is converted to 2^-S fixed point
// and added with 64-bit integer atomics, which makes the result independent of the order (deterministic run to run) and
// equal to the exactly summed value rounded to float once -- it differs from the reference's sequentially rounded sum
// by FP32 rounding only (tests: <= 1e-6 of the cell value).  HBM-bound: 12 B read + 8 scattered 8-B atomics per
// particle (L2-resident when particles are in tree order) + 12 B per cell for the conversion.
#include "common.cuh"

#include <math.h>

namespace haccsr {

struct CicWeights {
  int ix, iy, iz;
  float ab, de, gh;
};
__device__ __forceinline__ CicWeights cic_weights(float xx, float yy, float zz) {
  CicWeights w;
  w.ix = (int)floorf(xx); w.iy = (int)floorf(yy); w.iz = (int)floorf(zz);                    // :606-608
  w.ab = (float)(1.0 + (double)__fsub_rn((float)w.ix, xx));                                   // :614-616
  w.de = (float)(1.0 + (double)__fsub_rn((float)w.iy, yy));
  w.gh = (float)(1.0 + (double)__fsub_rn((float)w.iz, zz));
  return w;
}
// the 8 products of :618-625 / :696-703 for a leading float factor q (c or a grid value), in the reference's cell order
// (ix,iy,iz) (ix,jp,iz) (ix,jp,kp) (ix,iy,kp) (ip,iy,kp) (ip,jp,kp) (ip,jp,iz) (ip,iy,iz)
__device__ __forceinline__ double cic_term(int k, float q, const CicWeights &w) {
  const double ab1 = 1.0 - (double)w.ab, de1 = 1.0 - (double)w.de, gh1 = 1.0 - (double)w.gh;
  const float qab = __fmul_rn(q, w.ab);
  switch (k) {
    case 0: return (double)__fmul_rn(__fmul_rn(qab, w.de), w.gh);                // q*ab*de*gh: all float
    case 1: return ((double)qab * de1) * (double)w.gh;                           // q*ab*(1.0-de)*gh
    case 2: return ((double)qab * de1) * gh1;                                    // q*ab*(1.0-de)*(1.0-gh)
    case 3: return (double)__fmul_rn(qab, w.de) * gh1;                           // q*ab*de*(1.0-gh)
    case 4: return (((double)q * ab1) * (double)w.de) * gh1;                     // q*(1.0-ab)*de*(1.0-gh)
    case 5: return (((double)q * ab1) * de1) * gh1;                              // q*(1.0-ab)*(1.0-de)*(1.0-gh)
    case 6: return (((double)q * ab1) * de1) * (double)w.gh;                     // q*(1.0-ab)*(1.0-de)*gh
    default: return (((double)q * ab1) * (double)w.de) * (double)w.gh;           // q*(1.0-ab)*de*gh
  }
}
__device__ __forceinline__ long long cic_cell(int k, const CicWeights &w, int n0, int n1, int n2) {
  const int dx = (k >= 4) ? 1 : 0, dy = (k == 1 || k == 2 || k == 5 || k == 6) ? 1 : 0, dz = (k >= 2 && k <= 5) ? 1 : 0;
  const int x = w.ix + dx, y = w.iy + dy, z = w.iz + dz;
  if (x < 0 || x >= n0 || y < 0 || y >= n1 || z < 0 || z >= n2) return -1;                   // the `safe` slot (:391-394)
  return ((long long)x * n1 + y) * n2 + z;
}

__global__ void __launch_bounds__(256) k_cic(const float *__restrict__ x, const float *__restrict__ y,
                                             const float *__restrict__ z, long long n, int n0, int n1, int n2, float c,
                                             double scale, unsigned long long *__restrict__ acc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const CicWeights w = cic_weights(x[i], y[i], z[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long cell = cic_cell(k, w, n0, n1, n2);
      if (cell < 0) continue;
      const long long q = __double2ll_rn(cic_term(k, c, w) * scale);
      if (q) atomicAdd(acc + cell, (unsigned long long)q);
    }
  }
}
__global__ void __launch_bounds__(256) k_cic_finish(const unsigned long long *__restrict__ acc, long long ncell, double inv_scale,
                                                    float *__restrict__ rho) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x)
    rho[i] = (float)((double)(long long)acc[i] * inv_scale);
}

__global__ void __launch_bounds__(256) k_inverse_cic(const float *__restrict__ x, const float *__restrict__ y,
                                                     const float *__restrict__ z, float *__restrict__ v, long long n, int n0,
                                                     int n1, int n2, const float *__restrict__ grid, float tau, float fscal) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const CicWeights w = cic_weights(x[i], y[i], z[i]);
    float f = 0.f;                                                                             // :694
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long cell = cic_cell(k, w, n0, n1, n2);
      const float gv = cell >= 0 ? __ldg(grid + cell) : 0.0f;                                  // grad_phi[safe] = 0 (:677)
      const double term = cic_term(k, gv, w);
      f = (k == 0) ? __fadd_rn(f, (float)term) : (float)((double)f + term);                    // f += <float> / <double>
    }
    v[i] = __fadd_rn(v[i], __fmul_rn(__fmul_rn(f, fscal), tau));                               // :705
  }
}

static int grid_lin(const haccsr_ctx *c, long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)c->sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace haccsr

using namespace haccsr;

extern "C" {

int haccsr_cic(haccsr_ctx *c, const int32_t ng[3], float cfac, float *rho, int grid_on_device) {
  if (!c) { set_error("null context"); return 1; }
  if (!ng || !rho || ng[0] <= 0 || ng[1] <= 0 || ng[2] <= 0) { set_error("haccsr_cic: bad grid"); return 1; }
  if (!(cfac > 0.f) || !isfinite(cfac)) { set_error("haccsr_cic: the deposit factor must be positive and finite"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  const long long ncell = (long long)ng[0] * ng[1] * ng[2];
  HSR_TRY(c->cic_acc.ensure((size_t)ncell));
  cudaStream_t s = c->stream;
  HSR_CUDA(cudaMemsetAsync(c->cic_acc.p, 0, (size_t)ncell * sizeof(unsigned long long), s));
  // fixed point: a term is at most c; 2^40 steps per power of two above c leave 2^22 particles' worth of headroom per cell
  int e = 0;
  frexp((double)cfac, &e);
  const double scale = ldexp(1.0, 40 - e);
  const long long n = c->n_resident;
  if (n > 0) k_cic<<<grid_lin(c, n), 256, 0, s>>>(c->cur.x, c->cur.y, c->cur.z, n, ng[0], ng[1], ng[2], cfac, scale, c->cic_acc.p);
  float *out = rho;
  if (!grid_on_device) { HSR_TRY(c->cic_grid.ensure((size_t)ncell)); out = c->cic_grid.p; }
  k_cic_finish<<<grid_lin(c, ncell), 256, 0, s>>>(c->cic_acc.p, ncell, 1.0 / scale, out);
  HSR_CUDA(cudaGetLastError());
  if (!grid_on_device) HSR_CUDA(cudaMemcpyAsync(rho, out, (size_t)ncell * sizeof(float), cudaMemcpyDeviceToHost, s));
  HSR_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int haccsr_inverse_cic(haccsr_ctx *c, const int32_t ng[3], const float *grid, int grid_on_device, float tau, float fscal,
                       int comp) {
  if (!c) { set_error("null context"); return 1; }
  if (!ng || !grid || ng[0] <= 0 || ng[1] <= 0 || ng[2] <= 0) { set_error("haccsr_inverse_cic: bad grid"); return 1; }
  if (comp < 0 || comp > 3) { set_error("haccsr_inverse_cic: comp must be 0, 1, 2 (vx vy vz) or 3 (phi)"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  const long long ncell = (long long)ng[0] * ng[1] * ng[2];
  cudaStream_t s = c->stream;
  const float *g = grid;
  if (!grid_on_device) {
    HSR_TRY(c->cic_grid.ensure((size_t)ncell));
    HSR_CUDA(cudaMemcpyAsync(c->cic_grid.p, grid, (size_t)ncell * sizeof(float), cudaMemcpyHostToDevice, s));
    g = c->cic_grid.p;
  }
  float *v[4] = {c->cur.vx, c->cur.vy, c->cur.vz, c->cur.phi};
  const long long n = c->n_resident;
  if (n > 0) k_inverse_cic<<<grid_lin(c, n), 256, 0, s>>>(c->cur.x, c->cur.y, c->cur.z, v[comp], n, ng[0], ng[1], ng[2], g, tau, fscal);
  HSR_CUDA(cudaGetLastError());
  HSR_CUDA(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"
