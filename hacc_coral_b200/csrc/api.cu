// api.cu -- the C ABI of libhaccsr.so (include/haccsr.h): context, transfers, the kick orchestration and
// the small HBM-bound helpers around it (stream = Particles::map1, out-of-box compaction, mass fill).
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

namespace haccsr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int alloc_soa(Soa &s, int64_t n) {
  size_t m = (size_t)n + 64;
  float **f[8] = {&s.x, &s.y, &s.z, &s.vx, &s.vy, &s.vz, &s.mass, &s.phi};
  for (int i = 0; i < 8; ++i) HSR_CUDA(cudaMalloc((void **)f[i], m * sizeof(float)));
  HSR_CUDA(cudaMalloc((void **)&s.id, m * sizeof(int64_t)));
  HSR_CUDA(cudaMalloc((void **)&s.mask, m * sizeof(uint16_t)));
  return 0;
}
static void free_soa(Soa &s) {
  float *f[8] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, s.phi};
  for (int i = 0; i < 8; ++i) if (f[i]) cudaFree(f[i]);
  if (s.id) cudaFree(s.id);
  if (s.mask) cudaFree(s.mask);
  memset(&s, 0, sizeof(s));
}

// x += pt * v   (Particles::map1, reference src/cpu/Particles.cxx:745-755; pt = prefactor * tau).
// The reference evaluates x + (prefactor*tau)*v with a float multiply then a float add; kept unfused.
__global__ void __launch_bounds__(256) k_stream(float *__restrict__ x, float *__restrict__ y, float *__restrict__ z,
                                                const float *__restrict__ vx, const float *__restrict__ vy,
                                                const float *__restrict__ vz, float pt, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = __fadd_rn(x[i], __fmul_rn(pt, vx[i]));
    y[i] = __fadd_rn(y[i], __fmul_rn(pt, vy[i]));
    z[i] = __fadd_rn(z[i], __fmul_rn(pt, vz[i]));
  }
}

__global__ void __launch_bounds__(256) k_fill(float *__restrict__ a, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = v;
}

// in-box flag per particle: 0 <= floor(x) < hi in every dimension (Particles::resortParticles uses
// array_index on floor'ed coordinates; anything outside the local grid lands in the overflow cell Ng,
// reference src/cpu/Particles.cxx:434-443).
__global__ void __launch_bounds__(256) k_inbox_flags(const float *__restrict__ x, const float *__restrict__ y,
                                                     const float *__restrict__ z, float3 hi, long long n,
                                                     unsigned *__restrict__ flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float fx = floorf(x[i]), fy = floorf(y[i]), fz = floorf(z[i]);
    bool in = fx >= 0.f && fx < hi.x && fy >= 0.f && fy < hi.y && fz >= 0.f && fz < hi.z;
    flag[i] = in ? 1u : 0u;
  }
}
__global__ void __launch_bounds__(256) k_compact(Soa in, Soa out, const unsigned *__restrict__ flag,
                                                 const unsigned *__restrict__ pref, unsigned n_in, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned p = pref[i];
    long long d = flag[i] ? (long long)p : (long long)n_in + (i - (long long)p);
    out.x[d] = in.x[i]; out.y[d] = in.y[i]; out.z[d] = in.z[i];
    out.vx[d] = in.vx[i]; out.vy[d] = in.vy[i]; out.vz[d] = in.vz[i];
    out.mass[d] = in.mass[i]; out.phi[d] = in.phi[i]; out.id[d] = in.id[i]; out.mask[d] = in.mask[i];
  }
}

static int lin_grid(const haccsr_ctx *c, int64_t n) {
  int64_t g = (n + 255) / 256;
  int64_t cap = (int64_t)c->sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// stable two-way partition of all ten arrays: flagged particles first (in order), the rest behind them (in order)
int compact_by_flags(haccsr_ctx *c, const unsigned *flag, unsigned *pref, int64_t n, int64_t *n_kept) {
  HSR_TRY(scan_exclusive(c, flag, pref, n, c->d_counters + 12));
  HSR_CUDA(cudaMemcpyAsync(c->h_counters + 12, c->d_counters + 12, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t nin = c->h_counters[12];
  k_compact<<<lin_grid(c, n), 256, 0, c->stream>>>(c->cur, c->alt, flag, pref, (unsigned)nin, (long long)n);
  HSR_CUDA(cudaGetLastError());
  Soa t = c->cur; c->cur = c->alt; c->alt = t;
  c->order_epoch++;
  if (n_kept) *n_kept = nin;
  return 0;
}

}  // namespace haccsr

using namespace haccsr;

extern "C" {

const char *haccsr_last_error(void) { return g_err; }

int haccsr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
}

int haccsr_create(haccsr_ctx **out, int device, int64_t max_particles) {
  if (!out) { set_error("haccsr_create: out is NULL"); return 1; }
  *out = nullptr;
  if (max_particles < 0 || max_particles > 2000000000ll) { set_error("haccsr_create: bad max_particles"); return 1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); libhaccsr has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return 3;
  }
  if (device < 0 || device >= ndev) { set_error("haccsr_create: device %d out of range (have %d)", device, ndev); return 1; }
  cudaDeviceProp prop;
  HSR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libhaccsr is built for sm_100a only", device, prop.major, prop.minor);
    return 3;
  }
  HSR_CUDA(cudaSetDevice(device));
  haccsr_ctx *c = new (std::nothrow) haccsr_ctx();
  if (!c) { set_error("out of host memory"); return 2; }
  c->device = device; c->sm_count = prop.multiProcessorCount; c->cap = max_particles;
  if (const char *e = getenv("HACCSR_HOST_GROUPS")) { int g = atoi(e); c->host_groups = g < 1 ? 1 : (g > 8 ? 8 : g); }
  if (const char *e = getenv("HACCSR_HOST_LAST")) { float f = (float)atof(e); c->host_last_frac = (f >= 0.f && f < 1.f) ? f : 0.f; }
  if (const char *e = getenv("HACCSR_ITEM_POLICY")) c->item_policy = atoi(e) & 0x71;   // bit 0: remainder items; bits 4-6: largest chunk in groups (tuning)
  int rc = 0;
  do {
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = 2; break; }
    c->stream = c->own_stream;
    if ((rc = alloc_soa(c->cur, max_particles))) break;
    if ((rc = alloc_soa(c->alt, max_particles))) break;
    if (cudaMallocHost((void **)&c->h_state, sizeof(BuildState)) != cudaSuccess) { rc = 2; break; }
    if (cudaMalloc((void **)&c->d_state, sizeof(BuildState)) != cudaSuccess) { rc = 2; break; }
    if (cudaMallocHost((void **)&c->h_counters, 32 * sizeof(int64_t)) != cudaSuccess) { rc = 2; break; }
    if (cudaMalloc((void **)&c->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess) { rc = 2; break; }
    if (cudaMalloc((void **)&c->d_slotcount, 32 * sizeof(long long)) != cudaSuccess) { rc = 2; break; }
    for (int i = 0; i < 5; ++i) if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { rc = 2; break; }
    {
      // the copy stream also runs the small gather of phi / id / mask while the force kernel (130 k one-warp CTAs) owns the
      // machine: with equal priority its blocks queue behind the force kernel's and the downloads behind it start late
      int least = 0, greatest = 0;
      if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { rc = 2; break; }
      if (cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, greatest) != cudaSuccess) { rc = 2; break; }
    }
    if (cudaEventCreateWithFlags(&c->ev_up2, cudaEventDisableTiming) != cudaSuccess) { rc = 2; break; }
    if (cudaEventCreateWithFlags(&c->ev_vready, cudaEventDisableTiming) != cudaSuccess) { rc = 2; break; }
    if (cudaEventCreateWithFlags(&c->ev_built, cudaEventDisableTiming) != cudaSuccess) { rc = 2; break; }
    { bool bad = false;
      for (int g = 0; g < 8; ++g) if (cudaEventCreateWithFlags(&c->ev_grp[g], cudaEventDisableTiming) != cudaSuccess) bad = true;
      if (bad) { rc = 2; break; } }
    if (cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming) != cudaSuccess) { rc = 2; break; }
  } while (0);
  if (rc) {
    if (g_err[0] == 0 || rc == 2) set_error("haccsr_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    haccsr_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

int haccsr_destroy(haccsr_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_soa(c->cur); free_soa(c->alt);
  c->recA.release(); c->recB.release(); c->src4.release(); c->idxA.release(); c->idxB.release(); c->perm.release();
  c->nidA.release(); c->nidB.release(); c->nodes.release(); c->acc.release(); c->tot.release(); c->tile_desc.release();
  c->node_desc.release(); c->tilecount.release(); c->tilebase.release(); c->scratch_u32.release();
  c->n_ranges.release(); c->n_pseudo.release(); c->range_off.release(); c->pseudo_off.release(); c->list_len.release();
  c->ranges.release(); c->pool.release(); c->law_table.release(); c->item_cnt.release(); c->item_off.release(); c->items.release(); c->items_sorted.release(); c->lpt_hist.release(); c->refresh_slots.release();
  c->scan_tmp.release(); c->pp12.release(); c->cic_acc.release(); c->cic_grid.release();
  if (c->h_state) cudaFreeHost(c->h_state);
  if (c->d_state) cudaFree(c->d_state);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->d_counters) cudaFree(c->d_counters);
  if (c->d_slotcount) cudaFree(c->d_slotcount);
  c->refresh_cand.release(); c->xchg_send.release(); c->xchg_recv.release(); c->xchg_table.release();
  for (int i = 0; i < 5; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->ev_up2) cudaEventDestroy(c->ev_up2);
  if (c->ev_vready) cudaEventDestroy(c->ev_vready);
  for (int q = 0; q < 3; ++q) c->kick_a[q].release();
  if (c->ev_built) cudaEventDestroy(c->ev_built);
  for (int g = 0; g < 8; ++g) if (c->ev_grp[g]) cudaEventDestroy(c->ev_grp[g]);
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return 0;
}

int haccsr_set_stream(haccsr_ctx *c, void *cuda_stream) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return 0;
}

int haccsr_set_force_law(haccsr_ctx *c, int kind, const float *coeffs, int ncoef, float rsm, float rmax) {
  if (!c) { set_error("null context"); return 1; }
  if (kind != HACCSR_LAW_SR_POLY && kind != HACCSR_LAW_NEWTON && kind != HACCSR_LAW_SR_FIT && kind != HACCSR_LAW_SR_INTERP) {
    set_error("unsupported force law kind %d (supported: SR_POLY=0, SR_FIT=1, SR_INTERP=2, NEWTON=3); no CPU fallback", kind);
    return 1;
  }
  if (kind == HACCSR_LAW_SR_POLY && (ncoef < 1 || ncoef > 7 || !coeffs)) { set_error("SR_POLY needs 1..7 coefficients"); return 1; }
  if (kind == HACCSR_LAW_SR_FIT && (ncoef != 8 || !coeffs)) { set_error("SR_FIT needs the 8 constants b c d e f g h l"); return 1; }
  if (kind == HACCSR_LAW_SR_INTERP && (ncoef < 2 || ncoef > 4096 || !coeffs)) { set_error("SR_INTERP needs a table of 2..4096 samples"); return 1; }
  if (!(rmax > 0.f)) { set_error("rmax must be positive"); return 1; }
  // the reference's analytic fit is zero beyond FGrid::m_rmax whatever cutoff the tree is given (ForceLaw.cxx:32,70-80,187-192);
  // the device applies the caller's cutoff only, so a larger one is refused rather than evaluated differently
  if (kind == HACCSR_LAW_SR_FIT && rmax > 3.1163266f) {
    set_error("SR_FIT: rmax %g exceeds the range of the grid-force fit (3.116326355, ForceLaw.cxx:32)", rmax);
    return 1;
  }
  HSR_CUDA(cudaSetDevice(c->device));
  memset(&c->law, 0, sizeof(c->law));
  c->law.kind = kind;
  c->law.ncoef = (kind == HACCSR_LAW_SR_POLY || kind == HACCSR_LAW_SR_FIT) ? ncoef : 0;
  for (int i = 0; i < c->law.ncoef; ++i) c->law.a[i] = coeffs[i];
  if (kind == HACCSR_LAW_SR_INTERP) {
    // the evaluator's constants exactly as FGridEvalInterp derives them (ForceLaw.cxx:54-65,145-152):
    // r2_i = float(i * dr2) with dr2 = rmax^2 / (n - 1) in double; m_dr2, m_oodr2 in float
    const int n = ncoef;
    float *h = (float *)malloc(2 * (size_t)n * sizeof(float));
    if (!h) { set_error("out of host memory"); return 2; }
    const double dr2 = (double)(rmax * rmax) / (n - 1.0);
    for (int i = 0; i < n; ++i) { h[i] = coeffs[i]; h[n + i] = (float)(i * dr2); }
    c->law.ntab = n;
    c->law.tab_r2min = h[n]; c->law.tab_r2max = h[2 * n - 1];
    const float fdr2 = (float)((c->law.tab_r2max - c->law.tab_r2min) / (n - 1.0));
    c->law.tab_oodr2 = (float)(1.0 / fdr2);
    int rc = c->law_table.ensure(2 * (size_t)n);
    if (rc == 0 && cudaMemcpy(c->law_table.p, h, 2 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("copy of the interpolation table failed"); rc = 2;
    }
    free(h);
    if (rc) return rc;
  }
  c->law.rsm2 = (kind == HACCSR_LAW_NEWTON) ? 0.0f : rsm * rsm;   // ForceLaw.cxx:177 (m_rsm2 = rsm*rsm, float)
  c->law.rmax = rmax;
  c->law.rmax2 = rmax * rmax;                                     // RCBForceTree.cxx:582 (float product)
  c->law.smax = c->law.rmax2 + c->law.rsm2;
  if (kind == HACCSR_LAW_SR_POLY) {
    // HACCSR_ARITH_FUSED evaluates the polynomial in s = r2 + eps, eps = rsm2:  sum_j a_j (s - eps)^j = sum_k b_k s^k,
    // b_k = sum_{j>=k} a_j C(j,k) (-eps)^(j-k), accumulated in double from j = k upwards, rounded once to float.
    const double eps = (double)c->law.rsm2;
    for (int k = 0; k < ncoef; ++k) {
      double acc = 0.0, binom = 1.0, pw = 1.0;        // C(j,k), (-eps)^(j-k)
      for (int j = k; j < ncoef; ++j) {
        acc += (double)coeffs[j] * binom * pw;
        binom = binom * (double)(j + 1) / (double)(j + 1 - k);
        pw *= -eps;
      }
      c->law.b[k] = (float)acc;
    }
  }
  c->law_set = true;
  return 0;
}

int haccsr_set_culling(haccsr_ctx *c, int on) {
  if (!c) { set_error("null context"); return 1; }
  c->cull = on ? 1 : 0;
  return 0;
}

int haccsr_set_arithmetic(haccsr_ctx *c, int mode) {
  if (!c) { set_error("null context"); return 1; }
  if (mode != HACCSR_ARITH_FUSED && mode != HACCSR_ARITH_X86 && mode != HACCSR_ARITH_FUSED_RS3) { set_error("unknown arithmetic mode %d", mode); return 1; }
  c->arith = mode;
  return 0;
}

#define COPY_H2D(dst, src, bytes) do { if (src) HSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream)); } while (0)
#define COPY_D2H(dst, src, bytes) do { if (dst) HSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream)); } while (0)

int haccsr_upload(haccsr_ctx *c, int64_t n, const float *x, const float *y, const float *z, const float *vx,
                  const float *vy, const float *vz, const float *mass, const float *phi, const int64_t *id,
                  const uint16_t *mask) {
  if (!c) { set_error("null context"); return 1; }
  if (n < 0 || n > c->cap) { set_error("haccsr_upload: count %lld exceeds capacity %lld", (long long)n, (long long)c->cap); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  size_t fb = (size_t)n * sizeof(float);
  if (!x || !y || !z || !vx || !vy || !vz || !mass) { set_error("haccsr_upload: x y z vx vy vz mass are required"); return 1; }
  COPY_H2D(c->cur.x, x, fb); COPY_H2D(c->cur.y, y, fb); COPY_H2D(c->cur.z, z, fb);
  COPY_H2D(c->cur.vx, vx, fb); COPY_H2D(c->cur.vy, vy, fb); COPY_H2D(c->cur.vz, vz, fb);
  COPY_H2D(c->cur.mass, mass, fb);
  if (phi) { COPY_H2D(c->cur.phi, phi, fb); } else { HSR_CUDA(cudaMemsetAsync(c->cur.phi, 0, fb, c->stream)); }
  if (id) { COPY_H2D(c->cur.id, id, (size_t)n * sizeof(int64_t)); } else { HSR_CUDA(cudaMemsetAsync(c->cur.id, 0, (size_t)n * sizeof(int64_t), c->stream)); }
  if (mask) { COPY_H2D(c->cur.mask, mask, (size_t)n * sizeof(uint16_t)); } else { HSR_CUDA(cudaMemsetAsync(c->cur.mask, 0, (size_t)n * sizeof(uint16_t), c->stream)); }
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  c->n_resident = n;
  c->order_epoch++;
  return 0;
}

int haccsr_download(haccsr_ctx *c, int64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz,
                    float *mass, float *phi, int64_t *id, uint16_t *mask) {
  if (!c) { set_error("null context"); return 1; }
  if (n < 0 || n > c->n_resident) { set_error("haccsr_download: count %lld exceeds resident %lld", (long long)n, (long long)c->n_resident); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  size_t fb = (size_t)n * sizeof(float);
  COPY_D2H(x, c->cur.x, fb); COPY_D2H(y, c->cur.y, fb); COPY_D2H(z, c->cur.z, fb);
  COPY_D2H(vx, c->cur.vx, fb); COPY_D2H(vy, c->cur.vy, fb); COPY_D2H(vz, c->cur.vz, fb);
  COPY_D2H(mass, c->cur.mass, fb); COPY_D2H(phi, c->cur.phi, fb);
  COPY_D2H(id, c->cur.id, (size_t)n * sizeof(int64_t)); COPY_D2H(mask, c->cur.mask, (size_t)n * sizeof(uint16_t));
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int haccsr_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return 0;
  HSR_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return 0;
}
int haccsr_host_unregister(void *ptr) {
  if (!ptr) return 0;
  HSR_CUDA(cudaHostUnregister(ptr));
  return 0;
}

struct HostOut { float *x, *y, *z, *vx, *vy, *vz, *mass, *phi; int64_t *id; uint16_t *mask; };

}  // extern "C"
namespace haccsr {
int issue_host_out(haccsr_ctx *c) {
  if (!c->pending_ho) return 0;
  const HostOut *ho = (const HostOut *)c->pending_ho;
  c->pending_ho = nullptr;
  cudaStream_t cs = c->copy_stream, s = c->stream;
  const int64_t count = c->pending_count;
  const size_t fb = (size_t)count * sizeof(float);
  HSR_CUDA(cudaEventRecord(c->ev_built, s));
  HSR_CUDA(cudaStreamWaitEvent(cs, c->ev_built, 0));
  // velocities, phi, id, mask into tree order: behind their upload (same stream) and behind the build.  The build and the force
  // kernel do not wait for any of them: the force kernel leaves accelerations and apply_kick() kicks behind ev_vready (with
  // eight ranks sharing one host the six arrays take 33 ms to arrive, which used to sit between the build and the force kernel)
  HSR_TRY(gather_payload(c, cs));
  HSR_CUDA(cudaEventRecord(c->ev_vready, cs));
  HSR_CUDA(cudaMemcpyAsync(ho->x, c->cur.x, fb, cudaMemcpyDeviceToHost, cs));
  HSR_CUDA(cudaMemcpyAsync(ho->y, c->cur.y, fb, cudaMemcpyDeviceToHost, cs));
  HSR_CUDA(cudaMemcpyAsync(ho->z, c->cur.z, fb, cudaMemcpyDeviceToHost, cs));
  HSR_CUDA(cudaMemcpyAsync(ho->mass, c->cur.mass, fb, cudaMemcpyDeviceToHost, cs));
  if (ho->phi) HSR_CUDA(cudaMemcpyAsync(ho->phi, c->cur.phi, fb, cudaMemcpyDeviceToHost, cs));
  if (ho->id) HSR_CUDA(cudaMemcpyAsync(ho->id, c->cur.id, (size_t)count * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
  if (ho->mask) HSR_CUDA(cudaMemcpyAsync(ho->mask, c->cur.mask, (size_t)count * sizeof(uint16_t), cudaMemcpyDeviceToHost, cs));
  return 0;
}
}  // namespace haccsr
extern "C" {

static int kick_impl(haccsr_ctx *c, int64_t count, const float tree_lo[3], const float tree_hi[3], const float force_lo[3],
                     const float force_hi[3], float theta, int64_t ppn, int tdpts, float fcoeff,
                     const haccsr_kick_opts *opts, haccsr_stats *stats, const HostOut *ho) {
  if (!c) { set_error("null context"); return 1; }
  if (!c->law_set) { set_error("haccsr_kick: force law not set"); return 1; }
  if (tdpts != 1 && tdpts != 12) { set_error("haccsr_kick: tdpts must be 1 (RCBMonopoleForceTree, -R) or 12 (RCBQuadrupoleForceTree, -S); got %d", tdpts); return 1; }
  if (count < 0 || count > c->n_resident) { set_error("haccsr_kick: count %lld exceeds resident %lld", (long long)count, (long long)c->n_resident); return 1; }
  if (ppn < 1) { set_error("haccsr_kick: ppn must be >= 1"); return 1; }
  if (!tree_lo || !tree_hi || !force_lo || !force_hi) { set_error("haccsr_kick: null box"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  haccsr_stats local; memset(&local, 0, sizeof(local));
  haccsr_stats *st = stats ? stats : &local;
  memset(st, 0, sizeof(*st));
  st->particles = count;
  const bool count_cut = opts && opts->count_in_cutoff;
  const bool skip_force = opts && opts->skip_force;
  c->launches = 0; c->force_launches = 0;
  cudaStream_t s = c->stream;
  HSR_CUDA(cudaEventRecord(c->ev[0], s));
  HSR_TRY(build_tree(c, count, tree_lo, tree_hi, ppn, tdpts));
  HSR_CUDA(cudaEventRecord(c->ev[1], s));
  // host output: the build has permuted all ten arrays into c->cur and everything but the velocities is final now.  The
  // seven device->host copies are queued by issue_host_out() right before the force kernel is launched, not here: the
  // walk and the work-item setup read a few counters back, and a small device->host copy issued after 645 MB of them
  // waits in the same copy engine (measured: the walk took 11.8 ms instead of 0.7 ms).
  c->pending_ho = ho ? (const void *)ho : nullptr;
  c->pending_count = count;
  c->defer_kick = ho != nullptr;
  HSR_TRY(build_lists(c, force_lo, force_hi, theta, st));
  HSR_CUDA(cudaEventRecord(c->ev[2], s));
  if (!skip_force) {
    // with host output the force kernel runs as launches by particle range; the velocities of a range leave on the copy
    // stream while the next range is computed (force.cu: launch_force).  Two ranges, 90 % and 10 %: every launch has a tail
    // and only the last range's copy is exposed, so few launches and a short last range (measured per step: 110.3 ms; 85/15: 110.7; four equal ranges: 111.7; three: 111.8)
    c->force_groups = (ho && count >= (1 << 20)) ? c->host_groups : 1;
    {
      const int G = c->force_groups;
      const double last = (G > 1 && c->host_last_frac > 0.f) ? (double)c->host_last_frac : (G > 0 ? 1.0 / G : 1.0);
      for (int g = 0; g <= G; ++g) {
        const double f = g == G ? 1.0 : (G > 1 ? (1.0 - last) * (double)g / (double)(G - 1) : 0.0);
        c->group_lo[g] = (int64_t)((double)count * f);
      }
      c->group_lo[0] = 0; c->group_lo[G] = count;
    }
    c->ho_v[0] = ho ? ho->vx : nullptr; c->ho_v[1] = ho ? ho->vy : nullptr; c->ho_v[2] = ho ? ho->vz : nullptr;
    int rc = run_force(c, fcoeff, count_cut, st);
    if (rc == 0) rc = issue_host_out(c);     // no work items: nothing was launched, the copies are still pending
    const bool copied = c->force_groups > 1 && c->n_items > 0;
    c->force_groups = 1; c->ho_v[0] = c->ho_v[1] = c->ho_v[2] = nullptr;
    if (rc) return rc;
    HSR_CUDA(cudaEventRecord(c->ev[3], s));
    if (ho && !copied) {
      const size_t fb = (size_t)count * sizeof(float);
      HSR_CUDA(cudaStreamWaitEvent(s, c->ev_vready, 0));       // the velocities are permuted on the copy stream
      HSR_CUDA(cudaMemcpyAsync(ho->vx, c->cur.vx, fb, cudaMemcpyDeviceToHost, s));
      HSR_CUDA(cudaMemcpyAsync(ho->vy, c->cur.vy, fb, cudaMemcpyDeviceToHost, s));
      HSR_CUDA(cudaMemcpyAsync(ho->vz, c->cur.vz, fb, cudaMemcpyDeviceToHost, s));
    }
  } else {
    HSR_TRY(issue_host_out(c));
    HSR_CUDA(cudaEventRecord(c->ev[3], s));
  }
  c->defer_kick = false;
  if (ho) HSR_CUDA(cudaStreamSynchronize(c->copy_stream));
  HSR_CUDA(cudaStreamSynchronize(s));
  HSR_CUDA(cudaGetLastError());
  HSR_CUDA(cudaEventElapsedTime(&st->ms_build, c->ev[0], c->ev[1]));
  HSR_CUDA(cudaEventElapsedTime(&st->ms_walk, c->ev[1], c->ev[2]));
  HSR_CUDA(cudaEventElapsedTime(&st->ms_force, c->ev[2], c->ev[3]));
  HSR_CUDA(cudaEventElapsedTime(&st->ms_total, c->ev[0], c->ev[3]));
  st->force_launches = c->force_launches; st->total_launches = c->launches;
  return 0;
}

int haccsr_kick(haccsr_ctx *c, int64_t count, const float tree_lo[3], const float tree_hi[3], const float force_lo[3],
                const float force_hi[3], float theta, int64_t ppn, int tdpts, float fcoeff,
                const haccsr_kick_opts *opts, haccsr_stats *stats) {
  return kick_impl(c, count, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, tdpts, fcoeff, opts, stats, nullptr);
}

// upload -> kick -> download in one call, with the transfers the kernels do not depend on moved to a second
// stream: the tree build reads only x y z mass, so vx vy vz phi id mask arrive while it runs; the force kernel
// writes only vx vy vz, so the other seven arrays (already permuted by the build) leave while it runs.
int haccsr_kick_host(haccsr_ctx *c, int64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz, float *mass,
                     float *phi, int64_t *id, uint16_t *mask, const float tree_lo[3], const float tree_hi[3],
                     const float force_lo[3], const float force_hi[3], float theta, int64_t ppn, int tdpts, float fcoeff,
                     const haccsr_kick_opts *opts, haccsr_stats *stats) {
  if (!c) { set_error("null context"); return 1; }
  if (n < 0 || n > c->cap) { set_error("haccsr_kick_host: count %lld exceeds capacity %lld", (long long)n, (long long)c->cap); return 1; }
  if (!x || !y || !z || !vx || !vy || !vz || !mass) { set_error("haccsr_kick_host: x y z vx vy vz mass are required"); return 1; }
  if (opts && opts->skip_force) { set_error("haccsr_kick_host: skip_force is not supported here"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream, cs = c->copy_stream;
  const size_t fb = (size_t)n * sizeof(float);
  // the copy stream must not run ahead of work already queued on the main stream (previous users of `cur`)
  HSR_CUDA(cudaEventRecord(c->ev_main, s));
  HSR_CUDA(cudaStreamWaitEvent(cs, c->ev_main, 0));
  HSR_CUDA(cudaMemcpyAsync(c->cur.x, x, fb, cudaMemcpyHostToDevice, s));
  HSR_CUDA(cudaMemcpyAsync(c->cur.y, y, fb, cudaMemcpyHostToDevice, s));
  HSR_CUDA(cudaMemcpyAsync(c->cur.z, z, fb, cudaMemcpyHostToDevice, s));
  HSR_CUDA(cudaMemcpyAsync(c->cur.mass, mass, fb, cudaMemcpyHostToDevice, s));
  // the six arrays the build does not read queue up BEHIND the four it does: both streams share one PCIe direction,
  // and left to themselves the copy engines interleave them, doubling the time until the build can start
  HSR_CUDA(cudaEventRecord(c->ev_main, s));
  HSR_CUDA(cudaStreamWaitEvent(cs, c->ev_main, 0));
  HSR_CUDA(cudaMemcpyAsync(c->cur.vx, vx, fb, cudaMemcpyHostToDevice, cs));
  HSR_CUDA(cudaMemcpyAsync(c->cur.vy, vy, fb, cudaMemcpyHostToDevice, cs));
  HSR_CUDA(cudaMemcpyAsync(c->cur.vz, vz, fb, cudaMemcpyHostToDevice, cs));
  if (phi) HSR_CUDA(cudaMemcpyAsync(c->cur.phi, phi, fb, cudaMemcpyHostToDevice, cs));
  else HSR_CUDA(cudaMemsetAsync(c->cur.phi, 0, fb, cs));
  if (id) HSR_CUDA(cudaMemcpyAsync(c->cur.id, id, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  else HSR_CUDA(cudaMemsetAsync(c->cur.id, 0, (size_t)n * sizeof(int64_t), cs));
  if (mask) HSR_CUDA(cudaMemcpyAsync(c->cur.mask, mask, (size_t)n * sizeof(uint16_t), cudaMemcpyHostToDevice, cs));
  else HSR_CUDA(cudaMemsetAsync(c->cur.mask, 0, (size_t)n * sizeof(uint16_t), cs));
  HSR_CUDA(cudaEventRecord(c->ev_up2, cs));
  c->wait_up2 = true;
  c->n_resident = n;
  c->order_epoch++;
  // inside kick_impl the copy stream takes the seven arrays the force kernel leaves alone as soon as the build
  // has permuted them; the velocities follow the force kernel on the main stream
  const HostOut ho = {x, y, z, vx, vy, vz, mass, phi, id, mask};
  int rc = kick_impl(c, n, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, tdpts, fcoeff, opts, stats, &ho);
  if (c->wait_up2) { c->wait_up2 = false; cudaStreamWaitEvent(s, c->ev_up2, 0); }   // n == 0 or an early error
  cudaStreamSynchronize(cs);
  if (rc) c->n_resident = 0;      // the device arrays may be partly uploaded / partly permuted: nothing usable is resident
  return rc;
}

int haccsr_stream(haccsr_ctx *c, float pt) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  if (c->n_resident == 0) return 0;
  k_stream<<<lin_grid(c, c->n_resident), 256, 0, c->stream>>>(c->cur.x, c->cur.y, c->cur.z, c->cur.vx, c->cur.vy,
                                                              c->cur.vz, pt, (long long)c->n_resident);
  HSR_CUDA(cudaGetLastError());
  return 0;
}

int haccsr_fill_mass(haccsr_ctx *c, float value) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  if (c->n_resident == 0) return 0;
  k_fill<<<lin_grid(c, c->n_resident), 256, 0, c->stream>>>(c->cur.mass, value, (long long)c->n_resident);
  HSR_CUDA(cudaGetLastError());
  return 0;
}

int haccsr_partition_in_box(haccsr_ctx *c, const float hi[3], int64_t *count_in_box) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  const int64_t n = c->n_resident;
  if (n == 0) { if (count_in_box) *count_in_box = 0; return 0; }
  // flags / prefixes borrow the build's index buffers (they are rebuilt by every kick)
  HSR_TRY(c->idxA.ensure((size_t)n + 1)); HSR_TRY(c->idxB.ensure((size_t)n + 1));
  k_inbox_flags<<<lin_grid(c, n), 256, 0, c->stream>>>(c->cur.x, c->cur.y, c->cur.z, make_float3(hi[0], hi[1], hi[2]), (long long)n, c->idxA.p);
  int64_t nin = 0;
  HSR_TRY(compact_by_flags(c, c->idxA.p, c->idxB.p, n, &nin));
  if (count_in_box) *count_in_box = nin;
  return 0;
}

int haccsr_subcycle(haccsr_ctx *c, int nsub, float pt, const float box_hi[3], const float tree_lo[3], const float tree_hi[3],
                    const float force_lo[3], const float force_hi[3], float theta, int64_t ppn, int tdpts, float fcoeff,
                    haccsr_stats *stats) {
  if (!c) { set_error("null context"); return 1; }
  if (nsub < 1) { set_error("haccsr_subcycle: nsub must be >= 1"); return 1; }
  if (!box_hi) { set_error("haccsr_subcycle: null box"); return 1; }
  haccsr_stats sum; memset(&sum, 0, sizeof(sum));
  for (int s = 0; s < nsub; ++s) {
    haccsr_stats st;
    int64_t nin = 0;
    HSR_TRY(haccsr_stream(c, pt));                                  // Particles.cxx:1185-1186
    HSR_TRY(haccsr_partition_in_box(c, box_hi, &nin));              // :1243-1248 (resortParticles, Np = m_Np_last)
    HSR_TRY(haccsr_fill_mass(c, 1.0f));                             // :1256-1257
    HSR_TRY(haccsr_kick(c, nin, tree_lo, tree_hi, force_lo, force_hi, theta, ppn, tdpts, fcoeff, nullptr, &st));
    HSR_TRY(haccsr_stream(c, pt));                                  // :1194-1195
    const uint64_t pairs = sum.pairs_evaluated + st.pairs_evaluated;
    const float b = sum.ms_build + st.ms_build, w = sum.ms_walk + st.ms_walk, f = sum.ms_force + st.ms_force,
                t = sum.ms_total + st.ms_total;
    const int fl = sum.force_launches + st.force_launches, tl = sum.total_launches + st.total_launches + 5;
    sum = st;
    sum.pairs_evaluated = pairs; sum.ms_build = b; sum.ms_walk = w; sum.ms_force = f; sum.ms_total = t;
    sum.force_launches = fl; sum.total_launches = tl;
  }
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) *stats = sum;
  return 0;
}

int haccsr_map2_setup(const int32_t nglt[3], float edge, float gpscal, double fscal, double tau, double step_fraction,
                      haccsr_map2 *out) {
  if (!nglt || !out) { set_error("haccsr_map2_setup: null argument"); return 1; }
  int mx = nglt[0] > nglt[1] ? nglt[0] : nglt[1];
  if (nglt[2] > mx) mx = nglt[2];
  for (int k = 0; k < 3; ++k) {
    out->tree_lo[k] = 0.0f;
    out->tree_hi[k] = (float)(1.0 * mx);                              // Particles.cxx:1214-1215
    out->force_lo[k] = edge;                                          // :1219-1221
    out->force_hi[k] = (float)(1.0 * nglt[k] - (double)edge);         // :1222-1224 (double arithmetic, stored in POSVEL_T)
  }
  const float divscal = gpscal * gpscal * gpscal;                     // :1231 (float products)
  const float pi = (float)(4.0 * (double)atanf(1.0f));                // :1232 (POSVEL_T pi)
  out->fcoeff = (float)((double)divscal / 4.0 / (double)pi * fscal * tau * step_fraction);     // :1233
  return 0;
}

float haccsr_map1_factor(float pp, float tau, float adot, float alpha) {
  const float pf = powf(pp, (float)(1.0 + 1.0 / (double)alpha));                       // Particles.cxx:745
  const float prefactor = (float)(1.0 / (double)(alpha * adot * pf));                  // :746 (float product, double division)
  return prefactor * tau;                                                              // :752 (float)
}

int haccsr_particles_subcycle(haccsr_ctx *c, int nsub, const int32_t nglt[3], float edge, float gpscal, float alpha, double pp,
                              double adot, double tau, double tau2, double fscal, float theta, int64_t ppn, int tdpts,
                              haccsr_stats *stats) {
  if (!c) { set_error("null context"); return 1; }
  if (nsub < 1 || !nglt) { set_error("haccsr_particles_subcycle: bad argument"); return 1; }
  const double step_fraction = 1.0 / nsub;                                              // Particles.cxx:1180
  haccsr_map2 m;
  HSR_TRY(haccsr_map2_setup(nglt, edge, gpscal, fscal, tau, step_fraction, &m));
  // map1(gts->pp(), stepFraction*gts->tau2(), gts->adot()): the three arguments are converted to float at the call (:1186)
  const float pt = haccsr_map1_factor((float)pp, (float)(step_fraction * tau2), (float)adot, alpha);
  const float box_hi[3] = {(float)nglt[0], (float)nglt[1], (float)nglt[2]};
  return haccsr_subcycle(c, nsub, pt, box_hi, m.tree_lo, m.tree_hi, m.force_lo, m.force_hi, theta, ppn, tdpts, m.fcoeff, stats);
}

int haccsr_get_tree(haccsr_ctx *c, int64_t cap, int64_t *nodes, int32_t *count, int32_t *offset, int32_t *cl,
                    int32_t *cr, float *box10) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  if (nodes) *nodes = c->n_nodes;
  if (cap == 0) return 0;
  if (cap < c->n_nodes) { set_error("haccsr_get_tree: cap %lld < nodes %d", (long long)cap, c->n_nodes); return 1; }
  Node *h = (Node *)malloc((size_t)c->n_nodes * sizeof(Node));
  if (!h) { set_error("out of host memory"); return 2; }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) { free(h); set_error("haccsr_get_tree: stream synchronisation failed"); return 2; }
  cudaError_t e = cudaMemcpy(h, c->nodes.p, (size_t)c->n_nodes * sizeof(Node), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(h); set_error("cudaMemcpy nodes: %s", cudaGetErrorString(e)); return 2; }
  for (int i = 0; i < c->n_nodes; ++i) {
    if (count) count[i] = h[i].count;
    if (offset) offset[i] = h[i].offset;
    if (cl) cl[i] = h[i].cl;
    if (cr) cr[i] = h[i].cr;
    if (box10) {
      for (int k = 0; k < 3; ++k) { box10[10*i + k] = h[i].xmin[k]; box10[10*i + 3 + k] = h[i].xmax[k]; box10[10*i + 6 + k] = h[i].xc[k]; }
      box10[10*i + 9] = h[i].ppm;
    }
  }
  free(h);
  return 0;
}

int haccsr_get_pseudo_particles(haccsr_ctx *c, int64_t cap_nodes, float *pp /* 48 per node */) {
  if (!c) { set_error("null context"); return 1; }
  if (c->tdpts != 12) { set_error("haccsr_get_pseudo_particles: the last kick was not a quadrupole (tdpts = 12) kick"); return 1; }
  if (cap_nodes < c->n_nodes || !pp) { set_error("haccsr_get_pseudo_particles: buffer too small"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  HSR_CUDA(cudaMemcpy(pp, c->pp12.p, 12 * (size_t)c->n_nodes * sizeof(float4), cudaMemcpyDeviceToHost));
  return 0;
}

int haccsr_get_lists(haccsr_ctx *c, int64_t cap_nodes, int64_t cap_ranges, int64_t cap_pool, int64_t *n_nodes,
                     int64_t *n_ranges, int64_t *n_pool, uint32_t *range_off, uint32_t *ranges, float *pool) {
  if (!c) { set_error("null context"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  if (n_nodes) *n_nodes = c->n_nodes;
  if (n_ranges) *n_ranges = c->tot_ranges;
  if (n_pool) *n_pool = c->tot_pseudo;
  if (cap_nodes == 0 && cap_ranges == 0 && cap_pool == 0) return 0;
  if (cap_nodes < c->n_nodes + 1 || cap_ranges < c->tot_ranges || cap_pool < c->tot_pseudo) { set_error("haccsr_get_lists: buffers too small"); return 1; }
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  if (range_off) HSR_CUDA(cudaMemcpy(range_off, c->range_off.p, (size_t)(c->n_nodes + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (ranges && c->tot_ranges) HSR_CUDA(cudaMemcpy(ranges, c->ranges.p, (size_t)c->tot_ranges * sizeof(uint2), cudaMemcpyDeviceToHost));
  if (pool && c->tot_pseudo) HSR_CUDA(cudaMemcpy(pool, c->pool.p, (size_t)c->tot_pseudo * sizeof(float4), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
