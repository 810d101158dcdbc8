// refresh.cu -- device side of the overload (ghost-zone) refresh, sm_100a.
//
// What it replaces: the per-rank part of ParticleExchange::exchangeParticles as MC3Extras::refreshParticles
// drives it at refresh steps (reference src/simulation/MC3Extras.cxx:660-706): keep the alive particles
// (Particles::copyAliveIntoVectors, src/cpu/Particles.cxx:958-975), find the alive particles that are ghosts of
// each of the 26 neighbours (identifyExchangeParticles, src/halo_finder/ParticleExchange.cxx:542-574, slabs from
// calculateExchangeRegions :280-450), pack them per neighbour with the periodic shift applied to the position
// (exchange, :650-762), and append what the neighbours sent.  The transport between GPUs (one NCCL
// all-to-all-v of the packed buffer, or a device-local copy where the neighbour is the rank itself) is the
// host's job: hacc_coral_b200/refresh.py.
//
// Coordinates are the LOCAL grid units the force tree works in: a rank's alive region is [alo, ahi) and its
// overload shell is `ol` wide, so a particle sent towards direction s = (sx,sy,sz) arrives at x - s*(ahi-alo)
// in the receiver's frame (uniform decomposition: every rank has the same alive extent) -- the reference does
// the same thing in global box units plus a +-boxSize wrap (:672-673), which is this shift.
//
// Deterministic by construction: messages keep the senders' particle order (ballot/popc ranks, scans; no
// atomics on cursors), so a refresh followed by a kick is reproducible run to run.
#include "common.cuh"

namespace haccsr {

static constexpr int RT = 256;   // candidates per tile (one block of 8 warps)

struct RefreshGeom {
  float alo[3], ahi[3];   // alive region, local grid units
  float ol;               // overload width
  float ext[3];           // ahi - alo
};

// membership of one particle in the 27 direction slabs as a bit mask (bit d = (sx+1)*9 + (sy+1)*3 + (sz+1));
// comparisons are inclusive on both ends like ParticleExchange.cxx:560-565
__device__ __forceinline__ unsigned dir_mask(float x, float y, float z, const RefreshGeom &G) {
  const float p[3] = {x, y, z};
  unsigned low = 0, high = 0, full = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = G.alo[k], hi = G.ahi[k], mlo = __fadd_rn(lo, G.ol), mhi = __fsub_rn(hi, G.ol);
    if (p[k] >= lo && p[k] <= mlo) low |= 1u << k;
    if (p[k] >= mhi && p[k] <= hi) high |= 1u << k;
    if (p[k] >= lo && p[k] <= hi) full |= 1u << k;
  }
  unsigned m = 0;
#pragma unroll
  for (int d = 0; d < 27; ++d) {
    if (d == 13) continue;
    const int s[3] = {d / 9 - 1, (d / 3) % 3 - 1, d % 3 - 1};
    bool in = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned b = 1u << k;
      in = in && (s[k] < 0 ? (low & b) : (s[k] > 0 ? (high & b) : (full & b))) != 0;
    }
    if (in) m |= 1u << d;
  }
  return m;
}

// alive = inside [alo, ahi) in every dimension (Particles.cxx:975)
__global__ void __launch_bounds__(256) k_alive_flags(const float *__restrict__ x, const float *__restrict__ y,
                                                     const float *__restrict__ z, RefreshGeom G, long long n,
                                                     unsigned *__restrict__ flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool a = x[i] >= G.alo[0] && x[i] < G.ahi[0] && y[i] >= G.alo[1] && y[i] < G.ahi[1] && z[i] >= G.alo[2] && z[i] < G.ahi[2];
    flag[i] = a ? 1u : 0u;
  }
}

// shared = not strictly inside the inner region (ParticleExchange.cxx:549-556)
__global__ void __launch_bounds__(256) k_shared_flags(const float *__restrict__ x, const float *__restrict__ y,
                                                      const float *__restrict__ z, RefreshGeom G, long long n,
                                                      unsigned *__restrict__ flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p[3] = {x[i], y[i], z[i]};
    bool inner = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) inner = inner && (p[k] > __fadd_rn(G.alo[k], G.ol)) && (p[k] < __fsub_rn(G.ahi[k], G.ol));
    flag[i] = inner ? 0u : 1u;
  }
}

__global__ void __launch_bounds__(256) k_collect(const unsigned *__restrict__ flag, const unsigned *__restrict__ pref,
                                                 long long n, unsigned *__restrict__ cand) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (flag[i]) cand[pref[i]] = (unsigned)i;
}

// per tile of RT candidates and per slot: how many candidates belong to the slot's direction
__global__ void __launch_bounds__(RT) k_dir_count(const float *__restrict__ x, const float *__restrict__ y,
                                                  const float *__restrict__ z, const unsigned *__restrict__ cand, int m,
                                                  RefreshGeom G, const int *__restrict__ slot_of_dir, int ntiles,
                                                  unsigned *__restrict__ tilecount) {
  __shared__ unsigned s_cnt[27];
  const int t = threadIdx.x, c = blockIdx.x * RT + t;
  if (t < 27) s_cnt[t] = 0;
  __syncthreads();
  unsigned mask = 0;
  if (c < m) { const unsigned i = cand[c]; mask = dir_mask(x[i], y[i], z[i], G); }
#pragma unroll 1
  for (int d = 0; d < 27; ++d) {
    if (d == 13) continue;
    const unsigned b = __ballot_sync(0xffffffffu, (mask >> d) & 1u);
    if ((t & 31) == 0 && b) atomicAdd(&s_cnt[d], (unsigned)__popc(b));
  }
  __syncthreads();
  if (t < 27 && t != 13) tilecount[(size_t)slot_of_dir[t] * ntiles + blockIdx.x] = s_cnt[t];
}

__global__ void k_slot_counts(const unsigned *__restrict__ tilebase, int ntiles, const unsigned long long *__restrict__ total,
                              long long *__restrict__ out) {
  const int sl = threadIdx.x;
  if (sl >= 26) return;
  const long long b = tilebase[(size_t)sl * ntiles], nx = (sl == 25) ? (long long)*total : (long long)tilebase[(size_t)(sl + 1) * ntiles];
  out[sl] = nx - b;
}

struct PackLayout {
  long long byte_off[27];   // per SLOT: start of the message in the send buffer
  long long count[27];      // per SLOT: particles in the message
};

// message layout for n particles: id[n] (8 B) | x y z vx vy vz mass phi [n] each (4 B) | mask[n] (2 B), padded to 16 B
__host__ __device__ __forceinline__ long long message_bytes(long long n) {
  long long b = n * 8 + 8 * n * 4 + n * 2;
  return (b + 15) & ~15ll;
}

__global__ void __launch_bounds__(RT) k_dir_pack(Soa p, const unsigned *__restrict__ cand, int m, RefreshGeom G,
                                                 const int *__restrict__ slot_of_dir, int ntiles,
                                                 const unsigned *__restrict__ tilebase, PackLayout L,
                                                 unsigned char *__restrict__ sendbuf) {
  __shared__ unsigned s_warp[27][RT / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, c = blockIdx.x * RT + t;
  unsigned mask = 0, i = 0;
  float x = 0.f, y = 0.f, z = 0.f;
  if (c < m) { i = cand[c]; x = p.x[i]; y = p.y[i]; z = p.z[i]; mask = dir_mask(x, y, z, G); }
#pragma unroll 1
  for (int d = 0; d < 27; ++d) {
    const unsigned b = __ballot_sync(0xffffffffu, (mask >> d) & 1u);
    if (lane == 0) s_warp[d][w] = (unsigned)__popc(b);
  }
  __syncthreads();
  float vx = 0.f, vy = 0.f, vz = 0.f, ms = 0.f, ph = 0.f;
  int64_t id = 0;
  uint16_t mk = 0;
  if (mask) { vx = p.vx[i]; vy = p.vy[i]; vz = p.vz[i]; ms = p.mass[i]; ph = p.phi[i]; id = p.id[i]; mk = p.mask[i]; }
#pragma unroll 1
  for (int d = 0; d < 27; ++d) {           // warp-uniform loop: every lane takes part in the ballot
    const unsigned b = __ballot_sync(0xffffffffu, (mask >> d) & 1u);
    if (!((mask >> d) & 1u)) continue;
    const int s = slot_of_dir[d];
    unsigned before = 0;
    for (int q = 0; q < w; ++q) before += s_warp[d][q];
    // position inside the message: candidates of earlier tiles, earlier warps, earlier lanes
    const long long first = (long long)tilebase[(size_t)s * ntiles];
    const long long r = (long long)tilebase[(size_t)s * ntiles + blockIdx.x] - first + before + (unsigned)__popc(b & ((1u << lane) - 1u));
    const long long n = L.count[s];
    unsigned char *msg = sendbuf + L.byte_off[s];
    const int sx = d / 9 - 1, sy = (d / 3) % 3 - 1, sz = d % 3 - 1;
    reinterpret_cast<int64_t *>(msg)[r] = id;
    float *f = reinterpret_cast<float *>(msg + n * 8);
    f[0 * n + r] = __fsub_rn(x, (float)sx * G.ext[0]);
    f[1 * n + r] = __fsub_rn(y, (float)sy * G.ext[1]);
    f[2 * n + r] = __fsub_rn(z, (float)sz * G.ext[2]);
    f[3 * n + r] = vx; f[4 * n + r] = vy; f[5 * n + r] = vz; f[6 * n + r] = ms; f[7 * n + r] = ph;
    reinterpret_cast<uint16_t *>(msg + n * 8 + 8 * n * 4)[r] = mk;
  }
}

__global__ void __launch_bounds__(256) k_append(Soa p, long long at, const unsigned char *__restrict__ msg, long long n) {
  const float *f = reinterpret_cast<const float *>(msg + n * 8);
  const int64_t *id = reinterpret_cast<const int64_t *>(msg);
  const uint16_t *mk = reinterpret_cast<const uint16_t *>(msg + n * 8 + 8 * n * 4);
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long i = at + r;
    p.x[i] = f[0 * n + r]; p.y[i] = f[1 * n + r]; p.z[i] = f[2 * n + r];
    p.vx[i] = f[3 * n + r]; p.vy[i] = f[4 * n + r]; p.vz[i] = f[5 * n + r];
    p.mass[i] = f[6 * n + r]; p.phi[i] = f[7 * n + r];
    p.id[i] = id[r]; p.mask[i] = mk[r];
  }
}

static int lin_grid2(const haccsr_ctx *c, int64_t n) {
  int64_t g = (n + 255) / 256, cap = (int64_t)c->sm_count * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

static RefreshGeom make_geom(const float alo[3], const float ahi[3], float ol) {
  RefreshGeom G;
  for (int k = 0; k < 3; ++k) { G.alo[k] = alo[k]; G.ahi[k] = ahi[k]; G.ext[k] = ahi[k] - alo[k]; }
  G.ol = ol;
  return G;
}

int refresh_pack_async(haccsr_ctx *c, const int64_t byte_off_by_slot[27], void *sendbuf_device) {
  if (!c) { set_error("null context"); return 1; }
  if (!byte_off_by_slot) { set_error("haccsr_refresh_pack: null argument"); return 1; }
  if (c->refresh_epoch != c->order_epoch) {
    set_error("haccsr_refresh_pack: the particles were reordered (kick, upload, compaction or append) since haccsr_refresh_begin; "
              "its candidate list no longer describes them");
    return 1;
  }
  if (c->refresh_m == 0) return 0;
  if (!sendbuf_device) { set_error("haccsr_refresh_pack: null send buffer"); return 1; }
  HSR_CUDA(cudaSetDevice(c->device));
  const RefreshGeom G = make_geom(c->refresh_alo, c->refresh_ahi, c->refresh_ol);
  PackLayout L;
  for (int sl = 0; sl < 27; ++sl) { L.byte_off[sl] = sl < 26 ? byte_off_by_slot[sl] : 0; L.count[sl] = sl < 26 ? c->refresh_count[sl] : 0; }
  k_dir_pack<<<c->refresh_ntiles, RT, 0, c->stream>>>(c->cur, c->refresh_cand.p, (int)c->refresh_m, G, c->refresh_slots.p,
                                                      c->refresh_ntiles, c->tilebase.p, L, (unsigned char *)sendbuf_device);
  HSR_CUDA(cudaGetLastError());
  return 0;     // stream-ordered
}


}  // namespace haccsr

using namespace haccsr;

extern "C" {

int64_t haccsr_refresh_message_bytes(int64_t n) { return (int64_t)message_bytes(n); }

int64_t haccsr_resident(haccsr_ctx *c) { return c ? c->n_resident : -1; }

int haccsr_refresh_begin(haccsr_ctx *c, const float alive_lo[3], const float alive_hi[3], float ol,
                         const int32_t slot_of_dir[27], int64_t counts_by_slot[27], int64_t *n_alive) {
  if (!c) { set_error("null context"); return 1; }
  if (!alive_lo || !alive_hi || !slot_of_dir || !counts_by_slot) { set_error("haccsr_refresh_begin: null argument"); return 1; }
  for (int k = 0; k < 3; ++k)
    if (!(ol > 0.f) || !(alive_hi[k] - alive_lo[k] >= 2.f * ol)) { set_error("haccsr_refresh_begin: overload width must be positive and at most half the alive extent"); return 1; }
  bool seen[27] = {false};
  for (int d = 0; d < 27; ++d) {
    if (d == 13) continue;
    if (slot_of_dir[d] < 0 || slot_of_dir[d] >= 26 || seen[slot_of_dir[d]]) { set_error("haccsr_refresh_begin: slot_of_dir must be a permutation of 0..25"); return 1; }
    seen[slot_of_dir[d]] = true;
  }
  HSR_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const RefreshGeom G = make_geom(alive_lo, alive_hi, ol);
  // 1. drop the ghosts: stable compaction of the alive particles to the front
  int64_t n = c->n_resident, nal = 0;
  if (n >= 0x7fffffffll) { set_error("too many particles for 32-bit indexing"); return 1; }
  HSR_TRY(c->idxA.ensure((size_t)n + 1)); HSR_TRY(c->idxB.ensure((size_t)n + 1)); HSR_TRY(c->refresh_cand.ensure((size_t)n + 1));
  if (n > 0) {
    k_alive_flags<<<lin_grid2(c, n), 256, 0, s>>>(c->cur.x, c->cur.y, c->cur.z, G, (long long)n, c->idxA.p);
    HSR_TRY(compact_by_flags(c, c->idxA.p, c->idxB.p, n, &nal));
  }
  c->n_resident = nal;
  if (n_alive) *n_alive = nal;
  // 2. candidates = alive particles within `ol` of a face, in particle order
  int64_t m = 0;
  if (nal > 0) {
    k_shared_flags<<<lin_grid2(c, nal), 256, 0, s>>>(c->cur.x, c->cur.y, c->cur.z, G, (long long)nal, c->idxA.p);
    HSR_TRY(scan_exclusive(c, c->idxA.p, c->idxB.p, nal, c->d_counters + 12));
    HSR_CUDA(cudaMemcpyAsync(c->h_counters + 12, c->d_counters + 12, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    k_collect<<<lin_grid2(c, nal), 256, 0, s>>>(c->idxA.p, c->idxB.p, (long long)nal, c->refresh_cand.p);
    HSR_CUDA(cudaStreamSynchronize(s));
    m = c->h_counters[12];
  }
  c->refresh_m = m;
  c->refresh_epoch = c->order_epoch;
  c->refresh_ntiles = (int)((m + RT - 1) / RT);
  for (int k = 0; k < 3; ++k) { c->refresh_alo[k] = alive_lo[k]; c->refresh_ahi[k] = alive_hi[k]; }
  c->refresh_ol = ol;
  for (int d = 0; d < 27; ++d) { c->refresh_slot_of_dir[d] = slot_of_dir[d]; counts_by_slot[d] = 0; }
  if (m == 0) return 0;
  // 3. per-tile, per-slot counts and their exclusive scan in (slot, tile) order
  const int nt = c->refresh_ntiles;
  HSR_TRY(c->tilecount.ensure((size_t)26 * nt + 1)); HSR_TRY(c->tilebase.ensure((size_t)26 * nt + 1));
  HSR_TRY(c->refresh_slots.ensure(27));
  HSR_CUDA(cudaMemcpyAsync(c->refresh_slots.p, c->refresh_slot_of_dir, 27 * sizeof(int), cudaMemcpyHostToDevice, s));
  k_dir_count<<<nt, RT, 0, s>>>(c->cur.x, c->cur.y, c->cur.z, c->refresh_cand.p, (int)m, G, c->refresh_slots.p, nt, c->tilecount.p);
  HSR_TRY(scan_exclusive(c, c->tilecount.p, c->tilebase.p, (int64_t)26 * nt, c->d_counters + 13));
  // message sizes = differences of the scan at slot boundaries (26 values, read back through pinned memory)
  k_slot_counts<<<1, 32, 0, s>>>(c->tilebase.p, nt, c->d_counters + 13, c->d_slotcount);
  HSR_CUDA(cudaMemcpyAsync(c->h_counters, c->d_slotcount, 26 * sizeof(long long), cudaMemcpyDeviceToHost, s));
  HSR_CUDA(cudaStreamSynchronize(s));
  for (int sl = 0; sl < 26; ++sl) { counts_by_slot[sl] = c->h_counters[sl]; c->refresh_count[sl] = c->h_counters[sl]; }
  return 0;
}

int haccsr_refresh_pack(haccsr_ctx *c, const int64_t byte_off_by_slot[27], void *sendbuf_device) {
  HSR_TRY(refresh_pack_async(c, byte_off_by_slot, sendbuf_device));
  if (c && c->refresh_m) HSR_CUDA(cudaStreamSynchronize(c->stream));     // the caller's transport may run on any stream
  return 0;
}

int haccsr_refresh_append(haccsr_ctx *c, const void *message_device, int64_t n) {
  if (!c) { set_error("null context"); return 1; }
  if (n < 0) { set_error("haccsr_refresh_append: negative count"); return 1; }
  if (n == 0) return 0;
  if (!message_device) { set_error("haccsr_refresh_append: null message"); return 1; }
  if (c->n_resident + n > c->cap) {
    set_error("haccsr_refresh_append: %lld resident + %lld received exceed the context capacity %lld", (long long)c->n_resident,
              (long long)n, (long long)c->cap);
    return 1;
  }
  HSR_CUDA(cudaSetDevice(c->device));
  k_append<<<lin_grid2(c, n), 256, 0, c->stream>>>(c->cur, (long long)c->n_resident, (const unsigned char *)message_device, (long long)n);
  HSR_CUDA(cudaGetLastError());
  HSR_CUDA(cudaStreamSynchronize(c->stream));
  c->n_resident += n;
  return 0;      // appended particles take new indices; the alive ones keep theirs, so a begin / pack pair stays valid
}

}  // extern "C"
