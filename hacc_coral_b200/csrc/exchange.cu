// exchange.cu -- the overload (ghost-zone) refresh of one rank as ONE call of the C ABI, transport included.
//
// Replaces ParticleExchange::exchangeParticles as MC3Extras::refreshParticles drives it at refresh steps
// (reference src/simulation/MC3Extras.cxx:660-706; src/halo_finder/ParticleExchange.cxx:488-762) over the periodic Cartesian
// rank layout of Partition (src/halo_finder/Partition.cxx:121-137; rank = (px*ny + py)*nz + pz, MPI_Cart_create order).  The
// reference runs 13 paired MPI send/recv rounds with a barrier each (:630-632,729); here the device classifies and packs all
// 26 messages into one buffer ordered by destination rank (refresh.cu), ONE grouped ncclSend/ncclRecv moves it over
// NVLink (every peer is one NVSwitch hop, so nothing is gained by pairing rounds), and one kernel appends what arrived.
// Messages whose destination is the rank itself (periodic wrap on an axis of extent 1, ParticleExchange.cxx:676-695) never
// leave the GPU.  Host work per refresh: the plan (a few hundred integer operations), one read-back of the gathered
// count table, the launches.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, whichever copy the process already has): libhaccsr.so itself has no
// link-time dependency on it, and a context that never refreshes across ranks never needs it.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace haccsr {

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclCommCount) CommCount = nullptr;
  decltype(&ncclCommUserRank) CommUserRank = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 = not tried, 1 = ok, -1 = unavailable
  if (state == 1) return &api;
  if (state == -1) { set_error("NCCL is not available (libnccl.so.2 could not be loaded); there is no other transport"); return nullptr; }
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) { state = -1; set_error("NCCL is not available (dlopen libnccl.so.2: %s); there is no other transport", dlerror()); return nullptr; }
#define HSR_SYM(field, name)                                                            \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));          \
  if (!api.field) { state = -1; set_error("libnccl lacks %s", name); return nullptr; }
  HSR_SYM(GetUniqueId, "ncclGetUniqueId") HSR_SYM(CommInitRank, "ncclCommInitRank") HSR_SYM(CommDestroy, "ncclCommDestroy")
  HSR_SYM(CommCount, "ncclCommCount") HSR_SYM(CommUserRank, "ncclCommUserRank") HSR_SYM(GroupStart, "ncclGroupStart")
  HSR_SYM(GroupEnd, "ncclGroupEnd") HSR_SYM(Send, "ncclSend") HSR_SYM(Recv, "ncclRecv") HSR_SYM(AllGather, "ncclAllGather")
  HSR_SYM(GetErrorString, "ncclGetErrorString")
#undef HSR_SYM
  state = 1;
  return &api;
}

#define HSR_NCCL(api, call)                                                                              \
  do {                                                                                                   \
    ncclResult_t r__ = (call);                                                                           \
    if (r__ != ncclSuccess) {                                                                            \
      set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(r__));        \
      return 4;                                                                                          \
    }                                                                                                    \
  } while (0)

// ---- the Cartesian plan (host; mirrors Partition's neighbour table, Partition.cxx:140-260) -------------------------------
struct Cart {
  int dims[3];
  int size() const { return dims[0] * dims[1] * dims[2]; }
  void position(int rank, int p[3]) const { p[0] = rank / (dims[1] * dims[2]); p[1] = (rank / dims[2]) % dims[1]; p[2] = rank % dims[2]; }
  int rank_of(const int p[3]) const {
    int q[3];
    for (int k = 0; k < 3; ++k) q[k] = ((p[k] % dims[k]) + dims[k]) % dims[k];
    return (q[0] * dims[1] + q[1]) * dims[2] + q[2];
  }
  int neighbor(int rank, int d) const {
    int p[3];
    position(rank, p);
    p[0] += d / 9 - 1; p[1] += (d / 3) % 3 - 1; p[2] += d % 3 - 1;
    return rank_of(p);
  }
};
// message order of a rank: slot s = position of direction d when the 26 directions are sorted by (destination rank, d), so the
// messages for one destination are contiguous in the send buffer and both sides derive the same layout from the counts alone
struct Plan {
  int order[26];        // slot -> direction
  int dest[26];         // slot -> destination rank
  int slot_of_dir[27];
  Plan(const Cart &c, int rank) {
    int n = 0;
    for (int d = 0; d < 27; ++d) if (d != 13) order[n++] = d;
    std::stable_sort(order, order + 26, [&](int a, int b) { return c.neighbor(rank, a) < c.neighbor(rank, b); });
    for (int d = 0; d < 27; ++d) slot_of_dir[d] = 26;
    for (int s = 0; s < 26; ++s) { slot_of_dir[order[s]] = s; dest[s] = c.neighbor(rank, order[s]); }
  }
};

struct AppendEntry { long long byte_off, n, at; };   // message in the receive buffer, its particle count, first index in the arrays

// one launch appends every received message: blockIdx.y = message
__global__ void __launch_bounds__(256) k_append_all(Soa p, const unsigned char *__restrict__ recv, const AppendEntry *__restrict__ tab) {
  const AppendEntry e = tab[blockIdx.y];
  const long long n = e.n;
  const unsigned char *msg = recv + e.byte_off;
  const int64_t *id = reinterpret_cast<const int64_t *>(msg);
  const float *f = reinterpret_cast<const float *>(msg + n * 8);
  const uint16_t *mk = reinterpret_cast<const uint16_t *>(msg + n * 8 + 8 * n * 4);
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long i = e.at + r;
    p.x[i] = f[0 * n + r]; p.y[i] = f[1 * n + r]; p.z[i] = f[2 * n + r];
    p.vx[i] = f[3 * n + r]; p.vy[i] = f[4 * n + r]; p.vz[i] = f[5 * n + r];
    p.mass[i] = f[6 * n + r]; p.phi[i] = f[7 * n + r];
    p.id[i] = id[r]; p.mask[i] = mk[r];
  }
}

}  // namespace haccsr

using namespace haccsr;

extern "C" {

int haccsr_refresh_plan(const int32_t dims[3], int32_t rank, int32_t dir_of_slot[26], int32_t dest_of_slot[26]) {
  if (!dims || !dir_of_slot || !dest_of_slot) { set_error("haccsr_refresh_plan: null argument"); return 1; }
  Cart cart;
  for (int k = 0; k < 3; ++k) { cart.dims[k] = dims[k]; if (dims[k] < 1) { set_error("haccsr_refresh_plan: bad decomposition"); return 1; } }
  if (rank < 0 || rank >= cart.size()) { set_error("haccsr_refresh_plan: rank outside the decomposition"); return 1; }
  const Plan p(cart, rank);
  for (int q = 0; q < 26; ++q) { dir_of_slot[q] = p.order[q]; dest_of_slot[q] = p.dest[q]; }
  return 0;
}

int haccsr_nccl_unique_id(void *id128) {
  if (!id128) { set_error("haccsr_nccl_unique_id: null buffer"); return 1; }
  NcclApi *api = nccl_api();
  if (!api) return 3;
  static_assert(sizeof(ncclUniqueId) == HACCSR_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  HSR_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int haccsr_nccl_comm_create(void **comm, int device, int nranks, int rank, const void *id128) {
  if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("haccsr_nccl_comm_create: bad argument"); return 1; }
  NcclApi *api = nccl_api();
  if (!api) return 3;
  HSR_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t cm = nullptr;
  HSR_NCCL(api, api->CommInitRank(&cm, nranks, id, rank));
  *comm = (void *)cm;
  return 0;
}

int haccsr_nccl_comm_destroy(void *comm) {
  if (!comm) return 0;
  NcclApi *api = nccl_api();
  if (!api) return 3;
  HSR_NCCL(api, api->CommDestroy((ncclComm_t)comm));
  return 0;
}

int haccsr_refresh(haccsr_ctx *c, void *nccl_comm, const int32_t dims[3], int32_t rank, const float alive_lo[3],
                   const float alive_hi[3], float ol, haccsr_refresh_stats *stats) {
  if (!c) { set_error("null context"); return 1; }
  if (!dims || !alive_lo || !alive_hi) { set_error("haccsr_refresh: null argument"); return 1; }
  Cart cart;
  for (int k = 0; k < 3; ++k) { cart.dims[k] = dims[k]; if (dims[k] < 1) { set_error("haccsr_refresh: bad decomposition"); return 1; } }
  const int nranks = cart.size();
  if (rank < 0 || rank >= nranks) { set_error("haccsr_refresh: rank %d outside the %d x %d x %d decomposition", rank, dims[0], dims[1], dims[2]); return 1; }
  NcclApi *api = nullptr;
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  if (nranks > 1) {
    if (!comm) { set_error("haccsr_refresh: a decomposition of %d ranks needs an NCCL communicator", nranks); return 1; }
    api = nccl_api();
    if (!api) return 3;
    int cn = 0, cr = -1;
    HSR_NCCL(api, api->CommCount(comm, &cn));
    HSR_NCCL(api, api->CommUserRank(comm, &cr));
    if (cn != nranks || cr != rank) { set_error("haccsr_refresh: communicator is rank %d of %d, expected %d of %d", cr, cn, rank, nranks); return 1; }
  }
  HSR_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  cudaEvent_t e0 = c->ev[0], e1 = c->ev[1];
  HSR_CUDA(cudaEventRecord(e0, s));
  const Plan mine(cart, rank);
  // 1. drop the ghosts, classify the alive particles against the 26 slabs, message sizes (refresh.cu)
  int64_t counts[27], n_alive = 0;
  HSR_TRY(haccsr_refresh_begin(c, alive_lo, alive_hi, ol, mine.slot_of_dir, counts, &n_alive));
  // 2. every rank's row -- its 26 counts, its alive count and its capacity -- to every rank (tiny all-gather on the device),
  //    then one read-back.  With the last two every rank sees which ranks cannot hold their ghosts, so that all of them
  //    leave the call with the same status instead of some waiting in the exchange for a peer that has already returned.
  constexpr int ROW = 28;
  std::vector<long long> table((size_t)nranks * ROW);
  HSR_TRY(c->xchg_table.ensure((size_t)nranks * ROW + 32 + 3 * 26 * (size_t)nranks + 64));
  long long *d_table = c->xchg_table.p;                  // [nranks][ROW]
  long long *d_mine = d_table + (size_t)nranks * ROW;    // [ROW]
  AppendEntry *d_app = reinterpret_cast<AppendEntry *>(d_mine + 32);
  long long row[ROW];
  for (int q = 0; q < 26; ++q) row[q] = counts[q];
  row[26] = n_alive; row[27] = c->cap;
  if (nranks > 1) {
    HSR_CUDA(cudaMemcpyAsync(d_mine, row, ROW * sizeof(long long), cudaMemcpyHostToDevice, s));
    HSR_NCCL(api, api->AllGather(d_mine, d_table, ROW, ncclInt64, comm, s));
    HSR_CUDA(cudaMemcpyAsync(table.data(), d_table, table.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
    HSR_CUDA(cudaStreamSynchronize(s));
  } else {
    for (int q = 0; q < ROW; ++q) table[q] = row[q];
  }
  // 3. layouts: my messages sorted by destination; what each rank sends to me, in the order it sits in that rank's chunk
  int64_t off[27];
  std::vector<long long> send_bytes(nranks, 0), send_off(nranks + 1, 0), recv_bytes(nranks, 0), recv_off(nranks + 1, 0);
  long long pos = 0, sent = 0;
  for (int q = 0; q < 26; ++q) {
    off[q] = pos;
    const long long b = haccsr_refresh_message_bytes(counts[q]);
    pos += b; send_bytes[mine.dest[q]] += b; sent += counts[q];
  }
  off[26] = 0;
  const long long total_send = pos;
  for (int r = 0; r < nranks; ++r) send_off[r + 1] = send_off[r] + send_bytes[r];
  std::vector<AppendEntry> app;
  std::vector<long long> incoming(nranks, 0);     // ghosts every rank is about to receive
  long long rpos = 0, at = n_alive, ghosts = 0;
  for (int r = 0; r < nranks; ++r) {
    const Plan theirs(cart, r);
    recv_off[r] = rpos;
    for (int q = 0; q < 26; ++q) {
      const long long n = table[(size_t)r * ROW + q];
      incoming[theirs.dest[q]] += n;
      if (theirs.dest[q] != rank) continue;
      if (n > 0) { app.push_back({rpos, n, at}); at += n; ghosts += n; }
      rpos += haccsr_refresh_message_bytes(n);
    }
    recv_bytes[r] = rpos - recv_off[r];
  }
  recv_off[nranks] = rpos;
  const long long total_recv = rpos;
  for (int r = 0; r < nranks; ++r) {
    const long long alive_r = table[(size_t)r * ROW + 26], cap_r = table[(size_t)r * ROW + 27];
    if (alive_r + incoming[r] > cap_r) {      // the same verdict on every rank; the contexts keep their alive particles
      set_error("haccsr_refresh: %lld alive + %lld ghosts exceed the context capacity %lld on rank %d", alive_r, incoming[r], cap_r, r);
      return 1;
    }
  }
  HSR_TRY(c->xchg_send.ensure((size_t)total_send + 16));
  HSR_TRY(c->xchg_recv.ensure((size_t)total_recv + 16));
  // 4. pack, exchange, append -- all on the context's stream
  HSR_TRY(refresh_pack_async(c, off, c->xchg_send.p));
  if (send_bytes[rank] > 0)
    HSR_CUDA(cudaMemcpyAsync(c->xchg_recv.p + recv_off[rank], c->xchg_send.p + send_off[rank], (size_t)send_bytes[rank], cudaMemcpyDeviceToDevice, s));
  if (nranks > 1) {
    HSR_NCCL(api, api->GroupStart());
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) continue;
      if (send_bytes[r] > 0) HSR_NCCL(api, api->Send(c->xchg_send.p + send_off[r], (size_t)send_bytes[r], ncclUint8, r, comm, s));
      if (recv_bytes[r] > 0) HSR_NCCL(api, api->Recv(c->xchg_recv.p + recv_off[r], (size_t)recv_bytes[r], ncclUint8, r, comm, s));
    }
    HSR_NCCL(api, api->GroupEnd());
  }
  if (!app.empty()) {
    HSR_CUDA(cudaMemcpyAsync(d_app, app.data(), app.size() * sizeof(AppendEntry), cudaMemcpyHostToDevice, s));
    long long nmax = 0;
    for (const AppendEntry &e : app) nmax = std::max(nmax, e.n);
    long long gx = (nmax + 255) / 256;
    if (gx > 4 * c->sm_count) gx = 4 * c->sm_count;
    k_append_all<<<dim3((unsigned)gx, (unsigned)app.size()), 256, 0, s>>>(c->cur, c->xchg_recv.p, d_app);
    HSR_CUDA(cudaGetLastError());
  }
  HSR_CUDA(cudaEventRecord(e1, s));
  HSR_CUDA(cudaStreamSynchronize(s));     // app (host vector) must outlive the copy; the caller gets a finished refresh
  c->n_resident = at;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->alive = n_alive; stats->ghosts = ghosts; stats->sent = sent;
    stats->bytes_sent = total_send; stats->bytes_received = total_recv;
    stats->bytes_sent_remote = total_send - send_bytes[rank];
    stats->messages_received = (int32_t)app.size();
    HSR_CUDA(cudaEventElapsedTime(&stats->ms_total, e0, e1));
  }
  return 0;
}

}  // extern "C"
