/* include/haccsr.h -- C ABI of libhaccsr.so: HACC's short-range RCB-tree force on one B200.
 *
 * This is the drop-in boundary for the reference's short-range hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the reference root).  Plain pointers and sizes
 * only; all particle arrays use the reference's build types (-DID_64 -DPOSVEL_32 -DGRID_32,
 * src/halo_finder/include.mk:4 => POSVEL_T=float, ID_T=int64_t, MASK_T=uint16_t,
 * src/halo_finder/BasicDefinition.h:64-90).
 *
 * Life cycle (what RCBForceTree<1>'s constructor does in one call, src/halo_finder/RCBForceTree.cxx:335-450,
 * split so particles can stay resident on the GPU across the nsub sub-cycles of Particles::subCycle,
 * src/cpu/Particles.cxx:1176-1201):
 *
 *   haccsr_create -> haccsr_set_force_law -> haccsr_upload
 *        -> nsub x [ haccsr_stream, haccsr_kick, haccsr_stream ] -> haccsr_download -> haccsr_destroy
 *
 * There is no CPU fallback: every call fails with a non-zero status if no sm_100 device is usable.
 * All functions return 0 on success; on failure haccsr_last_error() describes the problem.
 * A context is bound to one device and must be used from one thread at a time.
 */
#ifndef HACCSR_H
#define HACCSR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HACCSR_VERSION 3

typedef struct haccsr_ctx haccsr_ctx;

/* Force-law kinds.  Replaces the ForceLaw* functor argument of the tree constructor
 * (src/halo_finder/ForceLaw.h:96-127). */
enum {
  /* ForceLawSR over a polynomial grid force: f(r2) = (r2+rsm^2)^-3/2 - sum_k a[k] r2^k, for r2 < rmax^2
   * (ForceLaw.cxx:137-141,187-192; BGQStep16.c:167-187).  ncoef <= 7. */
  HACCSR_LAW_SR_POLY = 0,
  /* ForceLawSR over the analytic grid-force fit FGridEvalFit (ForceLaw.cxx:39-51,70-80): coeffs = the eight
   * constants b c d e f g h l of FGrid (ForceLaw.cxx:23-31), ncoef = 8.  The default of run_hacc.sh (no -P). */
  HACCSR_LAW_SR_FIT = 1,
  /* ForceLawSR over FGridEvalInterp (ForceLaw.cxx:145-172): coeffs = the grid force tabulated at
   * r2 = i * rmax^2 / (ncoef - 1), i = 0 .. ncoef-1 (2 <= ncoef <= 4096), linear interpolation in r2. */
  HACCSR_LAW_SR_INTERP = 2,
  /* ForceLawNewton: f(r2) = r2^-3/2 (ForceLaw.h:102-107), used when the caller passes fl == NULL
   * (RCBForceTree.cxx:395-404); pairs with r2 == 0 contribute nothing (the guard of BGQStep16.c:183). */
  HACCSR_LAW_NEWTON = 3
};

/* Arithmetic of the pair kernel for the polynomial law (the other laws always use the X86 order).
 *   FUSED (default): multiply-adds contracted as in the reference's production kernel -- the QPX loop builds r2
 *          from three fused multiply-adds (src/halo_finder/BGQStep16.c:76-86) and any FMA-capable build of nbody1
 *          (RCBForceTree.cxx:608, gcc's default -ffp-contract=fast) contracts the same expression.  Here the chain
 *          is seeded with rsm^2, s = fma(dz,dz, fma(dy,dy, fma(dx,dx, rsm^2))), the polynomial is evaluated in s
 *          (re-expanded in double) and the cutoff is s < rmax^2 + rsm^2: 17 instead of 20 FMA-pipe operations a pair.
 *          Accelerations agree with the x86-64 reference build within FP32 rounding (tests: 1e-5 of the gross sum);
 *          pairs within one ulp of the cutoff may fall on the other side (their force is ~1e-5 of a typical pair's).
 *   X86:   r2 = (dx*dx + dy*dy) + dz*dz unfused, in the order of the x86-64 build of the reference: the set of
 *          pairs inside the cutoff is bit-identical to that build's.
 *   FUSED_RS3: FUSED with (r2 + rsm^2)^-3/2 formed as rsqrt(s*s*s) (the x86 order of that step) instead of rsqrt(s)^3: one more
 *          FMA-pipe operation a pair (17 instead of 16), and the MUFU.RSQ error enters once instead of three times, so the
 *          distance to the FP64 sum of the same pairs comes down to the CPU reference's own (profiles/parity_r2.json).
 *          Same set of in-cutoff pairs as FUSED. */
enum { HACCSR_ARITH_FUSED = 0, HACCSR_ARITH_X86 = 1, HACCSR_ARITH_FUSED_RS3 = 2 };

/* Mirrors RCBForceTree::printStats (RCBForceTree.cxx:460-511) plus the measurement fields of
 * SURVEY.md section 8(d). */
typedef struct haccsr_stats {
  int64_t particles;        /* count handed to the kick                                  */
  int64_t nodes;            /* tree.size()                                               */
  int64_t leaves;           /* leaves counted from node 1 on, as printStats does        */
  int64_t empty_leaves;     /* orphan nodes left by degenerate splits                    */
  int64_t max_ppn;          /* largest leaf                                              */
  double  mean_ppn;         /* mean particles per non-empty leaf                         */
  int64_t levels;           /* depth of the tree                                         */
  int64_t sink_leaves;      /* leaves that pass the force-box test (RCBForceTree.cxx:1166-1170) */
  int64_t list_ranges;      /* (start,count) source ranges emitted by the walk           */
  int64_t pseudo_particles; /* accepted monopoles copied into the pseudo-particle pool   */
  int64_t max_list;         /* longest interaction list (sources) of any sink leaf       */
  uint64_t pairs_evaluated; /* sum over sink leaves of count x list length               */
  uint64_t pairs_in_cutoff; /* pairs with 0 < r2 < rmax^2; only when count_in_cutoff != 0 */
  float ms_build;           /* device time: tree build incl. permuting the 10 arrays     */
  float ms_walk;            /* device time: interaction-list construction                */
  float ms_force;           /* device time: leaf-vs-list force kernel + kick             */
  float ms_total;           /* device time of the whole haccsr_kick call                 */
  int32_t force_launches;   /* kernel launches of the force kernel in this call          */
  int32_t total_launches;   /* all kernel launches in this call                          */
  uint64_t pairs_force_law; /* pairs for which the force law was executed; only when count_in_cutoff != 0.  Equal to
                               pairs_evaluated unless warp-level culling is on (haccsr_set_culling)            */
} haccsr_stats;

/* Per-call options of the kick; zero-initialise for defaults. */
typedef struct haccsr_kick_opts {
  int32_t count_in_cutoff;  /* 1: also count pairs with 0 < r2 < rmax^2 (slower kernel variant) */
  int32_t skip_force;       /* 1: build + walk only (used to time the phases separately)        */
  int32_t reserved[6];
} haccsr_kick_opts;

const char *haccsr_last_error(void);

/* Number of CUDA devices with compute capability 10.x visible to this process (0 => unusable). */
int haccsr_device_count(void);

/* Create a context on `device` able to hold `max_particles` particles.
 * Replaces: the allocation side of `new RCBMonopoleForceTree(...)` (src/cpu/Particles.cxx:1313) and the
 * bigchunk node pool (src/halo_finder/bigchunk.h:49-139). */
int haccsr_create(haccsr_ctx **out, int device, int64_t max_particles);
int haccsr_destroy(haccsr_ctx *ctx);

/* Run all work of this context on the given CUDA stream (a cudaStream_t passed as void*; NULL = the
 * context's own stream).  Lets a host framework time the kernels with events on its own stream. */
int haccsr_set_stream(haccsr_ctx *ctx, void *cuda_stream);

/* Replaces: the `ForceLaw *fl` and `POSVEL_T fsm` (= rmax), `POSVEL_T r` (= rsm) constructor arguments
 * (src/halo_finder/RCBForceTree.h:113-121). */
int haccsr_set_force_law(haccsr_ctx *ctx, int kind, const float *coeffs, int ncoef, float rsm, float rmax);

/* Select the pair-kernel arithmetic (HACCSR_ARITH_*); context-wide, default HACCSR_ARITH_FUSED. */
int haccsr_set_arithmetic(haccsr_ctx *ctx, int mode);

/* Warp-level culling in the pair kernel (fused arithmetic, polynomial law): after the cutoff test, a warp none of whose
 * lanes holds a pair inside the cutoff skips the force law for that source.  The reference evaluates every list pair in
 * full and multiplies 98 % of them by zero (RCBForceTree.cxx:612-613); skipping them changes no bit of
 * the result.  Off by default: with it on, the kernel no longer executes the fixed 30 flop per list pair that the
 * benchmark's roofline accounting assumes (haccsr_stats.pairs_force_law reports what was executed). */
int haccsr_set_culling(haccsr_ctx *ctx, int on);

/* Copy `count` particles from host arrays into the context (H2D).
 * Replaces: handing m_xArr ... m_maskArr to the constructor (src/cpu/Particles.cxx:1317-1327). */
int haccsr_upload(haccsr_ctx *ctx, int64_t count, const float *x, const float *y, const float *z,
                  const float *vx, const float *vy, const float *vz, const float *mass, const float *phi,
                  const int64_t *id, const uint16_t *mask);

/* Copy the context's particles back to host arrays (D2H), in the tree order of the last kick -- the
 * reference likewise leaves all 10 arrays permuted in place (RCBForceTree.cxx:623-672).  Any pointer
 * may be NULL to skip that array. */
int haccsr_download(haccsr_ctx *ctx, int64_t count, float *x, float *y, float *z, float *vx, float *vy,
                    float *vz, float *mass, float *phi, int64_t *id, uint16_t *mask);

/* Optional: page-lock / unlock a caller-owned host array so upload/download run at full PCIe rate. */
int haccsr_host_register(void *ptr, size_t bytes);
int haccsr_host_unregister(void *ptr);

/* The short-range kick on the resident particles: tree build (centre-of-mass recursive bisection,
 * leaves of <= ppn particles), interaction lists (opening angle theta), leaf-vs-list force kernel;
 * v += fcoeff * mass_i * sum_j mass_j f(r2) (x_j - x_i) for particles in leaves touching the force box.
 * Replaces: RCBForceTree<1>::RCBForceTree(...) in full (src/halo_finder/RCBForceTree.cxx:335-450):
 *   tree_lo/tree_hi   = minLoc/maxLoc, force_lo/force_hi = minForceLoc/maxForceLoc, theta = oa, ppn = nd,
 *   fcoeff = fcoeff; only the first `count` resident particles take part (Particles.cxx:1248).
 * tdpts = 1: RCBMonopoleForceTree (-R), an accepted node enters the list as one pseudo-particle at its centroid;
 * tdpts = 12: RCBQuadrupoleForceTree (-S), as 12 pseudo-particles on an icosahedron carrying its monopole, dipole and
 * quadrupole (RCBForceTree.cxx:229-272,519-569; RCBForceTree.h:202-203).  Other values are refused.
 * stats and opts may be NULL. */
int haccsr_kick(haccsr_ctx *ctx, int64_t count, const float tree_lo[3], const float tree_hi[3],
                const float force_lo[3], const float force_hi[3], float theta, int64_t ppn, int tdpts,
                float fcoeff, const haccsr_kick_opts *opts, haccsr_stats *stats);

/* haccsr_upload + haccsr_kick + haccsr_download in one call on caller-owned host arrays -- what the reference's
 * constructor does to the arrays it is handed (src/cpu/Particles.cxx:1313-1338): on return all ten arrays hold the
 * particles in tree order with vx vy vz kicked.  Transfers the kernels do not depend on run on a second stream
 * (the build reads only x y z mass; the force kernel writes only vx vy vz), so with page-locked arrays
 * (haccsr_host_register) most of the PCIe time hides behind the kernels.  phi, id, mask may be NULL. */
int haccsr_kick_host(haccsr_ctx *ctx, int64_t count, float *x, float *y, float *z, float *vx, float *vy, float *vz,
                     float *mass, float *phi, int64_t *id, uint16_t *mask, const float tree_lo[3],
                     const float tree_hi[3], const float force_lo[3], const float force_hi[3], float theta,
                     int64_t ppn, int tdpts, float fcoeff, const haccsr_kick_opts *opts, haccsr_stats *stats);

/* x += prefactor_tau * v for all resident particles.
 * Replaces: Particles::map1 (src/cpu/Particles.cxx:732-758); prefactor_tau = prefactor * tau there. */
int haccsr_stream(haccsr_ctx *ctx, float prefactor_tau);

/* Move particles outside [0,hi)^3 to the tail of the arrays (stable) and report how many remain in front.
 * Replaces: the out-of-box tail move of Particles::resortParticles (src/cpu/Particles.cxx:402-488,
 * m_Np_last :443); the per-cell counting sort is not reproduced because the tree re-sorts anyway. */
int haccsr_partition_in_box(haccsr_ctx *ctx, const float hi[3], int64_t *count_in_box);

/* mass[:] = value for all resident particles.  Replaces: Particles.cxx:1256-1257. */
int haccsr_fill_mass(haccsr_ctx *ctx, float value);

/* One full Particles::subCycle on the resident particles (src/cpu/Particles.cxx:1176-1201): nsub times
 *   [ stream(pt) ; move out-of-box particles to the tail ; mass = 1 ; kick(first count_in_box particles) ; stream(pt) ]
 * with pt = prefactor * (tau2 / nsub) of map1 (:745-755) and fcoeff = c of map2 (:1230-1233, already holding the
 * 1/nsub step fraction).  box_hi = the local grid extent nglt (Domain::ng_local_total) used by resortParticles.
 * Particles stay in HBM throughout: one haccsr_upload before and one haccsr_download after replace the 3*nsub
 * passes over host arrays of the reference.  stats (may be NULL) receives the LAST kick's stats with
 * pairs_evaluated, ms_build, ms_walk, ms_force, ms_total and the launch counts summed over the nsub kicks. */
int haccsr_subcycle(haccsr_ctx *ctx, int nsub, float prefactor_tau, const float box_hi[3], const float tree_lo[3],
                    const float tree_hi[3], const float force_lo[3], const float force_hi[3], float theta,
                    int64_t ppn, int tdpts, float fcoeff, haccsr_stats *stats);

/* ---- the glue of Particles::map2 / map1 / subCycle as code (host arithmetic only; usable without a GPU) -------------------------
 * haccsr_map2_setup computes what Particles::map2 hands to the tree (src/cpu/Particles.cxx:1205-1233), with the reference's own
 * promotions (float members, double literals and TimeStepper values, pi = 4.0*atanf(1.0) stored in a float):
 *   tree box   [0, max(nglt)]^3                                   (:1213-1216)
 *   force box  [edge, nglt - edge] per dimension                    (:1219-1228)
 *   fcoeff     c = gpscal^3 / 4.0 / pi * fscal * tau * step_fraction  (:1230-1233)
 * nglt = Domain::ng_local_total, edge = m_edge, gpscal = m_gpscal = float(ng) / float(np) (:152-153), fscal / tau =
 * TimeStepper::fscal() / tau() (double), step_fraction = 1.0 / nsub (:1180). */
typedef struct haccsr_map2 {
  float tree_lo[3], tree_hi[3], force_lo[3], force_hi[3];
  float fcoeff;
} haccsr_map2;
int haccsr_map2_setup(const int32_t nglt[3], float edge, float gpscal, double fscal, double tau, double step_fraction,
                      haccsr_map2 *out);
/* The factor Particles::map1 multiplies the velocities with, prefactor * tau in float (src/cpu/Particles.cxx:745-754):
 * pf = powf(pp, 1.0 + 1.0 / alpha), prefactor = 1.0 / (alpha * adot * pf); map1's arguments pp, tau, adot are floats. */
float haccsr_map1_factor(float pp, float tau, float adot, float alpha);
/* Particles::subCycle(gts) on the resident particles from the TimeStepper's scalars alone (src/cpu/Particles.cxx:1176-1201):
 * step_fraction = 1.0 / nsub; nsub x [ map1(pp, step_fraction * tau2, adot) ; map2(gts, step_fraction) ; map1(...) ] with the
 * boxes and c of haccsr_map2_setup.  The force law, rmax and rsm are those of haccsr_set_force_law (m_fl, m_fsrrmax, m_rsm). */
int haccsr_particles_subcycle(haccsr_ctx *ctx, int nsub, const int32_t nglt[3], float edge, float gpscal, float alpha,
                              double pp, double adot, double tau, double tau2, double fscal, float theta, int64_t ppn,
                              int tdpts, haccsr_stats *stats);

/* ---- PM coupling: the particle side of the long-range step -------------------------------------
 * The FFT Poisson solver stays the reference's; these two calls replace the particle loops either side of it so the
 * particles need not leave the GPU between sub-cycles.  Grids are ng[0]*ng[1]*ng[2] floats, index (ix*ng[1]+iy)*ng[2]+iz
 * (Particles::array_index, src/cpu/Particles.cxx:373-395), GRID_T = float (-DGRID_32); grid_on_device != 0 means the
 * pointer is a device pointer.
 *
 * haccsr_cic: cloud-in-cell deposit of all resident particles, rho (overwritten) = sum over particles of c * wx wy wz on
 * the 8 cells around each; cells outside the grid are dropped (the reference's `safe` slot).
 * Replaces: Particles::cic (src/cpu/Particles.cxx:589-643; c = m_gpscal^3).  Order-independent (fixed-point atomics), equal
 * to the reference's sequential float sums to FP32 rounding. */
int haccsr_cic(haccsr_ctx *ctx, const int32_t ng[3], float c, float *rho, int grid_on_device);
/* haccsr_inverse_cic: v[comp] += (CIC interpolation of grid at the particle) * fscal * tau for all resident particles;
 * comp 0, 1, 2 = vx, vy, vz, 3 = phi.  Replaces: Particles::inverse_cic(tau, fscal, comp) (src/cpu/Particles.cxx:647-714),
 * called once per gradient component by map2_poisson_backward_gradient (src/simulation/mc3.cxx:336-379).  Bit-identical
 * to the reference's loop (same weights, same promotions, same order of the 8 terms). */
int haccsr_inverse_cic(haccsr_ctx *ctx, const int32_t ng[3], const float *grid, int grid_on_device, float tau, float fscal,
                       int comp);

/* ---- overload (ghost-zone) refresh: the per-rank, on-device part ---------------------------------------------------
 * Replaces ParticleExchange::exchangeParticles as driven by MC3Extras::refreshParticles at refresh steps
 * (src/simulation/MC3Extras.cxx:660-706; src/halo_finder/ParticleExchange.cxx:488-762), in the local grid units
 * of the tree.  Directions are numbered d = (sx+1)*9 + (sy+1)*3 + (sz+1), s in {-1,0,1}^3, d != 13; the caller
 * maps each direction to a message SLOT (0..25) so that the send buffer is ordered by destination rank.
 *
 *   haccsr_refresh_begin   drops the ghosts (keeps x in [alive_lo, alive_hi), Particles.cxx:975), finds the alive
 *                          particles that are ghosts of each neighbour (inclusive slabs of width ol,
 *                          ParticleExchange.cxx:280-450,549-565) and returns the particle count of every message;
 *   haccsr_refresh_pack    writes the 26 messages into a device buffer at the byte offsets the caller chose, each
 *                          haccsr_refresh_message_bytes(n) long: id[n] | x y z vx vy vz mass phi [n] | mask[n];
 *                          positions are already in the receiver's frame (x - s * (alive_hi - alive_lo), the
 *                          reference's +-boxSize wrap of :672-673 included);
 *   (transport between ranks: one all-to-all-v of the buffer over NCCL -- hacc_coral_b200/refresh.py -- or MPI)
 *   haccsr_refresh_append  appends one received message (device pointer) behind the resident particles.
 * Messages keep the sender's particle order, so the result is deterministic. */
int64_t haccsr_refresh_message_bytes(int64_t n);
int haccsr_refresh_begin(haccsr_ctx *ctx, const float alive_lo[3], const float alive_hi[3], float ol,
                         const int32_t slot_of_dir[27], int64_t counts_by_slot[27], int64_t *n_alive);
int haccsr_refresh_pack(haccsr_ctx *ctx, const int64_t byte_off_by_slot[27], void *sendbuf_device);
int haccsr_refresh_append(haccsr_ctx *ctx, const void *message_device, int64_t n);
/* Number of particles currently resident in the context. */
int64_t haccsr_resident(haccsr_ctx *ctx);

/* ---- overload refresh as one call, transport included (what a C++ HACC rank calls at a refresh step) ------------------------
 * haccsr_refresh = haccsr_refresh_begin + plan + haccsr_refresh_pack + one grouped ncclSend/ncclRecv over NVLink + append.
 * Replaces: MC3Extras::refreshParticles -> ParticleExchange::exchangeParticles (src/simulation/MC3Extras.cxx:660-706,
 * src/halo_finder/ParticleExchange.cxx:488-762) for the particles resident in the context.  The ranks form the periodic
 * Cartesian grid dims[0] x dims[1] x dims[2] of Partition (src/halo_finder/Partition.cxx:121-137): rank = (px*dims[1] + py)*
 * dims[2] + pz, and nccl_comm is an ncclComm_t over exactly those ranks in that order (NULL is allowed for 1 x 1 x 1, where
 * every neighbour is the rank itself).  Collective: every rank of the communicator must call it.  On return the context
 * holds its alive particles (stable order) followed by the received ghosts in (source rank, direction) order --
 * deterministic, and bit-identical to extracting each rank's overloaded sub-volume from the global particle set.
 * NCCL is loaded at run time (libnccl.so.2); status 3 if it cannot be.  If the ghosts of ANY rank would not fit in that rank's
 * context (alive + incoming > capacity), every rank returns 1 with the same message before anything is exchanged (alive counts
 * and capacities travel with the message sizes), and every context is left holding its alive particles only. */
typedef struct haccsr_refresh_stats {
  int64_t alive, ghosts, sent;             /* particles kept, received, sent (a particle goes to up to 7 neighbours)       */
  int64_t bytes_sent, bytes_received;      /* packed message bytes, self messages included                                 */
  int64_t bytes_sent_remote;               /* bytes that crossed NVLink (bytes_sent minus the messages to the rank itself) */
  int32_t messages_received, reserved;
  float ms_total;                          /* device time of the whole call (events on the context's stream)               */
} haccsr_refresh_stats;
int haccsr_refresh(haccsr_ctx *ctx, void *nccl_comm, const int32_t dims[3], int32_t rank, const float alive_lo[3],
                   const float alive_hi[3], float ol, haccsr_refresh_stats *stats);
/* The message plan haccsr_refresh uses for `rank` (host arithmetic only, usable without a GPU): the 26 directions sorted by
 * (destination rank, direction), so that the messages for one destination are contiguous in the send buffer and both sides
 * derive the same layout from the counts alone.  dir_of_slot[s] = direction d = (sx+1)*9 + (sy+1)*3 + (sz+1) of message slot s,
 * dest_of_slot[s] = the rank it goes to (periodic Cartesian neighbour, Partition.cxx:140-260). */
int haccsr_refresh_plan(const int32_t dims[3], int32_t rank, int32_t dir_of_slot[26], int32_t dest_of_slot[26]);
/* Communicator helpers for hosts that do not link NCCL themselves: rank 0 obtains an id, distributes its 128 bytes by its own
 * means (MPI_Bcast in HACC; torch.distributed in bench.py), every rank creates its communicator on its device. */
#define HACCSR_NCCL_ID_BYTES 128
int haccsr_nccl_unique_id(void *id128);
int haccsr_nccl_comm_create(void **comm, int device, int nranks, int rank, const void *id128);
int haccsr_nccl_comm_destroy(void *comm);

/* ---- inspection (tests and tools): the tree and the lists of the last kick ---------------------- */
/* Node table, `cap` entries per array; box10 = xmin[3] xmax[3] xc[3] ppm per node
 * (TreeNode, src/halo_finder/RCBForceTree.h:131-147).  Returns the node count in *nodes. */
int haccsr_get_tree(haccsr_ctx *ctx, int64_t cap, int64_t *nodes, int32_t *count, int32_t *offset,
                    int32_t *cl, int32_t *cr, float *box10);
/* The 12 pseudo-particles (x, y, z, mass) of every node of the last kick made with tdpts = 12: 48 floats per node
 * (pppts<12> / pp<12>, RCBForceTree.cxx:525-569). */
int haccsr_get_pseudo_particles(haccsr_ctx *ctx, int64_t cap_nodes, float *pp48);
/* Interaction lists of the last kick.  For node k: ranges [range_off[k], range_off[k+1]) of
 * (start,count) pairs.  start < 2^31 indexes particles in tree order; start >= 2^31 indexes the
 * pseudo-particle pool (start - 2^31).  Query sizes first with cap_* = 0. */
int haccsr_get_lists(haccsr_ctx *ctx, int64_t cap_nodes, int64_t cap_ranges, int64_t cap_pool,
                     int64_t *n_nodes, int64_t *n_ranges, int64_t *n_pool, uint32_t *range_off,
                     uint32_t *ranges /* 2 per entry */, float *pool /* 4 per entry */);

#ifdef __cplusplus
}
#endif
#endif /* HACCSR_H */
