// force_kernels.cuh -- the pair kernels of libhaccsr and their launcher as templates (internal).  Included by force.cu (item set-up,
// dispatch; no instantiation there) and by force_v0.cu ... force_v4.cu, which instantiate one arithmetic variant each so that the
// five variants compile in parallel (one translation unit took 2 min 40 s).
#pragma once
// force.cu -- the leaf-vs-list short-range force kernel and kick, sm_100a.
//
// Computes, for every particle i of every sink leaf, the sum over the leaf's interaction list
//     a_i = sum_j m_j f(r2_ij) (x_j - x_i),   f(r2) = (r2 + rsm^2)^-3/2 - poly(r2)   for r2 < rmax^2, else 0
// and kicks v_i += fcoeff * m_i * a_i.  This is nbody1 (reference src/halo_finder/RCBForceTree.cxx:575-620)
// in the accumulate-then-scale form of the BG/Q kernel (src/halo_finder/BGQStep16.c:170-187 with
// RCBForceTree.cxx:594-596); the force law is ForceLawSR over FGridEvalPoly (ForceLaw.cxx:137-141,187-192).
//
// Mapping to the B200 (FP32-FMA-pipe bound; no tensor cores -- this is not a contraction):
//  * one WARP per work item = (sink leaf, chunk of <= 32*SMAX_ sinks); one warp per CTA, so the producer /
//    consumer hand-off needs no block barrier, only mbarriers and __syncwarp.
//  * each thread keeps S sinks (position + accumulator) in registers; S = ceil(chunk/32) is picked per item,
//    so a 300-particle leaf runs at 94 % lane utilisation instead of 59 % with a fixed 512-slot block.
//  * sources stream through a 4-stage shared-memory ring of 128-source float4 tiles.  Every list range is
//    a contiguous, 16-B aligned piece of the tree-ordered float4 array (or of the pseudo-particle pool),
//    so lane 0 moves it with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes);
//    the warp then reads each source once with a broadcast LDS.128 and applies it to its S sinks.
//  * r2 is formed with separate multiplies and adds in the reference's order (no FMA contraction), so the
//    set of pairs inside the cutoff is bit-identical to the CPU's; everything after r2 uses FMAs.
//  * (r2+rsm^2)^-3/2 = rsqrt((r2+rsm^2)^3) with MUFU.RSQ (rsqrt.approx.ftz), Horner polynomial in FMAs,
//    cutoff as one compare + select on the per-pair scalar.
// Reduction order (documented for the parity gate): each sink accumulates sequentially over its list in
// list order, in FP32, in one thread -- no cross-lane reduction is needed because sinks, not sources, are
// spread over lanes.
#include "common.cuh"

namespace haccsr {


static constexpr int SMAX_ = 8;        // max sink groups per thread (4 packed pairs): 8 x 32 = 256 sinks per work item
static constexpr int FTILE = 128;      // sources per shared-memory tile
static constexpr int FSTAGES = 4;      // ring depth

struct ForceParams {
  const WorkItem *items;
  const unsigned *range_off;   // per node
  const uint2 *ranges;
  const unsigned *list_len;    // per node
  const float4 *src4;
  const float4 *pool;
  float *vx, *vy, *vz;
  unsigned long long *incut;   // optional counter
  float a[8];                  // SR_POLY: a[0..6]; SR_FIT: b c d e f g h l of the analytic grid-force fit
  float b[8];                  // SR_POLY, fused arithmetic: MINUS the polynomial re-expanded in s = r2 + rsm^2
  float rsm2, rmax2, smax, fcoeff;   // smax = rmax2 + rsm2: the cutoff on s
  int unit_mass;               // 1: every particle mass is exactly 1.0f
  int defer;                   // 1: vx vy vz point at acceleration arrays; the kernel stores a_i there and apply_kick() kicks later
  const float *tab_f, *tab_r2; // SR_INTERP: grid force and its abscissae r2_i (device arrays of ntab floats)
  float tab_r2min, tab_r2max, tab_oodr2;
  int ntab;
};

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
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- the pair arithmetic ---------------------------------------------------------------------------------
// Packed FP32x2 (FFMA2 / FADD2 / FMUL2, new on sm_100): one instruction works on the pairs (source j, sink a)
// and (source j, sink b); the source coordinate is the instruction's broadcast scalar operand, the two sinks
// sit in an aligned register pair.  This halves the issue slots of the FMA-pipe work, which moves the bound
// from the issue port (25.4 instructions per pair in the scalar form, profiles/r1_force_ncu_summary.md) to the
// FMA pipe itself (21 lane-operations per pair).
// Rounding: every operation is round-to-nearest on each half, the same sequence as the scalar form, so a sink
// gets bit-identical results whether it is processed in a packed pair or in the scalar remainder group.
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it never does for the scalar .rn forms), so
// the three squares of r2 are formed with scalar mul.rn and only summed packed: r2 keeps the reference's
// unfused value (dx*dx + dy*dy) + dz*dz and the set of pairs inside the cutoff stays bit-identical to the CPU's.
struct SinkRegs2 { float2 nx, ny, nz, ax, ay, az; };   // two sinks: negated position, accumulators
struct SinkRegs1 { float nx, ny, nz, ax, ay, az; };

// UNITM: every source of the item has mass exactly 1.0f (HACC resets mass to 1 before each kick,
// Particles.cxx:1256-1257, and the item's list holds no pseudo-particle), so the multiply by m_j is skipped --
// x * 1.0f == x bit for bit, the result is unchanged.
// FUSED (HACCSR_ARITH_FUSED, SR_POLY only): the contracted arithmetic of the reference's production kernel --
// the QPX loop forms r2 with three multiply-adds (BGQStep16.c:76-86) -- taken one step further:
//   s  = fma(dz,dz, fma(dy,dy, fma(dx,dx, rsm^2)))        the chain is seeded with rsm^2, so it yields r2 + rsm^2;
//   -g = Horner in s with P.b = MINUS the law's polynomial re-expanded about -rsm^2 (in double, on the host);
//   f  = fma(rs*rs, rs, -g), rs = rsqrt(s)                 s^-3/2 and the subtraction in one multiply and one FMA;
//   cutoff s < P.smax (= rmax^2 + rsm^2) as the predicate of three scalar FFMAs per sink -- no select.
// 16 FMA-pipe operations and 2 other instructions (MUFU, FSETP) per pair instead of 20 + 3.  On sm_100 an FFMA2
// occupies the issue port for two cycles, so every instruction saved shows (tools/microbench_force.cu, modes 3/5/6/7:
// 58.7 / 70.5 / 72.7 / 75.4 % of the FP32 peak).  Oracle form FORM_FUSED restates exactly this sequence.
// RS3 (HACCSR_ARITH_FUSED_RS3): (r2 + rsm^2)^-3/2 as rsqrt(s*s*s) instead of rsqrt(s)^3 -- one more FMA-pipe operation a pair,
// a third of the rsqrt error (MUFU.RSQ's error enters once, halved by the -3/2... see DESIGN.md 3.1 for the measured distances)
template <int NC, bool GUARD0, bool COUNT, bool UNITM, bool RS3 = false>
__device__ __forceinline__ void interact2_fused(const float4 s, SinkRegs2 &k, const ForceParams &P, unsigned &cnt_a, unsigned &cnt_b) {
  const float2 dx = __fadd2_rn(make_float2(s.x, s.x), k.nx), dy = __fadd2_rn(make_float2(s.y, s.y), k.ny),
               dz = __fadd2_rn(make_float2(s.z, s.z), k.nz);
  float2 t = __ffma2_rn(dx, dx, make_float2(P.rsm2, P.rsm2));
  t = __ffma2_rn(dy, dy, t);
  t = __ffma2_rn(dz, dz, t);
  float2 p = make_float2(P.b[NC - 1], P.b[NC - 1]);
#pragma unroll
  for (int q = NC - 2; q >= 0; --q) p = __ffma2_rn(p, t, make_float2(P.b[q], P.b[q]));
  float2 f;
  if (RS3) {
    const float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
    f = __fadd2_rn(make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y)), p);
  } else {
    const float2 rs = make_float2(rsqrt_ftz(t.x), rsqrt_ftz(t.y));
    f = __ffma2_rn(__fmul2_rn(rs, rs), rs, p);
  }
  if (!UNITM) f = __fmul2_rn(f, make_float2(s.w, s.w));
  bool in_a = t.x < P.smax, in_b = t.y < P.smax;
  if (GUARD0) { in_a = in_a && (t.x > P.rsm2); in_b = in_b && (t.y > P.rsm2); }
  if (in_a) { k.ax.x = __fmaf_rn(f.x, dx.x, k.ax.x); k.ay.x = __fmaf_rn(f.x, dy.x, k.ay.x); k.az.x = __fmaf_rn(f.x, dz.x, k.az.x); }
  if (in_b) { k.ax.y = __fmaf_rn(f.y, dx.y, k.ax.y); k.ay.y = __fmaf_rn(f.y, dy.y, k.ay.y); k.az.y = __fmaf_rn(f.y, dz.y, k.az.y); }
  if (COUNT) { cnt_a += (in_a && t.x > P.rsm2) ? 1u : 0u; cnt_b += (in_b && t.y > P.rsm2) ? 1u : 0u; }
}
template <int NC, bool GUARD0, bool COUNT, bool UNITM, bool RS3 = false>
__device__ __forceinline__ void interact1_fused(const float4 s, SinkRegs1 &k, const ForceParams &P, unsigned &cnt) {
  const float dx = __fadd_rn(s.x, k.nx), dy = __fadd_rn(s.y, k.ny), dz = __fadd_rn(s.z, k.nz);
  const float t = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, P.rsm2)));
  float p = P.b[NC - 1];
#pragma unroll
  for (int q = NC - 2; q >= 0; --q) p = __fmaf_rn(p, t, P.b[q]);
  float f;
  if (RS3) {
    f = __fadd_rn(rsqrt_ftz(__fmul_rn(__fmul_rn(t, t), t)), p);
  } else {
    const float rs = rsqrt_ftz(t);
    f = __fmaf_rn(__fmul_rn(rs, rs), rs, p);
  }
  if (!UNITM) f = __fmul_rn(f, s.w);
  bool in = t < P.smax;
  if (GUARD0) in = in && (t > P.rsm2);
  if (in) { k.ax = __fmaf_rn(f, dx, k.ax); k.ay = __fmaf_rn(f, dy, k.ay); k.az = __fmaf_rn(f, dz, k.az); }
  if (COUNT) cnt += (in && t > P.rsm2) ? 1u : 0u;
}

// CULL (haccsr_set_culling, fused arithmetic only): the same sequence with a warp-level early exit.  Of the pairs the
// reference's lists make the kernel look at, 98 % lie outside the cutoff and contribute nothing; with
// sinks spread over lanes, the 64 sinks a packed instruction serves are neighbours in tree order, and for ~86 % of the
// sources none of them is inside the cutoff.  One vote after the cutoff test then skips the polynomial, the rsqrt and
// the accumulate for the whole warp.  Results are bit-identical to the unculled kernel: a skipped pair is exactly a
// pair whose accumulate predicate was false.  `nfull` counts the lane-pairs that ran the force law (COUNT variant).
template <int NC, bool GUARD0, bool COUNT, bool UNITM, bool RS3 = false>
__device__ __forceinline__ void interact2_cull(const float4 s, SinkRegs2 &k, const ForceParams &P, unsigned &cnt_a, unsigned &cnt_b,
                                               unsigned &nf_a, unsigned &nf_b) {
  const float2 dx = __fadd2_rn(make_float2(s.x, s.x), k.nx), dy = __fadd2_rn(make_float2(s.y, s.y), k.ny),
               dz = __fadd2_rn(make_float2(s.z, s.z), k.nz);
  float2 t = __ffma2_rn(dx, dx, make_float2(P.rsm2, P.rsm2));
  t = __ffma2_rn(dy, dy, t);
  t = __ffma2_rn(dz, dz, t);
  bool in_a = t.x < P.smax, in_b = t.y < P.smax;
  if (GUARD0) { in_a = in_a && (t.x > P.rsm2); in_b = in_b && (t.y > P.rsm2); }
  if (!__any_sync(0xffffffffu, in_a || in_b)) return;
  float2 p = make_float2(P.b[NC - 1], P.b[NC - 1]);
#pragma unroll
  for (int q = NC - 2; q >= 0; --q) p = __ffma2_rn(p, t, make_float2(P.b[q], P.b[q]));
  float2 f;
  if (RS3) {
    const float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
    f = __fadd2_rn(make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y)), p);
  } else {
    const float2 rs = make_float2(rsqrt_ftz(t.x), rsqrt_ftz(t.y));
    f = __ffma2_rn(__fmul2_rn(rs, rs), rs, p);
  }
  if (!UNITM) f = __fmul2_rn(f, make_float2(s.w, s.w));
  if (in_a) { k.ax.x = __fmaf_rn(f.x, dx.x, k.ax.x); k.ay.x = __fmaf_rn(f.x, dy.x, k.ay.x); k.az.x = __fmaf_rn(f.x, dz.x, k.az.x); }
  if (in_b) { k.ax.y = __fmaf_rn(f.y, dx.y, k.ax.y); k.ay.y = __fmaf_rn(f.y, dy.y, k.ay.y); k.az.y = __fmaf_rn(f.y, dz.y, k.az.y); }
  if (COUNT) { cnt_a += (in_a && t.x > P.rsm2) ? 1u : 0u; cnt_b += (in_b && t.y > P.rsm2) ? 1u : 0u; nf_a++; nf_b++; }
}
template <int NC, bool GUARD0, bool COUNT, bool UNITM, bool RS3 = false>
__device__ __forceinline__ void interact1_cull(const float4 s, SinkRegs1 &k, const ForceParams &P, unsigned &cnt, unsigned &nfull) {
  const float dx = __fadd_rn(s.x, k.nx), dy = __fadd_rn(s.y, k.ny), dz = __fadd_rn(s.z, k.nz);
  const float t = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, P.rsm2)));
  bool in = t < P.smax;
  if (GUARD0) in = in && (t > P.rsm2);
  if (!__any_sync(0xffffffffu, in)) return;
  float p = P.b[NC - 1];
#pragma unroll
  for (int q = NC - 2; q >= 0; --q) p = __fmaf_rn(p, t, P.b[q]);
  float f;
  if (RS3) {
    f = __fadd_rn(rsqrt_ftz(__fmul_rn(__fmul_rn(t, t), t)), p);
  } else {
    const float rs = rsqrt_ftz(t);
    f = __fmaf_rn(__fmul_rn(rs, rs), rs, p);
  }
  if (!UNITM) f = __fmul_rn(f, s.w);
  if (in) { k.ax = __fmaf_rn(f, dx, k.ax); k.ay = __fmaf_rn(f, dy, k.ay); k.az = __fmaf_rn(f, dz, k.az); }
  if (COUNT) { cnt += (in && t > P.rsm2) ? 1u : 0u; nfull += 1u; }
}

template <int NC, int LAW, bool GUARD0, bool COUNT, bool UNITM>
__device__ __forceinline__ void interact2(const float4 s, SinkRegs2 &k, const ForceParams &P, unsigned &cnt_a, unsigned &cnt_b) {
  const float2 dx = __fadd2_rn(make_float2(s.x, s.x), k.nx), dy = __fadd2_rn(make_float2(s.y, s.y), k.ny),
               dz = __fadd2_rn(make_float2(s.z, s.z), k.nz);
  // reference order, no contraction: (dx*dx + dy*dy) + dz*dz   (RCBForceTree.cxx:608, BGQStep16.c:176)
  const float2 qx = make_float2(__fmul_rn(dx.x, dx.x), __fmul_rn(dx.y, dx.y));
  const float2 qy = make_float2(__fmul_rn(dy.x, dy.x), __fmul_rn(dy.y, dy.y));
  const float2 qz = make_float2(__fmul_rn(dz.x, dz.x), __fmul_rn(dz.y, dz.y));
  const float2 r2 = __fadd2_rn(__fadd2_rn(qx, qy), qz);
  const float2 t = __fadd2_rn(r2, make_float2(P.rsm2, P.rsm2));
  const float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
  float2 f = make_float2(rsqrt_ftz(t3.x), rsqrt_ftz(t3.y));
  if (LAW == 0) {
    float2 p = make_float2(P.a[NC - 1], P.a[NC - 1]);
#pragma unroll
    for (int q = NC - 2; q >= 0; --q) p = __ffma2_rn(p, r2, make_float2(P.a[q], P.a[q]));
    f = __fadd2_rn(f, make_float2(-p.x, -p.y));
  }
  if (!UNITM) f = __fmul2_rn(f, make_float2(s.w, s.w));
  bool in_a = r2.x < P.rmax2, in_b = r2.y < P.rmax2;
  if (GUARD0) { in_a = in_a && (r2.x > 0.0f); in_b = in_b && (r2.y > 0.0f); }
  f.x = in_a ? f.x : 0.0f;
  f.y = in_b ? f.y : 0.0f;
  if (COUNT) { cnt_a += (in_a && r2.x > 0.0f) ? 1u : 0u; cnt_b += (in_b && r2.y > 0.0f) ? 1u : 0u; }
  k.ax = __ffma2_rn(f, dx, k.ax); k.ay = __ffma2_rn(f, dy, k.ay); k.az = __ffma2_rn(f, dz, k.az);
}

// Grid force g(r2) of the laws that only exist in scalar form (the per-pair cost is dominated by libm-grade
// transcendentals or a table gather, so packing buys nothing).
// LAW 2 = FGridEvalFit (reference ForceLaw.cxx:39-51,70-80), the reference's expression evaluated in float (the reference promotes
// `1.0 + d r2`, `-1.0 d r2` and `2.0/3.0 b^3` to double and nvcc may contract here: agreement to FP32 rounding, gated against the
// compiled reference's golden vectors in tests/test_gpu_parity.py::test_fit_and_interp_laws_match_reference; rmax beyond the
// fit's own range is refused by haccsr_set_force_law):
//   g = [tanh(br) - br/cosh^2(br) + c r^3 (1 + d r^2) exp(-d r^2) + e r^2 (f r^2 + g r^4 + l r^6) exp(-h r^2)] / r^3
// LAW 3 = FGridEvalInterp (ForceLaw.cxx:145-172): linear interpolation in r2, zero outside (r2min, r2max).
template <int LAW>
__device__ __forceinline__ float grid_force_general(float r2, const ForceParams &P) {
  if (LAW == 2) {
    const float b = P.a[0], c = P.a[1], d = P.a[2], e = P.a[3], ff = P.a[4], g = P.a[5], h = P.a[6], l = P.a[7];
    const float r = sqrtf(r2);
    if (!(r > 0.0f)) return c + (2.0f / 3.0f) * b * b * b;
    const float r4 = r2 * r2, r6 = r4 * r2;
    const float br = b * r;
    const float ch = coshf(br);
    const float num = tanhf(br) - br / ch / ch + c * r * r2 * (1.0f + d * r2) * expf(-d * r2) +
                      e * r2 * (ff * r2 + g * r4 + l * r6) * expf(-h * r2);
    return num / (r * r * r);
  } else {
    const bool in = (r2 > P.tab_r2min) && (r2 < P.tab_r2max);
    const int i = in ? (int)((r2 - P.tab_r2min) * P.tab_oodr2) : 0;
    const float f0 = __ldg(P.tab_f + i), f1 = __ldg(P.tab_f + i + 1), x0 = __ldg(P.tab_r2 + i);
    const float v = f0 + (r2 - x0) * P.tab_oodr2 * (f1 - f0);
    return in ? v : 0.0f;
  }
}

template <int NC, int LAW, bool GUARD0, bool COUNT, bool UNITM>
__device__ __forceinline__ void interact1(const float4 s, SinkRegs1 &k, const ForceParams &P, unsigned &cnt) {
  const float dx = __fadd_rn(s.x, k.nx), dy = __fadd_rn(s.y, k.ny), dz = __fadd_rn(s.z, k.nz);
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float t = __fadd_rn(r2, P.rsm2);
  float f = rsqrt_ftz(__fmul_rn(__fmul_rn(t, t), t));
  if (LAW == 0) {
    float p = P.a[NC - 1];
#pragma unroll
    for (int q = NC - 2; q >= 0; --q) p = __fmaf_rn(p, r2, P.a[q]);
    f = __fadd_rn(f, -p);
  } else if (LAW >= 2) {
    f = __fadd_rn(f, -grid_force_general<LAW>(r2, P));
  }
  if (!UNITM) f = __fmul_rn(f, s.w);
  bool in = r2 < P.rmax2;
  if (GUARD0) in = in && (r2 > 0.0f);
  f = in ? f : 0.0f;
  if (COUNT) cnt += (in && r2 > 0.0f) ? 1u : 0u;
  k.ax = __fmaf_rn(f, dx, k.ax); k.ay = __fmaf_rn(f, dy, k.ay); k.az = __fmaf_rn(f, dz, k.az);
}

struct Producer {
  const uint2 *ranges;
  unsigned ri, rend;      // current / end range index
  unsigned roff;          // sources already taken from the current range
  unsigned remaining;     // sources still to be fetched
};

// lane 0 only: post the copies of the next tile into `stage`.  Deliberately not inlined: one copy of this
// cold code instead of five per (S2, ODD) variant keeps the kernel's instruction footprint small.
static __device__ __noinline__ void produce_tile(Producer &pr, const ForceParams &P, float4 *tile, unsigned bar) {
  unsigned n = pr.remaining < (unsigned)FTILE ? pr.remaining : (unsigned)FTILE;
  if (n == 0) return;
  mbar_expect_tx(bar, n * 16u);
  unsigned pos = 0;
  while (pos < n) {
    uint2 r = __ldg(pr.ranges + pr.ri);
    unsigned take = r.y - pr.roff;
    if (take > n - pos) take = n - pos;
    const float4 *base = (r.x & POOL_FLAG) ? (P.pool + (r.x & ~POOL_FLAG)) : (P.src4 + r.x);
    bulk_g2s(smem_u32(tile + pos), base + pr.roff, take * 16u, bar);
    pos += take; pr.roff += take;
    if (pr.roff == r.y) { pr.ri++; pr.roff = 0; }
  }
  pr.remaining -= n;
}

// One work item: S = 2*S2 + ODD groups of 32 sinks (group g = sinks g*32 + lane of the chunk); groups 2k, 2k+1
// form packed pair k, an odd last group runs the scalar form.  Loop bodies are kept small on purpose (at most
// ~130 instructions): the eight (S2, ODD) variants together fit the 32 KB instruction cache, which the first
// version's 4x-unrolled bodies (72 KB) did not -- that showed as "no_instruction" stalls, worst on clustered
// snapshots where all eight variants are in flight on one SM.
template <int S2, int S1, int NC, int LAW, bool GUARD0, bool COUNT, bool UNITM, int FUSED>
__device__ __forceinline__ void run_item(const WorkItem it, const ForceParams &P, float4 (*tiles)[FTILE],
                                         unsigned long long *bars) {
  constexpr int S = 2 * S2 + S1;      // S1 scalar groups follow the S2 packed pairs
  constexpr int UNR = (LAW >= 2) ? 1 : ((S <= 2) ? 4 : ((S <= 4) ? 2 : 1));
  const int lane = threadIdx.x;
  SinkRegs2 k2[S2 > 0 ? S2 : 1];
  SinkRegs1 k1[S1 > 0 ? S1 : 1];
  const int node_off_sink = it.sink_begin;
#pragma unroll
  for (int g = 0; g < S; ++g) {
    int j = g * 32 + lane;
    // lanes past the end of the chunk re-use the chunk's first sink (finite numbers, result discarded)
    float4 s = __ldg(P.src4 + node_off_sink + (j < it.sink_count ? j : 0));
    if (g < 2 * S2) {
      SinkRegs2 &k = k2[g >> 1];
      if ((g & 1) == 0) { k.nx.x = -s.x; k.ny.x = -s.y; k.nz.x = -s.z; }
      else { k.nx.y = -s.x; k.ny.y = -s.y; k.nz.y = -s.z; }
    } else {
      SinkRegs1 &k = k1[g - 2 * S2];
      k.nx = -s.x; k.ny = -s.y; k.nz = -s.z;
    }
  }
#pragma unroll
  for (int k = 0; k < S2; ++k) k2[k].ax = k2[k].ay = k2[k].az = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < (S1 > 0 ? S1 : 1); ++k) k1[k].ax = k1[k].ay = k1[k].az = 0.f;
  Producer pr;
  pr.ranges = P.ranges; pr.ri = P.range_off[it.node]; pr.rend = P.range_off[it.node + 1];
  pr.roff = 0; pr.remaining = P.list_len[it.node];
  const unsigned total = pr.remaining;
  const unsigned ntiles = (total + FTILE - 1) / FTILE;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < FSTAGES; ++s) produce_tile(pr, P, tiles[s], smem_u32(&bars[s]));
  }
  unsigned cnt[S];
  unsigned nf[S];           // FUSED == 2, COUNT: pairs of this thread's sinks that ran the force law
#pragma unroll
  for (int g = 0; g < S; ++g) { cnt[g] = 0; nf[g] = 0; }
  for (unsigned t = 0; t < ntiles; ++t) {
    const int stage = t % FSTAGES;
    const unsigned parity = (t / FSTAGES) & 1u;
    mbar_wait(smem_u32(&bars[stage]), parity);
    const unsigned nsrc = (t + 1 == ntiles) ? (total - t * FTILE) : (unsigned)FTILE;
    const float4 *tile = tiles[stage];
#pragma unroll UNR
    for (unsigned j = 0; j < nsrc; ++j) {
      const float4 s = tile[j];
#pragma unroll
      for (int k = 0; k < S2; ++k) {
        if (FUSED == 2 || FUSED == 4) interact2_cull<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1], nf[2 * k], nf[2 * k + 1]);
        else if (FUSED) interact2_fused<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1]);
        else interact2<NC, LAW, GUARD0, COUNT, UNITM>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1]);
      }
#pragma unroll
      for (int k = 0; k < S1; ++k) {
        if (FUSED == 2 || FUSED == 4) interact1_cull<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k1[k], P, cnt[2 * S2 + k], nf[2 * S2 + k]);
        else if (FUSED) interact1_fused<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k1[k], P, cnt[2 * S2 + k]);
        else interact1<NC, LAW, GUARD0, COUNT, UNITM>(s, k1[k], P, cnt[2 * S2 + k]);
      }
    }
    __syncwarp();                       // every lane is done reading this stage
    if (lane == 0) produce_tile(pr, P, tiles[stage], smem_u32(&bars[stage]));
  }
  // kick: v += fcoeff * m_i * a   (RCBForceTree.cxx:594-596 / :615-617)
#pragma unroll
  for (int g = 0; g < S; ++g) {
    int j = g * 32 + lane;
    if (j < it.sink_count) {
      float ax, ay, az;
      if (g < 2 * S2) {
        const SinkRegs2 &k = k2[g >> 1];
        ax = (g & 1) ? k.ax.y : k.ax.x; ay = (g & 1) ? k.ay.y : k.ay.x; az = (g & 1) ? k.az.y : k.az.x;
      } else { const SinkRegs1 &k = k1[g - 2 * S2]; ax = k.ax; ay = k.ay; az = k.az; }
      int gi = node_off_sink + j;
      if (P.defer) { P.vx[gi] = ax; P.vy[gi] = ay; P.vz[gi] = az; continue; }     // haccsr_kick_host: the velocities are still on their way
      float c = P.fcoeff * __ldg(&P.src4[gi].w);
      P.vx[gi] = fmaf(c, ax, P.vx[gi]); P.vy[gi] = fmaf(c, ay, P.vy[gi]); P.vz[gi] = fmaf(c, az, P.vz[gi]);
    }
  }
  if (COUNT) {
    unsigned long long c64 = 0;   // padded lanes duplicate the chunk's first sink: not counted
#pragma unroll
    for (int g = 0; g < S; ++g) c64 += (g * 32 + lane < it.sink_count) ? cnt[g] : 0u;
    for (int o = 16; o > 0; o >>= 1) c64 += __shfl_down_sync(0xffffffffu, c64, o);
    if (lane == 0) atomicAdd(P.incut, c64);
    if (FUSED == 2 || FUSED == 4) {
      unsigned long long f64 = 0;
#pragma unroll
      for (int g = 0; g < S; ++g) f64 += (g * 32 + lane < it.sink_count) ? nf[g] : 0u;
      for (int o = 16; o > 0; o >>= 1) f64 += __shfl_down_sync(0xffffffffu, f64, o);
      if (lane == 0) atomicAdd(P.incut + 1, f64);
    }
  }
}

// Remainder items (kernel k_force_rem): the last r = count mod 32 sinks of a leaf (r <= REM_MAX).  As one more group of
// 32 sinks they would occupy a whole warp for the whole list with r/32 of the lanes doing useful work (a 328-particle
// leaf: 11 groups for 10.25 groups of sinks, 7 % of the kernel).  Here the roles are swapped: SOURCES are spread over the
// lanes, every lane holds the same REM_SINKS sinks, and the list is streamed once per batch of REM_SINKS sinks --
// ceil(r/8)/4 of a group's time.  A warp then consumes a 128-source tile in four iterations, so the kernel has its own,
// deeper ring (REM_STAGES tiles in flight per warp) and runs as a separate launch: the main kernel's code and register
// budget stay what they were.  Each lane accumulates the sources j = lane (mod 32) of every tile in list order; the 32
// partial sums of a sink are combined by a fixed xor-shuffle tree, so the result is deterministic (reduction order for
// the parity gate: 32 interleaved sequential sums, then the tree; all other sinks keep one sequential sum).
static constexpr int REM_SINKS = 8;     // sinks per batch (four packed pairs)
static constexpr int REM_MAX = 24;      // r > REM_MAX: four batches cost as much as a padded group
static constexpr int REM_STAGES = 8;    // ring depth of k_force_rem
static constexpr int ITEM_REM = 2;      // WorkItem::no_pseudo bit 1

struct ProducerRep { Producer pr; unsigned first_ri, total, passes_left; };
// lane 0 only: next tile of a list that is streamed `passes` times; a tile never straddles two passes
static __device__ __noinline__ void produce_tile_rep(ProducerRep &q, const ForceParams &P, float4 *tile, unsigned bar) {
  if (q.pr.remaining == 0) {
    if (q.passes_left == 0) return;
    q.passes_left--;
    q.pr.ri = q.first_ri; q.pr.roff = 0; q.pr.remaining = q.total;
  }
  produce_tile(q.pr, P, tile, bar);
}

template <int NST, int NC, int LAW, bool GUARD0, bool COUNT, bool UNITM, int FUSED>
__device__ __forceinline__ void run_item_rem(const WorkItem it, const ForceParams &P, float4 (*tiles)[FTILE],
                                             unsigned long long *bars, volatile float2 *sink_sm_v) {
  float2 *sink_sm = const_cast<float2 *>(sink_sm_v);
  constexpr int S2 = REM_SINKS / 2;
  const int lane = threadIdx.x;
  const unsigned total = P.list_len[it.node];
  const unsigned ntiles = (total + FTILE - 1) / FTILE;
  const int nbatch = (it.sink_count + REM_SINKS - 1) / REM_SINKS;
  ProducerRep q;
  q.pr.ranges = P.ranges; q.first_ri = P.range_off[it.node]; q.pr.rend = P.range_off[it.node + 1];
  q.pr.ri = q.first_ri; q.pr.roff = 0; q.pr.remaining = 0; q.total = total; q.passes_left = (unsigned)nbatch;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) produce_tile_rep(q, P, tiles[s], smem_u32(&bars[s]));
  }
  unsigned T = 0;                      // tiles consumed so far over all batches: ring stage and mbarrier parity
  unsigned long long c64 = 0, f64 = 0;
  for (int b = 0; b < nbatch; ++b) {
    SinkRegs2 k2[S2];
    // the batch's sinks go through shared memory in packed layout and come back as 64-bit loads, i.e. in aligned register
    // pairs: left to itself ptxas keeps the float4 loads and re-packs every operand pair with two MOVs inside the pair
    // loop (24 of 123 instructions per iteration)
    __syncwarp();
    if (lane < REM_SINKS) {
      const int j = b * REM_SINKS + lane;   // sinks past the end re-use the item's first sink (result discarded)
      const float4 s = __ldg(P.src4 + it.sink_begin + (j < it.sink_count ? j : 0));
      float *f = reinterpret_cast<float *>(sink_sm);
      f[(0 * S2 + (lane >> 1)) * 2 + (lane & 1)] = -s.x;
      f[(1 * S2 + (lane >> 1)) * 2 + (lane & 1)] = -s.y;
      f[(2 * S2 + (lane >> 1)) * 2 + (lane & 1)] = -s.z;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < S2; ++k) { k2[k].nx = sink_sm[0 * S2 + k]; k2[k].ny = sink_sm[1 * S2 + k]; k2[k].nz = sink_sm[2 * S2 + k]; }
#pragma unroll
    for (int k = 0; k < S2; ++k) k2[k].ax = k2[k].ay = k2[k].az = make_float2(0.f, 0.f);
    unsigned cnt[REM_SINKS], nf[REM_SINKS];
#pragma unroll
    for (int g = 0; g < REM_SINKS; ++g) { cnt[g] = 0; nf[g] = 0; }
    for (unsigned t = 0; t < ntiles; ++t, ++T) {
      const int stage = T % NST;
      const unsigned parity = (T / NST) & 1u;
      mbar_wait(smem_u32(&bars[stage]), parity);
      const unsigned nsrc = (t + 1 == ntiles) ? (total - t * FTILE) : (unsigned)FTILE;
      const float4 *tile = tiles[stage];
#pragma unroll 1
      for (unsigned j0 = 0; j0 < nsrc; j0 += 32) {
        const bool valid = j0 + lane < nsrc;
        // lanes past the end of the list see a source far outside every cutoff (finite r2, predicate false)
        const float4 s = valid ? tile[j0 + lane] : make_float4(3.0e15f, 3.0e15f, 3.0e15f, 0.f);
#pragma unroll
        for (int k = 0; k < S2; ++k) {
          if (FUSED == 2 || FUSED == 4) {
            unsigned na = 0, nb = 0;
            interact2_cull<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1], na, nb);
            if (COUNT && valid) { nf[2 * k] += na; nf[2 * k + 1] += nb; }
          } else if (FUSED) interact2_fused<NC, GUARD0, COUNT, UNITM, (FUSED >= 3)>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1]);
          else interact2<NC, LAW, GUARD0, COUNT, UNITM>(s, k2[k], P, cnt[2 * k], cnt[2 * k + 1]);
        }
      }
      __syncwarp();                       // every lane is done reading this stage
      if (lane == 0) produce_tile_rep(q, P, tiles[stage], smem_u32(&bars[stage]));
    }
    // the 32 partial sums of every sink: fixed xor tree, every lane ends up with the total
#pragma unroll
    for (int k = 0; k < S2; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        k2[k].ax.x = __fadd_rn(k2[k].ax.x, __shfl_xor_sync(0xffffffffu, k2[k].ax.x, o));
        k2[k].ax.y = __fadd_rn(k2[k].ax.y, __shfl_xor_sync(0xffffffffu, k2[k].ax.y, o));
        k2[k].ay.x = __fadd_rn(k2[k].ay.x, __shfl_xor_sync(0xffffffffu, k2[k].ay.x, o));
        k2[k].ay.y = __fadd_rn(k2[k].ay.y, __shfl_xor_sync(0xffffffffu, k2[k].ay.y, o));
        k2[k].az.x = __fadd_rn(k2[k].az.x, __shfl_xor_sync(0xffffffffu, k2[k].az.x, o));
        k2[k].az.y = __fadd_rn(k2[k].az.y, __shfl_xor_sync(0xffffffffu, k2[k].az.y, o));
      }
    }
    // kick: lane g writes sink g of the batch   (RCBForceTree.cxx:594-596 / :615-617)
    float ax = k2[0].ax.x, ay = k2[0].ay.x, az = k2[0].az.x;
#pragma unroll
    for (int g = 1; g < REM_SINKS; ++g) {
      if (lane == g) {
        ax = (g & 1) ? k2[g >> 1].ax.y : k2[g >> 1].ax.x; ay = (g & 1) ? k2[g >> 1].ay.y : k2[g >> 1].ay.x;
        az = (g & 1) ? k2[g >> 1].az.y : k2[g >> 1].az.x;
      }
    }
    const int j = b * REM_SINKS + lane;
    if (lane < REM_SINKS && j < it.sink_count) {
      const int gi = it.sink_begin + j;
      if (P.defer) {
        P.vx[gi] = ax; P.vy[gi] = ay; P.vz[gi] = az;
      } else {
        const float c = P.fcoeff * __ldg(&P.src4[gi].w);
        P.vx[gi] = fmaf(c, ax, P.vx[gi]); P.vy[gi] = fmaf(c, ay, P.vy[gi]); P.vz[gi] = fmaf(c, az, P.vz[gi]);
      }
    }
    if (COUNT) {
#pragma unroll
      for (int g = 0; g < REM_SINKS; ++g)
        if (b * REM_SINKS + g < it.sink_count) { c64 += cnt[g]; f64 += nf[g]; }
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) { c64 += __shfl_down_sync(0xffffffffu, c64, o); f64 += __shfl_down_sync(0xffffffffu, f64, o); }
    if (lane == 0) { atomicAdd(P.incut, c64); if (FUSED == 2 || FUSED == 4) atomicAdd(P.incut + 1, f64); }
  }
}

template <int NC, int LAW, bool GUARD0, bool COUNT, bool UNITM, int FUSED>
__device__ __forceinline__ void dispatch_item(int S, const WorkItem it, const ForceParams &P, float4 (*tiles)[FTILE],
                                              unsigned long long *bars) {
  switch (S) {
    case 1: run_item<0, 1, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 2: run_item<1, 0, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 3: run_item<1, 1, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 4: run_item<2, 0, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 5: run_item<2, 1, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 6: run_item<3, 0, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    case 7: run_item<3, 1, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
    default: run_item<4, 0, NC, LAW, GUARD0, COUNT, UNITM, FUSED>(it, P, tiles, bars); break;
  }
}

template <int NC, int LAW, bool GUARD0, bool COUNT, int FUSED>
__global__ void __launch_bounds__(32) k_force(const __grid_constant__ ForceParams P, int n_items) {
  __shared__ __align__(128) float4 tiles[FSTAGES][FTILE];
  __shared__ __align__(8) unsigned long long bars[FSTAGES];
  const int item = blockIdx.x;
  if (item >= n_items) return;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < FSTAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const WorkItem it = P.items[item];
  const int S = (it.sink_count + 31) / 32;
  if (LAW >= 2) {   // fit / interpolated laws: scalar form only, two code variants
    if (S <= 2) run_item<0, 2, NC, LAW, GUARD0, COUNT, false, false>(it, P, tiles, bars);
    else run_item<0, SMAX_, NC, LAW, GUARD0, COUNT, false, false>(it, P, tiles, bars);
    return;
  }
  if (P.unit_mass && (it.no_pseudo & 1)) dispatch_item<NC, LAW, GUARD0, COUNT, true, FUSED>(S, it, P, tiles, bars);
  else dispatch_item<NC, LAW, GUARD0, COUNT, false, FUSED>(S, it, P, tiles, bars);
}

// the remainder items of the same item array (LAW 0 / 1 only: the packed pair arithmetic)
template <int NC, int LAW, bool GUARD0, bool COUNT, int FUSED>
__global__ void __launch_bounds__(32) k_force_rem(const __grid_constant__ ForceParams P, int n_items) {
  __shared__ __align__(128) float4 tiles[REM_STAGES][FTILE];
  __shared__ __align__(8) unsigned long long bars[REM_STAGES];
  __shared__ __align__(16) float2 sink_sm[3 * REM_SINKS / 2];
  const int item = blockIdx.x;
  if (item >= n_items) return;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < REM_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const WorkItem it = P.items[item];
  if (P.unit_mass && (it.no_pseudo & 1)) run_item_rem<REM_STAGES, NC, LAW, GUARD0, COUNT, true, FUSED>(it, P, tiles, bars, sink_sm);
  else run_item_rem<REM_STAGES, NC, LAW, GUARD0, COUNT, false, FUSED>(it, P, tiles, bars, sink_sm);
}

// ---- work items -------------------------------------------------------------------------------------------
// The G = count / 32 full groups of a sink leaf are cut into ceil(G / SMAX_) chunks of nearly equal size; the
// r = count mod 32 sinks left over become one remainder item (run_item_rem) when r <= REM_MAX and the law has a
// packed form, else one more (padded) group.
struct LeafCut { int groups, chunks, rem; };
// policy 1: remainder items (default); 0: the r leftover sinks always become one more padded group
__device__ __forceinline__ LeafCut leaf_cut(int count, int policy) {
  LeafCut c;
  c.groups = count / 32; c.rem = count % 32;
  if (!(policy & 1) || c.rem > REM_MAX) { c.groups += c.rem ? 1 : 0; c.rem = 0; }
  const int maxg = (policy >> 4) ? (policy >> 4) : SMAX_;      // tuning: largest chunk in groups (<= SMAX_)
  c.chunks = (c.groups + maxg - 1) / maxg;
  return c;
}
// groups of chunk q (0 <= q < chunks): balanced.  (Cutting by whole packed pairs -- even group counts, the odd group
// runs the scalar form -- was measured and makes no difference: 94.1 vs 93.5 ms.)
__device__ __forceinline__ int chunk_groups(const LeafCut &c, int q) {
  return c.groups / c.chunks + (q < c.groups % c.chunks ? 1 : 0);
}
template <int NC, int LAW, bool GUARD0, int FUSED>
int launch_force(haccsr_ctx *c, const ForceParams &P0, int n_items, bool count) {
  // per group of items (a single group unless haccsr_kick_host asked for range-wise velocity copies): one launch for the
  // chunk items, one for the remainder items
  (void)n_items;       // the item ranges of the launches come from haccsr_ctx::seg_off
  const int groups = c->force_groups;
  for (int g = 0; g < groups; ++g) {
    for (int rem = 0; rem < 2; ++rem) {
      const int b = (int)c->seg_off[2 * g + rem], e = (int)c->seg_off[2 * g + rem + 1];
      if (e <= b) continue;
      ForceParams P = P0;
      P.items = P0.items + b;
      if (rem) {
        if (LAW <= 1) {
          if (count) k_force_rem<NC, (LAW <= 1 ? LAW : 0), GUARD0, true, FUSED><<<e - b, 32, 0, c->stream>>>(P, e - b);
          else k_force_rem<NC, (LAW <= 1 ? LAW : 0), GUARD0, false, FUSED><<<e - b, 32, 0, c->stream>>>(P, e - b);
        } else { set_error("remainder items for a law without a packed kernel"); return 1; }
      } else {
        if (count) k_force<NC, LAW, GUARD0, true, FUSED><<<e - b, 32, 0, c->stream>>>(P, e - b);
        else k_force<NC, LAW, GUARD0, false, FUSED><<<e - b, 32, 0, c->stream>>>(P, e - b);
      }
      c->launches++;
      if (!rem) c->force_launches++;     // haccsr_stats::force_launches counts the groups (k_force launches)
      HSR_CUDA(cudaGetLastError());
    }
    if (P0.defer) HSR_TRY(apply_kick(c, P0.fcoeff, c->group_lo[g], c->group_lo[g + 1]));
    if (groups > 1 && c->ho_v[0]) {
      // velocities of particles [lo, hi) are final once every launch up to this one is done
      const int64_t lo = c->group_lo[g], hi = c->group_lo[g + 1];
      HSR_CUDA(cudaEventRecord(c->ev_grp[g], c->stream));
      HSR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_grp[g], 0));
      float *dv[3] = {c->cur.vx, c->cur.vy, c->cur.vz};
      for (int q = 0; q < 3; ++q)
        HSR_CUDA(cudaMemcpyAsync(c->ho_v[q] + lo, dv[q] + lo, (size_t)(hi - lo) * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  return 0;
}

}  // namespace haccsr
