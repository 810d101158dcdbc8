/* oracle/haccsr_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by hacc_coral_b200/).
 *
 * A plain-C CPU restatement of the reference's short-range hot path, written from the rules in the
 * reference (file:line cited at each function; all paths relative to the reference root):
 *   tree build   src/halo_finder/RCBForceTree.cxx:778-890 (node), :689-775 (split), :623-672 (partition),
 *                src/halo_finder/BGQCM.c:181-212 (tight box + centre of mass)
 *   walk         src/halo_finder/RCBForceTree.cxx:923-1144 (interaction list), :1150-1197 (leaf scheduler)
 *   force kernel src/halo_finder/RCBForceTree.cxx:601-618 with ForceLaw.cxx:137-141,187-192 (generic form)
 *                src/halo_finder/BGQStep16.c:170-187 (BG/Q scalar-tail form)
 * It exists because the compiled reference (oracle/_ref) (a) aborts when a list exceeds VMAX=16384,
 * (b) never exposes its interaction lists and (c) cannot travel as source.  PARITY IS PINNED: tests/
 * check that this restatement reproduces the compiled reference bit-for-bit (kicked velocities, node
 * table, permutation) on seeded snapshots, and against fixtures in tests/golden/ generated from it.
 *
 * Monopole (TDPTS = 1) and quadrupole (TDPTS = 12 pseudo-particles on an icosahedron, :229-272,519-569;
 * selected with orc_set_tdpts).  Arithmetic is kept in the reference's order, in float, with no FMA
 * contraction (build with -ffp-contract=off), so results are comparable bit-for-bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t count, offset, cl, cr;
  float ppm, tdr;          /* ppm: the monopole mass (TDPTS = 1) */
  float xmin[3], xmax[3], xc[3];
  float ppm12[12];         /* pseudo-particle masses for TDPTS = 12 */
} orc_node;

/* TDPTS of the RCBForceTree<TDPTS> instantiation being restated: 1 (RCBMonopoleForceTree, -R) or
 * 12 (RCBQuadrupoleForceTree, -S); RCBForceTree.h:202-203. */
static int g_tdpts = 1;
void orc_set_tdpts(int t) { g_tdpts = (t == 12) ? 12 : 1; }
static const float g_ppc = 0.9f;   /* ppContract, the constructor's `ppc` default (RCBForceTree.h:123) */

/* The 12-point spherical 4-design of Hardin & Sloane (vertices of an icosahedron) in the point order of
 * RCBForceTree.cxx:229-272: p = 0.525731112119134, q = 0.85065080835204. */
#define ICO_P 0.525731112119134f
#define ICO_Q 0.85065080835204f
static const float g_d12x[12] = { 0, 0, ICO_P, -ICO_P, ICO_Q, -ICO_Q, 0, 0, -ICO_P, ICO_P, -ICO_Q, ICO_Q };
static const float g_d12y[12] = { ICO_Q, ICO_Q, 0, 0, ICO_P, ICO_P, -ICO_Q, -ICO_Q, 0, 0, -ICO_P, -ICO_P };
static const float g_d12z[12] = { ICO_P, -ICO_P, ICO_Q, ICO_Q, 0, 0, -ICO_P, ICO_P, -ICO_Q, -ICO_Q, 0, 0 };

/* :519-523 */
static float orc_pptdr(const float *xmin, const float *xmax, const float *xc) {
  return fminf(xmax[0] - xc[0], fminf(xmax[1] - xc[1], fminf(xmax[2] - xc[2], fminf(xc[0] - xmin[0],
               fminf(xc[1] - xmin[1], xc[2] - xmin[2])))));
}
/* :525-534 */
static void orc_pppts12(float tdr, const float *xc, float *ppx, float *ppy, float *ppz) {
  for (int i = 0; i < 12; ++i) {
    ppx[i] = tdr * g_d12x[i] + xc[0];
    ppy[i] = tdr * g_d12y[i] + xc[1];
    ppz[i] = tdr * g_d12z[i] + xc[2];
  }
}
/* :536-569 with TDPTS = 12: monopole + dipole + quadrupole weights of `count` sources onto the 12 points.
 * The 0.5 of :561 is a double literal: the product is formed in double and rounded to float once. */
static void orc_pp12(int64_t count, const float *xx, const float *yy, const float *zz, const float *mass,
                     const float *xc, const float *ppx, const float *ppy, const float *ppz, float *ppm, float tdr) {
  float K = 12;
  float odr0 = 1 / K;
  for (int64_t i = 0; i < count; ++i) {
    float xi = xx[i] - xc[0], yi = yy[i] - xc[1], zi = zz[i] - xc[2];
    float ri = sqrtf(xi*xi + yi*yi + zi*zi);
    for (int j = 0; j < 12; ++j) {
      float xj = ppx[j] - xc[0], yj = ppy[j] - xc[1], zj = ppz[j] - xc[2];
      float rj2 = xj*xj + yj*yj + zj*zj;
      float odr1 = 0, odr2 = 0;
      if (rj2 != 0) {
        float rj = sqrtf(rj2);
        float aij = (xi*xj + yi*yj + zi*zj) / (ri*rj);
        odr1 = (3/K)*(ri/tdr)*aij;
        odr2 = (float)((5/K)*(ri/tdr)*(ri/tdr)*0.5*(3*aij*aij - 1));
      }
      ppm[j] += mass[i]*(odr0 + odr1 + odr2);
    }
  }
}

typedef struct {
  /* inputs */
  int64_t n;
  float *x, *y, *z, *m;   /* working copies, permuted in place into tree order */
  int64_t *perm;          /* perm[i] = original index of the particle now at i  */
  int64_t ppn;
  /* node pool */
  orc_node *node;
  int64_t nnode, cap;
} orc_tree;

/* ---- BGQCM.c:181-212 : tight bounding box (fminf/fmaxf) and mass-weighted centroid in double.
 * The product w*x is formed in float and then added to a double accumulator (:203-206). */
static void orc_cm(int64_t cnt, const float *xx, const float *yy, const float *zz, const float *mass,
                   float *xmin, float *xmax, float *xc) {
  double x = 0, y = 0, z = 0, m = 0;
  for (int64_t i = 0; i < cnt; ++i) {
    if (i == 0) {
      xmin[0] = xmax[0] = xx[0]; xmin[1] = xmax[1] = yy[0]; xmin[2] = xmax[2] = zz[0];
    } else {
      xmin[0] = fminf(xmin[0], xx[i]); xmax[0] = fmaxf(xmax[0], xx[i]);
      xmin[1] = fminf(xmin[1], yy[i]); xmax[1] = fmaxf(xmax[1], yy[i]);
      xmin[2] = fminf(xmin[2], zz[i]); xmax[2] = fmaxf(xmax[2], zz[i]);
    }
    float w = mass[i];
    float wx = w * xx[i], wy = w * yy[i], wz = w * zz[i];
    x += wx; y += wy; z += wz; m += w;
  }
  xc[0] = (float)(x / m); xc[1] = (float)(y / m); xc[2] = (float)(z / m);
}

/* ---- RCBForceTree.cxx:623-672 : pivot partition.  Indices with key < pv are gathered in order (:638-645),
 * then element idx[j] is swapped with element j for j < is (:648-669).  Left block keeps input order. */
static int64_t orc_partition(orc_tree *t, int64_t off, int64_t n, int d, float pv, int64_t *idx) {
  float *key = (d == 0 ? t->x : d == 1 ? t->y : t->z) + off;
  float *x = t->x + off, *y = t->y + off, *z = t->z + off, *m = t->m + off;
  int64_t *p = t->perm + off;
  int64_t is = 0;
  for (int64_t i = 0; i < n; ++i) if (key[i] < pv) idx[is++] = i;
  for (int64_t j = 0; j < is; ++j) {
    int64_t i = idx[j];
    float f; int64_t q;
    f = x[i]; x[i] = x[j]; x[j] = f;
    f = y[i]; y[i] = y[j]; y[j] = f;
    f = z[i]; z[i] = z[j]; z[j] = f;
    f = m[i]; m[i] = m[j]; m[j] = f;
    q = p[i]; p[i] = p[j]; p[j] = q;
  }
  return is;
}

static int64_t orc_alloc2(orc_tree *t) {
  if (t->nnode + 2 > t->cap) {
    t->cap = t->cap * 2 + 2;
    t->node = (orc_node *)realloc(t->node, (size_t)t->cap * sizeof(orc_node));
  }
  int64_t a = t->nnode;
  memset(&t->node[a], 0, 2 * sizeof(orc_node));   /* :825 */
  t->nnode += 2;
  return a;
}

/* ---- RCBForceTree.cxx:778-890 (unthreaded order: node, then left subtree, then right subtree). */
static void orc_build_node(orc_tree *t, int64_t tl, int64_t *idx) {
  int64_t cnt = t->node[tl].count, off = t->node[tl].offset;
  orc_cm(cnt, t->x + off, t->y + off, t->z + off, t->m + off, t->node[tl].xmin, t->node[tl].xmax,
         t->node[tl].xc);                                                       /* :785-786 */
  if (cnt <= t->ppn) {                                                          /* :788 leaf */
    float s = 0.0f;
    /* pp<1> with the design point at the centre: ppm += mass*(1/K + 0 + 0), K = 1 (:536-569) */
    if (cnt > 1) for (int64_t i = 0; i < cnt; ++i) s += t->m[off + i] * (1.0f + 0.0f + 0.0f);
    t->node[tl].ppm = s;
    if (g_tdpts == 12) {                                                        /* :789-797 */
      orc_node *nd = &t->node[tl];
      nd->tdr = g_ppc * orc_pptdr(nd->xmin, nd->xmax, nd->xc);
      memset(nd->ppm12, 0, sizeof(nd->ppm12));
      if (cnt > 12) {
        float ppx[12], ppy[12], ppz[12];
        orc_pppts12(nd->tdr, nd->xc, ppx, ppy, ppz);
        orc_pp12(cnt, t->x + off, t->y + off, t->z + off, t->m + off, nd->xc, ppx, ppy, ppz, nd->ppm12, nd->tdr);
      }
    }
    return;
  }
  int64_t cl = orc_alloc2(t), cr = cl + 1;                                      /* :808-809 */
  orc_node *nd = &t->node[tl];
  for (int k = 0; k < 3; ++k) {                                                 /* :830-835 */
    t->node[cl].xmin[k] = t->node[cr].xmin[k] = nd->xmin[k];
    t->node[cl].xmax[k] = t->node[cr].xmax[k] = nd->xmax[k];
  }
  float xl0 = nd->xmax[0] - nd->xmin[0], xl1 = nd->xmax[1] - nd->xmin[1], xl2 = nd->xmax[2] - nd->xmin[2];
  int d = (xl0 > xl1 && xl0 > xl2) ? 0 : (xl1 > xl2) ? 1 : 2;                   /* :844-852 */
  float split = nd->xc[d];                                                      /* :720 */
  int64_t is = orc_partition(t, off, cnt, d, split, idx);                       /* :721-725 */
  if (!(is == 0 || is == cnt)) {                                                /* :727-729 */
    t->node[cl].count = is;
    t->node[cr].count = cnt - is;
    /* both counts > 0 here */
    t->node[tl].cl = cl; t->node[cl].offset = off;      t->node[cl].xmax[d] = split;   /* :744-747 */
    orc_build_node(t, cl, idx);
    t->node[tl].cr = cr; t->node[cr].offset = off + is; t->node[cr].xmin[d] = split;   /* :760-763 */
    orc_build_node(t, cr, idx);
  }
  /* parent moment from the children (:856-889): children with <= 1 particle contribute the particle
   * mass itself, others their ppm.  After a degenerate split both children are empty and ppm stays 0. */
  float s = 0.0f;
  for (int side = 0; side < 2; ++side) {
    int64_t c = side ? cr : cl;
    if (t->node[c].count > 0) {
      if (t->node[c].count <= 1) s += t->m[t->node[c].offset] * (1.0f + 0.0f + 0.0f);
      else s += t->node[c].ppm * (1.0f + 0.0f + 0.0f);
    }
  }
  t->node[tl].ppm = s;
  if (g_tdpts == 12) {                                                          /* :856-889 */
    orc_node *nd = &t->node[tl];
    float ppx[12], ppy[12], ppz[12];
    nd->tdr = g_ppc * orc_pptdr(nd->xmin, nd->xmax, nd->xc);
    orc_pppts12(nd->tdr, nd->xc, ppx, ppy, ppz);
    memset(nd->ppm12, 0, sizeof(nd->ppm12));
    for (int side = 0; side < 2; ++side) {
      const orc_node *ch = &t->node[side ? cr : cl];
      if (ch->count <= 0) continue;
      if (ch->count <= 12) {
        int64_t oc = ch->offset;
        orc_pp12(ch->count, t->x + oc, t->y + oc, t->z + oc, t->m + oc, nd->xc, ppx, ppy, ppz, nd->ppm12, nd->tdr);
      } else {
        float cx[12], cy[12], cz[12];
        orc_pppts12(ch->tdr, ch->xc, cx, cy, cz);
        orc_pp12(12, cx, cy, cz, ch->ppm12, nd->xc, ppx, ppy, ppz, nd->ppm12, nd->tdr);
      }
    }
  }
}

/* ---- force laws --------------------------------------------------------------------------- */
typedef struct {
  int kind;           /* 0: polynomial grid force (ncoef <= 7), 3: newton */
  float a[7];
  float rsm2, r2min, r2max;
} orc_law;

/* ForceLawSR::f_over_r over FGridEvalPoly::eval (ForceLaw.cxx:137-141, 187-192) */
static inline float orc_f_over_r(const orc_law *L, float r2) {
  if (L->kind == 3) return (float)(1.0 / r2 / sqrt(r2));  /* ForceLawNewton (ForceLaw.h:106) */
  const float *a = L->a;
  float poly = a[0] + r2*(a[1] + r2*(a[2] + r2*(a[3] + r2*(a[4] + r2*(a[5] + r2*a[6])))));
  poly = poly * (r2 >= L->r2min) * (r2 <= L->r2max);
  float ret = powf(r2 + L->rsm2, -1.5f) - poly;
  ret *= (r2 >= L->r2min) * (r2 <= L->r2max);
  return ret;
}

/* ---- the interaction list of one sink leaf -------------------------------------------------- */
typedef struct { int64_t *node; uint8_t *pseudo; int64_t n, cap; } orc_list;

static void orc_list_push(orc_list *l, int64_t node, int pseudo) {
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 256;
    l->node = (int64_t *)realloc(l->node, (size_t)l->cap * sizeof(int64_t));
    l->pseudo = (uint8_t *)realloc(l->pseudo, (size_t)l->cap);
  }
  l->node[l->n] = node; l->pseudo[l->n] = (uint8_t)pseudo; l->n++;
}

/* RCBForceTree.cxx:923-1144.  `anc` = sorted ancestor indices of tl (the `parents` vector).
 * Emits list entries in the reference's order: (node, pseudo=1) = the node's pseudo-particle,
 * (node, pseudo=0) = all real particles of the node; tl itself comes last (:1126-1139). */
static void orc_walk(const orc_tree *t, int64_t tl, const int64_t *anc, int nanc, float rmax,
                     float tan_oa, orc_list *out, int64_t **stk, int64_t *stkcap) {
  const orc_node *T = t->node;
  float rmax2 = rmax * rmax;
  int64_t sp = 0;
#define PUSH(v) do { if (sp == *stkcap) { *stkcap = *stkcap ? 2 * *stkcap : 256; \
    *stk = (int64_t *)realloc(*stk, (size_t)*stkcap * sizeof(int64_t)); } (*stk)[sp++] = (v); } while (0)
  PUSH(0);
  while (sp > 0) {
    int64_t tln = (*stk)[--sp];
    if (tln < tl) {                                                            /* :947-962 */
      int isp = 0;
      for (int lo = 0, hi = nanc - 1; lo <= hi;) {
        int mid = (lo + hi) / 2;
        if (anc[mid] == tln) { isp = 1; break; }
        if (anc[mid] < tln) lo = mid + 1; else hi = mid - 1;
      }
      if (isp) {
        int64_t cr = T[tln].cr, cl = T[tln].cl;
        if (cl != tl && cl > 0 && T[cl].count > 0) PUSH(cl);
        if (cr != tl && cr > 0 && T[cr].count > 0) PUSH(cr);
        continue;
      }
    }
    float dx = T[tln].xc[0] - T[tl].xc[0], dy = T[tln].xc[1] - T[tl].xc[1], dz = T[tln].xc[2] - T[tl].xc[2];
    float dist2 = dx*dx + dy*dy + dz*dz;                                       /* :965-968 */
    float sx = T[tln].xmax[0] - T[tln].xmin[0], sy = T[tln].xmax[1] - T[tln].xmin[1], sz = T[tln].xmax[2] - T[tln].xmin[2];
    float l2 = fminf(sx*sx, fminf(sy*sy, sz*sz));                              /* :970-973 */
    float dtt2 = dist2 * tan_oa * tan_oa;                                      /* :975 */
    int big;
    if (l2 > dtt2) big = 1;
    else {
      /* :986-1018 with useRealOA=false: all four (i,j) corner pairs give the same diagonal^2 */
      big = 0;
      for (int i = 0; i < 2 && !big; ++i)
        for (int j = 0; j < 2; ++j) {
          float x1 = (i == 0 ? T[tln].xmin : T[tln].xmax)[0] - T[tl].xc[0];
          float y1 = (j == 0 ? T[tln].xmin : T[tln].xmax)[1] - T[tl].xc[1];
          float z1 = T[tln].xmin[2] - T[tl].xc[2];
          float x2 = (i == 0 ? T[tln].xmax : T[tln].xmin)[0] - T[tl].xc[0];
          float y2 = (j == 0 ? T[tln].xmax : T[tln].xmin)[1] - T[tl].xc[1];
          float z2 = T[tln].xmax[2] - T[tl].xc[2];
          float ddx = x1 - x2, ddy = y1 - y2, ddz = z1 - z2;
          float dh2 = ddx*ddx + ddy*ddy + ddz*ddz;
          if (dh2 > dtt2) { big = 1; break; }
        }
    }
    if (!big) {
      if (dist2 > rmax2) continue;                                             /* :1024-1029 */
      if (T[tln].count <= g_tdpts) orc_list_push(out, tln, 0);                 /* :1033-1049 */
      else orc_list_push(out, tln, 1);                                         /* :1053-1062 */
      continue;
    } else if (T[tln].cr == 0 && T[tln].cl == 0) {                             /* :1063-1080 */
      orc_list_push(out, tln, 0);
      continue;
    }
    int64_t ch[2] = { T[tln].cl, T[tln].cr };                                  /* :1085-1123 */
    for (int s = 0; s < 2; ++s) {
      int64_t c = ch[s];
      if (c > 0 && T[c].count > 0) {
        int close = 1;
        for (int i = 0; i < 3; ++i) {
          float dist = 0;
          if (T[tl].xmax[i] < T[c].xmin[i]) dist = T[c].xmin[i] - T[tl].xmax[i];
          else if (T[tl].xmin[i] > T[c].xmax[i]) dist = T[tl].xmin[i] - T[c].xmax[i];
          if (dist > rmax) { close = 0; break; }
        }
        if (close) PUSH(c);
      }
    }
  }
#undef PUSH
  orc_list_push(out, tl, 0);                                                   /* :1126-1139 */
}

/* ---- public entry point -------------------------------------------------------------------- */
typedef struct {
  int64_t nodes, leaves, empty_leaves, max_ppn, sink_leaves, max_list;
  double mean_ppn;
  uint64_t pairs_eval, pairs_incut;
} orc_stats;

typedef struct orc_result {
  orc_tree t;
  /* CSR of interaction lists per processed sink leaf */
  int64_t nsink; int64_t *sink_leaf; int64_t *list_off; int64_t *list_node; uint8_t *list_pseudo;
} orc_result;

static void orc_collect_leaves(const orc_tree *t, const float *flo, const float *fhi, int64_t **leaf,
                               int64_t **ancoff, int64_t **anc, int64_t *nleaf) {
  /* RCBForceTree.cxx:1150-1197: DFS, children pushed left then right (right visited first),
   * ancestor list = the DFS path; inside test :1166-1170 on the leaf's tight box. */
  int64_t cap = 1024, nl = 0, acap = 4096, na = 0;
  *leaf = (int64_t *)malloc(cap * sizeof(int64_t));
  *ancoff = (int64_t *)malloc((cap + 1) * sizeof(int64_t));
  *anc = (int64_t *)malloc(acap * sizeof(int64_t));
  int64_t *q = (int64_t *)malloc((size_t)(t->nnode + 2) * sizeof(int64_t)), qn = 0;
  int64_t *par = (int64_t *)malloc((size_t)(t->nnode + 2) * sizeof(int64_t)), pn = 0;
  q[qn++] = 0;
  (*ancoff)[0] = 0;
  while (qn > 0) {
    int64_t tl = q[qn - 1];
    const orc_node *nd = &t->node[tl];
    if (nd->cr == 0 && nd->cl == 0) {
      qn--;
      int inside = 1;
      for (int i = 0; i < 3; ++i)
        inside &= (nd->xmax[i] < fhi[i] && nd->xmax[i] > flo[i]) || (nd->xmin[i] < fhi[i] && nd->xmin[i] > flo[i]);
      if (inside) {
        if (nl == cap) { cap *= 2; *leaf = (int64_t *)realloc(*leaf, cap * sizeof(int64_t));
                         *ancoff = (int64_t *)realloc(*ancoff, (cap + 1) * sizeof(int64_t)); }
        while (na + pn > acap) { acap *= 2; *anc = (int64_t *)realloc(*anc, acap * sizeof(int64_t)); }
        memcpy(*anc + na, par, (size_t)pn * sizeof(int64_t));
        na += pn;
        (*leaf)[nl] = tl; (*ancoff)[nl + 1] = na; nl++;
      }
    } else if (pn > 0 && par[pn - 1] == tl) {
      pn--; qn--;
    } else {
      if (nd->cl > 0) q[qn++] = nd->cl;
      if (nd->cr > 0) q[qn++] = nd->cr;
      par[pn++] = tl;
    }
  }
  free(q); free(par);
  *nleaf = nl;
}

/* The polynomial sum_j a_j r2^j re-expanded in s = r2 + eps: b_k = sum_{j>=k} a_j C(j,k) (-eps)^(j-k), in double,
 * rounded once (the same loop as haccsr_set_force_law, hacc_coral_b200/csrc/api.cu). */
static void orc_shifted_poly(const float *a, int ncoef, float eps_f, float *b) {
  const double eps = (double)eps_f;
  for (int k = 0; k < ncoef; ++k) {
    double acc = 0.0, binom = 1.0, pw = 1.0;
    for (int j = k; j < ncoef; ++j) {
      acc += (double)a[j] * binom * pw;
      binom = binom * (double)(j + 1) / (double)(j + 1 - k);
      pw *= -eps;
    }
    b[k] = (float)acc;
  }
}

/* kernel_form: 0 generic nbody1 (RCBForceTree.cxx:604-618), 1 BG/Q scalar tail (BGQStep16.c:170-187 +
 * RCBForceTree.cxx:594-596), 2 = form 1 evaluated and accumulated in double (the "FP64 sum of the same
 * pair set" used to put both implementations' rounding errors on one scale), 3 = like 2 but returns in vx
 * the gross sum  sum_j |f_ij||d_ij|  (vy = vz = 0): the magnitude FP32 summation error scales with;
 * 4 = the fused arithmetic of libhaccsr's default mode (see below). */
orc_result *orc_run(int64_t n, const float *x, const float *y, const float *z, const float *mass,
                    float *vx, float *vy, float *vz, /* in tree order on return: v += kick */
                    const float *boxes /* treeLo treeHi forceLo forceHi */, int law_kind, const float *coef,
                    int ncoef, float rsm, float rmax, float theta, int64_t ppn, float fcoeff,
                    int kernel_form, int do_force, int keep_lists, orc_stats *st) {
  orc_result *R = (orc_result *)calloc(1, sizeof(orc_result));
  orc_tree *t = &R->t;
  t->n = n; t->ppn = ppn;
  t->x = (float *)malloc((size_t)(n + 1) * 4); t->y = (float *)malloc((size_t)(n + 1) * 4);
  t->z = (float *)malloc((size_t)(n + 1) * 4); t->m = (float *)malloc((size_t)(n + 1) * 4);
  t->perm = (int64_t *)malloc((size_t)(n + 1) * 8);
  memcpy(t->x, x, (size_t)n * 4); memcpy(t->y, y, (size_t)n * 4); memcpy(t->z, z, (size_t)n * 4);
  memcpy(t->m, mass, (size_t)n * 4);
  for (int64_t i = 0; i < n; ++i) t->perm[i] = i;
  t->cap = 1024; t->node = (orc_node *)malloc((size_t)t->cap * sizeof(orc_node)); t->nnode = 1;
  memset(&t->node[0], 0, sizeof(orc_node));
  t->node[0].count = n; t->node[0].offset = 0;                                  /* :900-910 */
  for (int k = 0; k < 3; ++k) { t->node[0].xmin[k] = boxes[k]; t->node[0].xmax[k] = boxes[3 + k]; }
  int64_t *idx = (int64_t *)malloc((size_t)(n + 1) * 8);
  if (n > 0) orc_build_node(t, 0, idx);
  free(idx);

  /* velocities follow the particles into tree order (the reference permutes all 10 arrays) */
  {
    float *tv = (float *)malloc((size_t)(n + 1) * 4);
    float *vv[3] = { vx, vy, vz };
    for (int c = 0; c < 3; ++c) {
      for (int64_t i = 0; i < n; ++i) tv[i] = vv[c][t->perm[i]];
      memcpy(vv[c], tv, (size_t)n * 4);
    }
    free(tv);
  }

  orc_law L; memset(&L, 0, sizeof(L));
  L.kind = law_kind;
  for (int i = 0; i < 7; ++i) L.a[i] = (coef && i < ncoef) ? coef[i] : 0.0f;
  L.rsm2 = rsm * rsm; L.r2min = 0.0f; L.r2max = rmax * rmax;
  float tan_oa = tanf(theta);                                                   /* :381 */
  float rmax2 = rmax * rmax;

  int64_t *leaf, *ancoff, *anc, nleaf;
  orc_collect_leaves(t, boxes + 6, boxes + 9, &leaf, &ancoff, &anc, &nleaf);

  orc_list *lists = (orc_list *)calloc((size_t)(nleaf ? nleaf : 1), sizeof(orc_list));
  uint64_t pe = 0, pc = 0; int64_t maxlist = 0;
#pragma omp parallel reduction(+:pe,pc) reduction(max:maxlist)
  {
    int64_t *stk = 0, stkcap = 0;
    float *nx = 0, *ny = 0, *nz = 0, *nm = 0; int64_t ncap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int64_t li = 0; li < nleaf; ++li) {
      int64_t tl = leaf[li];
      orc_list *l = &lists[li];
      orc_walk(t, tl, anc + ancoff[li], (int)(ancoff[li + 1] - ancoff[li]), rmax, tan_oa, l, &stk, &stkcap);
      /* materialise the source list like the reference does (nx,ny,nz,nm; :1033-1080,1126-1139) */
      int64_t size = 0;
      for (int64_t e = 0; e < l->n; ++e) size += l->pseudo[e] ? g_tdpts : t->node[l->node[e]].count;
      if (size > maxlist) maxlist = size;
      if (do_force) {
        if (size > ncap) {
          ncap = size * 2;
          nx = (float *)realloc(nx, (size_t)ncap * 4); ny = (float *)realloc(ny, (size_t)ncap * 4);
          nz = (float *)realloc(nz, (size_t)ncap * 4); nm = (float *)realloc(nm, (size_t)ncap * 4);
        }
        int64_t s = 0;
        for (int64_t e = 0; e < l->n; ++e) {
          const orc_node *nd = &t->node[l->node[e]];
          if (l->pseudo[e] && g_tdpts == 12) {                                  /* :1053-1062 */
            orc_pppts12(nd->tdr, nd->xc, nx + s, ny + s, nz + s);
            for (int q = 0; q < 12; ++q) nm[s + q] = nd->ppm12[q];
            s += 12;
          } else if (l->pseudo[e]) { nx[s] = nd->xc[0]; ny[s] = nd->xc[1]; nz[s] = nd->xc[2]; nm[s] = nd->ppm; s++; }
          else for (int64_t i = 0; i < nd->count; ++i) {
            nx[s] = t->x[nd->offset + i]; ny[s] = t->y[nd->offset + i]; nz[s] = t->z[nd->offset + i];
            nm[s] = t->m[nd->offset + i]; s++;
          }
        }
        int64_t off = t->node[tl].offset, cnt = t->node[tl].count;
        for (int64_t i = 0; i < cnt; ++i) {
          float xi = t->x[off + i], yi = t->y[off + i], zi = t->z[off + i], mi = t->m[off + i];
          if (kernel_form == 0) {
            float ax = vx[off + i], ay = vy[off + i], az = vz[off + i];
            for (int64_t j = 0; j < size; ++j) {
              float dx = nx[j] - xi, dy = ny[j] - yi, dz = nz[j] - zi;
              float dist2 = dx*dx + dy*dy + dz*dz;
              float f_over_r = mi * nm[j] * orc_f_over_r(&L, dist2);
              float updateq = 1.0f; updateq *= (dist2 < rmax2);
              ax += updateq*fcoeff*f_over_r*dx; ay += updateq*fcoeff*f_over_r*dy; az += updateq*fcoeff*f_over_r*dz;
              pc += (dist2 > 0.0f && dist2 < rmax2);
            }
            vx[off + i] = ax; vy[off + i] = ay; vz[off + i] = az;
          } else if (kernel_form == 4) {
            /* FORM_FUSED: the contracted arithmetic of libhaccsr's HACCSR_ARITH_FUSED mode (include/haccsr.h), i.e. the
             * multiply-add chain of the QPX loop (BGQStep16.c:76-86) seeded with rsm^2, polynomial and cutoff in
             * s = r2 + rsm^2, accumulate-then-scale as RCBForceTree.cxx:594-596.  Pins that mode's in-cutoff pair set. */
            float ax = 0.0f, ay = 0.0f, az = 0.0f;
            float bb[8]; orc_shifted_poly(L.a, ncoef, L.rsm2, bb);
            const int nc = ncoef <= 6 ? 6 : 7;
            for (int k = ncoef; k < 8; ++k) bb[k] = 0.0f;
            const float smax = rmax2 + L.rsm2;
            for (int64_t j = 0; j < size; ++j) {
              float dx = nx[j] - xi, dy = ny[j] - yi, dz = nz[j] - zi;
              float s = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, L.rsm2)));
              float pl = -bb[nc - 1];
              for (int q = nc - 2; q >= 0; --q) pl = fmaf(pl, s, -bb[q]);      /* pl = -g(s) */
              float rs = (float)(1.0 / sqrt((double)s));
              float f = fmaf(rs * rs, rs, pl);
              f = f * nm[j];
              int in = s < smax;
              if (in) { ax = fmaf(f, dx, ax); ay = fmaf(f, dy, ay); az = fmaf(f, dz, az); }
              pc += (in && s > L.rsm2);
            }
            float c = fcoeff * mi;
            vx[off + i] = fmaf(c, ax, vx[off + i]); vy[off + i] = fmaf(c, ay, vy[off + i]); vz[off + i] = fmaf(c, az, vz[off + i]);
          } else if (kernel_form == 1) {
            float ax = 0.0f, ay = 0.0f, az = 0.0f;
            const float *a = L.a;
            for (int64_t j = 0; j < size; ++j) {
              float dx = nx[j] - xi, dy = ny[j] - yi, dz = nz[j] - zi;
              float r2 = dx*dx + dy*dy + dz*dz;
              float m = (r2 < rmax2) ? nm[j] : 0.0f;
              float f = (float)(pow(r2 + L.rsm2, -1.5) - (a[0] + r2*(a[1] + r2*(a[2] + r2*(a[3] + r2*(a[4] + r2*a[5]))))));
              f = (r2 > 0.0f) ? m * f : 0.0f;
              ax = ax + f * dx; ay = ay + f * dy; az = az + f * dz;
              pc += (r2 > 0.0f && r2 < rmax2);
            }
            vx[off + i] = vx[off + i] + ax * fcoeff; vy[off + i] = vy[off + i] + ay * fcoeff; vz[off + i] = vz[off + i] + az * fcoeff;
          } else {
            double ax = 0, ay = 0, az = 0, gross = 0;
            for (int64_t j = 0; j < size; ++j) {
              double dx = (double)nx[j] - xi, dy = (double)ny[j] - yi, dz = (double)nz[j] - zi;
              double r2 = dx*dx + dy*dy + dz*dz;
              /* same pair set as the float kernels: the cutoff decision is taken on the float r2 */
              float fdx = nx[j] - xi, fdy = ny[j] - yi, fdz = nz[j] - zi;
              float fr2 = fdx*fdx + fdy*fdy + fdz*fdz;
              if (!(fr2 < rmax2) || !(fr2 > 0.0f)) continue;
              double poly = 0; for (int k = 6; k >= 0; --k) poly = poly * r2 + (double)L.a[k];
              double f = (double)mi * nm[j] * (pow(r2 + (double)L.rsm2, -1.5) - poly);
              ax += f * dx; ay += f * dy; az += f * dz;
              gross += fabs(f) * sqrt(r2);
              pc++;
            }
            if (kernel_form == 3) {   /* gross = sum_j |f_ij| |d_ij|: the scale FP32 rounding error is relative to */
              vx[off + i] = (float)(gross * fabs((double)fcoeff)); vy[off + i] = 0.0f; vz[off + i] = 0.0f;
            } else {
              vx[off + i] = (float)(vx[off + i] + ax * fcoeff); vy[off + i] = (float)(vy[off + i] + ay * fcoeff);
              vz[off + i] = (float)(vz[off + i] + az * fcoeff);
            }
          }
        }
      }
      pe += (uint64_t)size * (uint64_t)t->node[tl].count;
      if (!keep_lists) { free(l->node); free(l->pseudo); l->node = 0; l->pseudo = 0; l->cap = 0; }
    }
    free(stk); free(nx); free(ny); free(nz); free(nm);
  }

  if (keep_lists) {
    R->nsink = nleaf;
    R->sink_leaf = (int64_t *)malloc((size_t)(nleaf + 1) * 8);
    R->list_off = (int64_t *)malloc((size_t)(nleaf + 1) * 8);
    int64_t tot = 0;
    for (int64_t li = 0; li < nleaf; ++li) { R->sink_leaf[li] = leaf[li]; R->list_off[li] = tot; tot += lists[li].n; }
    R->list_off[nleaf] = tot;
    R->list_node = (int64_t *)malloc((size_t)(tot + 1) * 8);
    R->list_pseudo = (uint8_t *)malloc((size_t)(tot + 1));
    for (int64_t li = 0; li < nleaf; ++li) {
      memcpy(R->list_node + R->list_off[li], lists[li].node, (size_t)lists[li].n * 8);
      memcpy(R->list_pseudo + R->list_off[li], lists[li].pseudo, (size_t)lists[li].n);
      free(lists[li].node); free(lists[li].pseudo);
    }
  }
  free(lists);

  if (st) {
    memset(st, 0, sizeof(*st));
    st->nodes = t->nnode; st->sink_leaves = nleaf; st->max_list = maxlist;
    int64_t nz = 0, parts = 0;
    for (int64_t i = 1; i < t->nnode; ++i)       /* census as printStats (:468-478), from node 1 */
      if (t->node[i].cl == 0 && t->node[i].cr == 0) {
        if (t->node[i].count > 0) { nz++; parts += t->node[i].count; if (t->node[i].count > st->max_ppn) st->max_ppn = t->node[i].count; }
        else st->empty_leaves++;
      }
    st->leaves = nz + st->empty_leaves;
    st->mean_ppn = nz ? parts / (double)nz : 0.0;
    st->pairs_eval = pe; st->pairs_incut = do_force ? pc : 0;
  }
  free(leaf); free(ancoff); free(anc);
  return R;
}

int64_t orc_nnode(const orc_result *R) { return R->t.nnode; }
int64_t orc_nsink(const orc_result *R) { return R->nsink; }
int64_t orc_nlist(const orc_result *R) { return R->nsink ? R->list_off[R->nsink] : 0; }

void orc_get_perm(const orc_result *R, int64_t *perm) { memcpy(perm, R->t.perm, (size_t)R->t.n * 8); }

void orc_get_nodes(const orc_result *R, int64_t *count, int64_t *offset, int64_t *cl, int64_t *cr, float *box10) {
  for (int64_t i = 0; i < R->t.nnode; ++i) {
    const orc_node *nd = &R->t.node[i];
    count[i] = nd->count; offset[i] = nd->offset; cl[i] = nd->cl; cr[i] = nd->cr;
    for (int k = 0; k < 3; ++k) { box10[10*i + k] = nd->xmin[k]; box10[10*i + 3 + k] = nd->xmax[k]; box10[10*i + 6 + k] = nd->xc[k]; }
    box10[10*i + 9] = nd->ppm;
  }
}

/* tdr and the 12 pseudo-particle masses per node (13 floats per node; TDPTS = 12 runs). */
void orc_get_pp12(const orc_result *R, float *pp13) {
  for (int64_t i = 0; i < R->t.nnode; ++i) {
    pp13[13*i] = R->t.node[i].tdr;
    for (int q = 0; q < 12; ++q) pp13[13*i + 1 + q] = R->t.node[i].ppm12[q];
  }
}

void orc_get_lists(const orc_result *R, int64_t *sink_leaf, int64_t *list_off, int64_t *list_node, uint8_t *list_pseudo) {
  memcpy(sink_leaf, R->sink_leaf, (size_t)R->nsink * 8);
  memcpy(list_off, R->list_off, (size_t)(R->nsink + 1) * 8);
  memcpy(list_node, R->list_node, (size_t)R->list_off[R->nsink] * 8);
  memcpy(list_pseudo, R->list_pseudo, (size_t)R->list_off[R->nsink]);
}

/* f_over_r of the restated force law at n values of r2 (for force-law parity tests). */
void orc_force_law_eval(int law_kind, const float *coef, int ncoef, float rsm, float rmax, int64_t n,
                        const float *r2, float *out) {
  orc_law L; memset(&L, 0, sizeof(L));
  L.kind = law_kind;
  for (int i = 0; i < 7; ++i) L.a[i] = (coef && i < ncoef) ? coef[i] : 0.0f;
  L.rsm2 = rsm * rsm; L.r2min = 0.0f; L.r2max = rmax * rmax;
  for (int64_t i = 0; i < n; ++i) out[i] = orc_f_over_r(&L, r2[i]);
}

void orc_free(orc_result *R) {
  if (!R) return;
  free(R->t.x); free(R->t.y); free(R->t.z); free(R->t.m); free(R->t.perm); free(R->t.node);
  free(R->sink_leaf); free(R->list_off); free(R->list_node); free(R->list_pseudo);
  free(R);
}

/* FP64 direct sum over ALL particles within the cutoff for selected sinks (tree-free check, the
 * role of the direct sum in the reference's ForceTreeTest.cxx:277-302). */
void orc_direct_sum(int64_t n, const float *x, const float *y, const float *z, const float *mass,
                    int64_t nsel, const int64_t *sel, const float *coef, int ncoef, float rsm, float rmax,
                    double *ax, double *ay, double *az) {
  double a[7]; for (int k = 0; k < 7; ++k) a[k] = (coef && k < ncoef) ? (double)coef[k] : 0.0;
  float rmax2 = rmax * rmax; double rsm2 = (double)(rsm * rsm);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < nsel; ++s) {
    int64_t i = sel[s];
    double sx = 0, sy = 0, sz = 0;
    for (int64_t j = 0; j < n; ++j) {
      float fdx = x[j] - x[i], fdy = y[j] - y[i], fdz = z[j] - z[i];
      float fr2 = fdx*fdx + fdy*fdy + fdz*fdz;
      if (!(fr2 < rmax2) || !(fr2 > 0.0f)) continue;
      double dx = (double)x[j] - x[i], dy = (double)y[j] - y[i], dz = (double)z[j] - z[i];
      double r2 = dx*dx + dy*dy + dz*dz;
      double poly = 0; for (int k = 6; k >= 0; --k) poly = poly * r2 + a[k];
      double f = (double)mass[i] * mass[j] * (pow(r2 + rsm2, -1.5) - poly);
      sx += f * dx; sy += f * dy; sz += f * dz;
    }
    ax[s] = sx; ay[s] = sy; az[s] = sz;
  }
}

/* ---- PM coupling: Particles::cic / Particles::inverse_cic (src/cpu/Particles.cxx:589-643, 647-714) ---------------------
 * PINNED: bit-for-bit equal to the reference's own loops.  Particles.cxx as a whole needs MPI, Domain and the initializer
 * headers, but the three functions involved do not: oracle/build_ref.sh cuts them out of the file where it lies and compiles
 * them between a small scaffold (oracle/stubs/cic_pre.h, cic_post.cxx) into oracle/_ref/libhaccref_cic.so; the fixture
 * tests/golden/cic_ref_clustered12k.npz holds that library's outputs (tests/golden/make_golden_cic.py) and
 * tests/test_cic_cpu.py asserts equality with the fixture and, where the library is present, live.  Restated line by line
 * under C's promotion rules (`1.0` literals are double). */
static int64_t orc_array_index(int xx, int yy, int zz, const int *ng, int64_t safe) {          /* :373-395 */
  if (xx >= 0 && xx < ng[0] && yy >= 0 && yy < ng[1] && zz >= 0 && zz < ng[2]) return ((int64_t)xx * ng[1] + yy) * ng[2] + zz;
  return safe;
}
/* rho: ng0*ng1*ng2 + 1 floats (the last one is the `safe` overflow slot) */
void orc_cic(int64_t np, const float *x, const float *y, const float *z, const int *ng, float c, float *rhoArr) {
  const int64_t Ng = (int64_t)ng[0] * ng[1] * ng[2], safe = Ng;
  memset(rhoArr, 0, sizeof(float) * (size_t)(Ng + 1));
  for (int64_t nn = 0; nn < np; nn++) {
    float xx = x[nn], yy = y[nn], zz = z[nn];
    int ix = (int)floorf(xx), iy = (int)floorf(yy), iz = (int)floorf(zz);
    int ip = ix + 1, jp = iy + 1, kp = iz + 1;
    float ab = 1.0 + (ix - xx), de = 1.0 + (iy - yy), gh = 1.0 + (iz - zz);
    rhoArr[orc_array_index(ix, iy, iz, ng, safe)] += c*ab*de*gh;
    rhoArr[orc_array_index(ix, jp, iz, ng, safe)] += c*ab*(1.0-de)*gh;
    rhoArr[orc_array_index(ix, jp, kp, ng, safe)] += c*ab*(1.0-de)*(1.0-gh);
    rhoArr[orc_array_index(ix, iy, kp, ng, safe)] += c*ab*de*(1.0-gh);
    rhoArr[orc_array_index(ip, iy, kp, ng, safe)] += c*(1.0-ab)*de*(1.0-gh);
    rhoArr[orc_array_index(ip, jp, kp, ng, safe)] += c*(1.0-ab)*(1.0-de)*(1.0-gh);
    rhoArr[orc_array_index(ip, jp, iz, ng, safe)] += c*(1.0-ab)*(1.0-de)*gh;
    rhoArr[orc_array_index(ip, iy, iz, ng, safe)] += c*(1.0-ab)*de*gh;
  }
}
/* grad_phi: ng0*ng1*ng2 + 1 floats; the overflow slot is set to 0 (:677) */
void orc_inverse_cic(int64_t np, const float *x, const float *y, const float *z, float *vel, const int *ng,
                     float *grad_phi, float tau, float fscal) {
  const int64_t Ng = (int64_t)ng[0] * ng[1] * ng[2], safe = Ng;
  grad_phi[safe] = 0.0;
  for (int64_t nn = 0; nn < np; nn++) {
    float xx = x[nn], yy = y[nn], zz = z[nn];
    int ix = (int)floorf(xx), iy = (int)floorf(yy), iz = (int)floorf(zz);
    int ip = ix + 1, jp = iy + 1, kp = iz + 1;
    float ab = 1.0 + (ix - xx), de = 1.0 + (iy - yy), gh = 1.0 + (iz - zz);
    float f = 0;
    f += grad_phi[orc_array_index(ix, iy, iz, ng, safe)]*ab*de*gh;
    f += grad_phi[orc_array_index(ix, jp, iz, ng, safe)]*ab*(1.0-de)*gh;
    f += grad_phi[orc_array_index(ix, jp, kp, ng, safe)]*ab*(1.0-de)*(1.0-gh);
    f += grad_phi[orc_array_index(ix, iy, kp, ng, safe)]*ab*de*(1.0-gh);
    f += grad_phi[orc_array_index(ip, iy, kp, ng, safe)]*(1.0-ab)*de*(1.0-gh);
    f += grad_phi[orc_array_index(ip, jp, kp, ng, safe)]*(1.0-ab)*(1.0-de)*(1.0-gh);
    f += grad_phi[orc_array_index(ip, jp, iz, ng, safe)]*(1.0-ab)*(1.0-de)*gh;
    f += grad_phi[orc_array_index(ip, iy, iz, ng, safe)]*(1.0-ab)*de*gh;
    vel[nn] += f*fscal*tau;
  }
}
