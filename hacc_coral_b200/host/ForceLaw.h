// host/ForceLaw.h -- header-compatible stand-in for the reference's src/halo_finder/ForceLaw.h.
//
// The reference passes the short-range force law to the tree as a virtual functor (`ForceLaw *fl`,
// RCBForceTree.h:119) whose parameters are protected members (ForceLaw.h:61-65,121-126).  A GPU kernel
// cannot call a host virtual function per pair, so this replacement keeps the same class names, constructors
// and public methods (callers such as Particles.cxx:161-183 and ForceTreeTest.cxx:85-93 compile unchanged)
// and adds ONE virtual, haccsr_describe(), through which the facade reads the law's parameters and hands
// them to libhaccsr (haccsr_set_force_law).  A user-defined subclass that does not describe itself is
// refused loudly by the facade -- there is no per-pair CPU fallback.
//
// The host-side evaluators (f_over_r, eval, fgor) are kept because callers use them directly, e.g. the
// direct-sum check of ForceTreeTest.cxx:277-296.  Formulas: ForceLaw.cxx:23-51 (grid-force fit and its
// constants), :96-141 (polynomials), :145-172 (interpolation), :175-192 (short-range law).
#ifndef FORCELAW_H
#define FORCELAW_H

#include <math.h>
#include <stdlib.h>

#include <vector>

#ifndef HACCSR_POLY_ORDER
#define HACCSR_POLY_ORDER 6   // the reference ships POLY_ORDER 6 (ForceLaw.cxx:8); 5 = the BG/Q kernel's law
#endif

// What the device needs to know about a law (mirrors the arguments of haccsr_set_force_law).
struct HaccsrLawDescription {
  int kind;                    // HACCSR_LAW_* of include/haccsr.h, or -1 = cannot be described
  int ncoef;                   // SR_POLY: coefficients in use
  float coef[8];               // SR_POLY: a[k]; SR_FIT: b c d e f g h l of the analytic fit
  float rsm;                   // Plummer softening of ForceLawSR
  float rmax;                  // cutoff radius
  std::vector<float> table;    // SR_INTERP: tabulated grid force at r2 = i * rmax^2 / (n - 1)
  HaccsrLawDescription() : kind(-1), ncoef(0), rsm(0.f), rmax(0.f) { for (int i = 0; i < 8; ++i) coef[i] = 0.f; }
};

// Analytic grid force over r (the spectrally filtered PM force a particle already feels), ForceLaw.cxx:23-51.
class FGrid {
 public:
  FGrid() : m_b(0.72f), m_c(0.01f), m_d(0.27f), m_e(0.0001f), m_f(360.0f), m_g(100.0f), m_h(0.67f), m_l(17.0f),
            m_rmax(3.116326355f) {}
  ~FGrid() {}

  float fgor(float r) {
    if (!(r >= 0.0f) || r > m_rmax) return 0.0f;
    const float f0 = m_c + 2.0 / 3.0 * m_b * m_b * m_b;
    if (r <= 0.0f) return f0;
    const float r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
    const float ch = coshf(m_b * r);
    const float num = tanhf(m_b * r) - m_b * r / ch / ch + m_c * r * r2 * (1.0 + m_d * r2) * expf(-1.0 * m_d * r2) +
                      m_e * r2 * (m_f * r2 + m_g * r4 + m_l * r6) * expf(-1.0 * m_h * r2);
    return num / (r * r * r);
  }
  // n samples of fgor at equally spaced r2 in [0, rmax^2]; caller frees with free() (ForceLaw.cxx:54-67)
  void fgor_r2_interp(int nInterp, float **r2, float **f) {
    *r2 = (float *)malloc(nInterp * sizeof(float));
    *f = (float *)malloc(nInterp * sizeof(float));
    const double step = (m_rmax * m_rmax) / (nInterp - 1.0);
    for (int i = 0; i < nInterp; ++i) {
      (*r2)[i] = i * step;
      (*f)[i] = fgor(sqrt(i * step));
    }
  }
  float rmax() { return m_rmax; }
  void haccsr_constants(float out[8]) const {
    out[0] = m_b; out[1] = m_c; out[2] = m_d; out[3] = m_e; out[4] = m_f; out[5] = m_g; out[6] = m_h; out[7] = m_l;
  }

 protected:
  float m_b, m_c, m_d, m_e, m_f, m_g, m_h, m_l, m_rmax;
};

class FGridEval {
 public:
  FGridEval() {}
  virtual ~FGridEval() {}
  virtual float eval(float r2) = 0;
  virtual float r2min() = 0;
  virtual float r2max() = 0;
  // fills kind / coefficients / table / rmax; returns false if the evaluator cannot run on the device
  virtual bool haccsr_describe(HaccsrLawDescription &) { return false; }
};

class FGridEvalFit : public FGridEval {
 public:
  explicit FGridEvalFit(FGrid *fg) : m_fg(fg) {}
  ~FGridEvalFit() {}
  float eval(float r2) { return m_fg->fgor(sqrt(r2)); }
  float r2min() { return 0.0; }
  float r2max() { return m_fg->rmax() * m_fg->rmax(); }
  bool haccsr_describe(HaccsrLawDescription &d) {
    d.kind = 1; d.ncoef = 8; m_fg->haccsr_constants(d.coef); d.rmax = m_fg->rmax();
    return true;
  }

 protected:
  FGrid *m_fg;
};

class FGridEvalPoly : public FGridEval {
 public:
  explicit FGridEvalPoly(FGrid *fg) : m_fg(fg), m_r2min(0.0f), m_r2max(fg->rmax() * fg->rmax()) {
    // ForceLaw.cxx:96-127: fits of the grid force in powers of r2
    static const float order6[7] = {0.271431f, -0.0783394f, 0.0133122f, -0.00159485f, 0.000132336f, -0.00000663394f, 0.000000147305f};
    static const float order5[7] = {0.269327f, -0.0750978f, 0.0114808f, -0.00109313f, 0.0000605491f, -0.00000147177f, 0.0f};
    static const float order4[7] = {0.263729f, -0.0686285f, 0.00882248f, -0.000592487f, 0.0000164622f, 0.0f, 0.0f};
    const float *src = (HACCSR_POLY_ORDER == 6) ? order6 : ((HACCSR_POLY_ORDER == 5) ? order5 : order4);
    for (int i = 0; i < 7; ++i) m_a[i] = src[i];
  }
  // the north-star law with explicit coefficients (a[0..n-1], n <= 7), e.g. the BG/Q table of BGQStep16.c:167
  FGridEvalPoly(FGrid *fg, const float *a, int n) : m_fg(fg), m_r2min(0.0f), m_r2max(fg->rmax() * fg->rmax()) {
    for (int i = 0; i < 7; ++i) m_a[i] = (i < n) ? a[i] : 0.0f;
  }
  ~FGridEvalPoly() {}
  float eval(float r2) {
    float v = m_a[6];
    for (int k = 5; k >= 0; --k) v = m_a[k] + r2 * v;
    return v * (r2 >= m_r2min) * (r2 <= m_r2max);
  }
  float r2min() { return 0.0; }
  float r2max() { return m_fg->rmax() * m_fg->rmax(); }
  bool haccsr_describe(HaccsrLawDescription &d) {
    d.kind = 0; d.ncoef = 7;
    while (d.ncoef > 1 && m_a[d.ncoef - 1] == 0.0f) d.ncoef--;
    for (int i = 0; i < 7; ++i) d.coef[i] = m_a[i];
    d.rmax = m_fg->rmax();
    return true;
  }

 protected:
  FGrid *m_fg;
  float m_r2min, m_r2max;
  float m_a[7];
};

class FGridEvalInterp : public FGridEval {
 public:
  FGridEvalInterp(FGrid *fg, int nInterp) : m_nInterp(nInterp), m_rmax(fg->rmax()) {
    fg->fgor_r2_interp(m_nInterp, &m_r2, &m_f);
    m_r2min = m_r2[0];
    m_r2max = m_r2[m_nInterp - 1];
    m_dr2 = (m_r2max - m_r2min) / (m_nInterp - 1.0);
    m_oodr2 = 1.0 / m_dr2;
  }
  ~FGridEvalInterp() { free(m_r2); free(m_f); }
  float eval(float r2) {   // ForceLaw.cxx:164-172: linear in r2 between samples, zero outside (r2min, r2max)
    const int in = (r2 > m_r2min) * (r2 < m_r2max);
    const int i = int((r2 - m_r2min) * m_oodr2) * in;
    return (1.0f * in) * (m_f[i] + (r2 - m_r2[i]) * m_oodr2 * (m_f[i + 1] - m_f[i]));
  }
  float r2min() { return m_r2min; }
  float r2max() { return m_r2max; }
  int nInterp() { return m_nInterp; }
  float *r2() { return m_r2; }
  float *f() { return m_f; }
  bool haccsr_describe(HaccsrLawDescription &d) {
    d.kind = 2; d.rmax = m_rmax;
    d.table.assign(m_f, m_f + m_nInterp);
    return true;
  }

 protected:
  float *m_r2;
  float *m_f;
  float m_r2min, m_r2max, m_dr2, m_oodr2;
  int m_nInterp;
  float m_rmax;
};

class ForceLaw {
 public:
  ForceLaw() {}
  virtual ~ForceLaw() {}
  virtual float f_over_r(float r2) = 0;
  virtual bool haccsr_describe(HaccsrLawDescription &) { return false; }
};

class ForceLawNewton : public ForceLaw {
 public:
  ~ForceLawNewton() {}
  float f_over_r(float r2) { return 1.0 / r2 / sqrt(r2); }
  bool haccsr_describe(HaccsrLawDescription &d) { d.kind = 3; d.rsm = 0.f; return true; }
};

// (r2 + rsm^2)^-3/2 minus the grid force, inside [r2min, r2max] (ForceLaw.cxx:175-192)
class ForceLawSR : public ForceLaw {
 public:
  ForceLawSR(FGridEval *fgore, float rsm)
      : m_rsm(rsm), m_rsm2(rsm * rsm), m_r2min(fgore->r2min()), m_r2max(fgore->r2max()), m_fgore(fgore) {}
  ~ForceLawSR() {}
  float f_over_r(float r2) {
    float v = powf(r2 + m_rsm2, -1.5) - m_fgore->eval(r2);
    v *= (r2 >= m_r2min) * (r2 <= m_r2max);
    return v;
  }
  bool haccsr_describe(HaccsrLawDescription &d) {
    if (!m_fgore->haccsr_describe(d)) return false;
    d.rsm = m_rsm;
    return true;
  }

 protected:
  float m_rsm, m_rsm2, m_r2min, m_r2max;
  FGridEval *m_fgore;
};

#endif  // FORCELAW_H
