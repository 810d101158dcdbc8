// tests/cxx/facade_test.cxx -- the reference's ForceTreeTest scenario as a pass/fail program, compiled against
// the facade headers (hacc_coral_b200/host/RCBForceTree.h, ForceLaw.h) exactly the way the reference's own test
// compiles against its headers (reference src/halo_finder/ForceTreeTest.cxx:77-105,188-237,277-302).
//
// A random sphere of nSphere unit-mass particles plus one test particle; the tree constructor kicks every
// particle; the kick of the test particle is compared with the direct O(N) sum evaluated on the host with the
// SAME ForceLaw object (fl->f_over_r), for the force laws and tree parameters given on the command line:
//   facade_test <law: fit|poly|interp|newton> <theta> <ppn> <nSphere> <trials> <tolerance> [quad]
// (`quad` builds RCBQuadrupoleForceTree, the reference's -S, instead of RCBMonopoleForceTree)
// Exit status 0 = every trial within tolerance.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "RCBForceTree.h"

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s law theta ppn nSphere trials tol\n", argv[0]); return 2; }
  const char *law = argv[1];
  const float theta = atof(argv[2]);
  const int ppn = atoi(argv[3]), nSphere = atoi(argv[4]), trials = atoi(argv[5]);
  const double tol = atof(argv[6]);
  const bool quad = argc > 7 && !strcmp(argv[7], "quad");
  const float L = 20.0f, rSphere = 4.0f, rsm = 0.1f;      // ForceTreeTest.cxx:77 uses rsm = 0.1

  FGrid fg;
  FGridEval *ev = 0;
  ForceLaw *fl = 0;
  if (!strcmp(law, "fit")) ev = new FGridEvalFit(&fg);
  else if (!strcmp(law, "poly")) ev = new FGridEvalPoly(&fg);
  else if (!strcmp(law, "interp")) ev = new FGridEvalInterp(&fg, 1024);
  if (ev) fl = new ForceLawSR(ev, rsm);
  else if (!strcmp(law, "newton")) fl = new ForceLawNewton();
  else { fprintf(stderr, "unknown law %s\n", law); return 2; }

  const int Np = nSphere + 1;
  POSVEL_T *x = new POSVEL_T[Np], *y = new POSVEL_T[Np], *z = new POSVEL_T[Np];
  POSVEL_T *vx = new POSVEL_T[Np], *vy = new POSVEL_T[Np], *vz = new POSVEL_T[Np];
  POSVEL_T *mass = new POSVEL_T[Np], *phi = new POSVEL_T[Np];
  ID_T *id = new ID_T[Np];
  MASK_T *mask = new MASK_T[Np];
  srand48(20240607);
  int bad = 0;
  double worst = 0.0;
  for (int t = 0; t < trials; ++t) {
    // sphere centre well inside the box, test particle anywhere in the box
    float c[3], p[3];
    for (int k = 0; k < 3; ++k) { c[k] = rSphere + (L - 2 * rSphere) * drand48(); p[k] = L * drand48(); }
    int n = 0;
    while (n < nSphere) {
      float u = 2 * drand48() - 1, v = 2 * drand48() - 1, w = 2 * drand48() - 1;
      if (u * u + v * v + w * w > 1.0f) continue;
      x[n] = c[0] + rSphere * u; y[n] = c[1] + rSphere * v; z[n] = c[2] + rSphere * w; ++n;
    }
    x[n] = p[0]; y[n] = p[1]; z[n] = p[2];
    for (int i = 0; i < Np; ++i) { vx[i] = vy[i] = vz[i] = 0.f; mass[i] = 1.f; phi[i] = 0.f; id[i] = i; mask[i] = 0; }
    float zero[3] = {0.f, 0.f, 0.f}, top[3] = {L, L, L};

    if (quad) {
      RCBQuadrupoleForceTree *sft = new RCBQuadrupoleForceTree(zero, top, zero, top, Np, x, y, z, vx, vy, vz, mass, phi, id,
                                                               mask, 1.0, fg.rmax(), rsm, theta, ppn, 2, 128, fl, 1.0f);
      delete sft;
    } else {
      RCBMonopoleForceTree *sft = new RCBMonopoleForceTree(zero, top, zero, top, Np, x, y, z, vx, vy, vz, mass, phi, id, mask,
                                                           1.0, fg.rmax(), rsm, theta, ppn, 2, 128, fl, 1.0f);
      delete sft;
    }

    // the tree reordered the arrays: find the test particle by id (ForceTreeTest.cxx:268-272 searches by x)
    int pidx = 0;
    while (id[pidx] != nSphere) ++pidx;
    double d[3] = {0, 0, 0};
    const float rmax2 = fg.rmax() * fg.rmax();
    for (int i = 0; i < Np; ++i) {
      if (i == pidx) continue;
      float dx = x[i] - x[pidx], dy = y[i] - y[pidx], dz = z[i] - z[pidx];
      float r2 = dx * dx + dy * dy + dz * dz;
      if (!(r2 < rmax2)) continue;                       // the kernel's cutoff (RCBForceTree.cxx:612)
      double f = fl->f_over_r(r2);
      d[0] += dx * f; d[1] += dy * f; d[2] += dz * f;
    }
    const double fd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double e = sqrt((vx[pidx] - d[0]) * (vx[pidx] - d[0]) + (vy[pidx] - d[1]) * (vy[pidx] - d[1]) +
                          (vz[pidx] - d[2]) * (vz[pidx] - d[2]));
    // error relative to the direct force, floored for test particles that feel almost nothing
    const double rel = e / (fd > 1e-3 ? fd : 1e-3);
    if (rel > worst) worst = rel;
    if (!(rel <= tol)) { ++bad; printf("trial %d: tree (%g %g %g) direct (%g %g %g) rel %g\n", t, vx[pidx], vy[pidx], vz[pidx], d[0], d[1], d[2], rel); }
  }
  printf("facade_test %s law=%s theta=%g ppn=%d nSphere=%d trials=%d: worst relative error %.3e (tolerance %.1e) -> %s\n", quad ? "quadrupole" : "monopole", law,
         theta, ppn, nSphere, trials, worst, tol, bad ? "FAIL" : "ok");
  haccsr_facade_release();
  delete fl; delete ev;
  return bad ? 1 : 0;
}
