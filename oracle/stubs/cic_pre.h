// oracle/stubs/cic_pre.h -- TEST INFRASTRUCTURE ONLY (ours, not reference text).
// Scaffolding in front of which oracle/build_ref.sh pipes three functions of the reference's src/cpu/Particles.cxx, taken
// from the file where it lies (Particles::array_index :371-395, Particles::cic :589-643, Particles::inverse_cic :647-714),
// so that the reference's own CIC loops compile without MPI: just enough of class Particles and of Domain for those
// bodies, with the reference's build types (-DPOSVEL_32 -DGRID_32: POSVEL_T = GRID_T = float, FLOOR = floorf,
// src/simulation/mc3types.h:11).
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DIMENSION 3
#define FLOOR floorf
typedef float POSVEL_T;
typedef float GRID_T;

struct Domain {
  static int s_ng[3];
  static void ng_local_total(int ng[3]) { for (int k = 0; k < 3; ++k) ng[k] = s_ng[k]; }
  static int Ng_local_total() { return s_ng[0] * s_ng[1] * s_ng[2]; }
};
int Domain::s_ng[3] = {0, 0, 0};

class Particles {
 public:
  int array_index(int xx, int yy, int zz, int ng[DIMENSION], int lo[DIMENSION], int hi[DIMENSION], int safe);
  void cic();
  void inverse_cic(float tau, float fscal, int comp);
  POSVEL_T *m_xArr, *m_yArr, *m_zArr, *m_vxArr, *m_vyArr, *m_vzArr, *m_phiArr;
  GRID_T *m_field;
  int m_Np_local_total;
  float m_gpscal;
  int m_coords_localQ;
};
