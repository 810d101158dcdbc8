// oracle/stubs/cic_post.cxx -- TEST INFRASTRUCTURE ONLY (ours): C entry points over the reference's CIC loops compiled just above.
extern "C" {
// rho: ng[0]*ng[1]*ng[2] + 1 floats (the last one is the reference's overflow slot `safe`); gpscal^3 is the deposit weight c
int ref_cic(int64_t np, float *x, float *y, float *z, const int32_t ng[3], float gpscal, float *rho) {
  for (int k = 0; k < 3; ++k) Domain::s_ng[k] = ng[k];
  Particles P;
  P.m_xArr = x; P.m_yArr = y; P.m_zArr = z; P.m_field = rho; P.m_Np_local_total = (int)np; P.m_gpscal = gpscal;
  P.m_coords_localQ = 1;
  P.cic();
  return 0;
}
// v += interpolation of grid (ng[0]*ng[1]*ng[2] + 1 floats) * fscal * tau; comp picks the array exactly like the reference
int ref_inverse_cic(int64_t np, float *x, float *y, float *z, float *vx, float *vy, float *vz, float *phi, const int32_t ng[3],
                    float *grid, float tau, float fscal, int comp) {
  for (int k = 0; k < 3; ++k) Domain::s_ng[k] = ng[k];
  Particles P;
  P.m_xArr = x; P.m_yArr = y; P.m_zArr = z; P.m_vxArr = vx; P.m_vyArr = vy; P.m_vzArr = vz; P.m_phiArr = phi;
  P.m_field = grid; P.m_Np_local_total = (int)np; P.m_coords_localQ = 1;
  P.inverse_cic(tau, fscal, comp);
  return 0;
}
}
