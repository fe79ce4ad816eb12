// geometry_host.cpp - TEST-ONLY host build of the device math in openslam_g2o_b200/csrc/geometry.cuh so that the
// exact formulas the kernels run can be compared with the oracle on a box without a GPU.  Never shipped.
#include <cstring>

#include "geometry.cuh"

using namespace g2o_b200::geo;

extern "C" {
void gh_se2(const double* xi, const double* xj, const double* z, double* e, double* A, double* B) {
  SE2 zi = se2_inv(SE2{z[0], z[1], z[2]});
  SE2 a{xi[0], xi[1], xi[2]}, b{xj[0], xj[1], xj[2]};
  se2_error(a, b, zi, e);
  se2_jacobians(a, b, zi, A, B);
}
static Iso to_iso(const double* p) { Iso r; memcpy(r.R, p, 72); memcpy(r.t, p + 9, 24); return r; }
void gh_se3(const double* Xi, const double* Xj, const double* Z, double* e, double* Ji, double* Jj) {
  Iso zi = iso_inverse(to_iso(Z));
  double e2[6];
  se3_error(to_iso(Xi), to_iso(Xj), zi, e);
  se3_jacobians(to_iso(Xi), to_iso(Xj), zi, Ji, Jj, e2);
  // (Zi Xi^-1) Xj vs Zi (Xi^-1 Xj): the two association orders the reference itself uses in computeError /
  // linearizeOplus; they must agree to rounding
  for (int i = 0; i < 6; ++i) if (fabs(e[i] - e2[i]) > 1e-11 * (1.0 + fabs(e[i]))) e[i] = 1e300;
}
void gh_p2mc(const double* cam, const double* X, const double* z, double* e, double* Jp, double* Jc) {
  double der[16];
  cam_derive(cam, der);
  p2mc_error(der, X, z, e);
  p2mc_jacobians(der, cam, X, Jp, Jc);
}
// est: t3 q4 f f cx cy b (world -> camera)
void gh_xyz2uv(const double* est, const double* X, const double* z, double* e, double* Jp, double* Jc) {
  double der[16];
  ba_derive<1>(est, der);
  ba_error<1>(der, X, z, e);
  ba_jacobians<1>(der, est, X, Jp, Jc);
}
void gh_oplus(int kind, double* est, const double* u) {
  if (kind == 0) se2_oplus(est, u);
  else if (kind == 1) se3_oplus(est, u, false);
  else if (kind == 2) cam_oplus(est, u);
  else if (kind == 4) expmap_oplus(est, u);
  else { est[0] += u[0]; est[1] += u[1]; est[2] += u[2]; }
}
void gh_se2_xy(const double* x, const double* l, const double* z, double* e, double* A, double* B) {
  SE2 a{x[0], x[1], x[2]};
  se2_xy_error(a, l, z, e);
  se2_xy_jacobians(a, l, A, B);
}
void gh_se3_xyz(const double* X, const double* offset, const double* l, const double* z, double* e, double* A, double* B) {
  se3_xyz_error(to_iso(X), to_iso(offset), l, z, e);
  se3_xyz_jacobians(to_iso(X), to_iso(offset), l, A, B);
}
void gh_inverse3(const double* m, double* r) { inverse3(m, r); }
void gh_dq_dR(const double* R, double* dq) { dq_dR(R, dq); }
}
