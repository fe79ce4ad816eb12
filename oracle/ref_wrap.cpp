// oracle/ref_wrap.cpp - TEST INFRASTRUCTURE ONLY: C entry points into REFERENCE code that compiles from where it lies under
// /root/reference behind the Eigen shim of oracle/stub (no reference source is copied):
//   g2o/core/robust_kernel.cpp, robust_kernel_impl.cpp, robust_kernel_factory.cpp   the robust kernels themselves (object code)
//   g2o/types/slam2d/se2.h                                                         the SE2 class (inline, instantiated here)
// tests/test_oracle.py holds the oracle's restatements (and through them the device math) against these, bit for bit.
#include "g2o/core/robust_kernel_impl.h"
#include "g2o/types/slam2d/se2.h"

extern "C" {

// RobustKernel{Huber, PseudoHuber, Cauchy, Saturated, DCS}::robustify (core/robust_kernel_impl.cpp:65-126), numbered like
// B200_ROBUST_* / the oracle: 1 Huber, 2 PseudoHuber, 3 Cauchy, 4 Saturated, 5 DCS
int ref_robustify(int kind, double delta, double e2, double* rho3) {
  // one object per kernel type and thread, re-used (the oracle calls this once per edge)
  static thread_local g2o::RobustKernelHuber huber;
  static thread_local g2o::RobustKernelPseudoHuber pseudoHuber;
  static thread_local g2o::RobustKernelCauchy cauchy;
  static thread_local g2o::RobustKernelSaturated saturated;
  static thread_local g2o::RobustKernelDCS dcs;
  g2o::RobustKernel* k = 0;
  switch (kind) {
    case 1: k = &huber; break;
    case 2: k = &pseudoHuber; break;
    case 3: k = &cauchy; break;
    case 4: k = &saturated; break;
    case 5: k = &dcs; break;
    default: return -1;
  }
  k->setDelta(delta);
  Eigen::Vector3d rho(0., 0., 0.);
  k->robustify(e2, rho);
  rho3[0] = rho[0]; rho3[1] = rho[1]; rho3[2] = rho[2];
  return 0;
}

// g2o::SE2 (types/slam2d/se2.h): composition, inverse, action on a point; all as [x y theta]
void ref_se2_mul(const double* a, const double* b, double* r) {
  const g2o::SE2 c = g2o::SE2(a[0], a[1], a[2]) * g2o::SE2(b[0], b[1], b[2]);
  r[0] = c[0]; r[1] = c[1]; r[2] = c[2];
}
void ref_se2_inverse(const double* a, double* r) {
  const g2o::SE2 c = g2o::SE2(a[0], a[1], a[2]).inverse();
  r[0] = c[0]; r[1] = c[1]; r[2] = c[2];
}
void ref_se2_apply(const double* a, const double* p, double* r) {
  const Eigen::Vector2d q = g2o::SE2(a[0], a[1], a[2]) * Eigen::Vector2d(p[0], p[1]);
  r[0] = q[0]; r[1] = q[1];
}
// EdgeSE2::computeError (types/slam2d/edge_se2.h:46-52) written with the reference's SE2 class:
// delta = inverseMeasurement * (v1^-1 * v2), error = delta.toVector(); inverseMeasurement = measurement.inverse() (:55-58)
void ref_edge_se2_error(const double* v1, const double* v2, const double* z, double* e) {
  const g2o::SE2 x1(v1[0], v1[1], v1[2]), x2(v2[0], v2[1], v2[2]);
  const g2o::SE2 zi = g2o::SE2(z[0], z[1], z[2]).inverse();
  const g2o::SE2 delta = zi * (x1.inverse() * x2);
  const Eigen::Vector3d v = delta.toVector();
  e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
}
// EdgeSE2PointXY::computeError (types/slam2d/edge_se2_pointxy.h:46-51): (v1^-1 * l2) - measurement
void ref_edge_se2_xy_error(const double* v1, const double* l2, const double* z, double* e) {
  const Eigen::Vector2d q = (g2o::SE2(v1[0], v1[1], v1[2]).inverse() * Eigen::Vector2d(l2[0], l2[1])) - Eigen::Vector2d(z[0], z[1]);
  e[0] = q[0]; e[1] = q[1];
}

}  // extern "C"
