// oracle.cpp - CPU ORACLE for the g2o LM/GN hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// A from-scratch, Eigen-free restatement of the reference's CPU path, in the reference's operation
// order (edges in internalId order, upper-triangular block Hessian, column-major blocks), linked
// against the reference's own vendored CSparse (oracle/_ref/libg2o_csparse_ref.so) for cs_amd,
// cs_symperm/etree/post/counts, and csparse_extension::cs_cholsolsymb.
//
// Every function cites the reference file:line it restates.  Paths are relative to
// /root/reference/g2o unless they start with EXTERNAL/.
//
// Parity pinning status: pinned by BASELINE.md section 2 known answers (block-AMD permutation hashes and
// nnz(L) for the four in-tree datasets, tests/test_oracle.py); CHOLMOD flavour: "parity unpinned".
#include "oracle.h"

#include <Eigen/Core>  // oracle/stub/Eigen/Core (shim, see there)
namespace g2o { namespace internal {   // g2o/types/slam3d/dquat2mat.h:9 - the reference's own object code (oracle/_ref)
void compute_dq_dR(Eigen::Matrix<double, 3, 9>& dq_dR, const double& r11, const double& r21, const double& r31, const double& r12,
                   const double& r22, const double& r32, const double& r13, const double& r23, const double& r33);
} }

#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <time.h>
#include <tr1/unordered_map>
#include <functional>
#include <vector>

#ifndef NCOMPLEX
#define NCOMPLEX
#endif
#include <cs.h>  // reference header, found via -I$(REF)/EXTERNAL/csparse at oracle build time

// solvers/csparse/csparse_helper.h:41-42 (compiled from the reference into libg2o_csparse_ref.so)
// oracle/ref_wrap.cpp -> oracle/_ref/libg2o_ref_wrap.so: the reference's robust kernels (core/robust_kernel_impl.cpp) and
// SE2 class (types/slam2d/se2.h), compiled from where they lie behind the Eigen shim of oracle/stub
extern "C" {
int ref_robustify(int kind, double delta, double e2, double* rho3);
void ref_edge_se2_error(const double* v1, const double* v2, const double* z, double* e);
void ref_edge_se2_xy_error(const double* v1, const double* l2, const double* z, double* e);
}
namespace g2o { namespace csparse_extension {
int cs_cholsolsymb(const cs* A, double* b, const css* S, double* workspace, int* work);
csn* cs_chol_workspace(const cs* A, const css* S, int* cin, double* xin);  // csparse_helper.h:37
} }

namespace {

// stuff/timeutil.h:107 get_monotonic_time
inline double now() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// stuff/misc.h:94-106
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}

// ---------------------------------------------------------------------------------------------
// tiny column-major dense helpers (stand-ins for the Eigen fixed-size expressions)
// ---------------------------------------------------------------------------------------------
// C(RxC) = A(RxK) * B(KxC)
template <int R, int K, int C>
inline void mm(const double* A, const double* B, double* Cm) {
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[r + k * R] * B[k + c * K];
      Cm[r + c * R] = s;
    }
}
// C(RxC) = A^T (A is KxR) * B (KxC)
template <int R, int K, int C>
inline void mtm(const double* A, const double* B, double* Cm) {
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[k + r * K] * B[k + c * K];
      Cm[r + c * R] = s;
    }
}
// C(RxC) = A (RxK) * B^T (B is CxK)
template <int R, int K, int C>
inline void mmt(const double* A, const double* B, double* Cm) {
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[r + k * R] * B[c + k * C];
      Cm[r + c * R] = s;
    }
}

// Eigen Quaterniond::toRotationMatrix (Eigen/src/Geometry/Quaternion.h); q = (x,y,z,w); R col-major
inline void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
  R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
  R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Eigen Quaterniond(Matrix3d) (restated in-tree at types/slam3d/test_mat2quat_jacobian.cpp:42-83)
inline void R_to_quat(const double* R, double* q) {
  auto m = [&](int r, int c) { return R[r + 3 * c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}
inline void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

// Isometry3d stored as [R col-major 9 | t 3]
struct Iso { double R[9]; double t[3]; };
inline Iso iso_inverse(const Iso& a) {  // Eigen Transform::inverse(Isometry): [R^T | -R^T t]
  Iso r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.R[i + 3 * j] = a.R[j + 3 * i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += -r.R[i + 3 * k] * a.t[k];
    r.t[i] = s;
  }
  return r;
}
inline Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  mm<3, 3, 3>(a.R, b.R, r.R);
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += a.R[i + 3 * k] * b.t[k];
    r.t[i] = s + a.t[i];
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// graph objects
// ---------------------------------------------------------------------------------------------
constexpr int EST_MAX = 64;  // CAM: t3 q4 K5 | w2n 12 | w2i 12 | dRdx 9 dRdy 9 dRdz 9 = 63

struct Edge;
struct Vertex {
  int kind = 0, id = 0, dim = 0;
  bool fixed = false, marginalized = false;
  int hessianIndex = -1, colInHessian = -1;
  int numOplusCalls = 0;
  double est[EST_MAX];
  double b[6];
  double* H = nullptr;  // mapped diagonal block (dim x dim, col-major), base_vertex.h mapHessianMemory
  std::vector<std::array<double, EST_MAX>> backup;
  std::vector<Edge*> edges;
};

struct Edge {
  int kind = 0, D = 0, internalId = 0;
  Vertex* v[2] = {nullptr, nullptr};
  double meas[12];     // SE2 [x y th]; SE3 Iso of Z; P2MC [u v]
  double invMeas[12];  // cached inverse measurement (edge_se2.h:55-58, edge_se3 setMeasurement)
  double info[36];     // D x D col-major
  double err[6];
  double Ji[36], Jj[36];  // D x Di, D x Dj col-major
  double* H = nullptr;    // mapped off-diagonal block
  bool transposed = false;  // _hessianRowMajor (base_binary_edge.hpp:207-218)
  int rkKind = 0;           // robustKernel(): 0 none, 1 Huber, 2 PseudoHuber, 3 Cauchy, 4 Saturated, 5 DCS
  double rkDelta = 1.0;     // RobustKernel::_delta
  int paramId = -1;         // EdgeProjectXYZ2UV: id of its CameraParameters (types_six_dof_expmap.cpp:241-256)
  double camPar[4] = {1, 0, 0, 0.5};  // focal_length, principle_point x y, baseline of that parameter
  double offset[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};  // EdgeSE3PointXYZ: ParameterSE3Offset::offset() as Iso
};

inline int vertex_dim(int kind) { return kind == ORC_VERTEX_SE2 ? 3 : kind == ORC_VERTEX_XYZ ? 3 : kind == ORC_VERTEX_XY ? 2 : 6; }
inline int edge_dim(int kind) {  // P2MC, XYZ2UV, SE2_XY: 2
  return kind == ORC_EDGE_SE2 ? 3 : kind == ORC_EDGE_SE3 ? 6 : kind == ORC_EDGE_SE3_XYZ ? 3 : 2;
}

// ---- SE2 (types/slam2d/se2.h:41-119) ----
struct SE2 { double x, y, th; };
inline SE2 se2_mul(const SE2& a, const SE2& b) {  // se2.h:66-78
  SE2 r;
  double c = cos(a.th), s = sin(a.th);
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}
inline SE2 se2_inv(const SE2& a) {  // se2.h:86-96
  SE2 r;
  r.th = normalize_theta(-a.th);
  double c = cos(r.th), s = sin(r.th);
  double tx = a.x * -1., ty = a.y * -1.;
  r.x = c * tx - s * ty;
  r.y = s * tx + c * ty;
  return r;
}

// ---- SBACam derived quantities (types/sba/sbacam.h:120-181) ----
// est: t[0..3) q[3..7) K[7..12)=fx fy cx cy baseline | w2n[12..24) | w2i[24..36) | dRdx[36..45) dRdy dRdz
inline void cam_refresh(double* est) {
  double R[9];
  quat_to_R(est + 3, R);
  double* w2n = est + 12;  // 3x4 col-major
  // transformW2F (sbacam.h:120-130): m.block<3,3> = R^T ; m.col(3) = -m * [t;1] with col(3) zero
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) w2n[r + 3 * c] = R[c + 3 * r];
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += w2n[r + 3 * k] * est[k];
    s += 0.0 * 1.0;
    w2n[r + 9] = -s;
  }
  // setProjection (sbacam.h:159): w2i = Kcam * w2n
  const double fx = est[7], fy = est[8], cx = est[9], cy = est[10];
  double* w2i = est + 24;
  for (int c = 0; c < 4; ++c) {
    w2i[0 + 3 * c] = fx * w2n[0 + 3 * c] + 0.0 * w2n[1 + 3 * c] + cx * w2n[2 + 3 * c];
    w2i[1 + 3 * c] = 0.0 * w2n[0 + 3 * c] + fy * w2n[1 + 3 * c] + cy * w2n[2 + 3 * c];
    w2i[2 + 3 * c] = 0.0 * w2n[0 + 3 * c] + 0.0 * w2n[1 + 3 * c] + 1.0 * w2n[2 + 3 * c];
  }
  // setDr (sbacam.h:162-181): dRdx = dRidx * w2n.block<3,3>(0,0) etc.
  static const double dRidx[9] = {0, 0, 0, 0, 0, -2, 0, 2, 0};   // col-major of [[0,0,0],[0,0,2],[0,-2,0]]
  static const double dRidy[9] = {0, 0, 2, 0, 0, 0, -2, 0, 0};   // [[0,0,-2],[0,0,0],[2,0,0]]
  static const double dRidz[9] = {0, -2, 0, 2, 0, 0, 0, 0, 0};   // [[0,2,0],[-2,0,0],[0,0,0]]
  mm<3, 3, 3>(dRidx, w2n, est + 36);
  mm<3, 3, 3>(dRidy, w2n, est + 45);
  mm<3, 3, 3>(dRidz, w2n, est + 54);
}
// SE3Quat::normalizeRotation (types/slam3d/se3quat.h:280-285)
inline void se3quat_normalize_rotation(double* q) {
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] *= -1;
  quat_normalize(q);
}


// ---- SE3Quat (types/slam3d/se3quat.h:40-300), stored as [t3 | q(xyzw)4] ----
// Eigen Quaterniond * Vector3d (QuaternionBase::_transformVector): v + w*uv + qv x uv, uv = 2 (qv x v)
inline void quat_rotate(const double* q, const double* v, double* r) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int i = 0; i < 3; ++i) uv[i] += uv[i];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) r[i] = v[i] + q[3] * uv[i] + c[i];
}
// Eigen quaternion product a*b
inline void quat_mul(const double* a, const double* b, double* r) {
  r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
// SE3Quat::inverse (se3quat.h:125-130): r = conj(r), t = r * (-t); no normalisation
inline void se3quat_inverse(const double* a, double* r) {
  double q[4] = {-a[3], -a[4], -a[5], a[6]};
  double mt[3] = {a[0] * -1., a[1] * -1., a[2] * -1.};
  quat_rotate(q, mt, r);
  for (int i = 0; i < 4; ++i) r[3 + i] = q[i];
}
// SE3Quat::operator* (se3quat.h:103-109)
inline void se3quat_mul(const double* a, const double* b, double* r) {
  double rt[3];
  quat_rotate(a + 3, b, rt);
  for (int i = 0; i < 3; ++i) r[i] = a[i] + rt[i];
  quat_mul(a + 3, b + 3, r + 3);
  se3quat_normalize_rotation(r + 3);
}
// SE3Quat::exp (se3quat.h:216-252): update = [omega ; upsilon]
inline void se3quat_exp(const double* u, double* r) {
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  // skew (se3_ops.hpp:27-38), col-major
  const double Om[9] = {0, om[2], -om[1], -om[2], 0, om[0], om[1], -om[0], 0};
  double Om2[9], R[9], V[9];
  mm<3, 3, 3>(Om, Om, Om2);
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
    for (int i = 0; i < 9; ++i) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  R_to_quat(R, r + 3);
  mm<3, 3, 1>(V, up, r);
  se3quat_normalize_rotation(r + 3);  // SE3Quat(Quaterniond, Vector3d) ctor
}

// ---------------------------------------------------------------------------------------------
// per-type error / Jacobian / oplus
// ---------------------------------------------------------------------------------------------
// types/slam3d/isometry3d_mappings.cpp:38-44, 77-83, 93-99  toVectorMQT
inline void to_vector_mqt(const Iso& d, double* v) {
  double q[4];
  R_to_quat(d.R, q);
  quat_normalize(q);
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] *= -1;
  v[0] = d.t[0]; v[1] = d.t[1]; v[2] = d.t[2];
  v[3] = q[0]; v[4] = q[1]; v[5] = q[2];
}

void compute_error(Edge* e) {
  Vertex* v0 = e->v[0];
  Vertex* v1 = e->v[1];
  switch (e->kind) {
    case ORC_EDGE_SE2: {  // types/slam2d/edge_se2.h:46-52, evaluated by the reference's own SE2 class (oracle/ref_wrap.cpp)
      ref_edge_se2_error(v0->est, v1->est, e->meas, e->err);
      break;
    }
    case ORC_EDGE_SE3: {  // types/slam3d/edge_se3.cpp:48-53
      Iso Xi, Xj, Zi;
      memcpy(&Xi, v0->est, sizeof(Iso)); memcpy(&Xj, v1->est, sizeof(Iso));
      memcpy(&Zi, e->invMeas, sizeof(Iso));
      Iso d = iso_mul(iso_mul(Zi, iso_inverse(Xi)), Xj);
      to_vector_mqt(d, e->err);
      break;
    }
    case ORC_EDGE_P2MC: {  // types/sba/types_sba.h:170-192
      const double* pt = v0->est;
      const double* w2i = v1->est + 24;
      double p[3];
      for (int r = 0; r < 3; ++r)
        p[r] = w2i[r] * pt[0] + w2i[r + 3] * pt[1] + w2i[r + 6] * pt[2] + w2i[r + 9] * 1.0;
      e->err[0] = p[0] / p[2] - e->meas[0];
      e->err[1] = p[1] / p[2] - e->meas[1];
      break;
    }
    case ORC_EDGE_XYZ2UV: {  // types/sba/types_six_dof_expmap.h:143-150, cam_map types_six_dof_expmap.cpp:65-71
      double p[3];
      quat_rotate(v1->est + 3, v0->est, p);  // SE3Quat::map: _r*xyz + _t
      for (int i = 0; i < 3; ++i) p[i] += v1->est[i];
      const double proj[2] = {p[0] / p[2], p[1] / p[2]};
      e->err[0] = e->meas[0] - (proj[0] * e->camPar[0] + e->camPar[1]);
      e->err[1] = e->meas[1] - (proj[1] * e->camPar[0] + e->camPar[2]);
      break;
    }
    case ORC_EDGE_SE2_XY: {  // types/slam2d/edge_se2_pointxy.h:46-51: (v1^-1 * l2) - z, by the reference's own SE2 class
      ref_edge_se2_xy_error(v0->est, v1->est, e->meas, e->err);
      break;
    }
    case ORC_EDGE_SE3_XYZ: {  // types/slam3d/edge_se3_pointxyz.cpp:98-108; cache: parameter_se3_offset.cpp:75-80
      Iso X, O; memcpy(&X, v0->est, sizeof(Iso)); memcpy(&O, e->offset, sizeof(Iso));
      Iso w2n = iso_inverse(iso_mul(X, O));
      for (int r = 0; r < 3; ++r)
        e->err[r] = (w2n.R[r] * v1->est[0] + w2n.R[r + 3] * v1->est[1] + w2n.R[r + 6] * v1->est[2] + w2n.t[r]) - e->meas[r];
      break;
    }
  }
}

// types/slam3d/dquat2mat.cpp:9-59 + dquat2mat_maxima_generated.cpp:1-165.
// dq (3x9 col-major), columns ordered r00 r10 r20 r01 r11 r21 r02 r12 r22.
// The oracle calls the REFERENCE's own function (compiled unmodified into oracle/_ref/libg2o_slam3d_ref.so behind the
// Eigen::Matrix shim of oracle/stub/Eigen/Core); the restatement below stays as the cross-check of tests/test_oracle.py
// (bit-equal on all four branches) and as the formula sheet of the device code (csrc/geometry.cuh).
void compute_dq_dR(double* dq, const double* R) {
  Eigen::Matrix<double, 3, 9> M;
  g2o::internal::compute_dq_dR(M, R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]);
  for (int i = 0; i < 27; ++i) dq[i] = M.v[i];
}
void compute_dq_dR_restated(double* dq, const double* R) {
  const double r00 = R[0], r10 = R[1], r20 = R[2], r01 = R[3], r11 = R[4], r21 = R[5], r02 = R[6],
               r12 = R[7], r22 = R[8];
  for (int i = 0; i < 27; ++i) dq[i] = 0;
  auto D = [&](int r, int c) -> double& { return dq[r + 3 * c]; };
  double S, qw;
  double tr = r00 + r11 + r22;
  if (tr > 0) {
    S = sqrt(tr + 1.0) * 2;
    qw = 0.25 * S;
    S *= .25;
    double a1 = 1 / pow(S, 3), a2 = -0.03125 * (r21 - r12) * a1, a3 = 1 / S, a4 = 0.25 * a3,
           a5 = -0.25 * a3, a6 = 0.03125 * (r20 - r02) * a1, a7 = -0.03125 * (r10 - r01) * a1;
    D(0, 0) = a2; D(0, 4) = a2; D(0, 5) = a4; D(0, 7) = a5; D(0, 8) = a2;
    D(1, 0) = a6; D(1, 2) = a5; D(1, 4) = a6; D(1, 6) = a4; D(1, 8) = a6;
    D(2, 0) = a7; D(2, 1) = a4; D(2, 3) = a5; D(2, 4) = a7; D(2, 8) = a7;
  } else if ((r00 > r11) & (r00 > r22)) {
    S = sqrt(1.0 + r00 - r11 - r22) * 2;
    qw = (r21 - r12) / S;
    S *= .25;
    double a1 = 1 / S, a2 = -0.125 * a1, a3 = 1 / pow(S, 3), a4 = r10 + r01, a5 = 0.25 * a1,
           a6 = 0.03125 * a3 * a4, a7 = r20 + r02, a8 = 0.03125 * a3 * a7;
    D(0, 0) = 0.125 * a1; D(0, 4) = a2; D(0, 8) = a2;
    D(1, 0) = -0.03125 * a3 * a4; D(1, 1) = a5; D(1, 3) = a5; D(1, 4) = a6; D(1, 8) = a6;
    D(2, 0) = -0.03125 * a3 * a7; D(2, 2) = a5; D(2, 4) = a8; D(2, 6) = a5; D(2, 8) = a8;
  } else if (r11 > r22) {
    S = sqrt(1.0 + r11 - r00 - r22) * 2;
    qw = (r02 - r20) / S;
    S *= .25;
    double a1 = 1 / pow(S, 3), a2 = r10 + r01, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4,
           a6 = -0.125 * a4, a7 = r21 + r12, a8 = 0.03125 * a1 * a7;
    D(0, 0) = a3; D(0, 1) = a5; D(0, 3) = a5; D(0, 4) = -0.03125 * a1 * a2; D(0, 8) = a3;
    D(1, 0) = a6; D(1, 4) = 0.125 * a4; D(1, 8) = a6;
    D(2, 0) = a8; D(2, 4) = -0.03125 * a1 * a7; D(2, 5) = a5; D(2, 7) = a5; D(2, 8) = a8;
  } else {
    S = sqrt(1.0 + r22 - r00 - r11) * 2;
    qw = (r10 - r01) / S;
    S *= .25;
    double a1 = 1 / pow(S, 3), a2 = r20 + r02, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4,
           a6 = r21 + r12, a7 = 0.03125 * a1 * a6, a8 = -0.125 * a4;
    D(0, 0) = a3; D(0, 2) = a5; D(0, 4) = a3; D(0, 6) = a5; D(0, 8) = -0.03125 * a1 * a2;
    D(1, 0) = a7; D(1, 4) = a7; D(1, 5) = a5; D(1, 7) = a5; D(1, 8) = -0.03125 * a1 * a6;
    D(2, 0) = a8; D(2, 4) = a8; D(2, 8) = 0.125 * a4;
  }
  if (qw <= 0) for (int i = 0; i < 27; ++i) dq[i] *= -1;
}

void linearize(Edge* e) {
  Vertex* v0 = e->v[0];
  Vertex* v1 = e->v[1];
  switch (e->kind) {
    case ORC_EDGE_SE2: {  // types/slam2d/edge_se2.cpp:76-99
      double thetai = v0->est[2];
      double dtx = v1->est[0] - v0->est[0], dty = v1->est[1] - v0->est[1];
      double si = sin(thetai), ci = cos(thetai);
      double A0[9], B0[9];
      auto A = [&](int r, int c) -> double& { return A0[r + 3 * c]; };
      auto B = [&](int r, int c) -> double& { return B0[r + 3 * c]; };
      A(0, 0) = -ci; A(0, 1) = -si; A(0, 2) = -si * dtx + ci * dty;
      A(1, 0) = si;  A(1, 1) = -ci; A(1, 2) = -ci * dtx - si * dty;
      A(2, 0) = 0;   A(2, 1) = 0;   A(2, 2) = -1;
      B(0, 0) = ci;  B(0, 1) = si;  B(0, 2) = 0;
      B(1, 0) = -si; B(1, 1) = ci;  B(1, 2) = 0;
      B(2, 0) = 0;   B(2, 1) = 0;   B(2, 2) = 1;
      double z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      double c = cos(e->invMeas[2]), s = sin(e->invMeas[2]);
      z[0] = c; z[3] = -s; z[1] = s; z[4] = c; z[8] = 1.;
      mm<3, 3, 3>(z, A0, e->Ji);
      mm<3, 3, 3>(z, B0, e->Jj);
      break;
    }
    case ORC_EDGE_SE3: {  // types/slam3d/edge_se3.cpp:63-75 -> isometry3d_gradients.h:194-265
      Iso Xi, Xj, Z;
      memcpy(&Xi, v0->est, sizeof(Iso)); memcpy(&Xj, v1->est, sizeof(Iso));
      memcpy(&Z, e->meas, sizeof(Iso));
      const Iso A = iso_inverse(Z);
      const Iso B = iso_mul(iso_inverse(Xi), Xj);
      const Iso E = iso_mul(A, B);
      const double *Re = E.R, *Ra = A.R, *Rb = B.R, *tb = B.t;
      double dq[27];
      compute_dq_dR(dq, Re);
      double* Ji = e->Ji; double* Jj = e->Jj;
      for (int i = 0; i < 36; ++i) Ji[i] = Jj[i] = 0;
      auto JI = [&](int r, int c) -> double& { return Ji[r + 6 * c]; };
      auto JJ = [&](int r, int c) -> double& { return Jj[r + 6 * c]; };
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { JI(r, c) = -Ra[r + 3 * c]; JJ(r, c) = Re[r + 3 * c]; }
      {  // dte/dqi = Ra * skewT(tb); skewT rows [0,-z,y],[z,0,-x],[-y,x,0] with doubled entries
        const double x = 2 * tb[0], y = 2 * tb[1], z = 2 * tb[2];
        double S[9] = {0, z, -y, -z, 0, x, y, -x, 0};  // col-major
        double T[9];
        mm<3, 3, 3>(Ra, S, T);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) JI(r, 3 + c) = T[r + 3 * c];
      }
      double buf[27];  // M (9x3 col-major): column k = vec(Mk)
      {  // dre/dqi : skewT(Sxt,Syt,Szt,Rb) isometry3d_gradients.h:71-84
        const double r11 = 2 * Rb[0], r12 = 2 * Rb[3], r13 = 2 * Rb[6], r21 = 2 * Rb[1], r22 = 2 * Rb[4],
                     r23 = 2 * Rb[7], r31 = 2 * Rb[2], r32 = 2 * Rb[5], r33 = 2 * Rb[8];
        // row-major listings from the reference, transposed into col-major storage
        double Sxt[9] = {0, r31, -r21, 0, r32, -r22, 0, r33, -r23};
        double Syt[9] = {-r31, 0, r11, -r32, 0, r12, -r33, 0, r13};
        double Szt[9] = {r21, -r11, 0, r22, -r12, 0, r23, -r13, 0};
        mm<3, 3, 3>(Ra, Sxt, buf); mm<3, 3, 3>(Ra, Syt, buf + 9); mm<3, 3, 3>(Ra, Szt, buf + 18);
        double Q[9];
        mm<3, 9, 3>(dq, buf, Q);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) JI(3 + r, 3 + c) = Q[r + 3 * c];
      }
      {  // dre/dqj : skew(Sx,Sy,Sz,I) isometry3d_gradients.h:58-69
        const double r11 = 2, r12 = 0, r13 = 0, r21 = 0, r22 = 2, r23 = 0, r31 = 0, r32 = 0, r33 = 2;
        double Sx[9] = {0, -r31, r21, 0, -r32, r22, 0, -r33, r23};
        double Sy[9] = {r31, 0, -r11, r32, 0, -r12, r33, 0, -r13};
        double Sz[9] = {-r21, r11, 0, -r22, r12, 0, -r23, r13, 0};
        mm<3, 3, 3>(Re, Sx, buf); mm<3, 3, 3>(Re, Sy, buf + 9); mm<3, 3, 3>(Re, Sz, buf + 18);
        double Q[9];
        mm<3, 9, 3>(dq, buf, Q);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) JJ(3 + r, 3 + c) = Q[r + 3 * c];
      }
      break;
    }
    case ORC_EDGE_P2MC: {  // types/sba/types_sba.cpp:334-403
      const double* cam = v1->est;
      const double* w2n = cam + 12;
      const double* pt = v0->est;
      double pc[3];
      for (int r = 0; r < 3; ++r)
        pc[r] = w2n[r] * pt[0] + w2n[r + 3] * pt[1] + w2n[r + 6] * pt[2] + w2n[r + 9] * 1.0;
      double px = pc[0], py = pc[1], pz = pc[2];
      double ipz2 = 1.0 / (pz * pz);
      if (std::isnan(ipz2)) { fprintf(stderr, "[SetJac] infinite jac\n"); abort(); }
      double ipz2fx = ipz2 * cam[7], ipz2fy = ipz2 * cam[8];
      double pwt[3] = {pt[0] - cam[0], pt[1] - cam[1], pt[2] - cam[2]};
      double* Jxi = e->Ji;  // 2x3
      double* Jxj = e->Jj;  // 2x6
      auto setcol = [&](double* J, int c, const double* dp) {
        J[0 + 2 * c] = (pz * dp[0] - px * dp[2]) * ipz2fx;
        J[1 + 2 * c] = (pz * dp[1] - py * dp[2]) * ipz2fy;
      };
      double dp[3];
      mm<3, 3, 1>(cam + 36, pwt, dp); setcol(Jxj, 3, dp);
      mm<3, 3, 1>(cam + 45, pwt, dp); setcol(Jxj, 4, dp);
      mm<3, 3, 1>(cam + 54, pwt, dp); setcol(Jxj, 5, dp);
      for (int k = 0; k < 3; ++k) {
        for (int r = 0; r < 3; ++r) dp[r] = -w2n[r + 3 * k];
        setcol(Jxj, k, dp);
      }
      for (int k = 0; k < 3; ++k) {
        for (int r = 0; r < 3; ++r) dp[r] = w2n[r + 3 * k];
        setcol(Jxi, k, dp);
      }
      break;
    }
    case ORC_EDGE_XYZ2UV: {  // types/sba/types_six_dof_expmap.cpp:288-326
      double p[3];
      quat_rotate(v1->est + 3, v0->est, p);
      for (int i = 0; i < 3; ++i) p[i] += v1->est[i];
      const double x = p[0], y = p[1], z = p[2], z_2 = z * z, f = e->camPar[0];
      const double tmp[6] = {f, 0, 0, f, -x / z * f, -y / z * f};  // 2x3 col-major
      double R[9], tR[6];
      quat_to_R(v1->est + 3, R);
      mm<2, 3, 3>(tmp, R, tR);
      for (int i = 0; i < 6; ++i) e->Ji[i] = -1. / z * tR[i];
      double* J = e->Jj;  // 2x6 col-major
      J[0] = x * y / z_2 * f;        J[1] = (1 + y * y / z_2) * f;
      J[2] = -(1 + (x * x / z_2)) * f; J[3] = -x * y / z_2 * f;
      J[4] = y / z * f;              J[5] = -x / z * f;
      J[6] = -1. / z * f;            J[7] = 0;
      J[8] = 0;                      J[9] = -1. / z * f;
      J[10] = x / z_2 * f;           J[11] = y / z_2 * f;
      break;
    }
    case ORC_EDGE_SE2_XY: {  // types/slam2d/edge_se2_pointxy.cpp:67-95
      const double x1 = v0->est[0], y1 = v0->est[1], th1 = v0->est[2];
      const double x2 = v1->est[0], y2 = v1->est[1];
      double aux_1 = cos(th1);
      double aux_2 = -aux_1;
      double aux_3 = sin(th1);
      double* Ji = e->Ji;  // 2 x 3
      double* Jj = e->Jj;  // 2 x 2
      Ji[0 + 2 * 0] = aux_2;
      Ji[0 + 2 * 1] = -aux_3;
      Ji[0 + 2 * 2] = aux_1 * y2 - aux_1 * y1 - aux_3 * x2 + aux_3 * x1;
      Ji[1 + 2 * 0] = aux_3;
      Ji[1 + 2 * 1] = aux_2;
      Ji[1 + 2 * 2] = -aux_3 * y2 + aux_3 * y1 - aux_1 * x2 + aux_1 * x1;
      Jj[0 + 2 * 0] = aux_1;
      Jj[0 + 2 * 1] = aux_3;
      Jj[1 + 2 * 0] = -aux_3;
      Jj[1 + 2 * 1] = aux_1;
      break;
    }
    case ORC_EDGE_SE3_XYZ: {  // types/slam3d/edge_se3_pointxyz.cpp:110-135
      Iso X, O; memcpy(&X, v0->est, sizeof(Iso)); memcpy(&O, e->offset, sizeof(Iso));
      Iso w2l = iso_inverse(X);
      double Zcam[3];
      for (int r = 0; r < 3; ++r) Zcam[r] = w2l.R[r] * v1->est[0] + w2l.R[r + 3] * v1->est[1] + w2l.R[r + 6] * v1->est[2] + w2l.t[r];
      double J[27];  // 3 x 9; constructor: fill(0), block(0,0) = -I
      for (int i = 0; i < 27; ++i) J[i] = 0;
      J[0 + 3 * 0] = -1; J[1 + 3 * 1] = -1; J[2 + 3 * 2] = -1;
      J[0 + 3 * 4] = -2 * Zcam[2];
      J[0 + 3 * 5] = 2 * Zcam[1];
      J[1 + 3 * 3] = 2 * Zcam[2];
      J[1 + 3 * 5] = -2 * Zcam[0];
      J[2 + 3 * 3] = -2 * Zcam[1];
      J[2 + 3 * 4] = 2 * Zcam[0];
      for (int i = 0; i < 9; ++i) J[18 + i] = w2l.R[i];
      Iso Oi = iso_inverse(O);
      double Jhom[27];
      mm<3, 3, 9>(Oi.R, J, Jhom);
      for (int i = 0; i < 18; ++i) e->Ji[i] = Jhom[i];
      for (int i = 0; i < 9; ++i) e->Jj[i] = Jhom[18 + i];
      break;
    }
  }
}

void oplus(Vertex* v, const double* u) {
  switch (v->kind) {
    case ORC_VERTEX_SE2: {  // types/slam2d/vertex_se2.h:51-58
      v->est[0] += u[0]; v->est[1] += u[1];
      v->est[2] = normalize_theta(v->est[2] + u[2]);
      break;
    }
    case ORC_VERTEX_SE3: {  // types/slam3d/vertex_se3.h:107-116, isometry3d_mappings.cpp:84-91,117-122
      Iso inc;
      double w = 1 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
      if (w < 0) {
        for (int i = 0; i < 9; ++i) inc.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
      } else {
        w = sqrt(w);
        double q[4] = {u[3], u[4], u[5], w};
        quat_to_R(q, inc.R);
      }
      inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
      Iso cur; memcpy(&cur, v->est, sizeof(Iso));
      Iso r = iso_mul(cur, inc);
      if (++v->numOplusCalls > 1000) {  // orthogonalizeAfter, isometry3d_mappings.h:86-91
        v->numOplusCalls = 0;
        double E[9], RE[9];
        mtm<3, 3, 3>(r.R, r.R, E);
        E[0] -= 1; E[4] -= 1; E[8] -= 1;
        mm<3, 3, 3>(r.R, E, RE);
        for (int i = 0; i < 9; ++i) r.R[i] -= 0.5 * RE[i];
      }
      memcpy(v->est, &r, sizeof(Iso));
      break;
    }
    case ORC_VERTEX_CAM: {  // types/sba/types_sba.h:93-100, sbacam.h:101-117
      double* t = v->est; double* q = v->est + 3;
      t[0] += u[0]; t[1] += u[1]; t[2] += u[2];
      double qr[4] = {u[3], u[4], u[5], 0};
      qr[3] = sqrt(1.0 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]));
      // Eigen quaternion product a*b, a=_r, b=qr
      double a[4] = {q[0], q[1], q[2], q[3]};
      double r[4];
      r[3] = a[3] * qr[3] - a[0] * qr[0] - a[1] * qr[1] - a[2] * qr[2];
      r[0] = a[3] * qr[0] + a[0] * qr[3] + a[1] * qr[2] - a[2] * qr[1];
      r[1] = a[3] * qr[1] + a[1] * qr[3] + a[2] * qr[0] - a[0] * qr[2];
      r[2] = a[3] * qr[2] + a[2] * qr[3] + a[0] * qr[1] - a[1] * qr[0];
      quat_normalize(r);
      for (int i = 0; i < 4; ++i) q[i] = r[i];
      cam_refresh(v->est);
      break;
    }
    case ORC_VERTEX_SE3_EXPMAP: {  // types/sba/types_six_dof_expmap.h:101-104: exp(update) * estimate
      double inc[7], r[7];
      se3quat_exp(u, inc);
      se3quat_mul(inc, v->est, r);
      for (int i = 0; i < 7; ++i) v->est[i] = r[i];
      break;
    }
    case ORC_VERTEX_XYZ: {  // types/sba/types_sba.h:151-155, types/slam3d/vertex_pointxyz.h:49-52
      v->est[0] += u[0]; v->est[1] += u[1]; v->est[2] += u[2];
      break;
    }
    case ORC_VERTEX_XY: {  // types/slam2d/vertex_point_xy.h:76-80
      v->est[0] += u[0]; v->est[1] += u[1];
      break;
    }
  }
}

// core/base_edge.h:58-61
inline double edge_chi2(const Edge* e) {
  const int D = e->D;
  double s = 0;
  for (int r = 0; r < D; ++r) {
    double t = 0;
    for (int k = 0; k < D; ++k) t += e->info[r + D * k] * e->err[k];
    s += e->err[r] * t;
  }
  return s;
}
// core/robust_kernel_impl.cpp:65-126: rho = [rho(e2), rho'(e2), rho''(e2)]
// The oracle calls the REFERENCE's own kernels (core/robust_kernel_impl.cpp compiled unmodified into
// oracle/_ref/libg2o_ref_wrap.so, entered through oracle/ref_wrap.cpp); the restatement below stays as the cross-check of
// tests/test_oracle.py (bit-equal) and as the formula sheet of the device code (csrc/kernels.cuh: robustify).
inline void robustify_restated(int kind, double delta, double e2, double rho[3]);
inline void robustify(int kind, double delta, double e2, double rho[3]) {
  if (kind >= 1 && kind <= 5) { ref_robustify(kind, delta, e2, rho); return; }
  robustify_restated(kind, delta, e2, rho);
}
inline void robustify_restated(int kind, double delta, double e2, double rho[3]) {
  const double dsqr = delta * delta;
  switch (kind) {
    case 1:  // RobustKernelHuber :65-79
      if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
      else { double sqrte = sqrt(e2); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
      break;
    case 2: {  // RobustKernelPseudoHuber :81-90
      double dsqrReci = 1. / dsqr, aux1 = dsqrReci * e2 + 1.0, aux2 = sqrt(aux1);
      rho[0] = 2 * dsqr * (aux2 - 1); rho[1] = 1. / aux2; rho[2] = -0.5 * dsqrReci * rho[1] / aux1;
      break;
    }
    case 3: {  // RobustKernelCauchy :92-100
      double dsqrReci = 1. / dsqr, aux = dsqrReci * e2 + 1.0;
      rho[0] = dsqr * log(aux); rho[1] = 1. / aux; rho[2] = -dsqrReci * rho[1] * rho[1];
      break;
    }
    case 4:  // RobustKernelSaturated :102-114
      if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; } else { rho[0] = dsqr; rho[1] = 0.; rho[2] = 0.; }
      break;
    case 5: {  // RobustKernelDCS :116-127 (delta is phi)
      double scale = (2.0 * delta) / (delta + e2);
      if (scale >= 1.0) scale = 1.0;
      rho[0] = scale * e2 * scale; rho[1] = scale * scale; rho[2] = 0;
      break;
    }
    default: rho[0] = e2; rho[1] = 1.; rho[2] = 0.;
  }
}

// core/base_binary_edge.hpp:54-120
template <int D, int Di, int Dj>
void construct_quadratic_form_t(Edge* e) {
  Vertex* from = e->v[0];
  Vertex* to = e->v[1];
  const double* A = e->Ji;  // D x Di
  const double* B = e->Jj;  // D x Dj
  const bool fromNotFixed = !from->fixed, toNotFixed = !to->fixed;
  if (!(fromNotFixed || toNotFixed)) return;
  const double* omega = e->info;
  double omega_r[D];
  for (int r = 0; r < D; ++r) {
    double s = 0;
    for (int k = 0; k < D; ++k) s += omega[r + D * k] * e->err[k];
    omega_r[r] = -s;
  }
  double weightedOmega[D * D];
  if (e->rkKind != 0) {  // :91-113 robust (weighted) error according to some kernel
    double rho[3];
    robustify(e->rkKind, e->rkDelta, edge_chi2(e), rho);
    for (int i = 0; i < D * D; ++i) weightedOmega[i] = rho[1] * e->info[i];  // robustInformation, base_edge.h:96-102
    for (int r = 0; r < D; ++r) omega_r[r] *= rho[1];
    omega = weightedOmega;  // A^T wO A, A^T wO B (resp. B^T wO A), B^T wO B: the products below with omega = wO
  }
  if (fromNotFixed) {
    double AtO[Di * D];  // Di x D
    mtm<Di, D, D>(A, omega, AtO);
    double tb[Di];
    mtm<Di, D, 1>(A, omega_r, tb);
    for (int i = 0; i < Di; ++i) from->b[i] += tb[i];
    double AtOA[Di * Di];
    mm<Di, D, Di>(AtO, A, AtOA);
    for (int i = 0; i < Di * Di; ++i) from->H[i] += AtOA[i];
    if (toNotFixed) {
      if (e->transposed) {  // _hessianTransposed (Dj x Di) += B^T * AtO^T
        double T[Dj * Di];
        for (int c = 0; c < Di; ++c)
          for (int r = 0; r < Dj; ++r) {
            double s = 0;
            for (int k = 0; k < D; ++k) s += B[k + D * r] * AtO[c + Di * k];
            T[r + Dj * c] = s;
          }
        for (int i = 0; i < Dj * Di; ++i) e->H[i] += T[i];
      } else {  // _hessian (Di x Dj) += AtO * B
        double T[Di * Dj];
        mm<Di, D, Dj>(AtO, B, T);
        for (int i = 0; i < Di * Dj; ++i) e->H[i] += T[i];
      }
    }
  }
  if (toNotFixed) {
    double tb[Dj];
    mtm<Dj, D, 1>(B, omega_r, tb);
    for (int i = 0; i < Dj; ++i) to->b[i] += tb[i];
    double BtO[Dj * D];
    mtm<Dj, D, D>(B, omega, BtO);
    double BtOB[Dj * Dj];
    mm<Dj, D, Dj>(BtO, B, BtOB);
    for (int i = 0; i < Dj * Dj; ++i) to->H[i] += BtOB[i];
  }
}
void construct_quadratic_form(Edge* e) {
  switch (e->kind) {
    case ORC_EDGE_SE2: construct_quadratic_form_t<3, 3, 3>(e); break;
    case ORC_EDGE_SE3: construct_quadratic_form_t<6, 6, 6>(e); break;
    case ORC_EDGE_P2MC: construct_quadratic_form_t<2, 3, 6>(e); break;
    case ORC_EDGE_XYZ2UV: construct_quadratic_form_t<2, 3, 6>(e); break;
    case ORC_EDGE_SE2_XY: construct_quadratic_form_t<2, 3, 2>(e); break;
    case ORC_EDGE_SE3_XYZ: construct_quadratic_form_t<3, 6, 3>(e); break;
  }
}
// ---------------------------------------------------------------------------------------------
// SparseBlockMatrix (core/sparse_block_matrix.h:61-220): per block column a std::map<row, block*>
// ---------------------------------------------------------------------------------------------
struct BlockPool {  // bump allocator for the block payloads
  std::vector<std::unique_ptr<double[]>> chunks;
  size_t used = 0, cap = 0;
  double* get(size_t n) {
    if (used + n > cap) {
      cap = std::max<size_t>(n, 1 << 20);
      chunks.emplace_back(new double[cap]);
      used = 0;
    }
    double* p = chunks.back().get() + used;
    used += n;
    return p;
  }
};
struct SBM {
  int rdim = 0, cdim = 0, nrows = 0, ncols = 0;  // uniform block dims (fix* solvers); #block rows / cols
  // ... or, for the variable-size flavour (SparseBlockMatrix<MatrixXd> of BlockSolverX, `*_var` solvers): cumulative
  // offsets, base[i] = first scalar row / column of block i, base[n] = total (sparse_block_matrix.h:198-204 keeps the
  // END offsets in _rowBlockIndices; same information)
  std::vector<int> rbase, cbase;
  std::vector<std::map<int, double*>> cols;
  BlockPool pool;
  int rd(int r) const { return rbase.empty() ? rdim : rbase[r + 1] - rbase[r]; }
  int cd(int c) const { return cbase.empty() ? cdim : cbase[c + 1] - cbase[c]; }
  int rb(int r) const { return rbase.empty() ? r * rdim : rbase[r]; }
  int cb(int c) const { return cbase.empty() ? c * cdim : cbase[c]; }
  int maxdim() const { int m = cdim; for (int c = 0; c < ncols && !cbase.empty(); ++c) m = std::max(m, cd(c)); return m; }
  void resize(int nr, int nc, int rd_, int cd_) { nrows = nr; ncols = nc; rdim = rd_; cdim = cd_; rbase.clear(); cbase.clear(); cols.assign(nc, {}); }
  void resize_var(const std::vector<int>& base) {  // square, block i spans [base[i], base[i+1])
    nrows = ncols = (int)base.size() - 1; rdim = cdim = 0; rbase = base; cbase = base; cols.assign(ncols, {});
  }
  double* block(int r, int c, bool alloc) {
    auto it = cols[c].find(r);
    if (it != cols[c].end()) return it->second;
    if (!alloc) return nullptr;
    const int sz = rd(r) * cd(c);
    double* p = pool.get((size_t)sz);
    for (int i = 0; i < sz; ++i) p[i] = 0;
    cols[c][r] = p;
    return p;
  }
  void clear() {  // sparse_block_matrix.hpp:67-83 (zero every block, keep structure)
    for (int c = 0; c < ncols; ++c) for (auto& kv : cols[c]) memset(kv.second, 0, sizeof(double) * rd(kv.first) * cd(c));
  }
  size_t nonZeroBlocks() const { size_t n = 0; for (auto& c : cols) n += c.size(); return n; }
};
struct CCSCol { int row; double* block; };

// ---------------------------------------------------------------------------------------------
// LinearSolverCSparse (solvers/csparse/linear_solver_csparse.h:70-345)
// ---------------------------------------------------------------------------------------------
struct LinearSolverCSparseO {
  bool blockOrdering = true;
  css* S = nullptr;
  cs A{};  // scalar CCS, upper
  std::vector<int> Ap, Ai; std::vector<double> Ax;
  std::vector<int> blockPerm;   // P of cs_amd on the block pattern (or scalar P)
  std::vector<double> work; std::vector<int> iwork;
  double timeSymbolic = 0, timeNumeric = 0;
  ~LinearSolverCSparseO() { if (S) cs_sfree(S); }
  void init() { if (S) { cs_sfree(S); S = nullptr; } }

  // core/sparse_block_matrix_ccs.h:143-199 fillCCS(upperTriangle=true)
  void fill(const SBM& M, bool onlyValues) {
    const int n = M.cb(M.ncols);
    if (!onlyValues) {
      size_t nz = 0;
      for (int c = 0; c < M.ncols; ++c)
        for (auto& kv : M.cols[c]) nz += (kv.first == c) ? M.cd(c) * (M.cd(c) + 1) / 2 : M.rd(kv.first) * M.cd(c);
      Ap.assign(n + 1, 0); Ai.assign(nz, 0); Ax.assign(nz, 0);
    }
    int nz = 0;
    for (int i = 0; i < M.ncols; ++i) {
      int cstart = M.cb(i);
      const int dc = M.cd(i);
      for (int c = 0; c < dc; ++c) {
        if (!onlyValues) Ap[cstart + c] = nz;
        for (auto& kv : M.cols[i]) {
          int rstart = M.rb(kv.first);
          const int d = M.rd(kv.first);
          int elems = d;
          if (rstart == cstart) elems = c + 1;
          for (int r = 0; r < elems; ++r) {
            Ax[nz] = kv.second[r + d * c];
            if (!onlyValues) Ai[nz] = rstart + r;
            ++nz;
          }
        }
      }
    }
    if (!onlyValues) Ap[n] = nz;
    A.nzmax = (int)Ai.size(); A.m = A.n = n; A.p = Ap.data(); A.i = Ai.data(); A.x = Ax.data(); A.nz = -1;
  }

  // linear_solver_csparse.h:246-300
  void computeSymbolic(const SBM& M) {
    double t = now();
    const int n = A.n;
    if (!blockOrdering) {
      S = cs_schol(1, &A);
      blockPerm.clear();
    } else {
      // core/sparse_block_matrix.hpp:519-545 fillBlockStructure (upper incl. diagonal)
      std::vector<int> Bp(M.ncols + 1), Bi;
      for (int c = 0; c < M.ncols; ++c) {
        Bp[c] = (int)Bi.size();
        for (auto& kv : M.cols[c]) if (kv.first <= c) Bi.push_back(kv.first);
      }
      Bp[M.ncols] = (int)Bi.size();
      cs aux{};
      aux.nzmax = (int)Bi.size(); aux.m = aux.n = M.ncols; aux.p = Bp.data(); aux.i = Bi.data(); aux.x = nullptr; aux.nz = -1;
      int* P = cs_amd(1, &aux);
      blockPerm.assign(P, P + M.ncols);
      std::vector<int> scalarPerm(n);
      size_t idx = 0;
      for (int i = 0; i < M.ncols; ++i) {
        int base = M.cb(P[i]);  // linear_solver_csparse.h:275-283: rowBaseOfBlock / rowsOfBlock of the permuted block
        for (int j = 0; j < M.cd(P[i]); ++j) scalarPerm[idx++] = base++;
      }
      cs_free(P);
      S = (css*)cs_calloc(1, sizeof(css));
      S->pinv = cs_pinv(scalarPerm.data(), n);
      cs* C = cs_symperm(&A, S->pinv, 0);
      S->parent = cs_etree(C, 0);
      int* post = cs_post(S->parent, n);
      int* c = cs_counts(C, S->parent, post, 0);
      cs_free(post);
      cs_spfree(C);
      S->cp = (int*)cs_malloc(n + 1, sizeof(int));
      S->unz = S->lnz = cs_cumsum(S->cp, c, n);
      cs_free(c);
      if (S->lnz < 0) { cs_sfree(S); S = nullptr; }
    }
    timeSymbolic = now() - t;
  }

  // linear_solver_csparse.h:106-142
  bool solve(const SBM& M, double* x, const double* b) {
    fill(M, S != nullptr);
    if (!S) computeSymbolic(M);
    if (!S) return false;
    const int n = A.n;
    if ((int)work.size() < n) { work.assign(2 * n, 0); iwork.assign(4 * n, 0); }
    double t = now();
    if (x != b) memcpy(x, b, n * sizeof(double));
    int ok = g2o::csparse_extension::cs_cholsolsymb(&A, x, S, work.data(), iwork.data());
    timeNumeric = now() - t;
    return ok != 0;
  }

  // linear_solver_csparse.h:190-225 solvePattern + core/marginal_covariance_cholesky.cpp:55-100, 152-214:
  // numeric factor through cs_chol_workspace, then the recursive sparse-inverse formula with a memo
  // (std::map stands in for the reference's unordered_map: same values, no iteration-order dependence)
  bool solvePattern(const SBM& M, int nblocks, const int* brow, const int* bcol, double* out) {
    fill(M, S != nullptr);
    if (!S) computeSymbolic(M);
    if (!S) return false;
    const int n = A.n, d = M.maxdim();  // output blocks are d x d (variable sizes: padded with zeros)
    if ((int)work.size() < n) { work.assign(2 * n, 0); iwork.assign(4 * n, 0); }
    csn* N = g2o::csparse_extension::cs_chol_workspace(&A, S, iwork.data(), work.data());
    if (!N) return false;
    const int* Lp = N->L->p; const int* Li = N->L->i; const double* Lx = N->L->x;
    const int* perm = S->pinv;
    std::vector<double> diag(n);
    for (int r = 0; r < n; ++r) diag[r] = 1.0 / Lx[Lp[r]];  // :63-68
    std::map<long long, double> memo;
    std::function<double(int, int)> entry = [&](int r, int c) -> double {  // computeEntry :71-100 (r <= c)
      const long long idx = (long long)r * n + c;
      auto it = memo.find(idx);
      if (it != memo.end()) return it->second;
      double s = 0.;
      for (int j = Lp[r] + 1; j < Lp[r + 1]; ++j) {
        const int rr = Li[j];
        const double val = rr < c ? entry(rr, c) : entry(c, rr);
        s += val * Lx[j];
      }
      const double result = r == c ? diag[r] * (diag[r] - s) : -s * diag[r];
      memo[idx] = result;
      return result;
    };
    struct Elem { int r, c; };
    std::vector<Elem> todo;
    for (int q = 0; q < nblocks; ++q)
      for (int ir = 0; ir < M.rd(brow[q]); ++ir)
        for (int ic = 0; ic < M.cd(bcol[q]); ++ic) {
          int r = perm[M.rb(brow[q]) + ir], c = perm[M.cb(bcol[q]) + ic];
          if (r > c) std::swap(r, c);
          todo.push_back({r, c});
        }
    // sort the elems to reduce the recursive calls (:39-43: descending column, then descending row)
    std::sort(todo.begin(), todo.end(), [](const Elem& a2, const Elem& b2) { return a2.c > b2.c || (a2.c == b2.c && a2.r > b2.r); });
    for (const Elem& e : todo) entry(e.r, e.c);
    for (int q = 0; q < nblocks; ++q)
      for (int ir = 0; ir < M.rd(brow[q]); ++ir)
        for (int ic = 0; ic < M.cd(bcol[q]); ++ic) {
          int r = perm[M.rb(brow[q]) + ir], c = perm[M.cb(bcol[q]) + ic];
          if (r > c) std::swap(r, c);
          out[(size_t)q * d * d + ir + (size_t)ic * d] = memo[(long long)r * n + c];  // column-major block
        }
    cs_nfree(N);
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// the optimizer: OptimizableGraph + SparseOptimizer + BlockSolver + GN/LM
// ---------------------------------------------------------------------------------------------
}  // namespace

static int pcg_solve(int nb, int d, const int* colptr, const int* rowidx, const double* values, double* x, const double* b,
                     double tolerance, int absolute_tolerance, int max_iter, double* residual_io);

struct oracle_graph {
  typedef std::tr1::unordered_map<int, Vertex*> VertexIDMap;  // core/hyper_graph.h VertexIDMap
  // LinearSolverPCG as the BlockSolver's linear solver (solvers/pcg/solver_pcg.cpp: `*_pcg*`): oracle_set_linear_solver
  bool usePCG = false;
  double pcgTolerance = 1e-6, pcgResidual = -1.0;
  int pcgAbsolute = 1, pcgMaxIter = -1, pcgIterations = 0;
  VertexIDMap vertices;
  std::vector<std::unique_ptr<Vertex>> vstore;
  std::vector<std::unique_ptr<Edge>> edges;  // addEdge order = internalId
  int rkKind = 0;            // robust kernel given to every edge (g2o.cpp:322-336)
  std::map<int, std::array<double, 4>> cameraParameters;  // PARAMS_CAMERAPARAMETERS id -> f cx cy baseline
  std::map<int, std::array<double, 12>> se3Offsets;       // PARAMS_SE3OFFSET id -> offset as Iso [R col-major | t]
  double rkDelta = 1.0;
  // SparseOptimizer
  std::vector<Vertex*> activeVertices, ivMap;
  std::vector<Edge*> activeEdges;
  // BlockSolver (core/block_solver.h:43-186)
  bool doSchur = false;
  int numPoses = 0, numLandmarks = 0, sizePoses = 0, sizeLandmarks = 0, poseDim = 0, landmarkDim = 0;
  SBM Hpp, Hll, Hpl, Hschur;
  std::vector<std::vector<CCSCol>> HplCCS;               // per landmark column, ascending pose row
  std::vector<std::vector<CCSCol>> HschurTransposedCCS;  // per row i1: (i2, block) ascending
  std::vector<double> DInvSchur;
  std::vector<double> x, b, coefficients, bschur, diagBackupPose, diagBackupLandmark;
  LinearSolverCSparseO linearSolver;
  // LM state (core/optimization_algorithm_levenberg.h)
  double currentLambda = -1, ni = 2;
  int levenbergIterations = 0;
  double timeSchur = 0, timeLinearSolver = 0;

  Vertex* vertex(int id) { auto it = vertices.find(id); return it == vertices.end() ? nullptr : it->second; }
};

namespace {
typedef oracle_graph G;

Vertex* add_vertex(G* g, int kind, int id) {
  if (g->vertex(id)) return nullptr;  // HyperGraph::addVertex fails on duplicate ids
  g->vstore.emplace_back(new Vertex());
  Vertex* v = g->vstore.back().get();
  v->kind = kind; v->id = id; v->dim = vertex_dim(kind);
  for (int i = 0; i < EST_MAX; ++i) v->est[i] = 0;
  g->vertices[id] = v;
  return v;
}

// per-type read(): payload = numbers after the id
bool vertex_read(Vertex* v, const double* p, int n) {
  switch (v->kind) {
    case ORC_VERTEX_SE2:  // types/slam2d/vertex_se2.cpp:41-47
      if (n < 3) return false;
      v->est[0] = p[0]; v->est[1] = p[1]; v->est[2] = p[2];
      return true;
    case ORC_VERTEX_SE3: {  // types/slam3d/vertex_se3.cpp:44-51 ; fromVectorQT (no normalisation)
      if (n < 7) return false;
      double q[4] = {p[3], p[4], p[5], p[6]};
      quat_to_R(q, v->est);
      v->est[9] = p[0]; v->est[10] = p[1]; v->est[11] = p[2];
      return true;
    }
    case ORC_VERTEX_CAM: {  // types/sba/types_sba.cpp:74-112
      if (n < 7) return false;
      v->est[0] = p[0]; v->est[1] = p[1]; v->est[2] = p[2];
      double q[4] = {p[3], p[4], p[5], p[6]};
      quat_normalize(q);
      se3quat_normalize_rotation(q);  // SBACam(r,t) -> SE3Quat(q,t) ctor
      for (int i = 0; i < 4; ++i) v->est[3 + i] = q[i];
      if (n >= 12) { for (int i = 0; i < 5; ++i) v->est[7 + i] = p[7 + i]; }
      else { v->est[7] = 300; v->est[8] = 300; v->est[9] = 320; v->est[10] = 320; v->est[11] = 0.1; }
      cam_refresh(v->est);
      return true;
    }
    case ORC_VERTEX_SE3_EXPMAP: {  // types/sba/types_six_dof_expmap.cpp:76-84: file holds cam2world, estimate = inverse
      if (n < 7) return false;
      double c2w[7];
      for (int i = 0; i < 7; ++i) c2w[i] = p[i];  // SE3Quat::fromVector: no normalisation (se3quat.h:157-160)
      se3quat_inverse(c2w, v->est);
      return true;
    }
    case ORC_VERTEX_XYZ:  // types/sba/types_sba.cpp:180-186; VertexPointXYZ: types/slam3d/vertex_pointxyz.cpp read
      if (n < 3) return false;
      v->est[0] = p[0]; v->est[1] = p[1]; v->est[2] = p[2];
      return true;
    case ORC_VERTEX_XY:  // types/slam2d/vertex_point_xy.cpp read
      if (n < 2) return false;
      v->est[0] = p[0]; v->est[1] = p[1];
      return true;
  }
  return false;
}

bool edge_read(Edge* e, const double* p, int n) {
  const int D = e->D;
  for (int i = 0; i < 36; ++i) e->info[i] = 0;
  for (int i = 0; i < D; ++i) e->info[i + D * i] = 1;
  switch (e->kind) {
    case ORC_EDGE_SE2: {  // types/slam2d/edge_se2.cpp:40-53
      if (n < 9) return false;
      e->meas[0] = p[0]; e->meas[1] = p[1]; e->meas[2] = p[2];
      SE2 zi = se2_inv(SE2{p[0], p[1], p[2]});
      e->invMeas[0] = zi.x; e->invMeas[1] = zi.y; e->invMeas[2] = zi.th;
      int k = 3;
      for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) { e->info[i + 3 * j] = p[k]; e->info[j + 3 * i] = p[k]; ++k; }
      return true;
    }
    case ORC_EDGE_SE3: {  // types/slam3d/edge_se3.cpp:14-36
      if (n < 7) return false;
      double q[4] = {p[3], p[4], p[5], p[6]};
      quat_normalize(q);  // Vector4d::MapType(meas.data()+3).normalize()
      Iso Z;
      quat_to_R(q, Z.R);
      Z.t[0] = p[0]; Z.t[1] = p[1]; Z.t[2] = p[2];
      memcpy(e->meas, &Z, sizeof(Iso));
      Iso Zi = iso_inverse(Z);
      memcpy(e->invMeas, &Zi, sizeof(Iso));
      int k = 7;
      for (int i = 0; i < 6 && k < n; ++i) for (int j = i; j < 6 && k < n; ++j) { e->info[i + 6 * j] = p[k]; e->info[j + 6 * i] = p[k]; ++k; }
      return true;
    }
    case ORC_EDGE_P2MC: {  // types/sba/types_sba.cpp:204-213 (information forced to identity)
      if (n < 2) return false;
      e->meas[0] = p[0]; e->meas[1] = p[1];
      return true;
    }
    case ORC_EDGE_XYZ2UV: {  // types/sba/types_six_dof_expmap.cpp:241-256: paramId u v i00 i01 i11
      if (n < 6) return false;
      e->paramId = (int)p[0];
      e->meas[0] = p[1]; e->meas[1] = p[2];
      int k = 3;
      for (int i = 0; i < 2; ++i) for (int j = i; j < 2; ++j) { e->info[i + 2 * j] = p[k]; e->info[j + 2 * i] = p[k]; ++k; }
      return true;
    }
    case ORC_EDGE_SE2_XY: {  // types/slam2d/edge_se2_pointxy.cpp:41-47: x y i00 i01 i11
      if (n < 5) return false;
      e->meas[0] = p[0]; e->meas[1] = p[1];
      e->info[0] = p[2]; e->info[2] = p[3]; e->info[3] = p[4]; e->info[1] = p[3];
      return true;
    }
    case ORC_EDGE_SE3_XYZ: {  // types/slam3d/edge_se3_pointxyz.cpp:62-84: paramId x y z + upper triangle of the information
      if (n < 4) return false;
      e->paramId = (int)p[0];
      e->meas[0] = p[1]; e->meas[1] = p[2]; e->meas[2] = p[3];
      int k = 4;
      for (int i = 0; i < 3 && k < n; ++i) for (int j = i; j < 3 && k < n; ++j) { e->info[i + 3 * j] = p[k]; e->info[j + 3 * i] = p[k]; ++k; }
      return true;
    }
  }
  return false;
}

// EdgeSE2::initialEstimate / EdgeSE3::initialEstimate for vertices created by load(createEdges=true)
void initial_estimate_to(Edge* e) {  // to = from * meas
  if (e->kind == ORC_EDGE_SE2) {
    SE2 f{e->v[0]->est[0], e->v[0]->est[1], e->v[0]->est[2]};
    SE2 r = se2_mul(f, SE2{e->meas[0], e->meas[1], e->meas[2]});
    e->v[1]->est[0] = r.x; e->v[1]->est[1] = r.y; e->v[1]->est[2] = r.th;
  } else if (e->kind == ORC_EDGE_SE3) {
    Iso f, z; memcpy(&f, e->v[0]->est, sizeof(Iso)); memcpy(&z, e->meas, sizeof(Iso));
    Iso r = iso_mul(f, z); memcpy(e->v[1]->est, &r, sizeof(Iso));
  } else if (e->kind == ORC_EDGE_SE2_XY) {  // types/slam2d/edge_se2_pointxy.cpp:55-64: vj = vi * measurement
    const double c = cos(e->v[0]->est[2]), s = sin(e->v[0]->est[2]);
    e->v[1]->est[0] = c * e->meas[0] - s * e->meas[1] + e->v[0]->est[0];
    e->v[1]->est[1] = s * e->meas[0] + c * e->meas[1] + e->v[0]->est[1];
  } else if (e->kind == ORC_EDGE_SE3_XYZ) {  // types/slam3d/edge_se3_pointxyz.cpp initialEstimate: cam->estimate() * (offset * z)
    Iso X, O; memcpy(&X, e->v[0]->est, sizeof(Iso)); memcpy(&O, e->offset, sizeof(Iso));
    Iso n2w = iso_mul(X, O);
    for (int r = 0; r < 3; ++r) e->v[1]->est[r] = n2w.R[r] * e->meas[0] + n2w.R[r + 3] * e->meas[1] + n2w.R[r + 6] * e->meas[2] + n2w.t[r];
  }
}
void initial_estimate_from(Edge* e) {  // from = to * meas^-1 (the landmark edges cannot initialise their pose: left at the origin)
  if (e->kind == ORC_EDGE_SE2) {
    SE2 t{e->v[1]->est[0], e->v[1]->est[1], e->v[1]->est[2]};
    SE2 r = se2_mul(t, SE2{e->invMeas[0], e->invMeas[1], e->invMeas[2]});
    e->v[0]->est[0] = r.x; e->v[0]->est[1] = r.y; e->v[0]->est[2] = r.th;
  } else if (e->kind == ORC_EDGE_SE3) {
    Iso t, z; memcpy(&t, e->v[1]->est, sizeof(Iso)); memcpy(&z, e->meas, sizeof(Iso));
    Iso r = iso_mul(t, iso_inverse(z)); memcpy(e->v[0]->est, &r, sizeof(Iso));
  }
}
void set_to_origin(Vertex* v) {
  for (int i = 0; i < EST_MAX; ++i) v->est[i] = 0;
  if (v->kind == ORC_VERTEX_SE3) { v->est[0] = v->est[4] = v->est[8] = 1; }
  if (v->kind == ORC_VERTEX_SE3_EXPMAP) v->est[6] = 1;  // SE3Quat()
  if (v->kind == ORC_VERTEX_CAM) { v->est[6] = 1; v->est[7] = 1; v->est[8] = 1; v->est[9] = 0.5; v->est[10] = 0.5; cam_refresh(v->est); }
}

int add_edge(G* g, int kind, int id1, int id2, const double* payload, int n) {
  // core/optimizable_graph.cpp:454-520 (binary edges, createEdges = true)
  static const int vk0[6] = {ORC_VERTEX_SE2, ORC_VERTEX_SE3, ORC_VERTEX_XYZ, ORC_VERTEX_XYZ, ORC_VERTEX_SE2, ORC_VERTEX_SE3};
  static const int vk1[6] = {ORC_VERTEX_SE2, ORC_VERTEX_SE3, ORC_VERTEX_CAM, ORC_VERTEX_SE3_EXPMAP, ORC_VERTEX_XY, ORC_VERTEX_XYZ};
  if (kind == ORC_EDGE_XYZ2UV) {  // OptimizableGraph::addEdge -> resolveParameters: an unknown parameter id rejects the edge
    if (n < 1 || !g->cameraParameters.count((int)payload[0])) return -1;
  }
  if (kind == ORC_EDGE_SE3_XYZ) {
    if (n < 1 || !g->se3Offsets.count((int)payload[0])) return -1;
  }
  Vertex* from = g->vertex(id1);
  Vertex* to = g->vertex(id2);
  int doInit = 0;
  if (!from) { from = add_vertex(g, vk0[kind], id1); set_to_origin(from); doInit = 2; }
  if (!to) { to = add_vertex(g, vk1[kind], id2); set_to_origin(to); doInit = 1; }
  if (from->kind != vk0[kind] || to->kind != vk1[kind]) return -1;
  g->edges.emplace_back(new Edge());
  Edge* e = g->edges.back().get();
  e->kind = kind; e->D = edge_dim(kind);
  e->rkKind = g->rkKind; e->rkDelta = g->rkDelta;
  e->v[0] = from; e->v[1] = to;
  e->internalId = (int)g->edges.size() - 1;
  for (int i = 0; i < 6; ++i) e->err[i] = 0;
  if (!edge_read(e, payload, n)) { g->edges.pop_back(); return -1; }
  if (kind == ORC_EDGE_XYZ2UV) { const auto& cp = g->cameraParameters[e->paramId]; for (int i = 0; i < 4; ++i) e->camPar[i] = cp[i]; }
  if (kind == ORC_EDGE_SE3_XYZ) { const auto& of = g->se3Offsets[e->paramId]; for (int i = 0; i < 12; ++i) e->offset[i] = of[i]; }
  from->edges.push_back(e);
  if (to != from) to->edges.push_back(e);
  if (doInit == 1) initial_estimate_to(e);
  if (doInit == 2) initial_estimate_from(e);
  return 0;
}

// ---- apps/g2o_cli/g2o.cpp:272-320 + core/sparse_optimizer.cpp:116-164 ----
bool gauge_freedom(G* g) {
  if (g->vertices.empty()) return false;
  int maxDim = 0;
  for (auto& kv : g->vertices) maxDim = std::max(maxDim, kv.second->dim);
  for (auto& kv : g->vertices) {
    Vertex* v = kv.second;
    if (v->dim == maxDim) {
      if (v->fixed) return false;
      // unary full-dimension priors: none of the configured edge types is unary
    }
  }
  return true;
}
Vertex* find_gauge(G* g) {
  if (g->vertices.empty()) return nullptr;
  int maxDim = 0;
  for (auto& kv : g->vertices) maxDim = std::max(maxDim, kv.second->dim);
  for (auto& kv : g->vertices) if (kv.second->dim == maxDim) return kv.second;
  return nullptr;
}

// ---- core/sparse_optimizer.cpp:199-267, 166-190 ----
bool initialize_optimization(G* g) {
  if (g->edges.empty()) return false;
  for (Vertex* v : g->ivMap) v->hessianIndex = -1;
  g->ivMap.clear();
  g->activeVertices.clear();
  g->activeEdges.clear();
  std::vector<char> edgeActive(g->edges.size(), 0);
  for (auto& kv : g->vertices) {
    Vertex* v = kv.second;
    int levelEdges = 0;
    for (Edge* e : v->edges) {
      bool allFixed = e->v[0]->fixed && e->v[1]->fixed;
      if (!allFixed) { edgeActive[e->internalId] = 1; levelEdges++; }
    }
    if (levelEdges) g->activeVertices.push_back(v);
  }
  for (size_t k = 0; k < g->edges.size(); ++k) if (edgeActive[k]) g->activeEdges.push_back(g->edges[k].get());
  // sortVectorContainers: vertices by id, edges by internalId (already)
  std::sort(g->activeVertices.begin(), g->activeVertices.end(), [](Vertex* a, Vertex* b) { return a->id < b->id; });
  // buildIndexMapping
  if (g->activeVertices.empty()) return false;
  g->ivMap.resize(g->activeVertices.size());
  size_t i = 0;
  for (int k = 0; k < 2; ++k)
    for (Vertex* v : g->activeVertices) {
      if (!v->fixed) {
        if ((int)v->marginalized == k) { v->hessianIndex = (int)i; g->ivMap[i] = v; ++i; }
      } else {
        v->hessianIndex = -1;
      }
    }
  g->ivMap.resize(i);
  return true;
}

// ---- core/block_solver.hpp:142-295 ----
bool build_structure(G* g) {
  g->numPoses = g->numLandmarks = g->sizePoses = g->sizeLandmarks = 0;
  g->poseDim = g->landmarkDim = 0;
  for (Vertex* v : g->ivMap) {
    if (!v->marginalized) { v->colInHessian = g->sizePoses; g->sizePoses += v->dim; ++g->numPoses; g->poseDim = v->dim; }
    else { v->colInHessian = g->sizeLandmarks; g->sizeLandmarks += v->dim; ++g->numLandmarks; g->landmarkDim = v->dim; }
  }
  const int pd = g->poseDim, ld = g->landmarkDim ? g->landmarkDim : 3;
  g->Hpp = SBM(); g->Hll = SBM(); g->Hpl = SBM(); g->Hschur = SBM();
  // BlockSolverX (`*_var` solvers, nothing marginalized): the blocks of Hpp take the dimension of their vertices
  // (block_solver.hpp:156-176 collects blockPoseIndices); uniform dimensions stay on the fixed-size layout
  bool variable = false;
  std::vector<int> poseBase{0};
  for (Vertex* v : g->ivMap) if (!v->marginalized) { poseBase.push_back(poseBase.back() + v->dim); variable |= v->dim != pd; }
  if (variable && g->doSchur) return false;  // mixed pose dimensions under a Schur complement: not restated
  if (variable) { g->Hpp.resize_var(poseBase); for (Vertex* v : g->ivMap) g->poseDim = std::max(g->poseDim, v->dim); }
  else g->Hpp.resize(g->numPoses, g->numPoses, pd, pd);
  if (g->doSchur) {
    g->Hschur.resize(g->numPoses, g->numPoses, pd, pd);
    g->Hll.resize(g->numLandmarks, g->numLandmarks, ld, ld);
    g->Hpl.resize(g->numPoses, g->numLandmarks, pd, ld);
  }
  const int total = g->sizePoses + g->sizeLandmarks;
  g->x.assign(total, 0); g->b.assign(total, 0);
  g->coefficients.assign(total, 0); g->bschur.assign(g->sizePoses, 0);
  int poseIdx = 0, landmarkIdx = 0;
  for (Vertex* v : g->ivMap) {
    if (!v->marginalized) { v->H = g->Hpp.block(poseIdx, poseIdx, true); ++poseIdx; }
    else { v->H = g->Hll.block(landmarkIdx, landmarkIdx, true); ++landmarkIdx; }
  }
  std::vector<std::set<int>> schurLookup;  // SparseBlockMatrixHashMap: per column set of rows
  if (g->doSchur) schurLookup.resize(g->numPoses);
  for (Edge* e : g->activeEdges) {
    Vertex* v1 = e->v[0]; Vertex* v2 = e->v[1];
    int ind1 = v1->hessianIndex, ind2 = v2->hessianIndex;
    if (ind1 == -1 || ind2 == -1) continue;
    bool transposedBlock = ind1 > ind2;
    if (transposedBlock) std::swap(ind1, ind2);
    if (!v1->marginalized && !v2->marginalized) {
      e->H = g->Hpp.block(ind1, ind2, true);
      e->transposed = transposedBlock;
      if (g->doSchur) schurLookup[ind2].insert(ind1);
    } else if (v1->marginalized && v2->marginalized) {
      e->H = g->Hll.block(ind1 - g->numPoses, ind2 - g->numPoses, true);
      e->transposed = false;
    } else {
      if (v1->marginalized) {
        e->H = g->Hpl.block(v2->hessianIndex, v1->hessianIndex - g->numPoses, true);
        e->transposed = true;
      } else {
        e->H = g->Hpl.block(v1->hessianIndex, v2->hessianIndex - g->numPoses, true);
        e->transposed = false;
      }
    }
  }
  if (!g->doSchur) return true;
  g->DInvSchur.assign((size_t)g->numLandmarks * ld * ld, 0);
  g->HplCCS.assign(g->numLandmarks, {});
  for (int c = 0; c < g->numLandmarks; ++c)
    for (auto& kv : g->Hpl.cols[c]) g->HplCCS[c].push_back(CCSCol{kv.first, kv.second});
  for (Vertex* v : g->ivMap) {
    if (!v->marginalized) continue;
    for (Edge* e1 : v->edges)
      for (int i = 0; i < 2; ++i) {
        Vertex* v1 = e1->v[i];
        if (v1->hessianIndex == -1 || v1 == v) continue;
        for (Edge* e2 : v->edges)
          for (int j = 0; j < 2; ++j) {
            Vertex* v2 = e2->v[j];
            if (v2->hessianIndex == -1 || v2 == v) continue;
            int i1 = v1->hessianIndex, i2 = v2->hessianIndex;
            if (i1 <= i2) schurLookup[i2].insert(i1);
          }
      }
  }
  for (int c = 0; c < g->numPoses; ++c) for (int r : schurLookup[c]) g->Hschur.block(r, c, true);
  g->HschurTransposedCCS.assign(g->numPoses, {});
  for (int c = 0; c < g->numPoses; ++c)
    for (auto& kv : g->Hschur.cols[c]) g->HschurTransposedCCS[kv.first].push_back(CCSCol{c, kv.second});
  return true;
}

// ---- core/sparse_optimizer.cpp:61-114 ----
double compute_active_errors(G* g) {
  for (Edge* e : g->activeEdges) compute_error(e);
  double chi = 0.0;
  for (Edge* e : g->activeEdges) {  // activeRobustChi2, sparse_optimizer.cpp:100-114
    if (e->rkKind != 0) { double rho[3]; robustify(e->rkKind, e->rkDelta, edge_chi2(e), rho); chi += rho[0]; }
    else chi += edge_chi2(e);
  }
  return chi;
}

// ---- core/block_solver.hpp:501-560 ----
void build_system(G* g) {
  for (Vertex* v : g->ivMap) for (int i = 0; i < 6; ++i) v->b[i] = 0;
  g->Hpp.clear();
  if (g->doSchur) { g->Hll.clear(); g->Hpl.clear(); }
  for (Edge* e : g->activeEdges) { linearize(e); construct_quadratic_form(e); }
  for (Vertex* v : g->ivMap) {
    int iBase = v->colInHessian;
    if (v->marginalized) iBase += g->sizePoses;
    for (int i = 0; i < v->dim; ++i) g->b[iBase + i] = v->b[i];
  }
}

// ---- core/block_solver.hpp:563-604 ----
void set_lambda(G* g, double lambda, bool backup) {
  const int ld = g->landmarkDim;
  if (backup) { g->diagBackupPose.resize((size_t)g->sizePoses); g->diagBackupLandmark.resize((size_t)g->numLandmarks * ld); }
  for (int i = 0; i < g->numPoses; ++i) {
    double* blk = g->Hpp.block(i, i, false);
    const int pd = g->Hpp.cd(i), b0 = g->Hpp.cb(i);
    for (int k = 0; k < pd; ++k) { if (backup) g->diagBackupPose[(size_t)b0 + k] = blk[k + pd * k]; blk[k + pd * k] += lambda; }
  }
  for (int i = 0; i < g->numLandmarks; ++i) {
    double* blk = g->Hll.block(i, i, false);
    for (int k = 0; k < ld; ++k) { if (backup) g->diagBackupLandmark[(size_t)i * ld + k] = blk[k + ld * k]; blk[k + ld * k] += lambda; }
  }
}
void restore_diagonal(G* g) {
  const int ld = g->landmarkDim;
  for (int i = 0; i < g->numPoses; ++i) {
    double* blk = g->Hpp.block(i, i, false);
    const int pd = g->Hpp.cd(i), b0 = g->Hpp.cb(i);
    for (int k = 0; k < pd; ++k) blk[k + pd * k] = g->diagBackupPose[(size_t)b0 + k];
  }
  for (int i = 0; i < g->numLandmarks; ++i) {
    double* blk = g->Hll.block(i, i, false);
    for (int k = 0; k < ld; ++k) blk[k + ld * k] = g->diagBackupLandmark[(size_t)i * ld + k];
  }
}

// Eigen fixed 3x3 inverse (Eigen/src/LU/Inverse.h compute_inverse<.,.,3>): cofactors / determinant
inline void inverse3(const double* m, double* r) {
  auto M = [&](int i, int j) { return m[i + 3 * j]; };
  auto cof = [&](int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
  };
  double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  double det = c0 * M(0, 0) + c1 * M(1, 0) + c2 * M(2, 0);
  double invdet = 1.0 / det;
  r[0 + 3 * 0] = c0 * invdet; r[0 + 3 * 1] = c1 * invdet; r[0 + 3 * 2] = c2 * invdet;
  r[1 + 3 * 0] = cof(0, 1) * invdet; r[1 + 3 * 1] = cof(1, 1) * invdet; r[1 + 3 * 2] = cof(2, 1) * invdet;
  r[2 + 3 * 0] = cof(0, 2) * invdet; r[2 + 3 * 1] = cof(1, 2) * invdet; r[2 + 3 * 2] = cof(2, 2) * invdet;
}

// ---- core/block_solver.hpp:354-486 ----
// _linearSolver->solve(A, x, b): LinearSolverCSparse, or LinearSolverPCG (solvers/pcg/linear_solver_pcg.hpp:79-160; always true)
bool linear_solve(G* g, const SBM& M, double* x, const double* b) {
  if (!g->usePCG) return g->linearSolver.solve(M, x, b);
  if (!M.cbase.empty()) return false;  // variable block sizes under PCG: not restated
  std::vector<int> colptr(M.ncols + 1, 0), rowidx;
  std::vector<double> values;
  const int sz = M.rdim * M.cdim;
  for (int c = 0; c < M.ncols; ++c) {
    for (auto& kv : M.cols[c]) {
      if (kv.first > c) continue;
      rowidx.push_back(kv.first);
      values.insert(values.end(), kv.second, kv.second + sz);
    }
    colptr[c + 1] = (int)rowidx.size();
  }
  std::vector<double> rhs(b, b + (size_t)M.ncols * M.cdim);  // x and b may alias in the callers
  g->pcgIterations = pcg_solve(M.ncols, M.cdim, colptr.data(), rowidx.data(), values.data(), x, rhs.data(), g->pcgTolerance,
                               g->pcgAbsolute, g->pcgMaxIter, &g->pcgResidual);
  return true;
}

bool solve(G* g) {
  if (!g->doSchur) {
    double t = now();
    bool ok = linear_solve(g, g->Hpp, g->x.data(), g->b.data());
    g->timeLinearSolver = now() - t;
    return ok;
  }
  double t = now();
  const int pd = g->poseDim, ld = g->landmarkDim;
  assert(pd == 6 && ld == 3);
  g->Hschur.clear();
  for (int c = 0; c < g->numPoses; ++c)  // _Hpp->add(_Hschur)
    for (auto& kv : g->Hpp.cols[c]) {
      double* d = g->Hschur.block(kv.first, c, false);
      for (int i = 0; i < pd * pd; ++i) d[i] += kv.second[i];
    }
  memset(g->coefficients.data(), 0, g->sizePoses * sizeof(double));
  for (int l = 0; l < g->numLandmarks; ++l) {
    const double* D = g->Hll.cols[l].begin()->second;
    double* Dinv = &g->DInvSchur[(size_t)l * 9];
    inverse3(D, Dinv);
    double db0[3], db[3];
    for (int j = 0; j < 3; ++j) db0[j] = g->b[l * 3 + g->sizePoses + j];
    mm<3, 3, 1>(Dinv, db0, db);
    const std::vector<CCSCol>& col = g->HplCCS[l];
    for (size_t o = 0; o < col.size(); ++o) {
      int i1 = col[o].row;
      const double* Bi = col[o].block;  // 6x3
      double BDinv[18];
      mm<6, 3, 3>(Bi, Dinv, BDinv);
      double Bb[6];
      mm<6, 3, 1>(Bi, db, Bb);
      for (int k = 0; k < 6; ++k) g->coefficients[i1 * 6 + k] += Bb[k];
      auto target = g->HschurTransposedCCS[i1].begin();
      for (size_t in = o; in < col.size(); ++in) {  // lower_bound(i1) == o since rows are unique+sorted
        int i2 = col[in].row;
        const double* Bj = col[in].block;
        while (target->row < i2) ++target;
        double* H = target->block;
        double T[36];
        mmt<6, 3, 6>(BDinv, Bj, T);
        for (int k = 0; k < 36; ++k) H[k] -= T[k];
      }
    }
  }
  memcpy(g->bschur.data(), g->b.data(), g->sizePoses * sizeof(double));
  for (int i = 0; i < g->sizePoses; ++i) g->bschur[i] -= g->coefficients[i];
  g->timeSchur = now() - t;
  t = now();
  bool solvedPoses = linear_solve(g, g->Hschur, g->x.data(), g->bschur.data());
  g->timeLinearSolver = now() - t;
  if (!solvedPoses) return false;
  double* xp = g->x.data();
  double* cp = g->coefficients.data();
  double* xl = g->x.data() + g->sizePoses;
  double* cl = g->coefficients.data() + g->sizePoses;
  const double* bl = g->b.data() + g->sizePoses;
  for (int i = 0; i < g->sizePoses; ++i) cp[i] = -xp[i];
  memcpy(cl, bl, g->sizeLandmarks * sizeof(double));
  for (int l = 0; l < g->numLandmarks; ++l)  // _HplCCS->rightMultiply(cl, cp): cl += B^T cp
    for (const CCSCol& rb : g->HplCCS[l]) {
      double tmp[3];
      mtm<3, 6, 1>(rb.block, cp + rb.row * 6, tmp);
      for (int k = 0; k < 3; ++k) cl[l * 3 + k] += tmp[k];
    }
  memset(xl, 0, g->sizeLandmarks * sizeof(double));
  for (int l = 0; l < g->numLandmarks; ++l) {  // _DInvSchur->multiply(xl, cl)
    double tmp[3];
    mm<3, 3, 1>(&g->DInvSchur[(size_t)l * 9], cl + l * 3, tmp);
    for (int k = 0; k < 3; ++k) xl[l * 3 + k] += tmp[k];
  }
  return true;
}

void update(G* g) {  // core/sparse_optimizer.cpp:421-434
  const double* u = g->x.data();
  for (Vertex* v : g->ivMap) { oplus(v, u); u += v->dim; }
}
void push(G* g) { for (Vertex* v : g->activeVertices) { std::array<double, EST_MAX> a; memcpy(a.data(), v->est, sizeof(v->est)); v->backup.push_back(a); } }
void pop(G* g) { for (Vertex* v : g->activeVertices) { memcpy(v->est, v->backup.back().data(), sizeof(v->est)); v->backup.pop_back(); } }
void discard_top(G* g) { for (Vertex* v : g->activeVertices) v->backup.pop_back(); }

bool algorithm_init(G* g) {  // core/optimization_algorithm_with_hessian.cpp:50-73 + block_solver.hpp:606-620
  bool useSchur = false;
  for (Vertex* v : g->activeVertices) if (v->marginalized) { useSchur = true; break; }
  g->doSchur = useSchur;
  g->linearSolver.init();
  g->pcgResidual = -1.0;  // LinearSolverPCG::init (linear_solver_pcg.h:64-71)
  return true;
}

double lambda_init(G* g) {  // core/optimization_algorithm_levenberg.cpp:149-163
  double maxDiagonal = 0.;
  for (Vertex* v : g->ivMap)
    for (int j = 0; j < v->dim; ++j) maxDiagonal = std::max(fabs(v->H[j + v->dim * j]), maxDiagonal);
  return 1e-5 * maxDiagonal;
}
double compute_scale(G* g) {  // core/optimization_algorithm_levenberg.cpp:165-172
  double scale = 0.;
  for (size_t j = 0; j < g->x.size(); ++j) scale += g->x[j] * (g->currentLambda * g->x[j] + g->b[j]);
  return scale;
}

// core/optimization_algorithm_gauss_newton.cpp:50-93
int solve_gn(G* g, int iteration, oracle_iter_stats* st) {
  double t = now();
  compute_active_errors(g);
  st->time_residuals = now() - t;
  if (iteration == 0) { if (!build_structure(g)) return -1; }
  t = now();
  build_system(g);
  st->time_quadratic_form = now() - t;
  t = now();
  bool ok = solve(g);
  st->time_linear_solution = now() - t;
  t = now();
  update(g);
  st->time_update = now() - t;
  return ok ? 1 : -1;
}

// core/optimization_algorithm_levenberg.cpp:57-147
int solve_lm(G* g, int iteration, oracle_iter_stats* st) {
  if (iteration == 0) { if (!build_structure(g)) return -1; }
  double t = now();
  double currentChi = compute_active_errors(g);
  st->time_residuals = now() - t;
  t = now();
  double tempChi = currentChi;
  build_system(g);
  st->time_quadratic_form = now() - t;
  if (iteration == 0) { g->currentLambda = lambda_init(g); g->ni = 2; }
  double rho = 0;
  int& qmax = g->levenbergIterations;
  qmax = 0;
  do {
    push(g);
    t = now();
    set_lambda(g, g->currentLambda, true);
    bool ok2 = solve(g);
    st->time_linear_solution += now() - t;
    st->time_schur += g->timeSchur; st->time_linear_solver += g->timeLinearSolver;
    t = now();
    update(g);
    st->time_update = now() - t;
    restore_diagonal(g);
    tempChi = compute_active_errors(g);
    if (!ok2) tempChi = DBL_MAX;
    rho = (currentChi - tempChi);
    double scale = compute_scale(g);
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - pow((2 * rho - 1), 3);
      alpha = std::min(alpha, 2. / 3.);
      double scaleFactor = std::max(1. / 3., alpha);
      g->currentLambda *= scaleFactor;
      g->ni = 2;
      currentChi = tempChi;
      discard_top(g);
    } else {
      g->currentLambda *= g->ni;
      g->ni *= 2;
      pop(g);
    }
    qmax++;
  } while (rho < 0 && qmax < 10);
  if (qmax == 10 || rho == 0) return 2;
  return 1;
}

}  // namespace

// =============================================================================================
// C interface
// =============================================================================================
// ---------------------------------------------------------------------------------------------
// LinearSolverPCG<MatrixType>::solve, solvers/pcg/linear_solver_pcg.hpp:79-197 - restated on plain arrays in the
// reference's operation order: linear structure of the off-diagonal upper blocks in column order (:86-106), mult() =
// diagonal part first, then per upper block dest[row] += B src[col]; dest[col] += B^T src[row] (:173-196), the
// recurrences and the stopping rule of :117-150.  *residual_io carries _residual from one solve to the next.
// ---------------------------------------------------------------------------------------------
static void pcg_inverse_spd(const double* B, int d, double* J) {  // it->second->inverse() (:95): Gauss-Jordan, partial pivoting
  std::vector<double> a((size_t)d * 2 * d, 0.0);
  for (int r = 0; r < d; ++r) { for (int c = 0; c < d; ++c) a[(size_t)r * 2 * d + c] = B[r + c * d]; a[(size_t)r * 2 * d + d + r] = 1.0; }
  for (int k = 0; k < d; ++k) {
    int p = k;
    for (int r = k + 1; r < d; ++r) if (fabs(a[(size_t)r * 2 * d + k]) > fabs(a[(size_t)p * 2 * d + k])) p = r;
    if (p != k) for (int c = 0; c < 2 * d; ++c) std::swap(a[(size_t)p * 2 * d + c], a[(size_t)k * 2 * d + c]);
    const double inv = 1.0 / a[(size_t)k * 2 * d + k];
    for (int c = 0; c < 2 * d; ++c) a[(size_t)k * 2 * d + c] *= inv;
    for (int r = 0; r < d; ++r) {
      if (r == k) continue;
      const double f = a[(size_t)r * 2 * d + k];
      if (f != 0.0) for (int c = 0; c < 2 * d; ++c) a[(size_t)r * 2 * d + c] -= f * a[(size_t)k * 2 * d + c];
    }
  }
  for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) J[r + c * d] = a[(size_t)r * 2 * d + d + c];
}

static int pcg_solve(int nb, int d, const int* colptr, const int* rowidx, const double* values, double* x, const double* b,
                     double tolerance, int absolute_tolerance, int max_iter, double* residual_io) {
  const int n = nb * d;
  std::vector<const double*> diag(nb, nullptr), sparse;
  std::vector<std::pair<int, int>> indices;   // (row offset, column offset) of every off-diagonal upper block
  std::vector<double> J((size_t)nb * d * d);
  for (int c = 0; c < nb; ++c)
    for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
      const double* B = values + (size_t)q * d * d;
      if (rowidx[q] == c) { diag[c] = B; pcg_inverse_spd(B, d, &J[(size_t)c * d * d]); break; }
      indices.push_back({rowidx[q] * d, c * d});
      sparse.push_back(B);
    }
  auto multDiag = [&](const double* const* blocks, const double* Jflat, const double* src, double* dest) {
    for (int i = 0; i < nb; ++i) {
      const double* B = blocks ? blocks[i] : Jflat + (size_t)i * d * d;
      for (int r = 0; r < d; ++r) {
        double s = 0;
        for (int c = 0; c < d; ++c) s += B[r + c * d] * src[i * d + c];
        dest[i * d + r] = s;
      }
    }
  };
  auto mult = [&](const double* src, double* dest) {
    multDiag(diag.data(), nullptr, src, dest);
    for (size_t i = 0; i < sparse.size(); ++i) {
      const int ro = indices[i].first, co = indices[i].second;
      const double* B = sparse[i];
      for (int r = 0; r < d; ++r) { double s = 0; for (int c = 0; c < d; ++c) s += B[r + c * d] * src[co + c]; dest[ro + r] += s; }
      for (int c = 0; c < d; ++c) { double s = 0; for (int r = 0; r < d; ++r) s += B[r + c * d] * src[ro + r]; dest[co + c] += s; }
    }
  };
  auto dot = [&](const std::vector<double>& u, const std::vector<double>& v) { double s = 0; for (int i = 0; i < n; ++i) s += u[i] * v[i]; return s; };
  std::vector<double> r(b, b + n), dv(n, 0.0), q(n, 0.0), s(n, 0.0);
  for (int i = 0; i < n; ++i) x[i] = 0.0;
  multDiag(nullptr, J.data(), r.data(), dv.data());
  double dn = dot(r, dv);
  double d0 = tolerance * dn;
  if (absolute_tolerance && *residual_io > 0.0 && *residual_io > d0) d0 = *residual_io;
  const int maxIter = max_iter < 0 ? n : max_iter;
  int iteration;
  for (iteration = 0; iteration < maxIter; ++iteration) {
    if (dn <= d0) break;
    mult(dv.data(), q.data());
    const double a = dn / dot(dv, q);
    for (int i = 0; i < n; ++i) x[i] += a * dv[i];
    for (int i = 0; i < n; ++i) r[i] -= a * q[i];
    multDiag(nullptr, J.data(), r.data(), s.data());
    const double dold = dn;
    dn = dot(r, s);
    const double ba = dn / dold;
    for (int i = 0; i < n; ++i) dv[i] = s[i] + ba * dv[i];
  }
  *residual_io = 0.5 * dn;
  return iteration;
}

extern "C" {
// LinearSolverPCG on an upper-triangular block CCS matrix (blocks d*d column-major); returns the iteration count
int oracle_pcg_solve(int nb, int d, const int* colptr, const int* rowidx, const double* values, double* x, const double* b,
                     double tolerance, int absolute_tolerance, int max_iter, double* residual_io) {
  return pcg_solve(nb, d, colptr, rowidx, values, x, b, tolerance, absolute_tolerance, max_iter, residual_io);
}
// test hooks: the reference's compute_dq_dR (which = 0) and the restatement (which = 1) on one rotation (col-major 3x3)
void oracle_dq_dR(const double* R, int which, double* dq27) { if (which == 0) compute_dq_dR(dq27, R); else compute_dq_dR_restated(dq27, R); }

oracle_graph* oracle_new(void) { return new oracle_graph(); }
void oracle_free(oracle_graph* g) { delete g; }

int oracle_add_vertex(oracle_graph* g, int kind, int id, const double* payload, int n) {
  Vertex* v = add_vertex(g, kind, id);
  if (!v) return -1;
  return vertex_read(v, payload, n) ? 0 : -1;
}
int oracle_add_edge(oracle_graph* g, int kind, int id1, int id2, const double* payload, int n) {
  return add_edge(g, kind, id1, id2, payload, n);
}
int oracle_add_vertices(oracle_graph* g, int kind, int n, const int* ids, const double* payload, int stride) {
  for (int i = 0; i < n; ++i) { int rc = oracle_add_vertex(g, kind, ids[i], payload + (size_t)i * stride, stride); if (rc) return rc; }
  return 0;
}
int oracle_add_edges(oracle_graph* g, int kind, int n, const int* id1, const int* id2, const double* payload, int stride) {
  for (int i = 0; i < n; ++i) { int rc = add_edge(g, kind, id1[i], id2[i], payload + (size_t)i * stride, stride); if (rc) return rc; }
  return 0;
}
int oracle_set_fixed(oracle_graph* g, int id, int fixed) {
  Vertex* v = g->vertex(id);
  if (!v) return -1;
  v->fixed = fixed != 0;
  return 0;
}

int oracle_add_camera_parameters(oracle_graph* g, int id, double focal_length, double cx, double cy, double baseline) {
  if (g->cameraParameters.count(id)) return -1;  // ParameterContainer::addParameter refuses duplicate ids
  g->cameraParameters[id] = {focal_length, cx, cy, baseline};
  return 0;
}

// ParameterSE3Offset::read (types/slam3d/parameter_se3_offset.cpp:47-56): x y z qx qy qz qw, quaternion normalised
int oracle_add_se3_offset(oracle_graph* g, int id, const double* xyzq) {
  if (!g || !xyzq || g->se3Offsets.count(id)) return -1;
  double q[4] = {xyzq[3], xyzq[4], xyzq[5], xyzq[6]};
  quat_normalize(q);
  std::array<double, 12> o;
  quat_to_R(q, o.data());
  o[9] = xyzq[0]; o[10] = xyzq[1]; o[11] = xyzq[2];
  g->se3Offsets[id] = o;
  return 0;
}

int oracle_load(oracle_graph* g, const char* path) {
  std::ifstream is(path);
  if (!is) return -1;
  std::string line, token;
  std::vector<double> nums;
  while (std::getline(is, line)) {
    std::stringstream ss(line);
    token.clear();
    ss >> token;
    if (token.empty() || token[0] == '#') continue;
    if (token == "FIX") { int id; while (ss >> id) oracle_set_fixed(g, id, 1); continue; }
    int vkind = -1, ekind = -1;
    if (token == "VERTEX_SE2") vkind = ORC_VERTEX_SE2;
    else if (token == "VERTEX_SE3:QUAT") vkind = ORC_VERTEX_SE3;
    else if (token == "VERTEX_CAM") vkind = ORC_VERTEX_CAM;
    else if (token == "VERTEX_XYZ") vkind = ORC_VERTEX_XYZ;
    else if (token == "EDGE_SE2") ekind = ORC_EDGE_SE2;
    else if (token == "EDGE_SE3:QUAT") ekind = ORC_EDGE_SE3;
    else if (token == "EDGE_PROJECT_P2MC") ekind = ORC_EDGE_P2MC;
    else if (token == "VERTEX_SE3:EXPMAP") vkind = ORC_VERTEX_SE3_EXPMAP;
    else if (token == "VERTEX_XY") vkind = ORC_VERTEX_XY;
    else if (token == "VERTEX_TRACKXYZ") vkind = ORC_VERTEX_XYZ;
    else if (token == "EDGE_SE2_XY") ekind = ORC_EDGE_SE2_XY;
    else if (token == "EDGE_SE3_TRACKXYZ") ekind = ORC_EDGE_SE3_XYZ;
    else if (token == "PARAMS_SE3OFFSET") {  // types/slam3d/parameter_se3_offset.cpp:47-56
      int id; double o[7];
      if (ss >> id >> o[0] >> o[1] >> o[2] >> o[3] >> o[4] >> o[5] >> o[6]) oracle_add_se3_offset(g, id, o);
      continue;
    }
    else if (token == "EDGE_PROJECT_XYZ2UV:EXPMAP") ekind = ORC_EDGE_XYZ2UV;
    else if (token == "PARAMS_CAMERAPARAMETERS") {  // optimizable_graph.cpp:398-415 + CameraParameters::read
      int id; double f, cx, cy, bl;
      if (ss >> id >> f >> cx >> cy >> bl) oracle_add_camera_parameters(g, id, f, cx, cy, bl);
      continue;
    }
    else continue;  // unknown tags are skipped (optimizable_graph.cpp:417-423)
    nums.clear();
    if (vkind >= 0) {
      int id; ss >> id;
      double d; while (ss >> d) nums.push_back(d);
      oracle_add_vertex(g, vkind, id, nums.data(), (int)nums.size());
    } else {
      int id1, id2; ss >> id1 >> id2;
      double d; while (ss >> d) nums.push_back(d);
      oracle_add_edge(g, ekind, id1, id2, nums.data(), (int)nums.size());
    }
  }
  return 0;
}

int oracle_setup_cli(oracle_graph* g, int requires_marginalize) {
  if (g->vertices.empty()) return -2;
  bool gf = gauge_freedom(g);
  Vertex* gauge = find_gauge(g);
  int ret = -1;
  if (gf) {
    if (!gauge) return -2;
    gauge->fixed = true;
    ret = gauge->id;
  }
  if (requires_marginalize) {
    int maxDim = 0, minDim = 1 << 30;
    for (auto& kv : g->vertices) { maxDim = std::max(maxDim, kv.second->dim); minDim = std::min(minDim, kv.second->dim); }
    if (maxDim != minDim)
      for (auto& kv : g->vertices) if (kv.second->dim != maxDim) kv.second->marginalized = true;
  }
  return ret;
}
int oracle_initialize(oracle_graph* g) { return initialize_optimization(g) ? 0 : -1; }
// Solver::computeMarginals (core/block_solver.hpp:490-499): blocks of Hpp^-1, column-major d x d each
int oracle_compute_marginals(oracle_graph* g, int nblocks, const int* rows, const int* cols, double* out) {
  return g->linearSolver.solvePattern(g->Hpp, nblocks, rows, cols, out) ? 0 : -1;
}
void oracle_robustify(int kind, double delta, double e2, double* rho3) { robustify(kind, delta, e2, rho3); }
void oracle_robustify_restated(int kind, double delta, double e2, double* rho3) { robustify_restated(kind, delta, e2, rho3); }
// the restated SE2 algebra / 2D edge errors (what compute_error used before it called the reference's SE2 class), for the
// cross-check against oracle/_ref
void oracle_se2_restated(int what, const double* a, const double* b, const double* z, double* out) {
  const SE2 A{a[0], a[1], a[2]};
  if (what == 0) { const SE2 r = se2_mul(A, SE2{b[0], b[1], b[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.th; }
  else if (what == 1) { const SE2 r = se2_inv(A); out[0] = r.x; out[1] = r.y; out[2] = r.th; }
  else if (what == 2) {  // EdgeSE2::computeError
    const SE2 d = se2_mul(se2_inv(SE2{z[0], z[1], z[2]}), se2_mul(se2_inv(A), SE2{b[0], b[1], b[2]}));
    out[0] = d.x; out[1] = d.y; out[2] = d.th;
  } else {               // EdgeSE2PointXY::computeError
    const SE2 xi = se2_inv(A);
    const double c = cos(xi.th), s = sin(xi.th);
    out[0] = (c * b[0] - s * b[1] + xi.x) - z[0];
    out[1] = (s * b[0] + c * b[1] + xi.y) - z[1];
  }
}
// apps/g2o_cli/g2o.cpp:322-336: one kernel of the given width on every edge
int oracle_set_robust_kernel(oracle_graph* g, int kind, double delta) {
  if (kind < 0 || kind > 5 || !(delta > 0)) return -1;
  g->rkKind = kind; g->rkDelta = delta;
  for (auto& e : g->edges) { e->rkKind = kind; e->rkDelta = delta; }
  return 0;
}
// Edge::setRobustKernel on one edge (edge k in addEdge order)
int oracle_set_edge_robust_kernel(oracle_graph* g, int k, int kind, double delta) {
  if (!g || k < 0 || k >= (int)g->edges.size() || kind < 0 || kind > 5) return -1;
  g->edges[k]->rkKind = kind; g->edges[k]->rkDelta = delta;
  return 0;
}
void oracle_set_block_ordering(oracle_graph* g, int bo) { g->linearSolver.blockOrdering = bo != 0; }
// kind 0: LinearSolverCSparse, 1: LinearSolverPCG (tolerance / absolute tolerance / max iterations as its setters)
int oracle_set_linear_solver(oracle_graph* g, int kind, double tolerance, int absolute_tolerance, int max_iterations) {
  if (!g || kind < 0 || kind > 1) return -1;
  g->usePCG = kind == 1;
  g->pcgTolerance = tolerance; g->pcgAbsolute = absolute_tolerance; g->pcgMaxIter = max_iterations; g->pcgResidual = -1.0;
  return 0;
}
int oracle_pcg_iterations(oracle_graph* g) { return g ? g->pcgIterations : -1; }

int oracle_optimize(oracle_graph* g, int algorithm, int iterations, oracle_iter_stats* stats) {
  if (g->ivMap.empty()) return -1;
  if (!algorithm_init(g)) return -1;
  int cj = 0;
  int result = 1;
  bool ok = true;
  for (int i = 0; i < iterations && ok; ++i) {
    oracle_iter_stats local; memset(&local, 0, sizeof(local));
    oracle_iter_stats* st = stats ? &stats[i] : &local;
    memset(st, 0, sizeof(*st));
    st->iteration = i;
    double ts = now();
    g->linearSolver.timeSymbolic = 0;
    result = (algorithm == ORC_GN) ? solve_gn(g, i, st) : solve_lm(g, i, st);
    ok = (result == 1);
    st->time_iteration = now() - ts;   // the hot path proper (without the stats-only chi2 below)
    st->chi2 = compute_active_errors(g);
    st->result = result;
    st->lambda = g->currentLambda;
    st->levenberg_iterations = g->levenbergIterations;
    st->time_symbolic = g->linearSolver.timeSymbolic;
    st->time_numeric = g->linearSolver.timeNumeric;
    if (algorithm == ORC_GN) { st->time_schur = g->timeSchur; st->time_linear_solver = g->timeLinearSolver; }
    ++cj;
  }
  if (result == -1) return 0;
  return cj;
}

// one OptimizationAlgorithmLevenberg::solve(iteration) (bench: time iterations of a continuing optimisation)
int oracle_lm_iteration(oracle_graph* g, int iteration, oracle_iter_stats* st) {
  oracle_iter_stats local;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  st->iteration = iteration;
  double ts = now();
  int r = solve_lm(g, iteration, st);
  st->time_iteration = now() - ts;
  st->result = r;
  st->lambda = g->currentLambda;
  st->levenberg_iterations = g->levenbergIterations;
  return r;
}
// one OptimizationAlgorithm{GaussNewton,Levenberg}::solve(iteration) of a continuing optimisation, timed like
// oracle_optimize times it; the chi2 of the state after the iteration is evaluated outside the timed part (what
// SparseOptimizer::optimize does for its statistics, core/sparse_optimizer.cpp:392-397).  iteration 0 runs
// OptimizationAlgorithmWithHessian::init first (sparse_optimizer.cpp:365).
int oracle_iteration(oracle_graph* g, int algorithm, int iteration, oracle_iter_stats* st) {
  oracle_iter_stats local;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  if (g->ivMap.empty()) return -1;
  if (iteration == 0 && !algorithm_init(g)) return -1;
  st->iteration = iteration;
  g->linearSolver.timeSymbolic = 0;
  double ts = now();
  int r = (algorithm == ORC_GN) ? solve_gn(g, iteration, st) : solve_lm(g, iteration, st);
  st->time_iteration = now() - ts;
  st->chi2 = compute_active_errors(g);
  st->result = r;
  st->lambda = g->currentLambda;
  st->levenberg_iterations = g->levenbergIterations;
  st->time_symbolic = g->linearSolver.timeSymbolic;
  st->time_numeric = g->linearSolver.timeNumeric;
  return r;
}
int oracle_algorithm_init(oracle_graph* g) { return algorithm_init(g) ? 0 : -1; }
int oracle_build_structure(oracle_graph* g) { return build_structure(g) ? 0 : -1; }
double oracle_compute_active_errors(oracle_graph* g) { return compute_active_errors(g); }
int oracle_build_system(oracle_graph* g) { build_system(g); return 0; }
double oracle_lambda_init(oracle_graph* g) { return lambda_init(g); }
int oracle_set_lambda(oracle_graph* g, double lambda, int backup) { g->currentLambda = lambda; set_lambda(g, lambda, backup != 0); return 0; }
int oracle_restore_diagonal(oracle_graph* g) { restore_diagonal(g); return 0; }
int oracle_solve(oracle_graph* g) { return solve(g) ? 1 : 0; }
int oracle_update(oracle_graph* g) { update(g); return 0; }
int oracle_push(oracle_graph* g) { push(g); return 0; }
int oracle_pop(oracle_graph* g) { pop(g); return 0; }
int oracle_discard_top(oracle_graph* g) { discard_top(g); return 0; }

int oracle_dims(oracle_graph* g, int* d) {
  d[0] = g->numPoses; d[1] = g->numLandmarks; d[2] = g->sizePoses; d[3] = g->sizeLandmarks;
  d[4] = (int)g->activeEdges.size(); d[5] = (int)g->activeVertices.size(); d[6] = g->poseDim; d[7] = g->landmarkDim;
  return 0;
}
int oracle_get_b(oracle_graph* g, double* b) { memcpy(b, g->b.data(), g->b.size() * sizeof(double)); return (int)g->b.size(); }
int oracle_get_x(oracle_graph* g, double* x) { memcpy(x, g->x.data(), g->x.size() * sizeof(double)); return (int)g->x.size(); }
int oracle_set_x(oracle_graph* g, const double* x) { memcpy(g->x.data(), x, g->x.size() * sizeof(double)); return (int)g->x.size(); }
int oracle_get_errors(oracle_graph* g, double* err) {
  size_t k = 0;
  for (Edge* e : g->activeEdges) for (int i = 0; i < e->D; ++i) err[k++] = e->err[i];
  return (int)k;
}
static int canonical_estimate(const Vertex* v, double* out) {
  switch (v->kind) {
    case ORC_VERTEX_SE2: memcpy(out, v->est, 3 * sizeof(double)); return 3;
    case ORC_VERTEX_SE3: memcpy(out, v->est, 12 * sizeof(double)); return 12;
    case ORC_VERTEX_CAM: memcpy(out, v->est, 12 * sizeof(double)); return 12;
    case ORC_VERTEX_XYZ: memcpy(out, v->est, 3 * sizeof(double)); return 3;
    case ORC_VERTEX_SE3_EXPMAP: memcpy(out, v->est, 7 * sizeof(double)); return 7;
    case ORC_VERTEX_XY: memcpy(out, v->est, 2 * sizeof(double)); return 2;
  }
  return -1;
}
int oracle_get_estimate(oracle_graph* g, int id, double* out) {
  Vertex* v = g->vertex(id);
  if (!v) return -1;
  return canonical_estimate(v, out);
}
int oracle_vertex_count(oracle_graph* g) { return (int)g->vertices.size(); }
int oracle_get_vertices(oracle_graph* g, int* ids, int* kinds, int* hidx, int* flags) {
  std::vector<Vertex*> vs;
  for (auto& kv : g->vertices) vs.push_back(kv.second);
  std::sort(vs.begin(), vs.end(), [](Vertex* a, Vertex* b) { return a->id < b->id; });
  for (size_t i = 0; i < vs.size(); ++i) {
    ids[i] = vs[i]->id; kinds[i] = vs[i]->kind; hidx[i] = vs[i]->hessianIndex;
    flags[i] = (vs[i]->fixed ? 1 : 0) | (vs[i]->marginalized ? 2 : 0);
  }
  return (int)vs.size();
}
int oracle_edge_count(oracle_graph* g) { return (int)g->edges.size(); }
int oracle_get_edge(oracle_graph* g, int k, int* kind, int* id1, int* id2, double* meas, double* info) {
  if (k < 0 || k >= (int)g->edges.size()) return -1;
  Edge* e = g->edges[k].get();
  *kind = e->kind; *id1 = e->v[0]->id; *id2 = e->v[1]->id;
  int nm = e->kind == ORC_EDGE_SE2 ? 3 : e->kind == ORC_EDGE_SE3 ? 12 : e->kind == ORC_EDGE_SE3_XYZ ? 3 : 2;
  memcpy(meas, e->meas, nm * sizeof(double));
  memcpy(info, e->info, e->D * e->D * sizeof(double));
  return 0;
}
int oracle_get_blocks(oracle_graph* g, int which, int* rows, int* cols, double* values) {
  SBM* M = which == 0 ? &g->Hpp : which == 1 ? &g->Hll : which == 2 ? &g->Hpl : &g->Hschur;
  int n = 0;
  if (!M->cbase.empty()) {  // variable block sizes: every block padded with zeros to maxdim x maxdim
    const int D = M->maxdim();
    for (int c = 0; c < M->ncols; ++c)
      for (auto& kv : M->cols[c]) {
        if (rows) {
          rows[n] = kv.first; cols[n] = c;
          double* dst = values + (size_t)n * D * D;
          for (int i = 0; i < D * D; ++i) dst[i] = 0;
          const int dr = M->rd(kv.first), dc = M->cd(c);
          for (int cc = 0; cc < dc; ++cc) for (int rr = 0; rr < dr; ++rr) dst[rr + D * cc] = kv.second[rr + dr * cc];
        }
        ++n;
      }
    return n;
  }
  const int sz = M->rdim * M->cdim;
  for (int c = 0; c < M->ncols; ++c)
    for (auto& kv : M->cols[c]) {
      if (rows) { rows[n] = kv.first; cols[n] = c; memcpy(values + (size_t)n * sz, kv.second, sz * sizeof(double)); }
      ++n;
    }
  return n;
}
int oracle_get_bschur(oracle_graph* g, double* out) { memcpy(out, g->bschur.data(), g->bschur.size() * sizeof(double)); return (int)g->bschur.size(); }
int oracle_get_block_perm(oracle_graph* g, int* perm) {
  auto& P = g->linearSolver.blockPerm;
  if (perm) memcpy(perm, P.data(), P.size() * sizeof(int));
  return (int)P.size();
}
int64_t oracle_get_lnz(oracle_graph* g) { return g->linearSolver.S ? (int64_t)g->linearSolver.S->lnz : -1; }

int oracle_cs_amd(int n, const int* colptr, const int* rowidx, int* perm) {
  cs aux{};
  aux.nzmax = colptr[n]; aux.m = aux.n = n; aux.p = const_cast<int*>(colptr); aux.i = const_cast<int*>(rowidx);
  aux.x = nullptr; aux.nz = -1;
  int* P = cs_amd(1, &aux);
  if (!P) return -1;
  memcpy(perm, P, n * sizeof(int));
  cs_free(P);
  return 0;
}
int64_t oracle_scalar_amd_lnz(oracle_graph* g) {
  SBM& M = g->doSchur ? g->Hschur : g->Hpp;
  LinearSolverCSparseO ls;
  ls.fill(M, false);
  css* S = cs_schol(1, &ls.A);
  if (!S) return -1;
  int64_t r = (int64_t)S->lnz;
  cs_sfree(S);
  return r;
}

}  // extern "C"
