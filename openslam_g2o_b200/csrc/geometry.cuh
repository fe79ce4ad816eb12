// geometry.cuh - per-edge / per-vertex math of the three configured g2o types, as device functions.
//
// Everything is FP64, column-major, and follows the reference's formulas as coded (not the textbook
// forms); each function cites the file:line it computes the same quantity as.  Functions are
// __host__ __device__ only so that tests/ can exercise the very same code on the CPU against the oracle
// when no GPU is present - the product itself only ever calls them from kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define G2O_HD __host__ __device__ __forceinline__
#else
#define G2O_HD inline
#endif

namespace g2o_b200 {
namespace geo {

constexpr double kPi = 3.14159265358979323846;

// stuff/misc.h:94-106
G2O_HD double normalize_theta(double theta) {
  if (theta >= -kPi && theta < kPi) return theta;
  double multiplier = floor(theta / (2 * kPi));
  theta = theta - multiplier * 2 * kPi;
  if (theta >= kPi) theta -= 2 * kPi;
  if (theta < -kPi) theta += 2 * kPi;
  return theta;
}

// ------------------------------------------------------------------ small dense helpers
template <int R, int K, int C>
G2O_HD void mm(const double* A, const double* B, double* Cm) {  // C = A(RxK) B(KxC)
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[r + k * R] * B[k + c * K];
      Cm[r + c * R] = s;
    }
}
template <int R, int K, int C>
G2O_HD void mtm(const double* A, const double* B, double* Cm) {  // C = A^T(A is KxR) B(KxC)
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[k + r * K] * B[k + c * K];
      Cm[r + c * R] = s;
    }
}

// Eigen Quaterniond::toRotationMatrix; q = (x y z w); R col-major
G2O_HD void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
  R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
  R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// ------------------------------------------------------------------ SE2 (types/slam2d)
struct SE2 { double x, y, th; };
G2O_HD SE2 se2_mul(const SE2& a, const SE2& b) {  // se2.h:66-78
  SE2 r;
  double s, c;
  sincos(a.th, &s, &c);
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}
G2O_HD SE2 se2_inv(const SE2& a) {  // se2.h:86-96
  SE2 r;
  r.th = normalize_theta(-a.th);
  double s, c;
  sincos(r.th, &s, &c);
  double tx = -a.x, ty = -a.y;
  r.x = c * tx - s * ty;
  r.y = s * tx + c * ty;
  return r;
}
// EdgeSE2::computeError (edge_se2.h:46-52); zi = cached inverse measurement
G2O_HD void se2_error(const SE2& xi, const SE2& xj, const SE2& zi, double* e) {
  SE2 d = se2_mul(zi, se2_mul(se2_inv(xi), xj));
  e[0] = d.x; e[1] = d.y; e[2] = d.th;
}
// EdgeSE2::linearizeOplus (edge_se2.cpp:76-99): A = d e/d xi, B = d e/d xj (3x3 col-major)
G2O_HD void se2_jacobians(const SE2& xi, const SE2& xj, const SE2& zi, double* A, double* B) {
  const double dtx = xj.x - xi.x, dty = xj.y - xi.y;
  double si, ci, sz, cz;
  sincos(xi.th, &si, &ci);
  sincos(zi.th, &sz, &cz);
  // A0, B0 before the left multiplication with diag(R(zi.th), 1)
  const double a00 = -ci, a01 = -si, a02 = -si * dtx + ci * dty;
  const double a10 = si, a11 = -ci, a12 = -ci * dtx - si * dty;
  const double b00 = ci, b01 = si, b10 = -si, b11 = ci;
  A[0] = cz * a00 - sz * a10; A[3] = cz * a01 - sz * a11; A[6] = cz * a02 - sz * a12;
  A[1] = sz * a00 + cz * a10; A[4] = sz * a01 + cz * a11; A[7] = sz * a02 + cz * a12;
  A[2] = 0; A[5] = 0; A[8] = -1;
  B[0] = cz * b00 - sz * b10; B[3] = cz * b01 - sz * b11; B[6] = 0;
  B[1] = sz * b00 + cz * b10; B[4] = sz * b01 + cz * b11; B[7] = 0;
  B[2] = 0; B[5] = 0; B[8] = 1;
}
G2O_HD void se2_oplus(double* est, const double* u) {  // vertex_se2.h:51-58
  est[0] += u[0];
  est[1] += u[1];
  est[2] = normalize_theta(est[2] + u[2]);
}

// ------------------------------------------------------------------ SE3 as Isometry3d [R col-major | t]
struct Iso { double R[9]; double t[3]; };
G2O_HD Iso iso_inverse(const Iso& a) {
  Iso r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.R[i + 3 * j] = a.R[j + 3 * i];
#pragma unroll
  for (int i = 0; i < 3; ++i) r.t[i] = -(r.R[i] * a.t[0] + r.R[i + 3] * a.t[1] + r.R[i + 6] * a.t[2]);
  return r;
}
G2O_HD Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  mm<3, 3, 3>(a.R, b.R, r.R);
#pragma unroll
  for (int i = 0; i < 3; ++i) r.t[i] = (a.R[i] * b.t[0] + a.R[i + 3] * b.t[1] + a.R[i + 6] * b.t[2]) + a.t[i];
  return r;
}
// internal::toVectorMQT (isometry3d_mappings.cpp:38-44,77-83,93-99): [t ; xyz of normalised quaternion, w>=0]
G2O_HD void iso_to_vector_mqt(const Iso& d, double* v) {
  const double* R = d.R;
  double q[4];
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[5] - R[7]) * t;
    q[1] = (R[6] - R[2]) * t;
    q[2] = (R[1] - R[3]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i + 3 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i + 3 * i] - R[j + 3 * j] - R[k + 3 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k + 3 * j] - R[j + 3 * k]) * t;
    q[j] = (R[j + 3 * i] + R[i + 3 * j]) * t;
    q[k] = (R[k + 3 * i] + R[i + 3 * k]) * t;
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double sgn = (q[3] / n < 0) ? -1.0 : 1.0;
  v[0] = d.t[0]; v[1] = d.t[1]; v[2] = d.t[2];
  v[3] = sgn * (q[0] / n); v[4] = sgn * (q[1] / n); v[5] = sgn * (q[2] / n);
}
// EdgeSE3::computeError (edge_se3.cpp:48-53); Zi = cached inverse measurement
G2O_HD void se3_error(const Iso& Xi, const Iso& Xj, const Iso& Zi, double* e) {
  Iso d = iso_mul(iso_mul(Zi, iso_inverse(Xi)), Xj);
  iso_to_vector_mqt(d, e);
}
// compute_dq_dR (dquat2mat.cpp:9-59, dquat2mat_maxima_generated.cpp:1-165); dq 3x9 col-major
G2O_HD void dq_dR(const double* R, double* dq) {
  const double r00 = R[0], r10 = R[1], r20 = R[2], r01 = R[3], r11 = R[4], r21 = R[5], r02 = R[6], r12 = R[7], r22 = R[8];
#pragma unroll
  for (int i = 0; i < 27; ++i) dq[i] = 0;
#define DQ(r, c) dq[(r) + 3 * (c)]
  double S, qw;
  const double tr = r00 + r11 + r22;
  if (tr > 0) {
    S = sqrt(tr + 1.0) * 2;
    qw = 0.25 * S;
    S *= .25;
    const double a3 = 1 / S, a1 = 1 / (S * S * S), a2 = -0.03125 * (r21 - r12) * a1, a4 = 0.25 * a3, a5 = -0.25 * a3,
                 a6 = 0.03125 * (r20 - r02) * a1, a7 = -0.03125 * (r10 - r01) * a1;
    DQ(0, 0) = a2; DQ(0, 4) = a2; DQ(0, 5) = a4; DQ(0, 7) = a5; DQ(0, 8) = a2;
    DQ(1, 0) = a6; DQ(1, 2) = a5; DQ(1, 4) = a6; DQ(1, 6) = a4; DQ(1, 8) = a6;
    DQ(2, 0) = a7; DQ(2, 1) = a4; DQ(2, 3) = a5; DQ(2, 4) = a7; DQ(2, 8) = a7;
  } else if ((r00 > r11) & (r00 > r22)) {
    S = sqrt(1.0 + r00 - r11 - r22) * 2;
    qw = (r21 - r12) / S;
    S *= .25;
    const double a1 = 1 / S, a2 = -0.125 * a1, a3 = 1 / (S * S * S), a4 = r10 + r01, a5 = 0.25 * a1, a6 = 0.03125 * a3 * a4,
                 a7 = r20 + r02, a8 = 0.03125 * a3 * a7;
    DQ(0, 0) = 0.125 * a1; DQ(0, 4) = a2; DQ(0, 8) = a2;
    DQ(1, 0) = -0.03125 * a3 * a4; DQ(1, 1) = a5; DQ(1, 3) = a5; DQ(1, 4) = a6; DQ(1, 8) = a6;
    DQ(2, 0) = -0.03125 * a3 * a7; DQ(2, 2) = a5; DQ(2, 4) = a8; DQ(2, 6) = a5; DQ(2, 8) = a8;
  } else if (r11 > r22) {
    S = sqrt(1.0 + r11 - r00 - r22) * 2;
    qw = (r02 - r20) / S;
    S *= .25;
    const double a1 = 1 / (S * S * S), a2 = r10 + r01, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4, a6 = -0.125 * a4,
                 a7 = r21 + r12, a8 = 0.03125 * a1 * a7;
    DQ(0, 0) = a3; DQ(0, 1) = a5; DQ(0, 3) = a5; DQ(0, 4) = -0.03125 * a1 * a2; DQ(0, 8) = a3;
    DQ(1, 0) = a6; DQ(1, 4) = 0.125 * a4; DQ(1, 8) = a6;
    DQ(2, 0) = a8; DQ(2, 4) = -0.03125 * a1 * a7; DQ(2, 5) = a5; DQ(2, 7) = a5; DQ(2, 8) = a8;
  } else {
    S = sqrt(1.0 + r22 - r00 - r11) * 2;
    qw = (r10 - r01) / S;
    S *= .25;
    const double a1 = 1 / (S * S * S), a2 = r20 + r02, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4, a6 = r21 + r12,
                 a7 = 0.03125 * a1 * a6, a8 = -0.125 * a4;
    DQ(0, 0) = a3; DQ(0, 2) = a5; DQ(0, 4) = a3; DQ(0, 6) = a5; DQ(0, 8) = -0.03125 * a1 * a2;
    DQ(1, 0) = a7; DQ(1, 4) = a7; DQ(1, 5) = a5; DQ(1, 7) = a5; DQ(1, 8) = -0.03125 * a1 * a6;
    DQ(2, 0) = a8; DQ(2, 4) = a8; DQ(2, 8) = 0.125 * a4;
  }
#undef DQ
  if (qw <= 0) {
#pragma unroll
    for (int i = 0; i < 27; ++i) dq[i] = -dq[i];
  }
}
// computeEdgeSE3Gradient (isometry3d_gradients.h:194-265): Ji, Jj 6x6 col-major; Z = measurement (NOT inverse).
// Also returns the error of the same linearisation point in e (toVectorMQT(E)).
G2O_HD void se3_jacobians(const Iso& Xi, const Iso& Xj, const Iso& Zinv, double* Ji, double* Jj, double* e) {
  const Iso& A = Zinv;
  const Iso B = iso_mul(iso_inverse(Xi), Xj);
  const Iso E = iso_mul(A, B);
  iso_to_vector_mqt(E, e);
  const double* Ra = A.R; const double* Rb = B.R; const double* Re = E.R; const double* tb = B.t;
  double dq[27];
  dq_dR(Re, dq);
#pragma unroll
  for (int i = 0; i < 36; ++i) { Ji[i] = 0; Jj[i] = 0; }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) { Ji[r + 6 * c] = -Ra[r + 3 * c]; Jj[r + 6 * c] = Re[r + 3 * c]; }
  {  // dte/dqi = Ra * skewT(tb)
    const double x = 2 * tb[0], y = 2 * tb[1], z = 2 * tb[2];
    const double S[9] = {0, z, -y, -z, 0, x, y, -x, 0};
    double T[9];
    mm<3, 3, 3>(Ra, S, T);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) Ji[r + 6 * (3 + c)] = T[r + 3 * c];
  }
  double M[27];
  {  // dre/dqi
    const double r11 = 2 * Rb[0], r12 = 2 * Rb[3], r13 = 2 * Rb[6], r21 = 2 * Rb[1], r22 = 2 * Rb[4], r23 = 2 * Rb[7],
                 r31 = 2 * Rb[2], r32 = 2 * Rb[5], r33 = 2 * Rb[8];
    const double Sxt[9] = {0, r31, -r21, 0, r32, -r22, 0, r33, -r23};
    const double Syt[9] = {-r31, 0, r11, -r32, 0, r12, -r33, 0, r13};
    const double Szt[9] = {r21, -r11, 0, r22, -r12, 0, r23, -r13, 0};
    mm<3, 3, 3>(Ra, Sxt, M); mm<3, 3, 3>(Ra, Syt, M + 9); mm<3, 3, 3>(Ra, Szt, M + 18);
    double Q[9];
    mm<3, 9, 3>(dq, M, Q);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) Ji[3 + r + 6 * (3 + c)] = Q[r + 3 * c];
  }
  {  // dre/dqj : Re * skew(.., I)
    const double Sx[9] = {0, 0, 0, 0, 0, 2, 0, -2, 0};
    const double Sy[9] = {0, 0, -2, 0, 0, 0, 2, 0, 0};
    const double Sz[9] = {0, 2, 0, -2, 0, 0, 0, 0, 0};
    mm<3, 3, 3>(Re, Sx, M); mm<3, 3, 3>(Re, Sy, M + 9); mm<3, 3, 3>(Re, Sz, M + 18);
    double Q[9];
    mm<3, 9, 3>(dq, M, Q);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) Jj[3 + r + 6 * (3 + c)] = Q[r + 3 * c];
  }
}
// VertexSE3::oplusImpl (vertex_se3.h:107-116) + fromCompactQuaternion (isometry3d_mappings.cpp:84-91)
G2O_HD void se3_oplus(double* est, const double* u, bool orthogonalize) {
  Iso inc;
  double w = 1 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  if (w < 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) inc.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    w = sqrt(w);
    const double q[4] = {u[3], u[4], u[5], w};
    quat_to_R(q, inc.R);
  }
  inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
  Iso cur;
#pragma unroll
  for (int i = 0; i < 9; ++i) cur.R[i] = est[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) cur.t[i] = est[9 + i];
  Iso r = iso_mul(cur, inc);
  if (orthogonalize) {  // approximateNearestOrthogonalMatrix (isometry3d_mappings.h:86-91)
    double E[9], RE[9];
    mtm<3, 3, 3>(r.R, r.R, E);
    E[0] -= 1; E[4] -= 1; E[8] -= 1;
    mm<3, 3, 3>(r.R, E, RE);
#pragma unroll
    for (int i = 0; i < 9; ++i) r.R[i] -= 0.5 * RE[i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) est[i] = r.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) est[9 + i] = r.t[i];
}

// ------------------------------------------------------------------ SBA camera (types/sba)
// derived block per camera: w2n (3x4 col-major, 12) | fx fy cx cy
// SBACam::setTransform/transformW2F (sbacam.h:120-130,155)
G2O_HD void cam_derive(const double* est /* t3 q4 fx fy cx cy b */, double* der /*16*/) {
  double R[9];
  quat_to_R(est + 3, R);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) der[r + 3 * c] = R[c + 3 * r];
#pragma unroll
  for (int r = 0; r < 3; ++r) der[r + 9] = -(der[r] * est[0] + der[r + 3] * est[1] + der[r + 6] * est[2]);
  der[12] = est[7]; der[13] = est[8]; der[14] = est[9]; der[15] = est[10];
}
// EdgeProjectP2MC::computeError (types_sba.h:170-192): e = proj(w2i [X;1]) - z, w2i = K w2n (sbacam.h:159)
G2O_HD void p2mc_error(const double* der, const double* X, const double* z, double* e) {
  double pn[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) pn[r] = der[r] * X[0] + der[r + 3] * X[1] + der[r + 6] * X[2] + der[r + 9];
  // w2i row0 = fx*w2n row0 + cx*w2n row2, row1 = fy*row1 + cy*row2, row2 = row2
  const double p0 = der[12] * pn[0] + der[14] * pn[2];
  const double p1 = der[13] * pn[1] + der[15] * pn[2];
  e[0] = p0 / pn[2] - z[0];
  e[1] = p1 / pn[2] - z[1];
}
// EdgeProjectP2MC::linearizeOplus (types_sba.cpp:334-403): Jp 2x3 (point), Jc 2x6 (camera), col-major.
// cam_t = camera translation (est[0..3)).  dRdx/y/z = dRid{x,y,z} * w2n(:,0:3) (sbacam.h:162-181).
G2O_HD void p2mc_jacobians(const double* der, const double* cam_t, const double* X, double* Jp, double* Jc) {
  double pc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) pc[r] = der[r] * X[0] + der[r + 3] * X[1] + der[r + 6] * X[2] + der[r + 9];
  const double px = pc[0], py = pc[1], pz = pc[2];
  const double ipz2 = 1.0 / (pz * pz);
  const double ipz2fx = ipz2 * der[12], ipz2fy = ipz2 * der[13];
  const double pw[3] = {X[0] - cam_t[0], X[1] - cam_t[1], X[2] - cam_t[2]};
  // rows of W = w2n(:,0:3) applied to pw
  double Wp[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) Wp[r] = der[r] * pw[0] + der[r + 3] * pw[1] + der[r + 6] * pw[2];
  // dRidx*W*pw = (0, 2*Wp2, -2*Wp1); dRidy*W*pw = (-2*Wp2, 0, 2*Wp0); dRidz*W*pw = (2*Wp1, -2*Wp0, 0)
  const double dqx[3] = {0.0, 2 * Wp[2], -2 * Wp[1]};
  const double dqy[3] = {-2 * Wp[2], 0.0, 2 * Wp[0]};
  const double dqz[3] = {2 * Wp[1], -2 * Wp[0], 0.0};
#define SETCOL(J, c, d0, d1, d2)                       \
  J[0 + 2 * (c)] = (pz * (d0) - px * (d2)) * ipz2fx;   \
  J[1 + 2 * (c)] = (pz * (d1) - py * (d2)) * ipz2fy;
  SETCOL(Jc, 3, dqx[0], dqx[1], dqx[2])
  SETCOL(Jc, 4, dqy[0], dqy[1], dqy[2])
  SETCOL(Jc, 5, dqz[0], dqz[1], dqz[2])
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    SETCOL(Jc, k, -der[0 + 3 * k], -der[1 + 3 * k], -der[2 + 3 * k])
    SETCOL(Jp, k, der[0 + 3 * k], der[1 + 3 * k], der[2 + 3 * k])
  }
#undef SETCOL
}
// SBACam::update (sbacam.h:101-117) through VertexCam::oplusImpl (types_sba.h:93-100)
G2O_HD void cam_oplus(double* est, const double* u) {
  est[0] += u[0]; est[1] += u[1]; est[2] += u[2];
  const double bx = u[3], by = u[4], bz = u[5];
  const double bw = sqrt(1.0 - (bx * bx + by * by + bz * bz));
  const double ax = est[3], ay = est[4], az = est[5], aw = est[6];
  double rw = aw * bw - ax * bx - ay * by - az * bz;
  double rx = aw * bx + ax * bw + ay * bz - az * by;
  double ry = aw * by + ay * bw + az * bx - ax * bz;
  double rz = aw * bz + az * bw + ax * by - ay * bx;
  const double n = sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
  est[3] = rx / n; est[4] = ry / n; est[5] = rz / n; est[6] = rw / n;
}

// ------------------------------------------------------------------ expmap camera (types/sba/types_six_dof_expmap, SE3Quat)
// estimate per pose: t3 | q(xyzw)4 (world -> camera, SE3Quat) | f f cx cy baseline (the CameraParameters of its edges)
// derived block: [R(q) | t] (3x4 col-major, 12) | f f cx cy - the same shape as the SBACam block (w2n | K), so both
// camera models run through the same kernels.
G2O_HD void expmap_derive(const double* est, double* der /*16*/) {
  quat_to_R(est + 3, der);
  der[9] = est[0]; der[10] = est[1]; der[11] = est[2];
  der[12] = est[7]; der[13] = est[8]; der[14] = est[9]; der[15] = est[10];
}
// EdgeProjectXYZ2UV::computeError (types_six_dof_expmap.h:143-150): e = obs - cam_map(T.map(X)),
// cam_map (types_six_dof_expmap.cpp:65-71): proj * focal_length + principle_point
G2O_HD void xyz2uv_error(const double* der, const double* X, const double* z, double* e) {
  double p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = der[r] * X[0] + der[r + 3] * X[1] + der[r + 6] * X[2] + der[r + 9];
  e[0] = z[0] - (p[0] / p[2] * der[12] + der[14]);
  e[1] = z[1] - (p[1] / p[2] * der[13] + der[15]);
}
// EdgeProjectXYZ2UV::linearizeOplus (types_six_dof_expmap.cpp:288-326): Jp 2x3 (point), Jc 2x6 (pose: omega, upsilon)
G2O_HD void xyz2uv_jacobians(const double* der, const double* X, double* Jp, double* Jc) {
  double p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = der[r] * X[0] + der[r + 3] * X[1] + der[r + 6] * X[2] + der[r + 9];
  const double x = p[0], y = p[1], z = p[2], z_2 = z * z, f = der[12];
  const double t02 = -x / z * f, t12 = -y / z * f, miz = -1. / z;
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // -1/z * tmp * R,  tmp = [f 0 -x/z f ; 0 f -y/z f]
    Jp[0 + 2 * c] = miz * (f * der[0 + 3 * c] + t02 * der[2 + 3 * c]);
    Jp[1 + 2 * c] = miz * (f * der[1 + 3 * c] + t12 * der[2 + 3 * c]);
  }
  Jc[0] = x * y / z_2 * f;          Jc[1] = (1 + y * y / z_2) * f;
  Jc[2] = -(1 + (x * x / z_2)) * f; Jc[3] = -x * y / z_2 * f;
  Jc[4] = y / z * f;                Jc[5] = -x / z * f;
  Jc[6] = miz * f;                  Jc[7] = 0;
  Jc[8] = 0;                        Jc[9] = miz * f;
  Jc[10] = x / z_2 * f;             Jc[11] = y / z_2 * f;
}
// Eigen Quaterniond(Matrix3d) (same branches as iso_to_vector_mqt above, without the normalisation)
G2O_HD void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[5] - R[7]) * t;
    q[1] = (R[6] - R[2]) * t;
    q[2] = (R[1] - R[3]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i + 3 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i + 3 * i] - R[j + 3 * j] - R[k + 3 * k] + 1.0);
    double qq[4];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    qq[3] = (R[k + 3 * j] - R[j + 3 * k]) * t;
    qq[j] = (R[j + 3 * i] + R[i + 3 * j]) * t;
    qq[k] = (R[k + 3 * i] + R[i + 3 * k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
  }
}
// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:101-104): estimate <- SE3Quat::exp(update) * estimate;
// exp: se3quat.h:216-252 (update = [omega ; upsilon]); product + normalizeRotation: se3quat.h:103-109, 280-285
G2O_HD void expmap_oplus(double* est, const double* u) {
  const double ox = u[0], oy = u[1], oz = u[2];
  const double theta = sqrt(ox * ox + oy * oy + oz * oz);
  const double Om[9] = {0, oz, -oy, -oz, 0, ox, oy, -ox, 0};  // skew(omega), col-major
  double Om2[9], R[9], V[9];
  mm<3, 3, 3>(Om, Om, Om2);
  double a, b, c;
  if (theta < 0.00001) { a = 1; b = 1; c = 1; }
  else {
    const double sn = sin(theta), cs = cos(theta);
    a = sn / theta; b = (1 - cs) / (theta * theta); c = (theta - sn) / (theta * theta * theta);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * Om[i] + b * Om2[i];
    V[i] = (theta < 0.00001) ? R[i] : I + b * Om[i] + c * Om2[i];
  }
  double qe[4], te[3];
  R_to_quat(R, qe);
#pragma unroll
  for (int r = 0; r < 3; ++r) te[r] = V[r] * u[3] + V[r + 3] * u[4] + V[r + 6] * u[5];
  if (qe[3] < 0) { qe[0] = -qe[0]; qe[1] = -qe[1]; qe[2] = -qe[2]; qe[3] = -qe[3]; }
  double n = sqrt(qe[0] * qe[0] + qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3]);
  qe[0] /= n; qe[1] /= n; qe[2] /= n; qe[3] /= n;
  // t <- te + qe * t (Eigen quaternion * vector: v + w uv + qv x uv, uv = 2 qv x v)
  const double vx = est[0], vy = est[1], vz = est[2];
  double ux = qe[1] * vz - qe[2] * vy, uy = qe[2] * vx - qe[0] * vz, uz = qe[0] * vy - qe[1] * vx;
  ux += ux; uy += uy; uz += uz;
  est[0] = te[0] + (vx + qe[3] * ux + (qe[1] * uz - qe[2] * uy));
  est[1] = te[1] + (vy + qe[3] * uy + (qe[2] * ux - qe[0] * uz));
  est[2] = te[2] + (vz + qe[3] * uz + (qe[0] * uy - qe[1] * ux));
  // q <- qe * q, normalizeRotation
  const double bx = est[3], by = est[4], bz = est[5], bw = est[6];
  double rw = qe[3] * bw - qe[0] * bx - qe[1] * by - qe[2] * bz;
  double rx = qe[3] * bx + qe[0] * bw + qe[1] * bz - qe[2] * by;
  double ry = qe[3] * by + qe[1] * bw + qe[2] * bx - qe[0] * bz;
  double rz = qe[3] * bz + qe[2] * bw + qe[0] * by - qe[1] * bx;
  if (rw < 0) { rx = -rx; ry = -ry; rz = -rz; rw = -rw; }
  n = sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
  est[3] = rx / n; est[4] = ry / n; est[5] = rz / n; est[6] = rw / n;
}
// the two camera models behind one call (MODEL 0: SBACam / P2MC, 1: SE3 expmap / XYZ2UV)
template <int MODEL> G2O_HD void ba_derive(const double* est, double* der) { if (MODEL == 0) cam_derive(est, der); else expmap_derive(est, der); }
template <int MODEL> G2O_HD void ba_error(const double* der, const double* X, const double* z, double* e) {
  if (MODEL == 0) p2mc_error(der, X, z, e); else xyz2uv_error(der, X, z, e);
}
template <int MODEL> G2O_HD void ba_jacobians(const double* der, const double* cam_t, const double* X, double* Jp, double* Jc) {
  if (MODEL == 0) p2mc_jacobians(der, cam_t, X, Jp, Jc); else xyz2uv_jacobians(der, X, Jp, Jc);
}
template <int MODEL> G2O_HD void ba_oplus(double* est, const double* u) { if (MODEL == 0) cam_oplus(est, u); else expmap_oplus(est, u); }

// ------------------------------------------------------------------ landmark SLAM edges
// EdgeSE2PointXY::computeError (types/slam2d/edge_se2_pointxy.h:46-51): e = x^-1 * l - z, SE2 * Vector2d = R(th) v + t
G2O_HD void se2_xy_error(const SE2& x, const double* l, const double* z, double* e) {
  const SE2 xi = se2_inv(x);
  double s, c;
  sincos(xi.th, &s, &c);
  e[0] = (c * l[0] - s * l[1] + xi.x) - z[0];
  e[1] = (s * l[0] + c * l[1] + xi.y) - z[1];
}
// EdgeSE2PointXY::linearizeOplus (types/slam2d/edge_se2_pointxy.cpp:67-95): A 2x3 (pose), B 2x2 (point), col-major
G2O_HD void se2_xy_jacobians(const SE2& x, const double* l, double* A, double* B) {
  const double x1 = x.x, y1 = x.y, x2 = l[0], y2 = l[1];
  double aux_3, aux_1;
  sincos(x.th, &aux_3, &aux_1);
  const double aux_2 = -aux_1;
  A[0] = aux_2; A[2] = -aux_3; A[4] = aux_1 * y2 - aux_1 * y1 - aux_3 * x2 + aux_3 * x1;
  A[1] = aux_3; A[3] = aux_2;  A[5] = -aux_3 * y2 + aux_3 * y1 - aux_1 * x2 + aux_1 * x1;
  B[0] = aux_1; B[2] = aux_3;
  B[1] = -aux_3; B[3] = aux_1;
}
// EdgeSE3PointXYZ::computeError (types/slam3d/edge_se3_pointxyz.cpp:98-108): e = w2n * l - z with the cache
// w2n = (X * offset)^-1 (parameter_se3_offset.cpp:75-80)
G2O_HD void se3_xyz_error(const Iso& X, const Iso& offset, const double* l, const double* z, double* e) {
  const Iso w2n = iso_inverse(iso_mul(X, offset));
#pragma unroll
  for (int r = 0; r < 3; ++r) e[r] = (w2n.R[r] * l[0] + w2n.R[r + 3] * l[1] + w2n.R[r + 6] * l[2] + w2n.t[r]) - z[r];
}
// EdgeSE3PointXYZ::linearizeOplus (types/slam3d/edge_se3_pointxyz.cpp:110-135): J = [-I | 2 [Zcam]x^T-pattern | R(w2l)],
// Zcam = w2l * l, w2l = X^-1; Jhom = R(offset^-1) * J; A = Jhom(:, 0:6) 3x6, B = Jhom(:, 6:9) 3x3, col-major
G2O_HD void se3_xyz_jacobians(const Iso& X, const Iso& offset, const double* l, double* A, double* B) {
  const Iso w2l = iso_inverse(X);
  double Z[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) Z[r] = w2l.R[r] * l[0] + w2l.R[r + 3] * l[1] + w2l.R[r + 6] * l[2] + w2l.t[r];
  double J[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) J[i] = 0.0;
  J[0] = -1.0; J[4] = -1.0; J[8] = -1.0;
  J[0 + 3 * 4] = -2 * Z[2]; J[0 + 3 * 5] = 2 * Z[1];
  J[1 + 3 * 3] = 2 * Z[2];  J[1 + 3 * 5] = -2 * Z[0];
  J[2 + 3 * 3] = -2 * Z[1]; J[2 + 3 * 4] = 2 * Z[0];
#pragma unroll
  for (int i = 0; i < 9; ++i) J[18 + i] = w2l.R[i];
  const Iso oi = iso_inverse(offset);
  double Jh[27];
  mm<3, 3, 9>(oi.R, J, Jh);
#pragma unroll
  for (int i = 0; i < 18; ++i) A[i] = Jh[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) B[i] = Jh[18 + i];
}

// Eigen fixed-size 3x3 inverse (cofactors / determinant), used at block_solver.hpp:389
G2O_HD void inverse3(const double* m, double* r) {
#define M_(i, j) m[(i) + 3 * (j)]
#define COF(i, j) (M_(((i) + 1) % 3, ((j) + 1) % 3) * M_(((i) + 2) % 3, ((j) + 2) % 3) - M_(((i) + 1) % 3, ((j) + 2) % 3) * M_(((i) + 2) % 3, ((j) + 1) % 3))
  const double c0 = COF(0, 0), c1 = COF(1, 0), c2 = COF(2, 0);
  const double det = c0 * M_(0, 0) + c1 * M_(1, 0) + c2 * M_(2, 0);
  const double invdet = 1.0 / det;
  r[0] = c0 * invdet; r[3] = c1 * invdet; r[6] = c2 * invdet;
  r[1] = COF(0, 1) * invdet; r[4] = COF(1, 1) * invdet; r[7] = COF(2, 1) * invdet;
  r[2] = COF(0, 2) * invdet; r[5] = COF(1, 2) * invdet; r[8] = COF(2, 2) * invdet;
#undef COF
#undef M_
}

}  // namespace geo
}  // namespace g2o_b200
