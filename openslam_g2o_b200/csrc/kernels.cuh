// kernels.cuh - sm_100a kernels of the per-iteration hot path except the Cholesky (chol.cu):
//   errors + chi2            (SparseOptimizer::computeActiveErrors/activeRobustChi2, core/sparse_optimizer.cpp:61-114)
//   linearize + accumulate   (BlockSolver::buildSystem, core/block_solver.hpp:501-560;
//                             BaseBinaryEdge::constructQuadraticForm, core/base_binary_edge.hpp:54-120)
//   Schur complement         (BlockSolver::solve, core/block_solver.hpp:367-439)
//   landmark back-substitution (core/block_solver.hpp:461-481)
//   oplus update             (SparseOptimizer::update, core/sparse_optimizer.cpp:421-434)
//   LM scalars               (computeLambdaInit / computeScale, core/optimization_algorithm_levenberg.cpp:149-172)
//
// Accumulation is by ordered gather (no atomics): every Hessian block / b segment is written by one
// thread group that sums its contributions in ascending edge order, so results are run-to-run
// bit-identical and follow the reference's summation order.
#pragma once
#include <cuda_runtime.h>

#include "geometry.cuh"

namespace g2o_b200 {
namespace k {

using namespace geo;

// ------------------------------------------------------------------ deterministic reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block sum for blockDim.x <= 1024, result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) r += sh[i];
  }
  return r;
}
// out[0] = sum(partials[0..n)) in a fixed order; single block
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = block_sum(s);
  if (threadIdx.x == 0) out[0] = s;
}

// ------------------------------------------------------------------ pose graphs (SE2 / SE3)
// edge arrays are SoA: meas[f*E + e], info[f*E + e] (upper triangle, row-major order)
template <int D>
__device__ __forceinline__ void load_info(const double* __restrict__ info, int E, int e, double* W /*DxD full*/) {
  int f = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      const double v = info[(long long)f * E + e];
      W[i + D * j] = v;
      W[j + D * i] = v;
      ++f;
    }
}
template <int D>
__device__ __forceinline__ double chi2_of(const double* W, const double* e) {
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) t += W[r + D * c] * e[c];
    s += e[r] * t;
  }
  return s;
}

__device__ __forceinline__ SE2 load_se2(const double* __restrict__ est, int v) {
  const double4 q = *reinterpret_cast<const double4*>(est + 4ll * v);
  return SE2{q.x, q.y, q.z};
}
__device__ __forceinline__ Iso load_iso(const double* __restrict__ est, int v) {
  Iso X;
  const double2* p = reinterpret_cast<const double2*>(est + 12ll * v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { double2 a = p[i]; X.R[2 * i] = a.x; X.R[2 * i + 1] = a.y; }
  { double2 a = p[4]; X.R[8] = a.x; X.t[0] = a.y; }
  { double2 a = p[5]; X.t[1] = a.x; X.t[2] = a.y; }
  return X;
}
__device__ __forceinline__ Iso load_iso_soa(const double* __restrict__ meas, int E, int e) {
  Iso Z;
#pragma unroll
  for (int i = 0; i < 9; ++i) Z.R[i] = meas[(long long)i * E + e];
#pragma unroll
  for (int i = 0; i < 3; ++i) Z.t[i] = meas[(long long)(9 + i) * E + e];
  return Z;
}

// ------------------------------------------------------------------ robust kernels
// One kernel for every edge (what `g2o -robustKernel NAME -robustKernelWidth W` sets up, apps/g2o_cli/g2o.cpp:322-336).
// rho(e2) and rho'(e2) as core/robust_kernel_impl.cpp:65-126; the quadratic form uses the first-order weight only:
// information and omega_r are scaled by rho' (base_edge.h:96-102, base_binary_edge.hpp:91-113), chi2 sums rho
// (sparse_optimizer.cpp:100-114).
// struct Robust { int kind; double delta; kinds, deltas }: common.h.  kinds != nullptr: per-edge kernels
// (OptimizableGraph::Edge::setRobustKernel on individual edges), indexed by the device edge index
__device__ __forceinline__ Robust robust_at(const Robust& rk, int e) {
  if (rk.kinds == nullptr) return rk;
  return Robust{(int)rk.kinds[e], rk.deltas[e], nullptr, nullptr};
}
__device__ __forceinline__ void robustify(const Robust& rk, double e2, double& rho0, double& rho1) {
  const double dsqr = rk.delta * rk.delta;
  switch (rk.kind) {
    case 1:
      if (e2 <= dsqr) { rho0 = e2; rho1 = 1.0; }
      else { const double sq = sqrt(e2); rho0 = 2 * sq * rk.delta - dsqr; rho1 = rk.delta / sq; }
      break;
    case 2: {
      const double aux1 = (1.0 / dsqr) * e2 + 1.0, aux2 = sqrt(aux1);
      rho0 = 2 * dsqr * (aux2 - 1); rho1 = 1.0 / aux2;
      break;
    }
    case 3: {
      const double aux = (1.0 / dsqr) * e2 + 1.0;
      rho0 = dsqr * log(aux); rho1 = 1.0 / aux;
      break;
    }
    case 4:
      if (e2 <= dsqr) { rho0 = e2; rho1 = 1.0; } else { rho0 = dsqr; rho1 = 0.0; }
      break;
    case 5: {
      double scale = (2.0 * rk.delta) / (rk.delta + e2);
      if (scale >= 1.0) scale = 1.0;
      rho0 = scale * e2 * scale; rho1 = scale * scale;
      break;
    }
    default: rho0 = e2; rho1 = 1.0;
  }
}

// KIND 0 = SE2 (D=3), 1 = SE3 (D=6)
template <int KIND>
__global__ void pg_chi2_kernel(int E, const int* __restrict__ v0, const int* __restrict__ v1,
                               const double* __restrict__ est, const double* __restrict__ meas,
                               const double* __restrict__ info, Robust rk, double* __restrict__ partials) {
  constexpr int D = KIND == 0 ? 3 : 6;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (e < E) {
    double err[D], W[D * D];
    if (KIND == 0) {
      const SE2 zi{meas[e], meas[(long long)E + e], meas[2ll * E + e]};
      se2_error(load_se2(est, v0[e]), load_se2(est, v1[e]), zi, err);
    } else {
      se3_error(load_iso(est, v0[e]), load_iso(est, v1[e]), load_iso_soa(meas, E, e), err);
    }
    load_info<D>(info, E, e, W);
    chi = chi2_of<D>(W, err);
    const Robust rke = robust_at(rk, e);
    if (rke.kind) { double r1; robustify(rke, chi, chi, r1); }
  }
  chi = block_sum(chi);
  if (threadIdx.x == 0) partials[blockIdx.x] = chi;
}

// per edge staging record: [Hii D*D | Hjj D*D | Hij D*D (destination orientation) | bi D | bj D]
template <int KIND>
__global__ void __launch_bounds__(128)
pg_linearize_kernel(int E, const int* __restrict__ v0, const int* __restrict__ v1, const double* __restrict__ est,
                    const double* __restrict__ meas, const double* __restrict__ info,
                    const unsigned char* __restrict__ transposed, Robust rk, double* __restrict__ stage) {
  constexpr int D = KIND == 0 ? 3 : 6;
  constexpr int STRIDE = 3 * D * D + 2 * D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double A[D * D], B[D * D], err[D];
  if (KIND == 0) {
    const SE2 zi{meas[e], meas[(long long)E + e], meas[2ll * E + e]};
    const SE2 xi = load_se2(est, v0[e]), xj = load_se2(est, v1[e]);
    se2_error(xi, xj, zi, err);
    se2_jacobians(xi, xj, zi, A, B);
  } else {
    se3_jacobians(load_iso(est, v0[e]), load_iso(est, v1[e]), load_iso_soa(meas, E, e), A, B, err);
  }
  double W[D * D];
  load_info<D>(info, E, e, W);
  const Robust rke = robust_at(rk, e);
    if (rke.kind) {  // weightedOmega = rho' * information; omega_r = -weightedOmega * error
    double r0, r1;
    robustify(rke, chi2_of<D>(W, err), r0, r1);
#pragma unroll
    for (int i = 0; i < D * D; ++i) W[i] *= r1;
  }
  double omega_r[D];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) s += W[r + D * c] * err[c];
    omega_r[r] = -s;
  }
  double* out = stage + (long long)e * STRIDE;
  double T[D * D];  // AtO, then BtO
  mtm<D, D, D>(A, W, T);
  {
    double H[D * D];
    mm<D, D, D>(T, A, H);
#pragma unroll
    for (int i = 0; i < D * D; ++i) out[i] = H[i];
    mm<D, D, D>(T, B, H);  // AtO * B = block (i,j)
    if (transposed[e]) {
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) out[2 * D * D + c + D * r] = H[r + D * c];
    } else {
#pragma unroll
      for (int i = 0; i < D * D; ++i) out[2 * D * D + i] = H[i];
    }
  }
  mtm<D, D, D>(B, W, T);
  {
    double H[D * D];
    mm<D, D, D>(T, B, H);
#pragma unroll
    for (int i = 0; i < D * D; ++i) out[D * D + i] = H[i];
  }
  double bi[D], bj[D];
  mtm<D, D, 1>(A, omega_r, bi);
  mtm<D, D, 1>(B, omega_r, bj);
#pragma unroll
  for (int i = 0; i < D; ++i) { out[3 * D * D + i] = bi[i]; out[3 * D * D + D + i] = bj[i]; }
}

// ordered gather: dst[seg*LEN + k] = sum over sources s of stage[(id/5)*STRIDE + field_off[id%5] + k]
// source ids of a segment are ascending in edge order.  LEN = D*D (Hessian blocks) or D (b).
template <int D, int LEN>
__global__ void gather_segments_kernel(int nseg, const int* __restrict__ src_ptr, const int* __restrict__ src_id,
                                       const double* __restrict__ stage, double* __restrict__ dst) {
  constexpr int STRIDE = 3 * D * D + 2 * D;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nseg * LEN) return;
  const int seg = (int)(idx / LEN), kk = (int)(idx - (long long)seg * LEN);
  double s = 0.0;
  for (int p = src_ptr[seg]; p < src_ptr[seg + 1]; ++p) {
    const int id = src_id[p];
    const int e = id / 5, f = id - e * 5;
    const int off = f < 3 ? f * D * D : 3 * D * D + (f - 3) * D;
    s += stage[(long long)e * STRIDE + off + kk];
  }
  dst[idx] = s;
}

// ------------------------------------------------------------------ landmark SLAM (pose + point, nothing marginalized)
// FAM 0: EdgeSE2PointXY (pose SE2, D = 3; point 2; error 2), FAM 1: EdgeSE3PointXYZ (pose SE3, D = 6; point 3; error 3).
// The point's Hessian blocks are padded to the pose dimension D (g2o_b200.h), so these edges write the same staging
// record as the pose-pose edges and the ordered gather above sums both kinds.  v0 = pose vertex, v1 = point vertex.
struct SensorOffset { double m[12]; };   // ParameterSE3Offset::offset() as [R col-major | t]
template <int FAM> struct PlDims { static constexpr int D = FAM == 0 ? 3 : 6, LD = FAM == 0 ? 2 : 3, ED = FAM == 0 ? 2 : 3; };

template <int FAM>
__device__ __forceinline__ void pl_error(int e, int E, const int* __restrict__ v0, const int* __restrict__ v1,
                                         const double* __restrict__ pose_est, const double* __restrict__ lm_est,
                                         const double* __restrict__ meas, const SensorOffset& off, double* err,
                                         double* A, double* B, bool jac) {
  const double4 lq = *reinterpret_cast<const double4*>(lm_est + 4ll * v1[e]);
  const double l[3] = {lq.x, lq.y, lq.z};
  if (FAM == 0) {
    const double z[2] = {meas[e], meas[(long long)E + e]};
    const SE2 x = load_se2(pose_est, v0[e]);
    se2_xy_error(x, l, z, err);
    if (jac) se2_xy_jacobians(x, l, A, B);
  } else {
    const double z[3] = {meas[e], meas[(long long)E + e], meas[2ll * E + e]};
    const Iso X = load_iso(pose_est, v0[e]);
    Iso O;
#pragma unroll
    for (int i = 0; i < 9; ++i) O.R[i] = off.m[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) O.t[i] = off.m[9 + i];
    se3_xyz_error(X, O, l, z, err);
    if (jac) se3_xyz_jacobians(X, O, l, A, B);
  }
}

template <int FAM>
__global__ void pl_chi2_kernel(int E, const int* __restrict__ v0, const int* __restrict__ v1,
                               const double* __restrict__ pose_est, const double* __restrict__ lm_est,
                               const double* __restrict__ meas, const double* __restrict__ info, SensorOffset off,
                               Robust rk, double* __restrict__ partials) {
  constexpr int ED = PlDims<FAM>::ED;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (e < E) {
    double err[ED], W[ED * ED];
    pl_error<FAM>(e, E, v0, v1, pose_est, lm_est, meas, off, err, nullptr, nullptr, false);
    load_info<ED>(info, E, e, W);
    chi = chi2_of<ED>(W, err);
    const Robust rke = robust_at(rk, e);
    if (rke.kind) { double r1; robustify(rke, chi, chi, r1); }
  }
  chi = block_sum(chi);
  if (threadIdx.x == 0) partials[blockIdx.x] = chi;
}

// staging record of edge rec0 + e: [Hii D*D | Hjj D*D (top-left LD x LD) | Hij D*D in destination orientation:
// D x LD in the first LD columns, transposed: LD x D in the first LD rows | bi D | bj D (first LD)]; padding = 0
template <int FAM>
__global__ void __launch_bounds__(128)
pl_linearize_kernel(int E, int rec0, const int* __restrict__ v0, const int* __restrict__ v1,
                    const double* __restrict__ pose_est, const double* __restrict__ lm_est,
                    const double* __restrict__ meas, const double* __restrict__ info,
                    const unsigned char* __restrict__ transposed, SensorOffset off, Robust rk,
                    double* __restrict__ stage) {
  constexpr int D = PlDims<FAM>::D, LD = PlDims<FAM>::LD, ED = PlDims<FAM>::ED;
  constexpr int STRIDE = 3 * D * D + 2 * D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double A[ED * D], B[ED * LD], err[ED], W[ED * ED];
  pl_error<FAM>(e, E, v0, v1, pose_est, lm_est, meas, off, err, A, B, true);
  load_info<ED>(info, E, e, W);
  const Robust rke = robust_at(rk, e);
    if (rke.kind) {
    double r0, r1;
    robustify(rke, chi2_of<ED>(W, err), r0, r1);
#pragma unroll
    for (int i = 0; i < ED * ED; ++i) W[i] *= r1;
  }
  double omega_r[ED];
#pragma unroll
  for (int r = 0; r < ED; ++r) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < ED; ++c) s += W[r + ED * c] * err[c];
    omega_r[r] = -s;
  }
  double* out = stage + (long long)(rec0 + e) * STRIDE;
#pragma unroll 4
  for (int i = 0; i < STRIDE; ++i) out[i] = 0.0;
  double AtO[D * ED], BtO[LD * ED];
  mtm<D, ED, ED>(A, W, AtO);
  mtm<LD, ED, ED>(B, W, BtO);
  {
    double H[D * D];
    mm<D, ED, D>(AtO, A, H);
#pragma unroll
    for (int i = 0; i < D * D; ++i) out[i] = H[i];
  }
  {
    double H[LD * LD];
    mm<LD, ED, LD>(BtO, B, H);
#pragma unroll
    for (int c = 0; c < LD; ++c)
#pragma unroll
      for (int r = 0; r < LD; ++r) out[D * D + r + D * c] = H[r + LD * c];
  }
  {
    double H[D * LD];
    mm<D, ED, LD>(AtO, B, H);   // block (pose, point)
    const bool tr = transposed[rec0 + e];
#pragma unroll
    for (int c = 0; c < LD; ++c)
#pragma unroll
      for (int r = 0; r < D; ++r) out[2 * D * D + (tr ? c + D * r : r + D * c)] = H[r + D * c];
  }
  double bi[D], bj[LD];
  mtm<D, ED, 1>(A, omega_r, bi);
  mtm<LD, ED, 1>(B, omega_r, bj);
#pragma unroll
  for (int i = 0; i < D; ++i) out[3 * D * D + i] = bi[i];
#pragma unroll
  for (int i = 0; i < LD; ++i) out[3 * D * D + D + i] = bj[i];
}

// VertexPointXY / VertexPointXYZ::oplusImpl (types/slam2d/vertex_point_xy.h:76-80, types/slam3d/vertex_pointxyz.h:49-52)
// with the point's unknowns at x[hidx * D .. + LD)
template <int D, int LD>
__global__ void oplus_point_kernel(int n, const int* __restrict__ hidx, const double* __restrict__ x, double* __restrict__ est) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int h = hidx[v];
  if (h < 0) return;
#pragma unroll
  for (int k = 0; k < LD; ++k) est[4ll * v + k] += x[(long long)D * h + k];
}

// ------------------------------------------------------------------ bundle adjustment (camera + XYZ)
// MODEL 0: VertexCam / EdgeProjectP2MC (types_sba), MODEL 1: VertexSE3Expmap / EdgeProjectXYZ2UV (types_six_dof_expmap)
template <int MODEL>
__global__ void cam_derive_kernel(int n, const double* __restrict__ est, double* __restrict__ der) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double e[12], d[16];
#pragma unroll
  for (int k = 0; k < 12; ++k) e[k] = est[12ll * i + k];
  ba_derive<MODEL>(e, d);
#pragma unroll
  for (int k = 0; k < 16; ++k) der[16ll * i + k] = d[k];
}

__device__ __forceinline__ void load_der(const double* __restrict__ der, int c, double* d) {
  const double2* p = reinterpret_cast<const double2*>(der + 16ll * c);
#pragma unroll
  for (int i = 0; i < 8; ++i) { double2 a = __ldg(p + i); d[2 * i] = a.x; d[2 * i + 1] = a.y; }
}

// edges [e_begin, E) (e_begin > 0: the observations of fixed points, which the fused back-substitution kernel does not visit)
template <int MODEL>
__global__ void ba_chi2_kernel(int E, const int* __restrict__ e_pt, const int* __restrict__ e_cam,
                               const double* __restrict__ pt_est, const double* __restrict__ cam_der,
                               const double* __restrict__ meas, const double* __restrict__ info, Robust rk,
                               double* __restrict__ partials, int e_begin) {
  const int e = e_begin + blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (e < E) {
    double der[16];
    load_der(cam_der, e_cam[e], der);
    const double4 X4 = *reinterpret_cast<const double4*>(pt_est + 4ll * e_pt[e]);
    const double X[3] = {X4.x, X4.y, X4.z};
    const double z[2] = {meas[e], meas[(long long)E + e]};
    double err[2];
    ba_error<MODEL>(der, X, z, err);
    const double w0 = info[e], w1 = info[(long long)E + e], w2 = info[2ll * E + e];
    chi = err[0] * (w0 * err[0] + w1 * err[1]) + err[1] * (w1 * err[0] + w2 * err[1]);
    const Robust rke = robust_at(rk, e);
    if (rke.kind) { double r1; robustify(rke, chi, chi, r1); }
  }
  chi = block_sum(chi);
  if (threadIdx.x == 0) partials[blockIdx.x] = chi;
}

// one thread per landmark: Hll, b_l and one Hpl block per observation (edges of a landmark are contiguous)
template <int MODEL>
__global__ void __launch_bounds__(128)
ba_linearize_points_kernel(int nl, const int* __restrict__ lm_eptr, const int* __restrict__ lm_order,
                           const int* __restrict__ lm_vertex,
                           const int* __restrict__ e_cam, const int* __restrict__ e_hpl,
                           const unsigned char* __restrict__ e_first, const double* __restrict__ pt_est,
                           const double* __restrict__ cam_est, const double* __restrict__ cam_der,
                           const double* __restrict__ meas, const double* __restrict__ info, int E, Robust rk,
                           double* __restrict__ Hll, double* __restrict__ Hpl, double* __restrict__ b_l) {
  const int rank = blockIdx.x * blockDim.x + threadIdx.x;  // landmark rank: edges and Hpl slots are in this order
  if (rank >= nl) return;
  const int l = lm_order[rank];
  const double4 X4 = *reinterpret_cast<const double4*>(pt_est + 4ll * lm_vertex[l]);
  const double X[3] = {X4.x, X4.y, X4.z};
  double H[9], bl[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = 0.0;
  bl[0] = bl[1] = bl[2] = 0.0;
  for (int e = lm_eptr[rank]; e < lm_eptr[rank + 1]; ++e) {
    const int c = e_cam[e];
    double der[16];
    load_der(cam_der, c, der);
    const double ct[3] = {cam_est[12ll * c], cam_est[12ll * c + 1], cam_est[12ll * c + 2]};
    double Jp[6], Jc[12], err[2];
    ba_jacobians<MODEL>(der, ct, X, Jp, Jc);
    const double z[2] = {meas[e], meas[(long long)E + e]};
    ba_error<MODEL>(der, X, z, err);
    double w0 = info[e], w1 = info[(long long)E + e], w2 = info[2ll * E + e];
    const Robust rke = robust_at(rk, e);
    if (rke.kind) {
      double r0, r1;
      robustify(rke, err[0] * (w0 * err[0] + w1 * err[1]) + err[1] * (w1 * err[0] + w2 * err[1]), r0, r1);
      w0 *= r1; w1 *= r1; w2 *= r1;
    }
    // JpW = Jp^T W (3x2), omega_r = -W err
    double JpW[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      JpW[k] = Jp[2 * k] * w0 + Jp[2 * k + 1] * w1;
      JpW[k + 3] = Jp[2 * k] * w1 + Jp[2 * k + 1] * w2;
    }
    const double or0 = -(w0 * err[0] + w1 * err[1]), or1 = -(w1 * err[0] + w2 * err[1]);
#pragma unroll
    for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
      for (int r = 0; r < 3; ++r) H[r + 3 * c2] += JpW[r] * Jp[2 * c2] + JpW[r + 3] * Jp[2 * c2 + 1];
#pragma unroll
    for (int r = 0; r < 3; ++r) bl[r] += Jp[2 * r] * or0 + Jp[2 * r + 1] * or1;
    const int slot = e_hpl[e];
    if (slot >= 0) {  // Hpl(cam, l) (6x3) += Jc^T W Jp   (the transposed write of base_binary_edge.hpp:81-82)
      double2* dst = reinterpret_cast<double2*>(Hpl + 18ll * slot);  // 144-byte blocks: 16-byte vector stores
      const bool first = e_first[e] != 0;
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
        for (int r = 0; r < 6; r += 2) {
          double2 v;
          v.x = Jc[2 * r] * JpW[c2] + Jc[2 * r + 1] * JpW[c2 + 3];
          v.y = Jc[2 * r + 2] * JpW[c2] + Jc[2 * r + 3] * JpW[c2 + 3];
          if (!first) { const double2 o = dst[(r + 6 * c2) >> 1]; v.x += o.x; v.y += o.y; }
          dst[(r + 6 * c2) >> 1] = v;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) Hll[9ll * l + i] = H[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) b_l[3ll * l + i] = bl[i];
}

// The same pass with one LANE per observation (round 2).  A warp owns a "packet": consecutive landmark ranks whose
// observations number at most 32 (a landmark with more gets a packet of its own and is walked in chunks of 32).  Edge
// arrays are read coalesced (the per-landmark kernel above reads them with a stride of the landmark's degree), all 32
// lanes do Jacobian work, the 6x3 Hpl blocks of a chunk - consecutive slots - are staged in shared memory and leave as
// ONE contiguous run of full 16-byte-per-lane stores, and every landmark's Hll / b_l is summed by one lane in ascending
// edge order from the staged per-edge terms (fixed order: run-to-run bit-identical).
constexpr int kLinWarps = 4;
constexpr int kLinWarpDoubles = 32 * 18 + 32 * 12;   // Hpl blocks | per-edge Hll, b_l terms
template <int MODEL, int MINB>   // MINB CTAs per SM: the register budget (4: 128, 5: 96, 6: 80 registers per thread)
__global__ void __launch_bounds__(32 * kLinWarps, MINB)
ba_linearize_packets_kernel(int npk, const int4* __restrict__ packets, const int* __restrict__ lm_eptr,
                            const int* __restrict__ lm_order, const int* __restrict__ e_pt,
                            const int* __restrict__ e_cam, const int* __restrict__ e_hpl,
                            const unsigned char* __restrict__ e_first, const double* __restrict__ pt_est,
                            const double* __restrict__ cam_est, const double* __restrict__ cam_der,
                            const double* __restrict__ meas, const double* __restrict__ info, int E, Robust rk,
                            double* __restrict__ Hll, double* __restrict__ Hpl, double* __restrict__ b_l) {
  __shared__ __align__(16) double sm[kLinWarps * kLinWarpDoubles];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * kLinWarps + wid;
  if (p >= npk) return;   // whole warps leave: only __syncwarp below
  double* blk = sm + wid * kLinWarpDoubles;   // [32][18]
  double* hb = blk + 32 * 18;                 // [32][12]
  // {first rank, #ranks (<= 32), first edge, end edge}: one load, so that the edge arrays can be requested right away
  // (the kernel is bound by the latency of its dependent loads, not by bandwidth: measured, long-scoreboard stalls)
  const int4 pk = __ldg(packets + p);
  const int r0 = pk.x, nr = pk.y, eb = pk.z, ee = pk.w;
  // the landmark this lane sums for
  const bool owner = lane < nr;
  const int own_a = owner ? lm_eptr[r0 + lane] : 0, own_b = owner ? lm_eptr[r0 + lane + 1] : 0;
  const int own_l = owner ? lm_order[r0 + lane] : 0;
  double H[9], bl[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = 0.0;
  bl[0] = bl[1] = bl[2] = 0.0;
  for (int e0 = eb; e0 < ee; e0 += 32) {
    const int e = e0 + lane;
    const bool active = e < ee;
    const int slot = active ? e_hpl[e] : -1;
    const bool first = active && e_first[e] != 0;
    if (active) {
      const double4 X4 = *reinterpret_cast<const double4*>(pt_est + 4ll * e_pt[e]);
      const double X[3] = {X4.x, X4.y, X4.z};
      const int c = e_cam[e];
      double der[16];
      load_der(cam_der, c, der);
      const double ct[3] = {cam_est[12ll * c], cam_est[12ll * c + 1], cam_est[12ll * c + 2]};
      double Jp[6], Jc[12], err[2];
      ba_jacobians<MODEL>(der, ct, X, Jp, Jc);
      const double z[2] = {meas[e], meas[(long long)E + e]};
      ba_error<MODEL>(der, X, z, err);
      double w0 = info[e], w1 = info[(long long)E + e], w2 = info[2ll * E + e];
      const Robust rke = robust_at(rk, e);
    if (rke.kind) {
        double q0, q1;
        robustify(rke, err[0] * (w0 * err[0] + w1 * err[1]) + err[1] * (w1 * err[0] + w2 * err[1]), q0, q1);
        w0 *= q1; w1 *= q1; w2 *= q1;
      }
      double JpW[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        JpW[k] = Jp[2 * k] * w0 + Jp[2 * k + 1] * w1;
        JpW[k + 3] = Jp[2 * k] * w1 + Jp[2 * k + 1] * w2;
      }
      const double or0 = -(w0 * err[0] + w1 * err[1]), or1 = -(w1 * err[0] + w2 * err[1]);
      double* t = hb + lane * 12;
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
        for (int r = 0; r < 3; ++r) t[r + 3 * c2] = JpW[r] * Jp[2 * c2] + JpW[r + 3] * Jp[2 * c2 + 1];
#pragma unroll
      for (int r = 0; r < 3; ++r) t[9 + r] = Jp[2 * r] * or0 + Jp[2 * r + 1] * or1;
      if (slot >= 0) {  // Hpl(cam, l) (6x3) = Jc^T W Jp into this lane's staging row
        double2* d = reinterpret_cast<double2*>(blk + 18 * lane);
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
          for (int r = 0; r < 6; r += 2)
            d[(r + 6 * c2) >> 1] = make_double2(Jc[2 * r] * JpW[c2] + Jc[2 * r + 1] * JpW[c2 + 3],
                                                Jc[2 * r + 2] * JpW[c2] + Jc[2 * r + 3] * JpW[c2 + 3]);
      }
    }
    __syncwarp();   // the staging rows are visible to the whole warp
    // Hpl: the slots opened in this chunk are consecutive (slots are numbered in edge order)
    const unsigned open_mask = __ballot_sync(0xffffffffu, slot >= 0 && first);
    const int sb = __shfl_sync(0xffffffffu, slot, open_mask ? __ffs(open_mask) - 1 : 0);
    // duplicate observations (same landmark, same camera) add to the block their run opened, in edge order: the row of
    // the nearest opener below them in this chunk, or - opened in an earlier chunk - the block in global memory
    unsigned dup = __ballot_sync(0xffffffffu, slot >= 0 && !first);
    while (dup) {
      const int d = __ffs(dup) - 1;
      if (lane == d) {
        const unsigned below = open_mask & ((1u << d) - 1u);
        double* dst = below ? blk + 18 * (31 - __clz(below)) : Hpl + 18ll * slot;
        const double* src = blk + 18 * d;
#pragma unroll
        for (int i = 0; i < 18; ++i) dst[i] += src[i];
      }
      __syncwarp();
      dup &= dup - 1;
    }
    {
      // the opened blocks leave as one contiguous run: element i of the staging area = double2 (i % 9) of row i / 9
      double2* dst = reinterpret_cast<double2*>(Hpl + 18ll * sb);
      const double2* src = reinterpret_cast<const double2*>(blk);
#pragma unroll
      for (int it = 0; it < 9; ++it) {
        const int i = lane + 32 * it, row = i / 9, j = i - 9 * row;
        if (open_mask >> row & 1u) dst[9 * __popc(open_mask & ((1u << row) - 1u)) + j] = src[i];
      }
    }
    // Hll, b_l: the owner adds its landmark's terms of this chunk in ascending edge order
    if (owner) {
      const int a = max(own_a, e0), b = min(own_b, e0 + 32);
      for (int q = a; q < b; ++q) {
        const double* t = hb + (q - e0) * 12;
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] += t[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) bl[i] += t[9 + i];
      }
    }
    __syncwarp();   // the staging buffers are reused by the next chunk; orders the global Hpl stores for later duplicates
  }
  if (owner) {
    const int l = own_l;
#pragma unroll
    for (int i = 0; i < 9; ++i) Hll[9ll * l + i] = H[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) b_l[3ll * l + i] = bl[i];
  }
}

// one CTA per free camera: Hpp(i,i) and b_i as an ordered tree-sum over the camera's observations
template <int MODEL, int MINB>   // MINB CTAs per SM the register allocation is held to
__global__ void __launch_bounds__(128, MINB)
ba_linearize_cams_kernel(const int* __restrict__ cam_eptr, const int* __restrict__ cam_eidx,
                         const int* __restrict__ pose_vertex, const int* __restrict__ e_pt,
                         const double* __restrict__ pt_est, const double* __restrict__ cam_est,
                         const double* __restrict__ cam_der, const double* __restrict__ meas,
                         const double* __restrict__ info, int E, Robust rk, const int* __restrict__ hpp_diag_block,
                         double* __restrict__ Hpp, double* __restrict__ b_p) {
  const int i = blockIdx.x;
  const int c = pose_vertex[i];
  double der[16];
  load_der(cam_der, c, der);
  const double ct[3] = {cam_est[12ll * c], cam_est[12ll * c + 1], cam_est[12ll * c + 2]};
  double acc[27];  // 21 upper-triangular entries of Jc^T W Jc (column-wise) + 6 of b
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.0;
  for (int p = cam_eptr[i] + threadIdx.x; p < cam_eptr[i + 1]; p += blockDim.x) {
    const int e = cam_eidx[p];
    const double4 X4 = *reinterpret_cast<const double4*>(pt_est + 4ll * e_pt[e]);
    const double X[3] = {X4.x, X4.y, X4.z};
    double Jp[6], Jc[12], err[2];
    ba_jacobians<MODEL>(der, ct, X, Jp, Jc);
    const double z[2] = {meas[e], meas[(long long)E + e]};
    ba_error<MODEL>(der, X, z, err);
    double w0 = info[e], w1 = info[(long long)E + e], w2 = info[2ll * E + e];
    const Robust rke = robust_at(rk, e);
    if (rke.kind) {
      double r0, r1;
      robustify(rke, err[0] * (w0 * err[0] + w1 * err[1]) + err[1] * (w1 * err[0] + w2 * err[1]), r0, r1);
      w0 *= r1; w1 *= r1; w2 *= r1;
    }
    double JW[12];  // Jc^T W : 6x2
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      JW[k] = Jc[2 * k] * w0 + Jc[2 * k + 1] * w1;
      JW[k + 6] = Jc[2 * k] * w1 + Jc[2 * k + 1] * w2;
    }
    const double or0 = -(w0 * err[0] + w1 * err[1]), or1 = -(w1 * err[0] + w2 * err[1]);
    int q = 0;
#pragma unroll
    for (int cc = 0; cc < 6; ++cc)
#pragma unroll
      for (int r = 0; r <= cc; ++r) acc[q++] += JW[r] * Jc[2 * cc] + JW[r + 6] * Jc[2 * cc + 1];
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[21 + r] += Jc[2 * r] * or0 + Jc[2 * r + 1] * or1;
  }
  __shared__ double sh[4][27];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) sh[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    const int nw = blockDim.x >> 5;
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += sh[w][threadIdx.x];
    sh[0][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    const int r = threadIdx.x % 6, cc = threadIdx.x / 6;
    const int lo = r < cc ? r : cc, hi = r < cc ? cc : r;
    Hpp[36ll * hpp_diag_block[i] + threadIdx.x] = sh[0][hi * (hi + 1) / 2 + lo];
  } else if (threadIdx.x < 42) {
    b_p[6ll * i + (threadIdx.x - 36)] = sh[0][21 + threadIdx.x - 36];
  }
}

// ------------------------------------------------------------------ Schur complement
constexpr int kDinvStride = 10;  // 9 doubles padded to 80 B: rows stay 16-byte aligned for double2 loads

// S1: per landmark Dinv = (Hll + lambda I)^-1 (back-substitution) and, for the Schur reduction, the triangular factor
// W with Dinv = W W^T (W = R^-1, R = upper Cholesky factor of Hll + lambda I) and u = W^T b_l:
//   Hpl(i1,l) Dinv Hpl(i2,l)^T = (Hpl(i1,l) W)(Hpl(i2,l) W)^T,   Hpl(i1,l) Dinv b_l = (Hpl(i1,l) W) u
// so the reduction multiplies ONE transformed 6x3 block per observation instead of two blocks and a 3x3 inverse
// (block_solver.hpp:381-395 computes Dinv and Hpl*Dinv per landmark).
// Wu per landmark: w00 w01 w02 w11 w12 w22 u0 u1 u2 (+1 pad: 80 B, 16-byte aligned rows)
__global__ void schur_landmark_inverse_kernel(int nl, const double* __restrict__ Hll, const double* __restrict__ b_l,
                                              const double* __restrict__ lambda, double* __restrict__ Dinv,
                                              double* __restrict__ Wu) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nl) return;
  const double lam = *lambda;
  double D[9], Di[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) D[i] = Hll[9ll * l + i];
  D[0] += lam; D[4] += lam; D[8] += lam;
  inverse3(D, Di);
  const double b0 = b_l[3ll * l], b1 = b_l[3ll * l + 1], b2 = b_l[3ll * l + 2];
#pragma unroll
  for (int i = 0; i < 9; ++i) Dinv[kDinvStride * (long long)l + i] = Di[i];
  // R^T R = D (column-major symmetric: D[r + 3c])
  const double r00 = sqrt(D[0]);
  const double w00 = 1.0 / r00;
  const double r01 = D[3] * w00, r02 = D[6] * w00;
  const double r11 = sqrt(D[4] - r01 * r01);
  const double w11 = 1.0 / r11;
  const double r12 = (D[7] - r01 * r02) * w11;
  const double r22 = sqrt(D[8] - r02 * r02 - r12 * r12);
  const double w22 = 1.0 / r22;
  const double w01 = -r01 * w11 * w00;
  const double w12 = -r12 * w22 * w11;
  const double w02 = -(r01 * w12 + r02 * w22) * w00;
  double* o = Wu + kDinvStride * (long long)l;
  o[0] = w00; o[1] = w01; o[2] = w02; o[3] = w11; o[4] = w12; o[5] = w22;
  o[6] = w00 * b0;
  o[7] = w01 * b0 + w11 * b1;
  o[8] = w02 * b0 + w12 * b1 + w22 * b2;
}

// S2: Hschur(i1,i2) = hpp_scale*Hpp(i1,i2) + [i1==i2] lambda_scale*lambda I - sum_l (Hpl(i1,l) Dinv_l) Hpl(i2,l)^T
//     bschur_i1     = hpp_scale*b_i1 - sum_l Hpl(i1,l) db_l                      (block_solver.hpp:397-439)
// Both scales are 1 on a single GPU.  With landmark sharding every rank adds its PARTIAL Hpp / b_p (its own edges'
// share of the camera blocks) and rank 0 alone the lambda term: the sum over the ranks - the one all-reduce of
// [Hschur | bschur] - is the reduced system, no separate reduction of Hpp.
//
// Two kernels, no atomics, fixed summation order:
//  schur_range_kernel   one CTA per RANGE of landmarks.  Hpl slots are numbered so that a range is one contiguous
//                       piece of Hpl (read from HBM exactly once, coalesced, into shared memory together with the
//                       landmarks' Dinv / db and the range's contribution indices); landmarks are ordered by their
//                       camera lists, so a range feeds few distinct Hschur blocks.  The contributions of a range to
//                       one block form a SEGMENT (<= kSrSegMax products); 4 lanes share a segment, gather their
//                       operands from shared memory, and leave one 6x6 (+6) partial sum per segment.
//  schur_finish_kernel  one thread per entry of a block: adds the block's partial sums in segment order and
//                       subtracts them from the Hpp term.
constexpr int kSrThreads = 128;   // 32 groups of 4 lanes; 3 CTAs per SM (shared memory and registers)
constexpr int kSrLanes = 4;
constexpr int kSrSegMax = 64;     // products per segment
constexpr int kSrPartial = 42;    // 36 block entries + 6 right-hand-side entries (diagonal blocks)
constexpr int kSrMaxDynSmem = 226 * 1024;  // opt-in dynamic shared memory of schur_range_kernel (device maximum)

struct SchurRanges {
  const int* slot0;    // 2 nr: [first Hpl slot, end slot) of a range (a wide landmark may sit between two ranges)
  const int* lm_ptr;   // nr+1: into lm_ids / lm_slot
  const int* lm_ids;   // landmark (Wu index) of every landmark of the range
  const int* lm_slot;  // first Hpl slot of that landmark, relative to the range (+ one end entry per range)
  const int* seg_ptr;  // nr+1: segments of the range
  const int* seg_t;    // nseg: destination block
  const int *seg_cb, *seg_ce;  // nseg: contributions of the segment [cb, ce); a range starts at a multiple of 8
  const unsigned short *sc_a, *sc_b, *sc_l;  // per contribution: Hpl slots / landmark, relative to the range
  const unsigned char* t_diag;        // per block: 1 = diagonal
  int cap_slots, cap_lms, cap_contrib;  // shared-memory capacities; only the contribution indices may exceed theirs
                                        // (one landmark seen by > 50 cameras): they are then read from global memory
};

__global__ void __launch_bounds__(kSrThreads, 3)
schur_range_kernel(SchurRanges R, const double* __restrict__ Hpl, const double* __restrict__ Wu,
                   double* __restrict__ partial) {
  extern __shared__ __align__(16) double sr_sm[];
  double* sH = sr_sm;                                    // transformed blocks Hpl(i,l) W_l | u-products
  double* sW = sH + (size_t)R.cap_slots * 18;            // Wu rows of the range's landmarks
  int* sS = reinterpret_cast<int*>(sW + (size_t)R.cap_lms * kDinvStride);  // first slot per landmark (+ end)
  unsigned short* sA = reinterpret_cast<unsigned short*>(sS + ((R.cap_lms + 1 + 3) & ~3));
  unsigned short* sB = sA + R.cap_contrib;
  unsigned short* sL = sB + R.cap_contrib;
  unsigned short* sLm = sL + R.cap_contrib;              // landmark (range-local) of every slot
  __shared__ __align__(8) unsigned long long sr_bar;
  const int r = blockIdx.x, tid = threadIdx.x;
  const int s0 = R.slot0[2 * r], ns = R.slot0[2 * r + 1] - s0;
  const int l0 = R.lm_ptr[r], nlm = R.lm_ptr[r + 1] - l0;
  const int g0 = R.seg_ptr[r], g1 = R.seg_ptr[r + 1];
  // descriptor of this group's first segment: fetched while the staging copies are in flight
  // a segment is shared by 4 lanes that sit 8 apart (lane = 8 * sub + segment): the 8 lanes of a quarter-warp - one
  // 128-bit shared-memory request - then work on 8 different segments at the SAME position of their contribution lists,
  // i.e. on the same landmark when the segments are the camera pairs of one camera set (the common case: landmarks are
  // ranked by camera list).  Their operands are then a handful of neighbouring blocks, most of them shared: broadcasts
  // instead of the bank conflicts of four lanes walking four landmarks deg * 144 bytes apart
  const int seg_in_cta = (tid >> 5) * 8 + (tid & 7);
  int sgm = g0 + seg_in_cta;
  int seg_tt = 0, seg_b = 0, seg_e = 0;
  if (sgm < g1) { seg_tt = R.seg_t[sgm]; seg_b = R.seg_cb[sgm]; seg_e = R.seg_ce[sgm]; }
  const int c0 = R.seg_cb[g0], nc = R.seg_ce[g1 - 1] - c0;  // every range has at least one segment; c0 % 8 == 0
  if (ns > R.cap_slots || nlm > R.cap_lms) __trap();  // the host never builds such a range (solver.cu: build_structure)
  const bool stC = nc <= R.cap_contrib;
  // staging: the range's piece of Hpl and its contribution indices are contiguous in HBM -> bulk (TMA) copies
  // tracked by one mbarrier; the Wu rows of the range's landmarks are gathered with 16-byte cp.async
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&sr_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned hbytes = (unsigned)ns * 144u;
    const unsigned cbytes = stC ? (((unsigned)nc * 2u + 15u) & ~15u) : 0u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(hbytes + 3u * cbytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(sH)),
                 "l"(Hpl + 18ll * s0), "r"(hbytes), "r"(bar)
                 : "memory");
    if (cbytes) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(sA)),
                   "l"(R.sc_a + c0), "r"(cbytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(sB)),
                   "l"(R.sc_b + c0), "r"(cbytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(sL)),
                   "l"(R.sc_l + c0), "r"(cbytes), "r"(bar)
                   : "memory");
    }
  }
  for (int i = tid; i < nlm * 5; i += kSrThreads) {
    const int l = i / 5, k = i - l * 5;
    const double* src = Wu + kDinvStride * (long long)R.lm_ids[l0 + l] + 2 * k;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sW + 2 * i)), "l"(src) : "memory");
  }
  for (int i = tid; i <= nlm; i += kSrThreads)  // nlm+1 entries: lm_slot carries one end entry per range
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sS + i)), "l"(R.lm_slot + (l0 + r) + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
  }
  __syncthreads();
  // transform in place: every block Hpl(i,l) becomes Hpl(i,l) W_l.  One thread per SLOT with 128-bit accesses: the 8
  // lanes of a quarter-warp touch 8 consecutive 144-byte blocks = 8 different 16-byte bank groups (a thread per
  // landmark walks its slots with a lane stride of deg * 144 bytes: 8- to 16-way bank conflicts for the common even
  // degrees - most of the 41 % conflict wavefronts of the round-1 profile)
  for (int l = tid; l < nlm; l += kSrThreads)
    for (int q = sS[l]; q < sS[l + 1]; ++q) sLm[q] = (unsigned short)l;
  __syncthreads();
  for (int q = tid; q < ns; q += kSrThreads) {
    const double* wl = sW + kDinvStride * sLm[q];
    const double w00 = wl[0], w01 = wl[1], w02 = wl[2], w11 = wl[3], w12 = wl[4], w22 = wl[5];
    double2* A2 = reinterpret_cast<double2*>(sH + 18 * q);
    double A[18];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double2 v = A2[k]; A[2 * k] = v.x; A[2 * k + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double a0 = A[i], a1 = A[i + 6], a2 = A[i + 12];
      A[i] = a0 * w00;
      A[i + 6] = fma(a0, w01, a1 * w11);
      A[i + 12] = fma(a0, w02, fma(a1, w12, a2 * w22));
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) A2[k] = make_double2(A[2 * k], A[2 * k + 1]);
  }
  __syncthreads();
  const unsigned short* ia = stC ? sA : R.sc_a + c0;
  const unsigned short* ib = stC ? sB : R.sc_b + c0;
  const unsigned short* il = stC ? sL : R.sc_l + c0;
  const int sub = (tid & 31) >> 3;
  const unsigned gmask = 0x01010101u << (tid & 7);
  for (; sgm < g1; sgm += kSrThreads / kSrLanes) {
    if (sgm != g0 + seg_in_cta) { seg_tt = R.seg_t[sgm]; seg_b = R.seg_cb[sgm]; seg_e = R.seg_ce[sgm]; }
    const bool diag = R.t_diag[seg_tt] != 0;
    const int cb = seg_b - c0, ce = seg_e - c0;
    double acc[36], cacc[6];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) cacc[k] = 0.0;
    // indices of the next product are fetched while this one is being multiplied
    int c = cb + sub;
    int a = 0, b = 0, ln = 0;
    if (c < ce) { a = ia[c]; b = ib[c]; ln = il[c]; }
    for (; c < ce; c += kSrLanes) {
      const int cn = c + kSrLanes;
      int an = 0, bn = 0, lnn = 0;
      if (cn < ce) { an = ia[cn]; bn = ib[cn]; lnn = il[cn]; }
      const double2* Ba = reinterpret_cast<const double2*>(sH + 18 * a);
      const double2* Bb = reinterpret_cast<const double2*>(sH + 18 * b);
      double A[18];
#pragma unroll
      for (int k = 0; k < 9; ++k) { const double2 v = Ba[k]; A[2 * k] = v.x; A[2 * k + 1] = v.y; }
      if (diag) {
        const double* u = sW + kDinvStride * ln + 6;
        const double u0 = u[0], u1 = u[1], u2 = u[2];
#pragma unroll
        for (int q = 0; q < 6; ++q) cacc[q] += A[q] * u0 + A[q + 6] * u1 + A[q + 12] * u2;
      }
      // += (Hpl(i1,l) W)(Hpl(i2,l) W)^T, one column of the second block (6 values = 3 double2) at a time
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double bj[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double2 v = Bb[3 * j + k]; bj[2 * k] = v.x; bj[2 * k + 1] = v.y; }
#pragma unroll
        for (int c2 = 0; c2 < 6; ++c2)
#pragma unroll
          for (int q = 0; q < 6; ++q) acc[q + 6 * c2] = fma(A[q + 6 * j], bj[c2], acc[q + 6 * c2]);
      }
      a = an; b = bn; ln = lnn;
    }
    // the 4 lanes of the group add up in a fixed tree; afterwards lane `sub` stores the entries k = sub mod 4
    double* out = partial + (long long)sgm * kSrPartial;
#pragma unroll
    for (int k = 0; k < 36; ++k) {
      double v = acc[k];
      v += __shfl_xor_sync(gmask, v, 8);
      v += __shfl_xor_sync(gmask, v, 16);
      if ((k & 3) == sub) out[k] = v;
    }
    if (diag) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double v = cacc[k];
        v += __shfl_xor_sync(gmask, v, 8);
        v += __shfl_xor_sync(gmask, v, 16);
        if ((k & 3) == sub) out[36 + k] = v;
      }
    }
  }
}

// Landmarks seen by more cameras than one range CTA can stage (k > 1400: 200 KB of Hpl blocks) take this path instead of
// being refused (round 1): one THREAD per camera pair (a <= b) of the landmark computes (Hpl_a W)(Hpl_b W)^T (+ the
// right-hand-side term on the diagonal pairs) straight from global memory - the k blocks of such a landmark (k * 144
// bytes) stay in L2 - and leaves it as a segment of its own for schur_finish_kernel.  Pairs are enumerated like the
// range plan does (a ascending, b = a .. k-1), so segment = seg0 + pair index.
struct SchurWide {
  int n;                       // wide landmarks
  const long long* pair0;      // n+1: first pair of every wide landmark
  const int* lm;               // landmark (Wu index)
  const int* slot0;            // its first Hpl slot
  const int* deg;              // its number of slots k
  const long long* seg0;       // its first segment
};
__global__ void __launch_bounds__(128)
schur_wide_kernel(SchurWide Wd, long long npairs, const double* __restrict__ Hpl, const double* __restrict__ Wu,
                  double* __restrict__ partial) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  int w = 0;
  while (w + 1 < Wd.n && Wd.pair0[w + 1] <= p) ++w;   // a handful of wide landmarks at most
  const long long q = p - Wd.pair0[w];
  const int k = Wd.deg[w];
  // invert q = a*k - a*(a-1)/2 + (b - a): largest a with a*k - a*(a-1)/2 <= q
  int a = (int)(((2.0 * k + 1.0) - sqrt((2.0 * k + 1.0) * (2.0 * k + 1.0) - 8.0 * (double)q)) * 0.5);
  a = max(0, min(a, k - 1));
  while (a > 0 && (long long)a * k - (long long)a * (a - 1) / 2 > q) --a;
  while (a + 1 < k && (long long)(a + 1) * k - (long long)(a + 1) * a / 2 <= q) ++a;
  const int b = a + (int)(q - ((long long)a * k - (long long)a * (a - 1) / 2));
  const double* wl = Wu + kDinvStride * (long long)Wd.lm[w];
  const double w00 = wl[0], w01 = wl[1], w02 = wl[2], w11 = wl[3], w12 = wl[4], w22 = wl[5];
  double A[18], B[18];
  {
    const double2* A2 = reinterpret_cast<const double2*>(Hpl + 18ll * (Wd.slot0[w] + a));
    const double2* B2 = reinterpret_cast<const double2*>(Hpl + 18ll * (Wd.slot0[w] + b));
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double2 v = A2[i]; A[2 * i] = v.x; A[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double2 v = B2[i]; B[2 * i] = v.x; B[2 * i + 1] = v.y; }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {   // the same transform schur_range_kernel applies in place: X <- X W
    double a0 = A[i], a1 = A[i + 6], a2 = A[i + 12];
    A[i] = a0 * w00; A[i + 6] = fma(a0, w01, a1 * w11); A[i + 12] = fma(a0, w02, fma(a1, w12, a2 * w22));
    a0 = B[i]; a1 = B[i + 6]; a2 = B[i + 12];
    B[i] = a0 * w00; B[i + 6] = fma(a0, w01, a1 * w11); B[i + 12] = fma(a0, w02, fma(a1, w12, a2 * w22));
  }
  double* out = partial + (Wd.seg0[w] + q) * kSrPartial;
#pragma unroll
  for (int c2 = 0; c2 < 6; ++c2)
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fma(A[r + 6 * j], B[c2 + 6 * j], v);
      out[r + 6 * c2] = v;
    }
  if (a == b) {
    const double u0 = wl[6], u1 = wl[7], u2 = wl[8];
#pragma unroll
    for (int r = 0; r < 6; ++r) out[36 + r] = A[r] * u0 + A[r + 6] * u1 + A[r + 12] * u2;
  }
}

// 64 threads per block of Hschur (42 active), 4 blocks per CTA
__global__ void __launch_bounds__(256)
schur_finish_kernel(int nT, const int* __restrict__ t_row, const int* __restrict__ t_col, const int* __restrict__ t_hpp,
                    const int* __restrict__ tseg_ptr, const int* __restrict__ tseg_idx,
                    const double* __restrict__ partial, const double* __restrict__ Hpp, const double* __restrict__ b_p,
                    const double* __restrict__ lambda, double hpp_scale, double lambda_scale,
                    double* __restrict__ Hschur, double* __restrict__ bschur) {
  const int t = blockIdx.x * 4 + (threadIdx.x >> 6);
  const int e = threadIdx.x & 63;
  if (t >= nT || e >= kSrPartial) return;
  const int i1 = t_row[t];
  const bool diag = i1 == t_col[t];
  if (e >= 36 && !diag) return;
  double s0 = 0.0;
  const int q1 = tseg_ptr[t + 1];
  int q = tseg_ptr[t];
  for (; q + 3 < q1; q += 4) {  // 4 independent loads in flight, added in segment order
    const double v0 = partial[(long long)tseg_idx[q] * kSrPartial + e];
    const double v1 = partial[(long long)tseg_idx[q + 1] * kSrPartial + e];
    const double v2 = partial[(long long)tseg_idx[q + 2] * kSrPartial + e];
    const double v3 = partial[(long long)tseg_idx[q + 3] * kSrPartial + e];
    s0 = (((s0 + v0) + v1) + v2) + v3;
  }
  for (; q < q1; ++q) s0 += partial[(long long)tseg_idx[q] * kSrPartial + e];
  if (e < 36) {
    const int hb = t_hpp[t];
    double base = 0.0;
    // landmark-sharded: Hpp holds this rank's partial camera blocks (summed by the all-reduce of Hschur), lambda is
    // added by rank 0 only (lambda_scale); single GPU: both scales are 1
    if (hb >= 0) base = hpp_scale * Hpp[36ll * hb + e];
    if (diag && e % 7 == 0) base += lambda_scale * *lambda;
    Hschur[36ll * t + e] = base - s0;
  } else {
    bschur[6ll * i1 + (e - 36)] = hpp_scale * b_p[6ll * i1 + (e - 36)] - s0;
  }
}

// landmark back-substitution: x_l = Dinv (b_l - sum_e Hpl(e)^T x_cam(e))     (block_solver.hpp:461-481)
__global__ void ba_backsub_kernel(int nl, const int* __restrict__ lm_eptr, const int* __restrict__ lm_order,
                                  const int* __restrict__ e_hpl,
                                  const int* __restrict__ e_pose, const double* __restrict__ Hpl,
                                  const double* __restrict__ Dinv, const double* __restrict__ b_l,
                                  const double* __restrict__ x_p, double* __restrict__ x_l) {
  const int rank = blockIdx.x * blockDim.x + threadIdx.x;  // landmark rank: edges and Hpl slots are in this order
  if (rank >= nl) return;
  const int l = lm_order[rank];
  double c0 = b_l[3ll * l], c1 = b_l[3ll * l + 1], c2 = b_l[3ll * l + 2];
  int prev = -1;
  for (int e = lm_eptr[rank]; e < lm_eptr[rank + 1]; ++e) {
    const int slot = e_hpl[e];
    if (slot < 0 || slot == prev) continue;  // duplicate observations share one block
    prev = slot;
    const double2* B2 = reinterpret_cast<const double2*>(Hpl + 18ll * slot);  // 16-byte vector loads
    const double2* xp2 = reinterpret_cast<const double2*>(x_p + 6ll * e_pose[e]);
    double B[18], xv[6];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double2 v = __ldg(B2 + k); B[2 * k] = v.x; B[2 * k + 1] = v.y; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double2 v = xp2[k]; xv[2 * k] = -v.x; xv[2 * k + 1] = -v.y; }
    double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) { t0 += B[r] * xv[r]; t1 += B[r + 6] * xv[r]; t2 += B[r + 12] * xv[r]; }
    c0 += t0; c1 += t1; c2 += t2;
  }
  const double* Di = Dinv + kDinvStride * (long long)l;
#pragma unroll
  for (int r = 0; r < 3; ++r) x_l[3ll * l + r] = Di[r] * c0 + Di[r + 3] * c1 + Di[r + 6] * c2;
}

// The tail of an LM trial for the landmarks in ONE pass over their observations (one thread per landmark, rank order):
//   x_l = Dinv (b_l - sum_e Hpl(e)^T x_cam(e))                        back-substitution, block_solver.hpp:461-481
//   scale_l += x_l . (lambda x_l + b_l)                               computeScale, optimization_algorithm_levenberg.cpp:165-172
//   X_l <- X_l + x_l                                                  VertexSBAPointXYZ::oplusImpl, types_sba.h:151-155
//   chi2 += sum_e rho(e_e^T Omega_e e_e) at the NEW state             computeActiveErrors + activeRobustChi2
// (the cameras have been updated by oplus_cam_kernel before).  Replaces ba_backsub_kernel + oplus_xyz_kernel +
// ba_chi2_kernel + lm_scale_kernel(landmark part): the observations' indices and the landmark are read once instead of
// three times.  Block sums are fixed trees: bit-identical from run to run.
template <int MODEL>
__global__ void __launch_bounds__(128)
ba_backsub_update_kernel(int nl, const int* __restrict__ lm_eptr, const int* __restrict__ lm_order,
                         const int* __restrict__ lm_vertex, const int* __restrict__ e_cam, const int* __restrict__ e_hpl,
                         const int* __restrict__ e_pose, const double* __restrict__ Hpl, const double* __restrict__ Dinv,
                         const double* __restrict__ b_l, const double* __restrict__ x_p, double* __restrict__ x_l,
                         const double* __restrict__ lambda, double* __restrict__ pt_est, const double* __restrict__ cam_der,
                         const double* __restrict__ meas, const double* __restrict__ info, int E, Robust rk,
                         double* __restrict__ partial_chi2, double* __restrict__ partial_scale) {
  const int rank = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0, sc = 0.0;
  if (rank < nl) {
    const int l = lm_order[rank];
    const int e0 = lm_eptr[rank], e1 = lm_eptr[rank + 1];
    const double bl0 = b_l[3ll * l], bl1 = b_l[3ll * l + 1], bl2 = b_l[3ll * l + 2];
    double c0 = bl0, c1 = bl1, c2 = bl2;
    int prev = -1;
    for (int e = e0; e < e1; ++e) {
      const int slot = e_hpl[e];
      if (slot < 0 || slot == prev) continue;  // duplicate observations share one block
      prev = slot;
      const double2* B2 = reinterpret_cast<const double2*>(Hpl + 18ll * slot);
      const double2* xp2 = reinterpret_cast<const double2*>(x_p + 6ll * e_pose[e]);
      double B[18], xv[6];
#pragma unroll
      for (int k = 0; k < 9; ++k) { const double2 v = __ldg(B2 + k); B[2 * k] = v.x; B[2 * k + 1] = v.y; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double2 v = xp2[k]; xv[2 * k] = -v.x; xv[2 * k + 1] = -v.y; }
      double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) { t0 += B[r] * xv[r]; t1 += B[r + 6] * xv[r]; t2 += B[r + 12] * xv[r]; }
      c0 += t0; c1 += t1; c2 += t2;
    }
    const double* Di = Dinv + kDinvStride * (long long)l;
    double x[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) x[r] = Di[r] * c0 + Di[r + 3] * c1 + Di[r + 6] * c2;
#pragma unroll
    for (int r = 0; r < 3; ++r) x_l[3ll * l + r] = x[r];
    const double lam = *lambda;
    sc = x[0] * (lam * x[0] + bl0) + x[1] * (lam * x[1] + bl1) + x[2] * (lam * x[2] + bl2);
    double4* Xp = reinterpret_cast<double4*>(pt_est + 4ll * lm_vertex[l]);
    double4 X4 = *Xp;
    X4.x += x[0]; X4.y += x[1]; X4.z += x[2];
    *Xp = X4;
    const double X[3] = {X4.x, X4.y, X4.z};
    for (int e = e0; e < e1; ++e) {
      double der[16];
      load_der(cam_der, e_cam[e], der);
      const double z[2] = {meas[e], meas[(long long)E + e]};
      double err[2];
      ba_error<MODEL>(der, X, z, err);
      const double w0 = info[e], w1 = info[(long long)E + e], w2 = info[2ll * E + e];
      double ce = err[0] * (w0 * err[0] + w1 * err[1]) + err[1] * (w1 * err[0] + w2 * err[1]);
      const Robust rke = robust_at(rk, e);
    if (rke.kind) { double r1; robustify(rke, ce, ce, r1); }
      chi += ce;
    }
  }
  chi = block_sum(chi);
  if (threadIdx.x == 0) partial_chi2[blockIdx.x] = chi;
  sc = block_sum(sc);
  if (threadIdx.x == 0) partial_scale[blockIdx.x] = sc;
}

// A(i,i) += lambda I (+ extra: the unit diagonal of padding unknowns) on the diagonal blocks of a block matrix: what
// Solver::setLambda does in place (block_solver.hpp:563-604), applied to the copy the PCG solver reads
template <int D>
__global__ void add_block_diagonal_kernel(int np, const int* __restrict__ diag_block, const double* __restrict__ lambda,
                                          const double* __restrict__ extra, double* __restrict__ A) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= np * D) return;
  const int i = idx / D, r = idx - i * D;
  A[(long long)D * D * diag_block[i] + r + D * r] += *lambda + (extra ? extra[idx] : 0.0);
}

// ------------------------------------------------------------------ oplus updates
// hidx[v] = hessian index (-1 fixed); x is indexed by colInHessian = hidx*dim (poses) / sizePoses + ... (landmarks)
__global__ void oplus_se2_kernel(int n, const int* __restrict__ hidx, const double* __restrict__ x, double* __restrict__ est) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int h = hidx[v];
  if (h < 0) return;
  double e[3] = {est[4ll * v], est[4ll * v + 1], est[4ll * v + 2]};
  se2_oplus(e, x + 3ll * h);
  est[4ll * v] = e[0]; est[4ll * v + 1] = e[1]; est[4ll * v + 2] = e[2];
}
__global__ void oplus_se3_kernel(int n, const int* __restrict__ hidx, const double* __restrict__ x, double* __restrict__ est,
                                 int orthogonalize) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int h = hidx[v];
  if (h < 0) return;
  double e[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) e[k] = est[12ll * v + k];
  se3_oplus(e, x + 6ll * h, orthogonalize != 0);
#pragma unroll
  for (int k = 0; k < 12; ++k) est[12ll * v + k] = e[k];
}
template <int MODEL>
__global__ void oplus_cam_kernel(int n, const int* __restrict__ hidx, const double* __restrict__ x, double* __restrict__ est,
                                 double* __restrict__ der) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int h = hidx[v];
  if (h < 0) return;
  double e[12], d[16];
#pragma unroll
  for (int k = 0; k < 12; ++k) e[k] = est[12ll * v + k];
  ba_oplus<MODEL>(e, x + 6ll * h);
  ba_derive<MODEL>(e, d);
#pragma unroll
  for (int k = 3; k < 7; ++k) est[12ll * v + k] = e[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) est[12ll * v + k] = e[k];
#pragma unroll
  for (int k = 0; k < 16; ++k) der[16ll * v + k] = d[k];
}
// lidx[v] = landmark index (-1 fixed); x_l = x + sizePoses
__global__ void oplus_xyz_kernel(int n, const int* __restrict__ lidx, const double* __restrict__ x_l, double* __restrict__ est) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int l = lidx[v];
  if (l < 0) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) est[4ll * v + k] += x_l[3ll * l + k];
}

// ------------------------------------------------------------------ host <-> device estimate staging
// device rows are padded (3 -> 4 doubles) for aligned vector loads; the ABI layout is dense
__global__ void unpack_rows_kernel(int n, int ne, int st, const double* __restrict__ dense, double* __restrict__ padded) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * ne) return;
  const int v = (int)(i / ne), k = (int)(i - (long long)v * ne);
  padded[(long long)v * st + k] = dense[i];
}
__global__ void pack_rows_kernel(int n, int ne, int st, const double* __restrict__ padded, double* __restrict__ dense) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * ne) return;
  const int v = (int)(i / ne), k = (int)(i - (long long)v * ne);
  dense[i] = padded[(long long)v * st + k];
}

// ------------------------------------------------------------------ LM scalars
// partial sums of x_j (lambda x_j + b_j)        (optimization_algorithm_levenberg.cpp:165-172)
__global__ void lm_scale_kernel(int n, const double* __restrict__ x, const double* __restrict__ b,
                                const double* __restrict__ lambda, double* __restrict__ partials) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (j < n) v = x[j] * (*lambda * x[j] + b[j]);
  v = block_sum(v);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}
// max |H_jj| over a set of diagonal blocks     (optimization_algorithm_levenberg.cpp:149-163)
template <int D>
__global__ void max_diag_kernel(int nblocks, const int* __restrict__ block_index, const double* __restrict__ H,
                                double* __restrict__ partials) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (idx < nblocks * D) {
    const int k = idx / D, r = idx - k * D;
    const long long blk = block_index ? block_index[k] : k;
    v = fabs(H[blk * D * D + r + D * r]);
  }
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) m = fmax(m, sh[i]);
    partials[blockIdx.x] = m;
  }
}
__global__ void reduce_max_kernel(const double* __restrict__ partials, int n, double scale, double* __restrict__ out,
                                  int accumulate) {
  __shared__ double sh[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, partials[i]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  m = warp_max(m);
  if (lane == 0) sh[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) r = fmax(r, sh[i]);
    r *= scale;
    out[0] = accumulate ? fmax(out[0], r) : r;
  }
}
// partial maxima of |v[i]| (sharded lambda init: the reduced Hpp diagonal and the per-rank landmark maxima)
__global__ void absmax_kernel(int n, const double* __restrict__ v, double* __restrict__ partials) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  double m = idx < n ? fabs(v[idx]) : 0.0;
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  m = warp_max(m);
  if (lane == 0) sh[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) r = fmax(r, sh[i]);
    partials[blockIdx.x] = r;
  }
}
// sharded LM scale: the pose part is a partial sum too -> one slot for the all-reduce
__global__ void fold_scale_kernel(double* __restrict__ scalars) {
  scalars[1] += scalars[4];
  scalars[4] = 0.0;
}
// gather the diagonal entries of the indexed vertices into a dense vector (host mirror for v->hessian(j,j))
template <int D>
__global__ void extract_diag_kernel(int nblocks, const int* __restrict__ block_index, const double* __restrict__ H,
                                    double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblocks * D) return;
  const int k = idx / D, r = idx - k * D;
  const long long blk = block_index ? block_index[k] : k;
  out[idx] = H[blk * D * D + r + D * r];
}

}  // namespace k
}  // namespace g2o_b200
