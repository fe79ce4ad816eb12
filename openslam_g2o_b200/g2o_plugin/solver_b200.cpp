// solver_b200.cpp - g2o plugin "libg2o_solver_b200.so": registers {gn,lm}_{fix3_2,fix6_3,var}_b200 with g2o's
// OptimizationAlgorithmFactory (g2o/core/optimization_algorithm_factory.h:153-162), same naming scheme as
// g2o/solvers/cholmod/solver_cholmod.cpp:41-132 (+ suffix _b200s: Level 2, _b200ls: Level 1).  The `g2o` binary picks it up through its *_solver_*.so glob
// (apps/g2o_cli/g2o_common.cpp:82,133-167); programmatic users `new OptimizationAlgorithmB200(...)` directly.
//
// Level 3 (this file): OptimizationAlgorithmB200 keeps the whole LM / GN iteration on the GPU behind
// b200_algorithm_solve(); the host only sees a handful of scalars per trial and gets the estimates written back into
// the vertices at the end of every solve() (what SparseOptimizer::optimize reads when verbose / statistics are on,
// core/sparse_optimizer.cpp:392-411).
// Level 2 (this file): SolverB200 is a g2o::Solver (core/solver.h:44-149) for the stock OptimizationAlgorithm
// {GaussNewton,Levenberg}: buildSystem / setLambda / solve / computeMarginals on the GPU, errors and update stay on the
// host where those algorithms call them; x, b and the Hessian diagonal are mirrored on the host.
// Level 1 is linear_solver_b200.h (LinearSolverB200 under the stock BlockSolver).
//
// Compiles only inside a g2o source tree (needs Eigen + g2o headers; neither exists in the build container).
#include <iostream>
#include <vector>

#include "g2o/core/robust_kernel_impl.h"
#include "g2o/core/batch_stats.h"
#include "g2o/core/block_solver.h"
#include "g2o/core/optimization_algorithm.h"
#include "g2o/core/optimization_algorithm_factory.h"
#include "g2o/core/optimization_algorithm_gauss_newton.h"
#include "g2o/core/optimization_algorithm_levenberg.h"
#include "g2o/core/solver.h"
#include "g2o/core/sparse_optimizer.h"
#include "g2o/stuff/macros.h"
#include "g2o/types/sba/types_sba.h"
#include "g2o/types/sba/types_six_dof_expmap.h"
#include "g2o/types/slam2d/edge_se2.h"
#include "g2o/types/slam2d/edge_se2_pointxy.h"
#include "g2o/types/slam2d/vertex_point_xy.h"
#include "g2o/types/slam2d/vertex_se2.h"
#include "g2o/types/slam3d/edge_se3.h"
#include "g2o/types/slam3d/edge_se3_pointxyz.h"
#include "g2o/types/slam3d/parameter_se3_offset.h"
#include "g2o/types/slam3d/vertex_pointxyz.h"
#include "g2o/types/slam3d/vertex_se3.h"
#include "g2o_b200.h"
#include "linear_solver_b200.h"

namespace g2o {

// estimate of one vertex in the C-ABI layout (include/g2o_b200.h); returns its B200_VERTEX_* kind, -1 if unsupported
static int estimateLength(int kind) { return (kind == B200_VERTEX_SE2 || kind == B200_VERTEX_XYZ) ? 3 : kind == B200_VERTEX_XY ? 2 : 12; }
static int packEstimate(OptimizableGraph::Vertex* v, double* e) {
  int kind = -1;
  if (VertexSE2* p = dynamic_cast<VertexSE2*>(v)) {
    kind = B200_VERTEX_SE2;
    e[0] = p->estimate().translation().x(); e[1] = p->estimate().translation().y(); e[2] = p->estimate().rotation().angle();
  } else if (VertexSE3* p = dynamic_cast<VertexSE3*>(v)) {
    kind = B200_VERTEX_SE3;  // the state is the Isometry3d itself: R (col-major) | t, never re-quaternionised
    Eigen::Map<Eigen::Matrix3d>(e) = p->estimate().linear();
    Eigen::Map<Eigen::Vector3d>(e + 9) = p->estimate().translation();
  } else if (VertexCam* p = dynamic_cast<VertexCam*>(v)) {
    kind = B200_VERTEX_CAM;
    const SBACam& c = p->estimate();
    Eigen::Map<Eigen::Vector3d>(e) = c.translation();
    Eigen::Map<Eigen::Vector4d>(e + 3) = c.rotation().coeffs();  // x y z w
    e[7] = c.Kcam(0, 0); e[8] = c.Kcam(1, 1); e[9] = c.Kcam(0, 2); e[10] = c.Kcam(1, 2); e[11] = c.baseline;
  } else if (VertexSE3Expmap* p = dynamic_cast<VertexSE3Expmap*>(v)) {
    kind = B200_VERTEX_SE3_EXPMAP;  // world -> camera SE3Quat; the intrinsics [7..12) come from its edges' CameraParameters
    Eigen::Map<Eigen::Vector3d>(e) = p->estimate().translation();
    Eigen::Map<Eigen::Vector4d>(e + 3) = p->estimate().rotation().coeffs();  // x y z w
    e[7] = e[8] = e[9] = e[10] = e[11] = 0.;
  } else if (VertexSBAPointXYZ* p = dynamic_cast<VertexSBAPointXYZ*>(v)) {
    kind = B200_VERTEX_XYZ;
    Eigen::Map<Eigen::Vector3d>(e) = p->estimate();
  } else if (VertexPointXYZ* p = dynamic_cast<VertexPointXYZ*>(v)) {  // landmark SLAM in 3D: same state, same update
    kind = B200_VERTEX_XYZ;
    Eigen::Map<Eigen::Vector3d>(e) = p->estimate();
  } else if (VertexPointXY* p = dynamic_cast<VertexPointXY*>(v)) {
    kind = B200_VERTEX_XY;
    e[0] = p->estimate()[0]; e[1] = p->estimate()[1];
  }
  return kind;
}

// the pointer graph <-> SoA arrays of the C-ABI, shared by the Level-3 algorithm and the Level-2 solver
struct B200GraphBinding {
  // one-time packing of the pointer graph into SoA arrays (what BlockSolver::buildStructure walks,
  // core/block_solver.hpp:142-295).  Unknown vertex / edge types -> false (no CPU fallback).
  bool ingest(b200_ctx* _ctx, SparseOptimizer* _optimizer) {
    const SparseOptimizer::VertexContainer& verts = _optimizer->activeVertices();
    const SparseOptimizer::EdgeContainer& edges = _optimizer->activeEdges();
    std::vector<double>* est = _est;
    for (int k = 0; k < B200_NUM_VERTEX_KINDS; ++k) _est[k].clear();
    std::vector<int32_t> hidx[B200_NUM_VERTEX_KINDS];
    std::vector<uint8_t> marg[B200_NUM_VERTEX_KINDS];
    _slot.clear();
    for (int k = 0; k < B200_NUM_VERTEX_KINDS; ++k) _verts[k].clear();
    for (size_t i = 0; i < verts.size(); ++i) {
      OptimizableGraph::Vertex* v = verts[i];
      double e[12];
      const int kind = packEstimate(v, e);
      if (kind < 0) {
        std::cerr << "OptimizationAlgorithmB200: unsupported vertex type (id " << v->id() << ")" << std::endl;
        return false;
      }
      const int ne = estimateLength(kind);
      _slot[v] = static_cast<int>(hidx[kind].size());
      _verts[kind].push_back(v);
      est[kind].insert(est[kind].end(), e, e + ne);
      hidx[kind].push_back(v->hessianIndex());
      marg[kind].push_back(v->marginalized() ? 1 : 0);
    }
    _hasLandmarks = !hidx[B200_VERTEX_XYZ].empty();
    std::vector<int32_t> vi, vj;
    std::vector<double> meas, info;
    // landmark SLAM: the pose-landmark edges form a second edge set beside the pose-pose edges (g2o_b200.h)
    std::vector<int32_t> lvi, lvj;
    std::vector<double> lmeas, linfo;
    int lkind = -1;
    const ParameterSE3Offset* sensorOffset = 0;
    int ekind = -1;
    int robustKind = B200_ROBUST_NONE;
    double robustDelta = 1.0;
    // kernels that differ between edges (Edge::setRobustKernel on individual edges) go over per edge
    bool uniformKernel = true;
    std::vector<uint8_t> rkKinds, lrkKinds;
    std::vector<double> rkDeltas, lrkDeltas;
    for (size_t k = 0; k < edges.size(); ++k) {
      OptimizableGraph::Edge* e = edges[k];
      int kind = -1;
      if (EdgeSE2* p = dynamic_cast<EdgeSE2*>(e)) {
        kind = B200_EDGE_SE2;
        Eigen::Vector3d z = p->measurement().toVector();
        meas.insert(meas.end(), z.data(), z.data() + 3);
        info.insert(info.end(), p->information().data(), p->information().data() + 9);
      } else if (EdgeSE3* p = dynamic_cast<EdgeSE3*>(e)) {
        kind = B200_EDGE_SE3;
        double z[12];
        Eigen::Map<Eigen::Matrix3d>(z) = p->measurement().linear();
        Eigen::Map<Eigen::Vector3d>(z + 9) = p->measurement().translation();
        meas.insert(meas.end(), z, z + 12);
        info.insert(info.end(), p->information().data(), p->information().data() + 36);
      } else if (EdgeProjectP2MC* p = dynamic_cast<EdgeProjectP2MC*>(e)) {
        kind = B200_EDGE_P2MC;
        meas.insert(meas.end(), p->measurement().data(), p->measurement().data() + 2);
        info.insert(info.end(), p->information().data(), p->information().data() + 4);
      } else if (EdgeProjectXYZ2UV* p = dynamic_cast<EdgeProjectXYZ2UV*>(e)) {
        kind = B200_EDGE_XYZ2UV;
        meas.insert(meas.end(), p->measurement().data(), p->measurement().data() + 2);
        info.insert(info.end(), p->information().data(), p->information().data() + 4);
        // the pose row carries the CameraParameters of its edges (parameter(0), types_six_dof_expmap.h:146-147)
        const CameraParameters* cam = static_cast<const CameraParameters*>(p->parameter(0));
        double* row = &est[B200_VERTEX_SE3_EXPMAP][12 * _slot[static_cast<OptimizableGraph::Vertex*>(e->vertex(1))]];
        const double want[5] = {cam->focal_length, cam->focal_length, cam->principle_point[0], cam->principle_point[1], cam->baseline};
        for (int i = 0; i < 5; ++i) {
          if (row[7] != 0. && row[7 + i] != want[i]) {
            std::cerr << "OptimizationAlgorithmB200: edges of one pose use different CameraParameters" << std::endl;
            return false;
          }
        }
        for (int i = 0; i < 5; ++i) row[7 + i] = want[i];
      }
      else if (EdgeSE2PointXY* p = dynamic_cast<EdgeSE2PointXY*>(e)) {
        kind = B200_EDGE_SE2_XY;
        lmeas.insert(lmeas.end(), p->measurement().data(), p->measurement().data() + 2);
        linfo.insert(linfo.end(), p->information().data(), p->information().data() + 4);
      } else if (EdgeSE3PointXYZ* p = dynamic_cast<EdgeSE3PointXYZ*>(e)) {
        kind = B200_EDGE_SE3_XYZ;
        lmeas.insert(lmeas.end(), p->measurement().data(), p->measurement().data() + 3);
        linfo.insert(linfo.end(), p->information().data(), p->information().data() + 9);
        const ParameterSE3Offset* off = static_cast<const ParameterSE3Offset*>(p->parameter(0));
        if (sensorOffset && !sensorOffset->offset().isApprox(off->offset(), 0.)) {
          std::cerr << "OptimizationAlgorithmB200: EdgeSE3PointXYZ edges with different ParameterSE3Offset values" << std::endl;
          return false;
        }
        sensorOffset = off;
      }
      const bool landmarkEdge = kind == B200_EDGE_SE2_XY || kind == B200_EDGE_SE3_XYZ;
      if (kind < 0 || (!landmarkEdge && ekind >= 0 && kind != ekind) || (landmarkEdge && lkind >= 0 && kind != lkind)) {
        std::cerr << "OptimizationAlgorithmB200: unsupported / mixed edge types" << std::endl;
        return false;
      }
      // one robust kernel (type and width) for every edge - what `g2o -robustKernel` sets up (g2o.cpp:322-336)
      int rk = B200_ROBUST_NONE;
      double rkDelta = 1.0;
      if (RobustKernel* r = e->robustKernel()) {
        rkDelta = r->delta();
        if (dynamic_cast<RobustKernelHuber*>(r)) rk = B200_ROBUST_HUBER;
        else if (dynamic_cast<RobustKernelPseudoHuber*>(r)) rk = B200_ROBUST_PSEUDO_HUBER;
        else if (dynamic_cast<RobustKernelCauchy*>(r)) rk = B200_ROBUST_CAUCHY;
        else if (dynamic_cast<RobustKernelSaturated*>(r)) rk = B200_ROBUST_SATURATED;
        else if (dynamic_cast<RobustKernelDCS*>(r)) rk = B200_ROBUST_DCS;
        else rk = -1;
      }
      if (rk < 0) {
        std::cerr << "OptimizationAlgorithmB200: unsupported robust kernel type" << std::endl;
        return false;
      }
      if (k > 0 && (rk != robustKind || rkDelta != robustDelta)) uniformKernel = false;
      robustKind = rk; robustDelta = rkDelta;
      (landmarkEdge ? lrkKinds : rkKinds).push_back(static_cast<uint8_t>(rk));
      (landmarkEdge ? lrkDeltas : rkDeltas).push_back(rkDelta);
      if (landmarkEdge) {
        lkind = kind;
        lvi.push_back(_slot[static_cast<OptimizableGraph::Vertex*>(e->vertex(0))]);
        lvj.push_back(_slot[static_cast<OptimizableGraph::Vertex*>(e->vertex(1))]);
        continue;
      }
      ekind = kind;
      vi.push_back(_slot[static_cast<OptimizableGraph::Vertex*>(e->vertex(0))]);
      vj.push_back(_slot[static_cast<OptimizableGraph::Vertex*>(e->vertex(1))]);
    }
    if (vi.empty() && lvi.empty()) return false;
    if (lkind >= 0 && ekind < 0) ekind = lkind == B200_EDGE_SE2_XY ? B200_EDGE_SE2 : B200_EDGE_SE3;  // no odometry at all
    if (sensorOffset) {
      double iso[12];
      Eigen::Map<Eigen::Matrix3d>(iso) = sensorOffset->offset().linear();
      Eigen::Map<Eigen::Vector3d>(iso + 9) = sensorOffset->offset().translation();
      if (b200_set_sensor_offset(_ctx, iso) != B200_OK) return false;
    }
    for (int k = 0; k < B200_NUM_VERTEX_KINDS; ++k)
      if (!hidx[k].empty() && b200_set_vertices(_ctx, k, static_cast<int>(hidx[k].size()), &est[k][0], &hidx[k][0], &marg[k][0]) != B200_OK)
        return false;
    if (b200_set_edges(_ctx, ekind, static_cast<int>(vi.size()), vi.empty() ? 0 : &vi[0], vi.empty() ? 0 : &vj[0], vi.empty() ? 0 : &meas[0],
                       vi.empty() ? 0 : &info[0]) != B200_OK) return false;
    // the pose-landmark set (n = 0 forgets the set of a previous graph)
    if (b200_set_edges(_ctx, lkind >= 0 ? lkind : B200_EDGE_SE2_XY, static_cast<int>(lvi.size()), lvi.empty() ? 0 : &lvi[0], lvi.empty() ? 0 : &lvj[0],
                       lvi.empty() ? 0 : &lmeas[0], lvi.empty() ? 0 : &linfo[0]) != B200_OK) return false;
    if (uniformKernel) {
      if (b200_set_robust_kernel(_ctx, robustKind, robustDelta) != B200_OK) return false;
    } else {
      if (b200_set_robust_kernel(_ctx, B200_ROBUST_NONE, 1.0) != B200_OK) return false;
      if (!rkKinds.empty() && b200_set_edge_robust_kernels(_ctx, ekind, static_cast<int>(rkKinds.size()), &rkKinds[0], &rkDeltas[0]) != B200_OK) return false;
      if (!lrkKinds.empty() && b200_set_edge_robust_kernels(_ctx, lkind, static_cast<int>(lrkKinds.size()), &lrkKinds[0], &lrkDeltas[0]) != B200_OK) return false;
    }
    int rc = b200_build_structure(_ctx);
    if (rc != B200_OK) std::cerr << "OptimizationAlgorithmB200: " << b200_last_error(_ctx) << std::endl;
    return rc == B200_OK;
  }

  // device estimates -> vertices
  void writeBack(b200_ctx* _ctx) {
    std::vector<double> buf;
    for (int kind = 0; kind < B200_NUM_VERTEX_KINDS; ++kind) {
      const std::vector<OptimizableGraph::Vertex*>& vs = _verts[kind];
      if (vs.empty()) continue;
      const int ne = estimateLength(kind);
      buf.resize(vs.size() * ne);
      if (b200_get_estimates(_ctx, kind, &buf[0]) != B200_OK) continue;
      for (size_t i = 0; i < vs.size(); ++i) {
        const double* e = &buf[i * ne];
        if (vs[i]->fixed()) continue;
        if (kind == B200_VERTEX_SE2) static_cast<VertexSE2*>(vs[i])->setEstimate(SE2(e[0], e[1], e[2]));
        else if (kind == B200_VERTEX_XYZ) {
          if (VertexSBAPointXYZ* q = dynamic_cast<VertexSBAPointXYZ*>(vs[i])) q->setEstimate(Eigen::Vector3d(e[0], e[1], e[2]));
          else static_cast<VertexPointXYZ*>(vs[i])->setEstimate(Eigen::Vector3d(e[0], e[1], e[2]));
        } else if (kind == B200_VERTEX_XY) static_cast<VertexPointXY*>(vs[i])->setEstimate(Eigen::Vector2d(e[0], e[1]));
        else if (kind == B200_VERTEX_SE3) {
          Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
          T.linear() = Eigen::Map<const Eigen::Matrix3d>(e);
          T.translation() = Eigen::Map<const Eigen::Vector3d>(e + 9);
          static_cast<VertexSE3*>(vs[i])->setEstimate(T);
        } else if (kind == B200_VERTEX_SE3_EXPMAP) {
          SE3Quat T;  // the device keeps q normalised with w >= 0 (SE3Quat::normalizeRotation)
          T.setRotation(Eigen::Quaterniond(e[6], e[3], e[4], e[5]));
          T.setTranslation(Eigen::Vector3d(e[0], e[1], e[2]));
          static_cast<VertexSE3Expmap*>(vs[i])->setEstimate(T);
        } else {
          SBACam cam(Eigen::Quaterniond(e[6], e[3], e[4], e[5]), Eigen::Vector3d(e[0], e[1], e[2]));
          cam.setKcam(e[7], e[8], e[9], e[10], e[11]);
          static_cast<VertexCam*>(vs[i])->setEstimate(cam);
        }
      }
    }
  }

  // host vertices -> device estimates (Level 2: SparseOptimizer::update / push / pop run on the host between solves)
  bool pushEstimates(b200_ctx* ctx) {
    for (int kind = 0; kind < B200_NUM_VERTEX_KINDS; ++kind) {
      const std::vector<OptimizableGraph::Vertex*>& vs = _verts[kind];
      if (vs.empty()) continue;
      const int ne = estimateLength(kind);
      const int keep = kind == B200_VERTEX_SE3_EXPMAP ? 7 : ne;  // expmap rows keep the intrinsics found at ingest
      for (size_t i = 0; i < vs.size(); ++i) {
        double e[12];
        if (packEstimate(vs[i], e) != kind) return false;
        for (int k = 0; k < keep; ++k) _est[kind][i * ne + k] = e[k];
      }
      if (b200_set_estimates(ctx, kind, &_est[kind][0]) != B200_OK) return false;
    }
    return true;
  }

  bool _hasLandmarks;
  std::map<OptimizableGraph::Vertex*, int> _slot;
  std::vector<OptimizableGraph::Vertex*> _verts[B200_NUM_VERTEX_KINDS];
  std::vector<double> _est[B200_NUM_VERTEX_KINDS];  // rows as handed to b200_set_vertices
};

// Solver::computeMarginals (core/block_solver.hpp:490-499) -> LinearSolver::solvePattern -> MarginalCovarianceCholesky
static bool b200Marginals(b200_ctx* _ctx, SparseBlockMatrix<MatrixXd>& spinv, const std::vector<std::pair<int, int> >& blockIndices) {
  if (blockIndices.empty()) return true;
  int dims[8];
  if (b200_get_dims(_ctx, dims) != B200_OK) return false;
  const int d = dims[6];
  if (b200_build_system(_ctx) != B200_OK) return false;  // Hpp at the current estimates, no lambda
  std::vector<int32_t> rows(blockIndices.size()), cols(blockIndices.size());
  for (size_t q = 0; q < blockIndices.size(); ++q) { rows[q] = blockIndices[q].first; cols[q] = blockIndices[q].second; }
  std::vector<double> out(blockIndices.size() * d * d);
  if (b200_compute_marginals(_ctx, static_cast<int>(rows.size()), &rows[0], &cols[0], &out[0]) != B200_OK) {
    std::cerr << "OptimizationAlgorithmB200: " << b200_last_error(_ctx) << std::endl;
    return false;
  }
  // same layout as MarginalCovarianceCholesky::computeCovariance: uniform d x d blocks indexed by Hessian index
  std::vector<int> blockEnds(dims[0]);
  for (int i = 0; i < dims[0]; ++i) blockEnds[i] = (i + 1) * d;
  spinv = SparseBlockMatrix<MatrixXd>(&blockEnds[0], &blockEnds[0], dims[0], dims[0], true);
  for (size_t q = 0; q < blockIndices.size(); ++q) {
    MatrixXd* blk = spinv.block(rows[q], cols[q], true);
    *blk = Eigen::Map<const Eigen::MatrixXd>(&out[q * d * d], d, d);
  }
  return true;
}

class OptimizationAlgorithmB200 : public OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmB200(int algorithm, bool pcg = false) : OptimizationAlgorithm(), _ctx(0), _algorithm(algorithm), _pcg(pcg) {
    _device = _properties.makeProperty<Property<int> >("device", 0);
    _userLambdaInit = _properties.makeProperty<Property<double> >("initialLambda", 0.);
    _maxTrialsAfterFailure = _properties.makeProperty<Property<int> >("maxTrialsAfterFailure", 10);
    // 0: block AMD, the reference's ordering; k > 0: nested dissection with 2^k parts on top of it (b200_set_ordering)
    _ndLevels = _properties.makeProperty<Property<int> >("ndLevels", 0);
    // LinearSolverPCG's setters (solvers/pcg/linear_solver_pcg.h:75-82), used by the *_pcg*_b200 solvers
    _pcgTolerance = _properties.makeProperty<Property<double> >("pcgTolerance", 1e-6);
    _pcgAbsoluteTolerance = _properties.makeProperty<Property<bool> >("pcgAbsoluteTolerance", true);
    _pcgMaxIterations = _properties.makeProperty<Property<int> >("pcgMaxIterations", -1);
  }
  virtual ~OptimizationAlgorithmB200() { b200_destroy(_ctx); }

  // OptimizationAlgorithmWithHessian::init (core/optimization_algorithm_with_hessian.cpp:50-73): (re)ingest the graph
  virtual bool init(bool /*online*/ = false) {
    if (!_ctx && b200_create(_device->value(), &_ctx) != B200_OK) {
      std::cerr << "OptimizationAlgorithmB200: " << b200_last_error(0) << std::endl;  // no CPU fallback
      return false;
    }
    b200_set_lm_params(_ctx, _userLambdaInit->value(), _maxTrialsAfterFailure->value());
    b200_set_ordering(_ctx, _ndLevels->value());
    b200_set_linear_solver(_ctx, _pcg ? B200_LINEAR_SOLVER_PCG : B200_LINEAR_SOLVER_CHOLESKY, _pcgTolerance->value(),
                           _pcgAbsoluteTolerance->value() ? 1 : 0, _pcgMaxIterations->value());
    return _graph.ingest(_ctx, _optimizer);
  }

  static int terminateHook(void* self) { return static_cast<OptimizationAlgorithmB200*>(self)->_optimizer->terminate() ? 1 : 0; }

  virtual SolverResult solve(int iteration, bool /*online*/ = false) {
    b200_iter_stats st;
    G2OBatchStatistics* gsOn = G2OBatchStatistics::globalStats();
    // `g2o -stats`: per-phase CUDA-event timing (individual launches instead of graph replays), like the reference's
    // stats mode costs an extra error pass per iteration (core/sparse_optimizer.cpp:392-397)
    if (iteration == 0) {
      b200_set_profiling(_ctx, gsOn ? 1 : 0);
      b200_set_terminate(_ctx, &OptimizationAlgorithmB200::terminateHook, this);  // forceStopFlag, sparse_optimizer.h:189
    }
    int rc = b200_algorithm_solve(_ctx, _algorithm, iteration, &st);
    if (rc < 0) {  // hard error (bad structure, CUDA failure, out of memory): st is not meaningful
      std::cerr << "OptimizationAlgorithmB200: " << b200_last_error(_ctx) << std::endl;
      return Fail;
    }
    G2OBatchStatistics* gs = G2OBatchStatistics::globalStats();
    if (gs) {  // same fields the CPU path fills (core/batch_stats.h:40-77)
      gs->levenbergIterations = st.levenberg_iterations;
      gs->timeIteration = st.time_iteration;
      gs->timeResiduals = st.time_residuals;
      gs->timeQuadraticForm = st.time_quadratic_form;
      gs->timeSchurComplement = st.time_schur;
      gs->timeNumericDecomposition = st.time_numeric;
      gs->timeLinearSolver = st.time_linear_solver;
      gs->timeLinearSolution = st.time_linear_solution;
      gs->timeUpdate = st.time_update;
      gs->timeSymbolicDecomposition = st.time_symbolic;
      gs->choleskyNNZ = static_cast<size_t>(b200_get_factor_nnz(_ctx));
    }
    _lambda = st.lambda;
    _levenbergIterations = st.levenberg_iterations;
    _graph.writeBack(_ctx);
    return static_cast<SolverResult>(st.result);
  }

  // OptimizationAlgorithmWithHessian::computeMarginals -> Solver::computeMarginals (core/block_solver.hpp:490-499)
  virtual bool computeMarginals(SparseBlockMatrix<MatrixXd>& spinv, const std::vector<std::pair<int, int> >& blockIndices) {
    return b200Marginals(_ctx, spinv, blockIndices);
  }
  virtual bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  virtual void printVerbose(std::ostream& os) const {
    os << "\t schur= " << (_graph._hasLandmarks ? 1 : 0) << "\t lambda= " << FIXED(_lambda) << "\t levenbergIter= " << _levenbergIterations;
  }

 protected:
  b200_ctx* _ctx;
  int _algorithm;
  bool _pcg;
  B200GraphBinding _graph;
  double _lambda;
  int _levenbergIterations;
  Property<int>* _device;
  Property<int>* _ndLevels;
  Property<double>* _pcgTolerance;
  Property<bool>* _pcgAbsoluteTolerance;
  Property<int>* _pcgMaxIterations;
  Property<double>* _userLambdaInit;
  Property<int>* _maxTrialsAfterFailure;
};


// ----------------------------------------------------------------------------------------------- Level 2
// g2o::Solver (core/solver.h:44-149) under the stock OptimizationAlgorithm{GaussNewton,Levenberg}.  Those algorithms
// evaluate the errors and apply the update on the host (SparseOptimizer::computeActiveErrors / update / push / pop), so
// the estimates are re-sent before every buildSystem(); they read x() and b() for computeScale
// (optimization_algorithm_levenberg.cpp:165-172) and v->hessian(j,j) for computeLambdaInit (:149-163), which is why x, b
// and the diagonal of every vertex block are mirrored on the host.
class SolverB200 : public Solver {
 public:
  explicit SolverB200(int device = 0, int ndLevels = 0) : Solver(), _ctx(0), _device(device), _ndLevels(ndLevels), _writeDebug(true) {}
  virtual ~SolverB200() { b200_destroy(_ctx); }

  virtual bool init(SparseOptimizer* optimizer, bool /*online*/ = false) {
    _optimizer = optimizer;
    if (!_ctx && b200_create(_device, &_ctx) != B200_OK) {
      std::cerr << "SolverB200: " << b200_last_error(0) << std::endl;  // no CPU fallback
      return false;
    }
    return b200_set_ordering(_ctx, _ndLevels) == B200_OK;
  }
  // BlockSolver::buildStructure (core/block_solver.hpp:142-295): index mapping -> patterns -> symbolic factorisation
  virtual bool buildStructure(bool /*zeroBlocks*/ = false) {
    if (!_graph.ingest(_ctx, _optimizer)) return false;
    int dims[8];
    if (b200_get_dims(_ctx, dims) != B200_OK) return false;
    resizeVector(static_cast<size_t>(dims[2] + dims[3]));
    // host storage of the diagonal blocks, mapped into the vertices like BlockSolver does (block_solver.hpp:181-197)
    const std::vector<OptimizableGraph::Vertex*>& iv = _optimizer->indexMapping();
    size_t total = 0;
    for (size_t i = 0; i < iv.size(); ++i) total += static_cast<size_t>(iv[i]->dimension()) * iv[i]->dimension();
    _diagBlocks.assign(total, 0.);
    _diag.resize(_xSize);
    size_t off = 0;
    for (size_t i = 0; i < iv.size(); ++i) {
      iv[i]->mapHessianMemory(&_diagBlocks[off]);
      off += static_cast<size_t>(iv[i]->dimension()) * iv[i]->dimension();
    }
    return true;
  }
  virtual bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  // BlockSolver::buildSystem (core/block_solver.hpp:501-560) at the estimates the host vertices hold now
  virtual bool buildSystem() {
    if (!_graph.pushEstimates(_ctx) || b200_build_system(_ctx) != B200_OK) return report();
    if (b200_get_b(_ctx, _b) != B200_OK || b200_get_hessian_diagonal(_ctx, &_diag[0]) != B200_OK) return report();
    const std::vector<OptimizableGraph::Vertex*>& iv = _optimizer->indexMapping();
    size_t k = 0;
    for (size_t i = 0; i < iv.size(); ++i)
      for (int j = 0; j < iv[i]->dimension(); ++j) iv[i]->hessian(j, j) = _diag[k++];
    return true;
  }
  virtual bool solve() {
    const int rc = b200_solve(_ctx);
    if (rc < 0) return report();
    if (rc == B200_NOT_POSITIVE_DEFINITE) return false;  // LM raises lambda and tries again
    return b200_get_x(_ctx, _x) == B200_OK;
  }
  virtual bool computeMarginals(SparseBlockMatrix<MatrixXd>& spinv, const std::vector<std::pair<int, int> >& blockIndices) {
    return b200Marginals(_ctx, spinv, blockIndices);
  }
  virtual bool setLambda(double lambda, bool backup = false) { return b200_set_lambda(_ctx, lambda, backup ? 1 : 0) == B200_OK; }
  virtual void restoreDiagonal() { b200_restore_diagonal(_ctx); }
  virtual bool supportsSchur() { return true; }
  virtual bool schur() { return _graph._hasLandmarks; }
  virtual void setSchur(bool) {}  // decided by the graph: marginalized XYZ vertices <=> Schur complement
  virtual void setWriteDebug(bool b) { _writeDebug = b; }
  virtual bool writeDebug() const { return _writeDebug; }
  virtual bool saveHessian(const std::string&) const { return false; }

 protected:
  bool report() {
    std::cerr << "SolverB200: " << b200_last_error(_ctx) << std::endl;
    return false;
  }
  b200_ctx* _ctx;
  int _device, _ndLevels;
  bool _writeDebug;
  B200GraphBinding _graph;
  std::vector<double> _diagBlocks, _diag;
};

// ----------------------------------------------------------------------------------------------- registration
static OptimizationAlgorithm* createSolverB200(const std::string& fullSolverName) {
  const std::string method = fullSolverName.substr(0, 2);
  const std::string rest = fullSolverName.substr(3);
  if (rest == "fix3_2_b200" || rest == "fix6_3_b200" || rest == "var_b200")  // Level 3: whole iteration on the GPU
    return new OptimizationAlgorithmB200(method == "gn" ? B200_GAUSS_NEWTON : B200_LEVENBERG);
  if (rest == "pcg_b200" || rest == "pcg3_2_b200" || rest == "pcg6_3_b200")  // ... with LinearSolverPCG as the linear solver
    return new OptimizationAlgorithmB200(method == "gn" ? B200_GAUSS_NEWTON : B200_LEVENBERG, true);
  // Level 2: stock LM/GN control, errors and update on the host; system, Schur complement and Cholesky on the GPU
  // Level 1: stock BlockSolver + LM/GN on the host, only the linear solver on the GPU
  Solver* s = 0;
  if (rest == "fix3_2_b200s" || rest == "fix6_3_b200s") s = new SolverB200();
  else if (rest == "fix3_2_b200ls") s = new BlockSolver_3_2(new LinearSolverB200<BlockSolver_3_2::PoseMatrixType>());
  else if (rest == "fix6_3_b200ls") s = new BlockSolver_6_3(new LinearSolverB200<BlockSolver_6_3::PoseMatrixType>());
  else return 0;
  if (method == "gn") return new OptimizationAlgorithmGaussNewton(s);
  if (method == "lm") return new OptimizationAlgorithmLevenberg(s);
  delete s;
  return 0;
}

class B200SolverCreator : public AbstractOptimizationAlgorithmCreator {
 public:
  explicit B200SolverCreator(const OptimizationAlgorithmProperty& p) : AbstractOptimizationAlgorithmCreator(p) {}
  virtual OptimizationAlgorithm* construct() { return createSolverB200(property().name); }
};

G2O_REGISTER_OPTIMIZATION_LIBRARY(b200);

#define B200_REGISTER(name, desc, pd, ld) \
  G2O_REGISTER_OPTIMIZATION_ALGORITHM(name, new B200SolverCreator(OptimizationAlgorithmProperty(#name, desc, "B200", true, pd, ld)))

B200_REGISTER(gn_fix3_2_b200, "Gauss-Newton: device-resident solver on B200 (fixed blocksize)", 3, 2);
B200_REGISTER(gn_fix6_3_b200, "Gauss-Newton: device-resident solver on B200 (fixed blocksize)", 6, 3);
B200_REGISTER(lm_fix3_2_b200, "Levenberg: device-resident solver on B200 (fixed blocksize)", 3, 2);
B200_REGISTER(lm_fix6_3_b200, "Levenberg: device-resident solver on B200 (fixed blocksize)", 6, 3);
// variable block sizes (requiresMarginalize = false, like gn_var / lm_var of solvers/csparse/solver_csparse.cpp:53-55,102,108):
// pose graphs and landmark SLAM (SE2 + XY, SE3 + TRACKXYZ) with every vertex in one system - Level 3 only (the padded
// landmark blocks of the C-ABI do not map onto a host-side SparseBlockMatrix<MatrixXd>)
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_var_b200, new B200SolverCreator(OptimizationAlgorithmProperty("gn_var_b200", "Gauss-Newton: device-resident solver on B200 (variable blocksize)", "B200", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_var_b200, new B200SolverCreator(OptimizationAlgorithmProperty("lm_var_b200", "Levenberg: device-resident solver on B200 (variable blocksize)", "B200", false, Eigen::Dynamic, Eigen::Dynamic)));
// block-Jacobi PCG as the linear solver (solvers/pcg/solver_pcg.cpp: gn_pcg, gn_pcg3_2, gn_pcg6_3, lm_pcg, ...)
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_pcg_b200, new B200SolverCreator(OptimizationAlgorithmProperty("gn_pcg_b200", "Gauss-Newton: device-resident solver on B200, PCG (variable blocksize)", "B200", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_pcg_b200, new B200SolverCreator(OptimizationAlgorithmProperty("lm_pcg_b200", "Levenberg: device-resident solver on B200, PCG (variable blocksize)", "B200", false, Eigen::Dynamic, Eigen::Dynamic)));
B200_REGISTER(gn_pcg3_2_b200, "Gauss-Newton: device-resident solver on B200, PCG (fixed blocksize)", 3, 2);
B200_REGISTER(gn_pcg6_3_b200, "Gauss-Newton: device-resident solver on B200, PCG (fixed blocksize)", 6, 3);
B200_REGISTER(lm_pcg3_2_b200, "Levenberg: device-resident solver on B200, PCG (fixed blocksize)", 3, 2);
B200_REGISTER(lm_pcg6_3_b200, "Levenberg: device-resident solver on B200, PCG (fixed blocksize)", 6, 3);
B200_REGISTER(gn_fix3_2_b200s, "Gauss-Newton: B200 solver (system, Schur, Cholesky) under the stock algorithm", 3, 2);
B200_REGISTER(gn_fix6_3_b200s, "Gauss-Newton: B200 solver (system, Schur, Cholesky) under the stock algorithm", 6, 3);
B200_REGISTER(lm_fix3_2_b200s, "Levenberg: B200 solver (system, Schur, Cholesky) under the stock algorithm", 3, 2);
B200_REGISTER(lm_fix6_3_b200s, "Levenberg: B200 solver (system, Schur, Cholesky) under the stock algorithm", 6, 3);
B200_REGISTER(gn_fix3_2_b200ls, "Gauss-Newton: BlockSolver + B200 supernodal Cholesky", 3, 2);
B200_REGISTER(gn_fix6_3_b200ls, "Gauss-Newton: BlockSolver + B200 supernodal Cholesky", 6, 3);
B200_REGISTER(lm_fix3_2_b200ls, "Levenberg: BlockSolver + B200 supernodal Cholesky", 3, 2);
B200_REGISTER(lm_fix6_3_b200ls, "Levenberg: BlockSolver + B200 supernodal Cholesky", 6, 3);

}  // namespace g2o
