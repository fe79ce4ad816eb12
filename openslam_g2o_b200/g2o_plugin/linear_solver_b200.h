// linear_solver_b200.h - Level-1 adapter: g2o::LinearSolver<MatrixType> over the C-ABI (include/g2o_b200.h).
//
// Drop-in for LinearSolverCSparse / LinearSolverCholmod (g2o/solvers/{csparse,cholmod}): same class contract
// (g2o/core/linear_solver.h:40-81).  Compiles only inside a g2o source tree (needs Eigen + g2o headers, neither of
// which exists in the build container - see DESIGN.md section 2); everything numerical lives in libg2o_b200.so.
#ifndef G2O_LINEAR_SOLVER_B200_H
#define G2O_LINEAR_SOLVER_B200_H

#include <iostream>
#include <vector>

#include "g2o/core/batch_stats.h"
#include "g2o/core/linear_solver.h"
#include "g2o/core/matrix_structure.h"
#include "g2o/stuff/timeutil.h"
#include "g2o_b200.h"

namespace g2o {

template <typename MatrixType>
class LinearSolverB200 : public LinearSolver<MatrixType> {
 public:
  LinearSolverB200() : LinearSolver<MatrixType>(), _ls(0), _device(0), _writeDebug(true) {}
  virtual ~LinearSolverB200() { b200_ls_destroy(_ls); }

  //! GPU ordinal; takes effect at the next init()
  void setDevice(int device) { _device = device; }

  virtual bool init() {
    if (!_ls && b200_ls_create(_device, &_ls) != B200_OK) {
      std::cerr << "LinearSolverB200: " << b200_ls_last_error(0) << std::endl;  // no CPU fallback
      return false;
    }
    return b200_ls_init(_ls) == B200_OK;  // drops the symbolic factor like LinearSolverCSparse::init
  }

  bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b) {
    if (!_ls && !init()) return false;
    const int nBlocks = static_cast<int>(A.blockCols().size());
    const int d = A.colsOfBlock(0);
    for (int i = 0; i < nBlocks; ++i)
      if (A.colsOfBlock(i) != d) {
        std::cerr << "LinearSolverB200: only uniform block sizes (3 or 6) are supported" << std::endl;
        return false;
      }
    // upper block pattern incl. diagonal, ascending rows (core/sparse_block_matrix.hpp:519-545) ...
    A.fillBlockStructure(_structure);
    // ... and the block values in the same order, column-major like Eigen stores them
    _values.resize(static_cast<size_t>(_structure.Ap[nBlocks]) * d * d);
    double* out = &_values[0];
    for (int c = 0; c < nBlocks; ++c) {
      const typename SparseBlockMatrix<MatrixType>::IntBlockMap& col = A.blockCols()[c];
      for (typename SparseBlockMatrix<MatrixType>::IntBlockMap::const_iterator it = col.begin(); it != col.end(); ++it) {
        if (it->first > c) break;
        memcpy(out, it->second->data(), sizeof(double) * d * d);
        out += d * d;
      }
    }
    double t = get_monotonic_time();
    int rc = b200_ls_solve(_ls, nBlocks, d, _structure.Ap, _structure.Aii, &_values[0], x, b);
    if (rc < 0) std::cerr << "LinearSolverB200: " << b200_ls_last_error(_ls) << std::endl;
    G2OBatchStatistics* globalStats = G2OBatchStatistics::globalStats();
    if (globalStats) {
      globalStats->timeNumericDecomposition = get_monotonic_time() - t;
      globalStats->choleskyNNZ = static_cast<size_t>(b200_ls_get_factor_nnz(_ls));
    }
    return rc == B200_OK;  // B200_NOT_POSITIVE_DEFINITE -> false, exactly like the CSparse / CHOLMOD solvers
  }

  virtual bool writeDebug() const { return _writeDebug; }
  virtual void setWriteDebug(bool b) { _writeDebug = b; }

 protected:
  b200_linear_solver* _ls;
  int _device;
  bool _writeDebug;
  MatrixStructure _structure;
  std::vector<double> _values;
};

// Drop-in for LinearSolverPCG (g2o/solvers/pcg/linear_solver_pcg.h:47-98): same setters, same stopping rule, the
// iteration on the GPU (csrc/pcg.cuh).  Uncompiled here for the same reason as the class above.
template <typename MatrixType>
class LinearSolverB200PCG : public LinearSolverB200<MatrixType> {
 public:
  LinearSolverB200PCG() : LinearSolverB200<MatrixType>(), _tolerance(1e-6), _absoluteTolerance(true), _verbose(false), _maxIter(-1) {}

  bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b) {
    if (!this->_ls && !this->init()) return false;
    const int nBlocks = static_cast<int>(A.blockCols().size());
    const int d = A.colsOfBlock(0);
    for (int i = 0; i < nBlocks; ++i)
      if (A.colsOfBlock(i) != d) {
        std::cerr << "LinearSolverB200PCG: only uniform block sizes (3 or 6) are supported" << std::endl;
        return false;
      }
    A.fillBlockStructure(this->_structure);
    this->_values.resize(static_cast<size_t>(this->_structure.Ap[nBlocks]) * d * d);
    double* out = &this->_values[0];
    for (int c = 0; c < nBlocks; ++c) {
      const typename SparseBlockMatrix<MatrixType>::IntBlockMap& col = A.blockCols()[c];
      for (typename SparseBlockMatrix<MatrixType>::IntBlockMap::const_iterator it = col.begin(); it != col.end(); ++it) {
        if (it->first > c) break;
        memcpy(out, it->second->data(), sizeof(double) * d * d);
        out += d * d;
      }
    }
    int iterations = 0;
    double residual = 0.;
    int rc = b200_ls_solve_pcg(this->_ls, nBlocks, d, this->_structure.Ap, this->_structure.Aii, &this->_values[0], x, b, _tolerance,
                               _absoluteTolerance ? 1 : 0, _maxIter, &iterations, &residual);
    if (rc < 0) std::cerr << "LinearSolverB200PCG: " << b200_ls_last_error(this->_ls) << std::endl;
    if (_verbose) std::cerr << "residual[" << iterations << "]: " << 2. * residual << std::endl;
    G2OBatchStatistics* globalStats = G2OBatchStatistics::globalStats();
    if (globalStats) globalStats->iterationsLinearSolver = iterations;
    return rc == B200_OK;
  }

  double tolerance() const { return _tolerance; }
  void setTolerance(double tolerance) { _tolerance = tolerance; }
  int maxIterations() const { return _maxIter; }
  void setMaxIterations(int maxIter) { _maxIter = maxIter; }
  bool absoluteTolerance() const { return _absoluteTolerance; }
  void setAbsoluteTolerance(bool absoluteTolerance) { _absoluteTolerance = absoluteTolerance; }
  bool verbose() const { return _verbose; }
  void setVerbose(bool verbose) { _verbose = verbose; }

 protected:
  double _tolerance;
  bool _absoluteTolerance, _verbose;
  int _maxIter;
};

}  // namespace g2o
#endif
