// chol.h - supernodal left-looking sparse block Cholesky on the GPU (host-side interface).
//
// Replaces the numeric phase of the reference's linear solvers for the reduced pose system:
// LinearSolverCSparse::solve -> cs_cholsolsymb -> cs_chol_workspace + cs_lsolve/cs_ltsolve
// (solvers/csparse/linear_solver_csparse.h:106-142, solvers/csparse/csparse_helper.cpp:56-143) and
// LinearSolverCholmod::solve -> cholmod_factorize + cholmod_solve
// (solvers/cholmod/linear_solver_cholmod.h:115-154).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "common.h"
#include "symbolic.h"

namespace g2o_b200 {

struct CholDev;
struct CholPlanDev;

class CholeskyGpu {
 public:
  CholeskyGpu() = default;
  ~CholeskyGpu();
  // host: ordering + symbolic analysis; uploads the plan.  colptr/rowidx = upper block pattern.
  void analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt, cudaStream_t s);
  bool analyzed() const { return analyzed_; }
  void reset() { analyzed_ = false; }
  const SymbolicFactor& symbolic() const { return S_; }

  // numeric factorisation of A + lambda*I (A: device, input block order, d*d col-major per block) and, fused into
  // it, the forward substitution of the right-hand side d_b (the supernode's right-hand side rides along as one more
  // panel row).  d_lambda may be nullptr (lambda = 0).  Asynchronous on s: a few setup kernels + ONE persistent
  // dataflow kernel; the not-positive-definite outcome lands in the device flag behind status_ptr().
  void factor(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
              EventProfiler* prof = nullptr);
  // x = A^-1 b for the b given to the preceding factor() (device, length nb*d, original ordering): the backward
  // sweep (one persistent dataflow kernel) + un-permutation.  One solve() per factor().  Asynchronous on s.
  void solve(const double* d_b, double* d_x, cudaStream_t s, LaunchCounter* lc, EventProfiler* prof = nullptr);
  // extra[i] (device, length nb*d, ORIGINAL ordering; nullptr = none) is added to diagonal entry i by every factor():
  // the unit diagonal of padding unknowns (landmark blocks padded to the pose dimension, solver.cu)
  void set_diagonal_extra(const double* d_extra) { d_diag_extra_ = d_extra; }
  // the next factor() calls also leave the inverse diagonal blocks of the tail-chain links where sparse_inverse() reads
  // them (the solves take them from the chain's packed records): on for marginals, off on the LM path (measured: the
  // second copy costs 15 us per factorisation on the Venice-shaped graph)
  void keep_chain_inverses(bool on) { keep_chain_inverses_ = on; }
  int* status_ptr() { return d_counters_.p + 2; }  // device int: 0 ok, 1 not positive definite
  double* factor_values() { return d_L_.p; }
  // Sparse inverse subset of the matrix of the preceding factor(): every block of A^-1 on the pattern of L + L^T
  // (Takahashi recursion down the supernodal tree, sparse_inverse.cuh) - what MarginalCovarianceCholesky
  // (core/marginal_covariance_cholesky.cpp:55-214) evaluates entry by entry.  Asynchronous on s.
  void sparse_inverse(cudaStream_t s, LaunchCounter* lc);
  // host: where block (r, c) of A^-1 (ORIGINAL block indices) lives in the array sparse_inverse() filled; false when the
  // block is outside the pattern of the factor
  bool locate_inverse_block(int r, int c, long long* off, int* ld, bool* transposed) const;
  // device gather of n located blocks into d_out (n * d * d doubles, column-major blocks).  Asynchronous on s.
  void gather_inverse_blocks(int n, const long long* d_off, const int* d_ld, const unsigned char* d_trans, double* d_out,
                             cudaStream_t s, LaunchCounter* lc);

 private:
  bool analyzed_ = false;
  const double* d_diag_extra_ = nullptr;
  bool keep_chain_inverses_ = false;
  SymbolicFactor S_;
  DevBuf<int> d_sn_col0_, d_sn_ncol_, d_sn_nrow_, d_sn_rowptr_, d_sn_rows_;
  DevBuf<long long> d_sn_lptr_;
  DevBuf<int> d_rel_;
  DevBuf<int> d_task_ptr_, d_task_sn_, d_task_parent_;
  DevBuf<long long> d_a_dst_, d_diag_dst_;
  DevBuf<int> d_a_ld_, d_diag_ld_, d_perm_;
  DevBuf<unsigned char> d_a_trans_;
  DevBuf<int> d_tile_sn_, d_tile_r0_, d_tile_c0_, d_tile_work_ptr_, d_work_a0_, d_work_a1_, d_work_b0_, d_work_b1_;
  DevBuf<int> d_sn_tile_ptr_, d_sn_chunk_ptr_, d_chunk_sn_, d_chunk_b0_, d_chunk_nb_;
  DevBuf<int> d_group_tile_, d_group_w0_, d_group_w1_, d_group_slot_, d_group_rtile_, d_rtile_tile_, d_rtile_slot0_, d_rtile_nslots_;
  DevBuf<long long> d_sn_dinvptr_, d_sn_cptr_, d_work_koff_, d_work_reloff_;
  DevBuf<int> d_work_mk_, d_work_nk_, d_work_ksn_, d_fwd_ptr_, d_fwd_src_;
  DevBuf<int> d_flow_kind_, d_flow_arg_, d_sn_nupd_, d_sn_nchunk_;
  DevBuf<double> d_L_, d_Ldiag_, d_Dinv_, d_y_, d_z_, d_gscratch_, d_contrib_;
  DevBuf<int> d_counters_;  // task counters, status flag, completion counters (zeroed by every factor())
  // tail chain (chol_chain.cuh)
  DevBuf<int> d_chain_sn_, d_chain_mapptr_, d_chain_map_, d_chain_colptr_, d_chain_fwd_ptr_, d_chain_fwd_src_, d_chain_desc_;
  DevBuf<unsigned> d_chain_new_rows_;
  DevBuf<double> d_chain_pack_;  // per chain link: rows below the diagonal block (packed) | inverse diagonal block
  DevBuf<unsigned char> d_task_skip_;
  // sparse inverse (built at the first sparse_inverse() after analyze()): Z and Y^T with the geometry of L, work items
  // (supernode, block row) and supernodes per depth level of the supernodal tree
  DevBuf<double> d_Zinv_, d_Yt_;
  DevBuf<int> d_col2sn_, d_spinv_item_sn_, d_spinv_item_p_, d_spinv_level_sn_;
  std::vector<int> spinv_item_ptr_, spinv_level_ptr_;
  bool spinv_planned_ = false;
  size_t chain_smem_ = 0, chain_back_smem_ = 0;
  int chain_back_buf_doubles_ = 0;
  int cnt_upd_ = 0, cnt_chunk_ = 0, cnt_slot_ = 0, cnt_bdone_ = 0;
  size_t flow_smem_ = 0;
  int xb_doubles_ = 0, stage_doubles_ = 0, flow_grid_ = 1, back_grid_ = 1;
  int nblk_ = 0;
  CholDev dev() const;
  CholPlanDev plan() const;
  template <int D>
  void factor_t(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
                EventProfiler* prof);
  template <int D>
  void solve_t(double* x, cudaStream_t s, LaunchCounter* lc, EventProfiler* prof);
};

}  // namespace g2o_b200
