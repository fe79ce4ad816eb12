// pcg_host.h - host driver of the block-Jacobi PCG kernels (pcg.cuh), shared by the Level-1 linear solver
// (linear_solver_api.cu: b200_ls_solve_pcg) and the Level-2/3 solver context (solver.cu: b200_set_linear_solver).
// Mirrors LinearSolverPCG<MatrixType> (solvers/pcg/linear_solver_pcg.h:47-98, linear_solver_pcg.hpp:79-160): init() forgets
// the structure and the carried-over absolute residual, solve() builds the "linear structure" of a new pattern once.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "common.h"
#include "pcg.cuh"

namespace g2o_b200 {

class PcgGpu {
 public:
  // LinearSolverPCG::init(): _residual = -1, _indices / _sparseMat cleared (linear_solver_pcg.h:64-71)
  void init() { ready_ = false; residual_ = -1.0; }
  bool matches(int nb, int d, int nblk) const { return ready_ && nb_ == nb && d_ == d && nblk_ == nblk; }
  // symmetric block-row lists (ascending column inside a row) of an upper-triangular block CCS pattern: the "linear
  // structure" of linear_solver_pcg.hpp:86-106.  Returns false (err set) on a malformed pattern.
  bool analyze(int nb, int d, const int* colptr, const int* rowidx, cudaStream_t s, std::string* err) {
    const int nblk = colptr[nb];
    std::vector<int> rowptr(nb + 1, 0), diag(nb, -1);
    for (int c = 0; c < nb; ++c)
      for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
        const int r = rowidx[q];
        if (r < 0 || r > c) { if (err) *err = "upper-triangular block pattern expected (row <= column)"; return false; }
        rowptr[r + 1]++;
        if (r != c) rowptr[c + 1]++; else diag[c] = q;
      }
    for (int c = 0; c < nb; ++c) if (diag[c] < 0) { if (err) *err = "missing diagonal block"; return false; }
    for (int i = 0; i < nb; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int> eb(rowptr[nb]), ec(rowptr[nb]), fill(rowptr.begin(), rowptr.end() - 1);
    // row i first meets its transposed entries (columns c' < i come from block column i: rows r < i, ascending), then
    // its own upper entries in ascending column order: walk the columns in order and append
    for (int c = 0; c < nb; ++c)
      for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
        const int r = rowidx[q];
        if (r != c) { eb[fill[c]] = ~q; ec[fill[c]] = r; fill[c]++; }
      }
    for (int c = 0; c < nb; ++c)
      for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
        const int r = rowidx[q];
        eb[fill[r]] = q; ec[fill[r]] = c; fill[r]++;
      }
    rowptr_.upload(rowptr, s); ent_blk_.upload(eb, s); ent_col_.upload(ec, s); diag_blk_.upload(diag, s);
    const size_t n = (size_t)nb * d;
    J_.alloc((size_t)nb * d * d);
    x_.alloc(n); r_.alloc(n); s_.alloc(n); q_.alloc(n); d0_.alloc(n); d1_.alloc(n);
    partial_.alloc(ceil_div(n, (size_t)kPcgThreads) + 1);
    sc_.alloc(1); status_.alloc(1);
    if (!h_status_ && !host_only_flag()) B200_CUDA(cudaMallocHost((void**)&h_status_, 16 + sizeof(PcgScalars)));
    if (!host_only_flag()) B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
    nb_ = nb; d_ = d; nblk_ = nblk; ready_ = true;
    return true;
  }
  // x = A^-1 b for the device matrix dA (blocks in pattern order, d*d column-major) and the device right-hand side db.
  // tolerance / absolute / max_iter: setTolerance, setAbsoluteTolerance, setMaxIterations (-1: the number of rows).
  // Host round trip every 64 iterations (the stopping rule runs on the device).  Returns B200_OK or
  // B200_NOT_POSITIVE_DEFINITE (a diagonal block that is not positive definite).
  int solve(const double* dA, const double* db, double tolerance, int absolute, int max_iter, cudaStream_t s, LaunchCounter* lc,
            int* iterations, double* residual) {
    return d_ == 3 ? run<3>(dA, db, tolerance, absolute, max_iter, s, lc, iterations, residual)
                   : run<6>(dA, db, tolerance, absolute, max_iter, s, lc, iterations, residual);
  }
  const double* x() const { return x_.p; }
  double carried_residual() const { return residual_; }
  ~PcgGpu() { if (h_status_) cudaFreeHost(h_status_); }

 private:
  template <int D>
  int run(const double* dA, const double* db, double tolerance, int absolute, int max_iter, cudaStream_t s, LaunchCounter* lc,
          int* iterations, double* residual) {
    const int nb = nb_, n = nb * D;
    PcgDev P{nb, rowptr_.p, ent_blk_.p, ent_col_.p, diag_blk_.p, dA, J_.p, x_.p, r_.p, s_.p, q_.p, d0_.p, d1_.p, partial_.p, sc_.p};
    const int grid = (int)ceil_div(n, kPcgThreads);
    B200_CUDA(cudaMemsetAsync(status_.p, 0, sizeof(int), s));
    B200_CUDA(cudaMemsetAsync(sc_.p, 0, sizeof(PcgScalars), s));
    pcg_jacobi_kernel<D><<<(int)ceil_div(nb, 128), 128, 0, s>>>(nb, diag_blk_.p, dA, J_.p, status_.p);
    pcg_init_kernel<D><<<grid, kPcgThreads, 0, s>>>(P, db, tolerance, absolute, residual_, max_iter < 0 ? n : max_iter);
    if (lc) lc->n += 2;
    PcgScalars* h = reinterpret_cast<PcgScalars*>(h_status_ + 2);   // pinned: [status | pad | scalars]
    const int limit = max_iter < 0 ? n : max_iter;
    for (int done_iters = 0;;) {
      // a batch of iterations without a host round trip; after `done` the kernels return at once
      const int batch = std::min(64, std::max(1, limit - done_iters));
      for (int k = 0; k < batch; ++k) {
        pcg_spmv_kernel<D><<<grid, kPcgThreads, 0, s>>>(P);
        pcg_update_kernel<D><<<grid, kPcgThreads, 0, s>>>(P);
      }
      if (lc) lc->n += 2 * batch;
      B200_CUDA(cudaMemcpyAsync(h, sc_.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaMemcpyAsync(h_status_, status_.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      if (*h_status_) return 1 /* B200_NOT_POSITIVE_DEFINITE */;
      done_iters = h->iteration;
      if (h->done || done_iters >= limit) break;
      if (!(h->dn == h->dn)) break;  // NaN: indefinite system, CG broke down
    }
    residual_ = 0.5 * h->dn;
    if (iterations) *iterations = h->iteration;
    if (residual) *residual = residual_;
    return 0;
  }
  bool ready_ = false;
  int nb_ = 0, d_ = 0, nblk_ = 0;
  DevBuf<int> rowptr_, ent_blk_, ent_col_, diag_blk_, status_;
  DevBuf<double> J_, x_, r_, s_, q_, d0_, d1_, partial_;
  DevBuf<PcgScalars> sc_;
  double residual_ = -1.0;
  int* h_status_ = nullptr;
};

}  // namespace g2o_b200
