// linear_solver_api.cu - Level-1 boundary: g2o::LinearSolver<MatrixType> (core/linear_solver.h:40-81).
// The adapter flattens SparseBlockMatrix<MatrixType> (upper-triangular block CCS, core/sparse_block_matrix.h)
// into plain arrays; this file owns the device-side life cycle: analyse at the first solve after init(),
// then per call H2D(values,b) -> factor -> solve -> D2H(x), like LinearSolverCSparse::solve
// (solvers/csparse/linear_solver_csparse.h:106-142).
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "../../include/g2o_b200.h"
#include "block_amd.h"
#include "chol.h"
#include "pcg_host.h"

using namespace g2o_b200;

struct b200_linear_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  CholeskyGpu chol;
  DevBuf<double> dA, db, dx;
  LaunchCounter lc;
  int nb = 0, d = 0, nblk = 0;
  int* h_status = nullptr;
  PcgGpu pcg;   // LinearSolverPCG state (pcg_host.h): lists of the current pattern, vectors, _residual of the last solve
};

namespace {
std::string g_ls_error;
template <typename F>
int ls_guarded(b200_linear_solver* ls, F&& f) {
  if (!ls) return B200_ERR_INVALID;
  try {
    return f();
  } catch (const CudaError& e) {
    ls->err = describe(e);
    return (e.code == cudaErrorNoDevice || e.code == cudaErrorInsufficientDriver) ? B200_ERR_NO_DEVICE : B200_ERR_CUDA;
  } catch (const std::exception& e) {
    ls->err = e.what();
    return B200_ERR_INVALID;
  }
}
}  // namespace

extern "C" {

int b200_ls_create(int device, b200_linear_solver** out) {
  if (!out) return B200_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    g_ls_error = "no CUDA device available: the B200 linear solver has no CPU fallback";
    return B200_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { g_ls_error = "invalid device index"; return B200_ERR_INVALID; }
  b200_linear_solver* ls = new b200_linear_solver();
  ls->device = device;
  try {
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaStreamCreateWithFlags(&ls->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaMallocHost((void**)&ls->h_status, sizeof(int)));
  } catch (const CudaError& e) {
    g_ls_error = describe(e);
    delete ls;
    return B200_ERR_CUDA;
  }
  *out = ls;
  return B200_OK;
}

void b200_ls_destroy(b200_linear_solver* ls) {
  if (!ls) return;
  cudaSetDevice(ls->device);
  if (ls->stream) cudaStreamSynchronize(ls->stream);
  if (ls->h_status) cudaFreeHost(ls->h_status);
  cudaStream_t s = ls->stream;
  delete ls;
  if (s) cudaStreamDestroy(s);
}

int b200_ls_init(b200_linear_solver* ls) {
  if (!ls) return B200_ERR_INVALID;
  ls->chol.reset();
  ls->pcg.init();             // LinearSolverPCG::init(): _residual = -1, _indices / _sparseMat cleared (linear_solver_pcg.h:64-71)
  return B200_OK;
}

int b200_ls_solve_pcg(b200_linear_solver* ls, int nblocks, int block_dim, const int32_t* colptr, const int32_t* rowidx,
                      const double* values, double* x, const double* b, double tolerance, int absolute_tolerance,
                      int max_iterations, int32_t* iterations, double* residual) {
  return ls_guarded(ls, [&]() -> int {
    if (nblocks <= 0 || (block_dim != 3 && block_dim != 6) || !colptr || !rowidx || !values || !x || !b) {
      ls->err = "invalid arguments (block_dim must be 3 or 6)";
      return B200_ERR_INVALID;
    }
    B200_CUDA(cudaSetDevice(ls->device));
    cudaStream_t s = ls->stream;
    const int nblk = colptr[nblocks], d = block_dim;
    if (!ls->pcg.matches(nblocks, d, nblk)) {
      const double carried = ls->pcg.carried_residual();
      (void)carried;  // a new pattern keeps _residual (only init() resets it): PcgGpu::analyze does not touch it
      if (!ls->pcg.analyze(nblocks, d, colptr, rowidx, s, &ls->err)) return B200_ERR_INVALID;
    }
    const size_t n = (size_t)nblocks * d;
    ls->dA.upload(values, (size_t)nblk * d * d, s);
    ls->db.upload(b, n, s);
    int it = 0;
    double res = 0.0;
    const int rc = ls->pcg.solve(ls->dA.p, ls->db.p, tolerance, absolute_tolerance, max_iterations, s, &ls->lc, &it, &res);
    if (rc != B200_OK) return rc;
    B200_CUDA(cudaMemcpyAsync(x, ls->pcg.x(), n * sizeof(double), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (iterations) *iterations = it;
    if (residual) *residual = res;
    return B200_OK;
  });
}

int b200_ls_solve(b200_linear_solver* ls, int nblocks, int block_dim, const int32_t* colptr, const int32_t* rowidx,
                  const double* values, double* x, const double* b) {
  return ls_guarded(ls, [&]() -> int {
    if (nblocks <= 0 || (block_dim != 3 && block_dim != 6) || !colptr || !rowidx || !values || !x || !b) {
      ls->err = "invalid arguments (block_dim must be 3 or 6)";
      return B200_ERR_INVALID;
    }
    B200_CUDA(cudaSetDevice(ls->device));
    cudaStream_t s = ls->stream;
    const int nblk = colptr[nblocks];
    if (!ls->chol.analyzed() || ls->nb != nblocks || ls->d != block_dim || ls->nblk != nblk) {
      SymbolicOptions opt;
      if (const char* e = getenv("G2O_B200_ND_LEVELS")) opt.nd_levels = std::max(0, atoi(e));  // see b200_set_ordering
      if (const char* e = getenv("G2O_B200_CHAIN")) opt.chain = atoi(e) != 0;
      if (const char* e = getenv("G2O_B200_WIDE_TILES")) opt.wide_tiles = atoi(e);
      ls->chol.analyze(nblocks, block_dim, colptr, rowidx, opt, s);
      ls->nb = nblocks; ls->d = block_dim; ls->nblk = nblk;
    }
    const size_t n = (size_t)nblocks * block_dim;
    ls->dA.upload(values, (size_t)nblk * block_dim * block_dim, s);
    ls->db.upload(b, n, s);
    ls->dx.alloc(n);
    ls->chol.factor(ls->dA.p, nullptr, ls->db.p, s, &ls->lc);
    ls->chol.solve(ls->db.p, ls->dx.p, s, &ls->lc);
    B200_CUDA(cudaMemcpyAsync(ls->h_status, ls->chol.status_ptr(), sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (*ls->h_status) return B200_NOT_POSITIVE_DEFINITE;
    B200_CUDA(cudaMemcpyAsync(x, ls->dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    return B200_OK;
  });
}

int b200_ls_get_block_ordering(b200_linear_solver* ls, int32_t* perm) {
  if (!ls || !ls->chol.analyzed()) return B200_ERR_INVALID;
  const std::vector<int>& P = ls->chol.symbolic().perm;
  if (perm) memcpy(perm, P.data(), P.size() * sizeof(int));
  return (int)P.size();
}
int64_t b200_ls_get_factor_nnz(b200_linear_solver* ls) {
  return (ls && ls->chol.analyzed()) ? ls->chol.symbolic().scalar_lnz : -1;
}
const char* b200_ls_last_error(const b200_linear_solver* ls) { return ls ? ls->err.c_str() : g_ls_error.c_str(); }

int b200_block_amd(int nblocks, const int32_t* colptr, const int32_t* rowidx, int32_t* perm) {
  if (nblocks < 0 || !colptr || !rowidx || !perm) return B200_ERR_INVALID;
  std::vector<int> p = g2o_b200::block_amd(nblocks, colptr, rowidx);
  memcpy(perm, p.data(), p.size() * sizeof(int));
  return B200_OK;
}

}  // extern "C"
