// chol.cu - supernodal left-looking sparse block Cholesky, sm_100a kernels + host driver.  See chol.h.
//
// Layout in HBM: L is one array of doubles; supernode s owns a dense column-major panel of
// (nrow_s*d) x (ncol_s*d) at sn_lptr[s] whose first ncol_s block rows are the (lower-triangular)
// diagonal block.  All index arrays are 32-bit block indices; panel offsets are 64-bit.
//
// Determinism: every panel is written by exactly one CTA and the updates it pulls from its
// descendants are applied in a fixed order, so repeated factorizations are bit-identical (no atomics).
#include "chol.h"

#include <algorithm>

namespace g2o_b200 {

struct CholDev {
  const int *sn_col0, *sn_ncol, *sn_nrow, *sn_rowptr, *sn_rows;
  const long long* sn_lptr;
  const int *upd_ptr, *upd_k, *upd_p0, *upd_p1;
  const long long* upd_relptr;
  const int* rel;
  const int *task_ptr, *task_sn;
};

// ---------------------------------------------------------------------------------------------
// scatter A (+ lambda on the diagonal) into the zeroed panels
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void chol_scatter_kernel(const double* __restrict__ A, int nblk, const long long* __restrict__ dst,
                                    const int* __restrict__ ld, const unsigned char* __restrict__ trans,
                                    double* __restrict__ L) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblk * D * D) return;
  const int k = idx / (D * D);
  const int e = idx - k * D * D;
  const int c = e / D, r = e - c * D;
  const double v = A[idx];  // element (r,c) of block k
  const long long base = dst[k];
  const int l = ld[k];
  if (trans[k]) L[base + c + (long long)r * l] = v;
  else L[base + r + (long long)c * l] = v;
}

template <int D>
__global__ void chol_add_lambda_kernel(int nb, const long long* __restrict__ diag_dst, const int* __restrict__ diag_ld,
                                       const double* __restrict__ lambda, double* __restrict__ L) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * D) return;
  const int k = idx / D, r = idx - k * D;
  L[diag_dst[k] + r + (long long)r * diag_ld[k]] += *lambda;
}

// ---------------------------------------------------------------------------------------------
// numeric factorisation: one CTA per task (a task = sequence of supernodes, children first)
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ void supernode_pull_updates(const CholDev& P, double* __restrict__ L, int J) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int M = P.sn_nrow[J] * D;
  const int col0 = P.sn_col0[J];
  double* Pj = L + P.sn_lptr[J];
  const int u0 = P.upd_ptr[J], u1 = P.upd_ptr[J + 1];
  for (int u = u0; u < u1; ++u) {
    const int K = P.upd_k[u], p0 = P.upd_p0[u], p1 = P.upd_p1[u];
    const int nrK = P.sn_nrow[K];
    const int h = nrK - p0, w = p1 - p0;
    const int Mk = nrK * D, Nk = P.sn_ncol[K] * D;
    const double* Kp = L + P.sn_lptr[K];
    const int* krows = P.sn_rows + P.sn_rowptr[K];
    const int* rel = P.rel + P.upd_relptr[u];
    const int ntile = h * w;
    for (int idx = tid; idx < ntile; idx += nt) {
      const int b = idx / h;
      const int a = idx - b * h;
      if (a < b) continue;
      const double* ra = Kp + (p0 + a) * D;
      const double* rb = Kp + (p0 + b) * D;
      double acc[D][D];
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) acc[r][c] = 0.0;
      for (int k = 0; k < Nk; ++k) {
        double av[D], bv[D];
#pragma unroll
        for (int r = 0; r < D; ++r) { av[r] = ra[r]; bv[r] = rb[r]; }
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int c = 0; c < D; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        ra += Mk;
        rb += Mk;
      }
      const int lr = rel[a];
      const int lc = krows[p0 + b] - col0;
      double* dst = Pj + (long long)lr * D + (long long)(lc * D) * M;
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) dst[r + (long long)c * M] -= acc[r][c];
    }
    __syncthreads();
  }
}

// dense right-looking Cholesky of the N leading columns of an M x N panel (in place), CTA-wide
__device__ void panel_factor(double* __restrict__ Pj, int M, int N, int* status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < N; ++j) {
    __syncthreads();
    double djj = Pj[j + (long long)j * M];
    if (!(djj > 0.0)) {  // cs_chol_workspace: "if (d <= 0) not pos def" (csparse_helper.cpp:136); NaN too
      if (tid == 0) *status = 1;
      djj = 1.0;
    }
    const double s = sqrt(djj);
    __syncthreads();
    double* cj = Pj + (long long)j * M;
    for (int i = j + tid; i < M; i += nt) cj[i] = (i == j) ? s : cj[i] / s;
    __syncthreads();
    const int rows = M - j - 1;
    const int cols = N - j - 1;
    const long long total = (long long)rows * cols;
    for (long long idx = tid; idx < total; idx += nt) {
      const int c = (int)(idx / rows);
      const int i = (int)(idx - (long long)c * rows);
      if (i < c) continue;
      const int gi = j + 1 + i, gc = j + 1 + c;
      Pj[gi + (long long)gc * M] = fma(-cj[gi], cj[gc], Pj[gi + (long long)gc * M]);
    }
  }
  __syncthreads();
}

template <int D>
__global__ void chol_factor_kernel(CholDev P, double* __restrict__ L, int task0, int* status) {
  const int t = task0 + blockIdx.x;
  const int q0 = P.task_ptr[t], q1 = P.task_ptr[t + 1];
  for (int q = q0; q < q1; ++q) {
    const int J = P.task_sn[q];
    supernode_pull_updates<D>(P, L, J);
    panel_factor(L + P.sn_lptr[J], P.sn_nrow[J] * D, P.sn_ncol[J] * D, status);
  }
}

// ---------------------------------------------------------------------------------------------
// triangular solves on the permuted vector y (in place)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void chol_permute_in_kernel(int nb, const int* __restrict__ perm, const double* __restrict__ b,
                                       double* __restrict__ y) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * D) return;
  const int k = idx / D, r = idx - k * D;
  y[idx] = b[perm[k] * D + r];
}
template <int D>
__global__ void chol_permute_out_kernel(int nb, const int* __restrict__ perm, const double* __restrict__ y,
                                        double* __restrict__ x, const int* __restrict__ status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * D) return;
  if (*status != 0) return;  // failed factorisation: leave x untouched (reference keeps the stale _x)
  const int k = idx / D, r = idx - k * D;
  x[perm[k] * D + r] = y[idx];
}

template <int D>
__global__ void chol_forward_kernel(CholDev P, const double* __restrict__ L, double* __restrict__ y, int task0) {
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int q0 = P.task_ptr[t], q1 = P.task_ptr[t + 1];
  for (int q = q0; q < q1; ++q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
    double* yj = y + (long long)col0 * D;
    for (int u = P.upd_ptr[J]; u < P.upd_ptr[J + 1]; ++u) {
      const int K = P.upd_k[u], p0 = P.upd_p0[u], p1 = P.upd_p1[u];
      const int Mk = P.sn_nrow[K] * D, Nk = P.sn_ncol[K] * D;
      const double* Kp = L + P.sn_lptr[K];
      const int* krows = P.sn_rows + P.sn_rowptr[K];
      const double* yk = y + (long long)P.sn_col0[K] * D;
      const int nrow = (p1 - p0) * D;
      for (int i = tid; i < nrow; i += nt) {
        const int p = p0 + i / D, rr = i % D;
        const double* lrow = Kp + p * D + rr;
        double s = 0.0;
        for (int k = 0; k < Nk; ++k) s = fma(lrow[(long long)k * Mk], yk[k], s);
        yj[(krows[p] - col0) * D + rr] -= s;
      }
      __syncthreads();
    }
    const double* Pj = L + P.sn_lptr[J];
    for (int j = 0; j < N; ++j) {
      __syncthreads();
      const double v = yj[j] / Pj[j + (long long)j * M];
      __syncthreads();
      if (tid == 0) yj[j] = v;
      for (int i = j + 1 + tid; i < N; i += nt) yj[i] = fma(-Pj[i + (long long)j * M], v, yj[i]);
    }
    __syncthreads();
  }
}

template <int D>
__global__ void chol_backward_kernel(CholDev P, const double* __restrict__ L, double* __restrict__ y, int task0) {
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int q0 = P.task_ptr[t], q1 = P.task_ptr[t + 1];
  for (int q = q1 - 1; q >= q0; --q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
    const double* Pj = L + P.sn_lptr[J];
    const int* jrows = P.sn_rows + P.sn_rowptr[J];
    double* xj = y + (long long)col0 * D;
    for (int j = tid; j < N; j += nt) {
      const double* cj = Pj + (long long)j * M;
      double s = 0.0;
      for (int i = N; i < M; ++i) s = fma(cj[i], y[(long long)jrows[i / D] * D + (i % D)], s);
      xj[j] -= s;
    }
    for (int j = N - 1; j >= 0; --j) {
      __syncthreads();
      const double v = xj[j] / Pj[j + (long long)j * M];
      __syncthreads();
      if (tid == 0) xj[j] = v;
      for (int i = tid; i < j; i += nt) xj[i] = fma(-Pj[j + (long long)i * M], v, xj[i]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
CholeskyGpu::~CholeskyGpu() {}

void CholeskyGpu::analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt,
                          cudaStream_t s) {
  S_ = g2o_b200::analyze(nb, d, colptr, rowidx, opt);
  nblk_ = colptr[nb];
  d_sn_col0_.upload(S_.sn_col0, s); d_sn_ncol_.upload(S_.sn_ncol, s); d_sn_nrow_.upload(S_.sn_nrow, s);
  d_sn_rowptr_.upload(S_.sn_rowptr, s); d_sn_rows_.upload(S_.sn_rows, s);
  std::vector<long long> tmp(S_.sn_lptr.begin(), S_.sn_lptr.end());
  d_sn_lptr_.upload(tmp, s);
  d_upd_ptr_.upload(S_.upd_ptr, s); d_upd_k_.upload(S_.upd_k, s); d_upd_p0_.upload(S_.upd_p0, s);
  d_upd_p1_.upload(S_.upd_p1, s); d_rel_.upload(S_.rel, s);
  std::vector<long long> tmp2(S_.upd_relptr.begin(), S_.upd_relptr.end());
  d_upd_relptr_.upload(tmp2, s);
  d_task_ptr_.upload(S_.task_ptr, s); d_task_sn_.upload(S_.task_sn, s);
  std::vector<long long> tmp3(S_.a_dst.begin(), S_.a_dst.end());
  d_a_dst_.upload(tmp3, s);
  std::vector<long long> tmp4(S_.diag_dst.begin(), S_.diag_dst.end());
  d_diag_dst_.upload(tmp4, s);
  d_a_ld_.upload(S_.a_ld, s); d_diag_ld_.upload(S_.diag_ld, s); d_perm_.upload(S_.perm, s);
  d_a_trans_.upload(S_.a_trans, s);
  d_L_.alloc((size_t)S_.factor_doubles);
  d_y_.alloc((size_t)nb * d);
  d_status_.alloc(1);
  if (!host_only_flag()) B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
  // CTA size per level from the largest tile count any of its supernodes sees
  level_threads_.assign(S_.nlevels, 128);
  for (int l = 0; l < S_.nlevels; ++l) {
    long long big = 0;
    for (int t = S_.level_ptr[l]; t < S_.level_ptr[l + 1]; ++t)
      for (int q = S_.task_ptr[t]; q < S_.task_ptr[t + 1]; ++q) {
        int J = S_.task_sn[q];
        big = std::max<long long>(big, (long long)S_.sn_nrow[J] * S_.sn_ncol[J] * d);
      }
    level_threads_[l] = big > 4096 ? 512 : big > 512 ? 256 : 128;
  }
  analyzed_ = true;
}

template <int D>
static void factor_t(const SymbolicFactor& S, const CholDev& P, const std::vector<int>& lt, int nblk,
                     const double* dA, const double* d_lambda, const long long* a_dst, const int* a_ld,
                     const unsigned char* a_trans, const long long* diag_dst, const int* diag_ld, double* L, int* status,
                     cudaStream_t s, LaunchCounter* lc) {
  B200_CUDA(cudaMemsetAsync(L, 0, (size_t)S.factor_doubles * sizeof(double), s));
  B200_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  chol_scatter_kernel<D><<<ceil_div((int64_t)nblk * D * D, 256), 256, 0, s>>>(dA, nblk, a_dst, a_ld, a_trans, L);
  if (lc) lc->n++;
  if (d_lambda) {
    chol_add_lambda_kernel<D><<<ceil_div((int64_t)S.nb * D, 256), 256, 0, s>>>(S.nb, diag_dst, diag_ld, d_lambda, L);
    if (lc) lc->n++;
  }
  for (int l = 0; l < S.nlevels; ++l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    chol_factor_kernel<D><<<nt, lt[l], 0, s>>>(P, L, t0, status);
    if (lc) lc->n++;
  }
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::factor(const double* dA, const double* d_lambda, cudaStream_t s, LaunchCounter* lc) {
  CholDev P{d_sn_col0_.p, d_sn_ncol_.p, d_sn_nrow_.p, d_sn_rowptr_.p, d_sn_rows_.p, d_sn_lptr_.p, d_upd_ptr_.p,
            d_upd_k_.p,   d_upd_p0_.p,  d_upd_p1_.p,  d_upd_relptr_.p, d_rel_.p,    d_task_ptr_.p, d_task_sn_.p};
  if (S_.d == 3)
    factor_t<3>(S_, P, level_threads_, nblk_, dA, d_lambda, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, d_diag_dst_.p,
                d_diag_ld_.p, d_L_.p, d_status_.p, s, lc);
  else
    factor_t<6>(S_, P, level_threads_, nblk_, dA, d_lambda, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, d_diag_dst_.p,
                d_diag_ld_.p, d_L_.p, d_status_.p, s, lc);
}

template <int D>
static void solve_t(const SymbolicFactor& S, const CholDev& P, const int* perm, const double* L, double* y,
                    const double* b, double* x, const int* status, cudaStream_t s, LaunchCounter* lc) {
  const int n = S.nb * D;
  chol_permute_in_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, perm, b, y);
  if (lc) lc->n++;
  for (int l = 0; l < S.nlevels; ++l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    chol_forward_kernel<D><<<nt, 128, 0, s>>>(P, L, y, t0);
    if (lc) lc->n++;
  }
  for (int l = S.nlevels - 1; l >= 0; --l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    chol_backward_kernel<D><<<nt, 128, 0, s>>>(P, L, y, t0);
    if (lc) lc->n++;
  }
  chol_permute_out_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, perm, y, x, status);
  if (lc) lc->n++;
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::solve(const double* d_b, double* d_x, cudaStream_t s, LaunchCounter* lc) {
  CholDev P{d_sn_col0_.p, d_sn_ncol_.p, d_sn_nrow_.p, d_sn_rowptr_.p, d_sn_rows_.p, d_sn_lptr_.p, d_upd_ptr_.p,
            d_upd_k_.p,   d_upd_p0_.p,  d_upd_p1_.p,  d_upd_relptr_.p, d_rel_.p,    d_task_ptr_.p, d_task_sn_.p};
  if (S_.d == 3) solve_t<3>(S_, P, d_perm_.p, d_L_.p, d_y_.p, d_b, d_x, d_status_.p, s, lc);
  else solve_t<6>(S_, P, d_perm_.p, d_L_.p, d_y_.p, d_b, d_x, d_status_.p, s, lc);
}

}  // namespace g2o_b200
