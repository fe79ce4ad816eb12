// sparse_inverse.cuh - sparse inverse subset (Takahashi recursion) on the supernodal GPU factor.
//
// Replaces LinearSolverCSparse::solvePattern + MarginalCovarianceCholesky (solvers/csparse/linear_solver_csparse.h:190-225,
// core/marginal_covariance_cholesky.cpp:55-214): the reference evaluates
//     Z(r,c) = [r == c] / L(r,r)^2 - 1/L(r,r) * sum_{k > r, L(k,r) != 0} L(k,r) Z(k,c)
// entry by entry with a hash-map cache on the scalar factor.  Here the same recursion runs supernode by supernode from the
// root of the elimination tree down (every block row below a supernode's diagonal block is a column of an ANCESTOR, so all
// of Z_RR is final when the supernode is reached):
//     Y    = L21 L11^-1                      (spinv_prepare_kernel, all supernodes at once)
//     Z_RJ = - Z_RR Y                        (spinv_rows_kernel, one CTA per block row of a supernode)
//     Z_JJ = L11^-T L11^-1 - Y^T Z_RJ        (spinv_diag_kernel, one CTA per column of the diagonal part)
// One launch pair per depth level of the supernodal tree; cost = the factorisation's flops, once, for EVERY block on the
// pattern of L (all diagonal blocks, all blocks of edges) - instead of one factorisation + solve per requested scalar
// column.  Z has the geometry of L (spinv_lookup.h); Z_RR blocks are located by binary search in the ancestor's row list.
#pragma once
#include "spinv_lookup.h"

namespace g2o_b200 {

struct SpinvDev {
  CholDev P;
  const long long* sn_dinvptr;
  const int* col2sn;
};

constexpr int kSpinvThreads = 256;
constexpr int kSpinvQC = 8;   // Z_RR blocks staged per step of the row kernel

template <int D>
__global__ void __launch_bounds__(kSpinvThreads)
spinv_prepare_kernel(const __grid_constant__ SpinvDev V, const double* __restrict__ L, const double* __restrict__ Dinv,
                     double* __restrict__ Yt, double* __restrict__ Z) {
  const int J = blockIdx.x;
  const int M = V.P.sn_nrow[J] * D, N = V.P.sn_ncol[J] * D, B = M - N;
  const double* Lp = L + V.P.sn_lptr[J];
  const double* Di = Dinv + V.sn_dinvptr[J];   // Di[k + c*N] = L11^-1 (k,c), lower triangular
  double* Yp = Yt + V.P.sn_lptr[J];            // Yp[c + N*r] = Y(r,c): the row kernel walks it along c
  double* Zp = Z + V.P.sn_lptr[J];
  const long long total = (long long)B * N;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.y * blockDim.x) {
    const int c = (int)(i / B), r = (int)(i - (long long)c * B);
    double s = 0.0;
    for (int k = c; k < N; ++k) s = fma(Lp[N + r + (long long)k * M], Di[k + (long long)c * N], s);
    Yp[c + (long long)N * r] = s;
  }
  if (blockIdx.y == 0)
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) {   // seed of the diagonal part: L11^-T L11^-1
      const int b = i / N, a = i - b * N;
      double s = 0.0;
      for (int k = max(a, b); k < N; ++k) s = fma(Di[k + (long long)a * N], Di[k + (long long)b * N], s);
      Zp[a + (long long)b * M] = s;
    }
}

// one CTA = one block row p (below the diagonal block) of one supernode: Z_RJ[p, :] = - sum_q G[p,q] Y[q, :]
template <int D>
__global__ void __launch_bounds__(kSpinvThreads)
spinv_rows_kernel(const __grid_constant__ SpinvDev V, const int* __restrict__ item_sn, const int* __restrict__ item_p,
                  int item0, const double* __restrict__ Yt, double* __restrict__ Z) {
  __shared__ double Gs[kSpinvQC * D * D];
  const int it = item0 + blockIdx.x;
  const int J = item_sn[it], p = item_p[it];
  const int nr = V.P.sn_nrow[J], nc = V.P.sn_ncol[J];
  const int M = nr * D, N = nc * D;
  const int* jrows = V.P.sn_rows + V.P.sn_rowptr[J];
  const double* Yp = Yt + V.P.sn_lptr[J];
  double* Zp = Z + V.P.sn_lptr[J];
  const int gp = jrows[p];
  constexpr int kOut = (D * 72 + kSpinvThreads - 1) / kSpinvThreads;   // outputs per thread (N <= 72)
  double acc[kOut];
  int oi[kOut], oc[kOut];
#pragma unroll
  for (int o = 0; o < kOut; ++o) {
    const int idx = threadIdx.x + o * kSpinvThreads;
    acc[o] = 0.0;
    oc[o] = idx % N;
    oi[o] = idx < D * N ? idx / N : -1;
  }
  for (int q0 = nc; q0 < nr; q0 += kSpinvQC) {
    const int nq = min(kSpinvQC, nr - q0);
    __syncthreads();
    for (int e = threadIdx.x; e < nq * D * D; e += kSpinvThreads) {
      const int qq = e / (D * D), ij = e - qq * D * D, j = ij / D, i = ij - j * D;   // G[p,q](i,j)
      const int gq = jrows[q0 + qq];
      int ld;
      double v;
      if (gp >= gq) {
        const long long off = spinv_locate(gp, gq, D, V.col2sn, V.P.sn_col0, V.P.sn_ncol, V.P.sn_nrow, V.P.sn_rowptr, V.P.sn_rows, V.P.sn_lptr, &ld);
        v = Z[off + i + (long long)j * ld];
      } else {
        const long long off = spinv_locate(gq, gp, D, V.col2sn, V.P.sn_col0, V.P.sn_ncol, V.P.sn_nrow, V.P.sn_rowptr, V.P.sn_rows, V.P.sn_lptr, &ld);
        v = Z[off + j + (long long)i * ld];
      }
      Gs[qq * D * D + i * D + j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < kOut; ++o) {
      if (oi[o] < 0) continue;
      double s = acc[o];
      for (int qq = 0; qq < nq; ++qq) {
        const double* g = Gs + qq * D * D + oi[o] * D;
        const double* y = Yp + oc[o] + (long long)N * ((q0 + qq - nc) * D);
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(g[j], y[(long long)N * j], s);
      }
      acc[o] = s;
    }
  }
#pragma unroll
  for (int o = 0; o < kOut; ++o)
    if (oi[o] >= 0) Zp[p * D + oi[o] + (long long)oc[o] * M] = -acc[o];
}

// one CTA = one COLUMN b of the diagonal part of one supernode: Z_JJ(:, b) -= Y^T Z_RJ(:, b) (the seed L11^-T L11^-1 is
// already there).  Thread = (row a, one of G interleaved slices of the B rows below), four independent partial sums per
// thread, slices added in fixed order through shared memory.  (First version: one CTA per supernode, every thread a
// serial dot product of length B - 0.9 ms for ONE wide supernode near the root, most of the 14 ms sweep on sphere2500.)
template <int D>
__global__ void __launch_bounds__(kSpinvThreads)
spinv_diag_kernel(const __grid_constant__ SpinvDev V, const int* __restrict__ level_sn, int sn0,
                  const double* __restrict__ Yt, double* __restrict__ Z) {
  __shared__ double part[kSpinvThreads];
  const int J = level_sn[sn0 + blockIdx.x];
  const int M = V.P.sn_nrow[J] * D, N = V.P.sn_ncol[J] * D, B = M - N;
  const int b = blockIdx.y;
  if (b >= N) return;   // whole CTA
  const double* Yp = Yt + V.P.sn_lptr[J];
  double* Zp = Z + V.P.sn_lptr[J];
  const int G = kSpinvThreads / N;            // slices (N <= 72: G >= 3)
  const int a = threadIdx.x % N, g = threadIdx.x / N;
  const double* zc = Zp + N + (long long)b * M;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (g < G) {
    int r = g;
    for (; r + 3 * G < B; r += 4 * G) {
      s0 = fma(Yp[a + (long long)N * r], zc[r], s0);
      s1 = fma(Yp[a + (long long)N * (r + G)], zc[r + G], s1);
      s2 = fma(Yp[a + (long long)N * (r + 2 * G)], zc[r + 2 * G], s2);
      s3 = fma(Yp[a + (long long)N * (r + 3 * G)], zc[r + 3 * G], s3);
    }
    for (; r < B; r += G) s0 = fma(Yp[a + (long long)N * r], zc[r], s0);
  }
  part[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0.0;
    for (int q = 0; q < G; ++q) s += part[threadIdx.x + q * N];
    Zp[threadIdx.x + (long long)b * M] -= s;
  }
}

// requested blocks -> dense output (d*d doubles each, column-major); trans: the stored block is the transposed one
template <int D>
__global__ void spinv_gather_kernel(int n, const long long* __restrict__ off, const int* __restrict__ ld,
                                    const unsigned char* __restrict__ trans, const double* __restrict__ Z,
                                    double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * D * D) return;
  const int q = idx / (D * D), ij = idx - q * D * D, j = ij / D, i = ij - j * D;
  out[idx] = trans[q] ? Z[off[q] + j + (long long)i * ld[q]] : Z[off[q] + i + (long long)j * ld[q]];
}

}  // namespace g2o_b200
