// pcg.cuh - LinearSolverPCG (solvers/pcg/linear_solver_pcg.hpp:79-197, linear_solver_pcg.h:47-98) on the GPU:
// conjugate gradients with the block-Jacobi preconditioner J = blockdiag(A)^-1 on the upper-triangular block CCS
// matrix the BlockSolver hands to its linear solver.  Same recurrences, same stopping rule (dn <= tolerance * dn_0, or
// the absolute residual carried over from the previous solve), same iteration limit (n when maxIter < 0).
//
// Two kernels per iteration, no host round trip inside a batch of iterations:
//   pcg_spmv_kernel    d <- s + beta d (fused: the direction of this iteration is formed on the fly for the gathered
//                      columns and written once into the other d buffer), q = A d, partial sums of d.q;
//                      the last CTA to arrive adds the partials in index order: alpha = dn / d.q
//   pcg_update_kernel  x += alpha d, r -= alpha q, s = J r, partial sums of r.s; the last CTA: dn, beta, iteration
//                      count, and the `done` flag the next pcg_spmv_kernel reads first
// A x is a GATHER over a symmetric block-row list built once per pattern on the host (every upper block appears in its
// row as B and in its column's row as B^T), so every output row is owned by one thread and summed in a fixed order; the
// reference scatters (dest[row] += B x[col]; dest[col] += B^T x[row], linear_solver_pcg.hpp:181-196): same sums, other
// order - results agree to rounding, the iteration count to +-1 near the threshold.
#pragma once

namespace g2o_b200 {

struct PcgScalars {
  double dn, dq, alpha, beta, d0;
  int iteration, done, max_iter, pad;
  unsigned arrive[2];
};

struct PcgDev {
  int nb;
  const int *rowptr, *ent_blk, *ent_col;        // symmetric block-row lists; ent_blk < 0: transposed block ~ent_blk
  const int* diag_blk;                          // per block column: its diagonal block
  const double* A;                              // blocks, d*d column-major
  double *J, *x, *r, *s, *q, *d0buf, *d1buf, *partial;
  PcgScalars* sc;
};

constexpr int kPcgThreads = 192;  // multiple of 3 and 6: a block row never straddles two CTAs

template <int D>
__global__ void pcg_jacobi_kernel(int nb, const int* __restrict__ diag_blk, const double* __restrict__ A, double* __restrict__ J,
                                  int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  // inverse of the symmetric positive definite diagonal block through its Cholesky factor (in registers)
  double L[D][D], Z[D][D];
  const double* B = A + (long long)diag_blk[i] * D * D;
#pragma unroll
  for (int c = 0; c < D; ++c)
#pragma unroll
    for (int r = 0; r < D; ++r) L[r][c] = r >= c ? B[c + r * D] : 0.0;  // upper triangle of the stored block, mirrored
  bool bad = false;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double s = L[k][k];
    if (!(s > 0.0)) { bad = true; s = 1.0; }
    const double rs = 1.0 / sqrt(s);
    L[k][k] = s * rs;
#pragma unroll
    for (int r = k + 1; r < D; ++r) L[r][k] *= rs;
#pragma unroll
    for (int c = k + 1; c < D; ++c)
#pragma unroll
      for (int r = c; r < D; ++r) L[r][c] -= L[r][k] * L[c][k];
  }
  // Z = L^-1 (lower), J = Z^T Z
#pragma unroll
  for (int c = 0; c < D; ++c)
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if (r < c) { Z[r][c] = 0.0; continue; }
      double s = r == c ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k >= c && k < r) s -= L[r][k] * Z[k][c];
      Z[r][c] = s / L[r][r];
    }
  double* o = J + (long long)i * D * D;
#pragma unroll
  for (int c = 0; c < D; ++c)
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k >= r && k >= c) s += Z[k][r] * Z[k][c];
      o[r + c * D] = s;
    }
  if (bad) *status = 1;
}

__device__ __forceinline__ double pcg_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
  return s;  // valid in thread 0
}
// thread 0 of every CTA: leave the CTA's partial sum; returns true in the CTA that arrives last (its thread 0 then adds
// all partials in index order: a fixed summation order whatever the arrival order)
__device__ __forceinline__ bool pcg_arrive(double part, double* partial, unsigned* counter, double* total) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = part;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (s_last) {
      __threadfence();
      double s = 0.0;
      for (unsigned k = 0; k < gridDim.x; ++k) s += __ldcg(partial + k);
      *total = s;
      *counter = 0u;
    }
  }
  __syncthreads();
  return s_last != 0;
}

// r = b, s = J r, x = 0, dn = r.s; the last CTA sets the threshold: linear_solver_pcg.hpp:117-130
template <int D>
__global__ void __launch_bounds__(kPcgThreads) pcg_init_kernel(PcgDev P, const double* __restrict__ b, double tolerance,
                                                               int absolute, double prev_residual, int max_iter) {
  __shared__ double sh[kPcgThreads / 32];
  const int t = blockIdx.x * kPcgThreads + threadIdx.x;
  const int n = P.nb * D;
  double part = 0.0;
  if (t < n) {
    const int i = t / D, c = t - i * D;
    const double* Ji = P.J + (long long)i * D * D;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += Ji[c + k * D] * b[i * D + k];
    P.r[t] = b[t]; P.s[t] = s; P.x[t] = 0.0; P.q[t] = 0.0; P.d0buf[t] = 0.0; P.d1buf[t] = 0.0;
    part = b[t] * s;
  }
  part = pcg_block_sum(part, sh);
  double total = 0.0;
  if (pcg_arrive(part, P.partial, &P.sc->arrive[0], &total) && threadIdx.x == 0) {
    PcgScalars* sc = P.sc;
    sc->dn = total;
    double d0 = tolerance * total;
    if (absolute && prev_residual > 0.0 && prev_residual > d0) d0 = prev_residual;
    sc->d0 = d0; sc->beta = 0.0; sc->alpha = 0.0; sc->dq = 0.0;
    sc->iteration = 0; sc->max_iter = max_iter;
    sc->done = (total <= d0 || max_iter <= 0) ? 1 : 0;
  }
}

template <int D>
__global__ void __launch_bounds__(kPcgThreads) pcg_spmv_kernel(PcgDev P) {
  __shared__ double sh[kPcgThreads / 32];
  const PcgScalars* scr = P.sc;
  if (scr->done) return;
  const int par = scr->iteration & 1;
  const double beta = scr->beta;
  const double* dold = par ? P.d1buf : P.d0buf;
  double* dnew = par ? P.d0buf : P.d1buf;
  const int t = blockIdx.x * kPcgThreads + threadIdx.x;
  const int n = P.nb * D;
  double part = 0.0;
  if (t < n) {
    const int i = t / D, c = t - i * D;
    double acc = 0.0;
    for (int e = P.rowptr[i]; e < P.rowptr[i + 1]; ++e) {
      const int kb = P.ent_blk[e], j = P.ent_col[e];
      const double* B = P.A + (long long)(kb < 0 ? ~kb : kb) * D * D;
      const double* sj = P.s + (long long)j * D;
      const double* dj = dold + (long long)j * D;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double dk = fma(beta, dj[k], sj[k]);          // the new direction at column block j, entry k
        acc = fma(kb < 0 ? B[k + c * D] : B[c + k * D], dk, acc);
      }
    }
    const double dt = fma(beta, dold[t], P.s[t]);
    dnew[t] = dt;
    P.q[t] = acc;
    part = dt * acc;
  }
  part = pcg_block_sum(part, sh);
  double total = 0.0;
  if (pcg_arrive(part, P.partial, &P.sc->arrive[0], &total) && threadIdx.x == 0) {
    P.sc->dq = total;
    P.sc->alpha = P.sc->dn / total;
  }
}

template <int D>
__global__ void __launch_bounds__(kPcgThreads) pcg_update_kernel(PcgDev P) {
  __shared__ double sh[kPcgThreads / 32];
  __shared__ double s_r[kPcgThreads];
  const PcgScalars* scr = P.sc;
  if (scr->done) return;
  const int par = scr->iteration & 1;
  const double alpha = scr->alpha;
  const double* dcur = par ? P.d0buf : P.d1buf;   // the buffer pcg_spmv_kernel has just written
  const int t = blockIdx.x * kPcgThreads + threadIdx.x;
  const int n = P.nb * D;
  double rn = 0.0;
  if (t < n) {
    P.x[t] = fma(alpha, dcur[t], P.x[t]);
    rn = fma(-alpha, P.q[t], P.r[t]);
    P.r[t] = rn;
  }
  s_r[threadIdx.x] = rn;
  __syncthreads();
  double part = 0.0;
  if (t < n) {
    const int i = t / D, c = t - i * D;
    const double* Ji = P.J + (long long)i * D * D;
    const double* rb = s_r + (threadIdx.x - c);   // kPcgThreads is a multiple of D: the block's D entries sit in this CTA
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s = fma(Ji[c + k * D], rb[k], s);
    P.s[t] = s;
    part = rn * s;
  }
  part = pcg_block_sum(part, sh);
  double total = 0.0;
  if (pcg_arrive(part, P.partial, &P.sc->arrive[1], &total) && threadIdx.x == 0) {
    PcgScalars* sc = P.sc;
    const double dold = sc->dn;
    sc->dn = total;
    sc->beta = total / dold;
    sc->iteration += 1;
    if (total <= sc->d0 || sc->iteration >= sc->max_iter) sc->done = 1;
  }
}

}  // namespace g2o_b200
