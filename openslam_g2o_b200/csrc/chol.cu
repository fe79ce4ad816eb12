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
// numeric factorisation
//   update : one CTA per destination tile (48 x 48 scalars) pulls every update piece that lands in the tile,
//            accumulates them in shared memory in a fixed order and subtracts the sum from the panel once
//   factor : one CTA per (supernode, row chunk): diagonal block + chunk rows staged in shared memory, blocked
//            right-looking Cholesky with one panel row per thread (pivot block factored redundantly in registers)
// Small subtrees run both phases for all their supernodes inside one CTA (fused kernel); the top of the tree is
// level-scheduled with the two phases as separate multi-CTA kernels.
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 48;        // scalar rows / cols of a destination tile
constexpr int kMaxPanelCols = 96;
constexpr int kCholThreads = 512;
constexpr int kUpdateSmemDoubles = kTile * kTile + 2 * kTile * kMaxPanelCols;  // acc | A rows | B rows

struct CholPlanDev {
  const int *tile_sn, *tile_r0, *tile_c0, *tile_work_ptr;
  const int *work_u, *work_a0, *work_a1, *work_b0, *work_b1;
  const int *sn_tile_ptr, *sn_chunk_ptr, *chunk_sn, *chunk_b0, *chunk_nb;
  const long long *sn_dinvptr, *sn_cptr;
  const long long *work_koff, *work_reloff;
  const int *work_mk, *work_nk, *fwd_ptr, *fwd_src;
};

// acc (shared, kTile x kTile) = sum over work items [w0,w1) of the tile, in list order.
// Operand rows of the updating panel are staged in shared memory with coalesced loads; every thread then owns
// 3x3 micro tiles of the product.
template <int D>
__device__ void accumulate_items(const CholDev& P, const CholPlanDev& Q, const double* __restrict__ L, int w0, int w1,
                                 int R0, int C0, double* __restrict__ acc, double* __restrict__ As,
                                 double* __restrict__ Bs) {
  constexpr int S = D / 3;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < kTile * kTile; i += nt) acc[i] = 0.0;
  for (int wi = w0; wi < w1; ++wi) {
    // one level of indirection: everything the item needs sits in flat per-item arrays
    const int a0 = Q.work_a0[wi], a1 = Q.work_a1[wi], b0 = Q.work_b0[wi], b1 = Q.work_b1[wi];
    const int Mk = Q.work_mk[wi], Nk = Q.work_nk[wi];
    const double* Kp = L + Q.work_koff[wi];
    const int* rel = P.rel + Q.work_reloff[wi];
    const int nA = (a1 - a0) * D, nB = (b1 - b0) * D;
    __syncthreads();  // previous item's operands fully consumed, acc zeroing done
    {
      // stage both operand row blocks; 4 independent global loads in flight per thread
      const int totA = nA * Nk, tot = totA + nB * Nk;
      for (int i0 = tid; i0 < tot; i0 += 4 * nt) {
        double v[4];
        int dst[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * nt;
          dst[q] = -1;
          if (i < tot) {
            const bool isA = i < totA;
            const int ii = isA ? i : i - totA;
            const int nR = isA ? nA : nB;
            const int k = ii / nR, r = ii - k * nR;
            v[q] = Kp[(isA ? a0 : b0) * D + r + (long long)k * Mk];
            dst[q] = (isA ? 0 : kTile * kMaxPanelCols) + r + k * kTile;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (dst[q] >= 0) As[dst[q]] = v[q];  // Bs == As + kTile*kMaxPanelCols
      }
    }
    __syncthreads();
    const int na = (a1 - a0) * S, nb = (b1 - b0) * S;
    for (int idx = tid; idx < na * nb; idx += nt) {
      const int j = idx / na, i = idx - j * na;
      const int ma = a0 * S + i, mb = b0 * S + j;
      if (ma < mb) continue;
      const double* ra = As + i * 3;
      const double* rb = Bs + j * 3;
      double c00 = 0, c01 = 0, c02 = 0, c10 = 0, c11 = 0, c12 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll 4
      for (int k = 0; k < Nk; ++k) {
        const double x0 = ra[0], x1 = ra[1], x2 = ra[2];
        const double y0 = rb[0], y1 = rb[1], y2 = rb[2];
        c00 = fma(x0, y0, c00); c01 = fma(x0, y1, c01); c02 = fma(x0, y2, c02);
        c10 = fma(x1, y0, c10); c11 = fma(x1, y1, c11); c12 = fma(x1, y2, c12);
        c20 = fma(x2, y0, c20); c21 = fma(x2, y1, c21); c22 = fma(x2, y2, c22);
        ra += kTile;
        rb += kTile;
      }
      const int ab = ma / S, bb = mb / S;
      const int tr = (rel[ab] - R0) * D + (ma - ab * S) * 3;
      const int tc = (rel[bb] - C0) * D + (mb - bb * S) * 3;
      double* dst = acc + tr + tc * kTile;
      dst[0] += c00; dst[1] += c10; dst[2] += c20;
      dst[kTile] += c01; dst[kTile + 1] += c11; dst[kTile + 2] += c21;
      dst[2 * kTile] += c02; dst[2 * kTile + 1] += c12; dst[2 * kTile + 2] += c22;
    }
  }
  __syncthreads();
}

template <int D>
__device__ void subtract_tile(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, int tile,
                              const double* __restrict__ acc) {
  const int J = Q.tile_sn[tile], R0 = Q.tile_r0[tile], C0 = Q.tile_c0[tile];
  const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
  double* Pj = L + P.sn_lptr[J];
  const int rows = min(kTile, M - R0 * D), cols = min(kTile, N - C0 * D);
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int c = i / rows, r = i - c * rows;
    Pj[(long long)(R0 * D + r) + (long long)(C0 * D + c) * M] -= acc[r + c * kTile];
  }
}

#ifdef CHOL_TIMING
__device__ unsigned long long g_chol_timing[8];
#define TCK(i) do { if (threadIdx.x == 0) { unsigned long long _n = clock64(); atomicAdd(&g_chol_timing[i], _n - _t0); _t0 = _n; } } while (0)
#else
#define TCK(i) do {} while (0)
#endif

// 1/sqrt(s) to full double accuracy: single-precision seed (one MUFU) + two Newton steps in double; falls back to the
// library routine outside the float range.  Sits on the critical path of every block column.
__device__ __forceinline__ double fast_rsqrt(double s) {
  if (s < 1e-30 || s > 1e30) return rsqrt(s);
  double r = (double)rsqrtf((float)s);
  const double hs = -0.5 * s;
  r = r * fma(hs * r, r, 1.5);
  r = r * fma(hs * r, r, 1.5);
  return r;
}

template <int D>
__device__ void factor_chunk(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, double* __restrict__ Ldiag,
                             int chunk, bool write_diag, double* __restrict__ Sm, int* status,
                             const double* __restrict__ y, double* __restrict__ z, double* contrib) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int J = Q.chunk_sn[chunk];
  const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
  const int crow0 = Q.chunk_b0[chunk] * D, crows = Q.chunk_nb[chunk] * D;
  // rows staged: the diagonal block, this chunk's rows and - when the forward solve rides along - the right-hand
  // side of the supernode as one more row: its "triangular solve" IS the forward substitution y_J = L11^-1 t
  const bool rhs = y != nullptr;
  const int Rp = N + crows;
  const int R = Rp + (rhs ? 1 : 0);
  const int col0s = P.sn_col0[J] * D;
  double* Pj = L + P.sn_lptr[J];
#ifdef CHOL_TIMING
  unsigned long long _t0 = clock64();
#endif
  __syncthreads();
  {
    // warps split in two groups: one streams the panel in, the other gathers the right-hand side
    // t_c = (P b)_c - sum of the descendants' contributions (lanes stride the list, fixed shuffle tree)
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int ngather = rhs ? max(1, nw / 4) : 0;
    if (wid < ngather) {
      for (int c = wid; c < N; c += ngather) {
        const int g = col0s + c;
        const int e0 = Q.fwd_ptr[g], e1 = Q.fwd_ptr[g + 1];
        double part = 0.0;
        for (int e = e0 + lane; e < e1; e += 32) part += contrib[Q.fwd_src[e]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) Sm[Rp + c * R] = y[g] - part;
      }
    } else {
      const int t2 = tid - ngather * 32, n2 = nt - ngather * 32;
      for (int i = t2; i < Rp * N; i += n2) {
        const int c = i / Rp, r = i - c * Rp;
        const int gr = r < N ? r : crow0 + (r - N);
        Sm[r + c * R] = Pj[gr + (long long)c * M];
      }
    }
  }
  __syncthreads();
  TCK(0);
  const int row = tid;  // one panel row per thread (R <= 192 <= blockDim)
  const int ncb = N / D;
  bool bad = false;
  for (int jb = 0; jb < ncb; ++jb) {
    const int j0 = jb * D;
    // pivot block (already carries every earlier update): factor redundantly in registers
    double Lp[D][D], inv[D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) Lp[r][c] = Sm[(j0 + r) + (j0 + c) * R];
    // right-looking inside the block: as soon as column k is scaled the remaining entries are updated with
    // independent FMAs, so the dependent chain per column is rsqrt -> multiply -> one FMA
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double s = Lp[k][k];
      if (!(s > 0.0)) { bad = true; s = 1.0; }  // d <= 0: not positive definite (csparse_helper.cpp:136)
      const double rs = fast_rsqrt(s);
      inv[k] = rs;
      Lp[k][k] = s * rs;
#pragma unroll
      for (int r = k + 1; r < D; ++r) Lp[r][k] *= rs;
#pragma unroll
      for (int c = k + 1; c < D; ++c)
#pragma unroll
        for (int r = c; r < D; ++r) Lp[r][c] = fma(-Lp[r][k], Lp[c][k], Lp[r][c]);
    }
    TCK(1);
    double x[D];
    const bool below = row >= j0 + D && row < R;
    if (row >= j0 && row < j0 + D) {
      const int rr = row - j0;
#pragma unroll
      for (int r = 0; r < D; ++r)
        if (r == rr) {
#pragma unroll
          for (int c = 0; c <= r; ++c) Sm[row + (j0 + c) * R] = Lp[r][c];
        }
    } else if (below) {
#pragma unroll
      for (int c = 0; c < D; ++c) x[c] = Sm[row + (j0 + c) * R];
#pragma unroll
      for (int c = 0; c < D; ++c) {  // right-looking substitution: chain per column = multiply -> one FMA
        x[c] *= inv[c];
#pragma unroll
        for (int m = c + 1; m < D; ++m) x[m] = fma(-x[c], Lp[m][c], x[m]);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) Sm[row + (j0 + c) * R] = x[c];
    }
    TCK(2);
    __syncthreads();
    TCK(3);
    {
      // rank-D trailing update T(i,c) -= sum_k X(i,k) X(c,k), i >= j0+D, c in [j0+D, N): every thread takes
      // 2 rows x 4 columns per item (rows rp apart so that a warp walks consecutive rows); entries above the
      // diagonal are never read again, so they may be updated or skipped freely
      const int base = j0 + D;
      const int nrows = R - base, ncols = N - base;
      const int rp = (nrows + 1) >> 1, cq = (ncols + 3) >> 2;
      for (int item = tid; item < rp * cq; item += nt) {
        const int ci = item / rp, ri = item - ci * rp;
        const int r0 = base + ri, r1 = r0 + rp;
        const int c0 = base + 4 * ci;
        const bool has1 = r1 < R;
        if ((has1 ? r1 : r0) < N && c0 > (has1 ? r1 : r0)) continue;  // whole item above the diagonal
        // all loads first, then 8 independent FMA chains, then the stores: nothing in between can alias
        double x0[D], x1[D], l[4][D], v0[4], v1[4];
        const int rr1 = has1 ? r1 : r0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          x0[k] = Sm[r0 + (j0 + k) * R];
          x1[k] = Sm[rr1 + (j0 + k) * R];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = min(c0 + q, N - 1);
          v0[q] = Sm[r0 + c * R];
          v1[q] = Sm[rr1 + c * R];
#pragma unroll
          for (int k = 0; k < D; ++k) l[q][k] = Sm[c + (j0 + k) * R];
        }
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v0[q] = fma(-x0[k], l[q][k], v0[q]);
            v1[q] = fma(-x1[k], l[q][k], v1[q]);
          }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q < N) {
            Sm[r0 + (c0 + q) * R] = v0[q];
            if (has1) Sm[r1 + (c0 + q) * R] = v1[q];
          }
      }
    }
    TCK(4);
    __syncthreads();
    TCK(5);
  }
  if (bad && tid == 0) *status = 1;
  for (int i = tid; i < R * N; i += nt) {
    const int c = i / R, r = i - c * R;
    if (r < N) {
      // the factored diagonal block goes to its own array: sibling chunk CTAs are still reading the unfactored
      // block from the panel (every chunk factors it redundantly), so it must not be overwritten in place
      if (write_diag && r >= c) Ldiag[Q.sn_dinvptr[J] + r + (long long)c * N] = Sm[i];
    } else if (r < Rp) {
      Pj[crow0 + (r - N) + (long long)c * M] = Sm[i];
    } else if (write_diag) {
      z[col0s + c] = Sm[i];  // y_J goes to its own vector: sibling chunks still read (P b)_J from y
    }
  }
  if (rhs) {
    // c_J = L21 y_J for this chunk's rows: what every ancestor will subtract from its right-hand side
    double* cj = contrib + Q.sn_cptr[J] + (crow0 - N);
    for (int r = tid; r < crows; r += nt) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int c = 0;
      for (; c + 3 < N; c += 4) {
        s0 = fma(Sm[N + r + c * R], Sm[Rp + c * R], s0);
        s1 = fma(Sm[N + r + (c + 1) * R], Sm[Rp + (c + 1) * R], s1);
        s2 = fma(Sm[N + r + (c + 2) * R], Sm[Rp + (c + 2) * R], s2);
        s3 = fma(Sm[N + r + (c + 3) * R], Sm[Rp + (c + 3) * R], s3);
      }
      for (; c < N; ++c) s0 = fma(Sm[N + r + c * R], Sm[Rp + c * R], s0);
      cj[r] = (s0 + s1) + (s2 + s3);
    }
  }
  TCK(6);
}

template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_fused_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, double* __restrict__ Ldiag, int task0, int* status,
                  const double* __restrict__ y, double* __restrict__ z, double* contrib) {
  extern __shared__ __align__(16) double smem[];  // update operands and factor staging share the same bytes
  double* acc = smem;
  double* As = smem + kTile * kTile;
  double* Bs = As + kTile * kMaxPanelCols;
  const int t = task0 + blockIdx.x;
  for (int q = P.task_ptr[t]; q < P.task_ptr[t + 1]; ++q) {
    const int J = P.task_sn[q];
    for (int tile = Q.sn_tile_ptr[J]; tile < Q.sn_tile_ptr[J + 1]; ++tile) {
      const int w0 = Q.tile_work_ptr[tile], w1 = Q.tile_work_ptr[tile + 1];
      if (w0 == w1) continue;
      accumulate_items<D>(P, Q, L, w0, w1, Q.tile_r0[tile], Q.tile_c0[tile], acc, As, Bs);
      subtract_tile<D>(P, Q, L, tile, acc);
      __syncthreads();
    }
    const int c0 = Q.sn_chunk_ptr[J], c1 = Q.sn_chunk_ptr[J + 1];
    for (int ch = c0; ch < c1; ++ch) factor_chunk<D>(P, Q, L, Ldiag, ch, ch == c0, smem, status, y, z, contrib);
    __syncthreads();
  }
}

// split levels, phase 1: one CTA per (tile, split-K group)
template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_update_groups_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, const int* __restrict__ g_tile,
                          const int* __restrict__ g_w0, const int* __restrict__ g_w1, const int* __restrict__ g_slot,
                          double* __restrict__ scratch) {
  extern __shared__ __align__(16) double smem[];
  double* acc = smem;
  double* As = smem + kTile * kTile;
  double* Bs = As + kTile * kMaxPanelCols;
  const int g = blockIdx.x;
  const int tile = g_tile[g];
  accumulate_items<D>(P, Q, L, g_w0[g], g_w1[g], Q.tile_r0[tile], Q.tile_c0[tile], acc, As, Bs);
  const int slot = g_slot[g];
  if (slot < 0) {
    subtract_tile<D>(P, Q, L, tile, acc);
  } else {
    double* out = scratch + (long long)slot * kTile * kTile;
    for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) out[i] = acc[i];
  }
}
// split levels, phase 1b: tiles cut into several groups: add the partial sums in group order, subtract once
template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_reduce_tiles_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, const int* __restrict__ r_tile,
                         const int* __restrict__ r_slot0, const int* __restrict__ r_nslots,
                         const double* __restrict__ scratch) {
  __shared__ __align__(16) double acc[kTile * kTile];
  const int t = blockIdx.x;
  const double* in = scratch + (long long)r_slot0[t] * kTile * kTile;
  const int ns = r_nslots[t];
  for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < ns; ++g) s += in[(long long)g * kTile * kTile + i];
    acc[i] = s;
  }
  __syncthreads();
  subtract_tile<D>(P, Q, L, r_tile[t], acc);
}
template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_factor_chunks_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, double* __restrict__ Ldiag,
                          const int* __restrict__ chunks, int* status, const double* __restrict__ y,
                          double* __restrict__ z, double* contrib) {
  extern __shared__ __align__(16) double smem[];
  const int ch = chunks[blockIdx.x];
  factor_chunk<D>(P, Q, L, Ldiag, ch, ch == Q.sn_chunk_ptr[Q.chunk_sn[ch]], smem, status, y, z, contrib);
}

// inverse of every triangular diagonal block (one CTA per supernode): the solves become matrix-vector products.
// Zt holds the inverse transposed (Zt[j + k*N] = inv(k,j)) so that thread j walks its column conflict-free.
template <int D>
__global__ void __launch_bounds__(128)
chol_invert_diag_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ Ldiag, double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double sm[];
  const int J = blockIdx.x;
  const int N = P.sn_ncol[J] * D;
  double* Ls = sm;           // N*N, Ls[i + k*N]
  double* Zt = sm + N * N;   // N*N
  const double* Lj = Ldiag + Q.sn_dinvptr[J];
  double* out = Dinv + Q.sn_dinvptr[J];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) Ls[i] = Lj[i];
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    for (int i = j; i < N; ++i) {
      double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0;
      int k = j;
      for (; k + 1 < i; k += 2) {
        s0 = fma(-Ls[i + k * N], Zt[j + k * N], s0);
        s1 = fma(-Ls[i + (k + 1) * N], Zt[j + (k + 1) * N], s1);
      }
      if (k < i) s0 = fma(-Ls[i + k * N], Zt[j + k * N], s0);
      Zt[j + i * N] = (s0 + s1) / Ls[i + i * N];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int c = i / N, r = i - c * N;  // out(r,c) = inv(r,c) = Zt[c + r*N]
    out[i] = r >= c ? Zt[c + r * N] : 0.0;
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

// forward: y_J = Linv (P b - contributions of the descendants); every supernode then leaves c_J = L21 y_J in its own
// scratch segment, so ancestors only add precomputed numbers (per-row lists in a fixed order -> deterministic) and L
// is read once, coalesced, by its owner.
constexpr int kSolveThreads = 256;

template <int D>
__global__ void __launch_bounds__(kSolveThreads)
chol_forward_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ L, const double* __restrict__ Dinv,
                    double* __restrict__ y, double* __restrict__ contrib, int task0) {
  __shared__ double tvec[kMaxPanelCols];
  __shared__ double yv[kMaxPanelCols];
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = P.task_ptr[t]; q < P.task_ptr[t + 1]; ++q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
    double* yj = y + (long long)col0 * D;
    __syncthreads();  // contributions written by the previous supernode of this task are visible
    for (int i = tid; i < N; i += nt) {
      const int g = col0 * D + i;
      double s = yj[i];
      const int e0 = Q.fwd_ptr[g], e1 = Q.fwd_ptr[g + 1];
      for (int e = e0; e < e1; ++e) s -= contrib[Q.fwd_src[e]];
      tvec[i] = s;
    }
    __syncthreads();
    const double* Di = Dinv + Q.sn_dinvptr[J];
    for (int i = tid; i < N; i += nt) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = 0;
      for (; j + 3 <= i; j += 4) {
        s0 = fma(Di[i + (long long)j * N], tvec[j], s0);
        s1 = fma(Di[i + (long long)(j + 1) * N], tvec[j + 1], s1);
        s2 = fma(Di[i + (long long)(j + 2) * N], tvec[j + 2], s2);
        s3 = fma(Di[i + (long long)(j + 3) * N], tvec[j + 3], s3);
      }
      for (; j <= i; ++j) s0 = fma(Di[i + (long long)j * N], tvec[j], s0);
      const double v = (s0 + s1) + (s2 + s3);
      yj[i] = v;
      yv[i] = v;
    }
    __syncthreads();
    const double* Pj = L + P.sn_lptr[J];
    double* cj = contrib + Q.sn_cptr[J];
    for (int r = N + tid; r < M; r += nt) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = 0;
      for (; k + 3 < N; k += 4) {
        s0 = fma(Pj[r + (long long)k * M], yv[k], s0);
        s1 = fma(Pj[r + (long long)(k + 1) * M], yv[k + 1], s1);
        s2 = fma(Pj[r + (long long)(k + 2) * M], yv[k + 2], s2);
        s3 = fma(Pj[r + (long long)(k + 3) * M], yv[k + 3], s3);
      }
      for (; k < N; ++k) s0 = fma(Pj[r + (long long)k * M], yv[k], s0);
      cj[r - N] = (s0 + s1) + (s2 + s3);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kSolveThreads)
chol_backward_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ L, const double* __restrict__ Dinv,
                     double* __restrict__ y, int task0) {
  extern __shared__ __align__(16) double xb[];  // x at the rows below the diagonal block
  __shared__ double tvec[kMaxPanelCols];
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int q = P.task_ptr[t + 1] - 1; q >= P.task_ptr[t]; --q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int nc = P.sn_ncol[J], nr = P.sn_nrow[J];
    const int M = nr * D, N = nc * D;
    const double* Pj = L + P.sn_lptr[J];
    const int* jrows = P.sn_rows + P.sn_rowptr[J];
    double* xj = y + (long long)col0 * D;
    __syncthreads();
    for (int i = tid; i < M - N; i += nt) xb[i] = y[(long long)jrows[nc + i / D] * D + (i % D)];
    __syncthreads();
    // t = y_J - L21^T x_below : one warp per column, lanes stride the rows (coalesced), fixed-order shuffle tree
    for (int j = wid; j < N; j += nw) {
      const double* cj = Pj + (long long)j * M + N;
      double s0 = 0.0, s1 = 0.0;
      int i = lane;
      for (; i + 32 < M - N; i += 64) { s0 = fma(cj[i], xb[i], s0); s1 = fma(cj[i + 32], xb[i + 32], s1); }
      if (i < M - N) s0 = fma(cj[i], xb[i], s0);
      double s = s0 + s1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tvec[j] = xj[j] - s;
    }
    __syncthreads();
    const double* Di = Dinv + Q.sn_dinvptr[J];
    for (int i = tid; i < N; i += nt) {  // x_J = Linv^T t
      const double* ci = Di + (long long)i * N;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = i;
      for (; j + 3 < N; j += 4) {
        s0 = fma(ci[j], tvec[j], s0); s1 = fma(ci[j + 1], tvec[j + 1], s1);
        s2 = fma(ci[j + 2], tvec[j + 2], s2); s3 = fma(ci[j + 3], tvec[j + 3], s3);
      }
      for (; j < N; ++j) s0 = fma(ci[j], tvec[j], s0);
      xj[i] = (s0 + s1) + (s2 + s3);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
CholeskyGpu::~CholeskyGpu() {}

namespace {
template <typename T>
void up64(DevBuf<long long>& d, const std::vector<T>& v, cudaStream_t s, std::vector<std::vector<long long>>& keep) {
  keep.emplace_back(v.begin(), v.end());
  d.upload(keep.back(), s);
}
constexpr int kMaxDynSmem = 200 * 1024;
template <int D>
void set_smem_attrs() {
  static bool done = false;
  if (done) return;
  B200_CUDA(cudaFuncSetAttribute(chol_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_update_groups_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_factor_chunks_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_invert_diag_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_backward_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  done = true;
}
// profiler ids of the kernel groups inside the Cholesky (continue the numbering of solver.cu)
enum { PH_CH_SCATTER = 12, PH_CH_UPDATE = 13, PH_CH_REDUCE = 14, PH_CH_PANEL = 15, PH_CH_FUSED = 16, PH_CH_INVERT = 17,
       PH_CH_FORWARD = 18, PH_CH_BACKWARD = 19 };
}  // namespace

void CholeskyGpu::analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt,
                          cudaStream_t s) {
  S_ = g2o_b200::analyze(nb, d, colptr, rowidx, opt);
  nblk_ = colptr[nb];
  std::vector<std::vector<long long>> keep;
  d_sn_col0_.upload(S_.sn_col0, s); d_sn_ncol_.upload(S_.sn_ncol, s); d_sn_nrow_.upload(S_.sn_nrow, s);
  d_sn_rowptr_.upload(S_.sn_rowptr, s); d_sn_rows_.upload(S_.sn_rows, s);
  up64(d_sn_lptr_, S_.sn_lptr, s, keep);
  d_upd_ptr_.upload(S_.upd_ptr, s); d_upd_k_.upload(S_.upd_k, s); d_upd_p0_.upload(S_.upd_p0, s);
  d_upd_p1_.upload(S_.upd_p1, s); d_rel_.upload(S_.rel, s);
  up64(d_upd_relptr_, S_.upd_relptr, s, keep);
  d_task_ptr_.upload(S_.task_ptr, s); d_task_sn_.upload(S_.task_sn, s);
  up64(d_a_dst_, S_.a_dst, s, keep);
  up64(d_diag_dst_, S_.diag_dst, s, keep);
  d_a_ld_.upload(S_.a_ld, s); d_diag_ld_.upload(S_.diag_ld, s); d_perm_.upload(S_.perm, s);
  d_a_trans_.upload(S_.a_trans, s);
  d_tile_sn_.upload(S_.tile_sn, s); d_tile_r0_.upload(S_.tile_r0, s); d_tile_c0_.upload(S_.tile_c0, s);
  d_tile_work_ptr_.upload(S_.tile_work_ptr, s);
  d_work_u_.upload(S_.work_u, s); d_work_a0_.upload(S_.work_a0, s); d_work_a1_.upload(S_.work_a1, s);
  d_work_b0_.upload(S_.work_b0, s); d_work_b1_.upload(S_.work_b1, s);
  d_sn_tile_ptr_.upload(S_.sn_tile_ptr, s); d_sn_chunk_ptr_.upload(S_.sn_chunk_ptr, s);
  d_chunk_sn_.upload(S_.chunk_sn, s); d_chunk_b0_.upload(S_.chunk_b0, s); d_chunk_nb_.upload(S_.chunk_nb, s);
  d_level_chunks_.upload(S_.level_chunks, s);
  d_group_tile_.upload(S_.group_tile, s); d_group_w0_.upload(S_.group_w0, s); d_group_w1_.upload(S_.group_w1, s);
  d_group_slot_.upload(S_.group_slot, s);
  d_rtile_tile_.upload(S_.rtile_tile, s); d_rtile_slot0_.upload(S_.rtile_slot0, s); d_rtile_nslots_.upload(S_.rtile_nslots, s);
  up64(d_sn_dinvptr_, S_.sn_dinvptr, s, keep);
  up64(d_sn_cptr_, S_.sn_cptr, s, keep);
  up64(d_work_koff_, S_.work_koff, s, keep);
  up64(d_work_reloff_, S_.work_reloff, s, keep);
  d_work_mk_.upload(S_.work_mk, s); d_work_nk_.upload(S_.work_nk, s);
  d_fwd_ptr_.upload(S_.fwd_ptr, s); d_fwd_src_.upload(S_.fwd_src, s);
  d_L_.alloc((size_t)S_.factor_doubles);
  d_Dinv_.alloc((size_t)S_.dinv_doubles);
  d_Ldiag_.alloc((size_t)S_.dinv_doubles);
  d_gscratch_.alloc((size_t)std::max(S_.max_group_slots, 1) * kTile * kTile);
  d_contrib_.alloc((size_t)std::max<int64_t>(S_.sn_cptr[S_.nsn], 1));
  d_y_.alloc((size_t)nb * d);
  d_z_.alloc((size_t)nb * d);
  d_status_.alloc(1);
  if (!host_only_flag()) {
    B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
    if (d == 3) set_smem_attrs<3>(); else set_smem_attrs<6>();
  }
  analyzed_ = true;
}

CholDev CholeskyGpu::dev() const {
  return CholDev{d_sn_col0_.p, d_sn_ncol_.p, d_sn_nrow_.p, d_sn_rowptr_.p, d_sn_rows_.p, d_sn_lptr_.p, d_upd_ptr_.p,
                 d_upd_k_.p,   d_upd_p0_.p,  d_upd_p1_.p,  d_upd_relptr_.p, d_rel_.p,    d_task_ptr_.p, d_task_sn_.p};
}
CholPlanDev CholeskyGpu::plan() const {
  return CholPlanDev{d_tile_sn_.p, d_tile_r0_.p, d_tile_c0_.p, d_tile_work_ptr_.p, d_work_u_.p, d_work_a0_.p, d_work_a1_.p,
                     d_work_b0_.p, d_work_b1_.p, d_sn_tile_ptr_.p, d_sn_chunk_ptr_.p, d_chunk_sn_.p, d_chunk_b0_.p,
                     d_chunk_nb_.p, d_sn_dinvptr_.p, d_sn_cptr_.p, d_work_koff_.p, d_work_reloff_.p, d_work_mk_.p,
                     d_work_nk_.p, d_fwd_ptr_.p, d_fwd_src_.p};
}

template <int D>
void CholeskyGpu::factor_t(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
                           EventProfiler* prof) {
  const SymbolicFactor& S = S_;
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  double* L = d_L_.p;
  int* status = d_status_.p;
  auto count = [&](int n = 1) { if (lc) lc->n += n; };
  {
    ScopedPhase ph(prof, PH_CH_SCATTER);
    B200_CUDA(cudaMemsetAsync(L, 0, (size_t)S.factor_doubles * sizeof(double), s));
    B200_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
    chol_scatter_kernel<D><<<ceil_div((int64_t)nblk_ * D * D, 256), 256, 0, s>>>(dA, nblk_, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, L);
    count();
    if (d_lambda) {
      chol_add_lambda_kernel<D><<<ceil_div((int64_t)S.nb * D, 256), 256, 0, s>>>(S.nb, d_diag_dst_.p, d_diag_ld_.p, d_lambda, L);
      count();
    }
    if (d_b) {  // the forward substitution rides along with the factorisation (one more row per panel)
      chol_permute_in_kernel<D><<<ceil_div(S.nb * D, 256), 256, 0, s>>>(S.nb, d_perm_.p, d_b, d_y_.p);
      count();
    }
  }
  const double* yv = d_b ? d_y_.p : nullptr;
  double* zv = d_z_.p;
  double* contrib = d_contrib_.p;
  const size_t rhs_smem = d_b ? (size_t)S.max_ncol * D * sizeof(double) : 0;
  forward_done_ = d_b != nullptr;
  const size_t upd_smem = (size_t)kUpdateSmemDoubles * sizeof(double);
  for (int l = 0; l < S.nlevels; ++l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    if (S.level_kind[l] == 0) {
      ScopedPhase ph(prof, PH_CH_FUSED);
      const size_t smem = std::max(upd_smem, (size_t)S.level_smem[l] + rhs_smem);
      chol_fused_kernel<D><<<nt, kCholThreads, smem, s>>>(P, Q, L, d_Ldiag_.p, t0, status, yv, zv, contrib);
      count();
    } else {
      const int g0 = S.level_group_ptr[l], ng = S.level_group_ptr[l + 1] - g0;
      const int r0 = S.level_rtile_ptr[l], nr = S.level_rtile_ptr[l + 1] - r0;
      const int c0 = S.level_chunk_ptr[l], nchunk = S.level_chunk_ptr[l + 1] - c0;
      if (ng > 0) {
        ScopedPhase ph(prof, PH_CH_UPDATE);
        chol_update_groups_kernel<D><<<ng, kCholThreads, upd_smem, s>>>(P, Q, L, d_group_tile_.p + g0, d_group_w0_.p + g0,
                                                                       d_group_w1_.p + g0, d_group_slot_.p + g0, d_gscratch_.p);
        count();
      }
      if (nr > 0) {
        ScopedPhase ph(prof, PH_CH_REDUCE);
        chol_reduce_tiles_kernel<D><<<nr, kCholThreads, 0, s>>>(P, Q, L, d_rtile_tile_.p + r0, d_rtile_slot0_.p + r0,
                                                               d_rtile_nslots_.p + r0, d_gscratch_.p);
        count();
      }
      ScopedPhase ph(prof, PH_CH_PANEL);
      chol_factor_chunks_kernel<D><<<nchunk, kCholThreads, S.level_smem[l] + rhs_smem, s>>>(P, Q, L, d_Ldiag_.p, d_level_chunks_.p + c0, status, yv, zv, contrib);
      count();
    }
  }
  {
    ScopedPhase ph(prof, PH_CH_INVERT);
    const size_t ismem = 2 * (size_t)S.max_ncol * D * S.max_ncol * D * 8;
    chol_invert_diag_kernel<D><<<S.nsn, 128, ismem, s>>>(P, Q, d_Ldiag_.p, d_Dinv_.p);
    count();
  }
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::factor(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
                         EventProfiler* prof) {
  if (S_.d == 3) factor_t<3>(dA, d_lambda, d_b, s, lc, prof);
  else factor_t<6>(dA, d_lambda, d_b, s, lc, prof);
}

template <int D>
void CholeskyGpu::solve_t(const double* b, double* x, cudaStream_t s, LaunchCounter* lc, EventProfiler* prof) {
  const SymbolicFactor& S = S_;
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  const int n = S.nb * D;
  auto count = [&](int k = 1) { if (lc) lc->n += k; };
  double* y = d_z_.p;  // forward result / backward in place
  if (!forward_done_) {
    ScopedPhase ph(prof, PH_CH_FORWARD);
    chol_permute_in_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, d_perm_.p, b, y);
    count();
    for (int l = 0; l < S.nlevels; ++l) {
      const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
      if (nt == 0) continue;
      chol_forward_kernel<D><<<nt, kSolveThreads, 0, s>>>(P, Q, d_L_.p, d_Dinv_.p, y, d_contrib_.p, t0);
      count();
    }
  }
  {
    ScopedPhase ph(prof, PH_CH_BACKWARD);
    const size_t bsmem = (size_t)S.max_nrow * D * sizeof(double);
    for (int l = S.nlevels - 1; l >= 0; --l) {
      const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
      if (nt == 0) continue;
      chol_backward_kernel<D><<<nt, kSolveThreads, bsmem, s>>>(P, Q, d_L_.p, d_Dinv_.p, y, t0);
      count();
    }
    chol_permute_out_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, d_perm_.p, y, x, d_status_.p);
    count();
  }
  forward_done_ = false;
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::solve(const double* d_b, double* d_x, cudaStream_t s, LaunchCounter* lc, EventProfiler* prof) {
  if (S_.d == 3) solve_t<3>(d_b, d_x, s, lc, prof);
  else solve_t<6>(d_b, d_x, s, lc, prof);
}

}  // namespace g2o_b200

#ifdef CHOL_TIMING
extern "C" void b200_debug_chol_timing(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g2o_b200::g_chol_timing, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g2o_b200::g_chol_timing, z, sizeof(z)); }
}
#endif
