// chol_chain.cuh - the tail chain of the supernodal Cholesky (included by chol.cu; block dimension 6).
//
// Band-like reduced camera systems (a ring of cameras, a corridor) have an AMD elimination tree whose upper part is ONE
// path: supernode after supernode, each the only heavy child of the next.  The dataflow kernel pays a panel staging, an
// HBM store and two completion flags per link of such a path, and most of the step is that latency (round 1: 21 - 30 us
// per link).  Here the whole path is factored by ONE CTA that keeps the frontal matrix in REGISTERS and moves it up the
// chain (multifrontal style): thread (a, b), a > b, owns the 6 x 6 block between front rows a and b; warp 15 owns the
// diagonal blocks and the right-hand side.  A link
//   1. adds its panel (A + the updates of the supernodes below the chain, brought to HBM by the dataflow kernel and
//      prefetched into shared memory with cp.async during the previous link) to the blocks of its own columns,
//   2. eliminates its block columns right-looking - pivot block by the diagonal warp (in the shadow of the other
//      warps' rank-6 update), triangular solve of the column by its 30 owner threads, one broadcast of the finished
//      column through shared memory, rank-6 update of every remaining block straight in registers -
//   3. writes the factor columns to the same HBM panels the backward sweep reads, and
//   4. re-indexes what is left (the update matrix) into the row order of the next link through shared memory.
// Nothing is signalled, nothing is re-read from HBM; the forward substitution rides along in the diagonal warp.
//
// Determinism: every block is owned by one thread and updated in column order - bit-identical from run to run.
#pragma once

namespace g2o_b200 {

constexpr int kChR = 31;                          // front rows (block rows) the register file holds
constexpr int kChPairs = kChR * (kChR - 1) / 2;   // 465 off-diagonal blocks: threads 0..464, column-major
constexpr int kChDiag0 = 480;                     // threads 480..510: diagonal block + right-hand side of front row tid-480
constexpr int kChThreads = 512;
constexpr int kChLd = 38;                         // doubles per row of the column buffer: 36 + pad (conflict-free 128-bit loads)
constexpr int kChFixedDoubles = 2 * kChR * kChLd + 2 * 48 + 72;  // column buffer x2 | pivot record x2 | rhs of the link's columns
constexpr int kChDinvSmem = (kMaxPanelCols * (kMaxPanelCols + 1) + kMaxPanelCols * kMaxPanelCols) * (int)sizeof(double);

struct ChainDev {
  int nlinks;
  const int *sn, *mapptr, *map;
  const unsigned* new_rows;
  const int *colptr, *fwd_ptr, *fwd_src;
  int stage_doubles, remap_blocks;
};

__device__ __forceinline__ int ch_col_offset(int b) { return 30 * b - b * (b - 1) / 2; }  // first thread of column b

// cp.async of a contiguous run of doubles (16-byte pieces; n even, both sides 16-byte aligned)
__device__ __forceinline__ void ch_stage_async(double* dst, const double* src, int n) {
  for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kChThreads, 1)
chol_chain_kernel(const __grid_constant__ CholDev P, const __grid_constant__ ChainDev C, const long long* __restrict__ sn_dinvptr,
                  double* __restrict__ L, double* __restrict__ Ldiag, int* status, const double* __restrict__ y,
                  double* __restrict__ z, const double* __restrict__ contrib) {
  constexpr int D = 6;
  extern __shared__ __align__(16) double ch_sm[];
  double* Lk = ch_sm;                         // [2][kChR][kChLd]: the finished block column, X_a[c*6 + i]
  double* piv = Lk + 2 * kChR * kChLd;        // [2][48]: L_kk (i*6+c, lower) | 1/diag (6) | y_k (6)
  double* rhs0 = piv + 2 * 48;                // [72]: (P b) - contributions from below the chain, for the link's columns
  double* stage = rhs0 + 72;                  // the link's panel as it lies in HBM
  double* remapT = stage + C.stage_doubles;   // re-index buffer: off-diagonal blocks | diagonal blocks (with their right-hand sides)
  double* remapD = remapT + (size_t)C.remap_blocks * 36;
  __shared__ int s_inv[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // my block: (a, b) with a > b for tid < 465 (column-major), (a, a) for the diagonal warp
  int a = -1, b = -1;
  if (tid < kChPairs) {
    int col = 0;
    while (col + 1 < kChR - 1 && ch_col_offset(col + 1) <= tid) ++col;
    b = col;
    a = b + 1 + (tid - ch_col_offset(col));
  } else if (tid >= kChDiag0 && tid < kChDiag0 + kChR) {
    a = b = tid - kChDiag0;
  }
  const bool offd = tid < kChPairs, diag = a >= 0 && a == b;
  // the largest column index held by this warp: the warp skips an update when every block of it is already final
  const int last_pair_lane = min(31, kChPairs - 1 - warp * 32);  // warp 14 holds 17 pairs, warp 15 none
  const int warp_max_b = __shfl_sync(0xffffffffu, b, max(last_pair_lane, 0));
  // T[i + 6 * j]: row i of front row a, row j of front row b  (element (a*6+i, b*6+j) of the front).  The diagonal
  // threads use the lower triangle only; the right-hand side of their front row rides in six of the unused upper slots
  // (CH_RHS), so that it needs no registers of its own and moves with the block when the front is re-indexed
  double T[36];
#define CH_RHS(i) T[((i) < 5 ? 6 * ((i) + 1) : 13)]
#pragma unroll
  for (int q = 0; q < 36; ++q) T[q] = 0.0;

  {  // panel of the first link
    const int J = C.sn[0];
    ch_stage_async(stage, L + P.sn_lptr[J], P.sn_nrow[J] * D * P.sn_ncol[J] * D);
  }
  for (int j = 0; j < C.nlinks; ++j) {
    const int J = C.sn[j];
    const int nrow = P.sn_nrow[J], ncol = P.sn_ncol[J], M = nrow * D, N = ncol * D;
    const int col0s = P.sn_col0[J] * D;
    double* Pj = L + P.sn_lptr[J];
    double* Dj = Ldiag + sn_dinvptr[J];
    // right-hand side of the link's columns: (P b) minus what the supernodes below the chain contribute (fixed order)
    if (tid < N) {
      const int cc = C.colptr[j] + tid;
      double s = 0.0;
      for (int e = C.fwd_ptr[cc]; e < C.fwd_ptr[cc + 1]; ++e) s += __ldcg(contrib + C.fwd_src[e]);
      rhs0[tid] = y[col0s + tid] - s;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // 1. the panel joins the front
    if (offd && a < nrow && b < ncol) {
      const double* src = stage + (a * D) + (size_t)(b * D) * M;
#pragma unroll
      for (int jj = 0; jj < D; ++jj)
#pragma unroll
        for (int i = 0; i < D; ++i) T[i + D * jj] += src[i + (size_t)jj * M];
    } else if (diag && a < ncol) {
      const double* src = stage + (a * D) + (size_t)(a * D) * M;
#pragma unroll
      for (int jj = 0; jj < D; ++jj)
#pragma unroll
        for (int i = jj; i < D; ++i) T[i + D * jj] += src[i + (size_t)jj * M];  // lower triangle only
    }
    __syncthreads();
    if (j + 1 < C.nlinks) {  // the next panel flies in while this link is eliminated
      const int Jn = C.sn[j + 1];
      ch_stage_async(stage, L + P.sn_lptr[Jn], P.sn_nrow[Jn] * D * P.sn_ncol[Jn] * D);
    }
    // 2. block columns of the link
    for (int k = 0; k < ncol; ++k) {
      double* pv = piv + (k & 1) * 48;
      double* Lc = Lk + (k & 1) * kChR * kChLd;
      if (diag && a == k) {
        // pivot block: right-looking inside the block, dependent chain per column = rsqrt -> multiply -> one FMA
        bool bad = false;
        double r[D], yk[D];
#pragma unroll
        for (int i = 0; i < D; ++i) r[i] = CH_RHS(i) + rhs0[k * D + i];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double s = T[c + D * c];
          if (!(s > 0.0)) { bad = true; s = 1.0; }  // d <= 0: not positive definite (csparse_helper.cpp:136)
          const double ri = fast_rsqrt(s);
          pv[36 + c] = ri;
          T[c + D * c] = s * ri;
#pragma unroll
          for (int i = c + 1; i < D; ++i) T[i + D * c] *= ri;
#pragma unroll
          for (int c2 = c + 1; c2 < D; ++c2)
#pragma unroll
            for (int i = c2; i < D; ++i) T[i + D * c2] = fma(-T[i + D * c], T[c2 + D * c], T[i + D * c2]);
          // forward substitution rides along: y_c = (r_c - sum_{m<c} L(c,m) y_m) / L(c,c)
          double t = r[c];
#pragma unroll
          for (int m = 0; m < c; ++m) t = fma(-T[c + D * m], yk[m], t);
          yk[c] = t * ri;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
          pv[42 + c] = yk[c];
          __stcg(z + col0s + k * D + c, yk[c]);
#pragma unroll
          for (int i = c; i < D; ++i) {
            pv[i * D + c] = T[i + D * c];
            Dj[(k * D + i) + (size_t)(k * D + c) * N] = T[i + D * c];
          }
        }
        if (bad) *status = 1;
      }
      __syncthreads();
      if (offd && b == k && a < nrow) {
        // my block sits in the pivot's column: X = T L_kk^-T, row by row (chain per column = multiply -> one FMA)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double ri = pv[36 + c];
#pragma unroll
          for (int i = 0; i < D; ++i) T[i + D * c] *= ri;
#pragma unroll
          for (int m = c + 1; m < D; ++m) {
            const double l = pv[m * D + c];
#pragma unroll
            for (int i = 0; i < D; ++i) T[i + D * m] = fma(-T[i + D * c], l, T[i + D * m]);
          }
        }
        // broadcast copy for the update (X_a[c*6 + i]) and the factor itself: rows inside the diagonal block of the
        // supernode go to Ldiag, the others to the panel - the layout the backward sweep reads
        double2* xs = reinterpret_cast<double2*>(Lc + a * kChLd);
        double* dst = a < ncol ? Dj + (a * D) + (size_t)(k * D) * N : Pj + (a * D) + (size_t)(k * D) * M;
        const int ldd = a < ncol ? N : M;
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
          for (int i = 0; i < D; i += 2) {
            const double2 v = make_double2(T[i + D * c], T[i + 1 + D * c]);
            xs[(c * D + i) >> 1] = v;
            __stcg(reinterpret_cast<double2*>(dst + i + (size_t)c * ldd), v);
          }
      }
      __syncthreads();
      // rank-6 update of everything to the right of the pivot column
      if (offd) {
        if (warp_max_b > k && b > k && a < nrow) {
          const double2* xa = reinterpret_cast<const double2*>(Lc + a * kChLd);
          const double2* xb = reinterpret_cast<const double2*>(Lc + b * kChLd);
#pragma unroll 2
          for (int c = 0; c < D; ++c) {
            const double2 a01 = xa[3 * c], a23 = xa[3 * c + 1], a45 = xa[3 * c + 2];
            const double2 b01 = xb[3 * c], b23 = xb[3 * c + 1], b45 = xb[3 * c + 2];
            const double av[D] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
            const double bv[D] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y};
#pragma unroll
            for (int jj = 0; jj < D; ++jj)
#pragma unroll
              for (int i = 0; i < D; ++i) T[i + D * jj] = fma(-av[i], bv[jj], T[i + D * jj]);
          }
        }
      } else if (diag && a > k && a < nrow) {
        const double2* xa = reinterpret_cast<const double2*>(Lc + a * kChLd);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double2 a01 = xa[3 * c], a23 = xa[3 * c + 1], a45 = xa[3 * c + 2];
          const double av[D] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
          const double yc = pv[42 + c];
#pragma unroll
          for (int jj = 0; jj < D; ++jj)
#pragma unroll
            for (int i = jj; i < D; ++i) T[i + D * jj] = fma(-av[i], av[jj], T[i + D * jj]);
#pragma unroll
          for (int i = 0; i < D; ++i) CH_RHS(i) = fma(-av[i], yc, CH_RHS(i));
        }
      }
    }
    // 3. what is left of the front moves into the row order of the next link
    if (j + 1 < C.nlinks) {
      const int nbelow = nrow - ncol;
      const int* map = C.map + C.mapptr[j + 1];
      const int Jn = C.sn[j + 1];
      const int nrow_n = P.sn_nrow[Jn];
      if (tid < 32) s_inv[tid] = -1;
      __syncthreads();
      if (tid < nbelow) s_inv[map[tid]] = tid;
      if (offd && b >= ncol && a < nrow) {
        const int ap = a - ncol, bp = b - ncol;
        double2* o = reinterpret_cast<double2*>(remapT + (size_t)(bp * nbelow - bp * (bp + 1) / 2 + (ap - bp - 1)) * 36);
#pragma unroll
        for (int q = 0; q < 18; ++q) o[q] = make_double2(T[2 * q], T[2 * q + 1]);
      } else if (diag && a >= ncol && a < nrow) {
        const int ap = a - ncol;
#pragma unroll
        for (int q = 0; q < 36; ++q) remapD[ap * 36 + q] = T[q];  // block + right-hand side
      }
      __syncthreads();
      if (offd && a < nrow_n) {
        const int ia = s_inv[a], ib = s_inv[b];  // the map is monotone: ia > ib whenever both rows are inherited
        if (ia >= 0 && ib >= 0) {
          const double2* in = reinterpret_cast<const double2*>(remapT + (size_t)(ib * nbelow - ib * (ib + 1) / 2 + (ia - ib - 1)) * 36);
#pragma unroll
          for (int q = 0; q < 18; ++q) { const double2 v = in[q]; T[2 * q] = v.x; T[2 * q + 1] = v.y; }
        } else {
#pragma unroll
          for (int q = 0; q < 36; ++q) T[q] = 0.0;
        }
      } else if (diag && a < nrow_n) {
        const int ia = s_inv[a];
#pragma unroll
        for (int q = 0; q < 36; ++q) T[q] = ia >= 0 ? remapD[ia * 36 + q] : 0.0;
      }
      // (the next link starts with a barrier before anything of this buffer is written again)
    }
  }
}

// inverses of the factored diagonal blocks of the chain links (the backward sweep multiplies by them): one CTA per link,
// thread j builds column j of the inverse - the same recurrence the chunk tasks run after their completion signal
__global__ void __launch_bounds__(128)
chol_chain_dinv_kernel(const __grid_constant__ CholDev P, const int* __restrict__ chain_sn, const long long* __restrict__ sn_dinvptr,
                       const double* __restrict__ Ldiag, double* __restrict__ Dinv) {
  constexpr int D = 6;
  extern __shared__ __align__(16) double cd_sm[];
  double* Ls = cd_sm;                                       // L(i,k) at Ls[i + k*(N+1)]
  double* Zt = cd_sm + kMaxPanelCols * (kMaxPanelCols + 1);  // Zt[j + i*N] = inv(i,j)
  const int J = chain_sn[blockIdx.x];
  const int N = P.sn_ncol[J] * D;
  const double* Lj = Ldiag + sn_dinvptr[J];
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {
    const int c = q / N, r = q - c * N;
    Ls[r + c * (N + 1)] = r >= c ? Lj[q] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    for (int i = j; i < N; ++i) {
      double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0;
      int k = j;
      for (; k + 1 < i; k += 2) {
        s0 = fma(-Ls[i + k * (N + 1)], Zt[j + k * N], s0);
        s1 = fma(-Ls[i + (k + 1) * (N + 1)], Zt[j + (k + 1) * N], s1);
      }
      if (k < i) s0 = fma(-Ls[i + k * (N + 1)], Zt[j + k * N], s0);
      Zt[j + i * N] = (s0 + s1) / Ls[i + i * (N + 1)];
    }
  }
  __syncthreads();
  double* out = Dinv + sn_dinvptr[J];
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {
    const int c = q / N, r = q - c * N;  // out(r,c) = inv(r,c) = Zt[c + r*N]
    out[q] = r >= c ? Zt[c + r * N] : 0.0;
  }
}

// backward sweep along the chain, top link first: x_J = L_JJ^-T (y_J - L21^T x_below).  One CTA; the rows below the
// diagonal block and the inverse diagonal block of the NEXT link are prefetched (cp.async, double buffer) while the
// current one is computed, so a link costs a gather, two small matrix-vector products and three barriers.
__global__ void __launch_bounds__(kChThreads, 1)
chol_chain_backward_kernel(const __grid_constant__ CholDev P, int nlinks, const int* __restrict__ chain_sn,
                           const long long* __restrict__ sn_dinvptr, const double* __restrict__ L,
                           const double* __restrict__ Dinv, double* __restrict__ y, int buf_doubles, int nbuf) {
  constexpr int D = 6;
  extern __shared__ __align__(16) double cb_sm[];
  double* xb = cb_sm;                 // x at the rows below the diagonal block (<= 30 * 6)
  double* tv = xb + 192;              // kMaxPanelCols
  double* bufs = tv + kMaxPanelCols;  // nbuf x [L21 (B x N, ld B) | Dinv (N x N)]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  auto prefetch = [&](int j, double* dst) {
    const int J = chain_sn[j];
    const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D, B = M - N;
    const double* Pj = L + P.sn_lptr[J];
    const int hb = B >> 1;  // 16-byte pieces per column (B is a multiple of 6)
    for (int i = tid; i < hb * N; i += blockDim.x) {
      const int c = i / hb, r = (i - c * hb) * 2;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + r + c * B)), "l"(Pj + N + r + (size_t)c * M) : "memory");
    }
    const double* Di = Dinv + sn_dinvptr[J];
    for (int i = tid * 2; i < N * N; i += blockDim.x * 2)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + B * N + i)), "l"(Di + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(nlinks - 1, bufs);
  for (int j = nlinks - 1; j >= 0; --j) {
    const int J = chain_sn[j];
    const int nc = P.sn_ncol[J], nr = P.sn_nrow[J];
    const int M = nr * D, N = nc * D, B = M - N;
    const int* jrows = P.sn_rows + P.sn_rowptr[J];
    double* xj = y + (size_t)P.sn_col0[J] * D;
    double* cur = bufs + (size_t)((nlinks - 1 - j) % nbuf) * buf_doubles;
    if (nbuf > 1 && j > 0) prefetch(j - 1, bufs + (size_t)((nlinks - j) % nbuf) * buf_doubles);
    for (int i = tid; i < B; i += blockDim.x) xb[i] = __ldcg(y + ((size_t)jrows[nc + i / D] * D + (i % D)));
    if (nbuf > 1 && j > 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // t = y_J - L21^T x_below : one warp per column, lanes stride the rows, fixed-order shuffle tree
    for (int c = wid; c < N; c += nw) {
      const double* cj = cur + c * B;
      double s0 = 0.0, s1 = 0.0;
      int i = lane;
      for (; i + 32 < B; i += 64) { s0 = fma(cj[i], xb[i], s0); s1 = fma(cj[i + 32], xb[i + 32], s1); }
      if (i < B) s0 = fma(cj[i], xb[i], s0);
      double s = s0 + s1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tv[c] = __ldcg(xj + c) - s;
    }
    __syncthreads();
    // x_J = Linv^T t : 4 threads per entry, each a quarter of the column of the inverse, fixed shuffle tree
    {
      const double* Di = cur + B * N;
      const int i = tid >> 2, q = tid & 3;
      double s = 0.0;
      if (i < N)
        for (int jj = i + q; jj < N; jj += 4) s = fma(Di[(size_t)i * N + jj], tv[jj], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (i < N && q == 0) __stcg(xj + i, s);
    }
    __syncthreads();  // x_J is visible to the gathers of the links below; xb / tv / the buffer may be reused
    if (nbuf == 1 && j > 0) prefetch(j - 1, bufs);
  }
}

}  // namespace g2o_b200
