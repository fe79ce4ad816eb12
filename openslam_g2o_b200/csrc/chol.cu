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
constexpr int kCholThreads = 256;

struct CholPlanDev {
  const int *tile_sn, *tile_r0, *tile_c0, *tile_work_ptr;
  const int *work_u, *work_a0, *work_a1, *work_b0, *work_b1;
  const int *sn_tile_ptr, *sn_chunk_ptr, *chunk_sn, *chunk_b0, *chunk_nb;
  const long long* sn_dinvptr;
};

template <int D>
__device__ void update_tile(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, int tile,
                            double* __restrict__ acc /* kTile*kTile shared */) {
  constexpr int S = D / 3;  // 3x3 micro blocks per block edge
  const int tid = threadIdx.x, nt = blockDim.x;
  const int w0 = Q.tile_work_ptr[tile], w1 = Q.tile_work_ptr[tile + 1];
  if (w0 == w1) return;
  const int J = Q.tile_sn[tile], R0 = Q.tile_r0[tile], C0 = Q.tile_c0[tile];
  for (int i = tid; i < kTile * kTile; i += nt) acc[i] = 0.0;
  __syncthreads();
  for (int wi = w0; wi < w1; ++wi) {
    const int u = Q.work_u[wi];
    const int a0 = Q.work_a0[wi] * S, a1 = Q.work_a1[wi] * S, b0 = Q.work_b0[wi] * S, b1 = Q.work_b1[wi] * S;
    const int K = P.upd_k[u], p0 = P.upd_p0[u];
    const int Mk = P.sn_nrow[K] * D, Nk = P.sn_ncol[K] * D;
    const double* Kp = L + P.sn_lptr[K] + (long long)p0 * D;
    const int* rel = P.rel + P.upd_relptr[u];
    const int na = a1 - a0, nb = b1 - b0;
    for (int idx = tid; idx < na * nb; idx += nt) {
      const int mb = b0 + idx / na;
      const int ma = a0 + idx % na;
      if (ma < mb) continue;
      const double* ra = Kp + ma * 3;
      const double* rb = Kp + mb * 3;
      double c00 = 0, c01 = 0, c02 = 0, c10 = 0, c11 = 0, c12 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll 4
      for (int k = 0; k < Nk; ++k) {
        const double x0 = ra[0], x1 = ra[1], x2 = ra[2];
        const double y0 = rb[0], y1 = rb[1], y2 = rb[2];
        c00 = fma(x0, y0, c00); c01 = fma(x0, y1, c01); c02 = fma(x0, y2, c02);
        c10 = fma(x1, y0, c10); c11 = fma(x1, y1, c11); c12 = fma(x1, y2, c12);
        c20 = fma(x2, y0, c20); c21 = fma(x2, y1, c21); c22 = fma(x2, y2, c22);
        ra += Mk;
        rb += Mk;
      }
      const int ab = ma / S, bb = mb / S;
      const int tr = (rel[ab] - R0) * D + (ma - ab * S) * 3;
      const int tc = (rel[bb] - C0) * D + (mb - bb * S) * 3;
      double* dst = acc + tr + tc * kTile;
      dst[0] += c00; dst[1] += c10; dst[2] += c20;
      dst[kTile] += c01; dst[kTile + 1] += c11; dst[kTile + 2] += c21;
      dst[2 * kTile] += c02; dst[2 * kTile + 1] += c12; dst[2 * kTile + 2] += c22;
    }
    __syncthreads();
  }
  const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
  double* Pj = L + P.sn_lptr[J];
  const int rows = min(kTile, M - R0 * D), cols = min(kTile, N - C0 * D);
  for (int i = tid; i < rows * cols; i += nt) {
    const int c = i / rows, r = i - c * rows;
    Pj[(long long)(R0 * D + r) + (long long)(C0 * D + c) * M] -= acc[r + c * kTile];
  }
  __syncthreads();
}

template <int D>
__device__ void factor_chunk(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, double* __restrict__ Ldiag,
                             int chunk, bool write_diag, double* __restrict__ Sm, int* status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int J = Q.chunk_sn[chunk];
  const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
  const int crow0 = Q.chunk_b0[chunk] * D, crows = Q.chunk_nb[chunk] * D;
  const int R = N + crows;  // rows staged: the diagonal block, then this chunk's rows
  double* Pj = L + P.sn_lptr[J];
  for (int i = tid; i < R * N; i += nt) {
    const int c = i / R, r = i - c * R;
    const int gr = r < N ? r : crow0 + (r - N);
    Sm[i] = Pj[gr + (long long)c * M];
  }
  __syncthreads();
  const int row = tid;  // one panel row per thread (R <= 192 <= blockDim)
  const int ncb = N / D;
  bool bad = false;
  for (int jb = 0; jb < ncb; ++jb) {
    const int j0 = jb * D;
    // pivot block (already carries every earlier update): factor redundantly in registers
    double Lp[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) Lp[r][c] = Sm[(j0 + r) + (j0 + c) * R];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double s = Lp[k][k];
#pragma unroll
      for (int m = 0; m < k; ++m) s = fma(-Lp[k][m], Lp[k][m], s);
      if (!(s > 0.0)) { bad = true; s = 1.0; }  // d <= 0: not positive definite (csparse_helper.cpp:136)
      const double lkk = sqrt(s);
      Lp[k][k] = lkk;
#pragma unroll
      for (int r = k + 1; r < D; ++r) {
        double t = Lp[r][k];
#pragma unroll
        for (int m = 0; m < k; ++m) t = fma(-Lp[r][m], Lp[k][m], t);
        Lp[r][k] = t / lkk;
      }
    }
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
      for (int c = 0; c < D; ++c) {
        double t = x[c];
#pragma unroll
        for (int m = 0; m < c; ++m) t = fma(-x[m], Lp[c][m], t);
        x[c] = t / Lp[c][c];
      }
#pragma unroll
      for (int c = 0; c < D; ++c) Sm[row + (j0 + c) * R] = x[c];
    }
    __syncthreads();
    if (below) {
      const int cend = row < N ? row : N - 1;
      for (int c = j0 + D; c <= cend; ++c) {
        double v = Sm[row + c * R];
#pragma unroll
        for (int k = 0; k < D; ++k) v = fma(-x[k], Sm[c + (j0 + k) * R], v);
        Sm[row + c * R] = v;
      }
    }
    __syncthreads();
  }
  if (bad && tid == 0) *status = 1;
  for (int i = tid; i < R * N; i += nt) {
    const int c = i / R, r = i - c * R;
    if (r < N) {
      // the factored diagonal block goes to its own array: sibling chunk CTAs are still reading the unfactored
      // block from the panel (every chunk factors it redundantly), so it must not be overwritten in place
      if (write_diag && r >= c) Ldiag[Q.sn_dinvptr[J] + r + (long long)c * N] = Sm[i];
    } else {
      Pj[crow0 + (r - N) + (long long)c * M] = Sm[i];
    }
  }
  __syncthreads();
}

template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_fused_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, double* __restrict__ Ldiag, int task0, int* status) {
  extern __shared__ __align__(16) double smem[];
  double* acc = smem;                    // kTile*kTile
  double* Sm = smem + kTile * kTile;     // factor staging
  const int t = task0 + blockIdx.x;
  for (int q = P.task_ptr[t]; q < P.task_ptr[t + 1]; ++q) {
    const int J = P.task_sn[q];
    for (int tile = Q.sn_tile_ptr[J]; tile < Q.sn_tile_ptr[J + 1]; ++tile) update_tile<D>(P, Q, L, tile, acc);
    const int c0 = Q.sn_chunk_ptr[J], c1 = Q.sn_chunk_ptr[J + 1];
    for (int ch = c0; ch < c1; ++ch) factor_chunk<D>(P, Q, L, Ldiag, ch, ch == c0, Sm, status);
  }
}
template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_update_tiles_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, const int* __restrict__ tiles) {
  __shared__ __align__(16) double acc[kTile * kTile];
  update_tile<D>(P, Q, L, tiles[blockIdx.x], acc);
}
template <int D>
__global__ void __launch_bounds__(kCholThreads)
chol_factor_chunks_kernel(CholDev P, CholPlanDev Q, double* __restrict__ L, double* __restrict__ Ldiag,
                          const int* __restrict__ chunks, int* status) {
  extern __shared__ __align__(16) double smem[];
  const int ch = chunks[blockIdx.x];
  factor_chunk<D>(P, Q, L, Ldiag, ch, ch == Q.sn_chunk_ptr[Q.chunk_sn[ch]], smem, status);
}

// inverse of every triangular diagonal block (one CTA per supernode): the solves become matrix-vector products
template <int D>
__global__ void __launch_bounds__(128)
chol_invert_diag_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ Ldiag, double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double Ls[];  // N*N
  const int J = blockIdx.x;
  const int N = P.sn_ncol[J] * D;
  const double* Lj = Ldiag + Q.sn_dinvptr[J];
  double* out = Dinv + Q.sn_dinvptr[J];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) Ls[i] = Lj[i];
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    // column j of the inverse: solve L z = e_j by forward substitution, z stored in place in `out`
    double* z = out + (long long)j * N;
    for (int i = 0; i < j; ++i) z[i] = 0.0;
    for (int i = j; i < N; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) s = fma(-Ls[i + k * N], z[k], s);
      z[i] = s / Ls[i + i * N];
    }
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

constexpr int kMaxPanelCols = 96;

template <int D>
__global__ void __launch_bounds__(128)
chol_forward_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ L, const double* __restrict__ Dinv,
                    double* __restrict__ y, int task0) {
  __shared__ double tvec[kMaxPanelCols];
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = P.task_ptr[t]; q < P.task_ptr[t + 1]; ++q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int N = P.sn_ncol[J] * D;
    double* yj = y + (long long)col0 * D;
    for (int i = tid; i < N; i += nt) tvec[i] = yj[i];
    __syncthreads();
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
        tvec[(krows[p] - col0) * D + rr] -= s;
      }
      __syncthreads();
    }
    const double* Di = Dinv + Q.sn_dinvptr[J];
    for (int i = tid; i < N; i += nt) {
      double s = 0.0;
      for (int j = 0; j <= i; ++j) s = fma(Di[i + (long long)j * N], tvec[j], s);
      yj[i] = s;
    }
    __syncthreads();
  }
}

template <int D>
__global__ void __launch_bounds__(128)
chol_backward_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ L, const double* __restrict__ Dinv,
                     double* __restrict__ y, int task0) {
  __shared__ double tvec[kMaxPanelCols];
  const int t = task0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int q = P.task_ptr[t + 1] - 1; q >= P.task_ptr[t]; --q) {
    const int J = P.task_sn[q];
    const int col0 = P.sn_col0[J];
    const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
    const double* Pj = L + P.sn_lptr[J];
    const int* jrows = P.sn_rows + P.sn_rowptr[J];
    double* xj = y + (long long)col0 * D;
    // t = y_J - L21^T x_below : one warp per column, lanes stride the rows (coalesced), fixed-order shuffle tree
    for (int j = wid; j < N; j += nw) {
      const double* cj = Pj + (long long)j * M;
      double s = 0.0;
      for (int i = N + lane; i < M; i += 32) s = fma(cj[i], y[(long long)jrows[i / D] * D + (i % D)], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tvec[j] = xj[j] - s;
    }
    __syncthreads();
    const double* Di = Dinv + Q.sn_dinvptr[J];
    for (int i = tid; i < N; i += nt) {  // x_J = Linv^T t
      double s = 0.0;
      const double* ci = Di + (long long)i * N;
      for (int j = i; j < N; ++j) s = fma(ci[j], tvec[j], s);
      xj[i] = s;
    }
    __syncthreads();
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
  B200_CUDA(cudaFuncSetAttribute(chol_factor_chunks_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_invert_diag_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  done = true;
}
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
  d_level_tiles_.upload(S_.level_tiles, s); d_level_chunks_.upload(S_.level_chunks, s);
  up64(d_sn_dinvptr_, S_.sn_dinvptr, s, keep);
  d_L_.alloc((size_t)S_.factor_doubles);
  d_Dinv_.alloc((size_t)S_.dinv_doubles);
  d_Ldiag_.alloc((size_t)S_.dinv_doubles);
  d_y_.alloc((size_t)nb * d);
  d_status_.alloc(1);
  if (!host_only_flag()) {
    B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
    if (d == 3) set_smem_attrs<3>(); else set_smem_attrs<6>();
  }
  analyzed_ = true;
}

template <int D>
static void factor_t(const SymbolicFactor& S, const CholDev& P, const CholPlanDev& Q, int nblk, const double* dA,
                     const double* d_lambda, const long long* a_dst, const int* a_ld, const unsigned char* a_trans,
                     const long long* diag_dst, const int* diag_ld, double* L, double* Ldiag, double* Dinv, const int* level_tiles,
                     const int* level_chunks, int* status, cudaStream_t s, LaunchCounter* lc) {
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
    if (S.level_kind[l] == 0) {
      const size_t smem = (size_t)kTile * kTile * 8 + S.level_smem[l];
      chol_fused_kernel<D><<<nt, kCholThreads, smem, s>>>(P, Q, L, Ldiag, t0, status);
      if (lc) lc->n++;
    } else {
      const int ntile = S.level_tile_ptr[l + 1] - S.level_tile_ptr[l];
      const int nchunk = S.level_chunk_ptr[l + 1] - S.level_chunk_ptr[l];
      if (ntile > 0) {
        chol_update_tiles_kernel<D><<<ntile, kCholThreads, 0, s>>>(P, Q, L, level_tiles + S.level_tile_ptr[l]);
        if (lc) lc->n++;
      }
      chol_factor_chunks_kernel<D><<<nchunk, kCholThreads, S.level_smem[l], s>>>(P, Q, L, Ldiag, level_chunks + S.level_chunk_ptr[l], status);
      if (lc) lc->n++;
    }
  }
  const size_t ismem = (size_t)S.max_ncol * D * S.max_ncol * D * 8;
  chol_invert_diag_kernel<D><<<S.nsn, 128, ismem, s>>>(P, Q, Ldiag, Dinv);
  if (lc) lc->n++;
  B200_CUDA(cudaGetLastError());
}

CholDev CholeskyGpu::dev() const {
  return CholDev{d_sn_col0_.p, d_sn_ncol_.p, d_sn_nrow_.p, d_sn_rowptr_.p, d_sn_rows_.p, d_sn_lptr_.p, d_upd_ptr_.p,
                 d_upd_k_.p,   d_upd_p0_.p,  d_upd_p1_.p,  d_upd_relptr_.p, d_rel_.p,    d_task_ptr_.p, d_task_sn_.p};
}
CholPlanDev CholeskyGpu::plan() const {
  return CholPlanDev{d_tile_sn_.p, d_tile_r0_.p, d_tile_c0_.p, d_tile_work_ptr_.p, d_work_u_.p, d_work_a0_.p, d_work_a1_.p,
                     d_work_b0_.p, d_work_b1_.p, d_sn_tile_ptr_.p, d_sn_chunk_ptr_.p, d_chunk_sn_.p, d_chunk_b0_.p,
                     d_chunk_nb_.p, d_sn_dinvptr_.p};
}

void CholeskyGpu::factor(const double* dA, const double* d_lambda, cudaStream_t s, LaunchCounter* lc) {
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  if (S_.d == 3)
    factor_t<3>(S_, P, Q, nblk_, dA, d_lambda, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, d_diag_dst_.p, d_diag_ld_.p, d_L_.p,
                d_Ldiag_.p, d_Dinv_.p, d_level_tiles_.p, d_level_chunks_.p, d_status_.p, s, lc);
  else
    factor_t<6>(S_, P, Q, nblk_, dA, d_lambda, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, d_diag_dst_.p, d_diag_ld_.p, d_L_.p,
                d_Ldiag_.p, d_Dinv_.p, d_level_tiles_.p, d_level_chunks_.p, d_status_.p, s, lc);
}

template <int D>
static void solve_t(const SymbolicFactor& S, const CholDev& P, const CholPlanDev& Q, const int* perm, const double* L,
                    const double* Dinv, double* y, const double* b, double* x, const int* status, cudaStream_t s,
                    LaunchCounter* lc) {
  const int n = S.nb * D;
  chol_permute_in_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, perm, b, y);
  if (lc) lc->n++;
  for (int l = 0; l < S.nlevels; ++l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    chol_forward_kernel<D><<<nt, 128, 0, s>>>(P, Q, L, Dinv, y, t0);
    if (lc) lc->n++;
  }
  for (int l = S.nlevels - 1; l >= 0; --l) {
    const int t0 = S.level_ptr[l], nt = S.level_ptr[l + 1] - t0;
    if (nt == 0) continue;
    chol_backward_kernel<D><<<nt, 128, 0, s>>>(P, Q, L, Dinv, y, t0);
    if (lc) lc->n++;
  }
  chol_permute_out_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, perm, y, x, status);
  if (lc) lc->n++;
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::solve(const double* d_b, double* d_x, cudaStream_t s, LaunchCounter* lc) {
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  if (S_.d == 3) solve_t<3>(S_, P, Q, d_perm_.p, d_L_.p, d_Dinv_.p, d_y_.p, d_b, d_x, d_status_.p, s, lc);
  else solve_t<6>(S_, P, Q, d_perm_.p, d_L_.p, d_Dinv_.p, d_y_.p, d_b, d_x, d_status_.p, s, lc);
}

}  // namespace g2o_b200
