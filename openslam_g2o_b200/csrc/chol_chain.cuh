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

constexpr int kChR = 29;                          // front rows (block rows) the register file holds (29: 14 warps, 144 registers each)
constexpr int kChPairs = kChR * (kChR - 1) / 2;   // 406 off-diagonal blocks: threads 0..405, column-major
constexpr int kChDiag0 = (kChPairs + 31) / 32 * 32;  // 416: the diagonal warp - diagonal block + right-hand side of front row tid-416
constexpr int kChThreads = kChDiag0 + 32;         // 448
constexpr int kChWarps = kChThreads / 32;
constexpr int kChLd = 38;                         // doubles per row of the column buffer: 36 + pad (conflict-free 128-bit loads)
constexpr int kChFixedDoubles = 2 * kChR * kChLd + 2 * 48 + 72;  // column buffer x2 | pivot record x2 | rhs of the link's columns
constexpr int kChDinvSmem = (kMaxPanelCols * (kMaxPanelCols + 1) + kMaxPanelCols * kMaxPanelCols) * (int)sizeof(double);

// per-link descriptor (16 ints, built by CholeskyGpu::analyze): everything a link needs that would otherwise be a chain
// of dependent global loads.  Descriptors travel two links ahead of the computation (cp.async ring).
enum { CD_NROW = 0, CD_NCOL = 1, CD_COL0S = 2, CD_NFWD = 3, CD_LPTR = 4 /* 2 ints */, CD_DPTR = 6 /* 2 ints */, CD_MAPOFF = 8,
       CD_MAPCNT = 9, CD_COLPTR = 10, CD_PACK = 12 /* 2 ints */, CD_INTS = 16 };
__device__ __forceinline__ long long cd_i64(const int* d, int at) { return (long long)(unsigned)d[at] | ((long long)d[at + 1] << 32); }

#ifdef CHOL_TIMING
// cycle counters of the chain kernels (timing build): [i] sum, [16+i] events
//  0 link prologue (wait for the panel + add)   1 step: top -> barrier A (tid 0: waits for the pivot)   2 barrier A -> B (column solve)
//  3 own update (tid 0)   4 pivot block (pivot thread)   5 own update (tid 448: a warp that stays active)   6 re-index
//  7 backward: wait for the link's data   8 backward: the link's arithmetic
__device__ unsigned long long g_chain_timing[32];
#define CTK(i, cond) do { if (cond) { unsigned long long _n = clock64(); atomicAdd(&g_chain_timing[i], _n - _c0); atomicAdd(&g_chain_timing[16 + (i)], 1ull); _c0 = _n; } } while (0)
#define CTK_START unsigned long long _c0 = clock64()
#define CTK_RESET _c0 = clock64()
#else
#define CTK(i, cond) do {} while (0)
#define CTK_START do {} while (0)
#define CTK_RESET do {} while (0)
#endif

struct ChainDev {
  int nlinks;
  const int *desc, *map;
  const int *fwd_ptr, *fwd_src;
  int stage_doubles, remap_blocks;
};

__device__ __forceinline__ int ch_col_offset(int b) { return (kChR - 1) * b - b * (b - 1) / 2; }  // first thread of column b

// cp.async of a contiguous run (16-byte pieces; both sides 16-byte aligned).  No commit: the caller groups the pieces.
__device__ __forceinline__ void ch_async16(void* dst, const void* src, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared((char*)dst + i)), "l"((const char*)src + i) : "memory");
}
// TMA: one elected thread arms the mbarrier with the byte count and issues bulk copies (cp.async.bulk, UBLKCP in SASS);
// the copy engine moves the data and completes the barrier - no LSU issue slots, unlike per-thread cp.async
__device__ __forceinline__ void ch_mbar_init(unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
}
__device__ __forceinline__ void ch_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ch_tma(void* dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ch_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ch_async4(int* dst, const int* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)), "l"(src + i) : "memory");
}

constexpr int kChRemapLd = 38;  // doubles per block in the re-index buffer (36 + pad: conflict-free 128-bit accesses)

__global__ void __launch_bounds__(kChThreads, 1)
chol_chain_kernel(const __grid_constant__ ChainDev C, double* __restrict__ L, double* __restrict__ Ldiag, double* __restrict__ pack,
                  int* status, const double* __restrict__ y, double* __restrict__ z, const double* __restrict__ contrib) {
  constexpr int D = 6;
  extern __shared__ __align__(16) double ch_sm[];
  double* Lk = ch_sm;                         // [2][kChR][kChLd]: the finished block column, X_a[c*6 + i]
  double* piv = Lk + 2 * kChR * kChLd;        // [2][48]: L_kk (i*6+c, lower) | 1/diag (6) | y_k (6)
  double* rhs0 = piv + 2 * 48;                // [72]: (P b) - contributions from below the chain, for the link's columns
  double* stage = rhs0 + 72;                  // the link's panel as it lies in HBM | (P b) at the link's columns (72)
  double* remapT = stage + C.stage_doubles + 72;  // re-index buffer: off-diagonal blocks | diagonal blocks (with their right-hand sides)
  double* remapD = remapT + (size_t)C.remap_blocks * kChRemapLd;
  __shared__ int s_inv[32];
  __shared__ __align__(16) int s_desc[4][CD_INTS];   // descriptor ring: link j at slot j & 3
  __shared__ __align__(16) int s_map[2][32];         // re-index map INTO link j at slot j & 1
  __shared__ __align__(8) unsigned long long s_bar;  // completion of the staged panel (TMA)
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
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
  const bool diag_warp = warp == kChDiag0 / 32;
  // the largest column index held by this warp: the warp skips an update when every block of it is already final
  const int last_pair_lane = min(31, kChPairs - 1 - warp * 32);  // the last pair warp is partly filled, the diagonal warp holds none
  const int warp_max_b = __shfl_sync(0xffffffffu, b, max(last_pair_lane, 0));
  // T[i + 6 * j]: row i of front row a, row j of front row b  (element (a*6+i, b*6+j) of the front).  The diagonal
  // threads use the lower triangle only; the right-hand side of their front row rides in six of the unused upper slots
  // (CH_RHS), so that it needs no registers of its own and moves with the block when the front is re-indexed
  double T[36];
#define CH_RHS(i) T[((i) < 5 ? 6 * ((i) + 1) : 13)]
#pragma unroll
  for (int q = 0; q < 36; ++q) T[q] = 0.0;

  // everything link j+1 needs is requested while link j is eliminated: its panel and right-hand side by TMA (one
  // elected thread; needs descriptor j+1, which arrived one link earlier), the map into it and descriptor j+2 by
  // small cp.async copies
  auto request = [&](int jn) {  // jn = link whose data is requested; its descriptor is in shared memory
    const int* dn = s_desc[jn & 3];
    if (tid == 0) {
      const unsigned pbytes = (unsigned)(dn[CD_NROW] * D * dn[CD_NCOL] * D) * 8u, ybytes = (unsigned)(dn[CD_NCOL] * D) * 8u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was read through the generic proxy
      ch_mbar_expect(bar, pbytes + ybytes);
      ch_tma(stage, L + cd_i64(dn, CD_LPTR), pbytes, bar);
      ch_tma(stage + C.stage_doubles, y + dn[CD_COL0S], ybytes, bar);
    }
    if (warp == 1) {
      const int l = tid - 32;
      if (l < dn[CD_MAPCNT])
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(s_map[jn & 1] + l)), "l"(C.map + dn[CD_MAPOFF] + l) : "memory");
      if (jn + 1 < C.nlinks && l < CD_INTS / 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_desc[(jn + 1) & 3] + 4 * l)),
                     "l"(C.desc + (size_t)(jn + 1) * CD_INTS + 4 * l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tid == 0) { ch_mbar_init(bar); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < CD_INTS) s_desc[0][tid] = C.desc[tid];
  __syncthreads();
  request(0);
  CTK_START;
  for (int j = 0; j < C.nlinks; ++j) {
    CTK_RESET;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    ch_mbar_wait(bar, (unsigned)(j & 1));
    __syncthreads();
    const int* dj = s_desc[j & 3];
    const int nrow = dj[CD_NROW], ncol = dj[CD_NCOL], M = nrow * D, N = ncol * D;
    // right-hand side of the link's columns: (P b) minus what the supernodes below the chain contribute (fixed order)
    if (tid < N) {
      double s = 0.0;
      if (dj[CD_NFWD] > 0) {
        const int cc = dj[CD_COLPTR] + tid;
        for (int e = C.fwd_ptr[cc]; e < C.fwd_ptr[cc + 1]; ++e) s += __ldcg(contrib + C.fwd_src[e]);
      }
      rhs0[tid] = stage[C.stage_doubles + tid] - s;
    }
    // 1. the panel joins the front (128-bit loads: consecutive front rows are 48 bytes apart, conflict-free)
    if (offd && a < nrow && b < ncol) {
      const double* src = stage + (a * D) + (size_t)(b * D) * M;
#pragma unroll
      for (int jj = 0; jj < D; ++jj)
#pragma unroll
        for (int i = 0; i < D; i += 2) {
          const double2 v = *reinterpret_cast<const double2*>(src + i + (size_t)jj * M);
          T[i + D * jj] += v.x;
          T[i + 1 + D * jj] += v.y;
        }
    } else if (diag && a < ncol) {
      const double* src = stage + (a * D) + (size_t)(a * D) * M;
#pragma unroll
      for (int jj = 0; jj < D; ++jj)
#pragma unroll
        for (int i = jj; i < D; ++i) T[i + D * jj] += src[i + (size_t)jj * M];  // lower triangle only
    }
    __syncthreads();
    if (j + 1 < C.nlinks) request(j + 1);  // flies in while this link is eliminated
    CTK(0, tid == 0);
    // 2. block columns of the link.  Step k: everybody first applies block column k-1 to its block; the pivot block k is
    //    factored by the diagonal warp as soon as ITS update is done (the other warps start theirs a moment later, so
    //    that the diagonal warp has the FP64 pipe of its sub-partition to itself), the owners of column k wait for
    //    the pivot on a named barrier, solve, and publish the column; one full barrier per step.
    for (int k = 0; k < ncol; ++k) {
      double* pv = piv + (k & 1) * 48;
      double* Lc = Lk + (k & 1) * kChR * kChLd;
      // ---- pivot block k (diagonal warp) and the solve of block column k (its 1 or 2 owner warps: threads
      //      offset(k) .. offset(k) + kChR - 2 - k); everybody else goes straight to the barrier
      {
        const int cw0 = ch_col_offset(k) >> 5, cw1 = min((ch_col_offset(k) + kChR - 2 - k) >> 5, kChDiag0 / 32 - 1);
        const int n_colw = cw1 - cw0 + 1;
        if (diag_warp) {
          if (diag && a == k) {
            CTK_RESET;
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
#pragma unroll
              for (int i = c; i < D; ++i) pv[i * D + c] = T[i + D * c];
            }
            if (bad) *status = 1;
            CTK(4, true);
          }
          __syncwarp();
          asm volatile("bar.arrive 1, %0;" ::"r"(32 * (1 + n_colw)) : "memory");  // the pivot is in shared memory
          if (diag && a == k) {  // off the critical path: the factor block and y_k go to HBM
            double* Dj = Ldiag + cd_i64(dj, CD_DPTR);
            const int col0s = dj[CD_COL0S];
#pragma unroll
            for (int c = 0; c < D; ++c) {
              __stcg(z + col0s + k * D + c, pv[42 + c]);
#pragma unroll
              for (int i = c; i < D; ++i) __stcg(Dj + (k * D + i) + (size_t)(k * D + c) * N, T[i + D * c]);
            }
          }
        } else if (warp >= cw0 && warp <= cw1) {
          asm volatile("bar.sync 1, %0;" ::"r"(32 * (1 + n_colw)) : "memory");
          if (b == k && a < nrow) {
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
            double2* xs = reinterpret_cast<double2*>(Lc + a * kChLd);  // broadcast copy for the update: X_a[c*6 + i]
#pragma unroll
            for (int q = 0; q < 18; ++q) xs[q] = make_double2(T[2 * q], T[2 * q + 1]);
          }
        }
      }
      CTK(1, tid == 0);
      __syncthreads();
      CTK(2, tid == 0);
      if (tid == kChPairs - 1) CTK_RESET;
      if (offd && b == k && a < nrow) {
        // the factor itself, after the barrier: rows inside the diagonal block of the supernode go to Ldiag, the others
        // to the panel - the layout the backward sweep reads
        const int ldd = a < ncol ? N : M;
        double* dst = (a < ncol ? Ldiag + cd_i64(dj, CD_DPTR) : L + cd_i64(dj, CD_LPTR)) + (a * D) + (size_t)(k * D) * ldd;
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
          for (int i = 0; i < D; i += 2) __stcg(reinterpret_cast<double2*>(dst + i + (size_t)c * ldd), make_double2(T[i + D * c], T[i + 1 + D * c]));
        if (a >= ncol) {
          // ... and a second, PACKED copy of the rows below the diagonal block ((M-N) x N, contiguous per link, followed
          // by the inverse diagonal block): what the backward sweep of the chain fetches with one bulk copy per link
          const int B = M - N;
          double* dp = pack + cd_i64(dj, CD_PACK) + ((a - ncol) * D) + (size_t)(k * D) * B;
#pragma unroll
          for (int c = 0; c < D; ++c)
#pragma unroll
            for (int i = 0; i < D; i += 2) __stcg(reinterpret_cast<double2*>(dp + i + (size_t)c * B), make_double2(T[i + D * c], T[i + 1 + D * c]));
        }
      }
      // ---- rank-6 update of everything to the right of column k.  The diagonal warp goes first and alone on its
      //      sub-partition (the next pivot block is the head of the critical path); the owners of column k+1 need no
      //      head start either - they wait for that pivot anyway; all other warps start when the diagonal warp is done
      {
        int n_next = 0;
        bool next_col_warp = false;
        if (k + 1 < ncol) {
          const int nw0 = ch_col_offset(k + 1) >> 5, nw1 = min((ch_col_offset(k + 1) + kChR - 3 - k) >> 5, kChDiag0 / 32 - 1);
          n_next = nw1 - nw0 + 1;
          next_col_warp = warp >= nw0 && warp <= nw1;
        }
        if (!diag_warp && !next_col_warp) asm volatile("bar.sync 2, %0;" ::"r"(32 * (kChWarps - n_next)) : "memory");
        if (offd) {
          if (warp_max_b > k && b > k && a < nrow) {
            const double2* xa = reinterpret_cast<const double2*>(Lc + a * kChLd);
            const double2* xb = reinterpret_cast<const double2*>(Lc + b * kChLd);
#ifndef CH_UPDATE_VARIANT
#define CH_UPDATE_VARIANT 0
#endif
#if CH_UPDATE_VARIANT == 2
            double2 a01 = xa[0], a23 = xa[1], a45 = xa[2], b01 = xb[0], b23 = xb[1], b45 = xb[2];
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double av[D] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
              const double bv[D] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y};
              if (c + 1 < D) {  // the next column's operands fly while this one is multiplied
                a01 = xa[3 * c + 3]; a23 = xa[3 * c + 4]; a45 = xa[3 * c + 5];
                b01 = xb[3 * c + 3]; b23 = xb[3 * c + 4]; b45 = xb[3 * c + 5];
              }
#pragma unroll
              for (int jj = 0; jj < D; ++jj)
#pragma unroll
                for (int i = 0; i < D; ++i) T[i + D * jj] = fma(-av[i], bv[jj], T[i + D * jj]);
            }
#else
#if CH_UPDATE_VARIANT == 1
#pragma unroll 1
#else
#pragma unroll 2
#endif
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
#endif
          }
        } else if (diag_warp) {
          if (diag && a > k && a < nrow) {
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
          __syncwarp();
          asm volatile("bar.arrive 2, %0;" ::"r"(32 * (kChWarps - n_next)) : "memory");  // the other warps may start
        }
      }
      CTK(3, tid == 0);
      CTK(5, tid == kChPairs - 1);
    }
    // 3. what is left of the front moves into the row order of the next link
    if (j + 1 < C.nlinks) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");  // the map into the next link (and its descriptor) have long arrived
      const int nbelow = nrow - ncol;
      if (tid < 32) s_inv[tid] = -1;
      __syncthreads();
      const int* map = s_map[(j + 1) & 1];
      const int nrow_n = s_desc[(j + 1) & 3][CD_NROW];
      if (tid < nbelow) s_inv[map[tid]] = tid;
      if (offd && b >= ncol && a < nrow) {
        const int ap = a - ncol, bp = b - ncol;
        double2* o = reinterpret_cast<double2*>(remapT + (size_t)(bp * nbelow - bp * (bp + 1) / 2 + (ap - bp - 1)) * kChRemapLd);
#pragma unroll
        for (int q = 0; q < 18; ++q) o[q] = make_double2(T[2 * q], T[2 * q + 1]);
      } else if (diag && a >= ncol && a < nrow) {
        double2* o = reinterpret_cast<double2*>(remapD + (a - ncol) * kChRemapLd);
#pragma unroll
        for (int q = 0; q < 18; ++q) o[q] = make_double2(T[2 * q], T[2 * q + 1]);  // block + right-hand side
      }
      __syncthreads();
      {
        // receivers: inherited blocks come from the buffer, blocks of rows that enter the front here start from zero
        const int ia = a >= 0 && a < nrow_n ? s_inv[a] : -1, ib = b >= 0 && a < nrow_n ? s_inv[b] : -1;
        const bool inh = ia >= 0 && ib >= 0;  // the map is monotone: ia > ib for an off-diagonal block
        const double* in = offd ? remapT + (size_t)(inh ? ib * nbelow - ib * (ib + 1) / 2 + (ia - ib - 1) : 0) * kChRemapLd
                                : remapD + (inh ? ia : 0) * kChRemapLd;
        if ((offd || diag) && a < nrow_n) {
#pragma unroll
          for (int q = 0; q < 18; ++q) {
            const double2 v = reinterpret_cast<const double2*>(in)[q];
            T[2 * q] = inh ? v.x : 0.0;
            T[2 * q + 1] = inh ? v.y : 0.0;
          }
        }
      }
      // (the next link starts with a barrier before anything of this buffer is written again)
      CTK(6, tid == 0);
    }
  }
}

// inverses of the factored diagonal blocks of the chain links (the backward sweep multiplies by them): one CTA per link,
// thread j builds column j of the inverse - the same recurrence the chunk tasks run after their completion signal
__global__ void __launch_bounds__(128)
chol_chain_dinv_kernel(const __grid_constant__ CholDev P, const int* __restrict__ chain_sn, const long long* __restrict__ sn_dinvptr,
                       const double* __restrict__ Ldiag, const int* __restrict__ desc, double* __restrict__ pack,
                       double* __restrict__ Dinv) {
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
  const int* dj = desc + (size_t)blockIdx.x * CD_INTS;  // the inverse goes behind the packed rows of the link
  double* out = pack + cd_i64(dj, CD_PACK) + (size_t)(dj[CD_NROW] - dj[CD_NCOL]) * D * N;
  double* out2 = Dinv ? Dinv + sn_dinvptr[J] : nullptr;  // ... and, when the sparse inverse will follow (marginals), where the dataflow kernel keeps its inverses
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {
    const int c = q / N, r = q - c * N;  // out(r,c) = inv(r,c) = Zt[c + r*N]
    const double v = r >= c ? Zt[c + r * N] : 0.0;
    out[q] = v;
    if (Dinv) out2[q] = v;
  }
}

// backward sweep along the chain, top link first: x_J = L_JJ^-T (y_J - L21^T x_below).  One CTA.  Like the factorisation
// it keeps what the links share on chip: x at the rows of the current front lives in shared memory and is re-indexed
// from link to link with the same maps (no gather from HBM), and everything a link reads from HBM - the rows below its
// diagonal block and the inverse diagonal block (one packed record per link, written by the factorisation), the forward
// result, the map, the descriptor - is requested one link ahead (TMA bulk copies, double buffer).  A link costs four barriers and two small matrix-vector products.
__global__ void __launch_bounds__(kChThreads, 1)
chol_chain_backward_kernel(int nlinks, const int* __restrict__ desc, const int* __restrict__ chain_map,
                           const double* __restrict__ pack, double* __restrict__ y, int buf_doubles) {
  constexpr int D = 6;
  extern __shared__ __align__(16) double cb_sm[];
  double* xf = cb_sm;                     // [2][192]: x at the rows of the front (own columns first), per link parity
  double* tv = xf + 2 * 192;              // kMaxPanelCols
  double* bufs = tv + kMaxPanelCols;      // 2 x [L21 (B x N, ld B) | Dinv (N x N) | forward result at the link's columns (N)]
  __shared__ __align__(16) int s_desc[4][CD_INTS];
  __shared__ __align__(16) int s_map[2][32];   // map INTO link j+1 (positions of link j's below-rows), at slot j & 1
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  __shared__ __align__(8) unsigned long long s_bar[2];
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
  auto request = [&](int jn) {  // data of link jn (its descriptor is in shared memory), descriptor of link jn-1
    const int* dn = s_desc[jn & 3];
    const int M = dn[CD_NROW] * D, N = dn[CD_NCOL] * D, B = M - N;
    double* dst = bufs + (size_t)(jn & 1) * buf_doubles;
    const unsigned bar = bar0 + 8u * (unsigned)(jn & 1);
    if (tid == 0) {  // TMA: [rows below the diagonal block | inverse diagonal block] of the link in one bulk copy, the forward result
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      ch_mbar_expect(bar, (unsigned)(B * N + N * N + N) * 8u);
      ch_tma(dst, pack + cd_i64(dn, CD_PACK), (unsigned)(B * N + N * N) * 8u, bar);
      ch_tma(dst + B * N + N * N, y + dn[CD_COL0S], (unsigned)N * 8u, bar);
    }
    if (wid == 1) {
      if (jn + 1 < nlinks) {  // where this link's below-rows sit in the front of the link above
        const int* du = s_desc[(jn + 1) & 3];
        if (lane < du[CD_MAPCNT])
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(s_map[jn & 1] + lane)), "l"(chain_map + du[CD_MAPOFF] + lane) : "memory");
      }
      if (jn > 0 && lane < CD_INTS / 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_desc[(jn - 1) & 3] + 4 * lane)),
                     "l"(desc + (size_t)(jn - 1) * CD_INTS + 4 * lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tid == 0) { ch_mbar_init(bar0); ch_mbar_init(bar0 + 8u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < CD_INTS) s_desc[(nlinks - 1) & 3][tid] = desc[(size_t)(nlinks - 1) * CD_INTS + tid];
  __syncthreads();
  request(nlinks - 1);
  CTK_START;
  for (int j = nlinks - 1; j >= 0; --j) {
    CTK_RESET;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    ch_mbar_wait(bar0 + 8u * (unsigned)(j & 1), (unsigned)(((nlinks - 1 - j) >> 1) & 1));
    __syncthreads();  // link j's data and descriptor j-1 are in shared memory; the other buffer is free again
    CTK(7, tid == 0);
    if (j > 0) request(j - 1);
    const int* dj = s_desc[j & 3];
    const int nc = dj[CD_NCOL], nr = dj[CD_NROW];
    const int M = nr * D, N = nc * D, B = M - N;
    const double* cur = bufs + (size_t)(j & 1) * buf_doubles;
    double* xn = xf + (j & 1) * 192;             // this link's front
    const double* xo = xf + ((j + 1) & 1) * 192;  // the front of the link above
    if (tid < B) xn[N + tid] = xo[s_map[j & 1][tid / D] * D + tid % D];
    __syncthreads();
    CTK(9, tid == 0);
    // t = y_J - L21^T x_below : one warp per column, lanes stride the rows, fixed-order shuffle tree
    for (int c = wid; c < N; c += nw) {
      const double* cj = cur + c * B;
      double s0 = 0.0, s1 = 0.0;
      int i = lane;
      for (; i + 32 < B; i += 64) { s0 = fma(cj[i], xn[N + i], s0); s1 = fma(cj[i + 32], xn[N + i + 32], s1); }
      if (i < B) s0 = fma(cj[i], xn[N + i], s0);
      double s = s0 + s1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tv[c] = cur[B * N + N * N + c] - s;
    }
    CTK(10, tid == 0);
    __syncthreads();
    CTK(11, tid == 0);
    // x_J = Linv^T t : 4 threads per entry, each a quarter of the column of the inverse, fixed shuffle tree
    {
      const double* Di = cur + B * N;
      const int i = tid >> 2, q = tid & 3;
      double s = 0.0;
      if (i < N)
        for (int jj = i + q; jj < N; jj += 4) s = fma(Di[(size_t)i * N + jj], tv[jj], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (i < N && q == 0) {
        xn[i] = s;
        __stcg(y + dj[CD_COL0S] + i, s);  // for the tasks below the chain and the final un-permutation
      }
    }
    // (the barrier at the top of the next link orders xn / tv / the buffers)
    CTK(8, tid == 0);
  }
}

}  // namespace g2o_b200
