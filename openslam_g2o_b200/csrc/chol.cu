// chol.cu - supernodal left-looking sparse block Cholesky, sm_100a kernels + host driver.  See chol.h.
//
// Layout in HBM: L is one array of doubles; supernode s owns a dense column-major panel of
// (nrow_s*d) x (ncol_s*d) at sn_lptr[s] whose first ncol_s block rows are the (lower-triangular)
// diagonal block.  All index arrays are 32-bit block indices; panel offsets are 64-bit.
//
// Execution: ONE persistent dataflow kernel per factorisation (and one per backward sweep).  The numeric work is a
// list of tasks in level-major order (symbolic.h: flow_kind / flow_arg); a CTA takes the next task from a global
// counter and spins on the completion counters of what the task consumes.  Dependencies only point backwards in
// the list, so the earliest unfinished task is always held by a running CTA: no deadlock for any grid size, no
// level barriers, and the updates a supernode receives from early descendants are applied while the late ones are
// still being factored.
//
// Determinism: every panel tile is written by exactly one CTA and the updates it pulls from its descendants are
// applied in a fixed order, so repeated factorizations are bit-identical (no floating-point atomics).
#include "chol.h"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace g2o_b200 {

struct CholDev {
  const int *sn_col0, *sn_ncol, *sn_nrow, *sn_rowptr, *sn_rows;
  const long long* sn_lptr;
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

template <int D>
__global__ void chol_add_diag_extra_kernel(int nb, const long long* __restrict__ diag_dst, const int* __restrict__ diag_ld,
                                           const int* __restrict__ perm, const double* __restrict__ extra,
                                           double* __restrict__ L) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * D) return;
  const int k = idx / D, r = idx - k * D;   // k = permuted block column, perm[k] = its original index
  L[diag_dst[k] + r + (long long)r * diag_ld[k]] += extra[(long long)perm[k] * D + r];
}

// ---------------------------------------------------------------------------------------------
// numeric factorisation
//   GROUP  : one CTA per (destination tile of 48 x 48 scalars, split-K group): pulls the update pieces that land in
//            the tile, accumulates them in shared memory in list order, subtracts the sum from the panel (split
//            tiles: leaves a partial sum; the group that finishes last adds them in group order and subtracts)
//   CHUNK  : one CTA per (supernode, row chunk): diagonal block + chunk rows (+ the right-hand side as one more row)
//            staged in shared memory, then held in REGISTERS: warp = block column, lane = block row, one d x d block
//            per thread; blocked right-looking Cholesky with one barrier per block column
//   SUBTREE: small subtrees run both phases for all their supernodes inside one CTA
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 48;        // scalar rows / cols of a destination tile
constexpr int kMaxBlockCols = 12;  // block columns of a panel = warps of a CTA (symbolic.cpp caps the supernode width)
constexpr int kMaxPanelCols = 6 * kMaxBlockCols;
constexpr int kCholThreads = 32 * kMaxBlockCols;
constexpr int kSlices = 3;       // k slices of the tile product (kCholThreads = 128 register tiles x kSlices)
constexpr int kUpdateSmemDoubles = kTile * kTile + 2 * kTile * kMaxPanelCols + (kSlices - 1) * 128 * 18;  // acc | A rows | B rows | slice sums
constexpr int kLds = 33;         // lane stride of the staged panel: element (slot, i) of column c at (c*D+i)*kLds+slot

struct CholPlanDev {
  const int *tile_sn, *tile_r0, *tile_c0, *tile_work_ptr;
  const int *work_a0, *work_a1, *work_b0, *work_b1;
  const int *sn_tile_ptr, *sn_chunk_ptr, *chunk_sn, *chunk_b0, *chunk_nb;
  const long long *sn_dinvptr, *sn_cptr;
  const long long *work_koff, *work_reloff;
  const int *work_mk, *work_nk, *fwd_ptr, *fwd_src;
  const int* rel;
};

// dataflow state: task list, completion counters (zeroed before every factorisation)
struct CholFlowDev {
  const int *kind, *arg;
  int ntasks;
  int* next_task;
  int *upd_done, *chunk_done, *slot_done;
  const int *sn_nupd, *sn_nchunk, *work_ksn, *group_rtile;
  const int *g_tile, *g_w0, *g_w1, *g_slot;
  const int *r_tile, *r_slot0, *r_nslots;
  double* scratch;
};

constexpr unsigned kSpinLimit = 1u << 28;  // minutes of polling: far beyond any factorisation that fits the device
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// called by every thread of the CTA after its global writes: publishes them and bumps a completion counter
__device__ __forceinline__ void cta_signal(int* counter) {
  __syncthreads();
  // release at gpu scope by one thread after the CTA barrier: cumulative over the other threads' stores
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
}
// called by every thread: returns when *counter >= target (thread 0 spins, the barrier hands the acquire on)
__device__ __forceinline__ void cta_wait(const int* counter, int target) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire(counter) < target)
      if (++spins > kSpinLimit) __trap();  // a producer died: fail loudly instead of hanging the device
  }
  __syncthreads();
}

#ifdef CHOL_TIMING
__device__ unsigned long long g_chol_timing[32];
#define TCK(i) do { if (threadIdx.x == 0) { unsigned long long _n = clock64(); atomicAdd(&g_chol_timing[i], _n - _t0); atomicAdd(&g_chol_timing[16 + (i)], 1ull); _t0 = _n; } } while (0)
#define TCK_INIT unsigned long long _t0 = clock64()
// global-timer stamps per supernode (ns): [0] chunk saw its updates, [1] chunk signalled, [2] last group signalled an
// update of it, [3] first consumer saw it complete
__device__ unsigned long long g_chol_stamp[6][4096];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STAMP_SET(a, j) do { if (threadIdx.x == 0 && (j) < 4096) g_chol_stamp[a][j] = gtime(); } while (0)
#define STAMP_MAX(a, j) do { if (threadIdx.x == 0 && (j) < 4096) atomicMax(&g_chol_stamp[a][j], gtime()); } while (0)
#define STAMP_MIN(a, j) do { if (threadIdx.x == 0 && (j) < 4096) atomicMin(&g_chol_stamp[a][j], gtime()); } while (0)
#else
#define STAMP_SET(a, j) do {} while (0)
#define STAMP_MAX(a, j) do {} while (0)
#define STAMP_MIN(a, j) do {} while (0)
#define TCK(i) do {} while (0)
#define TCK_INIT do {} while (0)
#endif

// acc (shared, kTile x kTile) = sum over work items [w0,w1) of the tile, in list order.  Every item waits until its
// source supernode is completely stored (warp 0 polls up to 32 items at once).  Operand rows of the updating panel
// are staged in shared memory with coalesced L2 loads; every thread then owns 3x3 micro tiles of the product.
template <int D>
__device__ void accumulate_items(const CholPlanDev& Q, const CholFlowDev& F, const double* __restrict__ L, int w0, int w1,
                                 int R0, int C0, double* __restrict__ acc, double* __restrict__ As,
                                 double* __restrict__ Bs, int* s_nready) {
  constexpr int S = D / 3;
  const int tid = threadIdx.x, nt = blockDim.x;
  TCK_INIT;
  __syncthreads();  // whoever used the shared bytes before (a panel factorisation of the same task) is done
  for (int i = tid; i < kTile * kTile; i += nt) acc[i] = 0.0;
  int nready = w0;
  for (int wi = w0; wi < w1; ++wi) {
    // one level of indirection: everything the item needs sits in flat per-item arrays (static plan data:
    // fetched before the item's source is known to be complete, the loads overlap the polling)
    const int a0 = Q.work_a0[wi], a1 = Q.work_a1[wi], b0 = Q.work_b0[wi], b1 = Q.work_b1[wi];
    const int Mk = Q.work_mk[wi], Nk = Q.work_nk[wi];
    const double* Kp = L + Q.work_koff[wi];
    const int* rel = Q.rel + Q.work_reloff[wi];
    const int nA = (a1 - a0) * D, nB = (b1 - b0) * D;
    if (wi >= nready) {
      if (tid < 32) {
        const int idx = wi + tid;
        const int K = idx < w1 ? F.work_ksn[idx] : -1;
        const int target = K >= 0 ? F.sn_nchunk[K] : 0;
        int lead;
        unsigned spins = 0;
        do {
          if (++spins > kSpinLimit) __trap();
          const bool ok = K < 0 || ld_acquire(F.chunk_done + K) >= target;
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          lead = __ffs(~m) - 1;  // leading ready items; -1 when all 32 are
          if (lead < 0) lead = 32;
        } while (lead == 0);
        if (tid == 0) *s_nready = wi + lead;
      }
      __syncthreads();
      nready = *s_nready;
      TCK(0);
      STAMP_MIN(3, F.work_ksn[wi]);
      STAMP_MAX(5, F.work_ksn[wi]);
    }
    __syncthreads();  // previous item's operands fully consumed, acc zeroing done, s_nready read by everyone
    {
      // stage both operand row blocks with cp.async (no register staging: every copy of the item is in flight at
      // once).  The source panel is complete and this SM's L1 was invalidated by the acquire that observed its
      // completion counter, so the copies see the producers' stores.  d = 6: rows come in multiples of 48 bytes
      // -> 16-byte copies; d = 3: 8-byte copies.
      constexpr int V = (D % 2 == 0) ? 2 : 1;  // doubles per copy
      const int rA = nA / V, rB = nB / V;
      const int totA = rA * Nk, tot = totA + rB * Nk;
      for (int i = tid; i < tot; i += nt) {
        const bool isA = i < totA;
        const int ii = isA ? i : i - totA;
        const int nR = isA ? rA : rB;
        const int k = ii / nR, r = (ii - k * nR) * V;
        const double* src = Kp + ((isA ? a0 : b0) * D + r + (long long)k * Mk);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(As + (isA ? 0 : kTile * kMaxPanelCols) + r + k * kTile);
        if (V == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    TCK(11);
    // product: 6 x 3 register tiles (two 3-row units of the A side x one 3-column unit of the B side), the k range
    // cut into kSlices parts -> 8 x 16 tiles x 3 slices = all 384 threads.  A warp covers 8 A-side and 4 B-side
    // positions: its 128-bit operand loads are broadcasts (1-2 shared-memory wavefronts per instruction), so the
    // loop is bound by the FP64 pipe, not by shared memory.
    const int na = (a1 - a0) * S, nb = (b1 - b0) * S;   // 3-row units
    const int ks = tid >> 7, m = tid & 127;
    const int mi = m & 7, mj = m >> 3;
    const int u0 = 2 * mi, u1 = u0 + 1;
    const int mb = b0 * S + mj;
    // a tile entirely above the diagonal of the update is never read again: skip it
    const bool work = u0 < na && mj < nb && (a0 * S + min(u1, na - 1)) >= mb;
    double c[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i][0] = c[i][1] = c[i][2] = 0.0;
    if (work) {
      const int kc = (Nk + kSlices - 1) / kSlices;
      const int k0 = ks * kc, k1 = min(Nk, k0 + kc);
      const double2* ra = reinterpret_cast<const double2*>(As + mi * 6 + k0 * kTile);
      const double* rb = Bs + mj * 3 + k0 * kTile;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const double2 x01 = ra[0], x23 = ra[1], x45 = ra[2];
        const double y0 = rb[0], y1 = rb[1], y2 = rb[2];
        c[0][0] = fma(x01.x, y0, c[0][0]); c[0][1] = fma(x01.x, y1, c[0][1]); c[0][2] = fma(x01.x, y2, c[0][2]);
        c[1][0] = fma(x01.y, y0, c[1][0]); c[1][1] = fma(x01.y, y1, c[1][1]); c[1][2] = fma(x01.y, y2, c[1][2]);
        c[2][0] = fma(x23.x, y0, c[2][0]); c[2][1] = fma(x23.x, y1, c[2][1]); c[2][2] = fma(x23.x, y2, c[2][2]);
        c[3][0] = fma(x23.y, y0, c[3][0]); c[3][1] = fma(x23.y, y1, c[3][1]); c[3][2] = fma(x23.y, y2, c[3][2]);
        c[4][0] = fma(x45.x, y0, c[4][0]); c[4][1] = fma(x45.x, y1, c[4][1]); c[4][2] = fma(x45.x, y2, c[4][2]);
        c[5][0] = fma(x45.y, y0, c[5][0]); c[5][1] = fma(x45.y, y1, c[5][1]); c[5][2] = fma(x45.y, y2, c[5][2]);
        ra += kTile / 2;
        rb += kTile;
      }
    }
    TCK(12);
    // k slices 1.. leave their tiles in scratch; slice 0 adds them in slice order and scatters into the tile
    double* scr = Bs + kTile * kMaxPanelCols;  // (kSlices-1) x 128 x 18
    if (ks > 0 && work) {
      double* o = scr + ((ks - 1) * 128 + m) * 18;
#pragma unroll
      for (int i = 0; i < 6; ++i) { o[3 * i] = c[i][0]; o[3 * i + 1] = c[i][1]; o[3 * i + 2] = c[i][2]; }
    }
    __syncthreads();
    if (ks == 0 && work) {
#pragma unroll
      for (int q = 0; q < kSlices - 1; ++q) {
        const double* o = scr + (q * 128 + m) * 18;
#pragma unroll
        for (int i = 0; i < 6; ++i) { c[i][0] += o[3 * i]; c[i][1] += o[3 * i + 1]; c[i][2] += o[3 * i + 2]; }
      }
      const int bb = mb / S;
      const int tc = (rel[bb] - C0) * D + (mb - bb * S) * 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int u = u0 + h;
        const int ma = a0 * S + u;
        if (u < na && ma >= mb) {
          const int ab = ma / S;
          const int tr = (rel[ab] - R0) * D + (ma - ab * S) * 3;
          double* dst = acc + tr + tc * kTile;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            dst[i] += c[3 * h + i][0];
            dst[kTile + i] += c[3 * h + i][1];
            dst[2 * kTile + i] += c[3 * h + i][2];
          }
        }
      }
    }
    TCK(1);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// wide tiles (SymbolicFactor::wide): 96 rows x 72 columns = the whole width of a panel.  An update item multiplies up
// to 96 x 72 rows of the source panel by up to 72 x 72 of its rows (10 flops per byte fetched from L2 instead of 6 for
// the 48 x 48 tiles), and the product runs on the FP64 tensor path: mma.sync.m8n8k4.f64 (DMMA), 12 warps = 4 row
// groups x 3 column groups, every warp 3 x 3 fragments (24 x 24) in registers, 6 shared-memory fragment loads per 9
// DMMAs.  The k range of an item is cut into two halves of 36 = one pipeline stage each; the copies (cp.async, 16
// bytes) of the next stage - the second half, or the first half of the next item when its source supernode is already
// complete - fly while the current one is multiplied: one CTA barrier per stage.
// Operand layout in a stage: element (row, k) at k * ld + row with ld = 100 (A side) / 76 (B side), both = 4 mod 16, so
// the 32 lanes of a fragment load (row = lane / 4, k = lane % 4) hit 32 different 8-byte bank pairs.
// ---------------------------------------------------------------------------------------------
constexpr int kWR = 96, kWC = 72;          // scalar rows / columns of a wide destination tile
constexpr int kWLdA = 100, kWLdB = 76;
constexpr int kWKH = 36;                   // k extent of a stage
constexpr int kWStage = (kWLdA + kWLdB) * kWKH;
constexpr int kWAccLd = 100;               // leading dimension of the accumulator tile in shared memory
constexpr int kWMapInts = 2 * (kWR + kWC); // destination row / column of every product row / column, per item parity
constexpr int kWideSmemDoubles = kWAccLd * kWC + 2 * kWStage + kWMapInts / 2;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// the cp.async copies of the NEXT stage.  A thread owns up to 9 of them: column k = warp + 12 (c / 3), 16-byte pair
// j = lane + 32 (c % 3) of the A-side rows followed by the B-side rows.  They are issued one per k-step of the current
// product: a burst of ~100 LDGSTS per stage ahead of the DMMAs keeps the warps in the LSU queue while the tensor pipe
// idles (measured, tests/csrc/dmma_tile.cu: 5590 cycles per stage, against 4250 interleaved and a DMMA floor of 3888;
// cp.async.bulk per column segment is slower still, 6450: 72 copies of 576-768 bytes per stage)
struct WideCopy {
  const double *srcA, *srcB;   // first row of the A / B side in column 0 of the stage
  unsigned dst;                // shared address of the stage buffer
  int rA, rAB, nk, Mk;         // pairs of the A side, of both sides; columns (0: nothing to copy); source column stride
};
__device__ __forceinline__ void wide_copy(const WideCopy& cp, int c, int warp, int lane) {
  const int kk = c / 3, k = warp + 12 * kk, j = lane + 32 * (c - 3 * kk);
  if (k < cp.nk && j < cp.rAB) {
    const bool isA = j < cp.rA;
    const double* src = (isA ? cp.srcA + 2 * j : cp.srcB + 2 * (j - cp.rA)) + (long long)k * cp.Mk;
    const unsigned dst = cp.dst + 8u * (unsigned)(isA ? k * kWLdA + 2 * j : kWLdA * kWKH + k * kWLdB + 2 * (j - cp.rA));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
}
__device__ __forceinline__ void wide_copy_all(const WideCopy& cp, int warp, int lane) {
#pragma unroll
  for (int c = 0; c < 9; ++c) wide_copy(cp, c, warp, lane);
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// one stage of an item's product: c (3 x 3 fragments of this warp) += A(24 x 4 ks) * B(24 x 4 ks)^T, ks <= 9
template <bool FULL>
__device__ __forceinline__ void wide_product(double (&c)[3][3][2], const double* __restrict__ Ap, const double* __restrict__ Bp,
                                             int ks, unsigned mask, const WideCopy& cp, int warp, int lane) {
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    wide_copy(cp, s, warp, lane);
    if (s < ks) {
      const double a0 = Ap[0], a1 = Ap[8], a2 = Ap[16];
      const double b0 = Bp[0], b1 = Bp[8], b2 = Bp[16];
      if (FULL || (mask & 1u)) dmma_m8n8k4(c[0][0][0], c[0][0][1], a0, b0);
      if (FULL || (mask & 8u)) dmma_m8n8k4(c[1][0][0], c[1][0][1], a1, b0);
      if (FULL || (mask & 64u)) dmma_m8n8k4(c[2][0][0], c[2][0][1], a2, b0);
      if (FULL || (mask & 2u)) dmma_m8n8k4(c[0][1][0], c[0][1][1], a0, b1);
      if (FULL || (mask & 16u)) dmma_m8n8k4(c[1][1][0], c[1][1][1], a1, b1);
      if (FULL || (mask & 128u)) dmma_m8n8k4(c[2][1][0], c[2][1][1], a2, b1);
      if (FULL || (mask & 4u)) dmma_m8n8k4(c[0][2][0], c[0][2][1], a0, b2);
      if (FULL || (mask & 32u)) dmma_m8n8k4(c[1][2][0], c[1][2][1], a1, b2);
      if (FULL || (mask & 256u)) dmma_m8n8k4(c[2][2][0], c[2][2][1], a2, b2);
      Ap += 4 * kWLdA;
      Bp += 4 * kWLdB;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

struct WideItem {
  int a0, nA, b0, nB, Mk, Nk;   // block row offsets (relative to the update's first row), scalar extents
  const double* Kp;
  const int* rel;
};

template <int D>
__device__ __noinline__ void accumulate_items_wide(const CholPlanDev& Q, const CholFlowDev& F, const double* __restrict__ L, int w0, int w1,
                                      int R0, int C0, double* __restrict__ acc, double* __restrict__ stg,
                                      int* __restrict__ maps, int* s_nready /* [3] */) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = warp & 3, ng = warp >> 2, g = lane >> 2, q = lane & 3;
  TCK_INIT;
  __syncthreads();  // whoever used the shared bytes before (a panel factorisation of the same task) is done
  for (int i = tid; i < kWAccLd * kWC; i += kCholThreads) acc[i] = 0.0;
  if (w0 >= w1) { __syncthreads(); return; }

  auto load_item = [&](int wi) {
    WideItem it;
    const int a0 = Q.work_a0[wi], a1 = Q.work_a1[wi], b0 = Q.work_b0[wi], b1 = Q.work_b1[wi];
    it.a0 = a0; it.nA = (a1 - a0) * D; it.b0 = b0; it.nB = (b1 - b0) * D;
    it.Mk = Q.work_mk[wi]; it.Nk = Q.work_nk[wi];
    it.Kp = L + Q.work_koff[wi];
    it.rel = Q.rel + Q.work_reloff[wi];
    return it;
  };
  // warp 0: how many of the items wi, wi+1, ... (up to 32) have a complete source supernode; spin: until at least one
  auto poll = [&](int wi, bool spin, int* out) {
    const int idx = wi + lane;
    const int K = idx < w1 ? F.work_ksn[idx] : -1;
    const int target = K >= 0 ? F.sn_nchunk[K] : 0;
    int lead;
    unsigned spins = 0;
    do {
      if (++spins > kSpinLimit) __trap();
      const bool ok = K < 0 || ld_acquire(F.chunk_done + K) >= target;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      lead = __ffs(~m) - 1;  // leading ready items; -1 when all 32 are
      if (lead < 0) lead = 32;
    } while (spin && lead == 0);
    if (lane == 0) *out = wi + lead;
  };
  // copies of one stage: k columns [h * 36, ...) of both operand row blocks, the zero columns that pad k to a multiple
  // of 4, and (first half) the destination maps of the item
  // destination of my product row (tid < 96) / column (128 <= tid < 200) of an item: fetched one item ahead of the
  // issue that stores it into the shared maps (a global load there would stall the issuing warps for an L2 / HBM round trip)
  auto load_rel = [&](const WideItem& it) -> int {
    if (tid < kWR) return tid < it.nA ? (it.rel[it.a0 + tid / D] - R0) * D + tid % D : -1;
    const int cc = tid - 128;
    if (cc >= 0 && cc < kWC) return cc < it.nB ? (it.rel[it.b0 + cc / D] - C0) * D + cc % D : -1;
    return -1;
  };
  // one stage = k columns [h * 36, ...) of both operand row blocks: the copy descriptor, the zero columns that pad k to a
  // multiple of 4, and (first half) the destination maps of the item.  The copies themselves: wide_copy
  auto setup = [&](const WideItem& it, int wi, int h, int buf, int relv) -> WideCopy {
    double* As = stg + buf * kWStage;
    double* Bs = As + kWLdA * kWKH;
    const int k0 = h * kWKH, nk = min(it.Nk - k0, kWKH);
    WideCopy cp;
    cp.rA = it.nA >> 1; cp.rAB = cp.rA + (it.nB >> 1);   // 16-byte pairs (D = 6: rows come in multiples of 48 bytes)
    cp.srcA = it.Kp + (long long)it.a0 * D + (long long)k0 * it.Mk;
    cp.srcB = it.Kp + (long long)it.b0 * D + (long long)k0 * it.Mk;
    cp.dst = (unsigned)__cvta_generic_to_shared(As);
    cp.nk = nk; cp.Mk = it.Mk;
    if (nk & 3) {
      const int pad = 4 - (nk & 3);
      for (int i = tid; i < pad * (kWLdA + kWLdB); i += kCholThreads) {
        const int kk = i / (kWLdA + kWLdB), r = i - kk * (kWLdA + kWLdB);
        if (r < kWLdA) As[(nk + kk) * kWLdA + r] = 0.0;
        else Bs[(nk + kk) * kWLdB + (r - kWLdA)] = 0.0;
      }
    }
    if (h == 0) {
      int* mp = maps + (wi & 1) * (kWR + kWC);
      if (tid < kWR) mp[tid] = relv;
      else if (tid >= 128 && tid < 128 + kWC) mp[kWR + tid - 128] = relv;
    }
    return cp;
  };

  int nready = w0;
  if (warp == 0) poll(w0, true, s_nready + 2);
  __syncthreads();
  nready = s_nready[2];
  TCK(0);
  WideItem cur = load_item(w0), nxt = cur;
  wide_copy_all(setup(cur, w0, 0, 0, load_rel(cur)), warp, lane);
  int rel_nxt = -1;
  if (w0 + 1 < w1) { nxt = load_item(w0 + 1); rel_nxt = load_rel(nxt); }
  double c[3][3][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  int unit = 0;  // stages issued so far minus one: stage `unit` sits in buffer unit & 1
  bool asked = false;
  for (int wi = w0; wi < w1; ++wi) {
    // fragments this warp needs for the item: inside the operands and not entirely above the diagonal of the update
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int r0 = mg * 24 + 8 * i, c0 = ng * 24 + 8 * j;
        const bool need = r0 < cur.nA && c0 < cur.nB && cur.a0 + min(r0 + 7, cur.nA - 1) / D >= cur.b0 + c0 / D;
        mask |= (need ? 1u : 0u) << (3 * i + j);
      }
    const int nh = cur.Nk > kWKH ? 2 : 1;
    bool next_issued = false;
    for (int h = 0; h < nh; ++h, ++unit) {
      const bool last = h == nh - 1;
      const bool want_next = last && wi + 1 < w1;
      TCK(7);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      TCK(13);
      __syncthreads();  // stage `unit` has landed; every warp is done with stage unit-1 (and the scatter of item wi-1)
      TCK(14);
      if (asked) nready = max(nready, s_nready[(unit + 1) & 1]);   // the look taken during the previous stage
      // one look (no spinning) at the completion counters of the next sources whenever the item after the next is not
      // known to be complete: by warp 0 ahead of its share of this stage's product - its DMMAs then run while the
      // other two warps of its sub-partition have finished theirs, the pipe stays busy - and read one stage later
      asked = nready <= wi + 2 && nready < w1 && wi + 1 < w1;
      if (asked && warp == 0) poll(max(nready, wi + 1), false, s_nready + (unit & 1));
      WideCopy cp;
      cp.nk = 0;  // nothing to copy
      if (!last) cp = setup(cur, wi, 1, (unit + 1) & 1, -1);
      else if (want_next && wi + 1 < nready) { cp = setup(nxt, wi + 1, 0, (unit + 1) & 1, rel_nxt); next_issued = true; }
      TCK(11);
      {
        const double* Ap = stg + (unit & 1) * kWStage + q * kWLdA + mg * 24 + g;
        const double* Bp = stg + (unit & 1) * kWStage + kWLdA * kWKH + q * kWLdB + ng * 24 + g;
        const int ks = (min(cur.Nk - h * kWKH, kWKH) + 3) >> 2;
        if (mask == 0x1ffu) wide_product<true>(c, Ap, Bp, ks, mask, cp, warp, lane);   // the common case in large fronts: no predicates
        else if (mask) wide_product<false>(c, Ap, Bp, ks, mask, cp, warp, lane);
        else wide_copy_all(cp, warp, lane);
      }
      TCK(12);
    }
    // the item's product leaves the registers: element (r, c) of it goes to (map[r], map[c]) of the tile
    if (mask) {
      const int* mp = maps + (wi & 1) * (kWR + kWC);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int r = mg * 24 + 8 * i + g;
        const int tr = mp[r];
        const int ab = cur.a0 + r / D;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (mask & (1u << (3 * i + j))) {
            const int cc = ng * 24 + 8 * j + 2 * q;
            const int t0 = mp[kWR + cc], t1 = mp[kWR + cc + 1];
            if (tr >= 0 && t0 >= 0 && ab >= cur.b0 + cc / D) acc[tr + t0 * kWAccLd] += c[i][j][0];
            if (tr >= 0 && t1 >= 0 && ab >= cur.b0 + (cc + 1) / D) acc[tr + t1 * kWAccLd] += c[i][j][1];
          }
          c[i][j][0] = c[i][j][1] = 0.0;
        }
      }
    }
    TCK(1);
    if (wi + 1 < w1) {
      if (!next_issued) {  // the next source was not complete when this item started: wait for it now
        if (warp == 0) poll(wi + 1, true, s_nready + 2);
        __syncthreads();
        nready = max(nready, s_nready[2]);
        TCK(0);
        wide_copy_all(setup(nxt, wi + 1, 0, unit & 1, rel_nxt), warp, lane);
      }
      cur = nxt;
      if (wi + 2 < w1) { nxt = load_item(wi + 2); rel_nxt = load_rel(nxt); }
    }
  }
  __syncthreads();
}

template <int D, bool WIDE>
__device__ void subtract_tile(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, int tile,
                              const double* __restrict__ acc) {
  constexpr int TR = WIDE ? kWR : kTile, TC = WIDE ? kWC : kTile, LD = WIDE ? kWAccLd : kTile;
  const int J = Q.tile_sn[tile], R0 = Q.tile_r0[tile], C0 = Q.tile_c0[tile];
  const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
  double* Pj = L + P.sn_lptr[J];
  const int rows = min(TR, M - R0 * D), cols = min(TC, N - C0 * D);
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int c = i / rows, r = i - c * rows;
    double* p = Pj + ((long long)(R0 * D + r) + (long long)(C0 * D + c) * M);
    __stcg(p, __ldcg(p) - acc[r + c * LD]);
  }
}

// 1/sqrt(s) to full double accuracy: hardware seed (MUFU.RSQ64H, ~2^-22) + two Newton steps in double.  Sits on the
// critical path of every block column.
__device__ __forceinline__ double fast_rsqrt(double s) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  const double hs = -0.5 * s;
  r = r * fma(hs * r, r, 1.5);
  r = r * fma(hs * r, r, 1.5);
  return r;
}

struct ChunkGeom {
  int J, ncb, M, N, cnb, crow0, crows, Rp, rs, nslots, col0s;
  double* Pj;
};
template <int D>
__device__ __forceinline__ ChunkGeom chunk_geom(const CholDev& P, const CholPlanDev& Q, double* L, int chunk) {
  ChunkGeom g;
  g.J = Q.chunk_sn[chunk];
  g.ncb = P.sn_ncol[g.J];
  g.M = P.sn_nrow[g.J] * D;
  g.N = g.ncb * D;
  g.cnb = Q.chunk_nb[chunk];
  g.crow0 = Q.chunk_b0[chunk] * D;
  g.crows = g.cnb * D;
  g.Rp = g.N + g.crows;        // panel rows staged
  g.rs = g.ncb + g.cnb;        // slot of the right-hand side
  g.nslots = g.rs + 1;
  g.col0s = P.sn_col0[g.J] * D;
  g.Pj = L + P.sn_lptr[g.J];
  return g;
}
#define CHUNK_GEOM_LOCALS                                                                                     \
  const int tid = threadIdx.x, nt = blockDim.x;                                                               \
  const int lane = tid & 31, w = tid >> 5, nw = nt >> 5;                                                      \
  const int J = G.J, ncb = G.ncb, M = G.M, N = G.N, crow0 = G.crow0, crows = G.crows, Rp = G.Rp, rs = G.rs,  \
            nslots = G.nslots, col0s = G.col0s;                                                               \
  double* Pj = G.Pj;                                                                                          \
  double* piv = Sm + N * D * kLds; /* D*D factored pivot block | D reciprocal diagonal entries */             \
  (void)tid; (void)nt; (void)lane; (void)w; (void)nw; (void)J; (void)ncb; (void)M; (void)crow0; (void)crows; \
  (void)Rp; (void)rs; (void)nslots; (void)col0s; (void)Pj; (void)piv

// phase 1: plan prefetch, wait for the updates of the supernode, right-hand-side gather, panel -> shared memory
template <int D>
__device__ __forceinline__ void chunk_load(const ChunkGeom& G, const CholPlanDev& Q, double* __restrict__ Sm,
                                        const double* __restrict__ y, const double* contrib, const int* wait_counter,
                                        int wait_target) {
  CHUNK_GEOM_LOCALS;
  TCK_INIT;
  // before waiting for the updates of this supernode: everything that only depends on the (static) plan.
  // Right-hand side gather t_c = (P b)_c - sum of the descendants' contributions: thread c owns column c and
  // prefetches its list bounds and the first source indices
  constexpr int kPre = 8;
  __shared__ int s_ge0[kMaxPanelCols], s_ge1[kMaxPanelCols];
  int ge0 = 0, ge1 = 0, gsrc[kPre];
  double gy = 0.0;
  if (tid < N) {
    const int g = col0s + tid;
    ge0 = Q.fwd_ptr[g];
    ge1 = Q.fwd_ptr[g + 1];
    gy = y[g];
    s_ge0[tid] = ge0;
    s_ge1[tid] = ge1;
#pragma unroll
    for (int q = 0; q < kPre; ++q) gsrc[q] = ge0 + q < ge1 ? Q.fwd_src[ge0 + q] : -1;
  }
  if (wait_target > 0) {
    if (tid == 0) {
      unsigned spins = 0;
      while (ld_acquire(wait_counter) < wait_target)
        if (++spins > kSpinLimit) __trap();
    }
  }
  __syncthreads();
  TCK(6);
  STAMP_SET(0, J);
  {
    // one column per warp pass, lanes stride the rows: 8-byte cp.async straight into the slot layout (every copy
    // of the panel in flight at once; L1 was invalidated by the acquire above, so the tile updates are visible)
    for (int c = w; c < N; c += nw) {
      const double* src = Pj + (long long)c * M;
      double* dst = Sm + c * D * kLds;
      for (int r = lane; r < Rp; r += 32) {
        const int slot = r / D;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + (r - slot * D) * kLds + slot)),
                     "l"(src + (r < N ? r : crow0 + (r - N)))
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ... and while the copies fly: the right-hand side.  The threads owning a right-hand-side column add up their
    // contributions in list order
    if (tid < N) {
      double part = 0.0;
#pragma unroll
      for (int q = 0; q < kPre; ++q)
        if (gsrc[q] >= 0) part += __ldcg(contrib + gsrc[q]);
#pragma unroll
      for (int i = 1; i < D; ++i) Sm[(tid * D + i) * kLds + rs] = 0.0;
      if (ge1 - ge0 <= kPre) Sm[(tid * D) * kLds + rs] = gy - part;
    }
    // long lists (upper part of the tree): one warp per column, lanes stride the list, fixed shuffle tree
    for (int c = w; c < N; c += nw) {
      const int e0 = s_ge0[c], e1 = s_ge1[c];
      if (e1 - e0 <= kPre) continue;
      double part = 0.0;
      for (int e = e0 + lane; e < e1; e += 32) part += __ldcg(contrib + Q.fwd_src[e]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) Sm[(c * D) * kLds + rs] = y[col0s + c] - part;
    }
    TCK(8);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  TCK(2);
}

// phase 2: the register-resident factorisation (see above); leaves the finished panel in shared memory
template <int D>
__device__ __forceinline__ void chunk_core(const ChunkGeom& G, double* __restrict__ Sm, int* status,
                                           double* __restrict__ Ldiag, double* __restrict__ z, bool first) {
  CHUNK_GEOM_LOCALS;
  TCK_INIT;
  // lanes above the diagonal of the diagonal block hold nothing
  const bool active = w < ncb && lane < nslots && lane >= w;
  double B[D][D];  // B[i][j]: row i, column j of my block
  if (active) {
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int i = 0; i < D; ++i) B[i][j] = Sm[((w * D + j) * D + i) * kLds + lane];
  }
  for (int jb = 0; jb < ncb; ++jb) {
    if (w == jb) {
      if (lane == jb) {
        // right-looking inside the block: as soon as column k is scaled the remaining entries are updated with
        // independent FMAs, so the dependent chain per column is rsqrt -> multiply -> one FMA
        bool bad = false;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = B[k][k];
          if (!(s > 0.0)) { bad = true; s = 1.0; }  // d <= 0: not positive definite (csparse_helper.cpp:136)
          const double r = fast_rsqrt(s);
          piv[D * D + k] = r;
          B[k][k] = s * r;
#pragma unroll
          for (int i = k + 1; i < D; ++i) B[i][k] *= r;
#pragma unroll
          for (int c = k + 1; c < D; ++c)
#pragma unroll
            for (int i = c; i < D; ++i) B[i][c] = fma(-B[i][k], B[c][k], B[i][c]);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int c = 0; c <= i; ++c) piv[i * D + c] = B[i][c];
        if (bad) *status = 1;
      }
      __syncwarp();
      if (active && lane > jb) {
        // right-looking substitution per row: chain per column = multiply -> one FMA
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double r = piv[D * D + c];
#pragma unroll
          for (int i = 0; i < D; ++i) B[i][c] *= r;
#pragma unroll
          for (int m = c + 1; m < D; ++m) {
            const double l = piv[m * D + c];
#pragma unroll
            for (int i = 0; i < D; ++i) B[i][m] = fma(-B[i][c], l, B[i][m]);
          }
        }
      }
      if (active) {
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
          for (int i = 0; i < D; ++i) Sm[((jb * D + j) * D + i) * kLds + lane] = B[i][j];
      }
    }
    __syncthreads();
    if (w == jb && active) {
      // my block column is final and this warp has nothing left to do: its rows go to HBM now, in the shadow of
      // the remaining block columns.  Chunk rows -> panel; (first chunk only) diagonal block -> Ldiag: sibling chunk
      // CTAs are still reading the unfactored block from the panel, so it must not be overwritten in place;
      // forward-substitution result -> z: sibling chunks still read (P b)_J from y
      if (lane >= ncb && lane < rs) {
        double* dp = Pj + (crow0 + (lane - ncb) * D + (long long)(jb * D) * M);
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
          for (int i = 0; i < D; ++i) __stcg(dp + i + (long long)j * M, B[i][j]);
      } else if (first && lane < ncb) {
        double* dd = Ldiag + (lane * D + (long long)(jb * D) * N);
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
          for (int i = 0; i < D; ++i)
            if (lane > jb || i >= j) dd[i + (long long)j * N] = B[i][j];
      } else if (first && lane == rs) {
#pragma unroll
        for (int j = 0; j < D; ++j) z[jb * D + j] = B[0][j];
      }
    }
    // the warp of the next pivot column updates first: the others start when it is done, so that their updates
    // run in the shadow of its (latency-bound) pivot factorisation instead of competing for the FP64 pipe
    const int nupd_threads = 32 * (ncb - jb - 1);
    if (w > jb + 1 && w < ncb) asm volatile("bar.sync 1, %0;" ::"r"(nupd_threads) : "memory");
    if (active && w > jb) {
      // rank-D update of my block: B -= X(my block row) * X(block row w)^T over the finished block column jb
      const double* xa = Sm + (jb * D * D) * kLds + lane;
      const double* xb = Sm + (jb * D * D) * kLds + w;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double a[D], b[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          a[i] = xa[(k * D + i) * kLds];
          b[i] = xb[(k * D + i) * kLds];
        }
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
          for (int i = 0; i < D; ++i) B[i][j] = fma(-a[i], b[j], B[i][j]);
      }
    }
    if (w == jb + 1 && w < ncb) asm volatile("bar.arrive 1, %0;" ::"r"(nupd_threads) : "memory");
  }
  TCK(3);
}

// phase 3: panel rows, forward-substitution results and the contribution vector to global memory, completion signal,
// then (off the critical path) the inverse of the diagonal block
template <int D>
__device__ __forceinline__ void chunk_store(const ChunkGeom& G, const CholPlanDev& Q, double* __restrict__ Sm,
                                         double* __restrict__ Ldiag, double* __restrict__ Dinv, bool first,
                                         double* __restrict__ z, double* contrib, int* chunk_done) {
  CHUNK_GEOM_LOCALS;
  TCK_INIT;
  // the finished panel sits in shared memory (the last iteration ended with a barrier and no update)
  {
    TCK(9);
    // c_J = L21 y_J for this chunk's rows: what every ancestor will subtract from its right-hand side.
    // 4 threads per row, each a quarter of the columns, fixed shuffle tree
    double* cj = contrib + Q.sn_cptr[J] + (crow0 - N);
    for (int r0 = 0; r0 < crows; r0 += nt / 4) {
      const int r = r0 + (tid >> 2), q = tid & 3;
      double s = 0.0;
      if (r < crows) {
        const int lr = N + r, slot = lr / D, ii = lr - slot * D;
        for (int c = q; c < N; c += 4) s = fma(Sm[(c * D + ii) * kLds + slot], Sm[(c * D) * kLds + rs], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (r < crows && q == 0) __stcg(cj + r, s);
    }
  }
  TCK(10);
  cta_signal(chunk_done + J);
  STAMP_SET(1, J);
  TCK(4);
  if (first) {
    // off the critical path: inverse of the triangular diagonal block, so that the solves are matrix-vector
    // products.  Thread j builds column j of the inverse; Zt holds it transposed (Zt[j + i*N] = inv(i,j)) so that
    // the threads walk it conflict-free and the factor entries are warp-wide broadcasts.
    double* Zt = piv + D * D + D;
    for (int j = tid; j < N; j += nt) {
      for (int i = j; i < N; ++i) {
        const int si = i / D;
        const double* li = Sm + (i - si * D) * kLds + si;  // L(i,k) = li[k*D*kLds]
        double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0;
        int k = j;
        for (; k + 1 < i; k += 2) {
          s0 = fma(-li[k * D * kLds], Zt[j + k * N], s0);
          s1 = fma(-li[(k + 1) * D * kLds], Zt[j + (k + 1) * N], s1);
        }
        if (k < i) s0 = fma(-li[k * D * kLds], Zt[j + k * N], s0);
        Zt[j + i * N] = (s0 + s1) / li[i * D * kLds];
      }
    }
    __syncthreads();
    double* out = Dinv + Q.sn_dinvptr[J];
    for (int i = tid; i < N * N; i += nt) {
      const int c = i / N, r = i - c * N;  // out(r,c) = inv(r,c) = Zt[c + r*N]
      out[i] = r >= c ? Zt[c + r * N] : 0.0;
    }
    TCK(5);
  }
}

template <int D>
__device__ __noinline__ void factor_chunk(const CholDev& P, const CholPlanDev& Q, double* __restrict__ L, double* __restrict__ Ldiag,
                             double* __restrict__ Dinv, int chunk, bool first, double* __restrict__ Sm, int* status,
                             const double* __restrict__ y, double* __restrict__ z, double* contrib, int* chunk_done,
                             const int* wait_counter, int wait_target) {
  const ChunkGeom G = chunk_geom<D>(P, Q, L, chunk);
  chunk_load<D>(G, Q, Sm, y, contrib, wait_counter, wait_target);
  chunk_core<D>(G, Sm, status, Ldiag + Q.sn_dinvptr[G.J], z + G.col0s, first);
  chunk_store<D>(G, Q, Sm, Ldiag, Dinv, first, z, contrib, chunk_done);
}

template <int D, bool WIDE>
__global__ void __launch_bounds__(kCholThreads, 1)
chol_factor_flow_kernel(const __grid_constant__ CholDev P, const __grid_constant__ CholPlanDev Q,
                        const __grid_constant__ CholFlowDev F, double* __restrict__ L, double* __restrict__ Ldiag,
                        double* __restrict__ Dinv, int* status, const double* __restrict__ y, double* __restrict__ z,
                        double* contrib) {
  extern __shared__ __align__(16) double smem[];  // update operands and factor staging share the same bytes
  __shared__ int s_task, s_nready[3];
  constexpr int TR = WIDE ? kWR : kTile, TC = WIDE ? kWC : kTile, LD = WIDE ? kWAccLd : kTile;
  double* acc = smem;
  double* As = smem + (WIDE ? kWAccLd * kWC : kTile * kTile);   // wide: the two operand stages
  double* Bs = As + kTile * kMaxPanelCols;
  int* maps = reinterpret_cast<int*>(As + 2 * kWStage);         // wide only
  const int tid = threadIdx.x;
  auto accumulate = [&](int w0, int w1, int R0, int C0) {
    if constexpr (WIDE) accumulate_items_wide<D>(Q, F, L, w0, w1, R0, C0, acc, As, maps, s_nready);
    else accumulate_items<D>(Q, F, L, w0, w1, R0, C0, acc, As, Bs, s_nready);
  };
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(F.next_task, 1);
    __syncthreads();
    const int ti = s_task;
    if (ti >= F.ntasks) break;
    const int kind = F.kind[ti], arg = F.arg[ti];
    if (kind == 1) {  // GROUP
      const int tile = F.g_tile[arg];
      const int slot = F.g_slot[arg];
      const int J = Q.tile_sn[tile], R0 = Q.tile_r0[tile], C0 = Q.tile_c0[tile];
      // this task is the only writer of its tile: the values it will subtract from (A + lambda I, scattered by the
      // previous kernels) are fetched before the updates are even complete (narrow tiles: a group is a few
      // microseconds; a wide group runs for tens of microseconds and reads them at the end)
      constexpr int kPer = WIDE ? 1 : (kTile * kTile + kCholThreads - 1) / kCholThreads;
      double old[kPer];
      const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D;
      const int rows = min(TR, M - R0 * D), cols = min(TC, N - C0 * D);
      double* Pt = L + P.sn_lptr[J] + ((long long)R0 * D + (long long)C0 * D * M);
      if (!WIDE && slot < 0) {
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int i = tid + q * kCholThreads;
          const int c = i / rows, r = i - c * rows;
          old[q] = i < rows * cols ? __ldcg(Pt + (r + (long long)c * M)) : 0.0;
        }
      }
      accumulate(F.g_w0[arg], F.g_w1[arg], R0, C0);
      STAMP_MAX(4, J);
      if (slot < 0) {
        if constexpr (WIDE) {
          subtract_tile<D, true>(P, Q, L, tile, acc);
        } else {
#pragma unroll
          for (int q = 0; q < kPer; ++q) {
            const int i = tid + q * kCholThreads;
            const int c = i / rows, r = i - c * rows;
            if (i < rows * cols) __stcg(Pt + (r + (long long)c * M), old[q] - acc[r + c * kTile]);
          }
        }
        cta_signal(F.upd_done + J);
        STAMP_MAX(2, J);
      } else {
        double* out = F.scratch + (long long)slot * TR * TC;
        for (int i = tid; i < TR * TC; i += blockDim.x) {
          const int c = i / TR, r = i - c * TR;
          __stcg(out + i, acc[r + c * LD]);
        }
        // split tile: whichever group arrives last adds the partial sums in group order and subtracts once
        const int rt = F.group_rtile[arg];
        const int ns = F.r_nslots[rt];
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          s_nready[0] = atomicAdd(F.slot_done + rt, 1) == ns - 1;
          __threadfence();
        }
        __syncthreads();
        if (s_nready[0]) {
          const double* in = F.scratch + (long long)F.r_slot0[rt] * TR * TC;
          for (int i = tid; i < TR * TC; i += blockDim.x) {
            double s = 0.0;
            for (int g = 0; g < ns; ++g) s += __ldcg(in + ((long long)g * TR * TC + i));
            const int c = i / TR, r = i - c * TR;
            acc[r + c * LD] = s;
          }
          __syncthreads();
          subtract_tile<D, WIDE>(P, Q, L, tile, acc);
          cta_signal(F.upd_done + Q.tile_sn[tile]);
        }
      }
    } else if (kind == 3) {  // CHUNK
      const int J = Q.chunk_sn[arg];
      factor_chunk<D>(P, Q, L, Ldiag, Dinv, arg, arg == Q.sn_chunk_ptr[J], smem, status, y, z, contrib, F.chunk_done,
                      F.upd_done + J, F.sn_nupd[J]);
    } else {  // SUBTREE
      for (int q = P.task_ptr[arg]; q < P.task_ptr[arg + 1]; ++q) {
        const int J = P.task_sn[q];
        for (int tile = Q.sn_tile_ptr[J]; tile < Q.sn_tile_ptr[J + 1]; ++tile) {
          const int w0 = Q.tile_work_ptr[tile], w1 = Q.tile_work_ptr[tile + 1];
          if (w0 == w1) continue;
          accumulate(w0, w1, Q.tile_r0[tile], Q.tile_c0[tile]);
          subtract_tile<D, WIDE>(P, Q, L, tile, acc);
          __syncthreads();
        }
        const int c0 = Q.sn_chunk_ptr[J], c1 = Q.sn_chunk_ptr[J + 1];
        for (int ch = c0; ch < c1; ++ch)
          factor_chunk<D>(P, Q, L, Ldiag, Dinv, ch, ch == c0, smem, status, y, z, contrib, F.chunk_done, nullptr, 0);
      }
    }
  }
}

}  // namespace g2o_b200
#include "chol_chain.cuh"
#include "sparse_inverse.cuh"
namespace g2o_b200 {

// ---------------------------------------------------------------------------------------------
// triangular solves on the permuted vector (in place)
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

// backward sweep, dataflow: tasks from the top of the tree down; a task waits for the task holding the parent of its
// root supernode (which waited for its own parent, ...: every ancestor's x is final).  Before it waits it stages the
// rows below the diagonal block of its root panel and the inverse diagonal block in shared memory, so that the
// critical path per tree level is: flag -> gather x_below -> L21^T x_below -> Linv^T t -> flag.
constexpr int kSolveThreads = 512;

template <int D>
__global__ void __launch_bounds__(kSolveThreads, 1)
chol_backward_flow_kernel(CholDev P, CholPlanDev Q, const double* __restrict__ L, const double* __restrict__ Dinv,
                          double* __restrict__ y, int ntasks, const int* __restrict__ task_parent, int* next_task,
                          int* bdone, int xb_doubles, int stage_doubles, const unsigned char* __restrict__ task_skip) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  double* xb = smem;                 // x at the rows below the diagonal block
  double* tvec = xb + xb_doubles;    // kMaxPanelCols
  double* stage = tvec + kMaxPanelCols;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(next_task, 1);
    __syncthreads();
    if (s_task >= ntasks) break;
    const int t = ntasks - 1 - s_task;
    if (task_skip && task_skip[t]) {  // a link of the tail chain: chol_chain_backward_kernel has already solved it
      cta_signal(bdone + t);
      continue;
    }
    const int q_root = P.task_ptr[t + 1] - 1;
    bool staged = false;
    {
      const int J = P.task_sn[q_root];
      const int M = P.sn_nrow[J] * D, N = P.sn_ncol[J] * D, B = M - N;
      if (B * N + N * N <= stage_doubles) {
        staged = true;
        const double* Pj = L + P.sn_lptr[J];
        for (int i = tid; i < B * N; i += nt) {
          const int c = i / B, r = i - c * B;
          stage[i] = Pj[N + r + (long long)c * M];
        }
        const double* Di = Dinv + Q.sn_dinvptr[J];
        for (int i = tid; i < N * N; i += nt) stage[B * N + i] = Di[i];
      }
    }
    // ... and, still before the wait, where the x values of the root's rows will come from (static plan data)
    constexpr int kPreIdx = 2;
    long long gidx[kPreIdx];
    {
      const int J = P.task_sn[q_root];
      const int nc = P.sn_ncol[J], B = (P.sn_nrow[J] - nc) * D;
      const int* jrows = P.sn_rows + P.sn_rowptr[J];
#pragma unroll
      for (int q = 0; q < kPreIdx; ++q) {
        const int i = tid + q * nt;
        gidx[q] = i < B ? (long long)jrows[nc + i / D] * D + (i % D) : -1;
      }
    }
    const int tp = task_parent[t];
    if (tp >= 0) cta_wait(bdone + tp, 1);
    for (int q = q_root; q >= P.task_ptr[t]; --q) {
      const int J = P.task_sn[q];
      const int col0 = P.sn_col0[J];
      const int nc = P.sn_ncol[J], nr = P.sn_nrow[J];
      const int M = nr * D, N = nc * D, B = M - N;
      const bool st = staged && q == q_root;
      const double* Pj = L + P.sn_lptr[J];
      const int* jrows = P.sn_rows + P.sn_rowptr[J];
      double* xj = y + (long long)col0 * D;
      __syncthreads();
      // x at the rows below the diagonal block goes through shared memory in pieces of at most xb_doubles (a top
      // separator of any height: round 1 refused panels above ~4800 block rows); t_j is accumulated piece by piece in
      // a fixed order
      for (int b0 = 0; b0 == 0 || b0 < B; b0 += xb_doubles) {
        const int Bc = min(xb_doubles, B - b0);
        if (b0 > 0) __syncthreads();  // the previous piece of xb has been consumed
        if (q == q_root && b0 == 0) {
#pragma unroll
          for (int k = 0; k < kPreIdx; ++k)
            if (gidx[k] >= 0 && tid + k * nt < Bc) xb[tid + k * nt] = __ldcg(y + gidx[k]);
          for (int i = tid + kPreIdx * nt; i < Bc; i += nt) xb[i] = __ldcg(y + ((long long)jrows[nc + i / D] * D + (i % D)));
        } else {
          for (int i = tid; i < Bc; i += nt) xb[i] = __ldcg(y + ((long long)jrows[nc + (b0 + i) / D] * D + ((b0 + i) % D)));
        }
        __syncthreads();
        // t = y_J - L21^T x_below : one warp per column, lanes stride the rows, fixed-order shuffle tree
        for (int j = wid; j < N; j += nw) {
          const double* cj = (st ? stage + j * B : Pj + ((long long)j * M + N)) + b0;
          double s0 = 0.0, s1 = 0.0;
          int i = lane;
          for (; i + 32 < Bc; i += 64) { s0 = fma(cj[i], xb[i], s0); s1 = fma(cj[i + 32], xb[i + 32], s1); }
          if (i < Bc) s0 = fma(cj[i], xb[i], s0);
          double s = s0 + s1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) tvec[j] = (b0 == 0 ? xj[j] : tvec[j]) - s;
        }
      }
      __syncthreads();
      const double* Di = st ? stage + B * N : Dinv + Q.sn_dinvptr[J];
      for (int i = tid; i < N; i += nt) {  // x_J = Linv^T t
        const double* ci = Di + (long long)i * N;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = i;
        for (; j + 3 < N; j += 4) {
          s0 = fma(ci[j], tvec[j], s0); s1 = fma(ci[j + 1], tvec[j + 1], s1);
          s2 = fma(ci[j + 2], tvec[j + 2], s2); s3 = fma(ci[j + 3], tvec[j + 3], s3);
        }
        for (; j < N; ++j) s0 = fma(ci[j], tvec[j], s0);
        __stcg(xj + i, (s0 + s1) + (s2 + s3));
      }
    }
    cta_signal(bdone + t);
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
constexpr int kMaxDynSmem = 226 * 1024;  // 227 KB per CTA minus the static shared variables
// The opt-in is a per-DEVICE (per-context) function attribute: it is set on the device current at every analyze(), not
// once per process - a second solver context on another GPU of the same process needs it too.  Always the maximum,
// so that contexts with different panel sizes on one device cannot lower it under each other.
template <int D>
void set_smem_attrs() {
  B200_CUDA(cudaFuncSetAttribute(chol_factor_flow_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  if (D == 6) B200_CUDA(cudaFuncSetAttribute(chol_factor_flow_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_backward_flow_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_chain_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  B200_CUDA(cudaFuncSetAttribute(chol_chain_dinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChDinvSmem));
}
// profiler ids of the kernel groups inside the Cholesky (continue the numbering of solver.cu)
enum { PH_CH_SCATTER = 12, PH_CH_FLOW = 13, PH_CH_CHAIN = 15, PH_CH_CHAIN_BACKWARD = 16, PH_CH_BACKWARD = 19 };
}  // namespace

void CholeskyGpu::analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt,
                          cudaStream_t s) {
  S_ = g2o_b200::analyze(nb, d, colptr, rowidx, opt);
  nblk_ = colptr[nb];
  std::vector<std::vector<long long>> keep;
  d_sn_col0_.upload(S_.sn_col0, s); d_sn_ncol_.upload(S_.sn_ncol, s); d_sn_nrow_.upload(S_.sn_nrow, s);
  d_sn_rowptr_.upload(S_.sn_rowptr, s); d_sn_rows_.upload(S_.sn_rows, s);
  up64(d_sn_lptr_, S_.sn_lptr, s, keep);
  d_rel_.upload(S_.rel, s);
  d_task_ptr_.upload(S_.task_ptr, s); d_task_sn_.upload(S_.task_sn, s);
  up64(d_a_dst_, S_.a_dst, s, keep);
  up64(d_diag_dst_, S_.diag_dst, s, keep);
  d_a_ld_.upload(S_.a_ld, s); d_diag_ld_.upload(S_.diag_ld, s); d_perm_.upload(S_.perm, s);
  d_a_trans_.upload(S_.a_trans, s);
  d_tile_sn_.upload(S_.tile_sn, s); d_tile_r0_.upload(S_.tile_r0, s); d_tile_c0_.upload(S_.tile_c0, s);
  d_tile_work_ptr_.upload(S_.tile_work_ptr, s);
  d_work_a0_.upload(S_.work_a0, s); d_work_a1_.upload(S_.work_a1, s);
  d_work_b0_.upload(S_.work_b0, s); d_work_b1_.upload(S_.work_b1, s);
  d_sn_tile_ptr_.upload(S_.sn_tile_ptr, s); d_sn_chunk_ptr_.upload(S_.sn_chunk_ptr, s);
  d_chunk_sn_.upload(S_.chunk_sn, s); d_chunk_b0_.upload(S_.chunk_b0, s); d_chunk_nb_.upload(S_.chunk_nb, s);
  d_group_tile_.upload(S_.group_tile, s); d_group_w0_.upload(S_.group_w0, s); d_group_w1_.upload(S_.group_w1, s);
  d_group_slot_.upload(S_.group_slot, s); d_group_rtile_.upload(S_.group_rtile, s);
  d_rtile_tile_.upload(S_.rtile_tile, s); d_rtile_slot0_.upload(S_.rtile_slot0, s); d_rtile_nslots_.upload(S_.rtile_nslots, s);
  up64(d_sn_dinvptr_, S_.sn_dinvptr, s, keep);
  up64(d_sn_cptr_, S_.sn_cptr, s, keep);
  up64(d_work_koff_, S_.work_koff, s, keep);
  up64(d_work_reloff_, S_.work_reloff, s, keep);
  d_work_mk_.upload(S_.work_mk, s); d_work_nk_.upload(S_.work_nk, s); d_work_ksn_.upload(S_.work_ksn, s);
  d_fwd_ptr_.upload(S_.fwd_ptr, s); d_fwd_src_.upload(S_.fwd_src, s);
  d_flow_kind_.upload(S_.flow_kind, s); d_flow_arg_.upload(S_.flow_arg, s);
  d_sn_nupd_.upload(S_.sn_nupd, s); d_sn_nchunk_.upload(S_.sn_nchunk, s);
  d_task_parent_.upload(S_.task_parent, s);
  d_col2sn_.upload(S_.col2sn, s);
  spinv_planned_ = false;
  {  // tail chain
    d_chain_sn_.upload(S_.chain_sn, s); d_chain_mapptr_.upload(S_.chain_mapptr, s); d_chain_map_.upload(S_.chain_map, s);
    d_chain_new_rows_.upload(S_.chain_new_rows, s); d_chain_colptr_.upload(S_.chain_colptr, s);
    d_chain_fwd_ptr_.upload(S_.chain_fwd_ptr, s); d_chain_fwd_src_.upload(S_.chain_fwd_src, s);
    std::vector<unsigned char> skip(S_.task_on_chain.begin(), S_.task_on_chain.end());
    d_task_skip_.upload(skip, s);
    chain_smem_ = chain_back_smem_ = 0;
    if (!S_.chain_sn.empty()) {
      chain_smem_ = ((size_t)kChFixedDoubles + S_.chain_stage_doubles + 72 + (S_.chain_remap_blocks + kChR) * kChRemapLd) * sizeof(double);
      size_t need = 0;
      // per-link descriptors (chol_chain.cuh: CD_*)
      const int nl = (int)S_.chain_sn.size();
      std::vector<int> desc((size_t)nl * CD_INTS, 0);
      long long pack_total = 0;  // packed [rows below the diagonal block | inverse diagonal block] per link (backward sweep)
      for (int j = 0; j < nl; ++j) {
        const int J = S_.chain_sn[j];
        const size_t N = (size_t)S_.sn_ncol[J] * d, B = (size_t)(S_.sn_nrow[J] - S_.sn_ncol[J]) * d;
        need = std::max(need, B * N + N * N + N);
        int* dj = desc.data() + (size_t)j * CD_INTS;
        dj[CD_NROW] = S_.sn_nrow[J]; dj[CD_NCOL] = S_.sn_ncol[J]; dj[CD_COL0S] = S_.sn_col0[J] * d;
        dj[CD_NFWD] = S_.chain_fwd_ptr[S_.chain_colptr[j + 1]] - S_.chain_fwd_ptr[S_.chain_colptr[j]];
        const long long lp = S_.sn_lptr[J], dp = S_.sn_dinvptr[J];
        dj[CD_LPTR] = (int)(unsigned)(lp & 0xffffffffll); dj[CD_LPTR + 1] = (int)(lp >> 32);
        dj[CD_DPTR] = (int)(unsigned)(dp & 0xffffffffll); dj[CD_DPTR + 1] = (int)(dp >> 32);
        dj[CD_MAPOFF] = S_.chain_mapptr[j]; dj[CD_MAPCNT] = S_.chain_mapptr[j + 1] - S_.chain_mapptr[j];
        dj[CD_COLPTR] = S_.chain_colptr[j];
        dj[CD_PACK] = (int)(unsigned)(pack_total & 0xffffffffll); dj[CD_PACK + 1] = (int)(pack_total >> 32);
        pack_total += (long long)(B * N + N * N);
      }
      d_chain_pack_.alloc((size_t)pack_total);
      d_chain_desc_.upload(desc, s);
      need = (need + 1) & ~(size_t)1;
      chain_back_buf_doubles_ = (int)need;
      chain_back_smem_ = (2 * 192 + kMaxPanelCols + 2 * need) * sizeof(double);
    }
  }
  d_L_.alloc((size_t)S_.factor_doubles);
  d_Dinv_.alloc((size_t)S_.dinv_doubles);
  d_Ldiag_.alloc((size_t)S_.dinv_doubles);
  d_gscratch_.alloc((size_t)std::max(S_.max_group_slots, 1) * (S_.wide ? kWR * kWC : kTile * kTile));
  d_contrib_.alloc((size_t)std::max<int64_t>(S_.sn_cptr[S_.nsn], 1));
  d_y_.alloc((size_t)nb * d);
  d_z_.alloc((size_t)nb * d);
  // completion counters: [0] next factor task | [1] next backward task | status | upd_done | chunk_done | slot_done | bdone
  const int ntask = (int)S_.task_ptr.size() - 1;
  cnt_upd_ = 4;
  cnt_chunk_ = cnt_upd_ + S_.nsn;
  cnt_slot_ = cnt_chunk_ + S_.nsn;
  cnt_bdone_ = cnt_slot_ + (int)S_.rtile_tile.size();
  d_counters_.alloc((size_t)cnt_bdone_ + ntask);
  // shared memory of the two persistent kernels
  const int Nmax = S_.max_ncol * d;
  const size_t factor_doubles = (size_t)Nmax * d * kLds + d * d + d + (size_t)Nmax * Nmax;
  flow_smem_ = std::max((size_t)(S_.wide ? kWideSmemDoubles : kUpdateSmemDoubles), factor_doubles) * sizeof(double);
  xb_doubles_ = std::min((S_.max_nrow * d + 1) & ~1, 12288);   // taller panels go through in pieces (chol_backward_flow_kernel)
  if (const char* e = getenv("G2O_B200_XB_DOUBLES")) xb_doubles_ = std::max(48, std::min(xb_doubles_, atoi(e) & ~1));  // tests: force pieces
  stage_doubles_ = (int)((kMaxDynSmem - 1024) / sizeof(double)) - xb_doubles_ - kMaxPanelCols;
  if (!host_only_flag()) {
    B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
    if (flow_smem_ > (size_t)kMaxDynSmem || stage_doubles_ < 0) throw CudaError{cudaErrorInvalidValue, "panel too large for shared memory", __FILE__, __LINE__};
    if (chain_smem_ > (size_t)kMaxDynSmem || chain_back_smem_ > (size_t)kMaxDynSmem) throw CudaError{cudaErrorInvalidValue, "tail chain too large for shared memory (symbolic.cpp budget)", __FILE__, __LINE__};
    if (d == 3) set_smem_attrs<3>(); else set_smem_attrs<6>();
    int dev = 0, sms = 0, occ = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (d == 3) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_factor_flow_kernel<3, false>, kCholThreads, flow_smem_));
    else if (S_.wide) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_factor_flow_kernel<6, true>, kCholThreads, flow_smem_));
    else B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_factor_flow_kernel<6, false>, kCholThreads, flow_smem_));
    flow_grid_ = std::max(1, std::min((int)S_.flow_kind.size(), sms * std::max(occ, 1)));
    back_grid_ = std::max(1, std::min(ntask, sms));
    if (const char* e = getenv("G2O_B200_FLOW_GRID")) flow_grid_ = std::max(1, std::min(flow_grid_, atoi(e)));
  }
  analyzed_ = true;
}

CholDev CholeskyGpu::dev() const {
  return CholDev{d_sn_col0_.p, d_sn_ncol_.p, d_sn_nrow_.p, d_sn_rowptr_.p, d_sn_rows_.p, d_sn_lptr_.p, d_task_ptr_.p, d_task_sn_.p};
}
CholPlanDev CholeskyGpu::plan() const {
  return CholPlanDev{d_tile_sn_.p, d_tile_r0_.p, d_tile_c0_.p, d_tile_work_ptr_.p, d_work_a0_.p, d_work_a1_.p,
                     d_work_b0_.p, d_work_b1_.p, d_sn_tile_ptr_.p, d_sn_chunk_ptr_.p, d_chunk_sn_.p, d_chunk_b0_.p,
                     d_chunk_nb_.p, d_sn_dinvptr_.p, d_sn_cptr_.p, d_work_koff_.p, d_work_reloff_.p, d_work_mk_.p,
                     d_work_nk_.p, d_fwd_ptr_.p, d_fwd_src_.p, d_rel_.p};
}

template <int D>
void CholeskyGpu::factor_t(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
                           EventProfiler* prof) {
  const SymbolicFactor& S = S_;
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  double* L = d_L_.p;
  int* cnt = d_counters_.p;
  auto count = [&](int n = 1) { if (lc) lc->n += n; };
  {
    ScopedPhase ph(prof, PH_CH_SCATTER);
    B200_CUDA(cudaMemsetAsync(L, 0, (size_t)S.factor_doubles * sizeof(double), s));
    B200_CUDA(cudaMemsetAsync(cnt, 0, d_counters_.n * sizeof(int), s));
    chol_scatter_kernel<D><<<ceil_div((int64_t)nblk_ * D * D, 256), 256, 0, s>>>(dA, nblk_, d_a_dst_.p, d_a_ld_.p, d_a_trans_.p, L);
    count();
    if (d_lambda) {
      chol_add_lambda_kernel<D><<<ceil_div((int64_t)S.nb * D, 256), 256, 0, s>>>(S.nb, d_diag_dst_.p, d_diag_ld_.p, d_lambda, L);
      count();
    }
    if (d_diag_extra_) {
      chol_add_diag_extra_kernel<D><<<ceil_div((int64_t)S.nb * D, 256), 256, 0, s>>>(S.nb, d_diag_dst_.p, d_diag_ld_.p, d_perm_.p, d_diag_extra_, L);
      count();
    }
    // the forward substitution rides along with the factorisation (one more row per panel)
    chol_permute_in_kernel<D><<<ceil_div(S.nb * D, 256), 256, 0, s>>>(S.nb, d_perm_.p, d_b, d_y_.p);
    count();
  }
  {
    ScopedPhase ph(prof, PH_CH_FLOW);
    CholFlowDev F{d_flow_kind_.p, d_flow_arg_.p, (int)S.flow_kind.size(), cnt + 0, cnt + cnt_upd_, cnt + cnt_chunk_,
                  cnt + cnt_slot_, d_sn_nupd_.p, d_sn_nchunk_.p, d_work_ksn_.p, d_group_rtile_.p, d_group_tile_.p,
                  d_group_w0_.p, d_group_w1_.p, d_group_slot_.p, d_rtile_tile_.p, d_rtile_slot0_.p, d_rtile_nslots_.p,
                  d_gscratch_.p};
    if (!S.flow_kind.empty()) {
      if constexpr (D == 6) {
        if (S.wide) chol_factor_flow_kernel<6, true><<<flow_grid_, kCholThreads, flow_smem_, s>>>(P, Q, F, L, d_Ldiag_.p, d_Dinv_.p, cnt + 2, d_y_.p, d_z_.p, d_contrib_.p);
        else chol_factor_flow_kernel<6, false><<<flow_grid_, kCholThreads, flow_smem_, s>>>(P, Q, F, L, d_Ldiag_.p, d_Dinv_.p, cnt + 2, d_y_.p, d_z_.p, d_contrib_.p);
      } else {
        chol_factor_flow_kernel<D, false><<<flow_grid_, kCholThreads, flow_smem_, s>>>(P, Q, F, L, d_Ldiag_.p, d_Dinv_.p, cnt + 2, d_y_.p, d_z_.p, d_contrib_.p);
      }
      count();
    }
  }
  if (!S.chain_sn.empty()) {
    // the tail chain: everything below it is complete (same stream), its panels hold A + the updates from below
    ScopedPhase ph(prof, PH_CH_CHAIN);
    ChainDev C{(int)S.chain_sn.size(), d_chain_desc_.p, d_chain_map_.p, d_chain_fwd_ptr_.p, d_chain_fwd_src_.p,
               (int)S.chain_stage_doubles, (int)S.chain_remap_blocks};
    chol_chain_kernel<<<1, kChThreads, chain_smem_, s>>>(C, L, d_Ldiag_.p, d_chain_pack_.p, cnt + 2, d_y_.p, d_z_.p, d_contrib_.p);
    chol_chain_dinv_kernel<<<(int)S.chain_sn.size(), 128, kChDinvSmem, s>>>(P, d_chain_sn_.p, d_sn_dinvptr_.p, d_Ldiag_.p, d_chain_desc_.p, d_chain_pack_.p, keep_chain_inverses_ ? d_Dinv_.p : nullptr);
    count(2);
  }
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::factor(const double* dA, const double* d_lambda, const double* d_b, cudaStream_t s, LaunchCounter* lc,
                         EventProfiler* prof) {
  if (!d_b) throw CudaError{cudaErrorInvalidValue, "factor() needs the right-hand side", __FILE__, __LINE__};
  if (S_.d == 3) factor_t<3>(dA, d_lambda, d_b, s, lc, prof);
  else factor_t<6>(dA, d_lambda, d_b, s, lc, prof);
}

template <int D>
void CholeskyGpu::solve_t(double* x, cudaStream_t s, LaunchCounter* lc, EventProfiler* prof) {
  const SymbolicFactor& S = S_;
  const CholDev P = dev();
  const CholPlanDev Q = plan();
  const int n = S.nb * D;
  const int ntask = (int)S.task_ptr.size() - 1;
  int* cnt = d_counters_.p;
  auto count = [&](int k = 1) { if (lc) lc->n += k; };
  double* y = d_z_.p;  // forward result (written by the factorisation) / backward in place
  if (!S.chain_sn.empty()) {  // top of the tree first: the chain links, one CTA
    ScopedPhase ph(prof, PH_CH_CHAIN_BACKWARD);
    chol_chain_backward_kernel<<<1, kChThreads, chain_back_smem_, s>>>((int)S.chain_sn.size(), d_chain_desc_.p, d_chain_map_.p,
                                                                       d_chain_pack_.p, y, chain_back_buf_doubles_);
    count();
  }
  ScopedPhase ph(prof, PH_CH_BACKWARD);
  const size_t bsmem = ((size_t)xb_doubles_ + kMaxPanelCols + stage_doubles_) * sizeof(double);
  if ((int)S.chain_sn.size() < ntask) {
    chol_backward_flow_kernel<D><<<back_grid_, kSolveThreads, bsmem, s>>>(P, Q, d_L_.p, d_Dinv_.p, y, ntask, d_task_parent_.p,
                                                                          cnt + 1, cnt + cnt_bdone_, xb_doubles_, stage_doubles_,
                                                                          S.chain_sn.empty() ? nullptr : d_task_skip_.p);
    count();
  }
  chol_permute_out_kernel<D><<<ceil_div(n, 256), 256, 0, s>>>(S.nb, d_perm_.p, y, x, cnt + 2);
  count();
  B200_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// sparse inverse subset (sparse_inverse.cuh)
// ---------------------------------------------------------------------------------------------
void CholeskyGpu::sparse_inverse(cudaStream_t s, LaunchCounter* lc) {
  const SymbolicFactor& S = S_;
  if (!spinv_planned_) {
    // depth levels of the supernodal tree (root = 0); per level its supernodes and its (supernode, block row) items
    std::vector<int> depth(S.nsn, 0);
    int nlev = 0;
    for (int J = S.nsn - 1; J >= 0; --J) {
      depth[J] = S.sn_parent[J] < 0 ? 0 : depth[S.sn_parent[J]] + 1;
      nlev = std::max(nlev, depth[J] + 1);
    }
    spinv_level_ptr_.assign(nlev + 1, 0);
    spinv_item_ptr_.assign(nlev + 1, 0);
    for (int J = 0; J < S.nsn; ++J) {
      spinv_level_ptr_[depth[J] + 1]++;
      spinv_item_ptr_[depth[J] + 1] += S.sn_nrow[J] - S.sn_ncol[J];
    }
    for (int l = 0; l < nlev; ++l) { spinv_level_ptr_[l + 1] += spinv_level_ptr_[l]; spinv_item_ptr_[l + 1] += spinv_item_ptr_[l]; }
    std::vector<int> level_sn(S.nsn), item_sn(std::max(spinv_item_ptr_[nlev], 1)), item_p(std::max(spinv_item_ptr_[nlev], 1));
    std::vector<int> lf(spinv_level_ptr_.begin(), spinv_level_ptr_.end() - 1), itf(spinv_item_ptr_.begin(), spinv_item_ptr_.end() - 1);
    for (int J = 0; J < S.nsn; ++J) {
      level_sn[lf[depth[J]]++] = J;
      for (int p = S.sn_ncol[J]; p < S.sn_nrow[J]; ++p) { item_sn[itf[depth[J]]] = J; item_p[itf[depth[J]]++] = p; }
    }
    d_spinv_level_sn_.upload(level_sn, s); d_spinv_item_sn_.upload(item_sn, s); d_spinv_item_p_.upload(item_p, s);
    d_Zinv_.alloc((size_t)S.factor_doubles); d_Yt_.alloc((size_t)S.factor_doubles);
    B200_CUDA(cudaStreamSynchronize(s));  // the temporaries above die here
    spinv_planned_ = true;
  }
  const SpinvDev V{dev(), d_sn_dinvptr_.p, d_col2sn_.p};
  auto count = [&](int k = 1) { if (lc) lc->n += k; };
  auto run = [&](auto dtag) {
    constexpr int D = decltype(dtag)::value;
    spinv_prepare_kernel<D><<<dim3(S.nsn, S.max_nrow > 256 ? 16 : 2), kSpinvThreads, 0, s>>>(V, d_L_.p, d_Dinv_.p, d_Yt_.p, d_Zinv_.p);
    count();
    const int nlev = (int)spinv_level_ptr_.size() - 1;
    for (int l = 0; l < nlev; ++l) {
      const int ni = spinv_item_ptr_[l + 1] - spinv_item_ptr_[l], ns = spinv_level_ptr_[l + 1] - spinv_level_ptr_[l];
      if (ni > 0) {
        spinv_rows_kernel<D><<<ni, kSpinvThreads, 0, s>>>(V, d_spinv_item_sn_.p, d_spinv_item_p_.p, spinv_item_ptr_[l], d_Yt_.p, d_Zinv_.p);
        spinv_diag_kernel<D><<<dim3(ns, S.max_ncol * D), kSpinvThreads, 0, s>>>(V, d_spinv_level_sn_.p, spinv_level_ptr_[l], d_Yt_.p, d_Zinv_.p);
        count(2);
      }
    }
  };
  if (S.d == 3) run(std::integral_constant<int, 3>{}); else run(std::integral_constant<int, 6>{});
  B200_CUDA(cudaGetLastError());
}

bool CholeskyGpu::locate_inverse_block(int r, int c, long long* off, int* ld, bool* transposed) const {
  const SymbolicFactor& S = S_;
  int pr = S.pinv[r], pc = S.pinv[c];
  *transposed = pr < pc;
  if (pr < pc) std::swap(pr, pc);
  const long long o = spinv_locate(pr, pc, S.d, S.col2sn.data(), S.sn_col0.data(), S.sn_ncol.data(), S.sn_nrow.data(),
                                   S.sn_rowptr.data(), S.sn_rows.data(), reinterpret_cast<const long long*>(S.sn_lptr.data()), ld);
  *off = o;
  return o >= 0;
}

void CholeskyGpu::gather_inverse_blocks(int n, const long long* d_off, const int* d_ld, const unsigned char* d_trans,
                                        double* d_out, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return;
  if (S_.d == 3) spinv_gather_kernel<3><<<ceil_div((int64_t)n * 9, 256), 256, 0, s>>>(n, d_off, d_ld, d_trans, d_Zinv_.p, d_out);
  else spinv_gather_kernel<6><<<ceil_div((int64_t)n * 36, 256), 256, 0, s>>>(n, d_off, d_ld, d_trans, d_Zinv_.p, d_out);
  if (lc) lc->n++;
  B200_CUDA(cudaGetLastError());
}

void CholeskyGpu::solve(const double* /*d_b: consumed by factor()*/, double* d_x, cudaStream_t s, LaunchCounter* lc,
                        EventProfiler* prof) {
  if (S_.d == 3) solve_t<3>(d_x, s, lc, prof);
  else solve_t<6>(d_x, s, lc, prof);
}

}  // namespace g2o_b200

#ifdef CHOL_TIMING
extern "C" void b200_debug_chol_stamps(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g2o_b200::g_chol_stamp, sizeof(unsigned long long) * 6 * 4096);
  if (reset) {
    static unsigned long long z[6 * 4096];
    for (int i = 0; i < 6 * 4096; ++i) z[i] = i / 4096 == 3 ? ~0ull : 0ull;
    cudaMemcpyToSymbol(g2o_b200::g_chol_stamp, z, sizeof(z));
  }
}
extern "C" void b200_debug_chain_timing(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g2o_b200::g_chain_timing, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g2o_b200::g_chain_timing, z, sizeof(z)); }
}
extern "C" void b200_debug_chol_timing(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g2o_b200::g_chol_timing, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g2o_b200::g_chol_timing, z, sizeof(z)); }
}
#endif
