// host_exec.cpp - TEST-ONLY sequential executor of the supernodal schedule produced by
// openslam_g2o_b200/csrc/symbolic.cpp.  It validates the integer plan (row structures, update lists,
// relative indices, scatter plan, task/level order) on the CPU where there is no GPU; it is never part
// of libg2o_b200.so and the product never calls it.
#include <cmath>
#include <cstring>
#include <vector>

#include "block_amd.h"
#include "symbolic.h"

using namespace g2o_b200;

extern "C" int hx_block_amd(int n, const int* cp, const int* ri, int* perm) {
  auto p = block_amd(n, cp, ri);
  memcpy(perm, p.data(), n * sizeof(int));
  return 0;
}

// info[0]=nsn info[1]=ntasks info[2]=nlevels info[3]=scalar_lnz info[4]=factor_doubles info[5]=max_nrow info[6]=max_ncol
extern "C" int hx_analyze(int nb, int d, const int* cp, const int* ri, int max_cols, int relax, long long* info, int* perm_out) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0;
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  info[0] = S.nsn; info[1] = (long long)S.task_ptr.size() - 1; info[2] = S.nlevels; info[3] = S.scalar_lnz;
  info[4] = S.factor_doubles; info[5] = S.max_nrow; info[6] = S.max_ncol; info[7] = (long long)S.flops;
  if (perm_out) memcpy(perm_out, S.perm.data(), nb * sizeof(int));
  return 0;
}

extern "C" int hx_solve(int nb, int d, const int* cp, const int* ri, const double* vals, double lambda,
                        const double* b, double* x, int max_cols, int relax) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0;
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  std::vector<double> L(S.factor_doubles, 0.0);
  const int nblk = cp[nb];
  for (int k = 0; k < nblk; ++k)
    for (int c = 0; c < d; ++c)
      for (int r = 0; r < d; ++r) {
        double v = S.a_trans[k] ? vals[(size_t)k * d * d + c + r * d] : vals[(size_t)k * d * d + r + c * d];
        L[S.a_dst[k] + r + (int64_t)c * S.a_ld[k]] = v;
      }
  for (int k = 0; k < nb; ++k) for (int r = 0; r < d; ++r) L[S.diag_dst[k] + r + (int64_t)r * S.diag_ld[k]] += lambda;
  std::vector<char> done(S.nsn, 0);
  const int nt = (int)S.task_ptr.size() - 1;
  auto factor_sn = [&](int J) -> int {
    double* P = L.data() + S.sn_lptr[J];
    const int M = S.sn_nrow[J] * d, N = S.sn_ncol[J] * d;
    const int* jrows = S.sn_rows.data() + S.sn_rowptr[J];
    // the GPU's update plan: destination tiles, each with its ordered list of work items
    const int TB = S.tile_blocks;
    for (int tile = S.sn_tile_ptr[J]; tile < S.sn_tile_ptr[J + 1]; ++tile) {
      if (S.tile_sn[tile] != J) return -5;
      const int R0 = S.tile_r0[tile], C0 = S.tile_c0[tile];
      std::vector<double> acc((size_t)TB * d * TB * d, 0.0);
      const int TS = TB * d;
      for (int wi = S.tile_work_ptr[tile]; wi < S.tile_work_ptr[tile + 1]; ++wi) {
        const int u = S.work_u[wi];
        if (u < S.upd_ptr[J] || u >= S.upd_ptr[J + 1]) return -6;
        const int K = S.upd_k[u], p0 = S.upd_p0[u];
        if (!done[K]) return -2;  // schedule violation
        const double* Kp = L.data() + S.sn_lptr[K];
        const int Mk = S.sn_nrow[K] * d, Nk = S.sn_ncol[K] * d;
        const int* krows = S.sn_rows.data() + S.sn_rowptr[K];
        const int* rel = S.rel.data() + S.upd_relptr[u];
        for (int b = S.work_b0[wi]; b < S.work_b1[wi]; ++b)
          for (int a = S.work_a0[wi]; a < S.work_a1[wi]; ++a) {
            if (a < b) continue;
            if (jrows[rel[a]] != krows[p0 + a] || jrows[rel[b]] != krows[p0 + b]) return -3;
            const int tr = (rel[a] - R0) * d, tc = (rel[b] - C0) * d;
            if (tr < 0 || tr + d > TS || tc < 0 || tc + d > TS) return -7;
            for (int cc = 0; cc < d; ++cc)
              for (int rr = 0; rr < d; ++rr) {
                double s = 0;
                for (int k = 0; k < Nk; ++k) s += Kp[(p0 + a) * d + rr + (size_t)k * Mk] * Kp[(p0 + b) * d + cc + (size_t)k * Mk];
                acc[tr + rr + (size_t)(tc + cc) * TS] += s;
              }
          }
      }
      const int rows = std::min(TS, M - R0 * d), cols = std::min(TS, N - C0 * d);
      for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) P[R0 * d + r + (size_t)(C0 * d + c) * M] -= acc[r + (size_t)c * TS];
    }
    // the chunk plan must cover every block row below the diagonal block exactly once
    {
      int expect = S.sn_ncol[J];
      for (int ch = S.sn_chunk_ptr[J]; ch < S.sn_chunk_ptr[J + 1]; ++ch) {
        if (S.chunk_sn[ch] != J || S.chunk_b0[ch] != expect || S.chunk_nb[ch] > S.chunk_blocks) return -8;
        expect += S.chunk_nb[ch];
      }
      if (expect != S.sn_nrow[J]) return -8;
    }
    for (int j = 0; j < N; ++j) {
      double dj = P[j + (size_t)j * M];
      if (!(dj > 0)) return 1;
      dj = std::sqrt(dj);
      P[j + (size_t)j * M] = dj;
      for (int i = j + 1; i < M; ++i) P[i + (size_t)j * M] /= dj;
      for (int c = j + 1; c < N; ++c) {
        double f = P[c + (size_t)j * M];
        for (int i = c; i < M; ++i) P[i + (size_t)c * M] -= P[i + (size_t)j * M] * f;
      }
    }
    done[J] = 1;
    return 0;
  };
  for (int l = 0; l < S.nlevels; ++l)
    for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
      for (int q = S.task_ptr[t]; q < S.task_ptr[t + 1]; ++q) {
        int rc = factor_sn(S.task_sn[q]);
        if (rc) return rc;
      }
  for (int s = 0; s < S.nsn; ++s) if (!done[s]) return -4;
  (void)nt;
  // solve: y = P b ; forward (pull) ; backward ; x = P^T y
  const int n = nb * d;
  std::vector<double> y(n);
  for (int k = 0; k < nb; ++k) for (int r = 0; r < d; ++r) y[k * d + r] = b[S.perm[k] * d + r];
  for (int l = 0; l < S.nlevels; ++l)
    for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
      for (int q = S.task_ptr[t]; q < S.task_ptr[t + 1]; ++q) {
        int J = S.task_sn[q];
        const int M = S.sn_nrow[J] * d, N = S.sn_ncol[J] * d;
        double* yj = y.data() + S.sn_col0[J] * d;
        for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
          int K = S.upd_k[u];
          const double* Kp = L.data() + S.sn_lptr[K];
          const int Mk = S.sn_nrow[K] * d, Nk = S.sn_ncol[K] * d;
          const int* krows = S.sn_rows.data() + S.sn_rowptr[K];
          const double* yk = y.data() + S.sn_col0[K] * d;
          for (int p = S.upd_p0[u]; p < S.upd_p1[u]; ++p)
            for (int rr = 0; rr < d; ++rr) {
              double s = 0;
              for (int k = 0; k < Nk; ++k) s += Kp[p * d + rr + (size_t)k * Mk] * yk[k];
              yj[(krows[p] - S.sn_col0[J]) * d + rr] -= s;
            }
        }
        const double* P = L.data() + S.sn_lptr[J];
        for (int j = 0; j < N; ++j) {
          yj[j] /= P[j + (size_t)j * M];
          for (int i = j + 1; i < N; ++i) yj[i] -= P[i + (size_t)j * M] * yj[j];
        }
      }
  for (int l = S.nlevels - 1; l >= 0; --l)
    for (int t = S.level_ptr[l + 1] - 1; t >= S.level_ptr[l]; --t)
      for (int q = S.task_ptr[t + 1] - 1; q >= S.task_ptr[t]; --q) {
        int J = S.task_sn[q];
        const int M = S.sn_nrow[J] * d, N = S.sn_ncol[J] * d;
        const double* P = L.data() + S.sn_lptr[J];
        const int* jrows = S.sn_rows.data() + S.sn_rowptr[J];
        double* xj = y.data() + S.sn_col0[J] * d;
        for (int j = N - 1; j >= 0; --j) {
          double s = xj[j];
          for (int i = N; i < M; ++i) s -= P[i + (size_t)j * M] * y[jrows[i / d] * d + i % d];
          for (int i = j + 1; i < N; ++i) s -= P[i + (size_t)j * M] * xj[i];
          xj[j] = s / P[j + (size_t)j * M];
        }
      }
  for (int k = 0; k < nb; ++k) for (int r = 0; r < d; ++r) x[S.perm[k] * d + r] = y[k * d + r];
  return 0;
}
