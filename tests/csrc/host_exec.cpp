// host_exec.cpp - TEST-ONLY sequential executor of the supernodal schedule produced by
// openslam_g2o_b200/csrc/symbolic.cpp.  It validates the integer plan (row structures, update lists,
// relative indices, scatter plan, task/level order) on the CPU where there is no GPU; it is never part
// of libg2o_b200.so and the product never calls it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "block_amd.h"
#include "symbolic.h"
#include "spinv_lookup.h"

using namespace g2o_b200;

static int g_nd_levels = 0;
static double g_relax_frac = 0.25;
static int g_chain = 1, g_chain_min_links = 3, g_chain_max_rows = 31, g_wide = -1;
extern "C" void hx_set_wide(int w) { g_wide = w; }
extern "C" void hx_set_chain(int on, int min_links, int max_rows) { g_chain = on; g_chain_min_links = min_links; g_chain_max_rows = max_rows; }
static void apply_chain_options(SymbolicOptions& o) { o.chain = g_chain != 0; o.chain_min_links = g_chain_min_links; o.chain_max_rows = g_chain_max_rows; o.wide_tiles = g_wide; }
extern "C" void hx_set_nd_levels(int k) { g_nd_levels = k; }
extern "C" void hx_set_relax_frac(double f) { g_relax_frac = f; }
extern "C" int hx_block_amd(int n, const int* cp, const int* ri, int* perm) {
  auto p = block_amd(n, cp, ri);
  memcpy(perm, p.data(), n * sizeof(int));
  return 0;
}

// info[0]=nsn info[1]=ntasks info[2]=nlevels info[3]=scalar_lnz info[4]=factor_doubles info[5]=max_nrow info[6]=max_ncol
extern "C" int hx_analyze(int nb, int d, const int* cp, const int* ri, int max_cols, int relax, long long* info, int* perm_out) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0; o.nd_levels = g_nd_levels; o.nd_min_part = 4; o.relax_frac = g_relax_frac;
  apply_chain_options(o);
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  info[0] = S.nsn; info[1] = (long long)S.task_ptr.size() - 1; info[2] = S.nlevels; info[3] = S.scalar_lnz;
  info[4] = S.factor_doubles; info[5] = S.max_nrow; info[6] = S.max_ncol; info[7] = (long long)S.flops;
  info[8] = (long long)S.chain_sn.size();
  if (perm_out) memcpy(perm_out, S.perm.data(), nb * sizeof(int));
  return 0;
}

// numeric factorisation exactly along the plan (tiles / work items / chunks / tail chain); L11 stays in the panel
static int hx_factor_impl(const SymbolicFactor& S, int nb, int d, const int* cp, const double* vals, double lambda,
                          std::vector<double>& L) {
  L.assign(S.factor_doubles, 0.0);
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
    const int TB = S.tile_blocks, TBC = S.tile_blocks_c;
    for (int tile = S.sn_tile_ptr[J]; tile < S.sn_tile_ptr[J + 1]; ++tile) {
      if (S.tile_sn[tile] != J) return -5;
      const int R0 = S.tile_r0[tile], C0 = S.tile_c0[tile];
      std::vector<double> acc((size_t)TB * d * TBC * d, 0.0);
      const int TS = TB * d, TSC = TBC * d;
      for (int wi = S.tile_work_ptr[tile]; wi < S.tile_work_ptr[tile + 1]; ++wi) {
        const int u = S.work_u[wi];
        if (u < S.upd_ptr[J] || u >= S.upd_ptr[J + 1]) return -6;
        const int K = S.upd_k[u], p0 = S.upd_p0[u];
        if (S.sn_on_chain[K]) return -9;  // chain links never feed GROUP work items
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
            if (tr < 0 || tr + d > TS || tc < 0 || tc + d > TSC) return -7;
            for (int cc = 0; cc < d; ++cc)
              for (int rr = 0; rr < d; ++rr) {
                double s = 0;
                for (int k = 0; k < Nk; ++k) s += Kp[(p0 + a) * d + rr + (size_t)k * Mk] * Kp[(p0 + b) * d + cc + (size_t)k * Mk];
                acc[tr + rr + (size_t)(tc + cc) * TS] += s;
              }
          }
      }
      const int rows = std::min(TS, M - R0 * d), cols = std::min(TSC, N - C0 * d);
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
    if (S.sn_on_chain[J]) return 0;  // the panel now holds A + the updates of the supernodes below the chain
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
  // the tail chain, the way chol_chain_kernel runs it: ONE frontal matrix in logical (local row) order that moves
  // up the chain - link j adds its panel, eliminates its columns right-looking, and hands the remaining update
  // matrix to link j+1 through chain_map; rows flagged in chain_new_rows start from zero
  if (!S.chain_sn.empty()) {
    const int nl = (int)S.chain_sn.size();
    const int RM = 32 * d;
    std::vector<double> F((size_t)RM * RM, 0.0), G((size_t)RM * RM, 0.0);
    for (int j = 0; j < nl; ++j) {
      const int J = S.chain_sn[j];
      if (j + 1 < nl && S.sn_parent[J] != S.chain_sn[j + 1]) return -100;
      if (j + 1 == nl && S.sn_parent[J] != -1) return -101;
      const int nrow = S.sn_nrow[J], ncol = S.sn_ncol[J], M = nrow * d, N = ncol * d;
      if (nrow > 31) return -102;
      double* P = L.data() + S.sn_lptr[J];
      for (int a = 0; a < nrow; ++a)
        if (S.chain_new_rows[j] >> a & 1u)
          for (int b = 0; b < nrow; ++b)
            for (int i = 0; i < d; ++i) for (int jj = 0; jj < d; ++jj) { F[(a * d + i) + (size_t)(b * d + jj) * RM] = 0; F[(b * d + i) + (size_t)(a * d + jj) * RM] = 0; }
      for (int c = 0; c < N; ++c) for (int r = c; r < M; ++r) F[r + (size_t)c * RM] += P[r + (size_t)c * M];
      for (int c = 0; c < N; ++c) {
        double dj = F[c + (size_t)c * RM];
        if (!(dj > 0)) return 1;
        dj = std::sqrt(dj);
        F[c + (size_t)c * RM] = dj;
        for (int r = c + 1; r < M; ++r) F[r + (size_t)c * RM] /= dj;
        for (int c2 = c + 1; c2 < M; ++c2) {
          const double f = F[c2 + (size_t)c * RM];
          for (int r = c2; r < M; ++r) F[r + (size_t)c2 * RM] -= F[r + (size_t)c * RM] * f;
        }
      }
      for (int c = 0; c < N; ++c) for (int r = c; r < M; ++r) P[r + (size_t)c * M] = F[r + (size_t)c * RM];
      done[J] = 1;
      if (j + 1 < nl) {
        const int nbelow = nrow - ncol;
        const int* map = S.chain_map.data() + S.chain_mapptr[j + 1];
        if (S.chain_mapptr[j + 2] - S.chain_mapptr[j + 1] != nbelow) return -103;
        const int Jn = S.chain_sn[j + 1];
        const int* jr = S.sn_rows.data() + S.sn_rowptr[J];
        const int* nr = S.sn_rows.data() + S.sn_rowptr[Jn];
        std::fill(G.begin(), G.end(), 0.0);
        for (int a = 0; a < nbelow; ++a) {
          if (map[a] < 0 || map[a] >= S.sn_nrow[Jn] || nr[map[a]] != jr[ncol + a]) return -104;
          if (S.chain_new_rows[j + 1] >> map[a] & 1u) return -105;
          if (a > 0 && map[a] <= map[a - 1]) return -106;
          for (int b = 0; b <= a; ++b)
            for (int i = 0; i < d; ++i) for (int jj = 0; jj < d; ++jj)
              G[(map[a] * d + i) + (size_t)(map[b] * d + jj) * RM] = F[((ncol + a) * d + i) + (size_t)((ncol + b) * d + jj) * RM];
        }
        // every row of the next link is either inherited or flagged new
        unsigned inh = 0;
        for (int a = 0; a < nbelow; ++a) inh |= 1u << map[a];
        const unsigned all = (1u << S.sn_nrow[Jn]) - 1u;
        if ((inh | S.chain_new_rows[j + 1]) != all || (inh & S.chain_new_rows[j + 1])) return -107;
        F.swap(G);
      }
    }
  }
  for (int s = 0; s < S.nsn; ++s) if (!done[s]) return -4;
  (void)nt;
  return 0;
}

extern "C" int hx_solve(int nb, int d, const int* cp, const int* ri, const double* vals, double lambda,
                        const double* b, double* x, int max_cols, int relax) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0; o.nd_levels = g_nd_levels; o.nd_min_part = 4; o.relax_frac = g_relax_frac;
  apply_chain_options(o);
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  std::vector<double> L;
  if (int rc = hx_factor_impl(S, nb, d, cp, vals, lambda, L)) return rc;
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

// Sparse inverse subset along the plan of sparse_inverse.cuh (same recursion, same block lookup: spinv_lookup.h), top of
// the supernodal tree first.  out: the requested blocks (ORIGINAL block indices) of (A + lambda I)^-1, d*d column-major;
// found[q] = 0 when block q is outside the pattern of the factor (out left untouched).
extern "C" int hx_sparse_inverse(int nb, int d, const int* cp, const int* ri, const double* vals, double lambda, int nreq,
                                 const int* rows, const int* cols, double* out, int* found, int max_cols, int relax) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0; o.nd_levels = g_nd_levels; o.nd_min_part = 4; o.relax_frac = g_relax_frac;
  apply_chain_options(o);
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  std::vector<double> L;
  if (int rc = hx_factor_impl(S, nb, d, cp, vals, lambda, L)) return rc;
  std::vector<double> Z(S.factor_doubles, 0.0);
  const long long* lptr = reinterpret_cast<const long long*>(S.sn_lptr.data());
  auto locate = [&](int gp, int gq, int* ld) {
    return spinv_locate(gp, gq, d, S.col2sn.data(), S.sn_col0.data(), S.sn_ncol.data(), S.sn_nrow.data(), S.sn_rowptr.data(), S.sn_rows.data(), lptr, ld);
  };
  for (int J = S.nsn - 1; J >= 0; --J) {  // parents have larger indices: every ancestor is final
    if (S.sn_parent[J] >= 0 && S.sn_parent[J] <= J) return -200;
    const int nr = S.sn_nrow[J], nc = S.sn_ncol[J], M = nr * d, N = nc * d, B = M - N;
    const double* P = L.data() + S.sn_lptr[J];
    double* Zp = Z.data() + S.sn_lptr[J];
    const int* jrows = S.sn_rows.data() + S.sn_rowptr[J];
    std::vector<double> Di((size_t)N * N, 0.0), Y((size_t)B * N, 0.0);  // Di = L11^-1 (lower), Y = L21 Di
    for (int j = 0; j < N; ++j)
      for (int i = j; i < N; ++i) {
        double s = i == j ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) s -= P[i + (size_t)k * M] * Di[k + (size_t)j * N];
        Di[i + (size_t)j * N] = s / P[i + (size_t)i * M];
      }
    for (int c = 0; c < N; ++c)
      for (int r = 0; r < B; ++r) {
        double s = 0;
        for (int k = c; k < N; ++k) s += P[N + r + (size_t)k * M] * Di[k + (size_t)c * N];
        Y[r + (size_t)c * B] = s;
      }
    for (int p = nc; p < nr; ++p)
      for (int q = nc; q < nr; ++q) {
        const int gp = jrows[p], gq = jrows[q];
        int ld;
        const long long off = gp >= gq ? locate(gp, gq, &ld) : locate(gq, gp, &ld);
        if (off < 0) return -201;  // the pattern of L is closed under this recursion
        for (int i = 0; i < d; ++i)
          for (int j = 0; j < d; ++j) {
            const double g = gp >= gq ? Z[off + i + (size_t)j * ld] : Z[off + j + (size_t)i * ld];
            for (int c = 0; c < N; ++c) Zp[p * d + i + (size_t)c * M] -= g * Y[(q - nc) * d + j + (size_t)c * B];
          }
      }
    for (int b = 0; b < N; ++b)
      for (int a = 0; a < N; ++a) {
        double s = 0;
        for (int k = std::max(a, b); k < N; ++k) s += Di[k + (size_t)a * N] * Di[k + (size_t)b * N];
        for (int r = 0; r < B; ++r) s -= Y[r + (size_t)a * B] * Zp[N + r + (size_t)b * M];
        Zp[a + (size_t)b * M] = s;
      }
  }
  for (int q = 0; q < nreq; ++q) {
    int pr = S.pinv[rows[q]], pc = S.pinv[cols[q]];
    const bool tr = pr < pc;
    if (tr) std::swap(pr, pc);
    int ld;
    const long long off = locate(pr, pc, &ld);
    found[q] = off >= 0;
    if (off < 0) continue;
    for (int j = 0; j < d; ++j)
      for (int i = 0; i < d; ++i) out[(size_t)q * d * d + i + j * d] = tr ? Z[off + j + (size_t)i * ld] : Z[off + i + (size_t)j * ld];
  }
  return 0;
}

// Dataflow schedule check: walking flow_kind/flow_arg in list order, everything a task waits for must already be
// complete (dependencies only point backwards - the no-deadlock argument of chol.cu), every group / chunk must appear
// exactly once, every split tile must be completed by exactly one last group, and the completion targets (sn_nupd, sn_nchunk) must be reached exactly.
extern "C" int hx_check_flow(int nb, int d, const int* cp, const int* ri, int max_cols, int relax, int group_items) {
  SymbolicOptions o; o.max_panel_cols_scalar = max_cols; o.relax = relax != 0; o.nd_levels = g_nd_levels; o.nd_min_part = 4; o.relax_frac = g_relax_frac; o.group_items = group_items;
  apply_chain_options(o);
  SymbolicFactor S = analyze(nb, d, cp, ri, o);
  const int nt = (int)S.task_ptr.size() - 1;
  std::vector<int> upd(S.nsn, 0), chunk(S.nsn, 0), slot(S.rtile_tile.size(), 0);
  std::vector<char> seen_g(S.group_tile.size(), 0), seen_r(S.rtile_tile.size(), 0), seen_c(S.chunk_sn.size(), 0), seen_t(nt, 0);
  auto ready = [&](int K) { return chunk[K] == S.sn_nchunk[K]; };
  auto items_ready = [&](int w0, int w1) {
    for (int wi = w0; wi < w1; ++wi) if (S.sn_on_chain[S.work_ksn[wi]] || !ready(S.work_ksn[wi])) return false;
    return true;
  };
  for (size_t i = 0; i < S.flow_kind.size(); ++i) {
    const int a = S.flow_arg[i];
    switch (S.flow_kind[i]) {
      case 0: {
        if (a < 0 || a >= nt || seen_t[a]) return -10;
        seen_t[a] = 1;
        for (int q = S.task_ptr[a]; q < S.task_ptr[a + 1]; ++q) {
          const int J = S.task_sn[q];
          if (S.sn_nupd[J] != 0) return -11;  // subtree supernodes are updated by their own CTA
          for (int t = S.sn_tile_ptr[J]; t < S.sn_tile_ptr[J + 1]; ++t)
            if (!items_ready(S.tile_work_ptr[t], S.tile_work_ptr[t + 1])) return -12;
          for (int c = S.sn_chunk_ptr[J]; c < S.sn_chunk_ptr[J + 1]; ++c) { if (seen_c[c]) return -13; seen_c[c] = 1; chunk[J]++; }
        }
        break;
      }
      case 1: {
        if (a < 0 || a >= (int)seen_g.size() || seen_g[a]) return -20;
        seen_g[a] = 1;
        if (!items_ready(S.group_w0[a], S.group_w1[a])) return -21;
        const int r = S.group_rtile[a];
        if ((r < 0) != (S.group_slot[a] < 0)) return -22;
        if (r < 0) upd[S.tile_sn[S.group_tile[a]]]++;
        else {
          if (S.rtile_tile[r] != S.group_tile[a]) return -23;
          if (++slot[r] == S.rtile_nslots[r]) {  // the last group of a split tile reduces and subtracts
            if (seen_r[r]) return -30;
            seen_r[r] = 1;
            upd[S.tile_sn[S.rtile_tile[r]]]++;
          }
        }
        break;
      }
      case 3: {
        if (a < 0 || a >= (int)seen_c.size() || seen_c[a]) return -40;
        seen_c[a] = 1;
        const int J = S.chunk_sn[a];
        if (S.sn_on_chain[J]) return -42;
        if (upd[J] != S.sn_nupd[J]) return -41;
        chunk[J]++;
        break;
      }
      default: return -50;
    }
  }
  for (int J = 0; J < S.nsn; ++J) {
    if (S.sn_on_chain[J]) { if (chunk[J] != 0 || upd[J] != S.sn_nupd[J]) return -64; continue; }
    if (!ready(J) || upd[J] != S.sn_nupd[J]) return -60;
  }
  // forward-substitution lists of the chain = the general lists restricted to sources below the chain, same order
  if (!S.chain_sn.empty()) {
    auto source_of = [&](int entry) { return (int)(std::upper_bound(S.sn_cptr.begin(), S.sn_cptr.end(), (int64_t)entry) - S.sn_cptr.begin()) - 1; };
    for (size_t j = 0; j < S.chain_sn.size(); ++j) {
      const int J = S.chain_sn[j];
      for (int c = 0; c < S.sn_ncol[J] * d; ++c) {
        const int g = S.sn_col0[J] * d + c, cc = S.chain_colptr[j] + c;
        int q = S.chain_fwd_ptr[cc];
        for (int e = S.fwd_ptr[g]; e < S.fwd_ptr[g + 1]; ++e) {
          if (S.sn_on_chain[source_of(S.fwd_src[e])]) continue;
          if (q >= S.chain_fwd_ptr[cc + 1] || S.chain_fwd_src[q] != S.fwd_src[e]) return -65;
          ++q;
        }
        if (q != S.chain_fwd_ptr[cc + 1]) return -66;
      }
    }
  }
  for (char c : seen_g) if (!c) return -61;
  for (char c : seen_r) if (!c) return -62;
  for (size_t c = 0; c < seen_c.size(); ++c) if (!seen_c[c] && !S.sn_on_chain[S.chunk_sn[c]]) return -63;
  // scratch slots must be unique across the whole factorisation (levels overlap in the dataflow kernel)
  {
    std::vector<char> used(std::max(S.max_group_slots, 1), 0);
    for (size_t g = 0; g < S.group_slot.size(); ++g)
      if (S.group_slot[g] >= 0) { if (S.group_slot[g] >= S.max_group_slots || used[S.group_slot[g]]) return -70; used[S.group_slot[g]] = 1; }
  }
  // backward sweep: the parent task of a task must sit later in the (level-sorted) task list
  for (int t = 0; t < nt; ++t) if (S.task_parent[t] >= 0 && S.task_parent[t] <= t) return -80;
  // panel geometry the kernels rely on
  for (int J = 0; J < S.nsn; ++J) {
    if (S.sn_ncol[J] > 12) return -90;
    for (int c = S.sn_chunk_ptr[J]; c < S.sn_chunk_ptr[J + 1]; ++c) if (S.sn_ncol[J] + S.chunk_nb[c] + 1 > 32) return -91;
  }
  return 0;
}
