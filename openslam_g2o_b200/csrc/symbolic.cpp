// symbolic.cpp - see symbolic.h.  All of this is integer-only host work done once per structure.
#include "symbolic.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <numeric>

#include "block_amd.h"
#include "host_parallel.h"

namespace g2o_b200 {

SymbolicFactor analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt) {
  SymbolicFactor S;
  S.nb = nb;
  S.d = d;
  const int nblk = colptr[nb];

  // ---- 1. ordering (bit-exact twin of cs_amd on the block pattern)
  if (!opt.given_perm.empty()) S.perm = opt.given_perm;
  else if (opt.nd_levels > 0) S.perm = nested_dissection_order(nb, colptr, rowidx, opt.nd_levels, opt.nd_min_part);
  else S.perm = block_amd(nb, colptr, rowidx);
  S.pinv.assign(nb, 0);
  for (int k = 0; k < nb; ++k) S.pinv[S.perm[k]] = k;
  if (opt.nd_levels > 0 && opt.given_perm.empty()) {
    // the supernode detection below wants a post-ordered elimination tree (AMD delivers one; a concatenation of
    // independently ordered parts need not): post-order the tree of the permuted pattern and compose
    std::vector<int> parent(nb, -1), anc(nb, -1);
    std::vector<std::vector<int>> upper(nb);  // for column k (new index): rows i < k
    for (int c = 0; c < nb; ++c)
      for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
        const int r = rowidx[p];
        if (r == c) continue;
        const int i = S.pinv[r], j = S.pinv[c];
        upper[std::max(i, j)].push_back(std::min(i, j));
      }
    for (int k = 0; k < nb; ++k)
      for (int i0 : upper[k]) {
        int i = i0;
        while (i != -1 && i < k) {
          const int nxt = anc[i];
          anc[i] = k;
          if (nxt == -1) parent[i] = k;
          i = nxt;
        }
      }
    std::vector<int> head(nb, -1), next(nb, -1), post, stack;
    for (int j = nb - 1; j >= 0; --j)
      if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    post.reserve(nb);
    for (int root = 0; root < nb; ++root) {
      if (parent[root] >= 0) continue;
      stack.push_back(root);
      while (!stack.empty()) {
        const int v = stack.back();
        const int ch = head[v];
        if (ch >= 0) { head[v] = next[ch]; stack.push_back(ch); }
        else { post.push_back(v); stack.pop_back(); }
      }
    }
    std::vector<int> perm2(nb);
    for (int k = 0; k < nb; ++k) perm2[k] = S.perm[post[k]];
    S.perm.swap(perm2);
    for (int k = 0; k < nb; ++k) S.pinv[S.perm[k]] = k;
  }

  // ---- 2. permuted pattern, as "upper by column": for column k the rows i < k
  std::vector<int> up_ptr(nb + 1, 0);
  for (int c = 0; c < nb; ++c)
    for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
      int r = rowidx[p];
      if (r == c) continue;
      int i = S.pinv[r], j = S.pinv[c];
      up_ptr[std::max(i, j) + 1]++;
    }
  for (int k = 0; k < nb; ++k) up_ptr[k + 1] += up_ptr[k];
  std::vector<int> up_idx(up_ptr[nb]);
  {
    std::vector<int> fill(up_ptr.begin(), up_ptr.end() - 1);
    for (int c = 0; c < nb; ++c)
      for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
        int r = rowidx[p];
        if (r == c) continue;
        int i = S.pinv[r], j = S.pinv[c];
        up_idx[fill[std::max(i, j)]++] = std::min(i, j);
      }
  }

  // ---- 3. elimination tree (Liu, path compression) and exact column counts (row-subtree walks)
  S.parent.assign(nb, -1);
  {
    std::vector<int> anc(nb, -1);
    for (int k = 0; k < nb; ++k)
      for (int p = up_ptr[k]; p < up_ptr[k + 1]; ++p) {
        int i = up_idx[p];
        while (i != -1 && i < k) {
          int nxt = anc[i];
          anc[i] = k;
          if (nxt == -1) S.parent[i] = k;
          i = nxt;
        }
      }
  }
  S.colcount.assign(nb, 1);
  {
    std::vector<int> mark(nb, -1);
    for (int k = 0; k < nb; ++k) {
      mark[k] = k;
      for (int p = up_ptr[k]; p < up_ptr[k + 1]; ++p)
        for (int j = up_idx[p]; mark[j] != k; j = S.parent[j]) {
          S.colcount[j]++;
          mark[j] = k;
        }
    }
  }
  S.scalar_lnz = 0;
  for (int k = 0; k < nb; ++k)
    S.scalar_lnz += (int64_t)d * d * (S.colcount[k] - 1) + (int64_t)d * (d + 1) / 2;

  // ---- 4. supernodes: fundamental partition
  std::vector<int> first;  // first column of each supernode
  for (int j = 0; j < nb; ++j) {
    bool join = j > 0 && S.parent[j - 1] == j && S.colcount[j - 1] == S.colcount[j] + 1;
    if (!join) first.push_back(j);
  }
  int ns = (int)first.size();
  std::vector<int> nc(ns), nr(ns);
  for (int s = 0; s < ns; ++s) {
    int end = (s + 1 < ns) ? first[s + 1] : nb;
    nc[s] = end - first[s];
    nr[s] = S.colcount[first[s]];
  }
  // relaxed amalgamation of a supernode with its parent when they are adjacent in column order
  if (opt.relax && ns > 1) {
    std::vector<int> c2s(nb);
    for (int s = 0; s < ns; ++s) for (int j = first[s]; j < first[s] + nc[s]; ++j) c2s[j] = s;
    std::vector<int> merged_into(ns, -1);
    std::vector<double> zeros(ns, 0.0);
    for (int s = 0; s + 1 < ns; ++s) {
      int last = first[s] + nc[s] - 1;
      int pj = S.parent[last];
      if (pj < 0) continue;
      int p = c2s[pj];
      if (p != s + 1 || first[p] != last + 1) continue;  // parent must start right after s
      double add = (double)nc[s] * (nc[s] + nr[p] - nr[s]);
      double z = zeros[s] + zeros[p] + add;
      int mc = nc[s] + nc[p];
      double total = (double)mc * (nc[s] + nr[p]) - 0.5 * (double)mc * (mc - 1);
      double frac = z / std::max(total, 1.0);
      // amalgamation stops at the panel width: a merged supernode wider than one panel would be cut into equal
      // pieces again (16 block columns -> 8 + 8), i.e. more and narrower panels on the critical path than 12 + ...
      const int cap = std::max(1, std::min(opt.max_panel_cols_scalar / d, 12)) * d;
      bool ok = (mc * d <= 12) || (mc * d <= 48 && mc * d <= cap && frac < 0.6) || (mc * d <= cap && frac < opt.relax_frac) || frac < 0.05;
      if (!ok) continue;
      // merge s into p (p keeps its index; its first column moves down)
      first[p] = first[s];
      nr[p] = nc[s] + nr[p];
      nc[p] = mc;
      zeros[p] = z;
      merged_into[s] = p;
    }
    std::vector<int> f2, c2, r2;
    for (int s = 0; s < ns; ++s)
      if (merged_into[s] < 0) { f2.push_back(first[s]); c2.push_back(nc[s]); r2.push_back(nr[s]); }
    first.swap(f2); nc.swap(c2); nr.swap(r2);
    ns = (int)first.size();
  }
  // width cap: split wide supernodes into a chain of panels
  {
    const int wmax = std::max(1, std::min(opt.max_panel_cols_scalar / d, 12));  // one warp per block column, 12 warps per CTA
    std::vector<int> f2, c2;
    for (int s = 0; s < ns; ++s) {
      int rem = nc[s], c0 = first[s];
      int pieces = (rem + wmax - 1) / wmax;
      int base = rem / pieces, extra = rem % pieces;
      for (int q = 0; q < pieces; ++q) {
        int w = base + (q < extra ? 1 : 0);
        f2.push_back(c0); c2.push_back(w);
        c0 += w;
      }
    }
    first.swap(f2); nc.swap(c2);
    ns = (int)first.size();
  }
  S.nsn = ns;
  S.sn_col0 = first;
  S.sn_ncol = nc;
  S.col2sn.assign(nb, 0);
  for (int s = 0; s < ns; ++s) for (int j = first[s]; j < first[s] + nc[s]; ++j) S.col2sn[j] = s;
  S.sn_parent.assign(ns, -1);
  for (int s = 0; s < ns; ++s) {
    int pj = S.parent[first[s] + nc[s] - 1];
    S.sn_parent[s] = pj < 0 ? -1 : S.col2sn[pj];
  }

  // ---- 5. row structures: own columns, pattern of A below them, children's below-rows
  // lower pattern by column: column c holds rows i > c ("transpose" of up_*)
  std::vector<int> lo_ptr(nb + 1, 0);
  for (int k = 0; k < nb; ++k) for (int p = up_ptr[k]; p < up_ptr[k + 1]; ++p) lo_ptr[up_idx[p] + 1]++;
  for (int k = 0; k < nb; ++k) lo_ptr[k + 1] += lo_ptr[k];
  std::vector<int> lo_idx(lo_ptr[nb]);
  {
    std::vector<int> fill(lo_ptr.begin(), lo_ptr.end() - 1);
    for (int k = 0; k < nb; ++k) for (int p = up_ptr[k]; p < up_ptr[k + 1]; ++p) lo_idx[fill[up_idx[p]]++] = k;
  }
  std::vector<int> child_ptr(ns + 1, 0), child_idx(ns);
  for (int s = 0; s < ns; ++s) if (S.sn_parent[s] >= 0) child_ptr[S.sn_parent[s] + 1]++;
  for (int s = 0; s < ns; ++s) child_ptr[s + 1] += child_ptr[s];
  {
    std::vector<int> fill(child_ptr.begin(), child_ptr.end() - 1);
    for (int s = 0; s < ns; ++s) if (S.sn_parent[s] >= 0) child_idx[fill[S.sn_parent[s]]++] = s;
  }
  S.sn_rowptr.assign(ns + 1, 0);
  S.sn_nrow.assign(ns, 0);
  {
    std::vector<int> stamp(nb, -1);
    std::vector<int> tmp;
    for (int s = 0; s < ns; ++s) {
      tmp.clear();
      const int c0 = first[s], c1 = first[s] + nc[s];
      for (int j = c0; j < c1; ++j) stamp[j] = s;
      for (int j = c0; j < c1; ++j)
        for (int p = lo_ptr[j]; p < lo_ptr[j + 1]; ++p) {
          int i = lo_idx[p];
          if (stamp[i] != s) { stamp[i] = s; tmp.push_back(i); }
        }
      for (int q = child_ptr[s]; q < child_ptr[s + 1]; ++q) {
        int c = child_idx[q];
        for (int p = S.sn_rowptr[c] + S.sn_ncol[c]; p < S.sn_rowptr[c + 1]; ++p) {
          int i = S.sn_rows[p];
          if (stamp[i] != s) { stamp[i] = s; tmp.push_back(i); }
        }
      }
      std::sort(tmp.begin(), tmp.end());
      for (int j = c0; j < c1; ++j) S.sn_rows.push_back(j);
      S.sn_rows.insert(S.sn_rows.end(), tmp.begin(), tmp.end());
      S.sn_nrow[s] = nc[s] + (int)tmp.size();
      S.sn_rowptr[s + 1] = (int)S.sn_rows.size();
    }
  }
  S.sn_lptr.assign(ns + 1, 0);
  S.flops = 0;
  for (int s = 0; s < ns; ++s) {
    S.sn_lptr[s + 1] = S.sn_lptr[s] + (int64_t)S.sn_nrow[s] * d * S.sn_ncol[s] * d;
    S.max_nrow = std::max(S.max_nrow, S.sn_nrow[s]);
    S.max_ncol = std::max(S.max_ncol, S.sn_ncol[s]);
  }
  S.factor_doubles = S.sn_lptr[ns];

  // ---- 6. scatter plan for the input blocks
  S.a_dst.assign(nblk, 0); S.a_ld.assign(nblk, 0); S.a_trans.assign(nblk, 0);
  S.diag_dst.assign(nb, 0); S.diag_ld.assign(nb, 0);
  {
    // position of a block row inside a supernode's row list: binary search
    auto local_row = [&](int s, int row) {
      const int* b = S.sn_rows.data() + S.sn_rowptr[s];
      const int* e = S.sn_rows.data() + S.sn_rowptr[s + 1];
      const int* it = std::lower_bound(b, e, row);
      assert(it != e && *it == row);
      return (int)(it - b);
    };
    for (int c = 0; c < nb; ++c)
      for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
        int r = rowidx[p];
        int i = S.pinv[r], j = S.pinv[c];
        int col = std::min(i, j), row = std::max(i, j);
        int s = S.col2sn[col];
        int ld = S.sn_nrow[s] * d;
        int lr = local_row(s, row), lc = col - S.sn_col0[s];
        S.a_dst[p] = S.sn_lptr[s] + (int64_t)lr * d + (int64_t)lc * d * ld;
        S.a_ld[p] = ld;
        S.a_trans[p] = (i < j) ? 1 : 0;  // A(r,c) lands at C(i,j); the stored lower entry is C(j,i) = A(r,c)^T
      }
    for (int k = 0; k < nb; ++k) {
      int s = S.col2sn[k];
      int ld = S.sn_nrow[s] * d;
      int lc = k - S.sn_col0[s];
      S.diag_dst[k] = S.sn_lptr[s] + (int64_t)lc * d + (int64_t)lc * d * ld;
      S.diag_ld[k] = ld;
    }
  }

  // ---- 7. left-looking update lists with relative indices
  {
    std::vector<int> cnt(ns + 1, 0);
    for (int K = 0; K < ns; ++K) {
      const int* rows = S.sn_rows.data() + S.sn_rowptr[K];
      int prev = -1;
      for (int p = S.sn_ncol[K]; p < S.sn_nrow[K]; ++p) {
        int J = S.col2sn[rows[p]];
        if (J != prev) { cnt[J + 1]++; prev = J; }
      }
    }
    S.upd_ptr.assign(ns + 1, 0);
    for (int s = 0; s < ns; ++s) S.upd_ptr[s + 1] = S.upd_ptr[s] + cnt[s + 1];
    const int nu = S.upd_ptr[ns];
    S.upd_k.assign(nu, 0); S.upd_p0.assign(nu, 0); S.upd_p1.assign(nu, 0); S.upd_relptr.assign(nu + 1, 0);
    std::vector<int> fill(S.upd_ptr.begin(), S.upd_ptr.end() - 1);
    for (int K = 0; K < ns; ++K) {
      const int* rows = S.sn_rows.data() + S.sn_rowptr[K];
      int p = S.sn_ncol[K];
      while (p < S.sn_nrow[K]) {
        int J = S.col2sn[rows[p]];
        int q = p;
        while (q < S.sn_nrow[K] && S.col2sn[rows[q]] == J) ++q;
        int u = fill[J]++;
        S.upd_k[u] = K; S.upd_p0[u] = p; S.upd_p1[u] = q;
        p = q;
      }
    }
    int64_t tot = 0;
    for (int u = 0; u < nu; ++u) { S.upd_relptr[u] = tot; tot += S.sn_nrow[S.upd_k[u]] - S.upd_p0[u]; }
    S.upd_relptr[nu] = tot;
    S.rel.assign(tot, 0);
    std::vector<int> pos(nb, -1);
    for (int J = 0; J < ns; ++J) {
      const int* jr = S.sn_rows.data() + S.sn_rowptr[J];
      for (int q = 0; q < S.sn_nrow[J]; ++q) pos[jr[q]] = q;
      for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
        int K = S.upd_k[u];
        const int* kr = S.sn_rows.data() + S.sn_rowptr[K];
        int64_t o = S.upd_relptr[u];
        for (int p = S.upd_p0[u]; p < S.sn_nrow[K]; ++p) {
          assert(pos[kr[p]] >= 0);
          S.rel[o++] = pos[kr[p]];
        }
      }
      for (int q = 0; q < S.sn_nrow[J]; ++q) pos[jr[q]] = -1;
    }
  }

  // ---- 8. work estimates and the task / level schedule
  std::vector<double> work(ns, 0.0), sub(ns, 0.0);
  const double d3 = (double)d * d * d;
  for (int J = 0; J < ns; ++J) {
    double w = 0;
    for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
      int K = S.upd_k[u];
      w += 2.0 * (S.sn_nrow[K] - S.upd_p0[u]) * (double)(S.upd_p1[u] - S.upd_p0[u]) * S.sn_ncol[K] * d3;
    }
    double ncs = S.sn_ncol[J], nrs = S.sn_nrow[J];
    w += d3 * (ncs * ncs * ncs / 3.0 + (nrs - ncs) * ncs * ncs);
    work[J] = w;
    S.flops += w;
  }
  for (int s = 0; s < ns; ++s) {
    sub[s] += work[s];
    if (S.sn_parent[s] >= 0) sub[S.sn_parent[s]] += sub[s];
  }
  // (the cap: one CTA runs a subtree task sequentially at a few tens of GFLOP/s - beyond ~1 ms per task the CTAs that
  // finish their subtrees early wait for the stragglers; above it the supernodes are split into GROUP / CHUNK tasks)
  const double thresh = std::min(std::max(S.flops * opt.subtree_work_fraction, opt.subtree_min_flops),
                                 std::max(opt.subtree_max_flops, opt.subtree_min_flops));
  S.subtree_flops = 0;
  // task id per supernode: a maximal subtree with sub <= thresh becomes one task (rooted at `root`)
  std::vector<int> root(ns, -1);
  for (int s = ns - 1; s >= 0; --s) {
    int p = S.sn_parent[s];
    if (p >= 0 && root[p] >= 0) root[s] = root[p];                 // inside a subtree task
    else if (sub[s] <= thresh) { root[s] = s; S.subtree_flops += sub[s]; }   // new subtree task rooted here
    else root[s] = -1;                                             // own task, level-scheduled
  }
  std::vector<int> task_of(ns, -1);
  std::vector<std::vector<int>> tasks;
  for (int s = 0; s < ns; ++s) {
    if (root[s] == -1) { task_of[s] = (int)tasks.size(); tasks.push_back({s}); }
    else if (root[s] == s) { task_of[s] = (int)tasks.size(); tasks.push_back({}); }
  }
  for (int s = 0; s < ns; ++s) if (root[s] >= 0) { task_of[s] = task_of[root[s]]; tasks[task_of[s]].push_back(s); }
  const int nt = (int)tasks.size();
  std::vector<int> tlevel(nt, 0);
  for (int s = 0; s < ns; ++s) {  // ascending: children first
    int p = S.sn_parent[s];
    if (p < 0) continue;
    int ts = task_of[s], tp = task_of[p];
    if (ts != tp) tlevel[tp] = std::max(tlevel[tp], tlevel[ts] + 1);
  }
  S.nlevels = 0;
  for (int t = 0; t < nt; ++t) S.nlevels = std::max(S.nlevels, tlevel[t] + 1);
  std::vector<int> order(nt);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tlevel[a] < tlevel[b]; });
  S.level_ptr.assign(S.nlevels + 1, 0);
  for (int t = 0; t < nt; ++t) S.level_ptr[tlevel[t] + 1]++;
  for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
  S.task_ptr.assign(nt + 1, 0);
  for (int q = 0; q < nt; ++q) {
    const std::vector<int>& t = tasks[order[q]];
    S.task_ptr[q + 1] = S.task_ptr[q] + (int)t.size();
    S.task_sn.insert(S.task_sn.end(), t.begin(), t.end());
  }

  // ---- 8b. tail chain: from the root down along the heaviest child, as long as the supernode is a task of its own
  //          (not inside a small-subtree task) and its front fits the chain kernel
  S.sn_on_chain.assign(ns, 0);
  if (opt.chain && d == 6 && ns > 0 && S.sn_parent[ns - 1] < 0) {
    std::vector<int> heavy(ns, -1);
    for (int s = 0; s < ns; ++s) {  // ascending: the last assignment wins only if heavier
      const int p = S.sn_parent[s];
      if (p >= 0 && (heavy[p] < 0 || sub[s] > sub[heavy[p]])) heavy[p] = s;
    }
    std::vector<int> path;
    size_t stage_max = 0, remap_max = 0;
    for (int cur = ns - 1; cur >= 0; cur = heavy[cur]) {
      if (root[cur] != -1 || S.sn_nrow[cur] > opt.chain_max_rows || S.sn_ncol[cur] > 12) break;
      // shared memory of the chain kernel: the staged panel of this link and the re-index buffer of the link below
      // (its rows below the diagonal block, i.e. at most this link's rows)
      const size_t stage = (size_t)S.sn_nrow[cur] * d * S.sn_ncol[cur] * d;
      const size_t below = (size_t)(S.sn_nrow[cur] - S.sn_ncol[cur]);
      const size_t st2 = std::max(stage_max, stage), rm2 = std::max(remap_max, below * (below + 1) / 2);
      if ((st2 + rm2 * (d * d + d)) * sizeof(double) > opt.chain_smem_budget) break;
      stage_max = st2; remap_max = rm2;
      path.push_back(cur);
    }
    if ((int)path.size() >= opt.chain_min_links) {
      S.chain_sn.assign(path.rbegin(), path.rend());
      for (int J : S.chain_sn) S.sn_on_chain[J] = 1;
      // the re-index buffer holds the rows below the diagonal block of the link BELOW: recompute exactly
      S.chain_stage_doubles = 0; S.chain_remap_blocks = 0;
      for (int J : S.chain_sn) {
        S.chain_stage_doubles = std::max(S.chain_stage_doubles, (size_t)S.sn_nrow[J] * d * S.sn_ncol[J] * d);
        const size_t below = (size_t)(S.sn_nrow[J] - S.sn_ncol[J]);
        S.chain_remap_blocks = std::max(S.chain_remap_blocks, below * (below + 1) / 2);
        S.chain_flops += work[J];
      }
    }
  }
  const int nlinks = (int)S.chain_sn.size();
  if (nlinks > 0) {
    // re-index maps (= the relative indices of the update link -> parent) and the rows that enter at each link
    S.chain_mapptr.assign(nlinks + 1, 0);
    S.chain_new_rows.assign(nlinks, 0u);
    S.chain_colptr.assign(nlinks + 1, 0);
    for (int j = 0; j < nlinks; ++j) {
      const int J = S.chain_sn[j];
      S.chain_colptr[j + 1] = S.chain_colptr[j] + S.sn_ncol[J] * d;
      unsigned fresh = S.sn_nrow[J] >= 32 ? 0xffffffffu : ((1u << S.sn_nrow[J]) - 1u);
      if (j > 0) {
        const int K = S.chain_sn[j - 1];
        // the update K -> J starts at K's first row below its diagonal block (J is K's parent)
        int u = -1;
        for (int q = S.upd_ptr[J]; q < S.upd_ptr[J + 1]; ++q) if (S.upd_k[q] == K) { u = q; break; }
        assert(u >= 0 && S.upd_p0[u] == S.sn_ncol[K]);
        const int nbelow = S.sn_nrow[K] - S.sn_ncol[K];
        for (int a = 0; a < nbelow; ++a) {
          const int r = S.rel[S.upd_relptr[u] + a];
          S.chain_map.push_back(r);
          fresh &= ~(1u << r);
        }
      }
      S.chain_mapptr[j + 1] = (int)S.chain_map.size();
      S.chain_new_rows[j] = fresh;
    }
    // (maps are stored with the RECEIVING link: entry j tells where the rows below link j-1's diagonal block go)
  }

  // ---- 9. numeric plan: destination tiles + their work items, row chunks, level kinds
  // narrow (default): 48 x 48 scalars, DFMA register tiles.  wide (d = 6, throughput-bound factorisations): 96 rows x
  // the whole panel width (72 columns) - one column tile per panel, operands of an update item are 96 x 72 and 72 x 72
  // (10 flops per byte fetched from L2 instead of 6) and the product runs on the FP64 tensor path (chol.cu)
  S.wide = d == 6 && (opt.wide_tiles > 0 || (opt.wide_tiles < 0 && S.flops >= opt.wide_min_flops));
  const int TB = S.wide ? 16 : std::max(1, 48 / d);   // block rows of a tile
  const int TBC = S.wide ? 12 : TB;                   // block columns of a tile
  // a chunk CTA maps block rows to the 32 lanes of a warp: diagonal block + chunk rows + the right-hand-side row
  auto chunk_cap = [&](int J) { return 31 - S.sn_ncol[J]; };
  S.tile_blocks = TB;
  S.tile_blocks_c = TBC;
  S.chunk_blocks = 30;
  S.sn_tile_ptr.assign(ns + 1, 0);
  std::vector<std::vector<int>> tile_lookup(ns);  // [tr * nct + tc] -> tile id or -1
  std::vector<int> sn_nct(ns, 0);
  for (int J = 0; J < ns; ++J) {
    const int ntr = (S.sn_nrow[J] + TB - 1) / TB, nct = (S.sn_ncol[J] + TBC - 1) / TBC;
    sn_nct[J] = nct;
    tile_lookup[J].assign((size_t)ntr * nct, -1);
    for (int tr = 0; tr < ntr; ++tr)
      for (int tc = 0; tc < nct && tc * TBC <= tr * TB + TB - 1; ++tc) {  // tiles that touch the lower triangle
        tile_lookup[J][(size_t)tr * nct + tc] = (int)S.tile_sn.size();
        S.tile_sn.push_back(J); S.tile_r0.push_back(tr * TB); S.tile_c0.push_back(tc * TBC);
      }
    S.sn_tile_ptr[J + 1] = (int)S.tile_sn.size();
  }
  const int ntiles = (int)S.tile_sn.size();
  S.tile_work_ptr.assign(ntiles + 1, 0);
  {
    // the tiles of supernode J are only counted / filled by J: contiguous ranges of supernodes run concurrently
    // (host_parallel.h), the result does not depend on the number of threads
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<int> fill;
      if (pass == 1) {
        for (int t = 0; t < ntiles; ++t) S.tile_work_ptr[t + 1] += S.tile_work_ptr[t];
        const int nw = S.tile_work_ptr[ntiles];
        S.work_u.assign(nw, 0); S.work_a0.assign(nw, 0); S.work_a1.assign(nw, 0); S.work_b0.assign(nw, 0); S.work_b1.assign(nw, 0);
        fill.assign(S.tile_work_ptr.begin(), S.tile_work_ptr.end() - 1);
      }
      parallel_ranges((size_t)ns, range_count((size_t)ns, 256), [&](int, size_t Jb, size_t Je) {
      std::vector<int> bound, cbound;
      for (int J = (int)Jb; J < (int)Je; ++J) {
        const int ntr = (S.sn_nrow[J] + TB - 1) / TB, nct = sn_nct[J];
        for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
          const int K = S.upd_k[u];
          if (S.sn_on_chain[K]) continue;  // chain link -> chain link: stays in the chain CTA's registers
          const int h = S.sn_nrow[K] - S.upd_p0[u], w = S.upd_p1[u] - S.upd_p0[u];
          const int* rel = S.rel.data() + S.upd_relptr[u];
          bound.assign(ntr + 1, h);  // bound[t] = first a with rel[a] >= t*TB
          int a = 0;
          for (int t = 0; t <= ntr; ++t) {
            while (a < h && rel[a] < t * TB) ++a;
            bound[t] = a;
          }
          cbound.assign(nct + 1, h);  // cbound[t] = first a with rel[a] >= t*TBC
          a = 0;
          for (int t = 0; t <= nct; ++t) {
            while (a < h && rel[a] < t * TBC) ++a;
            cbound[t] = a;
          }
          for (int tc = 0; tc < nct; ++tc) {
            const int b0 = std::min(cbound[tc], w), b1 = std::min(cbound[tc + 1], w);
            if (b0 >= b1) continue;
            for (int tr = 0; tr < ntr; ++tr) {
              const int a0 = bound[tr], a1 = bound[tr + 1];
              if (a0 >= a1) continue;
              if (a1 - 1 < b0) continue;  // entirely above the diagonal of the update
              const int t = tile_lookup[J][(size_t)tr * nct + tc];
              if (t < 0) continue;        // (cannot happen: a row >= column entry lies in a tile touching the lower triangle)
              if (pass == 0) S.tile_work_ptr[t + 1]++;
              else { int q = fill[t]++; S.work_u[q] = u; S.work_a0[q] = a0; S.work_a1[q] = a1; S.work_b0[q] = b0; S.work_b1[q] = b1; }
            }
          }
        }
      }
      });
    }
  }
  if (opt.sort_items_by_level) {
    // inside a tile: items whose source supernode completes early first (ascending task level, then ascending
    // supernode), so that a group never waits for a late descendant while early ones are ready.  The order is part of
    // the plan: sums stay deterministic.
    auto lvl = [&](int q) { return tlevel[task_of[S.upd_k[S.work_u[q]]]]; };
    parallel_ranges((size_t)ntiles, range_count((size_t)ntiles, 1024), [&](int, size_t tb, size_t te) {
    std::vector<int> idx, tmp;
    for (int t = (int)tb; t < (int)te; ++t) {
      const int w0 = S.tile_work_ptr[t], w1 = S.tile_work_ptr[t + 1];
      if (w1 - w0 < 2) continue;
      idx.resize(w1 - w0);
      for (int q = w0; q < w1; ++q) idx[q - w0] = q;
      std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lvl(x) < lvl(y); });
      for (std::vector<int>* arr : {&S.work_u, &S.work_a0, &S.work_a1, &S.work_b0, &S.work_b1}) {
        tmp.resize(w1 - w0);
        for (int q = 0; q < w1 - w0; ++q) tmp[q] = (*arr)[idx[q]];
        std::copy(tmp.begin(), tmp.end(), arr->begin() + w0);
      }
    }
    });
  }
  {
    const int nw = (int)S.work_u.size();
    S.work_koff.resize(nw); S.work_reloff.resize(nw); S.work_mk.resize(nw); S.work_nk.resize(nw); S.work_ksn.resize(nw);
    parallel_ranges((size_t)nw, range_count((size_t)nw, (size_t)1 << 16), [&](int, size_t qb, size_t qe) {
      for (size_t q = qb; q < qe; ++q) {
        const int u = S.work_u[q], K = S.upd_k[u];
        S.work_koff[q] = S.sn_lptr[K] + (int64_t)S.upd_p0[u] * d;
        S.work_reloff[q] = S.upd_relptr[u];
        S.work_mk[q] = S.sn_nrow[K] * d;
        S.work_nk[q] = S.sn_ncol[K] * d;
        S.work_ksn[q] = K;
      }
    });
  }
  S.sn_chunk_ptr.assign(ns + 1, 0);
  S.sn_dinvptr.assign(ns + 1, 0);
  for (int J = 0; J < ns; ++J) {
    const int below = S.sn_nrow[J] - S.sn_ncol[J];
    int b0 = S.sn_ncol[J];
    if (below == 0) { S.chunk_sn.push_back(J); S.chunk_b0.push_back(b0); S.chunk_nb.push_back(0); }
    for (int rem = below; rem > 0;) {
      const int nbk = std::min(rem, chunk_cap(J));
      S.chunk_sn.push_back(J); S.chunk_b0.push_back(b0); S.chunk_nb.push_back(nbk);
      b0 += nbk; rem -= nbk;
    }
    S.sn_chunk_ptr[J + 1] = (int)S.chunk_sn.size();
    S.sn_dinvptr[J + 1] = S.sn_dinvptr[J] + (int64_t)S.sn_ncol[J] * d * S.sn_ncol[J] * d;
  }
  S.dinv_doubles = S.sn_dinvptr[ns];
  S.level_kind.assign(S.nlevels, 0);
  S.level_chunk_ptr.assign(S.nlevels + 1, 0);
  S.level_group_ptr.assign(S.nlevels + 1, 0);
  S.level_rtile_ptr.assign(S.nlevels + 1, 0);
  // wide tiles: a group of 32 items (the scratch of split tiles halves: 97 -> 43 GB at 1M poses; measured +2 % at 250k
  // poses, -2 % at 90k against 16)
  S.group_items = std::max(1, S.wide && opt.group_items == 16 ? opt.wide_group_items : opt.group_items);
  S.sn_nupd.assign(ns, 0);
  S.sn_nchunk.assign(ns, 0);
  for (int J = 0; J < ns; ++J) S.sn_nchunk[J] = S.sn_chunk_ptr[J + 1] - S.sn_chunk_ptr[J];
  int slots = 0;  // scratch slots are numbered globally
  S.sn_cptr.assign(ns + 1, 0);
  for (int J = 0; J < ns; ++J) S.sn_cptr[J + 1] = S.sn_cptr[J] + (int64_t)(S.sn_nrow[J] - S.sn_ncol[J]) * d;
  {  // forward-solve gather lists
    const int n = nb * d;
    S.fwd_ptr.assign(n + 1, 0);
    for (int J = 0; J < ns; ++J)
      for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
        const int K = S.upd_k[u];
        const int* kr = S.sn_rows.data() + S.sn_rowptr[K];
        for (int p = S.upd_p0[u]; p < S.upd_p1[u]; ++p)
          for (int rr = 0; rr < d; ++rr) S.fwd_ptr[kr[p] * d + rr + 1]++;
      }
    for (int i = 0; i < n; ++i) S.fwd_ptr[i + 1] += S.fwd_ptr[i];
    S.fwd_src.assign(S.fwd_ptr[n], 0);
    std::vector<int> fill(S.fwd_ptr.begin(), S.fwd_ptr.end() - 1);
    for (int J = 0; J < ns; ++J)  // ascending J, then ascending K inside: every row list is in (K, p) order
      for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
        const int K = S.upd_k[u];
        const int* kr = S.sn_rows.data() + S.sn_rowptr[K];
        for (int p = S.upd_p0[u]; p < S.upd_p1[u]; ++p)
          for (int rr = 0; rr < d; ++rr)
            S.fwd_src[fill[kr[p] * d + rr]++] = (int)(S.sn_cptr[K] + (int64_t)(p - S.sn_ncol[K]) * d + rr);
      }
  }
  if (nlinks > 0) {  // forward substitution of the chain: per scalar column the contributions of sources BELOW the chain
    S.chain_fwd_ptr.assign(S.chain_colptr[nlinks] + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<int> fill;
      if (pass == 1) {
        for (int i = 0; i < S.chain_colptr[nlinks]; ++i) S.chain_fwd_ptr[i + 1] += S.chain_fwd_ptr[i];
        S.chain_fwd_src.assign(S.chain_fwd_ptr[S.chain_colptr[nlinks]], 0);
        fill.assign(S.chain_fwd_ptr.begin(), S.chain_fwd_ptr.end() - 1);
      }
      for (int j = 0; j < nlinks; ++j) {
        const int J = S.chain_sn[j];
        for (int u = S.upd_ptr[J]; u < S.upd_ptr[J + 1]; ++u) {
          const int K = S.upd_k[u];
          if (S.sn_on_chain[K]) continue;
          const int* kr = S.sn_rows.data() + S.sn_rowptr[K];
          for (int p = S.upd_p0[u]; p < S.upd_p1[u]; ++p)
            for (int rr = 0; rr < d; ++rr) {
              const int col = S.chain_colptr[j] + (kr[p] - S.sn_col0[J]) * d + rr;
              if (pass == 0) S.chain_fwd_ptr[col + 1]++;
              else S.chain_fwd_src[fill[col]++] = (int)(S.sn_cptr[K] + (int64_t)(p - S.sn_ncol[K]) * d + rr);
            }
        }
      }
    }
  }
  S.task_on_chain.assign(nt, 0);
  for (int t = 0; t < nt; ++t)
    if (S.task_ptr[t + 1] - S.task_ptr[t] == 1 && S.sn_on_chain[S.task_sn[S.task_ptr[t]]]) S.task_on_chain[t] = 1;
  std::vector<int> group_level;  // per group: the level it is listed at
  for (int l = 0; l < S.nlevels; ++l) {
    // chain links are left to the chain kernel: the level keeps only the GROUP tasks that bring them the updates of
    // the supernodes below the chain
    bool singletons = true;
    int tiles = 0, chunks = 0, ntask = 0;
    for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
      if (S.task_on_chain[t]) continue;
      ++ntask;
      if (S.task_ptr[t + 1] - S.task_ptr[t] != 1) singletons = false;
      for (int q = S.task_ptr[t]; q < S.task_ptr[t + 1]; ++q) {
        const int J = S.task_sn[q];
        tiles += S.sn_tile_ptr[J + 1] - S.sn_tile_ptr[J];
        chunks += S.sn_chunk_ptr[J + 1] - S.sn_chunk_ptr[J];
      }
    }
    const bool split = singletons && (tiles > ntask || chunks > ntask);
    if (split) S.level_kind[l] = 1;
    {
      for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
        if (!split && !S.task_on_chain[t]) continue;  // kind 0: the task's own CTA applies its updates
        const int J = S.task_sn[S.task_ptr[t]];
        for (int q = S.sn_tile_ptr[J]; q < S.sn_tile_ptr[J + 1]; ++q) {
          const int w0 = S.tile_work_ptr[q], w1 = S.tile_work_ptr[q + 1];
          if (w1 == w0) continue;
          // (option split_late_items, off: the items that come from the level just below the destination - they can only
          // start when that level is complete - in groups of their own, so that no early item waits in a group that is
          // listed late.  Measured: 3-4 % slower, the early items are not what the levels wait for)
          std::vector<std::pair<int, int>> cuts;  // [begin, end) of every group
          {
            int ws = w1;
            if (opt.split_late_items && opt.sort_items_by_level)
              while (ws > w0 && tlevel[task_of[S.upd_k[S.work_u[ws - 1]]]] >= l - 1) --ws;
            auto chop = [&](int b, int e) {
              if (b >= e) return;
              const int n = (e - b + S.group_items - 1) / S.group_items;
              for (int gi = 0; gi < n; ++gi) cuts.push_back({b + (int)((long long)(e - b) * gi / n), b + (int)((long long)(e - b) * (gi + 1) / n)});
            };
            if (ws > w0 && ws < w1) { chop(w0, ws); chop(ws, w1); }
            else chop(w0, w1);
          }
          const int ng = (int)cuts.size();
          S.sn_nupd[J]++;
          // a group is listed as early as its sources allow: one level above the latest of them (as-soon-as-possible),
          // not at the level of its destination - when the factorisation reaches the top of the tree, where a level is
          // one or two panels, only the groups that really depend on the previous level are left
          // ... plus `group_slack` levels (default 0).  Inside a level the groups whose sources are two or more levels
          // back come first (ready when taken), then the groups that need the level just below (key 2 l + 1), then the
          // level's panel factorisations.  Measured on the 90k / 250k-pose spheres: as-soon-as-possible listing is worth
          // 2 %, one or two levels of slack cost 5 % - what the CTAs really waited for were the SUBTREE tasks (step 8)
          auto asap = [&](int g0, int g1) {
            if (!opt.groups_asap) return 2 * l + 1;
            int m = -1;
            for (int w = g0; w < g1; ++w) m = std::max(m, tlevel[task_of[S.upd_k[S.work_u[w]]]]);
            const int lv = std::min(l, m + 1 + std::max(0, opt.group_slack));
            return 2 * lv + (m == lv - 1 ? 1 : 0);
          };
          if (ng == 1) {
            S.group_tile.push_back(q); S.group_w0.push_back(w0); S.group_w1.push_back(w1); S.group_slot.push_back(-1);
            S.group_rtile.push_back(-1);
            group_level.push_back(asap(w0, w1));
          } else {
            S.rtile_tile.push_back(q); S.rtile_slot0.push_back(slots); S.rtile_nslots.push_back(ng);
            for (int gi = 0; gi < ng; ++gi) {
              S.group_tile.push_back(q);
              S.group_rtile.push_back((int)S.rtile_tile.size() - 1);
              S.group_w0.push_back(cuts[gi].first);
              S.group_w1.push_back(cuts[gi].second);
              S.group_slot.push_back(slots++);
              group_level.push_back(asap(S.group_w0.back(), S.group_w1.back()));
            }
          }
        }
        if (!S.task_on_chain[t])
          for (int q = S.sn_chunk_ptr[J]; q < S.sn_chunk_ptr[J + 1]; ++q) S.level_chunks.push_back(q);
      }
    }
    S.max_group_slots = slots;
    S.level_chunk_ptr[l + 1] = (int)S.level_chunks.size();
    S.level_group_ptr[l + 1] = (int)S.group_tile.size();
    S.level_rtile_ptr[l + 1] = (int)S.rtile_tile.size();
  }
  {  // groups bucketed by their listing key 2 * level + (needs the level just below) (stable: creation order inside a bucket)
    const int ngp = (int)S.group_tile.size();
    std::vector<int> cnt(2 * S.nlevels + 1, 0), idx(ngp);
    for (int g = 0; g < ngp; ++g) cnt[group_level[g] + 1]++;
    for (int l = 0; l < 2 * S.nlevels; ++l) cnt[l + 1] += cnt[l];
    for (int l = 0; l <= S.nlevels; ++l) S.level_group_ptr[l] = cnt[2 * l];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int g = 0; g < ngp; ++g) idx[fill[group_level[g]]++] = g;
    for (std::vector<int>* arr : {&S.group_tile, &S.group_w0, &S.group_w1, &S.group_slot, &S.group_rtile}) {
      std::vector<int> tmp(ngp);
      for (int g = 0; g < ngp; ++g) tmp[g] = (*arr)[idx[g]];
      arr->swap(tmp);
    }
  }
  // ---- 10. dataflow task list (level-major; inside a split level: groups, then chunks; the reduction of a split
  //          tile is done by whichever of its groups finishes last)
  for (int l = 0; l < S.nlevels; ++l) {
    if (S.level_kind[l] == 0)
      for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
        if (!S.task_on_chain[t]) { S.flow_kind.push_back(0); S.flow_arg.push_back(t); }
    // kind 1: every supernode of the level; kind 0: only the GROUP tasks whose destination is a chain link
    for (int g = S.level_group_ptr[l]; g < S.level_group_ptr[l + 1]; ++g) { S.flow_kind.push_back(1); S.flow_arg.push_back(g); }
    for (int c = S.level_chunk_ptr[l]; c < S.level_chunk_ptr[l + 1]; ++c) { S.flow_kind.push_back(3); S.flow_arg.push_back(S.level_chunks[c]); }
  }
  {
    std::vector<int> sn_task(ns, -1);
    for (int t = 0; t < nt; ++t) for (int q = S.task_ptr[t]; q < S.task_ptr[t + 1]; ++q) sn_task[S.task_sn[q]] = t;
    S.task_parent.assign(nt, -1);
    for (int t = 0; t < nt; ++t) {
      const int root_sn = S.task_sn[S.task_ptr[t + 1] - 1];  // ascending order inside a task: the root comes last
      const int p = S.sn_parent[root_sn];
      S.task_parent[t] = p < 0 ? -1 : sn_task[p];
    }
  }
  return S;
}

}  // namespace g2o_b200
