// nested_dissection.cpp - optional ordering for parallelism (see symbolic.h: SymbolicOptions::nd_levels).
//
// The reference orders the block pattern with AMD (cs_amd / CHOLMOD_AMD), which is what block_amd.cpp reproduces
// bit for bit and what this library uses by default.  On band-like reduced camera systems (a ring of cameras) AMD
// eliminates from the ends inwards: the elimination tree is one or two chains as long as the matrix, and a chain is a
// sequential dependency no GPU can parallelise.  Nested dissection cuts the graph with small vertex separators first
// (George's automatic nested dissection: the middle level of a breadth-first level structure rooted at a
// pseudo-peripheral node), orders the parts independently - with the same AMD - and the separators last: 2^levels
// independent subtrees instead of one chain, at the price of somewhat more fill.  Pure integer host work.
#include <algorithm>
#include <numeric>
#include <queue>

#include "block_amd.h"
#include "symbolic.h"

namespace g2o_b200 {
namespace {

struct Graph {
  int n;
  std::vector<int> ptr, adj;  // symmetric adjacency without self loops
};

// breadth-first level structure of the subgraph `mask == id` from `root`; returns the levels, visits in `order`
int bfs_levels(const Graph& G, const std::vector<int>& mask, int id, int root, std::vector<int>& level,
               std::vector<int>& order) {
  order.clear();
  order.push_back(root);
  level[root] = 0;
  int depth = 0;
  for (size_t h = 0; h < order.size(); ++h) {
    const int v = order[h];
    for (int p = G.ptr[v]; p < G.ptr[v + 1]; ++p) {
      const int w = G.adj[p];
      if (mask[w] != id || level[w] >= 0) continue;
      level[w] = level[v] + 1;
      depth = std::max(depth, level[w]);
      order.push_back(w);
    }
  }
  return depth;
}

// AMD of the subgraph induced by `nodes` (the same block_amd as everywhere else), appended to `perm`
void amd_of_subgraph(const Graph& G, const std::vector<int>& nodes, std::vector<int>& local, std::vector<int>& perm) {
  const int m = (int)nodes.size();
  if (m == 0) return;
  std::vector<int> sorted(nodes);
  std::sort(sorted.begin(), sorted.end());
  for (int i = 0; i < m; ++i) local[sorted[i]] = i;
  std::vector<int> cp(m + 1, 0), ri;
  for (int j = 0; j < m; ++j) {
    const int v = sorted[j];
    std::vector<int> rows;
    for (int p = G.ptr[v]; p < G.ptr[v + 1]; ++p) {
      const int w = G.adj[p];
      if (local[w] >= 0 && local[w] < j) rows.push_back(local[w]);
    }
    rows.push_back(j);
    std::sort(rows.begin(), rows.end());
    ri.insert(ri.end(), rows.begin(), rows.end());
    cp[j + 1] = (int)ri.size();
  }
  const std::vector<int> P = block_amd(m, cp.data(), ri.data());
  for (int k = 0; k < m; ++k) perm.push_back(sorted[P[k]]);
  for (int i = 0; i < m; ++i) local[sorted[i]] = -1;
}

struct Dissector {
  const Graph& G;
  int min_part;
  std::vector<int> mask, level, local, perm, order;
  int next_id = 1;
  Dissector(const Graph& g, int mp) : G(g), min_part(mp), mask(g.n, 0), level(g.n, -1), local(g.n, -1) {}

  void run(const std::vector<int>& nodes, int levels) {
    if (levels <= 0 || (int)nodes.size() < 2 * min_part) { amd_of_subgraph(G, nodes, local, perm); return; }
    const int id = next_id++;
    for (int v : nodes) mask[v] = id;
    // connected components: dissect each one on its own
    std::vector<std::vector<int>> comps;
    for (int v : nodes) level[v] = -1;
    for (int v : nodes) {
      if (level[v] >= 0) continue;
      bfs_levels(G, mask, id, v, level, order);
      comps.push_back(order);
    }
    if (comps.size() > 1) {
      for (auto& c : comps) {
        std::vector<int> cc(c);
        run(cc, levels);  // note: run() re-labels mask for its own nodes
      }
      return;
    }
    // pseudo-peripheral root: repeat BFS from the last vertex of the deepest level structure
    int root = nodes[0], depth = -1;
    for (int iter = 0; iter < 4; ++iter) {
      for (int v : nodes) level[v] = -1;
      const int d = bfs_levels(G, mask, id, root, level, order);
      if (d <= depth) break;
      depth = d;
      root = order.back();
    }
    for (int v : nodes) level[v] = -1;
    depth = bfs_levels(G, mask, id, root, level, order);
    if (depth < 2) { amd_of_subgraph(G, nodes, local, perm); return; }  // (nearly) complete graph: nothing to cut
    // separator = the level that balances the two sides
    std::vector<int> cnt(depth + 1, 0);
    for (int v : nodes) cnt[level[v]]++;
    int best = 1, below = cnt[0];
    long long best_cost = -1;
    for (int m = 1; m < depth; ++m) {
      const int above = (int)nodes.size() - below - cnt[m];
      const long long cost = (long long)std::abs(below - above) + 4ll * cnt[m];  // balance + separator size
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = m; }
      below += cnt[m];
    }
    std::vector<int> A, B, S;
    for (int v : nodes) (level[v] < best ? A : level[v] > best ? B : S).push_back(v);
    for (int v : nodes) level[v] = -1;
    run(A, levels - 1);
    run(B, levels - 1);
    amd_of_subgraph(G, S, local, perm);
  }
};

}  // namespace

// colptr/rowidx: upper block pattern (rows <= col).  levels: depth of the dissection (2^levels parts).
std::vector<int> nested_dissection_order(int nb, const int* colptr, const int* rowidx, int levels, int min_part) {
  Graph G;
  G.n = nb;
  G.ptr.assign(nb + 1, 0);
  for (int c = 0; c < nb; ++c)
    for (int p = colptr[c]; p < colptr[c + 1]; ++p)
      if (rowidx[p] != c) { G.ptr[c + 1]++; G.ptr[rowidx[p] + 1]++; }
  for (int i = 0; i < nb; ++i) G.ptr[i + 1] += G.ptr[i];
  G.adj.resize(G.ptr[nb]);
  {
    std::vector<int> f(G.ptr.begin(), G.ptr.end() - 1);
    for (int c = 0; c < nb; ++c)
      for (int p = colptr[c]; p < colptr[c + 1]; ++p)
        if (rowidx[p] != c) { G.adj[f[c]++] = rowidx[p]; G.adj[f[rowidx[p]]++] = c; }
  }
  Dissector D(G, std::max(1, min_part));
  std::vector<int> all(nb);
  std::iota(all.begin(), all.end(), 0);
  D.run(all, levels);
  return D.perm;
}

}  // namespace g2o_b200
