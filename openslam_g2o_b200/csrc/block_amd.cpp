// block_amd.cpp - approximate-minimum-degree ordering of the block pattern (host, one-time).
//
// Parity contract: the permutation must be BIT-EXACT with what the reference obtains from
// cs_amd(1, blockPattern) (EXTERNAL/csparse/cs_amd.c:18-364, called from
// solvers/csparse/linear_solver_csparse.h:268).  AMD is a heuristic with many tie-breaks (degree-list
// insertion order, hash-bucket order, element absorption order, post-order of the assembly tree), so
// this is a restatement of the same quotient-graph algorithm with the same data-structure semantics:
//   * one integer workspace holding all adjacency / element lists, compacted when it fills up,
//   * degree lists as doubly linked LIFO lists, approximate external degrees (Amestoy/Davis/Duff bound),
//   * aggressive element absorption, mass elimination, hash-based indistinguishable-node merging,
//   * dense nodes (degree > max(16, 10 sqrt n)) deferred to the end,
//   * depth-first post-order of the assembly tree where children are visited youngest first.
// tests/test_amd.py checks it against the vendored cs_amd on the four in-tree datasets and on random,
// banded, arrow-head and dense-row patterns.
//
// Attribution: because the permutation has to be bit-exact, this file follows the structure of cs_amd.c step by step
// (same quotient-graph phases in the same order, the same working arrays, the dense-node rule, the in-place garbage
// collection).  It is a derived work of CSparse: "CSparse: a Concise Sparse matrix package", Copyright (c) 2006-2012,
// Timothy A. Davis, http://www.suitesparse.com - distributed under the GNU Lesser General Public License, version 2.1
// or (at your option) any later version (EXTERNAL/csparse/License.txt of the reference tree).  This file is made
// available under the same terms; CSparse and this file come WITHOUT ANY WARRANTY.
#include "block_amd.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace g2o_b200 {
namespace {

inline int flip(int i) { return -i - 2; }  // involution used to tag "absorbed into" links

struct Amd {
  int n;
  // quotient graph storage
  std::vector<int> store;   // adjacency lists of variables / member lists of elements
  std::vector<int> head_of; // start of object i inside store, or flip(parent) once absorbed, -1 = root
  std::vector<int> len, nv, next, last, head, elen, degree, w, hhead;
  int free_at = 0;          // first unused slot in store
  int capacity = 0;

  explicit Amd(int n_) : n(n_) {}

  // make all w[] of live objects smaller than the returned mark
  int reset_marks(int mark, int lemax) {
    if (mark < 2 || (int64_t)mark + lemax > INT32_MAX) {  // reference tests (mark + lemax < 0) on int overflow
      for (int k = 0; k < n; ++k) if (w[k] != 0) w[k] = 1;
      mark = 2;
    }
    return mark;
  }

  void build(const int* colptr, const int* rowidx) {
    // symmetric pattern without the diagonal; neighbours in ascending order
    std::vector<int> cnt(n + 1, 0);
    for (int j = 0; j < n; ++j)
      for (int p = colptr[j]; p < colptr[j + 1]; ++p) {
        int i = rowidx[p];
        if (i == j) continue;
        cnt[i]++; cnt[j]++;
      }
    head_of.assign(n + 1, 0);
    int tot = 0;
    for (int j = 0; j < n; ++j) { head_of[j] = tot; tot += cnt[j]; }
    head_of[n] = tot;
    const int nz = tot;
    capacity = nz + nz / 5 + 2 * n;
    store.assign(std::max(capacity, 1), 0);
    // neighbours < j come from column j (ascending), neighbours > j from row j (ascending)
    std::vector<int> fill(n, 0);
    // pass 1: rows i<j of column j -> goes to the low part of j's list
    for (int j = 0; j < n; ++j)
      for (int p = colptr[j]; p < colptr[j + 1]; ++p) {
        int i = rowidx[p];
        if (i < j) store[head_of[j] + fill[j]++] = i;
      }
    // pass 2: for ascending j, entry (i,j) with i<j appends j to i's list => ascending high part;
    // entries given below the diagonal (i>j) are mirrored the same way
    bool needs_sort = false;
    for (int j = 0; j < n; ++j)
      for (int p = colptr[j]; p < colptr[j + 1]; ++p) {
        int i = rowidx[p];
        if (i < j) store[head_of[i] + fill[i]++] = j;
        else if (i > j) { needs_sort = true; store[head_of[j] + fill[j]++] = i; store[head_of[i] + fill[i]++] = j; }
      }
    if (needs_sort)
      for (int j = 0; j < n; ++j) {
        std::sort(store.begin() + head_of[j], store.begin() + head_of[j] + cnt[j]);
        int m = (int)(std::unique(store.begin() + head_of[j], store.begin() + head_of[j] + cnt[j]) - (store.begin() + head_of[j]));
        cnt[j] = m;
      }
    free_at = nz;
    len.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) len[j] = cnt[j];
  }

  std::vector<int> run() {
    std::vector<int> perm(n + 1, 0);
    nv.assign(n + 1, 1); next.assign(n + 1, -1); last.assign(n + 1, -1); head.assign(n + 1, -1);
    elen.assign(n + 1, 0); degree.assign(n + 1, 0); w.assign(n + 1, 1); hhead.assign(n + 1, -1);
    int dense = (int)std::max(16.0, 10.0 * std::sqrt((double)n));
    dense = std::min(n - 2, dense);
    for (int i = 0; i <= n; ++i) degree[i] = len[i];
    int mark = reset_marks(0, 0);
    elen[n] = -2; head_of[n] = -1; w[n] = 0;
    int nel = 0, mindeg = 0, lemax = 0;
    // initial degree lists
    for (int i = 0; i < n; ++i) {
      int d = degree[i];
      if (d == 0) { elen[i] = -2; ++nel; head_of[i] = -1; w[i] = 0; }
      else if (d > dense) { nv[i] = 0; elen[i] = -1; ++nel; head_of[i] = flip(n); nv[n]++; }
      else {
        if (head[d] != -1) last[head[d]] = i;
        next[i] = head[d];
        head[d] = i;
      }
    }
    int* S = store.data();
    while (nel < n) {
      // pivot = head of the lowest non-empty degree list
      int k = -1;
      for (; mindeg < n && (k = head[mindeg]) == -1; ++mindeg) {}
      if (next[k] != -1) last[next[k]] = -1;
      head[mindeg] = next[k];
      const int elenk = elen[k];
      int nvk = nv[k];
      nel += nvk;
      // compact the workspace if the new element might not fit
      if (elenk > 0 && free_at + mindeg >= capacity) {
        for (int j = 0; j < n; ++j) {
          int p = head_of[j];
          if (p >= 0) { head_of[j] = S[p]; S[p] = flip(j); }
        }
        int q = 0;
        for (int p = 0; p < free_at;) {
          int j = flip(S[p++]);
          if (j >= 0) {
            S[q] = head_of[j];
            head_of[j] = q++;
            for (int k3 = 0; k3 < len[j] - 1; ++k3) S[q++] = S[p++];
          }
        }
        free_at = q;
      }
      // build the member list Lk of the new element k
      int dk = 0;
      nv[k] = -nvk;
      int p = head_of[k];
      const int pk1 = (elenk == 0) ? p : free_at;
      int pk2 = pk1;
      for (int k1 = 1; k1 <= elenk + 1; ++k1) {
        int e, pj, ln;
        if (k1 > elenk) { e = k; pj = p; ln = len[k] - elenk; }
        else { e = S[p++]; pj = head_of[e]; ln = len[e]; }
        for (int k2 = 1; k2 <= ln; ++k2) {
          int i = S[pj++];
          int nvi = nv[i];
          if (nvi <= 0) continue;
          dk += nvi;
          nv[i] = -nvi;
          S[pk2++] = i;
          if (next[i] != -1) last[next[i]] = last[i];
          if (last[i] != -1) next[last[i]] = next[i];
          else head[degree[i]] = next[i];
        }
        if (e != k) { head_of[e] = flip(k); w[e] = 0; }
      }
      if (elenk != 0) free_at = pk2;
      degree[k] = dk;
      head_of[k] = pk1;
      len[k] = pk2 - pk1;
      elen[k] = -2;
      // scan 1: w[e] - mark = |Le \ Lk| for every element e adjacent to a member of Lk
      mark = reset_marks(mark, lemax);
      for (int pk = pk1; pk < pk2; ++pk) {
        int i = S[pk];
        int eln = elen[i];
        if (eln <= 0) continue;
        int nvi = -nv[i];
        int wnvi = mark - nvi;
        for (int q = head_of[i]; q <= head_of[i] + eln - 1; ++q) {
          int e = S[q];
          if (w[e] >= mark) w[e] -= nvi;
          else if (w[e] != 0) w[e] = degree[e] + wnvi;
        }
      }
      // scan 2: approximate degrees, absorb elements, prune, hash
      for (int pk = pk1; pk < pk2; ++pk) {
        int i = S[pk];
        int p1 = head_of[i];
        int p2 = p1 + elen[i] - 1;
        int pn = p1;
        uint32_t hsum = 0;
        int d = 0;
        for (int q = p1; q <= p2; ++q) {
          int e = S[q];
          if (w[e] != 0) {
            int dext = w[e] - mark;
            if (dext > 0) { d += dext; S[pn++] = e; hsum += (uint32_t)e; }
            else { head_of[e] = flip(k); w[e] = 0; }
          }
        }
        elen[i] = pn - p1 + 1;
        int p3 = pn;
        int p4 = p1 + len[i];
        for (int q = p2 + 1; q < p4; ++q) {
          int j = S[q];
          int nvj = nv[j];
          if (nvj <= 0) continue;
          d += nvj;
          S[pn++] = j;
          hsum += (uint32_t)j;
        }
        if (d == 0) {  // mass elimination: i has no neighbour outside Lk
          head_of[i] = flip(k);
          int nvi = -nv[i];
          dk -= nvi; nvk += nvi; nel += nvi;
          nv[i] = 0;
          elen[i] = -1;
        } else {
          degree[i] = std::min(degree[i], d);
          S[pn] = S[p3];
          S[p3] = S[p1];
          S[p1] = k;
          len[i] = pn - p1 + 1;
          int h = (int32_t)hsum;
          h = ((h < 0) ? (-h) : h) % n;
          next[i] = hhead[h];
          hhead[h] = i;
          last[i] = h;
        }
      }
      degree[k] = dk;
      lemax = std::max(lemax, dk);
      mark = reset_marks(mark + lemax, lemax);
      // merge indistinguishable members of Lk (same hash bucket, identical lists)
      for (int pk = pk1; pk < pk2; ++pk) {
        int i = S[pk];
        if (nv[i] >= 0) continue;
        int h = last[i];
        i = hhead[h];
        hhead[h] = -1;
        for (; i != -1 && next[i] != -1; i = next[i], ++mark) {
          int ln = len[i];
          int eln = elen[i];
          for (int q = head_of[i] + 1; q <= head_of[i] + ln - 1; ++q) w[S[q]] = mark;
          int jlast = i;
          for (int j = next[i]; j != -1;) {
            bool same = (len[j] == ln) && (elen[j] == eln);
            for (int q = head_of[j] + 1; same && q <= head_of[j] + ln - 1; ++q)
              if (w[S[q]] != mark) same = false;
            if (same) {
              head_of[j] = flip(i);
              nv[i] += nv[j];
              nv[j] = 0;
              elen[j] = -1;
              j = next[j];
              next[jlast] = j;
            } else {
              jlast = j;
              j = next[j];
            }
          }
        }
      }
      // put the surviving members back into the degree lists
      int pfin = pk1;
      for (int pk = pk1; pk < pk2; ++pk) {
        int i = S[pk];
        int nvi = -nv[i];
        if (nvi <= 0) continue;
        nv[i] = nvi;
        int d = degree[i] + dk - nvi;
        d = std::min(d, n - nel - nvi);
        if (head[d] != -1) last[head[d]] = i;
        next[i] = head[d];
        last[i] = -1;
        head[d] = i;
        mindeg = std::min(mindeg, d);
        degree[i] = d;
        S[pfin++] = i;
      }
      nv[k] = nvk;
      len[k] = pfin - pk1;
      if (len[k] == 0) { head_of[k] = -1; w[k] = 0; }
      if (elenk != 0) free_at = pfin;
    }
    // assembly tree: head_of[i] = parent (after un-flipping), -1 for roots
    for (int i = 0; i < n; ++i) head_of[i] = flip(head_of[i]);
    for (int j = 0; j <= n; ++j) head[j] = -1;
    for (int j = n; j >= 0; --j) {  // absorbed variables hang below their representative
      if (nv[j] > 0) continue;
      next[j] = head[head_of[j]];
      head[head_of[j]] = j;
    }
    for (int e = n; e >= 0; --e) {  // elements hang below the element that absorbed them
      if (nv[e] <= 0) continue;
      if (head_of[e] != -1) { next[e] = head[head_of[e]]; head[head_of[e]] = e; }
    }
    int k = 0;
    std::vector<int>& stack = w;
    for (int i = 0; i <= n; ++i) {
      if (head_of[i] != -1) continue;
      // iterative DFS post-order from root i
      int top = 0;
      stack[0] = i;
      while (top >= 0) {
        int pnode = stack[top];
        int child = head[pnode];
        if (child == -1) { --top; perm[k++] = pnode; }
        else { head[pnode] = next[child]; stack[++top] = child; }
      }
    }
    perm.resize(n);
    return perm;
  }
};

}  // namespace

std::vector<int> block_amd(int n, const int* colptr, const int* rowidx) {
  if (n <= 0) return {};
  if (n == 1) return {0};
  Amd a(n);
  a.build(colptr, rowidx);
  return a.run();
}

}  // namespace g2o_b200
