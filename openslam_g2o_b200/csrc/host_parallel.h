// host_parallel.h - contiguous-range fork/join for the host-side structure phase (graph ingest, Schur plan).
//
// Every use is written so that the RESULT does not depend on the number of threads: a thread owns a contiguous range
// of the output, offsets come from prefix sums, sorts use total orders.  b200_debug_upload_digest pins that
// (tests/test_graph_host.py).  G2O_B200_HOST_THREADS overrides the thread count (1 = everything inline).
// Rules for the bodies: plain host memory only - no CUDA calls, no DevBuf (host_only_flag() is thread-local and unset
// in the workers) - and nothing that throws (an exception escaping a worker terminates the process).
#pragma once
#if defined(__linux__)
#include <sched.h>
#endif
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <system_error>
#include <thread>
#include <vector>

namespace g2o_b200 {

inline int host_threads() {
  if (const char* e = getenv("G2O_B200_HOST_THREADS")) return std::max(1, atoi(e));
  unsigned n = std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;  // the cores this process may run on (taskset, container cpusets), not the cores of the machine
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min<unsigned>(n ? n : 1u, (unsigned)CPU_COUNT(&set));
#endif
  return (int)std::min(16u, std::max(1u, n));
}

// number of ranges [0, n) is cut into: at most host_threads(), at least `grain` items each
inline int range_count(size_t n, size_t grain) {
  if (const char* e = getenv("G2O_B200_HOST_GRAIN")) grain = (size_t)std::max(1, atoi(e));  // tests: many ranges on small inputs
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), n / std::max<size_t>(grain, 1)));
}
inline size_t range_begin(size_t n, int parts, int t) { return (size_t)((unsigned long long)n * (unsigned)t / (unsigned)parts); }

// f(t, begin, end) for t = 0..parts-1 over contiguous ranges of [0, n); inline when parts == 1
template <typename F>
inline void parallel_ranges(size_t n, int parts, F&& f) {
  if (parts <= 1) { f(0, (size_t)0, n); return; }
  std::vector<std::thread> th;
  th.reserve(parts - 1);
  int started = 1;  // ranges [1, started) run on their own thread
  try {
    for (; started < parts; ++started) {
      const int t = started;
      th.emplace_back([&f, n, parts, t] { f(t, range_begin(n, parts, t), range_begin(n, parts, t + 1)); });
    }
  } catch (const std::system_error&) {  // no more threads to be had (container limits): the rest runs here
  }
  f(0, (size_t)0, range_begin(n, parts, 1));
  for (int t = started; t < parts; ++t) f(t, range_begin(n, parts, t), range_begin(n, parts, t + 1));
  for (std::thread& x : th) x.join();
}

// sort with a TOTAL order (no equal elements): chunks sorted concurrently, then merged pairwise - same result as
// std::sort for any thread count
template <typename It, typename Cmp>
inline void parallel_sort(It first, It last, Cmp cmp, size_t grain = (size_t)1 << 16) {
  const size_t n = (size_t)(last - first);
  int parts = range_count(n, grain);
  if (parts <= 1) { std::sort(first, last, cmp); return; }
  std::vector<size_t> cut(parts + 1);
  for (int t = 0; t <= parts; ++t) cut[t] = range_begin(n, parts, t);
  parallel_ranges((size_t)parts, parts, [&](int, size_t b, size_t e) { for (size_t t = b; t < e; ++t) std::sort(first + cut[t], first + cut[t + 1], cmp); });
  while (parts > 1) {
    const int pairs = parts / 2;
    parallel_ranges((size_t)pairs, pairs, [&](int, size_t b, size_t e) {
      for (size_t p = b; p < e; ++p) std::inplace_merge(first + cut[2 * p], first + cut[2 * p + 1], first + cut[2 * p + 2], cmp);
    });
    std::vector<size_t> next;
    for (int t = 0; t <= parts; t += 2) next.push_back(cut[t]);
    if (parts % 2) next.push_back(cut[parts]);
    cut.swap(next);
    parts = (int)cut.size() - 1;
  }
}

}  // namespace g2o_b200
