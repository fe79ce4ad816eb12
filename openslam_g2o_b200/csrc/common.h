// common.h - CUDA error handling and device buffers shared by the .cu translation units
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace g2o_b200 {

struct CudaError {
  cudaError_t code;
  const char* what;
  const char* file;
  int line;
};

#define B200_CUDA(expr)                                                         \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) throw ::g2o_b200::CudaError{_e, #expr, __FILE__, __LINE__}; \
  } while (0)

inline std::string describe(const CudaError& e) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e.code, cudaGetErrorString(e.code), e.file,
           e.line, e.what);
  return buf;
}

// Host-only contexts (b200_create(-1, ..)) run the integer structure phase without a device so that ordering /
// pattern / sharding logic can be tested on a CPU box; every compute entry point still fails loudly there.
inline bool& host_only_flag() {
  static thread_local bool f = false;
  return f;
}

// FNV-1a over every array a host-only context "uploads": a digest of the complete device-side plan of the structure
// phase (b200_debug_upload_digest), used to check that host-side optimisations leave the plan bit-identical
inline uint64_t& upload_digest() {
  static thread_local uint64_t h = 1469598103934665603ull;
  return h;
}
inline void digest_bytes(const void* data, size_t bytes) {
  uint64_t h = upload_digest();
  const unsigned char* p = static_cast<const unsigned char*>(data);
  // 8 bytes per step: enough for a regression digest, ~1 GB/s
  size_t i = 0;
  for (; i + 8 <= bytes; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 1099511628211ull; }
  for (; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
  upload_digest() = (h ^ bytes) * 1099511628211ull;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;    // elements in use
  size_t cap = 0;  // elements allocated
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cap = 0;
  }
  void alloc(size_t count) {
    if (host_only_flag()) { n = count; return; }
    if (count <= cap && p) { n = count; return; }
    release();
    const size_t c = count == 0 ? 1 : count;
    B200_CUDA(cudaMalloc((void**)&p, c * sizeof(T)));
    cap = c;
    n = count;
  }
  void upload(const std::vector<T>& h, cudaStream_t s) {
    alloc(h.size());
    if (host_only_flag()) { digest_bytes(h.data(), h.size() * sizeof(T)); return; }
    if (!h.empty()) B200_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    if (host_only_flag()) { digest_bytes(h, count * sizeof(T)); return; }
    if (count) B200_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) {
    if (host_only_flag()) return;
    if (p && n) B200_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// CUDA-event pairs around kernel groups, resolved lazily (no sync while recording).  ids: see solver.cu / g2o_b200.h
struct EventProfiler {
  bool on = false;
  cudaStream_t stream = nullptr;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<Rec> recs;
  double seconds[24] = {};
  long long count[24] = {};
  cudaEvent_t event() {
    if (used == pool.size()) {
      cudaEvent_t e;
      B200_CUDA(cudaEventCreate(&e));
      pool.push_back(e);
    }
    return pool[used++];
  }
  void flush() {
    if (recs.empty()) return;
    cudaStreamSynchronize(stream);
    for (const Rec& r : recs) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { seconds[r.id] += ms * 1e-3; count[r.id]++; }
    }
    recs.clear();
    used = 0;
  }
  void reset() { flush(); for (int i = 0; i < 24; ++i) { seconds[i] = 0; count[i] = 0; } }
  void destroy() { for (cudaEvent_t e : pool) cudaEventDestroy(e); pool.clear(); }
};
struct ScopedPhase {
  EventProfiler* p;
  int id;
  cudaEvent_t a = nullptr;
  ScopedPhase(EventProfiler* prof, int phase) : p(prof), id(phase) {
    if (p && p->on) { a = p->event(); cudaEventRecord(a, p->stream); }
  }
  ~ScopedPhase() {
    if (p && p->on && a) {
      cudaEvent_t b = p->event();
      cudaEventRecord(b, p->stream);
      p->recs.push_back({id, a, b});
    }
  }
};

// robust kernel applied to every edge (kernels.cuh: robustify)
struct Robust {
  int kind;      // B200_ROBUST_*: 0 none, 1 Huber, 2 PseudoHuber, 3 Cauchy, 4 Saturated, 5 DCS
  double delta;
  // per-edge kernels (device arrays in device edge order; nullptr: the uniform kernel above)
  const unsigned char* kinds;
  const double* deltas;
};

// counts every kernel launch of this library (reported as bench "gpu_launches")
struct LaunchCounter {
  int64_t n = 0;
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace g2o_b200
