// fp64_pipes.cu - measured FP64 throughput of this GPU's two FP64 paths (test/measurement tool, not part of the product):
//   dfma  : plain DFMA, 8 independent accumulators per thread
//   dmma  : mma.sync.aligned.m8n8k4.f64 (FP64 tensor path, DMMA in SASS), 8 independent accumulator tiles per warp
// and the dependent-issue latency of DFMA / MUFU.RSQ64H (what bounds the pivot chain of the sparse factorisation).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu ; prints one JSON object.
#include <cuda_runtime.h>

#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
  for (int i = 0; i < iters; ++i) {
    c0 = fma(a, c0, b); c1 = fma(a, c1, b); c2 = fma(a, c2, b); c3 = fma(a, c3, b);
    c4 = fma(a, c4, b); c5 = fma(a, c5, b); c6 = fma(a, c6, b); c7 = fma(a, c7, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c[t][0] = threadIdx.x + t; c[t][1] = threadIdx.x - t; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one warp, one dependent chain: cycles per instruction
__global__ void latency_kernel(double* out, long long* cyc, int iters, double a, double b) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  double y = b + 2.0;
  for (int i = 0; i < iters; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y)); y = r + 1.5; }
  long long t2 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaMalloc(&cyc, 2 * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 16;
  double best[2] = {0, 0};
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 4; ++rep) {
      const int blocks = sms * 4, threads = 512;
      cudaEventRecord(e0);
      if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 0.999, 1e-3);
      else dmma_kernel<<<blocks, threads>>>(out, iters, 0.999, 1e-3);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      // dfma: 8 FMA per thread per iteration; dmma: 8 tiles x (8*8*4 FMA) per warp per iteration
      const double fma_count = which == 0 ? (double)blocks * threads * 8.0 * iters : (double)blocks * (threads / 32) * 8.0 * 256.0 * iters;
      const double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
      if (tf > best[which]) best[which] = tf;
    }
  latency_kernel<<<1, 32>>>(out, cyc, 4096, 0.999, 1e-3);
  long long h[2] = {0, 0};
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t err = cudaDeviceSynchronize();
  printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_dependent_cycles\": %.1f, "
         "\"rsqrt64h_plus_dadd_dependent_cycles\": %.1f, \"cuda_error\": \"%s\"}\n",
         sms, khz / 1e3, best[0], best[1], h[0] / 4096.0, h[1] / 4096.0, cudaGetErrorString(err));
  return 0;
}
