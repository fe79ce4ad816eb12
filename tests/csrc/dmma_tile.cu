// dmma_tile.cu - measurement tool (not part of the product): the 96 x 72 x 36 stage product of chol.cu's wide tiles in
// isolation (operands resident in shared memory, one barrier per stage, no copies) - what fraction of the DMMA peak the
// fragment-load pattern itself reaches.  Variants: 0 = as in chol.cu (compiler-scheduled), 1 = fragments of the next
// k-step loaded before the current one is multiplied (software pipelining in registers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_tile dmma_tile.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int LdA = 100, LdB = 76, KH = 36;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// V = 2 / 3: variant 0 / 1 plus the cp.async copies of the NEXT stage (two buffers) from a global panel with column stride Mk
template <int V>
__global__ void __launch_bounds__(384, 1) tile_kernel(double* out, int stages, long long* cyc, const double* __restrict__ G, int Mk, int nslabs, int noprod) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long mbar[2];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (V >= 2 ? 2 : 1) * (LdA + LdB) * KH; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, mg = warp & 3, ng = warp >> 2, g = lane >> 2, q = lane & 3;
  double c[3][3][2] = {};
  long long t0 = clock64();
  for (int st = 0; st < stages; ++st) {
    const int buf = (V >= 2) ? (st & 1) : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (V == 4) {
      // TMA: one warp issues 72 bulk copies (one per operand column segment), completion through an mbarrier per buffer
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar[buf]);
      if (st > 0) {
        unsigned done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"((unsigned)(((st - 1) >> 1) & 1)) : "memory");
      }
      __syncthreads();
      if (warp == 0) {
        const unsigned nbar = (unsigned)__cvta_generic_to_shared(&mbar[buf ^ 1]);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nbar), "r"((unsigned)((96 + 72) * KH * 8)) : "memory");
        __syncwarp();
        double* An = sm + (buf ^ 1) * (LdA + LdB) * KH; double* Bn = An + LdA * KH;
        const double* src0 = G + (size_t)((blockIdx.x * 7 + st) % nslabs) * 96;
        for (int seg = lane; seg < 2 * KH; seg += 32) {
          const bool isA = seg < KH; const int k = isA ? seg : seg - KH;
          const double* src = src0 + (size_t)k * Mk;
          double* dst = isA ? An + k * LdA : Bn + k * LdB;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"((unsigned)((isA ? 96 : 72) * 8)), "r"(nbar) : "memory");
        }
      }
    } else if (V == 5) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    } else if (V >= 2) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      double* An = sm + (buf ^ 1) * (LdA + LdB) * KH; double* Bn = An + LdA * KH;
      // a different 96-row slab of the panel per CTA and stage (L2 resident after the first pass)
      const double* src0 = G + (size_t)((blockIdx.x * 7 + st) % nslabs) * 96;
      for (int k = warp; k < KH; k += 12)
        for (int j = lane; j < 48 + 36; j += 32) {
          const bool isA = j < 48;
          const double* src = src0 + (size_t)k * Mk + (isA ? 2 * j : 2 * (j - 48));
          const double* dst = isA ? An + k * LdA + 2 * j : Bn + k * LdB + 2 * (j - 48);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const double* Ap = sm + buf * (LdA + LdB) * KH + q * LdA + mg * 24 + g;
    const double* Bp = sm + buf * (LdA + LdB) * KH + LdA * KH + q * LdB + ng * 24 + g;
    if (noprod) { if (V < 2) __syncthreads(); continue; }
    if (V == 5) {
      double* An = sm + (buf ^ 1) * (LdA + LdB) * KH; double* Bn = An + LdA * KH;
      const double* src0 = G + (size_t)((blockIdx.x * 7 + st) % nslabs) * 96;
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        {  // copy number s of this thread: column k = warp + 12 (s / 3), pair j = lane + 32 (s % 3)
          const int k = warp + 12 * (s / 3), j = lane + 32 * (s % 3);
          if (j < 48 + 36) {
            const bool isA = j < 48;
            const double* src = src0 + (size_t)k * Mk + (isA ? 2 * j : 2 * (j - 48));
            const double* dst = isA ? An + k * LdA + 2 * j : Bn + k * LdB + 2 * (j - 48);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
          }
        }
        const double a0 = Ap[0], a1 = Ap[8], a2 = Ap[16], b0 = Bp[0], b1 = Bp[8], b2 = Bp[16];
        dmma(c[0][0][0], c[0][0][1], a0, b0); dmma(c[1][0][0], c[1][0][1], a1, b0); dmma(c[2][0][0], c[2][0][1], a2, b0);
        dmma(c[0][1][0], c[0][1][1], a0, b1); dmma(c[1][1][0], c[1][1][1], a1, b1); dmma(c[2][1][0], c[2][1][1], a2, b1);
        dmma(c[0][2][0], c[0][2][1], a0, b2); dmma(c[1][2][0], c[1][2][1], a1, b2); dmma(c[2][2][0], c[2][2][1], a2, b2);
        Ap += 4 * LdA; Bp += 4 * LdB;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (V == 0 || V == 2 || V == 4) {
#pragma unroll 3
      for (int s = 0; s < 9; ++s) {
        const double a0 = Ap[0], a1 = Ap[8], a2 = Ap[16], b0 = Bp[0], b1 = Bp[8], b2 = Bp[16];
        dmma(c[0][0][0], c[0][0][1], a0, b0); dmma(c[1][0][0], c[1][0][1], a1, b0); dmma(c[2][0][0], c[2][0][1], a2, b0);
        dmma(c[0][1][0], c[0][1][1], a0, b1); dmma(c[1][1][0], c[1][1][1], a1, b1); dmma(c[2][1][0], c[2][1][1], a2, b1);
        dmma(c[0][2][0], c[0][2][1], a0, b2); dmma(c[1][2][0], c[1][2][1], a1, b2); dmma(c[2][2][0], c[2][2][1], a2, b2);
        Ap += 4 * LdA; Bp += 4 * LdB;
      }
    } else {
      double a0 = Ap[0], a1 = Ap[8], a2 = Ap[16], b0 = Bp[0], b1 = Bp[8], b2 = Bp[16];
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        double na0 = 0, na1 = 0, na2 = 0, nb0 = 0, nb1 = 0, nb2 = 0;
        if (s + 1 < 9) { Ap += 4 * LdA; Bp += 4 * LdB; na0 = Ap[0]; na1 = Ap[8]; na2 = Ap[16]; nb0 = Bp[0]; nb1 = Bp[8]; nb2 = Bp[16]; }
        dmma(c[0][0][0], c[0][0][1], a0, b0); dmma(c[1][0][0], c[1][0][1], a1, b0); dmma(c[2][0][0], c[2][0][1], a2, b0);
        dmma(c[0][1][0], c[0][1][1], a0, b1); dmma(c[1][1][0], c[1][1][1], a1, b1); dmma(c[2][1][0], c[2][1][1], a2, b1);
        dmma(c[0][2][0], c[0][2][1], a0, b2); dmma(c[1][2][0], c[1][2][1], a1, b2); dmma(c[2][2][0], c[2][2][1], a2, b2);
        a0 = na0; a1 = na1; a2 = na2; b0 = nb0; b1 = nb1; b2 = nb2;
      }
    }
    if (V < 2) __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (V == 4) {  // the last stage issued is still in flight
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar[stages & 1]);
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"((unsigned)(((stages - 1) >> 1) & 1)) : "memory");
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&cyc, 8);
  const int smem = 2 * (LdA + LdB) * KH * 8, stages = 2000;
  const int nslabs = 64;   // a 6144 x 36 panel (1.8 MB)
  double* G; cudaMalloc(&G, (size_t)6150 * KH * 8); cudaMemset(G, 0, (size_t)6150 * KH * 8);
  cudaFuncSetAttribute(tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("{");
  for (int cfg = 0; cfg < 4; cfg += 2) {
  const int noprod = cfg & 1; const int Mk = (cfg & 2) ? 6150 : 6144;
  printf("\"cfg_noprod%d_Mk%d\": {", noprod, Mk);
  for (int v = 0; v < 6; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) tile_kernel<0><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
      else if (v == 1) tile_kernel<1><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
      else if (v == 2) tile_kernel<2><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
      else if (v == 3) tile_kernel<3><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
      else if (v == 4) tile_kernel<4><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
      else tile_kernel<5><<<148, 384, smem>>>(out, stages, cyc, G, Mk, nslabs, noprod);
    }
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("\"variant%d_cycles_per_stage\": %.0f, ", v, (double)h / stages);
  }
  printf("\"_\": 0}, ");
  }
  printf("\"dmma_floor_cycles_per_stage\": 3888, \"cuda_error\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
