// Throughput probe: ex2.approx f32 vs f16x2 vs bf16x2 (elements per clock per SM), cvt packs, and the FMA-pipe polynomial.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; h[i] = 0x3c003c00u + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) { float y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i+1)&7])); a[i] += __uint_as_float(h[i] << 16); }
      if (MODE == 4) {  // polynomial exp2 on the FMA pipe: floor via magic add, degree-3 poly, exponent add
        float x = a[i];
        float fl = x + 12582912.f;            // round to nearest int (magic 1.5 * 2^23)
        float n = fl - 12582912.f;
        float f = x - n;
        float p = fmaf(f, 0.0555041f, 0.2402265f);
        p = fmaf(p, f, 0.6931472f);
        p = fmaf(p, f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23)) * 1e-3f - 0.5f;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int elems_per_op) {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 2, 512>>>(out, 10, 1.f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = 148.0 * 2 * 512 * 8.0 * iters;
  printf("%-28s %.3f ms  %.2f warp-ops/ns  = %.1f elems/clk/SM at %.2f GHz (max clock)\n", name, ms, ops / 32 / (ms * 1e6),
         ops * elems_per_op / (ms * 1e-3) / 148 / (clk * 1e3), clk * 1e-6);
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.rn.bf16x2.f32 (+fadd)", 2);
  run<4>("poly exp2 (FMA pipe)", 1);
  return 0;
}
