// Optimizer step of the reference's training recipe on flat parameter buffers (src/optim/build_optax.py:188-278 builds
//   clip_by_global_norm -> scale_by_adam(b1, b2, mu_dtype=bf16) -> add_decayed_weights(wd) -> scale(lr) -> schedule -> -1,
// configs/openvision.py:265-289: b1 = 0.9, b2 = 0.95, wd = 0.2 on the weight matrices, first moment kept in bf16):
//   g' = g * gscale                       (gscale = 1/world for the DP mean, times max_norm / max(||g||, max_norm))
//   mu = b1 mu + (1 - b1) g'              (used un-rounded for this step, stored as bf16)
//   nu = b2 nu + (1 - b2) g'^2
//   p  = p - lr * ( (mu / (1 - b1^t)) / (sqrt(nu / (1 - b2^t)) + eps) + wd * p )
// One pass over p, g, mu, nu: 128-bit vectorised, grid sized to the SM count, HBM-bound (fp32 p, g: 4+4+2+4 B read,
// 4+2+4 B written per element).
#include <cuda_bf16.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&f)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&f)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
}

// out += sum x^2 (fp32 accumulation per thread, block reduction, one atomicAdd per block)
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ x, long long n, float* __restrict__ out) {
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += 256LL * gridDim.x) {
    float f[4];
    load4<T>(x + 4 * i, f);
    acc = fmaf(f[0], f[0], acc); acc = fmaf(f[1], f[1], acc); acc = fmaf(f[2], f[2], acc); acc = fmaf(f[3], f[3], acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = static_cast<float>(x[4 * n4 + threadIdx.x]);
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    acc = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

struct AdamArgs {
  float lr, b1, b2, eps, wd, bc1, bc2, gscale, max_norm;
  const float* gnorm_sq;   // device scalar: sum of squares of the (unscaled) gradients of the whole model, or null
};

template <typename TP, typename TG>
__global__ void __launch_bounds__(256) adamw_kernel(TP* __restrict__ p, const TG* __restrict__ g, __nv_bfloat16* __restrict__ mu,
                                                    float* __restrict__ nu, long long n, const AdamArgs a) {
  float gscale = a.gscale;
  if (a.gnorm_sq != nullptr) {   // clip_by_global_norm on the scaled gradient
    const float norm = sqrtf(*a.gnorm_sq) * a.gscale;
    if (norm > a.max_norm) gscale *= a.max_norm / norm;
  }
  const float ib1 = 1.f / a.bc1, ib2 = 1.f / a.bc2;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += 256LL * gridDim.x) {
    float pf[4], gf[4], mf[4], nf[4];
    load4<TP>(p + 4 * i, pf);
    load4<TG>(g + 4 * i, gf);
    load4<__nv_bfloat16>(mu + 4 * i, mf);
    load4<float>(nu + 4 * i, nf);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gf[k] * gscale;
      mf[k] = fmaf(a.b1, mf[k], (1.f - a.b1) * gg);
      nf[k] = fmaf(a.b2, nf[k], (1.f - a.b2) * gg * gg);
      const float u = (mf[k] * ib1) / (sqrtf(nf[k] * ib2) + a.eps) + a.wd * pf[k];
      pf[k] = fmaf(-a.lr, u, pf[k]);
    }
    store4<TP>(p + 4 * i, pf);
    store4<__nv_bfloat16>(mu + 4 * i, mf);
    store4<float>(nu + 4 * i, nf);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // scalar tail
    const long long i = 4 * n4 + threadIdx.x;
    const float gg = static_cast<float>(g[i]) * gscale;
    const float m = fmaf(a.b1, __bfloat162float(mu[i]), (1.f - a.b1) * gg);
    const float v = fmaf(a.b2, nu[i], (1.f - a.b2) * gg * gg);
    const float pv = static_cast<float>(p[i]);
    const float u = (m * ib1) / (sqrtf(v * ib2) + a.eps) + a.wd * pv;
    p[i] = static_cast<TP>(fmaf(-a.lr, u, pv));
    mu[i] = __float2bfloat16_rn(m);
    nu[i] = v;
  }
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_sumsq(const void* x, int is_bf16, long long n, float* out, void* stream) {
  if (n <= 0) return set_error(OVK_ERR_SHAPE, "sumsq: empty input");
  if (reinterpret_cast<uintptr_t>(x) & 15) return set_error(OVK_ERR_ALIGN, "sumsq: x must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * num_sms();
  const int grid = static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  if (is_bf16) sumsq_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, out);
  else sumsq_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), n, out);
  return check_launch("sumsq_kernel");
}

extern "C" int ovk_adamw_step(void* p, int p_is_bf16, const void* g, int g_is_bf16, void* mu, float* nu, long long n, float lr,
                              float b1, float b2, float eps, float wd, int step, float gscale, const float* gnorm_sq,
                              float max_norm, void* stream) {
  if (n <= 0) return set_error(OVK_ERR_SHAPE, "adamw_step: empty parameter buffer");
  if (step < 1) return set_error(OVK_ERR_SHAPE, "adamw_step: step counts from 1");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mu) |
       reinterpret_cast<uintptr_t>(nu)) & 15)
    return set_error(OVK_ERR_ALIGN, "adamw_step: buffers must be 16-byte aligned");
  AdamArgs a;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd;
  a.bc1 = 1.f - powf(b1, static_cast<float>(step));
  a.bc2 = 1.f - powf(b2, static_cast<float>(step));
  a.gscale = gscale; a.gnorm_sq = gnorm_sq; a.max_norm = max_norm;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * num_sms();
  const int grid = static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  __nv_bfloat16* m = reinterpret_cast<__nv_bfloat16*>(mu);
  if (!p_is_bf16 && !g_is_bf16)
    adamw_kernel<float, float><<<grid, 256, 0, s>>>(reinterpret_cast<float*>(p), reinterpret_cast<const float*>(g), m, nu, n, a);
  else if (!p_is_bf16 && g_is_bf16)
    adamw_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<float*>(p), reinterpret_cast<const __nv_bfloat16*>(g), m, nu, n, a);
  else if (p_is_bf16 && !g_is_bf16)
    adamw_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(p), reinterpret_cast<const float*>(g), m, nu, n, a);
  else
    adamw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(p),
                                                                     reinterpret_cast<const __nv_bfloat16*>(g), m, nu, n, a);
  return check_launch("adamw_kernel");
}
