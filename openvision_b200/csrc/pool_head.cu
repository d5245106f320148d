// ovk_pool_head: the tail of VisionTransformer.forward (open_clip/transformer.py:599-607 _global_pool, :638-640 ln_post
// after pooling, :645-646 pooled @ proj) and CLIP.encode_image's F.normalize (model.py:267) as ONE kernel.
//
// A CTA owns PH_IMGS images.  Phase 1 streams their tokens once (128-bit loads; 'avg': mean over tokens 1.., 'tok': token 0, 'last': token L-1)
// into fp32 rows in shared memory; phase 2 is the LayerNorm of those rows (one warp per image, two-pass statistics); phase 3
// multiplies by proj [D, E] on the FMA pipe: every thread owns two output columns and reuses each proj element for all of the
// CTA's images (proj is read once per CTA out of L2); phase 4 optionally L2-normalises the rows.  The token stream is the
// only HBM traffic that matters (B * L * D * 2 bytes): the kernel is bandwidth-bound, two CTAs per SM overlap one CTA's
// proj phase with the other's token stream.  Everything between the token load and the output store stays fp32.
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int PH_THREADS = 512;
constexpr int PH_IMGS = 4;

__device__ __forceinline__ void ph_unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = bf16_lo(w[q]);
    f[2 * q + 1] = bf16_hi(w[q]);
  }
}
__device__ __forceinline__ float ph_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool OUT_F32>
__global__ void __launch_bounds__(PH_THREADS, 2)
pool_head_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float ln_eps, const __nv_bfloat16* __restrict__ proj, void* __restrict__ out, int B, int L, int D, int E,
                 int mode, int normalize, float norm_eps, int scratch_floats) {
  extern __shared__ __align__(16) float ph_smem[];
  float* h = ph_smem;                      // [PH_IMGS][D] pooled (then normalised) rows
  float* scratch = ph_smem + PH_IMGS * D;  // phase 1: [TY][D] partial sums; phases 3-4: [PH_IMGS][E] outputs
  __shared__ float inv_norm[PH_IMGS];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * PH_IMGS;
  const int nimg = min(PH_IMGS, B - b0);
  const int nvec = D >> 3;
  const int TXN = min(nvec, PH_THREADS);          // column-vector lanes
  const int TY = max(1, PH_THREADS / TXN);        // token row groups
  const int tx = tid % TXN, ty = tid / TXN;
  const int l0 = mode == 0 ? 1 : (mode == 2 ? L - 1 : 0);
  const int l1 = mode == 1 ? 1 : L;
  const float pool_scale = 1.f / static_cast<float>(l1 - l0);

  // ---------------------------------------------------------------- phase 1: pooling
  for (int img = 0; img < nimg; ++img) {
    const __nv_bfloat16* xb = x + static_cast<long long>(b0 + img) * L * D;
    if (ty < TY) {
      for (int v = tx; v < nvec; v += TXN) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int l = l0 + ty;
        for (; l + 3 * TY < l1; l += 4 * TY) {   // four independent 16-byte loads in flight per thread
          uint4 r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(l + u * TY) * D) + v);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[8];
            ph_unpack8(r[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        }
        for (; l < l1; l += TY) {
          float f[8];
          ph_unpack8(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(l) * D) + v), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
        float4* dst = reinterpret_cast<float4*>(scratch + ty * D + v * 8);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    __syncthreads();
    for (int d = tid; d < D; d += PH_THREADS) {   // row groups summed in a fixed order (deterministic)
      float s = 0.f;
      for (int g = 0; g < TY; ++g) s += scratch[g * D + d];
      h[img * D + d] = s * pool_scale;
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- phase 2: LayerNorm of the pooled rows (ln_post)
  if (gamma != nullptr) {
    if (warp < nimg) {
      float* hr = h + warp * D;
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s += hr[d];
      const float mean = ph_warp_sum(s) / static_cast<float>(D);
      float q = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float c = hr[d] - mean;
        q = fmaf(c, c, q);
      }
      const float rstd = rsqrtf(ph_warp_sum(q) / static_cast<float>(D) + ln_eps);
      for (int d = lane; d < D; d += 32) hr[d] = fmaf((hr[d] - mean) * rstd, __ldg(gamma + d), __ldg(beta + d));
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- phase 3: @ proj (or pass-through)
  const int W = proj != nullptr ? E : D;   // output width
  float* o = scratch;                      // [PH_IMGS][W]
  if (proj != nullptr) {
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(proj);   // bf16 pairs, row d holds E / 2 of them
    const int ep2 = E >> 1;
    for (int ep = tid; ep < ep2; ep += PH_THREADS) {
      float acc[PH_IMGS][2];
#pragma unroll
      for (int i = 0; i < PH_IMGS; ++i) acc[i][0] = acc[i][1] = 0.f;
      int d = 0;
      for (; d + 4 <= D; d += 4) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = __ldg(pw + static_cast<long long>(d + u) * ep2 + ep);
#pragma unroll
        for (int i = 0; i < PH_IMGS; ++i) {
          const float4 hv = *reinterpret_cast<const float4*>(h + i * D + d);
          const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[i][0] = fmaf(hh[u], bf16_lo(w[u]), acc[i][0]);
            acc[i][1] = fmaf(hh[u], bf16_hi(w[u]), acc[i][1]);
          }
        }
      }
      for (; d < D; ++d) {
        const uint32_t w = __ldg(pw + static_cast<long long>(d) * ep2 + ep);
#pragma unroll
        for (int i = 0; i < PH_IMGS; ++i) {
          acc[i][0] = fmaf(h[i * D + d], bf16_lo(w), acc[i][0]);
          acc[i][1] = fmaf(h[i * D + d], bf16_hi(w), acc[i][1]);
        }
      }
#pragma unroll
      for (int i = 0; i < PH_IMGS; ++i) *reinterpret_cast<float2*>(o + i * W + 2 * ep) = make_float2(acc[i][0], acc[i][1]);
    }
  } else {
    for (int i = tid; i < nimg * D; i += PH_THREADS) o[i] = h[i];
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 4: F.normalize, store
  if (warp < PH_IMGS) {
    float inv = 1.f;
    if (normalize && warp < nimg) {
      float q = 0.f;
      for (int e = lane; e < W; e += 32) q = fmaf(o[warp * W + e], o[warp * W + e], q);
      inv = 1.f / fmaxf(sqrtf(ph_warp_sum(q)), norm_eps);
    }
    if (lane == 0) inv_norm[warp] = inv;
  }
  __syncthreads();
  for (int i = tid; i < nimg * (W >> 1); i += PH_THREADS) {
    const int img = i / (W >> 1), c = (i - img * (W >> 1)) * 2;
    const float s = inv_norm[img];
    const float a = o[img * W + c] * s, bq = o[img * W + c + 1] * s;
    const long long idx = static_cast<long long>(b0 + img) * W + c;
    if constexpr (OUT_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + idx) = make_float2(a, bq);
    else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + idx) = pack_bf16x2(a, bq);
  }
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_pool_head(const void* x, int B, int L, int D, int mode, const float* gamma, const float* beta,
                             float ln_eps, const void* proj, int E, int normalize, float norm_eps, void* out, int out_is_f32,
                             void* stream) {
  if (B <= 0 || L <= 0 || D <= 0 || (D % 8)) return set_error(OVK_ERR_SHAPE, "pool_head: bad shape B=%d L=%d D=%d", B, L, D);
  if (mode < 0 || mode > 2) return set_error(OVK_ERR_SHAPE, "pool_head: mode must be 0 (avg), 1 (tok) or 2 (last)");
  if (mode == 0 && L < 2) return set_error(OVK_ERR_SHAPE, "pool_head: avg pooling needs L >= 2");
  if ((gamma == nullptr) != (beta == nullptr)) return set_error(OVK_ERR_SHAPE, "pool_head: gamma and beta come together");
  if (proj != nullptr && (E <= 0 || (E % 8))) return set_error(OVK_ERR_SHAPE, "pool_head: E must be a positive multiple of 8");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (proj != nullptr && (reinterpret_cast<uintptr_t>(proj) & 3)))
    return set_error(OVK_ERR_ALIGN, "pool_head: x must be 16-byte aligned");
  const int W = proj != nullptr ? E : D;
  const int nvec = D / 8;
  const int TXN = nvec < PH_THREADS ? nvec : PH_THREADS;
  const int TY = PH_THREADS / TXN > 1 ? PH_THREADS / TXN : 1;
  const int scratch = TY * D > PH_IMGS * W ? TY * D : PH_IMGS * W;
  const size_t smem = static_cast<size_t>(PH_IMGS * D + scratch) * sizeof(float);
  if (smem > 100 * 1024) return set_error(OVK_ERR_SHAPE, "pool_head: D=%d / E=%d too wide for the shared-memory rows", D, E);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (B + PH_IMGS - 1) / PH_IMGS;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(pool_head): %s", cudaGetErrorString(e));
    kern<<<grid, PH_THREADS, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, ln_eps,
                                        reinterpret_cast<const __nv_bfloat16*>(proj), out, B, L, D, E, mode, normalize, norm_eps,
                                        scratch);
    return check_launch("pool_head_kernel");
  };
  return out_is_f32 ? launch(pool_head_kernel<true>) : launch(pool_head_kernel<false>);
}
