// ovk_pool_head: the tail of VisionTransformer.forward (open_clip/transformer.py:599-607 _global_pool, :638-640 ln_post
// after pooling, :645-646 pooled @ proj) and CLIP.encode_image's F.normalize (model.py:267) as ONE kernel.
//
// Persistent, one CTA (1024 threads) per SM; the images are dealt to the CTAs in balanced contiguous ranges (B = 1024 on 148
// SMs: 6 or 7 each) and processed up to PH_IMGS at a time.  Phase 1 streams their tokens once (128-bit loads; 'avg': mean over tokens 1.., 'tok': token 0, 'last': token L-1)
// into fp32 rows in shared memory; phase 2 is the LayerNorm of those rows (one warp per image, two-pass statistics); phase 3
// multiplies by proj [D, E] on the FMA pipe: every thread owns four output columns of a d-slice and reuses each proj element for all of the
// CTA's images (proj is read once per CTA and pass out of L2, eight rows in flight per thread); phase 4 optionally
// L2-normalises the rows.  The token stream is the only HBM traffic that matters (B * L * D * 2 bytes).  Everything between
// the token load and the output store stays fp32.
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int PH_THREADS = 1024;
constexpr int PH_IMGS = 8;      // images per pass of a CTA

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
__global__ void __launch_bounds__(PH_THREADS, 1)
pool_head_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float ln_eps, const __nv_bfloat16* __restrict__ proj, void* __restrict__ out, int B, int L, int D, int E,
                 int mode, int normalize, float norm_eps) {
  extern __shared__ __align__(16) float ph_smem[];
  float* h = ph_smem;                      // [PH_IMGS][D] pooled (then normalised) rows
  float* scratch = ph_smem + PH_IMGS * D;  // phase 1: [slots][D] partial sums; phases 3-4: [PH_IMGS][W] outputs
  __shared__ float inv_norm[PH_IMGS];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // images are dealt to the CTAs in contiguous, balanced ranges (sizes differ by at most one)
  const int per = B / gridDim.x, extra = B % gridDim.x;
  const int cta = blockIdx.x;
  const int img_begin = cta * per + min(cta, extra);
  const int img_end = img_begin + per + (cta < extra ? 1 : 0);
  const int nvec = D >> 3;
  // phase-1 geometry: an image row is covered by wpi warps (32 column vectors of 16 bytes each); the CTA's 32 warps form
  // 32 / wpi slots, each streaming every TYg-th token row of one image — no synchronisation until the whole pass is pooled
  const int wpi = (nvec + 31) >> 5;
  const int slots = 32 / wpi;
  const int l0 = mode == 0 ? 1 : (mode == 2 ? L - 1 : 0);
  const int l1 = mode == 1 ? 1 : L;
  const float pool_scale = 1.f / static_cast<float>(l1 - l0);
  const int W = proj != nullptr ? E : D;   // output width

  for (int b0 = img_begin; b0 < img_end;) {
    const int nimg = min(min(PH_IMGS, slots), img_end - b0);
    // ---------------------------------------------------------------- phase 1: pooling
    {
      const int TYg = slots / nimg;                 // row groups per image
      const int slot = warp / wpi;
      const int img = slot / TYg, ty = slot - img * TYg;
      const int v = (warp - slot * wpi) * 32 + lane;
      if (slot < nimg * TYg && v < nvec) {
        const __nv_bfloat16* xb = x + static_cast<long long>(b0 + img) * L * D;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int l = l0 + ty;
        for (; l + 7 * TYg < l1; l += 8 * TYg) {   // eight independent 16-byte loads in flight per thread: 128 KB per SM
          uint4 r[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(l + u * TYg) * D) + v);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float f[8];
            ph_unpack8(r[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        }
        for (; l < l1; l += TYg) {
          float f[8];
          ph_unpack8(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(l) * D) + v), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
        float4* dst = reinterpret_cast<float4*>(scratch + slot * D + v * 8);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      __syncthreads();
      for (int i = tid; i < nimg * D; i += PH_THREADS) {   // row groups summed in a fixed order (deterministic)
        const int im = i / D, d = i - im * D;
        float sum = 0.f;
        for (int g = 0; g < TYg; ++g) sum += scratch[(im * TYg + g) * D + d];
        h[i] = sum * pool_scale;
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
    float* o = scratch;                      // [PH_IMGS][W]
    if (proj != nullptr) {
      // thread (q, sl): four output columns 4q..4q+3 (one 8-byte load per proj row) over the d-slice sl; eight rows of
      // proj in flight per thread, ~1000 threads: proj (D x E bf16, L2-resident) streams at the SM's L2 bandwidth instead of
      // one load latency per row.  The slices add their partial sums into o one after the other (fixed order).
      const int NQ = E >> 2;
      const int DS = max(1, min(8, PH_THREADS / NQ));
      const int q = tid % NQ, sl = tid / NQ;
      const int dper = ((D + DS - 1) / DS + 7) & ~7;
      float acc[PH_IMGS][4];
#pragma unroll
      for (int i = 0; i < PH_IMGS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const bool active = sl < DS;
      if (active) {
        const int d0 = sl * dper, d1 = min(D, d0 + dper);
        const uint2* pw = reinterpret_cast<const uint2*>(proj) + q;   // row d at pw + d * NQ
        int d = d0;
        for (; d + 8 <= d1; d += 8) {
          uint2 w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __ldg(pw + static_cast<long long>(d + u) * NQ);
#pragma unroll
          for (int i = 0; i < PH_IMGS; ++i) {
            if (i < nimg) {
              const float4 ha = *reinterpret_cast<const float4*>(h + i * D + d);
              const float4 hb = *reinterpret_cast<const float4*>(h + i * D + d + 4);
              const float hh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                acc[i][0] = fmaf(hh[u], bf16_lo(w[u].x), acc[i][0]);
                acc[i][1] = fmaf(hh[u], bf16_hi(w[u].x), acc[i][1]);
                acc[i][2] = fmaf(hh[u], bf16_lo(w[u].y), acc[i][2]);
                acc[i][3] = fmaf(hh[u], bf16_hi(w[u].y), acc[i][3]);
              }
            }
          }
        }
        for (; d < d1; ++d) {
          const uint2 w = __ldg(pw + static_cast<long long>(d) * NQ);
#pragma unroll
          for (int i = 0; i < PH_IMGS; ++i) {
            const float hv = h[i * D + d];
            acc[i][0] = fmaf(hv, bf16_lo(w.x), acc[i][0]);
            acc[i][1] = fmaf(hv, bf16_hi(w.x), acc[i][1]);
            acc[i][2] = fmaf(hv, bf16_lo(w.y), acc[i][2]);
            acc[i][3] = fmaf(hv, bf16_hi(w.y), acc[i][3]);
          }
        }
      }
      for (int turn = 0; turn < DS; ++turn) {
        if (active && sl == turn) {
#pragma unroll
          for (int i = 0; i < PH_IMGS; ++i) {
            float4* dst = reinterpret_cast<float4*>(o + i * W + 4 * q);
            float4 cur = turn == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
            cur.x += acc[i][0];
            cur.y += acc[i][1];
            cur.z += acc[i][2];
            cur.w += acc[i][3];
            *dst = cur;
          }
        }
        __syncthreads();
      }
    } else {
      for (int i = tid; i < nimg * D; i += PH_THREADS) o[i] = h[i];
      __syncthreads();
    }

    // ---------------------------------------------------------------- phase 4: F.normalize, store
    if (warp < PH_IMGS) {
      float inv = 1.f;
      if (normalize && warp < nimg) {
        float qq = 0.f;
        for (int e = lane; e < W; e += 32) qq = fmaf(o[warp * W + e], o[warp * W + e], qq);
        inv = 1.f / fmaxf(sqrtf(ph_warp_sum(qq)), norm_eps);
      }
      if (lane == 0) inv_norm[warp] = inv;
    }
    __syncthreads();
    for (int i = tid; i < nimg * (W >> 1); i += PH_THREADS) {
      const int img = i / (W >> 1), c = (i - img * (W >> 1)) * 2;
      const float sc = inv_norm[img];
      const float a = o[img * W + c] * sc, bq = o[img * W + c + 1] * sc;
      const long long idx = static_cast<long long>(b0 + img) * W + c;
      if constexpr (OUT_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + idx) = make_float2(a, bq);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + idx) = pack_bf16x2(a, bq);
    }
    __syncthreads();   // h / scratch are rewritten by the next pass
    b0 += nimg;
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
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (proj != nullptr && (reinterpret_cast<uintptr_t>(proj) & 7)))
    return set_error(OVK_ERR_ALIGN, "pool_head: x must be 16-byte and proj 8-byte aligned");
  const int W = proj != nullptr ? E : D;
  const int nvec = D / 8;
  const int wpi = (nvec + 31) / 32;
  if (wpi > 32) return set_error(OVK_ERR_SHAPE, "pool_head: D=%d too wide", D);
  const int slots = 32 / wpi;
  const int scratch = slots * D > PH_IMGS * W ? slots * D : PH_IMGS * W;
  const size_t smem = static_cast<size_t>(PH_IMGS * D + scratch) * sizeof(float);
  if (smem > 200 * 1024) return set_error(OVK_ERR_SHAPE, "pool_head: D=%d / E=%d too wide for the shared-memory rows", D, E);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // persistent: one CTA per SM, images dealt out in balanced contiguous ranges
  const int grid = B < num_sms() ? B : num_sms();
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(pool_head): %s", cudaGetErrorString(e));
    kern<<<grid, PH_THREADS, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, ln_eps,
                                        reinterpret_cast<const __nv_bfloat16*>(proj), out, B, L, D, E, mode, normalize, norm_eps);
    return check_launch("pool_head_kernel");
  };
  return out_is_f32 ? launch(pool_head_kernel<true>) : launch(pool_head_kernel<false>);
}
