// Weight packing for LayerNorm folded into the projection that consumes it (transformer.py:254-265: attn(ln_1(x)),
// mlp(ln_2(x)); LayerNorm: transformer.py:15-30).
//
//   ln(x) W^T + b = rstd * ((x - mu) (W.gamma)^T) + (W beta + b)
//                 = rstd * (x Wc^T) + d        with  Wc[n][k] = W[n][k] gamma[k] - mean_k(W[n][.] gamma[.]),  d = W beta + b
//
// because sum_k Wc[n][k] = 0 makes x Wc^T blind to the row mean of x: the GEMM reads the un-normalised residual stream,
// the mean never has to be subtracted, and the epilogue only scales by the row's rstd (ovk_gemm_bf16_ln).
//
// Wc is stored in bf16, and rounding breaks sum_k Wc = 0 by r_n ~ 2^-9 |w| sqrt(K), which would leak mu * rstd * r_n
// into the output.  After round-to-nearest the kernel therefore walks each row once more and moves a few entries to the
// ADJACENT bf16 value in the direction that cancels the residual (only entries whose rounding error already pointed
// that way by more than a quarter ulp, so no entry ends up further than 3/4 ulp from its exact value), until
// |sum_k Wc_bf16| is below half an ulp of the entries.  The leak drops to ~1e-4 * |mu| / sigma of the output scale.
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

template <typename TW>
__device__ __forceinline__ float ldw(const TW* p);
template <>
__device__ __forceinline__ float ldw<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldw<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }

// one warp per weight row
template <typename TW>
__global__ void __launch_bounds__(128) ln_pack_kernel(const TW* __restrict__ W, long long ldw_, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                                      __nv_bfloat16* __restrict__ Wc, long long ldwc, float* __restrict__ d,
                                                      int N, int K) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const TW* w = W + static_cast<long long>(n) * ldw_;
  unsigned short* q = reinterpret_cast<unsigned short*>(Wc + static_cast<long long>(n) * ldwc);
  // pass 1: row mean of W.gamma and d = W beta + b
  float sg = 0.f, sb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = ldw<TW>(w + k);
    sg = fmaf(wv, gamma[k], sg);
    sb = fmaf(wv, beta[k], sb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  const float mean = sg / static_cast<float>(K);
  if (lane == 0) d[n] = sb + (bias != nullptr ? bias[n] : 0.f);
  // pass 2: round to nearest, residual of the rounded row
  float r = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = fmaf(ldw<TW>(w + k), gamma[k], -mean);
    const unsigned short b = static_cast<unsigned short>(pack_bf16x2(v, 0.f) & 0xFFFFu);
    q[k] = b;
    r += bf16_bits_to_float(b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  __syncwarp();
  // pass 3: cancel the residual with one-ulp moves, entries visited in ascending k exactly as a serial walk would.  Each
  // lane prepares its entry of a 32-entry chunk for both directions (coalesced loads, no dependence on r); the warp then
  // steps through the chunk's candidates in lane order with the running residual r kept identical in every lane, so the
  // only serial part is two shuffles per candidate instead of a dependent global load per entry.
  for (int k0 = 0; k0 < K && r != 0.f; k0 += 32) {
    const int k = k0 + lane;
    uint32_t b = 0;
    float delta_dn = 0.f, delta_up = 0.f;
    bool ok_dn = false, ok_up = false;
    if (k < K) {
      b = q[k];
      if (!((b & 0x7FFFu) == 0 || (b & 0x7F80u) == 0x7F80u)) {   // zero, inf / nan: leave alone
        const float qv = bf16_bits_to_float(b);
        const float v = fmaf(ldw<TW>(w + k), gamma[k], -mean);
        const float e = v - qv;                                     // rounding error of this entry
        // the adjacent value towards -inf (used when r > 0) and towards +inf (r < 0)
        const uint32_t nb_dn = (qv > 0.f) ? b - 1 : b + 1;
        const uint32_t nb_up = (qv > 0.f) ? b + 1 : b - 1;
        if ((nb_dn & 0x7F80u) != 0x7F80u) {
          delta_dn = bf16_bits_to_float(nb_dn) - qv;
          ok_dn = !(e * delta_dn < 0.25f * delta_dn * delta_dn);    // only entries already off in that direction
        }
        if ((nb_up & 0x7F80u) != 0x7F80u) {
          delta_up = bf16_bits_to_float(nb_up) - qv;
          ok_up = !(e * delta_up < 0.25f * delta_up * delta_up);
        }
      }
    }
    // candidates that can still move r towards zero, re-evaluated after every move (r changes sign / size); entries at or
    // below the last visited lane are not revisited — the same decisions a serial walk over k makes, at one ballot + one
    // shuffle per APPLIED move instead of a step per entry
    unsigned visited = 0u;
    while (r != 0.f) {
      const bool down = r > 0.f;
      const float dsel = down ? delta_dn : delta_up;
      const bool osel = down ? ok_dn : ok_up;
      const unsigned cand = __ballot_sync(0xffffffffu, osel && fabsf(dsel) < 2.f * fabsf(r)) & ~visited;
      if (cand == 0u) break;
      const int j = __ffs(cand) - 1;
      const float dj = __shfl_sync(0xffffffffu, dsel, j);
      if (lane == j) {
        const float qv = bf16_bits_to_float(b);
        const uint32_t nb = ((qv > 0.f) == down) ? b - 1 : b + 1;
        q[k] = static_cast<unsigned short>(nb);
      }
      r += dj;
      visited |= (2u << j) - 1u;   // lanes 0..j
    }
  }
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_pack_ln_linear(const void* W, int w_is_f32, long long ldw, const float* gamma, const float* beta,
                                  const float* bias, void* Wc, long long ldwc, float* d, int N, int K, void* stream) {
  if (N <= 0 || K <= 0) return set_error(OVK_ERR_SHAPE, "pack_ln_linear: empty weight");
  if (W == nullptr || gamma == nullptr || beta == nullptr || Wc == nullptr || d == nullptr)
    return set_error(OVK_ERR_SHAPE, "pack_ln_linear: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (N + 3) / 4;
  if (w_is_f32)
    ln_pack_kernel<float><<<grid, 128, 0, s>>>(reinterpret_cast<const float*>(W), ldw, gamma, beta, bias,
                                               reinterpret_cast<__nv_bfloat16*>(Wc), ldwc, d, N, K);
  else
    ln_pack_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(W), ldw, gamma, beta, bias,
                                                       reinterpret_cast<__nv_bfloat16*>(Wc), ldwc, d, N, K);
  return check_launch("ln_pack_kernel");
}
