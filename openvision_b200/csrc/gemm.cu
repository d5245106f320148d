// ovk_gemm_bf16: C[M,N] (bf16) = epilogue( A[M,K] * B[N,K]^T ) on tcgen05 tensor cores.
// Replaces the reference's F.linear / addmm call sites (open_clip/transformer.py:225,233-235,250-252 via
// nn.MultiheadAttention in_proj/out_proj and mlp.c_fc/c_proj) with fused bias / exact-erf GELU / residual epilogues.
#include "gemm_core.cuh"
#include "host_utils.h"

namespace ovk {

// gelu(x) = x * Phi(x), exact-erf form (nn.GELU() default, transformer.py:234; JAX approximate=False vit.py:198-202).
// erfc via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7), branch-free, 2 MUFU + ~14 FMA-pipe ops.
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, u, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = fast_exp2(x * x * -0.72134752044448170f);  // exp(-x^2/2)
  const float q = 0.5f * p * e;                                // = 0.5*erfc(|x|/sqrt2) = Phi(-|x|)
  const float xq = x * q;
  return x >= 0.f ? x - xq : xq;
}
// QuickGELU: x * sigmoid(1.702 x) (transformer.py:33-36), only when quick_gelu=True.
__device__ __forceinline__ float gelu_quick(float x) { return x * __frcp_rn(1.0f + fast_exp2(x * -2.4554669595930157f)); }
// tanh-approximate GELU (text tower act_kwargs={'approximate':'tanh'}).
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  // tanh(k) = 1 - 2/(1+exp(2k))
  const float th = 1.0f - 2.0f * __frcp_rn(1.0f + fast_exp2(k * 2.8853900817779268f));
  return 0.5f * x * (1.0f + th);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias,
                 const __nv_bfloat16* residual, long long ldr, int M, int N, int K, int flags) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using L = GemmSmemLayout<BN>;
  GemmCtx<BN, L> cx(smem_raw);
  const uint32_t tmem_base = gemm_prologue(cx, &tmA, &tmB, &tmC);
  const int warp = threadIdx.x >> 5;

  if (warp == 0) {
    if (elect_one()) gemm_producer(cx, &tmA, &tmB, M, N, K);
  } else if (warp == 1) {
    if (elect_one()) gemm_mma_issuer(cx, tmem_base, M, N, K);
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> smem -> TMA store
    const int ew = warp - 4;            // TMEM lane quadrant
    const int et = threadIdx.x - 128;   // 0..127 = row inside the tile
    const uint32_t lane = lane_id();
    float* bias_s = reinterpret_cast<float*>(cx.epi_scratch());
    GemmSched sched(M, N, BN);
    const bool has_bias = (flags & OVK_EPI_BIAS) != 0;
    const bool has_res = (flags & OVK_EPI_RESIDUAL) != 0;
    const int act = flags & OVK_EPI_ACT_MASK;
    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int t = blockIdx.x; t < sched.total; t += gridDim.x, ++it) {
      const GemmTileInfo ti = sched.tile(t, BN);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // bias tile -> smem (visible after the barrier below)
      for (int j = et; j < BN; j += GEMM_EPI_THREADS) {
        const int n = ti.n0 + j;
        bias_s[j] = (has_bias && n < N) ? bias[n] : 0.f;
      }
      named_bar_sync(1, GEMM_EPI_THREADS);
      mbar_wait(&cx.tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const int row = ti.m0 + et;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c, ++chunk_ctr) {
        const int ncol0 = ti.n0 + c * 64;
        if (ncol0 >= N) {  // whole chunk out of range (uniform across the CTA): still release TMEM on the last chunk
          if (c == BN / 64 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&cx.tmem_empty[acc]);
          }
          continue;
        }
        uint4 res[8];
        if (has_res) {
          const bool rok = row < M;
          const uint4* rp = reinterpret_cast<const uint4*>(residual + static_cast<long long>(row) * ldr + ncol0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            res[j] = (rok && ncol0 + j * 8 < N) ? rp[j] : make_uint4(0, 0, 0, 0);
          }
        }
        uint32_t v[64];
        {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
          tmem_ld_x32(taddr + c * 64, v0);
          tmem_ld_x32(taddr + c * 64 + 32, v1);
          tmem_ld_wait();
        }
        if (c == BN / 64 - 1) {  // accumulator fully drained into registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cx.tmem_empty[acc]);
        }
        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x0 = __uint_as_float(v[2 * j]) + bias_s[c * 64 + 2 * j];
          float x1 = __uint_as_float(v[2 * j + 1]) + bias_s[c * 64 + 2 * j + 1];
          if (act == OVK_EPI_GELU_ERF) {
            x0 = gelu_erf(x0);
            x1 = gelu_erf(x1);
          } else if (act == OVK_EPI_GELU_TANH) {
            x0 = gelu_tanh(x0);
            x1 = gelu_tanh(x1);
          } else if (act == OVK_EPI_GELU_QUICK) {
            x0 = gelu_quick(x0);
            x1 = gelu_quick(x1);
          }
          if (has_res) {
            const uint32_t rv = reinterpret_cast<const uint32_t*>(res)[j];
            x0 += bf16_lo(rv);
            x1 += bf16_hi(rv);
          }
          packed[j] = pack_bf16x2(x0, x1);
        }
        // staging buffer (c & 1): wait until the TMA store issued two chunks ago has finished reading it
        uint8_t* sbuf = cx.c_stage(chunk_ctr & 1);
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, GEMM_EPI_THREADS);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint4*>(sbuf + sw128_offset(et, j)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, GEMM_EPI_THREADS);
        if (et == 0) {
          tma_store_2d(&tmC, sbuf, ncol0, ti.m0);
          tma_store_commit();
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }
  gemm_teardown(cx, tmem_base);
}

template <int BN>
static int launch_gemm(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                       int K, const float* bias, const void* residual, long long ldr, int flags, cudaStream_t stream) {
  using L = GemmSmemLayout<BN>;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, GEMM_BK, GEMM_BM))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, GEMM_BK, BN))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmC, C, N, M, ldc, 64, GEMM_BM))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, L::DYN_BYTES, stream>>>(
      tmA, tmB, tmC, bias, reinterpret_cast<const __nv_bfloat16*>(residual), ldr, M, N, K, flags);
  return check_launch("gemm_bf16_kernel");
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M,
                             int N, int K, const float* bias, const void* residual, long long ldr, int flags,
                             void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OVK_ERR_SHAPE, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (K % 8) || (N % 8))
    return set_error(OVK_ERR_ALIGN, "gemm: lda/ldb/ldc/K/N must be multiples of 8 elements (16 B TMA strides)");
  if ((flags & OVK_EPI_RESIDUAL) && (residual == nullptr || (ldr % 8)))
    return set_error(OVK_ERR_ALIGN, "gemm: residual epilogue needs a pointer and ldr %% 8 == 0");
  if ((flags & OVK_EPI_BIAS) && bias == nullptr) return set_error(OVK_ERR_SHAPE, "gemm: bias flag without pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (N <= 128) return launch_gemm<128>(A, lda, B, ldb, C, ldc, M, N, K, bias, residual, ldr, flags, s);
  return launch_gemm<256>(A, lda, B, ldb, C, ldc, M, N, K, bias, residual, ldr, flags, s);
}
