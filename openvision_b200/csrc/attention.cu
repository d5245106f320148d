// Fused multi-head self-attention forward on tcgen05: O = softmax(Q K^T * scale) V per (batch, head), flash-style.
// Replaces the core of nn.MultiheadAttention(need_weights=False, attn_mask=None) between in_proj and out_proj
// (open_clip/transformer.py:225,239-252; JAX twin src/models/common.py:53-200).  The online-softmax recurrence
// is the one stated in src/models/bpt.py:105-124 (running max m, running sum l, rescale of the accumulator).
//
// One CTA = one (128-query tile, head, batch).  S = Q K_j^T goes to TMEM (128 fp32 columns), the four softmax warps
// (thread <-> query row, i.e. TMEM lane) turn it into P (bf16) in shared memory in the K-major SWIZZLE_128B layout the
// second MMA wants, O += P V_j accumulates in TMEM (64 columns) with V consumed MN-major straight from the TMA tile,
// so neither K nor V is ever transposed in memory.  Q/K/V tiles are fetched with 4-D TMA boxes directly out of the
// packed in_proj output [B, L, 3, H, 64]; rows >= L are zero-filled by TMA and masked in the softmax.
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 192;  // warps 0-3 softmax, warp 4 TMA producer, warp 5 MMA issuer / TMEM owner
constexpr int ATT_TILE_BYTES = 128 * 128;  // [128 rows x 64 bf16]
constexpr int ATT_OFF_Q = 0;
constexpr int ATT_OFF_K = ATT_OFF_Q + ATT_TILE_BYTES;
constexpr int ATT_OFF_V = ATT_OFF_K + 2 * ATT_TILE_BYTES;
constexpr int ATT_OFF_P = ATT_OFF_V + 2 * ATT_TILE_BYTES;
constexpr int ATT_OFF_BAR = ATT_OFF_P + 2 * ATT_TILE_BYTES;
constexpr int ATT_NUM_BARS = 12;
constexpr int ATT_SMEM_BYTES = ATT_OFF_BAR + ATT_NUM_BARS * 8 + 16;
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O: [128,192)
constexpr uint32_t ATT_TMEM_S = 0;
constexpr uint32_t ATT_TMEM_O = 128;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                     float* __restrict__ lse_out, int L, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + ATT_OFF_Q;
  uint8_t* sK = smem + ATT_OFF_K;
  uint8_t* sV = smem + ATT_OFF_V;
  uint8_t* sP = smem + ATT_OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* pv_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ATT_OFF_BAR + ATT_NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (L + ATT_BKV - 1) / ATT_BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_4d(sQ, &tmQKV, q_full, 0, h, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 10);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_4d(sK + s * ATT_TILE_BYTES, &tmQKV, &k_full[s], 0, H + h, j * ATT_BKV, b);
        mbar_wait(&v_empty[s], ph ^ 1, 11);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_4d(sV + s * ATT_TILE_BYTES, &tmQKV, &v_full[s], 0, 2 * H + h, j * ATT_BKV, b);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      mbar_wait(q_full, 0, 12);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int valid = min(ATT_BKV, L - j * ATT_BKV);
        const int nblk = (valid + 15) & ~15;  // MMA N of S and K-extent of PV for this block
        const uint32_t k_addr = smem_u32(sK + s * ATT_TILE_BYTES);
        const uint32_t v_addr = smem_u32(sV + s * ATT_TILE_BYTES);
        // S = Q K_j^T
        mbar_wait(&k_full[s], ph, 13);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, nblk, 0, 0);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) {
          umma_bf16_ss(tmem_base + ATT_TMEM_S, umma_desc_kmajor_sw128(q_addr + k * 32),
                       umma_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
        // O += P V_j   (P: K-major [128 x nblk] in two 64-column swizzle atoms; V: MN-major [nblk x 64])
        mbar_wait(p_ready, j & 1, 14);
        mbar_wait(&v_full[s], ph, 15);
        tc_fence_after();
        constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_HD, 0, 1);
        const int ksteps = nblk / 16;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint32_t a = p_addr + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32;
          const uint32_t bb = v_addr + kk * 16 * 128;
          umma_bf16_ss(tmem_base + ATT_TMEM_O, umma_desc_kmajor_sw128(a), umma_desc_mnmajor_sw128(bb, ATT_TILE_BYTES),
                       idesc_pv, (j | kk) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warps: thread <-> query row
    const int r = threadIdx.x;  // 0..127 = row in tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY;  // running max (log2 domain, already scaled)
    float l = 0.f;        // running sum
    for (int j = 0; j < nkv; ++j) {
      const int valid = min(ATT_BKV, L - j * ATT_BKV);
      const int nblk = (valid + 15) & ~15;
      mbar_wait(s_full, j & 1, 16);
      tc_fence_after();
      // pass 1: row maximum over the valid columns
      float mx = -INFINITY;
      for (int c = 0; c < nblk; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_S + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m, mx * scale_log2);
      const float alpha = fast_exp2(m - m_new);  // first block: exp2(-inf) = 0
      // P buffer and the O accumulator are owned by PV(j-1) until it completes
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1, 17);
        tc_fence_after();
      }
      // pass 2: P = exp2(S*scale - m_new) -> bf16 -> swizzled smem; row sum in fp32
      float rowsum = 0.f;
      for (int c = 0; c < nblk; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_S + c, v);
        tmem_ld_wait();
        float p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e = fast_exp2(fmaf(__uint_as_float(v[i]), scale_log2, -m_new));
          p[i] = (c + i < valid) ? e : 0.f;
          rowsum += p[i];
        }
        uint8_t* atom = sP + (c >> 6) * ATT_TILE_BYTES;
        const uint32_t chunk = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(atom + sw128_offset(r, chunk)) =
            make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
        *reinterpret_cast<uint4*>(atom + sw128_offset(r, chunk + 1)) =
            make_uint4(pack_bf16x2(p[8], p[9]), pack_bf16x2(p[10], p[11]), pack_bf16x2(p[12], p[13]),
                       pack_bf16x2(p[14], p[15]));
      }
      l = l * alpha + rowsum;
      m = m_new;
      if (j > 0) {
        // rescale the running output: O *= alpha
        for (int c = 0; c < ATT_HD; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_O + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(tmem_base + t_lane + ATT_TMEM_O + c, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // -------------------------------------------------------------------- epilogue: O / l -> bf16 -> smem -> TMA store
    mbar_wait(pv_done, (nkv - 1) & 1, 18);
    tc_fence_after();
    const float inv_l = 1.f / l;
    uint8_t* sO = sQ;  // every MMA that read Q has completed (pv_done tracks all earlier MMAs of the issuing thread)
    for (int c = 0; c < ATT_HD; c += 16) {
      uint32_t o[16];
      tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_O + c, o);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(o[i]) * inv_l;
      const uint32_t chunk = c >> 3;
      *reinterpret_cast<uint4*>(sO + sw128_offset(r, chunk)) =
          make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      *reinterpret_cast<uint4*>(sO + sw128_offset(r, chunk + 1)) =
          make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                     pack_bf16x2(f[14], f[15]));
    }
    if (lse_out != nullptr && q0 + r < L) {
      lse_out[(static_cast<long long>(b) * H + h) * L + q0 + r] = (m + log2f(l)) * 0.69314718055994531f;
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_4d(&tmO, sO, 0, h, q0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale,
                                 void* stream) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(OVK_ERR_SHAPE, "attention: empty problem");
  if (hd != ATT_HD) return set_error(OVK_ERR_SHAPE, "attention: head dim %d not supported (this build: 64)", hd);
  if (B > 65535 || H > 65535) return set_error(OVK_ERR_SHAPE, "attention: B and H must be <= 65535");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmQKV, tmO;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {(uint32_t)hd, 1, ATT_BKV, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {(uint32_t)hd, 1, ATT_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((L + ATT_BQ - 1) / ATT_BQ, H, B);
  attention_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, s>>>(tmQKV, tmO, lse, L, H, scale * 1.4426950408889634f);
  return check_launch("attention_fwd_kernel");
}
