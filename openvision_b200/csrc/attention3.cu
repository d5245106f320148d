// Multi-head self-attention forward, head width 64, third-generation kernel: O = softmax(Q K^T * scale) V.
// Replaces the core of nn.MultiheadAttention(need_weights=False, attn_mask=None) between in_proj and out_proj
// (open_clip/transformer.py:225,239-252); online softmax as in src/models/bpt.py:105-124.
//
// Same decomposition as attention2.cu (one persistent CTA per SM, a work item = a PAIR of 128-row query tiles of one
// (head, image), K / V fetched once for both, P kept in TMEM as the A operand of P V, remainder token of L = 128 k + 1
// folded in through 16-column accumulators), with the two changes the round-1 profile asked for
// (profiles/r01_attention2_ncu_full.csv: softmax warps 29 % of their time waiting for S / O, exponentials issued at half
// the MUFU rate because every pair was followed by its dependent adds):
//   * the key dimension advances in HALF blocks of 64 keys and the two halves of a tile's 128 score columns are two
//     buffers: S(hb+1) is already in TMEM while the softmax group works on S(hb), and S(hb+2) is issued right behind
//     P V(hb).  A group never waits for the tensor pipe inside an item; the only ordering left is "P V(hb-1) has landed"
//     before the (rare) rescaling of O, signalled by a commit barrier behind every P V.
//   * the exponential pass is software-pipelined by hand: the 16 exponentials of chunk c are issued back to back, and the
//     sums / bf16 packing consume chunk c-1, so no instruction waits on a MUFU result that was issued just before it.
//   warps 0-3  : softmax group A — thread <-> query row of tile A       warp 8 : TMA producer
//   warps 4-7  : softmax group B — the same for tile B                   warp 9 : MMA issuer
// TMEM columns: S_A [0,128) = two 64-column buffers, S_B [128,256), O_A [256,320), O_B [320,384), remainder-token
// accumulators [384,464) as in attention2.cu.
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int A3_BQ = 128;
constexpr int A3_BKV = 128;           // keys per K / V shared-memory tile (one TMA box)
constexpr int A3_HB = 64;             // keys per softmax step (half a tile)
constexpr int A3_HD = 64;
constexpr int A3_THREADS = 320;
constexpr int A3_TILE = 128 * 128;    // [128 rows x 64 bf16], SWIZZLE_128B
constexpr int A3_HALF = 64 * 128;     // byte offset of rows 64.. inside a tile
constexpr int A3_NS = 3;              // K / V ring depth (tiles)
constexpr int A3_OFF_Q = 0;                               // [2 item buffers][2 tiles]
constexpr int A3_OFF_V = A3_OFF_Q + 4 * A3_TILE;          // V before K (the O_t^T MMA over-reads one tile past V)
constexpr int A3_OFF_K = A3_OFF_V + A3_NS * A3_TILE;
constexpr int A3_OFF_ROWS = A3_OFF_K + A3_NS * A3_TILE;   // per item buffer: k_t, v_t, q_t rows (128 B each)
constexpr int A3_OFF_PT = A3_OFF_ROWS + 2 * 384;          // p_t: bf16 [128] (+128 B that the MMA over-reads)
constexpr int A3_OFF_RED = A3_OFF_PT + 384;               // remainder-row scratch: [2][4] max, [4] sum, s_tt
constexpr int A3_OFF_BAR = A3_OFF_RED + 64;
constexpr int A3_B_QFULL = 0;     // [2]
constexpr int A3_B_QFREE = 2;     // [2]
constexpr int A3_B_KFULL = 4;     // [NS]
constexpr int A3_B_KEMPTY = 7;    // [NS]
constexpr int A3_B_VFULL = 10;    // [NS]
constexpr int A3_B_VEMPTY = 13;   // [NS]
constexpr int A3_B_SFULL = 16;    // [2 tiles][2 buffers]
// p_ready and pv_done alternate between two barriers per tile slot (even / odd half block): a softmax group may run one
// half block ahead of the MMA warp, and a parity wait cannot tell "two phases ahead" from "not yet"
constexpr int A3_B_PREADY = 20;   // [2 tiles][2]
constexpr int A3_B_PVDONE = 24;   // [2 tiles][2]  one arrival behind every P V of the tile slot
constexpr int A3_NUM_BARS = 28;
constexpr int A3_SMEM = A3_OFF_BAR + A3_NUM_BARS * 8 + 16;
constexpr uint32_t A3_T_S = 0, A3_T_O = 256, A3_T_SK = 384, A3_T_ST = 416, A3_T_OT = 432, A3_T_TT = 448;

__device__ __forceinline__ void umma_bf16_ts3(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// 16 exponentials p = 2^(s * scale - m) of one chunk, issued back to back
__device__ __forceinline__ void exp_chunk(const uint32_t (&x)[16], float scale_log2, float neg_m, float (&e)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = fast_exp2(fmaf(__uint_as_float(x[i]), scale_log2, neg_m));
}
// consume a chunk of exponentials: row-sum contribution and the 8 packed bf16 pairs
__device__ __forceinline__ float pack_chunk(const float (&e)[16], uint32_t (&pw)[8]) {
  float s0 = e[0] + e[1], s1 = e[2] + e[3], s2 = e[4] + e[5], s3 = e[6] + e[7];
  float s4 = e[8] + e[9], s5 = e[10] + e[11], s6 = e[12] + e[13], s7 = e[14] + e[15];
#pragma unroll
  for (int i = 0; i < 8; ++i) pw[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
  return ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// latency-critical waits (softmax group <-> MMA warp hand-offs): SPIN polls with test_wait instead of suspending the thread
// in hardware (the suspended form wakes up late; the polling form costs issue slots of the scheduler the warp lives on)
template <bool SPIN>
__device__ __forceinline__ void mbar_wait_x(uint64_t* bar, uint32_t parity, int code) {
  if constexpr (!SPIN) {
    mbar_wait(bar, parity, code);
  } else {
    uint32_t polls = 0;
    while (!mbar_test_wait(bar, parity)) {
      if (++polls > (1u << 27)) hang_trap(code);
    }
  }
}

// DUAL: one MMA-issuing warp per tile slot (warps 9 and 10) instead of one warp alternating between the two slots, so that
// neither softmax group ever waits for the other one's hand-off; K / V ring slots are then released by two arrivals.
template <bool DUAL, bool SPIN>
__global__ void __launch_bounds__(A3_THREADS + 32, 1)
attention_fwd3_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmRow, float* __restrict__ lse_out,
                      __nv_bfloat16* __restrict__ out, int L, int Lm, int H, int nq, int total_items, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention3: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A3_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + A3_OFF_BAR + A3_NUM_BARS * 8);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int npair = (nq + 1) >> 1;
  const int nkv = (Lm + A3_BKV - 1) / A3_BKV;
  const int nhb = (Lm + A3_HB - 1) / A3_HB;
  const bool tail = L > Lm;   // one remainder token (host guarantees L - Lm <= 1, and then Lm % 128 == 0)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmRow);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[A3_B_QFULL + i], 1);
      mbar_init(&bars[A3_B_QFREE + i], 2);
      mbar_init(&bars[A3_B_SFULL + 2 * i], 1);
      mbar_init(&bars[A3_B_SFULL + 2 * i + 1], 1);
      mbar_init(&bars[A3_B_PREADY + 2 * i], 128);
      mbar_init(&bars[A3_B_PREADY + 2 * i + 1], 128);
      mbar_init(&bars[A3_B_PVDONE + 2 * i], 1);
      mbar_init(&bars[A3_B_PVDONE + 2 * i + 1], 1);
    }
    for (int i = 0; i < A3_NS; ++i) {
      mbar_init(&bars[A3_B_KFULL + i], 1);
      mbar_init(&bars[A3_B_KEMPTY + i], DUAL ? 2 : 1);
      mbar_init(&bars[A3_B_VFULL + i], 1);
      mbar_init(&bars[A3_B_VEMPTY + i], DUAL ? 2 : 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      int n = 0, g = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int pr = item % npair, h = (item / npair) % H, b = item / (npair * H);
        const int buf = n & 1;
        const bool has_b = 2 * pr + 1 < nq;
        const bool titem = tail && (2 * pr == nq - 1 || 2 * pr + 1 == nq - 1);
        mbar_wait(&bars[A3_B_QFREE + buf], ((n >> 1) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(&bars[A3_B_QFULL + buf], (has_b ? 2 : 1) * A3_TILE + (tail ? 256 : 0) + (titem ? 128 : 0));
        tma_load_4d(smem + A3_OFF_Q + (2 * buf) * A3_TILE, &tmQKV, &bars[A3_B_QFULL + buf], 0, h, 2 * pr * A3_BQ, b);
        if (has_b)
          tma_load_4d(smem + A3_OFF_Q + (2 * buf + 1) * A3_TILE, &tmQKV, &bars[A3_B_QFULL + buf], 0, h, (2 * pr + 1) * A3_BQ, b);
        if (tail) {
          tma_load_4d(smem + A3_OFF_ROWS + buf * 384, &tmRow, &bars[A3_B_QFULL + buf], 0, H + h, Lm, b);
          tma_load_4d(smem + A3_OFF_ROWS + buf * 384 + 128, &tmRow, &bars[A3_B_QFULL + buf], 0, 2 * H + h, Lm, b);
          if (titem) tma_load_4d(smem + A3_OFF_ROWS + buf * 384 + 256, &tmRow, &bars[A3_B_QFULL + buf], 0, h, Lm, b);
        }
        for (int j = 0; j < nkv; ++j, ++g) {
          const int s = g % A3_NS;
          const uint32_t ph = (g / A3_NS) & 1;
          mbar_wait(&bars[A3_B_KEMPTY + s], ph ^ 1, 41);
          mbar_arrive_expect_tx(&bars[A3_B_KFULL + s], A3_TILE);
          tma_load_4d(smem + A3_OFF_K + s * A3_TILE, &tmQKV, &bars[A3_B_KFULL + s], 0, H + h, j * A3_BKV, b);
          mbar_wait(&bars[A3_B_VEMPTY + s], ph ^ 1, 42);
          mbar_arrive_expect_tx(&bars[A3_B_VFULL + s], A3_TILE);
          tma_load_4d(smem + A3_OFF_V + s * A3_TILE, &tmQKV, &bars[A3_B_VFULL + s], 0, 2 * H + h, j * A3_BKV, b);
        }
      }
    }
  } else if (warp == 9 || (DUAL && warp == 10)) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer(s)
      const int me = warp - 9;   // DUAL: the tile slot this warp serves
      int n = 0, g0 = 0;
      uint32_t p_par = 0;    // bit 2w + (hb & 1): parity of the next p_ready phase of that barrier
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int pr = item % npair;
        const int buf = n & 1;
        const int ntile = (2 * pr + 1 < nq) ? 2 : 1;
        const int tw = !tail ? -1 : (2 * pr == nq - 1 ? 0 : (2 * pr + 1 == nq - 1 ? 1 : -1));   // tile carrying the remainder query
        const uint32_t rows = smem_u32(smem + A3_OFF_ROWS + buf * 384);
        const int w_lo = DUAL ? me : 0, w_hi = DUAL ? min(me + 1, ntile) : ntile;
        if (DUAL && me >= ntile) {   // no tile for this warp in this item: keep the K / V ring's two-arrival protocol going
          for (int j = 0; j < nkv; ++j) {
            const int gg = g0 + j;
            const int s = gg % A3_NS;
            mbar_wait(&bars[A3_B_KFULL + s], (gg / A3_NS) & 1, 52);
            mbar_arrive(&bars[A3_B_KEMPTY + s]);
            mbar_wait(&bars[A3_B_VFULL + s], (gg / A3_NS) & 1, 53);
            mbar_arrive(&bars[A3_B_VEMPTY + s]);
          }
          g0 += nkv;
          continue;
        }
        mbar_wait_x<SPIN>(&bars[A3_B_QFULL + buf], (n >> 1) & 1, 43);
        // S_w(hb) = Q_w K(hb)^T into buffer hb & 1 of the tile's score columns (+ the remainder-token side products)
        auto issue_s = [&](int w, int hb) {
          const int j = hb >> 1, half = hb & 1;
          const int gg = g0 + j;
          const int s = gg % A3_NS;
          const int valid = min(A3_HB, Lm - hb * A3_HB);
          const int nblk = (valid + 15) & ~15;
          const uint32_t q_addr = smem_u32(smem + A3_OFF_Q + (2 * buf + w) * A3_TILE);
          const uint32_t k_tile = smem_u32(smem + A3_OFF_K + s * A3_TILE);
          const uint32_t k_addr = k_tile + half * A3_HALF;
          mbar_wait_x<SPIN>(&bars[A3_B_KFULL + s], (gg / A3_NS) & 1, 44);
          tc_fence_after();
          const uint32_t idesc_s = umma_idesc_bf16(A3_BQ, nblk, 0, 0);
#pragma unroll
          for (int k = 0; k < A3_HD / 16; ++k)
            umma_bf16_ss(tmem_base + A3_T_S + 128 * w + 64 * half, umma_desc_kmajor_sw128(q_addr + k * 32),
                         umma_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0);
          constexpr uint32_t idesc_16 = umma_idesc_bf16(128, 16, 0, 0);
          if (tail && hb == 0) {   // this tile's rows against the remainder key
#pragma unroll
            for (int k = 0; k < A3_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A3_T_SK + 16 * w, umma_desc_kmajor_sw128(q_addr + k * 32), umma_desc_row0(rows + k * 32),
                           idesc_16, k != 0);
          }
          if (w == tw && half == 0) {   // remainder query row against the 128 keys of this tile (transposed), and the remainder key
#pragma unroll
            for (int k = 0; k < A3_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A3_T_ST, umma_desc_kmajor_sw128(k_tile + k * 32), umma_desc_row0(rows + 256 + k * 32),
                           idesc_16, k != 0);
            if (hb == 0) {
#pragma unroll
              for (int k = 0; k < A3_HD / 16; ++k)
                umma_bf16_ss(tmem_base + A3_T_TT, umma_desc_row0(rows + k * 32), umma_desc_row0(rows + 256 + k * 32), idesc_16,
                             k != 0);
            }
          }
          if ((DUAL || w == ntile - 1) && (half == 1 || hb == nhb - 1)) umma_commit(&bars[A3_B_KEMPTY + s]);   // last read of this K tile
          umma_commit(&bars[A3_B_SFULL + 2 * w + half]);
        };
        for (int w = w_lo; w < w_hi; ++w) issue_s(w, 0);
        if (nhb > 1)
          for (int w = w_lo; w < w_hi; ++w) issue_s(w, 1);
        for (int hb = 0; hb < nhb; ++hb) {
          const int j = hb >> 1, half = hb & 1;
          const int gg = g0 + j;
          const int s = gg % A3_NS;
          const int valid = min(A3_HB, Lm - hb * A3_HB);
          const int ksteps = ((valid + 15) & ~15) / 16;
          const uint32_t v_tile = smem_u32(smem + A3_OFF_V + s * A3_TILE);
          const uint32_t v_addr = v_tile + half * A3_HALF;
          for (int w = w_lo; w < w_hi; ++w) {
            mbar_wait_x<SPIN>(&bars[A3_B_PREADY + 2 * w + half], (p_par >> (2 * w + half)) & 1, 45);   // P_w(hb) is in TMEM (O_w rescaled if needed)
            p_par ^= 1u << (2 * w + half);
            mbar_wait_x<SPIN>(&bars[A3_B_VFULL + s], (gg / A3_NS) & 1, 46);
            tc_fence_after();
            constexpr uint32_t idesc_pv = umma_idesc_bf16(A3_BQ, A3_HD, 0, 1);
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16_ts3(tmem_base + A3_T_O + 64 * w, tmem_base + A3_T_S + 128 * w + 64 * half + 8 * kk,
                            umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, A3_TILE), idesc_pv, (hb | kk) != 0);
            if (w == tw && half == 0) {   // O_t^T += V_j^T p_t^T over the tile's 128 keys (M runs over head dims, rows 64..127 padding)
              constexpr uint32_t idesc_ot = umma_idesc_bf16(128, 16, 1, 0);
              const uint32_t pt_addr = smem_u32(smem + A3_OFF_PT);
              for (int kk = 0; kk < A3_BKV / 16; ++kk)
                umma_bf16_ss(tmem_base + A3_T_OT, umma_desc_mnmajor_sw128(v_tile + kk * 16 * 128, A3_TILE),
                             umma_desc_row0(pt_addr + kk * 32), idesc_ot, (j | kk) != 0);
            }
            if ((DUAL || w == ntile - 1) && (half == 1 || hb == nhb - 1)) umma_commit(&bars[A3_B_VEMPTY + s]);   // last read of this V tile
            umma_commit(&bars[A3_B_PVDONE + 2 * w + half]);
            if (hb + 2 < nhb) issue_s(w, hb + 2);   // refills the buffer P V_w(hb) has just been queued to read
          }
        }
        g0 += nkv;
      }
    }
  } else if (warp < 8) {
    // -------------------------------------------------------------------- softmax groups
    const int w = warp >> 2;                 // tile slot (0 = A, 1 = B)
    const int quad = warp & 3;
    const int r = quad * 32 + static_cast<int>(lane);   // query row in the tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_sw = tmem_base + t_lane + A3_T_S + 128 * w;
    const uint32_t t_o = tmem_base + t_lane + A3_T_O + 64 * w;
    const uint32_t red = smem_u32(smem + A3_OFF_RED);
    const uint32_t bar_id = 1 + w;
    int n = 0;
    uint32_t s_par = 0;    // bit (hb & 1): parity of the next s_full phase of that score buffer
    uint32_t pv_base = 0;  // bit (hb & 1): parity of the first pv_done phase of that barrier in the current item
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
      const int pr = item % npair, h = (item / npair) % H, b = item / (npair * H);
      const int buf = n & 1;
      const int qt = 2 * pr + w;
      if (qt >= nq) {   // odd number of query tiles: group B sits this item out (but stays in step with the producer)
        if (r == 0) {
          mbar_wait(&bars[A3_B_QFULL + buf], (n >> 1) & 1, 49);
          mbar_arrive(&bars[A3_B_QFREE + buf]);
        }
        continue;
      }
      const int q0 = qt * A3_BQ;
      const bool titem = tail && qt == nq - 1;   // this group also carries the remainder query row
      float m_ref = -INFINITY, l = 0.f, s_tail = 0.f;
      float m_t = -INFINITY, l_t = 0.f;
      uint32_t u_tt = 0;   // q_t . k_t (thread 0 of the group)
      const float neg_inf = -INFINITY;
#pragma unroll 1
      for (int hb = 0; hb < nhb; ++hb) {
        const int half = hb & 1;
        const int valid = min(A3_HB, Lm - hb * A3_HB);
        const uint32_t t_s = t_sw + 64 * half;
        mbar_wait_x<SPIN>(&bars[A3_B_SFULL + 2 * w + half], (s_par >> half) & 1, 47);
        s_par ^= 1u << half;
        tc_fence_after();
        uint32_t u_sk = 0, u_st = 0;
        if (titem && hb == 0 && quad == 0) tmem_ld_x1(tmem_base + t_lane + A3_T_TT, u_tt);   // (the next item's S(0) rewrites it)
        if (tail && hb == 0) tmem_ld_x1(tmem_base + t_lane + A3_T_SK + 16 * w, u_sk);
        if (titem && half == 0) tmem_ld_x1(tmem_base + t_lane + A3_T_ST, u_st);
        // pass 1: row maximum of the 64 scores
        float mx = neg_inf;
        {
          uint32_t sv[64];
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
          tmem_ld_x32(t_s, s0);
          if (valid > 32) tmem_ld_x32(t_s + 32, s1);
          tmem_ld_wait();
          if (valid == A3_HB) {
            float m4[4] = {neg_inf, neg_inf, neg_inf, neg_inf};
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = max3f(m4[i & 3], __uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1]));
            mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) mx = (i < valid) ? fmaxf(mx, __uint_as_float(sv[i])) : mx;
          }
        }
        if (tail && hb == 0) s_tail = __uint_as_float(u_sk) * scale_log2;
        const float m_blk = mx * scale_log2;
        // lazy rescaling: the reference only moves when the maximum grew by more than 2^8.  O may only be touched once
        // P V(hb-1) has landed (S(hb) was issued behind P V(hb-2) only)
        if (hb == 0) {
          m_ref = m_blk;
        } else {
          const bool grow = m_blk > m_ref + 8.f;
          if (__any_sync(0xffffffffu, grow)) {
            mbar_wait_x<SPIN>(&bars[A3_B_PVDONE + 2 * w + (half ^ 1)], ((pv_base >> (half ^ 1)) + ((hb - 1) >> 1)) & 1, 50);
            tc_fence_after();
            const float alpha = grow ? fast_exp2(m_ref - m_blk) : 1.f;
            m_ref = grow ? m_blk : m_ref;
            l *= alpha;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_x16(t_o + 16 * c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(t_o + 16 * c, o);
            }
          }
        }
        // remainder query row (once per 128-key tile): this thread holds the score of key r of the tile
        if (titem && half == 0) {
          const int j = hb >> 1;
          const float st = __uint_as_float(u_st) * scale_log2;
          float wm = st;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
          if (lane == 0) sts_f32(red + (j & 1) * 16 + quad * 4, wm);
          named_bar_sync(bar_id, 128);
          const float4 w4 = lds_f32x4(red + (j & 1) * 16);
          const float mt = fmaxf(fmaxf(w4.x, w4.y), fmaxf(w4.z, w4.w));
          if (j == 0) {
            m_t = mt;
          } else if (mt > m_t + 8.f) {   // uniform over the group
            const float alpha = fast_exp2(m_t - mt);
            m_t = mt;
            l_t *= alpha;
            if (quad < 2) {              // O_t^T: lanes 0..63 = head dims, column 0
              uint32_t o1;
              tmem_ld_x1(tmem_base + t_lane + A3_T_OT, o1);
              tmem_ld_wait();
              tmem_st_x1(tmem_base + t_lane + A3_T_OT, __float_as_uint(__uint_as_float(o1) * alpha));
            }
          }
          const float p = fast_exp2(st - m_t);
          l_t += p;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(smem + A3_OFF_PT) + 2 * r),
                       "h"(static_cast<unsigned short>(pack_bf16x2(p, 0.f) & 0xFFFFu)) : "memory");
        }
        // pass 2: P = 2^(S*scale - m_ref) as packed bf16 pairs over the score columns already consumed: chunk c (score
        // columns [16c, 16c+16)) becomes words [8c, 8c+8) of the buffer
        const float neg_m = -m_ref;
        if (valid == A3_HB) {
          uint32_t xa[16], xb[16], pw[8];
          float ea[16], eb[16];
          tmem_ld_x16(t_s, xa);
          tmem_ld_x16(t_s + 16, xb);
          tmem_ld_wait();
          exp_chunk(xa, scale_log2, neg_m, ea);                 // chunk 0
          tmem_ld_x16(t_s + 32, xa);
          exp_chunk(xb, scale_log2, neg_m, eb);                 // chunk 1
          float acc = pack_chunk(ea, pw);                       // consume chunk 0
          tmem_st_x8(t_s, pw);
          tmem_ld_wait();
          tmem_ld_x16(t_s + 48, xb);
          exp_chunk(xa, scale_log2, neg_m, ea);                 // chunk 2
          acc += pack_chunk(eb, pw);                            // consume chunk 1
          tmem_st_x8(t_s + 8, pw);
          tmem_ld_wait();
          exp_chunk(xb, scale_log2, neg_m, eb);                 // chunk 3
          acc += pack_chunk(ea, pw);                            // consume chunk 2
          tmem_st_x8(t_s + 16, pw);
          acc += pack_chunk(eb, pw);                            // consume chunk 3
          tmem_st_x8(t_s + 24, pw);
          l += acc;
        } else {
          float acc = 0.f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            if (16 * c < valid) {
              uint32_t x[16], pw[8];
              tmem_ld_x16(t_s + 16 * c, x);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int k0 = 16 * c + 2 * i;
                float p0 = fast_exp2(fmaf(__uint_as_float(x[2 * i]), scale_log2, neg_m));
                float p1 = fast_exp2(fmaf(__uint_as_float(x[2 * i + 1]), scale_log2, neg_m));
                p0 = (k0 < valid) ? p0 : 0.f;
                p1 = (k0 + 1 < valid) ? p1 : 0.f;
                acc += p0 + p1;
                pw[i] = pack_bf16x2(p0, p1);
              }
              tmem_st_x8(t_s + 8 * c, pw);
            }
          }
          l += acc;
        }
        tmem_st_wait();
        if (titem && half == 0) fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[A3_B_PREADY + 2 * w + half]);
      }
      // ------------------------------------------------------------------ epilogue: O / l -> bf16 -> smem -> TMA store
      {
        const int e = (nhb - 1) & 1;   // the last P V; its barrier's previous phase (P V(nhb-3)) completed before S(nhb-1) did
        mbar_wait_x<SPIN>(&bars[A3_B_PVDONE + 2 * w + e], ((pv_base >> e) + ((nhb - 1) >> 1)) & 1, 48);
        pv_base ^= (((nhb + 1) >> 1) & 1) | (((nhb >> 1) & 1) << 1);
      }
      tc_fence_after();
      uint32_t o[64];
      {
        uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
        uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
        tmem_ld_x32(t_o, o0);
        tmem_ld_x32(t_o + 32, o1);
      }
      uint32_t u_ot = 0;
      if (titem && quad < 2) tmem_ld_x1(tmem_base + t_lane + A3_T_OT, u_ot);
      tmem_ld_wait();
      const uint32_t rows = smem_u32(smem + A3_OFF_ROWS + buf * 384);
      if (tail) {   // fold the remainder key in: one more online-softmax step, entirely in registers
        const float m_fin = fmaxf(m_ref, s_tail);
        const float a = fast_exp2(m_ref - m_fin);
        const float pt = fast_exp2(s_tail - m_fin);
        l = fmaf(l, a, pt);
        m_ref = m_fin;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 vv = lds128(rows + 128 + c * 16);
          const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[8 * c + 2 * q] = __float_as_uint(fmaf(__uint_as_float(o[8 * c + 2 * q]), a, pt * bf16_lo(w4[q])));
            o[8 * c + 2 * q + 1] = __float_as_uint(fmaf(__uint_as_float(o[8 * c + 2 * q + 1]), a, pt * bf16_hi(w4[q])));
          }
        }
      }
      const float inv_l = 1.f / l;
      const uint32_t stage = smem_u32(smem + A3_OFF_Q + (2 * buf + w) * A3_TILE);   // this tile's Q buffer is dead by now
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(stage + sw128_offset(r, c),
               make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l)));
      if (lse_out != nullptr && q0 + r < L)
        lse_out[(static_cast<long long>(b) * H + h) * L + q0 + r] = (m_ref + log2f(l)) * 0.69314718055994531f;
      if (titem) {   // finish the remainder query row: totals over the group, its own remainder key, normalise, store
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) l_t += __shfl_xor_sync(0xffffffffu, l_t, o2);
        if (lane == 0) sts_f32(red + 32 + quad * 4, l_t);
        if (r == 0) sts_f32(red + 48, __uint_as_float(u_tt) * scale_log2);
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (titem && quad < 2) {
        const float4 w4 = lds_f32x4(red + 32);
        const float s_tt = lds_f32(red + 48);
        const float m_fin = fmaxf(m_t, s_tt);
        const float a = fast_exp2(m_t - m_fin);
        const float pt = fast_exp2(s_tt - m_fin);
        const float l_all = fmaf((w4.x + w4.y) + (w4.z + w4.w), a, pt);
        unsigned short vb;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(vb) : "r"(rows + 128 + 2 * r));
        const float v = __uint_as_float(static_cast<uint32_t>(vb) << 16);
        const float ov = fmaf(__uint_as_float(u_ot), a, pt * v) / l_all;
        const float o_hi = __shfl_down_sync(0xffffffffu, ov, 1);
        const long long row = static_cast<long long>(b) * L + Lm;
        if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(out + (row * H + h) * A3_HD + r) = pack_bf16x2(ov, o_hi);
        if (r == 0 && lse_out != nullptr)
          lse_out[(static_cast<long long>(b) * H + h) * L + Lm] = (m_fin + log2f(l_all)) * 0.69314718055994531f;
      }
      if (r == 0) {
        tma_store_4d(&tmO, smem + A3_OFF_Q + (2 * buf + w) * A3_TILE, 0, h, q0, b);
        tma_store_commit();
        tma_store_wait_read<0>();   // the producer may refill this Q buffer
        mbar_arrive(&bars[A3_B_QFREE + buf]);
      }
    }
    if (r == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ovk

using namespace ovk;

// Called by ovk_attention_fwd for hd == 64 (attention.cu); returns OVK_OK or an error code.
int ovk_attention_fwd3_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s) {
  const int hd = A3_HD;
  CUtensorMap tmQKV, tmO, tmRow;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {A3_HD, 1, A3_BKV, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const uint32_t box1[4] = {A3_HD, 1, 1, 1};   // one head row (128 B), unswizzled: remainder key / value / query
    if ((rc = make_tmap_nd_bf16(&tmRow, qkv, 4, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {A3_HD, 1, A3_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  // OVK_ATT3_MODE (A/B switch, read per call): bit 0 = one MMA warp per tile slot, bit 1 = polling hand-off waits
  static const int default_mode = 0;
  const char* me = getenv("OVK_ATT3_MODE");
  const int mode = (me != nullptr && me[0] >= '0' && me[0] <= '3') ? me[0] - '0' : default_mode;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_fwd3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_fwd3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_fwd3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention3): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  const int tail = (L > A3_BQ && L % A3_BQ == 1) ? 1 : 0;   // cls + power-of-two grid: remainder token handled outside the tiles
  const int l_main = L - tail;
  const int nq = (l_main + A3_BQ - 1) / A3_BQ;
  const long long items = static_cast<long long>((nq + 1) / 2) * H * B;
  if (items > 0x7fffffffLL) return set_error(OVK_ERR_SHAPE, "attention: too many work items");
  const int grid = static_cast<int>(items < (long long)num_sms() ? items : (long long)num_sms());
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  const float sl2 = scale * 1.4426950408889634f;
  const int it = static_cast<int>(items);
  switch (mode) {
    case 1: attention_fwd3_kernel<true, false><<<grid, A3_THREADS + 32, A3_SMEM, s>>>(tmQKV, tmO, tmRow, lse, o, L, l_main, H, nq, it, sl2); break;
    case 2: attention_fwd3_kernel<false, true><<<grid, A3_THREADS, A3_SMEM, s>>>(tmQKV, tmO, tmRow, lse, o, L, l_main, H, nq, it, sl2); break;
    case 3: attention_fwd3_kernel<true, true><<<grid, A3_THREADS + 32, A3_SMEM, s>>>(tmQKV, tmO, tmRow, lse, o, L, l_main, H, nq, it, sl2); break;
    default: attention_fwd3_kernel<false, false><<<grid, A3_THREADS, A3_SMEM, s>>>(tmQKV, tmO, tmRow, lse, o, L, l_main, H, nq, it, sl2); break;
  }
  return check_launch("attention_fwd3_kernel");
}
