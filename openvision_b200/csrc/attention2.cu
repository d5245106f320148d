// Multi-head self-attention forward, head width 64, second-generation kernel: O = softmax(Q K^T * scale) V.
// Replaces the core of nn.MultiheadAttention(need_weights=False, attn_mask=None) between in_proj and out_proj
// (open_clip/transformer.py:225,239-252); online softmax as in src/models/bpt.py:105-124.
//
// One persistent CTA per SM (512 TMEM columns, 320 threads, up to 200 registers each).  A work item is a PAIR of 128-row query tiles of one
// (head, image): K / V blocks are fetched once for both tiles.
//   warps 0-3  : softmax group A — thread <-> query row of tile A (no cross-thread exchange, no block-wide barrier
//                inside the key loop; the row is read from TMEM once for the maximum and once for the exponentials)
//   warps 4-7  : softmax group B — the same for tile B
//   warp  8    : TMA producer (Q tiles double-buffered across items, K / V ring of 3 blocks)
//   warp  9    : MMA issuer.  S_w = Q_w K_j^T -> TMEM; P_w (bf16) is written back INTO the S_w columns and consumed from
//                TMEM as the A operand of O_w += P_w V_j (no shared-memory round trip for P).  The two tiles alternate on
//                the tensor core: while group A exponentiates S_A(j), the pipe runs P V_B(j-1) and S_B(j), and vice versa.
// TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  and five 16-column accumulators for the remainder
// token of L = 128 k + 1 sequences (cls + power-of-two grid), see below.
//
// Remainder token (L = 128k+1; its key AND its query row do not fill a tile):
//   * key:   s_k = Q_w k_t^T from an N = 16 MMA whose B operand is the single 128-byte key row (row-0-only descriptor);
//            each thread reads one TMEM word and folds the key into its output row in the epilogue.
//   * query: rides with the LAST tile of the head in transposed form: S_t^T = K_j q_t^T (lane = key) gives every thread
//            of that group one score, the probabilities go back as a 256-byte row-0-only operand, O_t^T = V_j^T p_t^T
//            accumulates in 16 columns (lane = head dim); q_t . k_t comes from a 1 x 1 MMA of the two single rows.
#include <cuda_bf16.h>
#include <stdint.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int A2_BQ = 128;
constexpr int A2_BKV = 128;
constexpr int A2_HD = 64;
constexpr int A2_THREADS = 320;   // 8 softmax warps + producer + MMA issuer: 204 registers per thread available
constexpr int A2_TILE = 128 * 128;   // [128 rows x 64 bf16], SWIZZLE_128B
constexpr int A2_NS = 3;             // K / V ring depth (blocks)
constexpr int A2_OFF_Q = 0;                          // [2 item buffers][2 tiles]
// V before K: the O_t^T MMA reads V as an M = 128 MN-major operand whose second (padding) panel lies one tile past
// the V block, which must still be inside the allocation
constexpr int A2_OFF_V = A2_OFF_Q + 4 * A2_TILE;
constexpr int A2_OFF_K = A2_OFF_V + A2_NS * A2_TILE;
constexpr int A2_OFF_ROWS = A2_OFF_K + A2_NS * A2_TILE;  // per item buffer: k_t, v_t, q_t rows (128 B each)
constexpr int A2_OFF_PT = A2_OFF_ROWS + 2 * 384;          // p_t: bf16 [128] (+128 B that the MMA over-reads)
constexpr int A2_OFF_RED = A2_OFF_PT + 384;               // reduction scratch of the remainder row: [2][4] max, [4] sum, s_tt
constexpr int A2_OFF_BAR = A2_OFF_RED + 64;
// barriers
constexpr int A2_B_QFULL = 0;     // [2]
constexpr int A2_B_QFREE = 2;     // [2]  both groups are done with the item's Q buffer (also used as output staging)
constexpr int A2_B_KFULL = 4;     // [NS]
constexpr int A2_B_KEMPTY = 7;    // [NS]
constexpr int A2_B_VFULL = 10;    // [NS]
constexpr int A2_B_VEMPTY = 13;   // [NS]
constexpr int A2_B_SFULL = 16;    // [2]  per tile
constexpr int A2_B_PREADY = 18;   // [2]
constexpr int A2_B_ODONE = 20;    // [2]
constexpr int A2_NUM_BARS = 22;
constexpr int A2_SMEM = A2_OFF_BAR + A2_NUM_BARS * 8 + 16;
constexpr uint32_t A2_T_S = 0, A2_T_O = 256, A2_T_SK = 384, A2_T_ST = 416, A2_T_OT = 432, A2_T_TT = 448;

// D[tmem] (+)= A[tmem] * B[smem]^T : A is the [128 x 16] bf16 block held as 8 TMEM columns of packed pairs
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(A2_THREADS, 1)
attention_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmRow, float* __restrict__ lse_out,
                      __nv_bfloat16* __restrict__ out, int L, int Lm, int H, int nq, int total_items, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention2: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + A2_OFF_BAR + A2_NUM_BARS * 8);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int npair = (nq + 1) >> 1;
  const int nkv = (Lm + A2_BKV - 1) / A2_BKV;
  const bool tail = L > Lm;   // one remainder token (host guarantees L - Lm <= 1)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmRow);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[A2_B_QFULL + i], 1);
      mbar_init(&bars[A2_B_QFREE + i], 2);
      mbar_init(&bars[A2_B_SFULL + i], 1);
      mbar_init(&bars[A2_B_PREADY + i], 128);
      mbar_init(&bars[A2_B_ODONE + i], 1);
    }
    for (int i = 0; i < A2_NS; ++i) {
      mbar_init(&bars[A2_B_KFULL + i], 1);
      mbar_init(&bars[A2_B_KEMPTY + i], 1);
      mbar_init(&bars[A2_B_VFULL + i], 1);
      mbar_init(&bars[A2_B_VEMPTY + i], 1);
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
        mbar_wait(&bars[A2_B_QFREE + buf], ((n >> 1) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(&bars[A2_B_QFULL + buf], (has_b ? 2 : 1) * A2_TILE + (tail ? 256 : 0) + (titem ? 128 : 0));
        tma_load_4d(smem + A2_OFF_Q + (2 * buf) * A2_TILE, &tmQKV, &bars[A2_B_QFULL + buf], 0, h, 2 * pr * A2_BQ, b);
        if (has_b)
          tma_load_4d(smem + A2_OFF_Q + (2 * buf + 1) * A2_TILE, &tmQKV, &bars[A2_B_QFULL + buf], 0, h, (2 * pr + 1) * A2_BQ, b);
        if (tail) {
          tma_load_4d(smem + A2_OFF_ROWS + buf * 384, &tmRow, &bars[A2_B_QFULL + buf], 0, H + h, Lm, b);
          tma_load_4d(smem + A2_OFF_ROWS + buf * 384 + 128, &tmRow, &bars[A2_B_QFULL + buf], 0, 2 * H + h, Lm, b);
          if (titem) tma_load_4d(smem + A2_OFF_ROWS + buf * 384 + 256, &tmRow, &bars[A2_B_QFULL + buf], 0, h, Lm, b);
        }
        for (int j = 0; j < nkv; ++j, ++g) {
          const int s = g % A2_NS;
          const uint32_t ph = (g / A2_NS) & 1;
          mbar_wait(&bars[A2_B_KEMPTY + s], ph ^ 1, 41);
          mbar_arrive_expect_tx(&bars[A2_B_KFULL + s], A2_TILE);
          tma_load_4d(smem + A2_OFF_K + s * A2_TILE, &tmQKV, &bars[A2_B_KFULL + s], 0, H + h, j * A2_BKV, b);
          mbar_wait(&bars[A2_B_VEMPTY + s], ph ^ 1, 42);
          mbar_arrive_expect_tx(&bars[A2_B_VFULL + s], A2_TILE);
          tma_load_4d(smem + A2_OFF_V + s * A2_TILE, &tmQKV, &bars[A2_B_VFULL + s], 0, 2 * H + h, j * A2_BKV, b);
        }
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      int n = 0, g0 = 0;
      int cnt[2] = {0, 0};   // key blocks processed so far per tile slot (phase of s_full / p_ready)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int pr = item % npair;
        const int buf = n & 1;
        const int ntile = (2 * pr + 1 < nq) ? 2 : 1;
        const int tw = !tail ? -1 : (2 * pr == nq - 1 ? 0 : (2 * pr + 1 == nq - 1 ? 1 : -1));   // tile carrying the remainder query
        const uint32_t rows = smem_u32(smem + A2_OFF_ROWS + buf * 384);
        mbar_wait(&bars[A2_B_QFULL + buf], (n >> 1) & 1, 43);
        uint64_t s_use = 0, v_use = 0;   // 2-bit counters per key block: tiles that have issued S(j) / P V(j)
        auto issue_s = [&](int w, int j) {
          const int gg = g0 + j;
          const int s = gg % A2_NS;
          const int valid = min(A2_BKV, Lm - j * A2_BKV);
          const int nblk = (valid + 15) & ~15;
          const uint32_t q_addr = smem_u32(smem + A2_OFF_Q + (2 * buf + w) * A2_TILE);
          const uint32_t k_addr = smem_u32(smem + A2_OFF_K + s * A2_TILE);
          mbar_wait(&bars[A2_B_KFULL + s], (gg / A2_NS) & 1, 44);
          tc_fence_after();
          const uint32_t idesc_s = umma_idesc_bf16(A2_BQ, nblk, 0, 0);
#pragma unroll
          for (int k = 0; k < A2_HD / 16; ++k)
            umma_bf16_ss(tmem_base + A2_T_S + 128 * w, umma_desc_kmajor_sw128(q_addr + k * 32),
                         umma_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0);
          constexpr uint32_t idesc_16 = umma_idesc_bf16(128, 16, 0, 0);
          if (tail && j == 0) {   // this tile's rows against the remainder key
#pragma unroll
            for (int k = 0; k < A2_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A2_T_SK + 16 * w, umma_desc_kmajor_sw128(q_addr + k * 32), umma_desc_row0(rows + k * 32),
                           idesc_16, k != 0);
          }
          if (w == tw) {          // remainder query row against this key block (transposed), and against the remainder key
#pragma unroll
            for (int k = 0; k < A2_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A2_T_ST, umma_desc_kmajor_sw128(k_addr + k * 32), umma_desc_row0(rows + 256 + k * 32),
                           idesc_16, k != 0);
            if (j == 0) {
#pragma unroll
              for (int k = 0; k < A2_HD / 16; ++k)
                umma_bf16_ss(tmem_base + A2_T_TT, umma_desc_row0(rows + k * 32), umma_desc_row0(rows + 256 + k * 32), idesc_16,
                             k != 0);
            }
          }
          s_use += 1ull << (2 * j);
          if (static_cast<int>((s_use >> (2 * j)) & 3) == ntile) umma_commit(&bars[A2_B_KEMPTY + s]);   // last reader of K_j
          umma_commit(&bars[A2_B_SFULL + w]);
        };
        for (int w = 0; w < ntile; ++w) issue_s(w, 0);
        // The two tiles strictly alternate (serving whichever group finishes first, through a polling loop, measured
        // slower: the groups have equal work and the poll delays both).
        for (int j = 0; j < nkv; ++j) {
          const int gg = g0 + j;
          const int s = gg % A2_NS;
          const int valid = min(A2_BKV, Lm - j * A2_BKV);
          const int ksteps = ((valid + 15) & ~15) / 16;
          const uint32_t v_addr = smem_u32(smem + A2_OFF_V + s * A2_TILE);
          for (int w = 0; w < ntile; ++w) {
            mbar_wait(&bars[A2_B_PREADY + w], cnt[w] & 1, 45);   // P_w(j) is in TMEM (and O_w is rescaled if needed)
            ++cnt[w];
            mbar_wait(&bars[A2_B_VFULL + s], (gg / A2_NS) & 1, 46);
            tc_fence_after();
            constexpr uint32_t idesc_pv = umma_idesc_bf16(A2_BQ, A2_HD, 0, 1);
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16_ts(tmem_base + A2_T_O + 64 * w, tmem_base + A2_T_S + 128 * w + 8 * kk,
                           umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, A2_TILE), idesc_pv, (j | kk) != 0);
            if (w == tw) {   // O_t^T += V_j^T p_t^T (A = V as an MN-major operand: M runs over head dims, rows 64..127 padding)
              constexpr uint32_t idesc_ot = umma_idesc_bf16(128, 16, 1, 0);
              const uint32_t pt_addr = smem_u32(smem + A2_OFF_PT);
              for (int kk = 0; kk < A2_BKV / 16; ++kk)
                umma_bf16_ss(tmem_base + A2_T_OT, umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, A2_TILE),
                             umma_desc_row0(pt_addr + kk * 32), idesc_ot, (j | kk) != 0);
            }
            v_use += 1ull << (2 * j);
            if (static_cast<int>((v_use >> (2 * j)) & 3) == ntile) umma_commit(&bars[A2_B_VEMPTY + s]);   // last reader of V_j
            if (j + 1 < nkv) issue_s(w, j + 1);                        // overwrites S_w / P_w: ordered after P V_w(j)
            else umma_commit(&bars[A2_B_ODONE + w]);
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
    const uint32_t t_s = tmem_base + t_lane + A2_T_S + 128 * w;
    const uint32_t t_o = tmem_base + t_lane + A2_T_O + 64 * w;
    const uint32_t red = smem_u32(smem + A2_OFF_RED);
    const uint32_t bar_id = 1 + w;
    int n = 0, cnt = 0, cnt_o = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
      const int pr = item % npair, h = (item / npair) % H, b = item / (npair * H);
      const int buf = n & 1;
      const int qt = 2 * pr + w;
      if (qt >= nq) {   // odd number of query tiles: group B sits this item out (but stays in step with the producer)
        if (r == 0) {
          mbar_wait(&bars[A2_B_QFULL + buf], (n >> 1) & 1, 49);
          mbar_arrive(&bars[A2_B_QFREE + buf]);
        }
        continue;
      }
      const int q0 = qt * A2_BQ;
      const bool titem = tail && qt == nq - 1;   // this group also carries the remainder query row
      float m_ref = -INFINITY, l = 0.f, s_tail = 0.f;
      float m_t = -INFINITY, l_t = 0.f;
      uint32_t u_tt = 0;   // q_t . k_t (thread 0 of the group)
      for (int j = 0; j < nkv; ++j) {
        const int valid = min(A2_BKV, Lm - j * A2_BKV);
        mbar_wait(&bars[A2_B_SFULL + w], cnt & 1, 47);
        ++cnt;
        tc_fence_after();
        uint32_t u_sk = 0, u_st = 0;
        if (titem && j == 0 && quad == 0) tmem_ld_x1(tmem_base + t_lane + A2_T_TT, u_tt);   // (the next item's S(0) rewrites it)
        if (tail && j == 0) tmem_ld_x1(tmem_base + t_lane + A2_T_SK + 16 * w, u_sk);
        if (titem) tmem_ld_x1(tmem_base + t_lane + A2_T_ST, u_st);
        // pass 1: row maximum, 32 columns at a time (the scores are read again for the exponentials: a 128-register score
        // row does not fit the 168-register budget; software-pipelining these loads measured slower)
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (32 * c < valid) {
            uint32_t sv[32];
            tmem_ld_x32(t_s + 32 * c, sv);
            tmem_ld_wait();
            if (32 * c + 32 <= valid) {
              float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
              for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[i]));
              mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = (32 * c + i < valid) ? fmaxf(mx, __uint_as_float(sv[i])) : mx;
            }
          }
        }
        if (tail && j == 0) s_tail = __uint_as_float(u_sk) * scale_log2;
        const float m_blk = mx * scale_log2;
        // lazy rescaling: the reference only moves when the maximum grew by more than 2^8 (P V(j-1) has completed:
        // s_full(j) was committed after it)
        if (j == 0) {
          m_ref = m_blk;
        } else {
          const bool grow = m_blk > m_ref + 8.f;
          if (__any_sync(0xffffffffu, grow)) {
            const float alpha = grow ? fast_exp2(m_ref - m_blk) : 1.f;
            m_ref = grow ? m_blk : m_ref;
            l *= alpha;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {   // 16 columns at a time: the score row stays in registers meanwhile
              uint32_t o[16];
              tmem_ld_x16(t_o + 16 * c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(t_o + 16 * c, o);
            }
          }
        }
        // remainder query row: this thread holds the score of key r of the block
        if (titem) {
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
              tmem_ld_x1(tmem_base + t_lane + A2_T_OT, o1);
              tmem_ld_wait();
              tmem_st_x1(tmem_base + t_lane + A2_T_OT, __float_as_uint(__uint_as_float(o1) * alpha));
            }
          }
          const float p = fast_exp2(st - m_t);
          l_t += p;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(smem + A2_OFF_PT) + 2 * r),
                       "h"(static_cast<unsigned short>(pack_bf16x2(p, 0.f) & 0xFFFFu)) : "memory");
        }
        // pass 2: P = 2^(S*scale - m_ref) as packed bf16 pairs, written over the S columns (64 words per row): chunk c
        // (score columns [64c, 64c+64)) becomes words [32c, 32c+32), i.e. columns that have already been read.
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (64 * c < valid) {
            uint32_t sv[64];
            {
              uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
              uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
              tmem_ld_x32(t_s + 64 * c, s0);
              tmem_ld_x32(t_s + 64 * c + 32, s1);
            }
            tmem_ld_wait();
            uint32_t pw[32];
            if (64 * c + 64 <= valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), scale_log2, -m_ref));
                const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, -m_ref));
                rs4[i & 3] += p0 + p1;
                pw[i] = pack_bf16x2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int k0 = 64 * c + 2 * i;
                float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), scale_log2, -m_ref));
                float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, -m_ref));
                p0 = (k0 < valid) ? p0 : 0.f;
                p1 = (k0 + 1 < valid) ? p1 : 0.f;
                rs4[i & 3] += p0 + p1;
                pw[i] = pack_bf16x2(p0, p1);
              }
            }
            tmem_st_x32(t_s + 32 * c, pw);
          }
        }
        l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        tmem_st_wait();
        if (titem) fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[A2_B_PREADY + w]);
      }
      // ------------------------------------------------------------------ epilogue: O / l -> bf16 -> smem -> TMA store
      mbar_wait(&bars[A2_B_ODONE + w], cnt_o & 1, 48);
      ++cnt_o;
      tc_fence_after();
      uint32_t o[64];
      {
        uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
        uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
        tmem_ld_x32(t_o, o0);
        tmem_ld_x32(t_o + 32, o1);
      }
      uint32_t u_ot = 0;
      if (titem && quad < 2) tmem_ld_x1(tmem_base + t_lane + A2_T_OT, u_ot);
      tmem_ld_wait();
      const uint32_t rows = smem_u32(smem + A2_OFF_ROWS + buf * 384);
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
      const uint32_t stage = smem_u32(smem + A2_OFF_Q + (2 * buf + w) * A2_TILE);   // this tile's Q buffer is dead by now
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
        if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(out + (row * H + h) * A2_HD + r) = pack_bf16x2(ov, o_hi);
        if (r == 0 && lse_out != nullptr)
          lse_out[(static_cast<long long>(b) * H + h) * L + Lm] = (m_fin + log2f(l_all)) * 0.69314718055994531f;
      }
      if (r == 0) {
        tma_store_4d(&tmO, smem + A2_OFF_Q + (2 * buf + w) * A2_TILE, 0, h, q0, b);
        tma_store_commit();
        tma_store_wait_read<0>();   // the producer may refill this Q buffer
        mbar_arrive(&bars[A2_B_QFREE + buf]);
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
int ovk_attention_fwd2_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s) {
  const int hd = A2_HD;
  CUtensorMap tmQKV, tmO, tmRow;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {A2_HD, 1, A2_BKV, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const uint32_t box1[4] = {A2_HD, 1, 1, 1};   // one head row (128 B), unswizzled: remainder key / value / query
    if ((rc = make_tmap_nd_bf16(&tmRow, qkv, 4, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {A2_HD, 1, A2_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention2): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  const int tail = (L > A2_BQ && L % A2_BQ == 1) ? 1 : 0;   // cls + power-of-two grid: remainder token handled outside the tiles
  const int l_main = L - tail;
  const int nq = (l_main + A2_BQ - 1) / A2_BQ;
  const long long items = static_cast<long long>((nq + 1) / 2) * H * B;
  if (items > 0x7fffffffLL) return set_error(OVK_ERR_SHAPE, "attention: too many work items");
  const int grid = static_cast<int>(items < (long long)num_sms() ? items : (long long)num_sms());
  attention_fwd2_kernel<<<grid, A2_THREADS, A2_SMEM, s>>>(tmQKV, tmO, tmRow, lse, reinterpret_cast<__nv_bfloat16*>(out), L, l_main, H,
                                                         nq, static_cast<int>(items), scale * 1.4426950408889634f);
  return check_launch("attention_fwd2_kernel");
}
