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
constexpr int ATT_SM_WARPS = 8;    // softmax warps: warp w owns TMEM lane quadrant w % 4 and key columns [64*(w/4), +64)
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 2);  // + warp 8 TMA producer, warp 9 MMA issuer / TMEM owner
constexpr int ATT_TILE_BYTES = 128 * 128;  // [128 rows x 64 bf16]
constexpr int ATT_NUM_BARS = 14;
constexpr int ATT_BT = 128 * 32;   // 4 KB narrow tile (head dims 64..79), SWIZZLE_32B
// Shared-memory layout.  hd = 64 (RB = 0): K / V double-buffered, 113 KB -> two CTAs per SM.
// Head widths 64 < hd <= 80 (RB = 16; H/14: 80, So400m/14: 72): the extra dims ride along as a second, narrow operand
// block: [128 rows x 32 B] tiles with 32-byte swizzle (TMA zero-fills dims >= hd), one more k-step in S = Q K^T and a
// second N = 16 accumulator for P V; K / V are single-buffered there so that two CTAs still fit per SM (98 KB).
template <int RB>
struct AttL {
  static constexpr int NS = RB ? 1 : 2;   // K / V ring depth
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + ATT_TILE_BYTES;
  static constexpr int OFF_V = OFF_K + NS * ATT_TILE_BYTES;
  static constexpr int OFF_P = OFF_V + NS * ATT_TILE_BYTES;
  // remainder token (L = 128 k + 1): its key row (128 B), its value row (double-buffered, 2 x 128 B), its query row
  // (128 B, B operand of the S_t^T MMA), the probabilities of the remainder query row (bf16 [128], B operand of the
  // O_t^T MMA; the MMA over-reads up to 128 B past it, into the scratch / barrier bytes) and two 16-byte scratch vectors
  static constexpr int OFF_TAIL = OFF_P + 2 * ATT_TILE_BYTES;
  static constexpr int OFF_QR = OFF_TAIL + 384;
  static constexpr int OFF_PT = OFF_QR + 128;
  static constexpr int OFF_XT = OFF_PT + 256;
  static constexpr int OFF_BAR = OFF_XT + 128;
  static constexpr int BASE_BYTES = OFF_BAR + ATT_NUM_BARS * 8 + 16;
  static constexpr int OFF_QB = (BASE_BYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_KB = OFF_QB + ATT_BT;
  static constexpr int OFF_VB = OFF_KB + NS * ATT_BT;
  static constexpr int OFF_OB = OFF_VB + NS * ATT_BT;   // output staging
  static constexpr int SMEM_BYTES = RB ? OFF_OB + ATT_BT : BASE_BYTES;
};
constexpr uint32_t ATT_TMEM_OB = 192;                              // 16 columns after O
constexpr uint32_t UMMA_SW32 = 6;
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) { return umma_desc(saddr, 16, 256, UMMA_SW32); }
// 16-byte chunk `chunk` (0..1) of row `row` in a [rows x 32 B] SWIZZLE_32B tile
__device__ __forceinline__ uint32_t sw32_offset(uint32_t row, uint32_t chunk) {
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}
constexpr int ATT_MAX_TAIL = 1;     // remainder key / query row (L mod 128 == 1: cls + power-of-two grid) handled outside the tiles
constexpr int ATT_TAIL_MAX_L = 1040;  // longest sequence the remainder-row kernel keeps scores for (16 KB of smem)     // remainder keys / query rows (L mod 128) handled outside the 128-wide tiles
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O: [128,192)  O_b: [192,208)  O_t^T: [208,224)  S_t^T: [224,240)  X: [240,242)
constexpr uint32_t ATT_TMEM_OT = 208;   // remainder query row: O_t^T (lane = head dim, column 0)
constexpr uint32_t ATT_TMEM_ST = 224;   // remainder query row: S_t^T (lane = key of the block, column 0)
constexpr uint32_t ATT_TMEM_X = 240;    // half-row maxima exchanged between the two threads of a row
constexpr uint32_t ATT_TMEM_SK = 192;   // scores of the tile's rows against the remainder key (column 0; hd = 64 only)
constexpr uint32_t ATT_TMEM_S = 0;
constexpr uint32_t ATT_TMEM_O = 128;

template <int RB>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmTail, const __grid_constant__ CUtensorMap tmQKVb,
                     const __grid_constant__ CUtensorMap tmOb, float* __restrict__ lse_out,
                     __nv_bfloat16* __restrict__ out, int tail_row_fused, int L, int Lm, int H, int nq, int total_items,
                     float scale_log2, int causal) {
  // PERSISTENT: each CTA walks work items (query tile, head, image) with stride gridDim.x, keeping its TMEM allocation,
  // barriers and the K/V TMA ring alive across items, so the next item's Q/K/V loads run under the current item's
  // softmax.  Query rows [0, nq * 128) are handled here; a short remainder of rows goes to attention_tail_kernel.
  using LL = AttL<RB>;
  constexpr int NS = LL::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + LL::OFF_Q;
  uint8_t* sK = smem + LL::OFF_K;
  uint8_t* sV = smem + LL::OFF_V;
  uint8_t* sP = smem + LL::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LL::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* pv_done = bars + 11;
  uint64_t* q_empty = bars + 12;
  uint64_t* o_free = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + LL::OFF_BAR + ATT_NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  // keys [0, Lm) go through the tensor-core blocks; the (few) keys [Lm, L) are folded in on the FMA pipe (see below)
  const int nkv = (Lm + ATT_BKV - 1) / ATT_BKV;
  const int ntail = L - Lm;
  // The remainder QUERY row (row Lm) rides along with the last query tile of its head (hd = 64 only): its scores come
  // out of one more small MMA per key block in TRANSPOSED form, S_t^T = K_j q_t^T (M = 128 keys, N = 16 of which one
  // column is real), so each softmax thread of column-half 0 holds the score of one key; the probabilities go back as
  // a 256-byte vector and O_t^T = V_j^T p_t^T (M = head dims, N = 16) accumulates in 16 more TMEM columns.
  const bool tail_q = RB == 0 && ntail > 0 && tail_row_fused != 0;

  if (warp == ATT_SM_WARPS && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmTail);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1 + (tail_q ? 2 : 0));  // MMA commit (+ two softmax warps' read of the remainder query / key rows)
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 32 * ATT_SM_WARPS);
    mbar_init(pv_done, 1);
    mbar_init(o_free, 32 * ATT_SM_WARPS);
    fence_mbar_init();
  }
  if (warp == ATT_SM_WARPS + 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == ATT_SM_WARPS) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      int n = 0, g = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int qt = item % nq, h = (item / nq) % H, b = item / (nq * H);
        mbar_wait(q_empty, (n & 1) ^ 1, 9);
        const bool titem = tail_q && qt == nq - 1;
        mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES + (ntail > 0 ? 256 : 0) + (titem ? 128 : 0) + (RB ? ATT_BT : 0));
        tma_load_4d(sQ, &tmQKV, q_full, 0, h, qt * ATT_BQ, b);
        if (RB) tma_load_4d(smem + LL::OFF_QB, &tmQKVb, q_full, 64, h, qt * ATT_BQ, b);
        if (ntail > 0) {  // remainder key row and its value row (1 x 128 B each, unswizzled)
          tma_load_4d(smem + LL::OFF_TAIL, &tmTail, q_full, 0, H + h, Lm, b);
          tma_load_4d(smem + LL::OFF_TAIL + 128 + (n & 1) * 128, &tmTail, q_full, 0, 2 * H + h, Lm, b);
          if (titem) tma_load_4d(smem + LL::OFF_QR, &tmTail, q_full, 0, h, Lm, b);
        }
        for (int j = 0; j < nkv; ++j, ++g) {
          const int s = g % NS;
          const uint32_t ph = (g / NS) & 1;
          mbar_wait(&k_empty[s], ph ^ 1, 10);
          mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES + (RB ? ATT_BT : 0));
          tma_load_4d(sK + s * ATT_TILE_BYTES, &tmQKV, &k_full[s], 0, H + h, j * ATT_BKV, b);
          if (RB) tma_load_4d(smem + LL::OFF_KB + s * ATT_BT, &tmQKVb, &k_full[s], 64, H + h, j * ATT_BKV, b);
          mbar_wait(&v_empty[s], ph ^ 1, 11);
          mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES + (RB ? ATT_BT : 0));
          tma_load_4d(sV + s * ATT_TILE_BYTES, &tmQKV, &v_full[s], 0, 2 * H + h, j * ATT_BKV, b);
          if (RB) tma_load_4d(smem + LL::OFF_VB + s * ATT_BT, &tmQKVb, &v_full[s], 64, 2 * H + h, j * ATT_BKV, b);
        }
      }
    }
  } else if (warp == ATT_SM_WARPS + 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      // Descriptors as `base low word + constant`: building both 64-bit descriptors of every MMA from addresses costs ~10
      // uniform-datapath instructions each, which for MMAs this small is more than their tensor time (measured in the attention
      // backward: profiles/r02_attention_bwd_analysis.md).  The address field counts 16-byte units and never carries out.
      constexpr uint64_t HI128 = 0x40004040ull << 32;   // SWIZZLE_128B, SBO 1024, descriptor version 1
      constexpr uint64_t HI32 = 0xC0004010ull << 32;    // SWIZZLE_32B, SBO 256
      constexpr uint64_t HIROW0 = 0x00004001ull << 32;  // no swizzle, SBO 16 (umma_desc_row0)
      constexpr uint32_t LBO16 = 1u << 16;
      constexpr uint32_t LBO_MN = (ATT_TILE_BYTES >> 4) << 16;
      auto d128 = [](uint32_t lo, uint32_t bytes) { return HI128 | (lo + (bytes >> 4)); };
      auto d32 = [](uint32_t lo, uint32_t bytes) { return HI32 | (lo + (bytes >> 4)); };
      auto drow0 = [](uint32_t lo, uint32_t bytes) { return HIROW0 | (lo + (bytes >> 4)); };
      const uint32_t q_lo = ((smem_u32(sQ) & 0x3FFFF) >> 4) | LBO16;
      const uint32_t p_lo = ((smem_u32(sP) & 0x3FFFF) >> 4) | LBO16;
      const uint32_t qb_lo = ((smem_u32(smem + LL::OFF_QB) & 0x3FFFF) >> 4) | LBO16;
      const uint32_t kt_lo = ((smem_u32(smem + LL::OFF_TAIL) & 0x3FFFF) >> 4) | LBO16;
      const uint32_t qr_lo = ((smem_u32(smem + LL::OFF_QR) & 0x3FFFF) >> 4) | LBO16;
      const uint32_t pt_lo = ((smem_u32(smem + LL::OFF_PT) & 0x3FFFF) >> 4) | LBO16;
      int n = 0, g = 0;
      // Issue order per item: S(0) | S(1) P V(0) | S(2) P V(1) | ... | P V(last).  S(j+1) goes out BEFORE P V(j): both
      // wait for the same event (the softmax warps are done with S(j)), and the softmax warps need S(j+1) first —
      // P V(j) only has to be finished before they overwrite P again.
      auto issue_s = [&](int j, int gg, bool titem) {
        const int s = gg % NS;
        const int valid = min(ATT_BKV, Lm - j * ATT_BKV);
        const int nblk = (valid + 15) & ~15;  // MMA N of S and K-extent of PV for this block
        const uint32_t k_lo = ((smem_u32(sK + s * ATT_TILE_BYTES) & 0x3FFFF) >> 4) | LBO16;
        mbar_wait(&k_full[s], (gg / NS) & 1, 13);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, nblk, 0, 0);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) umma_bf16_ss(tmem_base + ATT_TMEM_S, d128(q_lo, k * 32), d128(k_lo, k * 32), idesc_s, k != 0);
        if (RB)   // dims 64 .. 64 + RB: one more k-step from the narrow tiles
          umma_bf16_ss(tmem_base + ATT_TMEM_S, d32(qb_lo, 0), d32(((smem_u32(smem + LL::OFF_KB + s * ATT_BT) & 0x3FFFF) >> 4) | LBO16, 0),
                       idesc_s, 1);
        if (RB == 0 && ntail > 0 && j == 0) {   // scores against the remainder key: s_k = Q k_t^T (N = 16, column 0 is real)
          constexpr uint32_t idesc_sk = umma_idesc_bf16(128, 16, 0, 0);
#pragma unroll
          for (int k = 0; k < ATT_HD / 16; ++k) umma_bf16_ss(tmem_base + ATT_TMEM_SK, d128(q_lo, k * 32), drow0(kt_lo, k * 32), idesc_sk, k != 0);
        }
        if (titem) {   // S_t^T = K_j q_t^T
          constexpr uint32_t idesc_st = umma_idesc_bf16(128, 16, 0, 0);
#pragma unroll
          for (int k = 0; k < ATT_HD / 16; ++k) umma_bf16_ss(tmem_base + ATT_TMEM_ST, d128(k_lo, k * 32), drow0(qr_lo, k * 32), idesc_st, k != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
        if (j == nkv - 1) umma_commit(q_empty);  // Q tile no longer needed: the producer may fetch the next item's
      };
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const bool titem = tail_q && item % nq == nq - 1;
        mbar_wait(q_full, n & 1, 12);
        issue_s(0, g, titem);   // the S columns are free: p_ready of the previous item's last block was waited on below
        for (int j = 0; j < nkv; ++j, ++g) {
          const int s = g % NS;
          const uint32_t ph = (g / NS) & 1;
          const int valid = min(ATT_BKV, Lm - j * ATT_BKV);
          const int nblk = (valid + 15) & ~15;
          const uint32_t v_addr = smem_u32(sV + s * ATT_TILE_BYTES);
          mbar_wait(p_ready, g & 1, 14);    // softmax(j) done: S may be overwritten, P(j) is in shared memory
          if (j + 1 < nkv) issue_s(j + 1, g + 1, titem);
          // O += P V_j   (P: K-major [128 x nblk] in two 64-column swizzle atoms; V: MN-major [nblk x 64])
          mbar_wait(&v_full[s], ph, 15);
          if (j == 0) mbar_wait(o_free, (n & 1) ^ 1, 8);  // previous item's epilogue has read the O accumulator
          tc_fence_after();
          constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_HD, 0, 1);
          const int ksteps = nblk / 16;
          const uint32_t v_lo = ((v_addr & 0x3FFFF) >> 4) | LBO_MN;
          if (ksteps == 8) {   // full key block: straight-line
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ss(tmem_base + ATT_TMEM_O, d128(p_lo, (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32), d128(v_lo, kk * 2048), idesc_pv,
                           (j | kk) != 0);
          } else {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16_ss(tmem_base + ATT_TMEM_O, d128(p_lo, (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32), d128(v_lo, kk * 2048), idesc_pv,
                           (j | kk) != 0);
          }
          if (RB) {   // output dims 64 .. 64 + RB: N = 16 accumulator, V_b consumed MN-major from its 32-byte rows
            constexpr uint32_t idesc_pvb = umma_idesc_bf16(ATT_BQ, 16, 0, 1);
            const uint32_t vb_lo = ((smem_u32(smem + LL::OFF_VB + s * ATT_BT) & 0x3FFFF) >> 4) | LBO16;
            if (ksteps == 8) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma_bf16_ss(tmem_base + ATT_TMEM_OB, d128(p_lo, (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32), d32(vb_lo, kk * 512), idesc_pvb,
                             (j | kk) != 0);
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16_ss(tmem_base + ATT_TMEM_OB, d128(p_lo, (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32), d32(vb_lo, kk * 512), idesc_pvb,
                             (j | kk) != 0);
            }
          }
          if (titem) {   // O_t^T += V_j^T p_t^T  (A = V MN-major: M runs over head dims; rows 64..127 are padding)
            constexpr uint32_t idesc_ot = umma_idesc_bf16(128, 16, 1, 0);
#pragma unroll
            for (int kk = 0; kk < ATT_BKV / 16; ++kk)
              umma_bf16_ss(tmem_base + ATT_TMEM_OT, d128(v_lo, kk * 2048), drow0(pt_lo, kk * 32), idesc_ot, (j | kk) != 0);
          }
          umma_commit(&v_empty[s]);
          umma_commit(pv_done);
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warps
    // Two threads per query row: thread (quad, lane, half) owns TMEM lane 32*quad + lane and key columns
    // [64*half, 64*half + 64) of the block, plus output columns [32*half, +32).  One TMEM pass per key block (the 64
    // scores stay in registers between the max and the exp); the two halves agree on the exponent reference through a
    // 512-byte exchange.  Lazy rescaling: the reference only moves when the row maximum grew by more than 2^8, so the
    // O accumulator is rescaled (TMEM load / multiply / store) rarely instead of once per key block.
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int r = quad * 32 + lane;  // row in tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t p_base = smem_u32(sP) + half * ATT_TILE_BYTES;
    const uint32_t xt = smem_u32(smem + LL::OFF_XT);   // [0,16): warp maxima, [16,32): warp sums of the remainder row
    constexpr uint32_t SMT = 32 * ATT_SM_WARPS;
    int g = 0, n = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
      const int qt = item % nq, h = (item / nq) % H, b = item / (nq * H);
      const int q0 = qt * ATT_BQ;
      const bool rows_live = quad * 32 < L - q0;   // warp-uniform (a last tile of 1 row leaves 3 of 4 lane quadrants idle)
      // the previous item's output tile (staged in the P region) must have left shared memory
      if (threadIdx.x == 0) tma_store_wait_read<0>();
      named_bar_sync(1, SMT);
      float m_ref = -INFINITY;  // exponent reference (log2 domain, already scaled), identical in both halves of a row
      float l = 0.f;            // this half's running sum of 2^(s - m_ref)
      // Remainder key (the 257th token): its scores against the tile's rows come out of a small N = 16 MMA issued with
      // S(0) (one TMEM word per thread), and it is folded into the output in the epilogue — instead of a whole extra
      // TMA -> MMA -> softmax -> MMA round for a 1-column block.
      float s_tail = 0.f;
      // remainder query row (handled by the threads of column-half 0): reference, partial sum, score against the remainder key
      const bool titem = tail_q && qt == nq - 1 && half == 0;
      float m_t = -INFINITY, l_t = 0.f, s_tt = 0.f;
      if (tail_q && half == 0 && quad < 2) {   // q_t . k_t, two dims per lane (the warps that finish the remainder row)
        mbar_wait(q_full, n & 1, 19);
        if (titem) {
          uint32_t qa, ka;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(qa) : "r"(smem_u32(smem + LL::OFF_QR) + 4 * lane));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ka) : "r"(smem_u32(smem + LL::OFF_TAIL) + 4 * lane));
          float e = fmaf(bf16_lo(qa), bf16_lo(ka), bf16_hi(qa) * bf16_hi(ka));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
          s_tt = e * scale_log2;
        }
        if (lane == 0) mbar_arrive(q_empty);
      }
      for (int j = 0; j < nkv; ++j, ++g) {
        // valid columns of this half (may be <= 0).  Causal (the text tower's additive mask, transformer.py:757-763: -inf
        // above the diagonal): row q0 + r sees keys <= q0 + r, i.e. a per-thread column count; masked columns contribute 0
        // to P and to the row sum exactly like the columns past the end of the sequence.
        const int valid_u = min(ATT_BKV, Lm - j * ATT_BKV) - 64 * half;
        const int valid = causal ? min(valid_u, q0 + r - j * ATT_BKV - 64 * half + 1) : valid_u;
        const int nblk = ((min(ATT_BKV, Lm - j * ATT_BKV) + 15) & ~15) - 64 * half;
        mbar_wait(s_full, g & 1, 16);
        tc_fence_after();
        if (!rows_live) {   // all 32 rows of this warp lie past the end of the sequence: nothing they produce is stored
          named_bar_sync(2, SMT);
          tc_fence_before();
          mbar_arrive(p_ready);
          continue;
        }
        if (RB == 0 && ntail > 0 && j == 0) {   // this row against the remainder key (out of the small s_k MMA)
          uint32_t u;
          tmem_ld_x1(tmem_base + t_lane + ATT_TMEM_SK, u);
          tmem_ld_wait();
          s_tail = __uint_as_float(u) * scale_log2;
        }
        // pass 1: half-row maximum (two 32-column TMEM loads; the scores are re-read in pass 2 to stay within the
        // 96-register budget that two resident CTAs of 320 threads allow)
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          if (c < nblk) {
            uint32_t sv[32];
            tmem_ld_x32(tmem_base + t_lane + ATT_TMEM_S + 64 * half + c, sv);
            tmem_ld_wait();
            if (c + 32 <= valid) {
              float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
              for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[i]));
              mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else if (c < valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = (c + i < valid) ? fmaxf(mx, __uint_as_float(sv[i])) : mx;
            }
          }
        }
        float st_raw = 0.f;
        if (titem) {   // this thread's key (row r of the block) against the remainder query row
          uint32_t u;
          tmem_ld_x1(tmem_base + t_lane + ATT_TMEM_ST, u);
          tmem_ld_wait();
          st_raw = __uint_as_float(u) * scale_log2;
          float wm = st_raw;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
          if (lane == 0) sts_f32(xt + quad * 4, wm);
        }
        // exchange the half-row maxima through two spare TMEM columns of the row's lane (both threads of a row must
        // derive the SAME reference; any reference >= the true maximum - 2^8 is valid)
        tmem_st_x1(tmem_base + t_lane + ATT_TMEM_X + half, __float_as_uint(mx * scale_log2));
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(2, SMT);
        tc_fence_after();
        float m_tile;
        {
          uint32_t a, bb;
          tmem_ld_x2(tmem_base + t_lane + ATT_TMEM_X, a, bb);
          tmem_ld_wait();
          m_tile = fmaxf(__uint_as_float(a), __uint_as_float(bb));
        }
        // P buffer and the O accumulator are owned by PV(j-1) until it completes
        if (j > 0) {
          mbar_wait(pv_done, (g - 1) & 1, 17);
          tc_fence_after();
        }
        if (j == 0) {
          m_ref = m_tile;
        } else {
          const bool grow = m_tile > m_ref + 8.f;
          if (__any_sync(0xffffffffu, grow)) {  // rescale the running output of the rows whose reference moves
            const float alpha = grow ? fast_exp2(m_ref - m_tile) : 1.f;
            m_ref = grow ? m_tile : m_ref;
            l *= alpha;
            uint32_t o[32];
            tmem_ld_x32(tmem_base + t_lane + ATT_TMEM_O + 32 * half, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tmem_base + t_lane + ATT_TMEM_O + 32 * half, o);
            if (RB && half == 0) {
              uint32_t ob[16];
              tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_OB, ob);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ob[i] = __float_as_uint(__uint_as_float(ob[i]) * alpha);
              tmem_st_x16(tmem_base + t_lane + ATT_TMEM_OB, ob);
            }
            tmem_st_wait();
          }
        }
        if (titem) {   // online-softmax step of the remainder query row over this block's 128 keys (one per thread)
          const float4 w4 = lds_f32x4(xt);
          const float mt = fmaxf(fmaxf(w4.x, w4.y), fmaxf(w4.z, w4.w));
          if (j == 0) {
            m_t = mt;
          } else if (mt > m_t + 8.f) {   // uniform over the 128 threads
            const float alpha = fast_exp2(m_t - mt);
            m_t = mt;
            l_t *= alpha;
            if (quad < 2) {              // O_t^T lives in lanes 0..63 (head dims), column 0
              uint32_t o1;
              tmem_ld_x1(tmem_base + t_lane + ATT_TMEM_OT, o1);
              tmem_ld_wait();
              tmem_st_x1(tmem_base + t_lane + ATT_TMEM_OT, __float_as_uint(__uint_as_float(o1) * alpha));
              tmem_st_wait();
            }
          }
          const float p = fast_exp2(st_raw - m_t);
          l_t += p;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(smem + LL::OFF_PT) + 2 * r),
                       "h"(static_cast<unsigned short>(pack_bf16x2(p, 0.f) & 0xFFFFu)) : "memory");
        }
        // pass 2: P = 2^(S*scale - m_ref) -> bf16 -> swizzled smem (this half's 64-column atom); row sum in fp32
        float rowsum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          if (c < nblk) {
            uint32_t sv[32];
            tmem_ld_x32(tmem_base + t_lane + ATT_TMEM_S + 64 * half + c, sv);
            tmem_ld_wait();
            if (c + 32 <= valid) {   // full chunk (the common case): no per-element masking
              // packed fp32 pairs (FFMA2 / FADD2): half the FMA-pipe instructions of the chunk
              const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(-m_ref, -m_ref);
              uint64_t rs2[2] = {0ull, 0ull};   // two independent (even, odd) running sums
#pragma unroll
              for (int cc = 0; cc < 32; cc += 8) {
                float pv[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float lo, hi;
                  f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[cc + 2 * i]), __uint_as_float(sv[cc + 2 * i + 1])), sc2, nm2), lo, hi);
                  pv[2 * i] = fast_exp2(lo);
                  pv[2 * i + 1] = fast_exp2(hi);
                  rs2[i & 1] = f2_add(rs2[i & 1], f2_pack(pv[2 * i], pv[2 * i + 1]));
                }
                sts128(p_base + sw128_offset(r, (c + cc) >> 3),
                       make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7])));
              }
              {
                float lo, hi;
                f2_unpack(f2_add(rs2[0], rs2[1]), lo, hi);
                rowsum += lo + hi;
              }
            } else {
#pragma unroll
              for (int cc = 0; cc < 32; cc += 8) {
                if (c + cc < nblk) {
                  float pv[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float e = fast_exp2(fmaf(__uint_as_float(sv[cc + i]), scale_log2, -m_ref));
                    pv[i] = (c + cc + i < valid) ? e : 0.f;
                    rowsum += pv[i];
                  }
                  sts128(p_base + sw128_offset(r, (c + cc) >> 3),
                         make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7])));
                }
              }
            }
          }
        }
        l += rowsum;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_ready);
      }
      // ------------------------------------------------------------------ epilogue: O / l -> bf16 -> smem -> TMA store
      if (!rows_live) {   // keep the barrier protocol, skip the work (these rows are clipped by the TMA store)
        mbar_arrive(o_free);
        named_bar_sync(2, SMT);
        named_bar_sync(1, SMT);
        continue;
      }
      mbar_wait(pv_done, (g - 1) & 1, 18);
      tc_fence_after();
      uint32_t o[32];
      uint32_t ob[16];
      tmem_ld_x32(tmem_base + t_lane + ATT_TMEM_O + 32 * half, o);
      if (RB && half == 0) tmem_ld_x16(tmem_base + t_lane + ATT_TMEM_OB, ob);
      uint32_t ot = 0;
      if (titem && quad < 2) tmem_ld_x1(tmem_base + t_lane + ATT_TMEM_OT, ot);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);  // the accumulator may be overwritten by the next item's first P V
      if (titem) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l_t += __shfl_xor_sync(0xffffffffu, l_t, o);
        if (lane == 0) sts_f32(xt + 16 + quad * 4, l_t);
      }
      if (ntail > 0) {      // fold the remainder key in: one more online-softmax step, entirely in registers
        const uint32_t vt = smem_u32(smem + LL::OFF_TAIL + 128 + (n & 1) * 128) + 64 * half;
        const float m_fin = fmaxf(m_ref, s_tail);
        const float a = fast_exp2(m_ref - m_fin);
        const float pt = fast_exp2(s_tail - m_fin);
        l = fmaf(l, a, half == 0 ? pt : 0.f);
        m_ref = m_fin;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 vv = lds128(vt + c * 16);
          const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[8 * c + 2 * q] = __float_as_uint(fmaf(__uint_as_float(o[8 * c + 2 * q]), a, pt * bf16_lo(w4[q])));
            o[8 * c + 2 * q + 1] = __float_as_uint(fmaf(__uint_as_float(o[8 * c + 2 * q + 1]), a, pt * bf16_hi(w4[q])));
          }
        }
      }
      // combine the two halves' sums through the (now idle) second P atom
      const uint32_t lx = smem_u32(sP) + ATT_TILE_BYTES;
      sts_f32(lx + (half * 128 + r) * 4, l);
      named_bar_sync(2, SMT);
      const float l_tot = lds_f32(lx + r * 4) + lds_f32(lx + (128 + r) * 4);
      const float inv_l = 1.f / l_tot;
      if (titem && quad < 2) {   // remainder query row: fold its remainder key in, normalise, store head dim r
        const float4 w4 = lds_f32x4(xt + 16);
        const float m_fin = fmaxf(m_t, s_tt);
        const float a = fast_exp2(m_t - m_fin);
        const float pt = fast_exp2(s_tt - m_fin);
        const float l_all = fmaf((w4.x + w4.y) + (w4.z + w4.w), a, pt);
        unsigned short vb;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(vb) : "r"(smem_u32(smem + LL::OFF_TAIL + 128 + (n & 1) * 128) + 2 * r));
        const float v = __uint_as_float(static_cast<uint32_t>(vb) << 16);
        const float o = fmaf(__uint_as_float(ot), a, pt * v) / l_all;
        const float o_hi = __shfl_down_sync(0xffffffffu, o, 1);
        const long long row = static_cast<long long>(b) * L + Lm;
        if ((lane & 1) == 0)
          *reinterpret_cast<uint32_t*>(out + (row * H + h) * ATT_HD + r) = pack_bf16x2(o, o_hi);
        if (r == 0 && lse_out != nullptr)
          lse_out[(static_cast<long long>(b) * H + h) * L + Lm] = (m_fin + log2f(l_all)) * 0.69314718055994531f;
      }
      // staging in the first P atom (every MMA that read P has completed: pv_done)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts128(smem_u32(sP) + sw128_offset(r, 4 * half + c),
               make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l)));
      if (half == 0 && lse_out != nullptr && q0 + r < L) {
        lse_out[(static_cast<long long>(b) * H + h) * L + q0 + r] = (m_ref + log2f(l_tot)) * 0.69314718055994531f;
      }
      if (RB && half == 0) {
        const uint32_t obs = smem_u32(smem + LL::OFF_OB);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          sts128(obs + sw32_offset(r, c),
                 make_uint4(pack_bf16x2(__uint_as_float(ob[8 * c]) * inv_l, __uint_as_float(ob[8 * c + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 2]) * inv_l, __uint_as_float(ob[8 * c + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 4]) * inv_l, __uint_as_float(ob[8 * c + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 6]) * inv_l, __uint_as_float(ob[8 * c + 7]) * inv_l)));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, SMT);
      if (threadIdx.x == 0) {
        tma_store_4d(&tmO, sP, 0, h, q0, b);
        if (RB) tma_store_4d(&tmOb, smem + LL::OFF_OB, 64, h, q0, b);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ATT_SM_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ remainder rows
// The (few) query rows left over after the last full 128-row tile (L = 257 = 2*128 + 1 for a 16x16 grid + cls) are a
// [1 x L] x [L x 64] product per head: one warp per (row, head, image) on the FMA pipe, instead of a tcgen05 tile that
// would be 99% padding.  K / V rows are read coalesced: lane l holds 16-byte chunk (l % 8) of key row 4*step + l / 8,
// i.e. 512 contiguous-per-row bytes per load instruction; the 8 lanes of a row reduce their partial dot by shuffles.
__global__ void __launch_bounds__(128)
attention_tail_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out,
                      int L, int H, int row0, int nrows, float scale_log2) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= nrows * H) return;
  const int lane = threadIdx.x & 31;
  const int ch = lane & 7;    // 16-byte chunk of the 128-byte head row
  const int sub = lane >> 3;  // which of the 4 rows of a step
  const int row = row0 + w / H;
  const int h = w % H;
  const long long b = blockIdx.y;
  const long long tok = 3LL * H * ATT_HD;  // elements per token in the packed qkv tensor
  const __nv_bfloat16* base = qkv + b * L * tok + h * ATT_HD + ch * 8;
  float q[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + row * tok));
    q[0] = bf16_lo(u.x); q[1] = bf16_hi(u.x); q[2] = bf16_lo(u.y); q[3] = bf16_hi(u.y);
    q[4] = bf16_lo(u.z); q[5] = bf16_hi(u.z); q[6] = bf16_lo(u.w); q[7] = bf16_hi(u.w);
  }
  const int steps = (L + 3) / 4;
  // pass 1: scores (kept in shared memory: one float per key), running maximum
  __shared__ float sc[4][ATT_TAIL_MAX_L];
  float* my = sc[threadIdx.x >> 5];
  float mx = -INFINITY;
  const __nv_bfloat16* kb = base + H * ATT_HD;
#pragma unroll 4
  for (int t = 0; t < steps; ++t) {
    const int j = 4 * t + sub;
    float d = 0.f;
    if (j < L) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(kb + j * tok));
      d = fmaf(q[0], bf16_lo(u.x), d); d = fmaf(q[1], bf16_hi(u.x), d); d = fmaf(q[2], bf16_lo(u.y), d);
      d = fmaf(q[3], bf16_hi(u.y), d); d = fmaf(q[4], bf16_lo(u.z), d); d = fmaf(q[5], bf16_hi(u.z), d);
      d = fmaf(q[6], bf16_lo(u.w), d); d = fmaf(q[7], bf16_hi(u.w), d);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    d = (j < L) ? d * scale_log2 : -INFINITY;
    if (ch == 0 && j < L) my[j] = d;
    mx = fmaxf(mx, d);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  __syncwarp();
  // pass 2: P V with the same coalesced addressing; each lane accumulates its 8 output dims for its row subset
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float l = 0.f;
  const __nv_bfloat16* vb = base + 2 * H * ATT_HD;
#pragma unroll 4
  for (int t = 0; t < steps; ++t) {
    const int j = 4 * t + sub;
    if (j < L) {
      const float p = fast_exp2(my[j] - mx);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(vb + j * tok));
      l += p;
      acc[0] = fmaf(p, bf16_lo(u.x), acc[0]); acc[1] = fmaf(p, bf16_hi(u.x), acc[1]);
      acc[2] = fmaf(p, bf16_lo(u.y), acc[2]); acc[3] = fmaf(p, bf16_hi(u.y), acc[3]);
      acc[4] = fmaf(p, bf16_lo(u.z), acc[4]); acc[5] = fmaf(p, bf16_hi(u.z), acc[5]);
      acc[6] = fmaf(p, bf16_lo(u.w), acc[6]); acc[7] = fmaf(p, bf16_hi(u.w), acc[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  l += __shfl_xor_sync(0xffffffffu, l, 8);    // every chunk lane of a row group carries the same p: sum over the 4 groups
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  const float inv = 1.f / l;
  if (sub == 0) {
    *reinterpret_cast<uint4*>(out + (b * L + row) * (static_cast<long long>(H) * ATT_HD) + h * ATT_HD + ch * 8) =
        make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                   pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  }
  if (lse_out != nullptr && lane == 0) lse_out[(b * H + h) * L + row] = (mx + log2f(l)) * 0.69314718055994531f;
}

}  // namespace ovk

using namespace ovk;

int ovk_attention_fwd2_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s);
int ovk_attention_fwd3_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s);
int ovk_attention_fwd4_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s);

extern "C" int ovk_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale,
                                 void* stream) {
  return ovk_attention_fwd_ex(qkv, out, lse, B, L, H, hd, scale, 0, stream);
}

extern "C" int ovk_attention_fwd_ex(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale,
                                    int flags, void* stream) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(OVK_ERR_SHAPE, "attention: empty problem");
  if (flags & ~OVK_ATT_CAUSAL) return set_error(OVK_ERR_SHAPE, "attention: unknown flags 0x%x", flags);
  const int causal = (flags & OVK_ATT_CAUSAL) ? 1 : 0;
  if (hd < 64 || hd > 80 || (hd % 8))
    return set_error(OVK_ERR_SHAPE, "attention: head dim %d not supported (64, 72 or 80)", hd);
  const bool ext = hd > ATT_HD;   // 64 < hd <= 80: extra 16-dim operand block (zero-filled past hd by TMA)
  if (B > 65535 || H > 65535) return set_error(OVK_ERR_SHAPE, "attention: B and H must be <= 65535");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // head width 64 and an even number of 128-row query tiles (L = 257: two): the pair kernel of attention2.cu, which fetches
  // K / V once for two query tiles.  With an odd tile count one of its two softmax groups would idle for a whole item, and
  // this file's kernel (one tile per CTA, two CTAs per SM) is faster.  OVK_ATT_V1=1 forces this file's kernel (A/B runs).
  const char* att_env = getenv("OVK_ATT_V1");   // read per call so that tools/ab_probe.py can alternate the two kernels
  const bool force_v1 = att_env != nullptr && att_env[0] == '1';
  {
    const int t2 = (L > ATT_BQ && L % ATT_BQ == 1) ? 1 : 0;
    const int nq2 = (L - t2 + ATT_BQ - 1) / ATT_BQ;
    const char* odd_env = getenv("OVK_ATT_PAIR_ODD");   // =1: the pair kernels also for an odd tile count (tests of that path)
    const bool pair_odd = odd_env != nullptr && odd_env[0] == '1';
    if (!ext && !force_v1 && !causal && L <= 4096 && (nq2 % 2 == 0 || pair_odd)) {
      // attention4.cu (half-block double buffering, one MMA warp per tile slot, cross-item prefetch, deferred epilogue) by
      // default; OVK_ATT_VER=3 / 2 (or the older OVK_ATT_V2=1) select the earlier generations for A/B measurements
      const char* ver = getenv("OVK_ATT_VER");
      const char* v2 = getenv("OVK_ATT_V2");
      if ((v2 != nullptr && v2[0] == '1') || (ver != nullptr && ver[0] == '2')) return ovk_attention_fwd2_launch(qkv, out, lse, B, L, H, scale, s);
      if (ver != nullptr && ver[0] == '3') return ovk_attention_fwd3_launch(qkv, out, lse, B, L, H, scale, s);
      return ovk_attention_fwd4_launch(qkv, out, lse, B, L, H, scale, s);
    }
  }
  CUtensorMap tmQKV, tmO, tmTail, tmQKVb, tmOb;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {ATT_HD, 1, ATT_BKV, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const uint32_t box1[4] = {ATT_HD, 1, 1, 1};   // one head row (128 B), unswizzled: the remainder key / value
    if ((rc = make_tmap_nd_bf16(&tmTail, qkv, 4, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    const uint32_t boxb[4] = {16, 1, ATT_BKV, 1};  // dims 64..79 as [128 rows x 32 B] tiles
    if ((rc = make_tmap_nd_bf16(&tmQKVb, qkv, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {ATT_HD, 1, ATT_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const uint32_t boxb[4] = {16, 1, ATT_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmOb, out, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttL<0>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttL<16>::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  // a short remainder of query rows goes to the FMA-pipe tail kernel instead of a padded 128-row tile (hd = 64 only)
  int tail = L % ATT_BQ;
  if (ext || causal || L <= ATT_BQ || tail > ATT_MAX_TAIL || L > ATT_TAIL_MAX_L) tail = 0;   // (the remainder-token paths are unmasked)
  const int l_main = L - tail;
  const int nq = (l_main + ATT_BQ - 1) / ATT_BQ;
  const long long items = static_cast<long long>(nq) * H * B;
  if (items > 0x7fffffffLL) return set_error(OVK_ERR_SHAPE, "attention: too many work items");
  // the remainder query row rides inside the main kernel (OVK_ATT_TAIL_KERNEL=1 selects the stand-alone FMA-pipe kernel)
  static const bool tail_kernel = [] { const char* e = getenv("OVK_ATT_TAIL_KERNEL"); return e != nullptr && e[0] == '1'; }();
  const int fused_tail = (tail > 0 && !tail_kernel) ? 1 : 0;
  const int per_sm = 2;
  const int grid = static_cast<int>(items < (long long)per_sm * num_sms() ? items : (long long)per_sm * num_sms());
  if (ext)
    attention_fwd_kernel<16><<<grid, ATT_THREADS, AttL<16>::SMEM_BYTES, s>>>(tmQKV, tmO, tmTail, tmQKVb, tmOb, lse,
                                                                          reinterpret_cast<__nv_bfloat16*>(out), 0, L, l_main, H, nq,
                                                                          static_cast<int>(items), scale * 1.4426950408889634f, causal);
  else
    attention_fwd_kernel<0><<<grid, ATT_THREADS, AttL<0>::SMEM_BYTES, s>>>(tmQKV, tmO, tmTail, tmQKVb, tmOb, lse,
                                                                      reinterpret_cast<__nv_bfloat16*>(out), fused_tail, L, l_main, H, nq,
                                                                      static_cast<int>(items), scale * 1.4426950408889634f, causal);
  if ((rc = check_launch("attention_fwd_kernel"))) return rc;
  if (tail && !fused_tail) {
    dim3 tgrid((tail * H + 3) / 4, B);
    attention_tail_kernel<<<tgrid, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                reinterpret_cast<__nv_bfloat16*>(out), lse, L, H, l_main, tail,
                                                scale * 1.4426950408889634f);
    return check_launch("attention_tail_kernel");
  }
  return OVK_OK;
}
