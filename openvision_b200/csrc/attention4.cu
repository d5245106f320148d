// Multi-head self-attention forward, head width 64, fourth-generation kernel: O = softmax(Q K^T * scale) V.
// Replaces the core of nn.MultiheadAttention(need_weights=False, attn_mask=None) between in_proj and out_proj
// (open_clip/transformer.py:225,239-252); online softmax as in src/models/bpt.py:105-124.
//
// attention3.cu (64-key half blocks, two score buffers per tile, hand-pipelined exponentials) with the hand-offs that its
// profile still showed exposed (its stall profile: 11 % of a softmax warp's time waiting for the first S of
// an item, 11 % for the last P V) taken off the critical path:
//   * every tile slot is an independent pipeline with its OWN MMA-issuing warp (warps 9 and 10): the steps of a slot —
//     (item, half block) pairs — are numbered t = 0, 1, 2, ... across items; S(t+2) is issued right behind P V(t) whether or
//     not it belongs to the same item, into score buffer t & 1.  The first scores of the next item are therefore in TMEM
//     before the current item's last exponentials are done.
//   * the epilogue of an item (read O, normalise, store) is deferred until the softmax group has produced P for the first
//     half block of its NEXT item: the latency of the last P V hides behind that work, and P V'(0), which overwrites O, is
//     only released (p_ready) after the epilogue has read it.
//   warps 0-3 : softmax group A (thread <-> query row of tile A)      warp 8 : TMA producer
//   warps 4-7 : softmax group B                                        warps 9 / 10 : MMA issuers of slots A / B
// Barrier b of {s_full, p_ready, pv_done}[slot][t & 1] completes its phase (t >> 1) at step t, so every wait is
// "parity (t >> 1) & 1" with no per-barrier bookkeeping.  K / V ring slots are released by two arrivals (one per issuer; an
// issuer whose slot has no tile in an item arrives directly).
// TMEM columns: S_A [0,128) = two 64-column buffers, S_B [128,256), O_A [256,320), O_B [320,384), remainder-token
// accumulators [384,464) as in attention2.cu.
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int A4_BQ = 128;
constexpr int A4_BKV = 128;           // keys per K / V shared-memory tile (one TMA box)
constexpr int A4_HB = 64;             // keys per softmax step (half a tile)
constexpr int A4_HD = 64;
constexpr int A4_THREADS = 352;       // 8 softmax warps + producer + 2 MMA issuers
constexpr int A4_TILE = 128 * 128;    // [128 rows x 64 bf16], SWIZZLE_128B
constexpr int A4_HALF = 64 * 128;     // byte offset of rows 64.. inside a tile
#ifndef OVK_A4_NS
#define OVK_A4_NS 3
#endif
constexpr int A4_NS = OVK_A4_NS;      // K / V ring depth (tiles); 5 (all of shared memory) measured equal to 3
constexpr int A4_OFF_Q = 0;                               // [2 item buffers][2 tiles]
constexpr int A4_OFF_V = A4_OFF_Q + 4 * A4_TILE;          // V before K (the O_t^T MMA over-reads one tile past V)
constexpr int A4_OFF_K = A4_OFF_V + A4_NS * A4_TILE;
constexpr int A4_OFF_ROWS = A4_OFF_K + A4_NS * A4_TILE;   // per item buffer: k_t, v_t, q_t rows (128 B each)
constexpr int A4_OFF_PT = A4_OFF_ROWS + 2 * 384;          // p_t: bf16 [128] (+128 B that the MMA over-reads)
constexpr int A4_OFF_RED = A4_OFF_PT + 384;               // remainder-row scratch: [2][4] max, [4] sum, s_tt
constexpr int A4_OFF_BAR = A4_OFF_RED + 64;
constexpr int A4_B_QFULL = 0;     // [2]
constexpr int A4_B_QFREE = 2;     // [2]
constexpr int A4_B_KFULL = 4;                       // [NS]
constexpr int A4_B_KEMPTY = A4_B_KFULL + A4_NS;     // [NS]  two arrivals
constexpr int A4_B_VFULL = A4_B_KEMPTY + A4_NS;     // [NS]
constexpr int A4_B_VEMPTY = A4_B_VFULL + A4_NS;     // [NS]  two arrivals
constexpr int A4_B_SFULL = A4_B_VEMPTY + A4_NS;     // [2 slots][2]
constexpr int A4_B_PREADY = A4_B_SFULL + 4;         // [2 slots][2]
constexpr int A4_B_PVDONE = A4_B_PREADY + 4;        // [2 slots][2]
constexpr int A4_NUM_BARS = A4_B_PVDONE + 4;
constexpr int A4_SMEM = A4_OFF_BAR + A4_NUM_BARS * 8 + 16;
static_assert(A4_SMEM <= 232448, "attention4: shared memory");
constexpr uint32_t A4_T_S = 0, A4_T_O = 256, A4_T_SK = 384, A4_T_ST = 416, A4_T_OT = 432, A4_T_TT = 448;

__device__ __forceinline__ void umma_bf16_ts4(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8_4(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float max3f_4(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// 16 exponentials p = 2^(s * scale - m) of one chunk, issued back to back
__device__ __forceinline__ void exp_chunk_4(const uint32_t (&x)[16], float scale_log2, float neg_m, float (&e)[16]) {
#ifdef OVK_A4_NO_F32X2
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = fast_exp2(fmaf(__uint_as_float(x[i]), scale_log2, neg_m));
#else
  const uint64_t s2 = f2_pack(scale_log2, scale_log2), m2 = f2_pack(neg_m, neg_m);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    f2_unpack(f2_fma(f2_pack(__uint_as_float(x[2 * i]), __uint_as_float(x[2 * i + 1])), s2, m2), lo, hi);
    e[2 * i] = fast_exp2(lo);
    e[2 * i + 1] = fast_exp2(hi);
  }
#endif
}
// TRUNC variant (OVK_ATT4_TRUNC=1, A/B): the fp32 -> bf16 pack of P is the XU pipe's second customer (F2FP runs there, ~3
// clocks per warp instruction next to the 8 of every MUFU).  A byte permute on the ALU pipe keeps the upper halves of two
// fp32 words instead (truncation towards zero).  Truncation alone would shrink P by E[ulp / 2] = 2^-8 * E[1 / mantissa] =
// 0.2818 % on average, so the exponent argument carries +log2(1 + delta), delta = 0.002826: E[trunc(e (1 + delta))] = e.
// The fp32 row sum then holds (1 + delta) * sum e; the epilogue divides it out (and the lse subtracts the constant).
constexpr float A4_TRUNC_LOG2 = 0.0040714f;      // log2(1.002826)
constexpr float A4_TRUNC_INV = 0.99718196f;      // 1 / 1.002826
// consume a chunk of exponentials: row-sum contribution and the 8 packed bf16 pairs
template <bool TRUNC>
__device__ __forceinline__ float pack_chunk_4(const float (&e)[16], uint32_t (&pw)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    pw[i] = TRUNC ? __byte_perm(__float_as_uint(e[2 * i]), __float_as_uint(e[2 * i + 1]), 0x7632) : pack_bf16x2(e[2 * i], e[2 * i + 1]);
#ifdef OVK_A4_NO_F32X2
  float s0 = e[0] + e[1], s1 = e[2] + e[3], s2 = e[4] + e[5], s3 = e[6] + e[7];
  float s4 = e[8] + e[9], s5 = e[10] + e[11], s6 = e[12] + e[13], s7 = e[14] + e[15];
  return ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
#else
  const uint64_t a0 = f2_add(f2_pack(e[0], e[1]), f2_pack(e[2], e[3])), a1 = f2_add(f2_pack(e[4], e[5]), f2_pack(e[6], e[7]));
  const uint64_t a2 = f2_add(f2_pack(e[8], e[9]), f2_pack(e[10], e[11])), a3 = f2_add(f2_pack(e[12], e[13]), f2_pack(e[14], e[15]));
  float lo, hi;
  f2_unpack(f2_add(f2_add(a0, a1), f2_add(a2, a3)), lo, hi);
  return lo + hi;
#endif
}

// what the deferred epilogue of an item needs
struct A4Item {
  float m_ref, l, s_tail, m_t, l_t;
  uint32_t u_tt, t_last;
  int b, h, q0, buf;
  bool titem;
};

// DBG (only with -DOVK_ATT4_DEBUG_VARIANTS, tools/attn_knockout.py): knock-outs that make the RESULT WRONG but show what a
// softmax step spends its time on: bit 0 no exponentials / packing, bit 1 no row maximum, bit 2 no P stores, bit 3 no score loads,
// bit 4 no epilogue work, bit 5 no S / P V MMAs, bit 6 no HBM traffic (every item reads and writes head 0 of image 0).
// MODE specialises the softmax step for the shapes that matter (the kernel is bound by the instruction stream of its softmax
// warps, so dead branches cost): 0 generic; 1 Lm a multiple of 64 and no remainder token (L = 256, 512, ...): no column masks,
// no remainder-token code at all; 2 Lm a multiple of 64 with the remainder token (L = 257: the ViT-L/14@224 tower).
template <bool TRUNC, int DBG = 0, int MODE = 0>
__global__ void __launch_bounds__(A4_THREADS, 1)
attention_fwd4_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmRow, float* __restrict__ lse_out,
                      __nv_bfloat16* __restrict__ out, int L, int Lm, int H, int nq, int total_items, float scale_log2,
                      int stagger_ns) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention4: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A4_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + A4_OFF_BAR + A4_NUM_BARS * 8);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int npair = (nq + 1) >> 1;
  const int nkv = (Lm + A4_BKV - 1) / A4_BKV;
  const int nhb = (Lm + A4_HB - 1) / A4_HB;
  const bool tail = MODE == 1 ? false : MODE == 2 ? true : L > Lm;   // one remainder token (host guarantees L - Lm <= 1, and then Lm % 128 == 0)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmRow);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[A4_B_QFULL + i], 1);
      mbar_init(&bars[A4_B_QFREE + i], 2);
      for (int e = 0; e < 2; ++e) {
        mbar_init(&bars[A4_B_SFULL + 2 * i + e], 1);
        mbar_init(&bars[A4_B_PREADY + 2 * i + e], 128);
        mbar_init(&bars[A4_B_PVDONE + 2 * i + e], 1);
      }
    }
    for (int i = 0; i < A4_NS; ++i) {
      mbar_init(&bars[A4_B_KFULL + i], 1);
      mbar_init(&bars[A4_B_KEMPTY + i], 2);
      mbar_init(&bars[A4_B_VFULL + i], 1);
      mbar_init(&bars[A4_B_VEMPTY + i], 2);
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
        const int pr = item % npair;
        // DBG bit 6: every item loads head 0 of image 0 (L2 hits): the kernel without its HBM reads
        const int h = (DBG & 64) ? 0 : (item / npair) % H, b = (DBG & 64) ? 0 : item / (npair * H);
        const int buf = n & 1;
        const bool has_b = 2 * pr + 1 < nq;
        const bool titem = tail && (2 * pr == nq - 1 || 2 * pr + 1 == nq - 1);
        mbar_wait(&bars[A4_B_QFREE + buf], ((n >> 1) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(&bars[A4_B_QFULL + buf], (has_b ? 2 : 1) * A4_TILE + (tail ? 256 : 0) + (titem ? 128 : 0));
        tma_load_4d(smem + A4_OFF_Q + (2 * buf) * A4_TILE, &tmQKV, &bars[A4_B_QFULL + buf], 0, h, 2 * pr * A4_BQ, b);
        if (has_b)
          tma_load_4d(smem + A4_OFF_Q + (2 * buf + 1) * A4_TILE, &tmQKV, &bars[A4_B_QFULL + buf], 0, h, (2 * pr + 1) * A4_BQ, b);
        if (tail) {
          tma_load_4d(smem + A4_OFF_ROWS + buf * 384, &tmRow, &bars[A4_B_QFULL + buf], 0, H + h, Lm, b);
          tma_load_4d(smem + A4_OFF_ROWS + buf * 384 + 128, &tmRow, &bars[A4_B_QFULL + buf], 0, 2 * H + h, Lm, b);
          if (titem) tma_load_4d(smem + A4_OFF_ROWS + buf * 384 + 256, &tmRow, &bars[A4_B_QFULL + buf], 0, h, Lm, b);
        }
        for (int j = 0; j < nkv; ++j, ++g) {
          const int s = g % A4_NS;
          const uint32_t ph = (g / A4_NS) & 1;
          mbar_wait(&bars[A4_B_KEMPTY + s], ph ^ 1, 41);
          mbar_arrive_expect_tx(&bars[A4_B_KFULL + s], A4_TILE);
          tma_load_4d(smem + A4_OFF_K + s * A4_TILE, &tmQKV, &bars[A4_B_KFULL + s], 0, H + h, j * A4_BKV, b);
          mbar_wait(&bars[A4_B_VEMPTY + s], ph ^ 1, 42);
          mbar_arrive_expect_tx(&bars[A4_B_VFULL + s], A4_TILE);
          tma_load_4d(smem + A4_OFF_V + s * A4_TILE, &tmQKV, &bars[A4_B_VFULL + s], 0, 2 * H + h, j * A4_BKV, b);
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer of tile slot w
      // Two cursors walk the (item, half block) steps of this slot: the S cursor two steps ahead of the P V cursor, across
      // item boundaries.  An item in which the slot has no tile (odd number of query tiles) ends the run: the S cursor
      // never looks past it (its K tiles cannot arrive before this warp has released ring slots that the lagging P V
      // cursor still holds), the warp performs its share of the item's K / V ring releases, and a new run starts.
      //
      // This thread sits on the critical path of every step (p_ready -> P V -> the S two steps ahead), and it is ONE thread:
      // every instruction costs a full dependent-issue latency.  The round-2 profile of the first version of this loop
      // showed ~280 instructions per step (integer divisions for the item decode, ring-slot modulos, descriptor encoding),
      // i.e. more time than the softmax group needs for the step.  So: no division or modulo per step (counters advance
      // incrementally), descriptors are a constant high word plus a running low word, and everything a P V needs is
      // computed BEFORE the wait for its p_ready.
      const int w = warp - 9;
      const int stride = static_cast<int>(gridDim.x);
      const int stride_pr = stride % npair;
      const uint32_t tS = tmem_base + A4_T_S + 128 * w;
      const uint32_t tO = tmem_base + A4_T_O + 64 * w;
      // descriptor = (hi word, constant) : (lo word = (addr >> 4) | LBO << 16); adding bytes >> 4 to lo moves the start address
      const uint64_t dk0 = umma_desc_kmajor_sw128(0);
      const uint64_t dmn0 = umma_desc_mnmajor_sw128(0, A4_TILE);
      const uint64_t dr0 = umma_desc_row0(0);
      auto with_addr = [](uint64_t d0, uint32_t addr) { return d0 + static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
      const uint64_t dq0 = with_addr(dk0, smem_u32(smem + A4_OFF_Q + w * A4_TILE));        // (selects, not arrays: no local memory)
      const uint64_t dq1 = with_addr(dk0, smem_u32(smem + A4_OFF_Q + (2 + w) * A4_TILE));
      const uint64_t dk_ring = with_addr(dk0, smem_u32(smem + A4_OFF_K));
      const uint64_t dv_ring = with_addr(dmn0, smem_u32(smem + A4_OFF_V));
      const uint64_t drow0 = with_addr(dr0, smem_u32(smem + A4_OFF_ROWS));
      const uint64_t drow1 = with_addr(dr0, smem_u32(smem + A4_OFF_ROWS + 384));
      const uint64_t dpt = with_addr(dr0, smem_u32(smem + A4_OFF_PT));
      constexpr uint32_t TILE16 = A4_TILE >> 4, HALF16 = A4_HALF >> 4;
      constexpr uint32_t idesc_s64 = umma_idesc_bf16(A4_BQ, A4_HB, 0, 0);
      constexpr uint32_t idesc_16 = umma_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(A4_BQ, A4_HD, 0, 1);
      constexpr uint32_t idesc_ot = umma_idesc_bf16(128, 16, 1, 0);
      const int valid_last = Lm - (nhb - 1) * A4_HB;            // keys of the last half block
      const int nblk_last = (valid_last + 15) & ~15;
      const uint32_t idesc_s_last = umma_idesc_bf16(A4_BQ, nblk_last, 0, 0);
      const int ksteps_last = nblk_last / 16;

      // item-level cursor state (advanced without division): item index, its pair index, CTA-local ordinal n, ring position
      struct ItemCur {
        int item, pr, n;
        int slot;        // ring slot of the item's first K / V tile
        uint32_t ph;     // phase bit of that slot's current use
      };
      auto next_item = [&](ItemCur& c) {
        c.item += stride;
        c.pr += stride_pr;
        if (c.pr >= npair) c.pr -= npair;
        ++c.n;
        for (int j = 0; j < nkv; ++j) {   // nkv ring steps (nkv <= 32; two tiles per item for the 257-token towers)
          if (++c.slot == A4_NS) {
            c.slot = 0;
            c.ph ^= 1;
          }
        }
      };
      auto is_active = [&](const ItemCur& c) { return c.item < total_items && 2 * c.pr + w < nq; };

      ItemCur it{static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.x) % npair, 0, 0, 0};
      uint32_t t_s = 0, t_p = 0;   // step counters of the two cursors (they are equal whenever a run starts)
      while (it.item < total_items) {
        if (!is_active(it)) {   // this warp's two-arrival share of the item's ring slots (it reads none of them)
          int sl = it.slot;
          uint32_t ph = it.ph;
          for (int j = 0; j < nkv; ++j) {
            mbar_wait(&bars[A4_B_KFULL + sl], ph, 52);
            mbar_arrive(&bars[A4_B_KEMPTY + sl]);
            mbar_wait(&bars[A4_B_VFULL + sl], ph, 53);
            mbar_arrive(&bars[A4_B_VEMPTY + sl]);
            if (++sl == A4_NS) {
              sl = 0;
              ph ^= 1;
            }
          }
          next_item(it);
          continue;
        }
        // ---- a run of consecutive active items
        ItemCur cs = it, cp = it;          // S cursor / P V cursor: item level
        int hb_s = 0, hb_p = 0;            // half block inside the item
        int ks = it.slot, vs = it.slot;    // ring slot of the K tile S(hb_s) reads / of the V tile P V(hb_p) reads
        uint32_t kph = it.ph, vph = it.ph;
        bool s_live = true;

        auto issue_s = [&]() {
          const int half = hb_s & 1;
          const int buf = cs.n & 1;
          const bool last = hb_s == nhb - 1;
          const bool carry = tail && (2 * cs.pr + w == nq - 1);   // this tile also carries the remainder query row
          const uint32_t d_tmem = tS + 64 * (t_s & 1);
          const uint64_t a = buf ? dq1 : dq0;
          const uint64_t rw = buf ? drow1 : drow0;
          const uint64_t kt = dk_ring + static_cast<uint32_t>(ks) * TILE16;
          const uint64_t bd = kt + static_cast<uint32_t>(half) * HALF16;
          const uint32_t idesc = last ? idesc_s_last : idesc_s64;
          if (hb_s == 0) mbar_wait(&bars[A4_B_QFULL + buf], (cs.n >> 1) & 1, 43);
          if (half == 0) mbar_wait(&bars[A4_B_KFULL + ks], kph, 44);
          tc_fence_after();
          if constexpr ((DBG & 32) == 0) {
#pragma unroll
            for (int k = 0; k < A4_HD / 16; ++k) umma_bf16_ss(d_tmem, a + 2 * k, bd + 2 * k, idesc, k != 0);
          }
          if (tail && hb_s == 0) {   // this tile's rows against the remainder key
#pragma unroll
            for (int k = 0; k < A4_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A4_T_SK + 16 * w, a + 2 * k, rw + 2 * k, idesc_16, k != 0);
          }
          if (carry && half == 0) {   // remainder query row against the 128 keys of this tile (transposed), and the remainder key
#pragma unroll
            for (int k = 0; k < A4_HD / 16; ++k)
              umma_bf16_ss(tmem_base + A4_T_ST, kt + 2 * k, rw + 16 + 2 * k, idesc_16, k != 0);
            if (hb_s == 0) {
#pragma unroll
              for (int k = 0; k < A4_HD / 16; ++k)
                umma_bf16_ss(tmem_base + A4_T_TT, rw + 2 * k, rw + 16 + 2 * k, idesc_16, k != 0);
            }
          }
          if (half == 1 || last) umma_commit(&bars[A4_B_KEMPTY + ks]);   // this warp's last read of the K tile
          umma_commit(&bars[A4_B_SFULL + 2 * w + (t_s & 1)]);
          ++t_s;
          if (half == 1 || last) {   // next K tile
            if (++ks == A4_NS) {
              ks = 0;
              kph ^= 1;
            }
          }
          if (++hb_s == nhb) {
            hb_s = 0;
            next_item(cs);
            s_live = is_active(cs);
          }
        };

        issue_s();
        if (s_live) issue_s();
        while (true) {
          // ---- P V(t_p): operands first, then the wait, then nothing but the MMAs
          const int half = hb_p & 1;
          const bool last = hb_p == nhb - 1;
          const bool carry = tail && (2 * cp.pr + w == nq - 1);
          const uint32_t e = t_p & 1;
          const uint32_t a_tmem = tS + 64 * e;
          const uint64_t vt = dv_ring + static_cast<uint32_t>(vs) * TILE16;
          const uint64_t bd = vt + static_cast<uint32_t>(half) * HALF16;
          const int ksteps = last ? ksteps_last : A4_HB / 16;
          if (half == 0) mbar_wait(&bars[A4_B_VFULL + vs], vph, 46);
          mbar_wait(&bars[A4_B_PREADY + 2 * w + e], (t_p >> 1) & 1, 45);   // P(t) is in TMEM (and the previous O has been read)
          tc_fence_after();
          if constexpr ((DBG & 32) != 0) {
          } else if (ksteps == 4) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_ts4(tO, a_tmem + 8 * kk, bd + kk * 128, idesc_pv, (hb_p | kk) != 0);
          } else {
            for (int kk = 0; kk < ksteps; ++kk) umma_bf16_ts4(tO, a_tmem + 8 * kk, bd + kk * 128, idesc_pv, (hb_p | kk) != 0);
          }
          if (carry && half == 0) {   // O_t^T += V_j^T p_t^T over the tile's 128 keys (M runs over head dims, rows 64..127 padding)
#pragma unroll
            for (int kk = 0; kk < A4_BKV / 16; ++kk)
              umma_bf16_ss(tmem_base + A4_T_OT, vt + kk * 128, dpt + 2 * kk, idesc_ot, (hb_p | kk) != 0);
          }
          if (half == 1 || last) umma_commit(&bars[A4_B_VEMPTY + vs]);   // this warp's last read of the V tile
          umma_commit(&bars[A4_B_PVDONE + 2 * w + e]);
          ++t_p;
          if (s_live) issue_s();   // refills the buffer the P V just queued reads
          if (half == 1 || last) {
            if (++vs == A4_NS) {
              vs = 0;
              vph ^= 1;
            }
          }
          if (++hb_p == nhb) {
            hb_p = 0;
            next_item(cp);
            if (!is_active(cp)) break;
          }
        }
        it = cp;
      }
    }
  } else if (warp < 8) {
    // -------------------------------------------------------------------- softmax groups
    const int w = warp >> 2;                 // tile slot (0 = A, 1 = B)
    const int quad = warp & 3;
    const int r = quad * 32 + static_cast<int>(lane);   // query row in the tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_sw = tmem_base + t_lane + A4_T_S + 128 * w;
    const uint32_t t_o = tmem_base + t_lane + A4_T_O + 64 * w;
    const uint32_t red = smem_u32(smem + A4_OFF_RED);
    const uint32_t bar_id = 1 + w;
    const float neg_inf = -INFINITY;
    uint32_t t = 0;            // step counter of this slot
    bool pending = false;      // an item whose epilogue has not run yet
    A4Item prev{};

    // ---- epilogue of `prev`: O / l -> bf16 -> smem -> TMA store (+ the remainder query row of the head)
    // (Measured and dropped, tools/attn_ab.py against this version on the same box: handing the staging Q buffer back one
    // step later instead of waiting for the store's read right here: +2..5 %; pulling O out of TMEM before the p_ready
    // arrival and normalising / storing after it: +13 % (o[64] then lives across the arrival and spills at the 168-register
    // cap); polling instead of suspending barrier waits: +9 %; a 5-deep K / V ring: +-0; every thread storing its 128-byte
    // output row straight from registers (no staging tile / proxy fence / group barrier / TMA store): +-0.  tools/attn_knockout.py
    // shows why none of this moves the total: the exponentials run at the XU pipe's rate but nothing else overlaps them - the two
    // tile slots execute the same step at the same time, so both warps of a scheduler are in their exp phase together and in
    // their load / max / hand-off phase together.)
    auto epilogue = [&]() {
      const uint32_t tl = prev.t_last;
      mbar_wait(&bars[A4_B_PVDONE + 2 * w + (tl & 1)], (tl >> 1) & 1, 48);
      tc_fence_after();
      if constexpr ((DBG & 16) != 0) {   // barrier protocol only
        if (r == 0) mbar_arrive(&bars[A4_B_QFREE + prev.buf]);
        return;
      }
      uint32_t o[64];
      {
        uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
        uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
        tmem_ld_x32(t_o, o0);
        tmem_ld_x32(t_o + 32, o1);
      }
      uint32_t u_ot = 0;
      if (prev.titem && quad < 2) tmem_ld_x1(tmem_base + t_lane + A4_T_OT, u_ot);
      tmem_ld_wait();
      float m_ref = prev.m_ref, l = TRUNC ? prev.l * A4_TRUNC_INV : prev.l;
      const int b = prev.b, h = prev.h, q0 = prev.q0, buf = prev.buf;
      const bool titem = prev.titem;
      const uint32_t rows = smem_u32(smem + A4_OFF_ROWS + buf * 384);
      if (tail) {   // fold the remainder key in: one more online-softmax step, entirely in registers
        const float m_fin = fmaxf(m_ref, prev.s_tail);
        const float a = fast_exp2(m_ref - m_fin);
        const float pt = fast_exp2(prev.s_tail - m_fin);
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
      const uint32_t stage = smem_u32(smem + A4_OFF_Q + (2 * buf + w) * A4_TILE);   // this tile's Q buffer is dead by now
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(stage + sw128_offset(r, c),
               make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l),
                          pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l)));
      if (lse_out != nullptr && q0 + r < L)
        lse_out[(static_cast<long long>(b) * H + h) * L + q0 + r] = (m_ref + log2f(l)) * 0.69314718055994531f;
      float l_t = prev.l_t;
      if (titem) {   // finish the remainder query row: totals over the group, its own remainder key, normalise, store
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) l_t += __shfl_xor_sync(0xffffffffu, l_t, o2);
        if (lane == 0) sts_f32(red + 32 + quad * 4, l_t);
        if (r == 0) sts_f32(red + 48, __uint_as_float(prev.u_tt) * scale_log2);
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (titem && quad < 2) {
        const float4 w4 = lds_f32x4(red + 32);
        const float s_tt = lds_f32(red + 48);
        const float m_fin = fmaxf(prev.m_t, s_tt);
        const float a = fast_exp2(prev.m_t - m_fin);
        const float pt = fast_exp2(s_tt - m_fin);
        const float l_all = fmaf((w4.x + w4.y) + (w4.z + w4.w), a, pt);
        unsigned short vb;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(vb) : "r"(rows + 128 + 2 * r));
        const float v = __uint_as_float(static_cast<uint32_t>(vb) << 16);
        const float ov = fmaf(__uint_as_float(u_ot), a, pt * v) / l_all;
        const float o_hi = __shfl_down_sync(0xffffffffu, ov, 1);
        const long long row = static_cast<long long>(b) * L + Lm;
        if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(out + (row * H + h) * A4_HD + r) = pack_bf16x2(ov, o_hi);
        if (r == 0 && lse_out != nullptr)
          lse_out[(static_cast<long long>(b) * H + h) * L + Lm] = (m_fin + log2f(l_all)) * 0.69314718055994531f;
      }
      if (r == 0) {
        tma_store_4d(&tmO, smem + A4_OFF_Q + (2 * buf + w) * A4_TILE, 0, h, q0, b);
        tma_store_commit();
        tma_store_wait_read<0>();   // the producer may refill this Q buffer
        mbar_arrive(&bars[A4_B_QFREE + buf]);
      }
      // (the scratch words red+32.. / the staging tile are reused only after the next named barrier / the next TMA load)
    };

    // Start-up stagger of slot B (OVK_ATT4_STAGGER_NS, A/B): the two slots do the same work per step, so left alone both warps
    // of a scheduler sit in their exp phase (XU contended) and in their load / max / hand-off phase (XU idle) together.
    if (w == 1 && stagger_ns > 0) __nanosleep(static_cast<unsigned>(stagger_ns));
    for (int n = 0;; ++n) {
      const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
      const bool done = item >= total_items;
      const int pr = done ? 0 : item % npair;
      const int qt = 2 * pr + w;
      if (done || qt >= nq) {
        if (pending) {   // nothing to overlap the epilogue with
          epilogue();
          pending = false;
        }
        if (done) break;
        if (r == 0) {    // odd number of query tiles: this group sits the item out (but stays in step with the producer)
          mbar_wait(&bars[A4_B_QFULL + (n & 1)], (n >> 1) & 1, 49);
          mbar_arrive(&bars[A4_B_QFREE + (n & 1)]);
        }
        continue;
      }
      A4Item cur;
      cur.h = (DBG & 64) ? 0 : (item / npair) % H;     // (bit 6: stores go to one place as well)
      cur.b = (DBG & 64) ? 0 : item / (npair * H);
      cur.buf = n & 1;
      cur.q0 = qt * A4_BQ;
      cur.titem = tail && qt == nq - 1;   // this group also carries the remainder query row
      cur.m_ref = neg_inf;
      cur.l = 0.f;
      cur.s_tail = 0.f;
      cur.m_t = neg_inf;
      cur.l_t = 0.f;
      cur.u_tt = 0;   // q_t . k_t (thread 0 of the group)
      const bool titem = cur.titem;
#pragma unroll 1
      for (int hb = 0; hb < nhb; ++hb, ++t) {
        const int half = hb & 1;
        const uint32_t e = t & 1;
        const int valid = MODE != 0 ? A4_HB : min(A4_HB, Lm - hb * A4_HB);
        const uint32_t t_s = t_sw + 64 * e;
        mbar_wait(&bars[A4_B_SFULL + 2 * w + e], (t >> 1) & 1, 47);
        tc_fence_after();
        uint32_t u_sk = 0, u_st = 0;
        if (titem && hb == 0 && quad == 0) tmem_ld_x1(tmem_base + t_lane + A4_T_TT, cur.u_tt);
        if (tail && hb == 0) tmem_ld_x1(tmem_base + t_lane + A4_T_SK + 16 * w, u_sk);
        if (titem && half == 0) tmem_ld_x1(tmem_base + t_lane + A4_T_ST, u_st);
        // The 64 scores of the step are read from TMEM ONCE and stay in registers for both the row maximum and the
        // exponentials: TMEM reads run at ~64 B/clk per SM, so a second pass over S costs as much as all the exponentials
        // of the step (attention2 / attention3 read S twice and were bound by exactly that).
        uint32_t x[64];
        {
          uint32_t(&x0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&x[0]);
          uint32_t(&x1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&x[32]);
          if constexpr ((DBG & 8) == 0) {
            tmem_ld_x32(t_s, x0);
            if (valid > 32) tmem_ld_x32(t_s + 32, x1);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              x[i] = 0x3c000000u + lane + i;
              asm volatile("" : "+r"(x[i]));
            }
          }
        }
        float mx = neg_inf;
        if constexpr ((DBG & 2) != 0) {
          mx = __uint_as_float(x[0]);
        } else if (valid == A4_HB) {
          float m4[4] = {neg_inf, neg_inf, neg_inf, neg_inf};
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = max3f_4(m4[i & 3], __uint_as_float(x[2 * i]), __uint_as_float(x[2 * i + 1]));
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            x[i] = (i < valid) ? x[i] : 0xff800000u;   // -inf: contributes 0 to P and to the row sum
            mx = fmaxf(mx, __uint_as_float(x[i]));
          }
        }
        if (tail && hb == 0) cur.s_tail = __uint_as_float(u_sk) * scale_log2;
        const float m_blk = mx * scale_log2;
        // lazy rescaling: the reference only moves when the maximum grew by more than 2^8.  O may only be touched once
        // P V(t-1) has landed (S(t) was issued behind P V(t-2) only)
        if (hb == 0) {
          cur.m_ref = m_blk;
        } else {
          const bool grow = m_blk > cur.m_ref + 8.f;
          if (__any_sync(0xffffffffu, grow)) {
            mbar_wait(&bars[A4_B_PVDONE + 2 * w + (e ^ 1)], ((t - 1) >> 1) & 1, 50);
            tc_fence_after();
            const float alpha = grow ? fast_exp2(cur.m_ref - m_blk) : 1.f;
            cur.m_ref = grow ? m_blk : cur.m_ref;
            cur.l *= alpha;
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
            cur.m_t = mt;
          } else if (mt > cur.m_t + 8.f) {   // uniform over the group
            const float alpha = fast_exp2(cur.m_t - mt);
            cur.m_t = mt;
            cur.l_t *= alpha;
            if (quad < 2) {              // O_t^T: lanes 0..63 = head dims, column 0
              uint32_t o1;
              tmem_ld_x1(tmem_base + t_lane + A4_T_OT, o1);
              tmem_ld_wait();
              tmem_st_x1(tmem_base + t_lane + A4_T_OT, __float_as_uint(__uint_as_float(o1) * alpha));
            }
          }
          const float p = fast_exp2(st - cur.m_t);
          cur.l_t += p;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(smem + A4_OFF_PT) + 2 * r),
                       "h"(static_cast<unsigned short>(pack_bf16x2(p, 0.f) & 0xFFFFu)) : "memory");
        }
        // P = 2^(S*scale - m_ref) as packed bf16 pairs, written over the score columns: chunk c (score columns [16c, 16c+16))
        // becomes words [8c, 8c+8) of the buffer.  Hand-pipelined: the exponentials of chunk c+1 are issued before the sums /
        // packing of chunk c, so nothing waits on a MUFU result issued just before it.  (Masked columns hold -inf -> 0.)
        const float neg_m = TRUNC ? A4_TRUNC_LOG2 - cur.m_ref : -cur.m_ref;
        {
          uint32_t pw[8];
          float ea[16], eb[16];
          uint32_t(&c0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&x[0]);
          uint32_t(&c1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&x[16]);
          uint32_t(&c2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&x[32]);
          uint32_t(&c3)[16] = *reinterpret_cast<uint32_t(*)[16]>(&x[48]);
          if constexpr ((DBG & 1) != 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pw[i] = x[16 * c + i] ^ x[16 * c + 8 + i];
              if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s + 8 * c, pw);
              else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
            }
            cur.l += neg_m;
          } else {
          exp_chunk_4(c0, scale_log2, neg_m, ea);
          exp_chunk_4(c1, scale_log2, neg_m, eb);
          float acc = pack_chunk_4<TRUNC>(ea, pw);
          if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s, pw); else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
          if (valid > 32) {   // uniform
            exp_chunk_4(c2, scale_log2, neg_m, ea);
            acc += pack_chunk_4<TRUNC>(eb, pw);
            if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s + 8, pw); else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
            exp_chunk_4(c3, scale_log2, neg_m, eb);
            acc += pack_chunk_4<TRUNC>(ea, pw);
            if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s + 16, pw); else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
            acc += pack_chunk_4<TRUNC>(eb, pw);
            if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s + 24, pw); else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
          } else {
            acc += pack_chunk_4<TRUNC>(eb, pw);
            if constexpr ((DBG & 4) == 0) tmem_st_x8_4(t_s + 8, pw); else asm volatile("" ::"r"(pw[0]), "r"(pw[7]));
          }
          cur.l += acc;
          }
        }
        tmem_st_wait();
        if (titem && half == 0) fence_proxy_async_smem();
        // the previous item's epilogue runs here: its last P V has had this whole step to land, and P V(t), which overwrites
        // O, is not released before O has been read
        if (hb == 0 && pending) {
          epilogue();
          pending = false;
        }
        tc_fence_before();
        mbar_arrive(&bars[A4_B_PREADY + 2 * w + e]);
      }
      cur.t_last = t - 1;
      prev = cur;
      pending = true;
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
int ovk_attention_fwd4_launch(const void* qkv, void* out, float* lse, int B, int L, int H, float scale, cudaStream_t s) {
  const int hd = A4_HD;
  CUtensorMap tmQKV, tmO, tmRow;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {A4_HD, 1, A4_BKV, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const uint32_t box1[4] = {A4_HD, 1, 1, 1};   // one head row (128 B), unswizzled: remainder key / value / query
    if ((rc = make_tmap_nd_bf16(&tmRow, qkv, 4, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {A4_HD, 1, A4_BQ, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaSuccess;
    auto set_attr = [&](auto kern) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM);
    };
    set_attr(attention_fwd4_kernel<false, 0, 0>);
    set_attr(attention_fwd4_kernel<false, 0, 1>);
    set_attr(attention_fwd4_kernel<false, 0, 2>);
    set_attr(attention_fwd4_kernel<true, 0, 0>);
    set_attr(attention_fwd4_kernel<true, 0, 1>);
    set_attr(attention_fwd4_kernel<true, 0, 2>);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention4): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  const char* stg = getenv("OVK_ATT4_STAGGER_NS");
  const int stagger = stg != nullptr ? atoi(stg) : 0;
  const char* tr = getenv("OVK_ATT4_TRUNC");   // read per call (A/B runs alternate the two inside one process)
  const bool trunc = tr != nullptr && tr[0] == '1';
  const int tail = (L > A4_BQ && L % A4_BQ == 1) ? 1 : 0;   // cls + power-of-two grid: remainder token handled outside the tiles
  const int l_main = L - tail;
  const int nq = (l_main + A4_BQ - 1) / A4_BQ;
  const long long items = static_cast<long long>((nq + 1) / 2) * H * B;
  if (items > 0x7fffffffLL) return set_error(OVK_ERR_SHAPE, "attention: too many work items");
  const int grid = static_cast<int>(items < (long long)num_sms() ? items : (long long)num_sms());
#ifdef OVK_ATT4_DEBUG_VARIANTS
  {
    const char* dbg = getenv("OVK_ATT4_DBG");
    const int d = dbg != nullptr ? atoi(dbg) : 0;
#define OVK_A4_DBG_CASE(N)                                                                                                   \
    if (d == N) {                                                                                                              \
      cudaFuncSetAttribute(attention_fwd4_kernel<false, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM);             \
      attention_fwd4_kernel<false, N><<<grid, A4_THREADS, A4_SMEM, s>>>(tmQKV, tmO, tmRow, lse, reinterpret_cast<__nv_bfloat16*>(out), \
                                                                       L, l_main, H, nq, static_cast<int>(items),              \
                                                                       scale * 1.4426950408889634f, stagger);                 \
      return check_launch("attention_fwd4_kernel<dbg>");                                                                       \
    }
    OVK_A4_DBG_CASE(1) OVK_A4_DBG_CASE(2) OVK_A4_DBG_CASE(4) OVK_A4_DBG_CASE(8) OVK_A4_DBG_CASE(5) OVK_A4_DBG_CASE(7) OVK_A4_DBG_CASE(15)
    OVK_A4_DBG_CASE(16) OVK_A4_DBG_CASE(31) OVK_A4_DBG_CASE(47) OVK_A4_DBG_CASE(63) OVK_A4_DBG_CASE(64) OVK_A4_DBG_CASE(79) OVK_A4_DBG_CASE(127)
#undef OVK_A4_DBG_CASE
  }
#endif
  const char* gen = getenv("OVK_ATT4_GENERIC");   // =1: the unspecialised instance for every shape (A/B, tests)
  const int mode = (gen != nullptr && gen[0] == '1') || (l_main % A4_HB) != 0 ? 0 : (tail ? 2 : 1);
#define OVK_A4_LAUNCH(TR, MD)                                                                                                 \
  attention_fwd4_kernel<TR, 0, MD><<<grid, A4_THREADS, A4_SMEM, s>>>(tmQKV, tmO, tmRow, lse, reinterpret_cast<__nv_bfloat16*>(out), \
                                                                    L, l_main, H, nq, static_cast<int>(items),                 \
                                                                    scale * 1.4426950408889634f, stagger)
  if (trunc) {
    if (mode == 1) OVK_A4_LAUNCH(true, 1); else if (mode == 2) OVK_A4_LAUNCH(true, 2); else OVK_A4_LAUNCH(true, 0);
  } else {
    if (mode == 1) OVK_A4_LAUNCH(false, 1); else if (mode == 2) OVK_A4_LAUNCH(false, 2); else OVK_A4_LAUNCH(false, 0);
  }
#undef OVK_A4_LAUNCH
  return check_launch("attention_fwd4_kernel");
}
