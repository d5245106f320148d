// One-pass attention backward, TRANSPOSED score tiles (autograd of nn.MultiheadAttention's softmax(QK^T/sqrt(hd))V,
// open_clip/transformer.py:225,239-252).  Successor of attention_bwd.cu's MODE_FUSED kernel, whose softmax warps and tensor
// pipe took turns (the score and dP accumulators are single-buffered there: 512 TMEM columns do not hold two sets).
//
// Work item = (128-key tile j, head, image); key / value tiles stationary, query / dO tiles streamed.  Scores are computed
// TRANSPOSED, keys on the TMEM lanes, in two 64-query halves per streamed tile:
//     S^T_h = K_j Q_h^T            dP^T_h = V_j dO_h^T            (M = 128 keys, N = 64 queries, SS form)
//     P^T_h = exp2(S^T_h * scale*log2e - lse2[q]),  dS^T_h = P^T_h . (dP^T_h - delta[q])      (one thread per key row and
//                                                                   32 query columns; lse2 / delta broadcast from smem)
//     dV += P^T_h dO_h             dK += dS^T_h Q_h               (A = the bf16 tile written back INTO the S^T / dP^T columns:
//                                                                  TS form, no shared-memory round trip)
//     dQ_i(partial) = dS_i K_j                                    (A = dS^T staged in smem as [key rows][query cols], read as an
//                                                                  MN-major operand; drained through TMA reduce-add, fp32)
// The halves are the pipeline: while the compute warps work on half b of tile i the tensor pipe runs dV / dK of half a and the
// scores of half a of tile i+1 (256 TMEM columns hold both halves of S^T and dP^T, as before), so neither side waits for the
// other in steady state.  Per 128 x 128 score tile: 5 MMAs, 16 384 exponentials, 160 KB of TMEM reads.
// TMEM: S^T a|b 0..127, dP^T a|b 128..255, dK 256, dV 320, dQ 384, narrow (head dims 64..79) dK 448, dV 464, dQ 480.
#include <algorithm>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int T2_T = 128;
constexpr int T2_HD = 64;
constexpr int T2_TILE = T2_T * 128;            // 16 KB: [128 rows x 64 bf16]
// compute warps: NW per TMEM lane quarter (2 or 4): a warp owns 32 key rows and CG = 64 / NW query columns of a half
constexpr int t2_threads(int nw) { return 32 * (4 * nw + 2); }   // + TMA producer warp + MMA warp
constexpr int T2_OFF_K0 = 0;
constexpr int T2_OFF_V0 = 1 * T2_TILE;
constexpr int T2_OFF_Q = 2 * T2_TILE;          // x2 (streamed)
constexpr int T2_OFF_DO = 4 * T2_TILE;         // x2
constexpr int T2_OFF_DS = 6 * T2_TILE;         // dS^T bf16 [128 key rows][128 queries] as two 64-query atoms, x2 buffers
constexpr int T2_OFF_DQST = 10 * T2_TILE;      // fp32 staging of the dQ partial: one [32 rows x 128 B] region per compute warp
constexpr int T2_OFF_K1 = 12 * T2_TILE;        // second stationary set (hd = 64): the next item's K / V arrive under this one
constexpr int T2_OFF_V1 = 13 * T2_TILE;
// wide heads (64 < hd <= 80): narrow [128 rows x 32 B] SWIZZLE_32B tiles take the place of the second stationary set
constexpr int T2_BT = 128 * 32;
constexpr int T2_OFF_KB = 12 * T2_TILE;
constexpr int T2_OFF_VB = T2_OFF_KB + T2_BT;
constexpr int T2_OFF_QB = T2_OFF_VB + T2_BT;    // x2
constexpr int T2_OFF_DOB = T2_OFF_QB + 2 * T2_BT;  // x2
constexpr int T2_OFF_NST = T2_OFF_DOB + 2 * T2_BT; // x2: narrow dQ staging (dense 64-byte rows) / narrow dK, dV output tiles
constexpr int T2_OFF_STATS = 14 * T2_TILE;     // [2 slots][-lse*log2e | -delta][128] floats
constexpr int T2_OFF_BAR = T2_OFF_STATS + 2 * 2 * 128 * 4;
constexpr int T2_NUM_BARS = 15;
constexpr int T2_SMEM_BYTES = T2_OFF_BAR + T2_NUM_BARS * 8 + 16;
static_assert(T2_OFF_NST + 2 * T2_BT <= T2_OFF_STATS, "narrow tiles overflow the second stationary set");
static_assert(T2_SMEM_BYTES <= 232448, "shared memory budget");
constexpr uint32_t T2_TM_S = 0, T2_TM_DP = 128, T2_TM_DK = 256, T2_TM_DV = 320, T2_TM_DQ = 384;
constexpr uint32_t T2_TM_DKB = 448, T2_TM_DVB = 464, T2_TM_DQB = 480;
constexpr uint32_t T2_TM_K = 448, T2_TM_V = 480;   // hd = 64: K_j / V_j as TMEM-resident A operands (32 columns each)

__device__ __forceinline__ uint64_t t2_desc_sw32(uint32_t saddr) { return umma_desc(saddr, 16, 256, 6); }
__device__ __forceinline__ uint32_t t2_sw32_offset(uint32_t row, uint32_t chunk) {
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A is the [128 x 16] bf16 block held as 8 TMEM columns of packed pairs
__device__ __forceinline__ void t2_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared memory -> TMEM copy of a [128 rows x 256 bit] slab described like an MMA operand (one K = 16 step of a K-major bf16 tile
// becomes the 8 TMEM columns a TS-form MMA reads as its A block); ordered with the MMAs of the issuing thread
__device__ __forceinline__ void t2_tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// TMEM column (inside a half's 64 columns) of the bf16 A block of 16-query group kk: the two 32-column thread groups of a
// half each write their 16 packed columns at the START of their own 32 fp32 columns (nobody writes where another warp reads)
template <int CG>
__device__ __forceinline__ uint32_t t2_acol(int kk) {
  return CG == 32 ? static_cast<uint32_t>((kk >> 1) * 32 + (kk & 1) * 8) : static_cast<uint32_t>(kk * 16);
}
__device__ __forceinline__ void t2_tmem_ld(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_x32(taddr, r); }
__device__ __forceinline__ void t2_tmem_ld(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_x16(taddr, r); }
__device__ __forceinline__ void t2_tmem_st(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st_x16(taddr, r); }
__device__ __forceinline__ void t2_tmem_st(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16-byte chunk `chunk` of row `row` in a [rows x 64 B] SWIZZLE_64B tile (base 512-byte aligned)
__device__ __forceinline__ uint32_t t2_sw64_offset(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

__device__ __forceinline__ void t2_wait(uint64_t* bar, uint32_t parity, int code, bool spin) {
  if (spin) {
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
      if (++polls > (1u << 28)) hang_trap(code);
    }
  } else {
    mbar_wait(bar, parity, code);
  }
}

// debug (dbg bit 32): cycles spent in a wait, accumulated per wait site
__device__ __forceinline__ void t2_wait_t(uint64_t* bar, uint32_t parity, int code, bool spin, bool prof, long long& acc) {
  if (prof) {
    const long long t0 = clock64();
    t2_wait(bar, parity, code, spin);
    acc += clock64() - t0;
  } else {
    t2_wait(bar, parity, code, spin);
  }
}

template <int RB, int NW>
__global__ void __launch_bounds__(t2_threads(NW), 1)
attention_bwd_t_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmDQKV, const __grid_constant__ CUtensorMap tmQKVb,
                       const __grid_constant__ CUtensorMap tmDOb, const __grid_constant__ CUtensorMap tmDQKVb,
                       const __grid_constant__ CUtensorMap tmDQF, const __grid_constant__ CUtensorMap tmDQFb,
                       const float* __restrict__ stats_g, int Lp, int L, int H, float scale,
                       int total_items, int per, int causal, int Lm, const float* __restrict__ ws,
                       const __nv_bfloat16* __restrict__ qkv_g, const __nv_bfloat16* __restrict__ dout_g, int hd_g) {
  // Lm < L (= L - 1): the tiles cover tokens [0, Lm); the remainder token is attention_bwd_tail_kernel's (attention_bwd.cu),
  // which leaves ws[1][j] = P(t, j), ws[2][j] = dS(t, j) for the rank-1 terms of the dV / dK epilogue below.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention_bwd_t: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  constexpr int CW = 4 * NW;        // compute warps
  constexpr int CG = 64 / NW;       // query columns of a half per warp
  constexpr int NST = RB ? 1 : 2;   // stationary tile sets
  constexpr bool KVT = (RB == 0);   // K_j / V_j also live in TMEM (the 64 columns the narrow accumulators would take)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_OFF_BAR);
  uint64_t* st_full = bars + 0;    // [2] K_j / V_j of an item have landed
  uint64_t* st_empty = bars + 2;   // [2] ... and no MMA reads them any more
  uint64_t* sf = bars + 4;         // [2] streamed Q_i / dO_i tile + its statistics have landed
  uint64_t* se = bars + 6;         // [2] ... and are no longer read
  uint64_t* s_full = bars + 8;     // [2 halves] S^T_h and dP^T_h are in TMEM
  uint64_t* p_ready = bars + 10;   // [2 halves] the compute warps wrote P^T_h / dS^T_h (TMEM) and dS^T_h (smem)
  uint64_t* dq_full = bars + 12;   // the dQ partial of a streamed tile is in TMEM (and every earlier MMA has completed)
  uint64_t* dq_free = bars + 13;   // ... and has been read out
  uint64_t* acc_free = bars + 14;  // the epilogue has read dK / dV of the previous item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_OFF_BAR + T2_NUM_BARS * 8);
  float* stats = reinterpret_cast<float*>(smem + T2_OFF_STATS);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int nt = (Lm + T2_T - 1) / T2_T;
  const float s2 = scale * 1.4426950408889634f;
  const int item_first = blockIdx.x * per;
  const int item_last = min(total_items, item_first + per);
#ifdef OVK_ATTBWD_DEBUG
  // build with NVCCFLAGS_EXTRA=-DOVK_ATTBWD_DEBUG for tools/attn_bwd_knockout.py: knock-out bits (results wrong on purpose),
  // bit 16: spinning waits, bit 32: cycles per wait / issue site printed by block 0
  const int dbg = causal >> 8;
  const bool spin = (dbg & 16) != 0;
  const bool prof = (dbg & 32) != 0;
#else
  constexpr int dbg = 0;
  constexpr bool spin = false, prof = false;
#endif
  causal &= 1;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0, w7 = 0, w8 = 0;
  const long long t_start = clock64();
  auto off_k = [](int buf) { return buf == 0 ? T2_OFF_K0 : T2_OFF_K1; };
  auto off_v = [](int buf) { return buf == 0 ? T2_OFF_V0 : T2_OFF_V1; };

  if (warp == CW && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
    tma_prefetch_desc(&tmDQF);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
      mbar_init(&sf[i], 1);
      mbar_init(&se[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], CW);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_free, CW);
    mbar_init(acc_free, CW);
    fence_mbar_init();
  }
  if (warp == CW + 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The CTA's work as ONE sequence of score tiles G = (item n, streamed query tile it): the roles below never drain their
  // pipelines at an item boundary (the next item's scores are issued behind the last tile's dV / dK MMAs, and the compute
  // warps read out dK / dV of item n between the two halves of the first tile of item n+1).
  const int n_items = max(0, item_last - item_first);
  const int total_tiles = n_items * nt;

  if (warp == CW) {
    // ---------------------------------------------------------------------- producer (one thread): TMA tiles + statistics
    if (lane == 0) {
    int n = 0, it = 0, t0 = 0, h = 0, b = 0;
    long long bh = 0;
    for (int G = 0; G < total_tiles; ++G) {
      if (it == 0) {
        const int item = item_first + n;
        t0 = (item % nt) * T2_T, h = (item / nt) % H, b = item / (nt * H);
        bh = static_cast<long long>(b) * H + h;
        const int buf = n % NST;
        t2_wait_t(&st_empty[buf], ((n / NST) & 1) ^ 1, 40, spin, prof, w0);
        mbar_arrive_expect_tx(&st_full[buf], 2 * T2_TILE + (RB ? 2 * T2_BT : 0));
        tma_load_4d(smem + off_k(buf), &tmQKV, &st_full[buf], 0, H + h, t0, b);       // K_j
        tma_load_4d(smem + off_v(buf), &tmQKV, &st_full[buf], 0, 2 * H + h, t0, b);   // V_j
        if (RB) {
          tma_load_4d(smem + T2_OFF_KB, &tmQKVb, &st_full[buf], 64, H + h, t0, b);
          tma_load_4d(smem + T2_OFF_VB, &tmQKVb, &st_full[buf], 64, 2 * H + h, t0, b);
        }
        if (NST == 1 && n + 1 < n_items) {
          // one stationary set: the next item's K / V can only be fetched when this item's last MMA has completed, and the
          // scores of its first tile wait for them.  Pulling the tiles into L2 now takes the DRAM part out of that wait.
          const int item1 = item + 1;
          const int t1 = (item1 % nt) * T2_T, h1 = (item1 / nt) % H, b1 = item1 / (nt * H);
          tma_prefetch_l2_4d(&tmQKV, 0, H + h1, t1, b1);
          tma_prefetch_l2_4d(&tmQKV, 0, 2 * H + h1, t1, b1);
          tma_prefetch_l2_4d(&tmQKVb, 64, H + h1, t1, b1);
          tma_prefetch_l2_4d(&tmQKVb, 64, 2 * H + h1, t1, b1);
        }
      }
      const int s = G & 1;
      t2_wait_t(&se[s], ((G >> 1) & 1) ^ 1, 41, spin, prof, w1);
      mbar_arrive_expect_tx(&sf[s], 2 * T2_TILE + (RB ? 2 * T2_BT : 0) + 2 * 512);
      tma_load_4d(smem + T2_OFF_Q + s * T2_TILE, &tmQKV, &sf[s], 0, h, it * T2_T, b);    // Q_i
      tma_load_4d(smem + T2_OFF_DO + s * T2_TILE, &tmDO, &sf[s], 0, h, it * T2_T, b);    // dO_i
      if (RB) {
        tma_load_4d(smem + T2_OFF_QB + s * T2_BT, &tmQKVb, &sf[s], 64, h, it * T2_T, b);
        tma_load_4d(smem + T2_OFF_DOB + s * T2_BT, &tmDOb, &sf[s], 64, h, it * T2_T, b);
      }
      // per-query statistics of the tile as attention_bwd_delta_kernel left them: nl = -lse * log2(e), nd = -delta, padded to
      // whole tiles with (-inf, 0) so that queries past the end get p = 0:  x = s * s2 + nl,  dS = p * (dP + nd)
      bulk_load_1d(stats + s * 256, stats_g + (bh * 2 + 0) * Lp + it * T2_T, 512, &sf[s]);
      bulk_load_1d(stats + s * 256 + 128, stats_g + (bh * 2 + 1) * Lp + it * T2_T, 512, &sf[s]);
      if (prof && blockIdx.x == 0 && G == total_tiles - 1)
        printf("[t2 producer] total %lld clk; waits: st_empty %lld se %lld\n", clock64() - t_start, w0, w1);
      if (++it == nt) it = 0, ++n;
    }
    }
  } else if (warp == CW + 1) {
    if (elect_one() && total_tiles > 0) {
      // -------------------------------------------------------------------- MMA issuer
      // The issue loop of ONE thread is what paces this kernel when its MMAs are small (N = 64: 32 tensor clocks each): building
      // two 64-bit shared-memory descriptors per MMA from addresses costs ~10 uniform-datapath instructions (measured: ~105
      // clocks per MMA, 4 200 per tile).  Here only the LOW descriptor word moves: base words are formed once per tile and every
      // operand is `base + constant` (the address field counts 16-byte units and never carries out of its 14 bits).
      auto commit_t = [&](uint64_t* bar) {
        if (prof) {
          const long long c0 = clock64();
          umma_commit(bar);
          w5 += clock64() - c0;
        } else {
          umma_commit(bar);
        }
      };
      const uint32_t sb = smem_u32(smem);
      constexpr uint64_t HI128 = 0x40004040ull << 32;   // SWIZZLE_128B, SBO 1024, descriptor version 1
      constexpr uint64_t HI32 = 0xC0004010ull << 32;    // SWIZZLE_32B, SBO 256
      constexpr uint32_t LBO_K = 1u << 16;              // K-major: LBO 16 bytes (unused with swizzle)
      constexpr uint32_t LBO_MN = (T2_TILE >> 4) << 16; // MN-major: 64-element panels one tile apart
      auto dk = [&](uint32_t lo, uint32_t bytes) { return HI128 | (lo + (bytes >> 4)); };
      auto dn = [&](uint32_t lo, uint32_t bytes) { return HI32 | (lo + (bytes >> 4)); };
      const uint32_t kb_lo = ((sb + T2_OFF_KB) >> 4) | LBO_K, vb_lo = ((sb + T2_OFF_VB) >> 4) | LBO_K;
      constexpr uint32_t idesc_s = umma_idesc_bf16(T2_T, 64, 0, 0);   // all 64 columns: rows past the end are zero-filled
      constexpr uint32_t idesc = umma_idesc_bf16(T2_T, T2_HD, 0, 1);
      constexpr uint32_t idescb = umma_idesc_bf16(T2_T, 16, 0, 1);
      constexpr uint32_t idq = umma_idesc_bf16(T2_T, T2_HD, 1, 1);
      constexpr uint32_t idqb = umma_idesc_bf16(T2_T, 16, 1, 1);
      // S^T_h = K_j Q_h^T and dP^T_h = V_j dO_h^T of half hh of tile GG
      auto issue_sdp = [&](int GG, int nn, int it, int hh) {
        const int s = GG & 1, buf = nn % NST;
        const long long i0 = prof ? clock64() : 0;
        if (it * T2_T + hh * 64 < Lm && !(dbg & 128)) {
          const uint32_t k_lo = ((sb + off_k(buf)) >> 4) | LBO_K, v_lo = ((sb + off_v(buf)) >> 4) | LBO_K;
          const uint32_t q_lo = ((sb + T2_OFF_Q + s * T2_TILE + hh * 8192) >> 4) | LBO_K;
          const uint32_t do_lo = ((sb + T2_OFF_DO + s * T2_TILE + hh * 8192) >> 4) | LBO_K;
          const uint32_t ts = tmem_base + T2_TM_S + hh * 64, tp = tmem_base + T2_TM_DP + hh * 64;
          if (KVT) {
            // A = K_j / V_j resident in TMEM: an SS-form MMA of this size spends ~100 clocks fetching its 4 KB A block from
            // shared memory for 32 clocks of tensor work
#pragma unroll
            for (int k = 0; k < T2_HD / 16; ++k) t2_umma_ts(ts, tmem_base + T2_TM_K + 8 * k, dk(q_lo, k * 32), idesc_s, k != 0);
#pragma unroll
            for (int k = 0; k < T2_HD / 16; ++k) t2_umma_ts(tp, tmem_base + T2_TM_V + 8 * k, dk(do_lo, k * 32), idesc_s, k != 0);
          } else {
#pragma unroll
            for (int k = 0; k < T2_HD / 16; ++k) umma_bf16_ss(ts, dk(k_lo, k * 32), dk(q_lo, k * 32), idesc_s, k != 0);
            if (RB) umma_bf16_ss(ts, dn(kb_lo, 0), dn(((sb + T2_OFF_QB + s * T2_BT + hh * 2048) >> 4) | LBO_K, 0), idesc_s, 1);
#pragma unroll
            for (int k = 0; k < T2_HD / 16; ++k) umma_bf16_ss(tp, dk(v_lo, k * 32), dk(do_lo, k * 32), idesc_s, k != 0);
            if (RB) umma_bf16_ss(tp, dn(vb_lo, 0), dn(((sb + T2_OFF_DOB + s * T2_BT + hh * 2048) >> 4) | LBO_K, 0), idesc_s, 1);
          }
        }
        if (prof) w7 += clock64() - i0;
        commit_t(&s_full[hh]);
      };
      // operands of tile GG have landed (its streamed tile; for the first tile of an item also the stationary tiles)
      auto wait_operands = [&](int GG, int nn, int it) {
        if (it == 0) t2_wait_t(&st_full[nn % NST], (nn / NST) & 1, 42, spin, prof, w0);
        t2_wait_t(&sf[GG & 1], (GG >> 1) & 1, 43, spin, prof, w1);
        tc_fence_after();
        if (KVT && it == 0) {
          // the new item's K_j / V_j into TMEM, behind every score MMA of the previous item (issued earlier by this thread)
          const uint32_t k_lo = ((sb + off_k(nn % NST)) >> 4) | LBO_K, v_lo = ((sb + off_v(nn % NST)) >> 4) | LBO_K;
#pragma unroll
          for (int k = 0; k < T2_HD / 16; ++k) {
            t2_tmem_cp_128x256b(tmem_base + T2_TM_K + 8 * k, dk(k_lo, k * 32));
            t2_tmem_cp_128x256b(tmem_base + T2_TM_V + 8 * k, dk(v_lo, k * 32));
          }
        }
      };
      wait_operands(0, 0, 0);
      issue_sdp(0, 0, 0, 0);
      issue_sdp(0, 0, 0, 1);
      int n = 0, it = 0, jt = item_first % nt;   // item, streamed tile, key tile of the item
      for (int G = 0; G < total_tiles; ++G) {
        int n1 = n, it1 = it + 1;                // coordinates of tile G + 1
        if (it1 == nt) it1 = 0, ++n1;
        const int t0 = jt * T2_T;
        const int s = G & 1, buf = n % NST;
        const int nkv16 = (min(T2_T, Lm - t0) + 15) & ~15;
        // the next tile's scores go out behind this tile's dV / dK MMAs of the same half (same issuing thread: in order).
        // One stationary set (wide heads): the first tile of the next item has to wait for this item's last MMA instead.
        const bool next_now = G + 1 < total_tiles && (NST == 2 || it + 1 < nt);
        for (int hh = 0; hh < 2; ++hh) {
          t2_wait_t(&p_ready[hh], G & 1, 44, spin, prof, hh ? w3 : w2);
          if (it == 0 && hh == 0) t2_wait_t(acc_free, (n & 1) ^ 1, 45, spin, prof, w4);   // the previous item's dK / dV have been read out
          tc_fence_after();
          const int nq = min(64, Lm - (it * T2_T + hh * 64));
          const int ksteps = nq > 0 ? (nq + 15) >> 4 : 0;
          // dV += P^T_h dO_h, dK += dS^T_h Q_h: A from TMEM, B = the half's 64 rows of dO_i / Q_i as an MN-major operand
          const uint32_t do_lo = ((sb + T2_OFF_DO + s * T2_TILE + hh * 8192) >> 4) | LBO_MN;
          const uint32_t q_lo = ((sb + T2_OFF_Q + s * T2_TILE + hh * 8192) >> 4) | LBO_MN;
          const uint32_t dob_lo = ((sb + T2_OFF_DOB + s * T2_BT + hh * 2048) >> 4) | LBO_K;
          const uint32_t qb_lo = ((sb + T2_OFF_QB + s * T2_BT + hh * 2048) >> 4) | LBO_K;
          const uint32_t tp = tmem_base + T2_TM_S + hh * 64, tds = tmem_base + T2_TM_DP + hh * 64;
          const uint32_t first = (it | hh) != 0;
          const long long m0 = prof ? clock64() : 0;
          if (dbg & 64) {
          } else if (ksteps == 4) {   // the common case, straight-line
            t2_umma_ts(tmem_base + T2_TM_DV, tp + t2_acol<CG>(0), dk(do_lo, 0), idesc, first);
            t2_umma_ts(tmem_base + T2_TM_DV, tp + t2_acol<CG>(1), dk(do_lo, 2048), idesc, 1);
            t2_umma_ts(tmem_base + T2_TM_DV, tp + t2_acol<CG>(2), dk(do_lo, 4096), idesc, 1);
            t2_umma_ts(tmem_base + T2_TM_DV, tp + t2_acol<CG>(3), dk(do_lo, 6144), idesc, 1);
            if (RB) {
              t2_umma_ts(tmem_base + T2_TM_DVB, tp + t2_acol<CG>(0), dn(dob_lo, 0), idescb, first);
              t2_umma_ts(tmem_base + T2_TM_DVB, tp + t2_acol<CG>(1), dn(dob_lo, 512), idescb, 1);
              t2_umma_ts(tmem_base + T2_TM_DVB, tp + t2_acol<CG>(2), dn(dob_lo, 1024), idescb, 1);
              t2_umma_ts(tmem_base + T2_TM_DVB, tp + t2_acol<CG>(3), dn(dob_lo, 1536), idescb, 1);
            }
            t2_umma_ts(tmem_base + T2_TM_DK, tds + t2_acol<CG>(0), dk(q_lo, 0), idesc, first);
            t2_umma_ts(tmem_base + T2_TM_DK, tds + t2_acol<CG>(1), dk(q_lo, 2048), idesc, 1);
            t2_umma_ts(tmem_base + T2_TM_DK, tds + t2_acol<CG>(2), dk(q_lo, 4096), idesc, 1);
            t2_umma_ts(tmem_base + T2_TM_DK, tds + t2_acol<CG>(3), dk(q_lo, 6144), idesc, 1);
            if (RB) {
              t2_umma_ts(tmem_base + T2_TM_DKB, tds + t2_acol<CG>(0), dn(qb_lo, 0), idescb, first);
              t2_umma_ts(tmem_base + T2_TM_DKB, tds + t2_acol<CG>(1), dn(qb_lo, 512), idescb, 1);
              t2_umma_ts(tmem_base + T2_TM_DKB, tds + t2_acol<CG>(2), dn(qb_lo, 1024), idescb, 1);
              t2_umma_ts(tmem_base + T2_TM_DKB, tds + t2_acol<CG>(3), dn(qb_lo, 1536), idescb, 1);
            }
          } else {
            for (int kk = 0; kk < ksteps; ++kk) t2_umma_ts(tmem_base + T2_TM_DV, tp + t2_acol<CG>(kk), dk(do_lo, kk * 2048), idesc, first | kk);
            if (RB)
              for (int kk = 0; kk < ksteps; ++kk) t2_umma_ts(tmem_base + T2_TM_DVB, tp + t2_acol<CG>(kk), dn(dob_lo, kk * 512), idescb, first | kk);
            for (int kk = 0; kk < ksteps; ++kk) t2_umma_ts(tmem_base + T2_TM_DK, tds + t2_acol<CG>(kk), dk(q_lo, kk * 2048), idesc, first | kk);
            if (RB)
              for (int kk = 0; kk < ksteps; ++kk) t2_umma_ts(tmem_base + T2_TM_DKB, tds + t2_acol<CG>(kk), dn(qb_lo, kk * 512), idescb, first | kk);
          }
          if (prof) w6 += clock64() - m0;
          if (hh == 1) commit_t(&se[s]);   // last readers of Q_i / dO_i
          if (next_now) {
            if (hh == 0) wait_operands(G + 1, n1, it1);
            issue_sdp(G + 1, n1, it1, hh);
          }
        }
        // dQ_i (partial over this key tile) = dS_i K_j: A = dS^T in smem, rows = keys -> MN-major, M = 128 queries (two atoms).
        // The compute warps read the previous partial out right behind the half they have just finished.
        if (G > 0) {
          t2_wait_t(dq_free, (G - 1) & 1, 47, spin, prof, w4);
          tc_fence_after();
        }
        const long long q0t = prof ? clock64() : 0;
        {
          const uint32_t ds_lo = ((sb + T2_OFF_DS + (G & 1) * 2 * T2_TILE) >> 4) | LBO_MN;
          const uint32_t kmn_lo = ((sb + off_k(buf)) >> 4) | LBO_MN;
          if (nkv16 == T2_T && !(dbg & 8)) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(tmem_base + T2_TM_DQ, dk(ds_lo, kk * 2048), dk(kmn_lo, kk * 2048), idq, kk != 0);
            if (RB) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(tmem_base + T2_TM_DQB, dk(ds_lo, kk * 2048), dn(kb_lo, kk * 512), idqb, kk != 0);
            }
          } else {
            for (int kk = 0; kk < ((dbg & 8) ? 0 : nkv16 / 16); ++kk)
              umma_bf16_ss(tmem_base + T2_TM_DQ, dk(ds_lo, kk * 2048), dk(kmn_lo, kk * 2048), idq, kk != 0);
            if (RB)
              for (int kk = 0; kk < nkv16 / 16; ++kk)
                umma_bf16_ss(tmem_base + T2_TM_DQB, dk(ds_lo, kk * 2048), dn(kb_lo, kk * 512), idqb, kk != 0);
          }
        }
        if (prof) w8 += clock64() - q0t;
        commit_t(dq_full);
        if (it == nt - 1) commit_t(&st_empty[buf]);
        if (G + 1 < total_tiles && !next_now) {
          wait_operands(G + 1, n1, it1);
          issue_sdp(G + 1, n1, it1, 0);
          issue_sdp(G + 1, n1, it1, 1);
        }
        if (it1 == 0 && ++jt == nt) jt = 0;
        n = n1, it = it1;
      }
      if (prof && blockIdx.x == 0)
        printf("[t2 mma] tiles %d total %lld clk; waits: st_full %lld sf %lld p_ready0 %lld p_ready1 %lld acc/dq_free %lld; commits %lld dV/dK issue %lld S/dP issue %lld dQ issue %lld\n", total_tiles,
               clock64() - t_start, w0, w1, w2, w3, w4, w5, w6, w7, w8);
    }
  } else {
    // ---------------------------------------------------------------------- compute warps
    const int quad = warp & 3;
    const int csel = warp >> 2;           // which CG of a half's 64 query columns
    const int r = quad * 32 + lane;       // key row inside the tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const uint64_t sc2 = f2_pack(s2, s2);
    // per-warp arrival on a barrier of count CW: every lane has fenced its own TMEM / shared-memory accesses
    auto warp_arrive = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // dQ partial #q (of the query tile at q0 of (h, b)): TMEM -> fp32 rows in this warp's own staging region -> one TMA
    // reduce-add per warp
    auto drain_dq = [&](int q, int q0, int h, int b, bool first_key_tile) {
      t2_wait_t(dq_full, q & 1, 50, spin, prof, w0);
      tc_fence_after();
      if (q0 + quad * 32 < Lm && !(dbg & 2)) {   // warp-uniform: otherwise nothing but clipped rows
        uint32_t o[CG];
        t2_tmem_ld(tmem_base + t_lane + T2_TM_DQ + csel * CG, o);
        // the warp's previous store / reduce-add has COMPLETED (not only read its staging rows): the same warp owns this block
        // of the accumulator for every key tile, and its first key tile's plain store must land before the adds that follow
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
        const uint32_t region = T2_OFF_DQST + (csel * 4 + quad) * (32 * CG * 4);
        const uint32_t dst = smem_u32(smem + region);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < CG / 4; ++k)
          sts128(dst + (CG == 32 ? sw128_offset(lane, k) : t2_sw64_offset(lane, k)), make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
        if (RB && csel == 0) {
          uint32_t ob[16];
          tmem_ld_x16(tmem_base + t_lane + T2_TM_DQB, ob);
          tmem_ld_wait();
          const uint32_t dstb = smem_u32(smem + T2_OFF_NST + quad * (32 * 64)) + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) sts128(dstb + k * 16, make_uint4(ob[4 * k], ob[4 * k + 1], ob[4 * k + 2], ob[4 * k + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(dbg & 1)) {
          if (first_key_tile) {   // no zero-filled accumulator: the first key tile of an (image, head) writes, the others add
            tma_store_4d(&tmDQF, smem + region, csel * CG, h, q0 + quad * 32, b);
            if (RB && csel == 0) tma_store_4d(&tmDQFb, smem + T2_OFF_NST + quad * (32 * 64), 64, h, q0 + quad * 32, b);
          } else {
            tma_reduce_add_4d(&tmDQF, smem + region, csel * CG, h, q0 + quad * 32, b);
            if (RB && csel == 0) tma_reduce_add_4d(&tmDQFb, smem + T2_OFF_NST + quad * (32 * 64), 64, h, q0 + quad * 32, b);
          }
          tma_store_commit();
        }
      }
      warp_arrive(dq_free);
    };
    // Item epilogue: the item's last dQ partial (its dq_full also says that every MMA of the item has completed), then dK / dV:
    // accumulators -> (+ rank-1 terms of the remainder token) -> bf16 -> staging in dS buffer `sbuf` -> TMA store.
    auto item_epilogue = [&](int t0, int h, int b, int q_last, int sbuf) {
      const long long bh = static_cast<long long>(b) * H + h;
      drain_dq(q_last, (nt - 1) * T2_T, h, b, t0 == 0);
      if (RB) {                    // the narrow output tiles below reuse the narrow dQ staging
        if (lane == 0) tma_store_wait_read<0>();
        named_bar_sync(1, 32 * CW);
      }
      // this warp's share of the [128 keys x (dK | dV)] read-out: tensor `which` (0: dK, scaled; 1: dV), columns [col0, col0 + ncols)
      constexpr int ncols = 128 / NW;
      const int which = csel / (NW / 2), col0 = (csel % (NW / 2)) * ncols;
      const bool tail = ws != nullptr;
      float tcoef = 0.f;
      const __nv_bfloat16* tvec = nullptr;
      if (tail) {
        const float* wsb = ws + bh * 3 * Lm;
        const long long trow = static_cast<long long>(b) * L + Lm;   // the remainder token
        if (which == 0) {
          tcoef = wsb[2 * Lm + t0 + r];                              // dS(t, j)
          tvec = qkv_g + (trow * 3 * H + h) * hd_g;                  // q_t
        } else {
          tcoef = wsb[Lm + t0 + r];                                  // P(t, j)
          tvec = dout_g + (trow * H + h) * hd_g;                     // dO_t
        }
      }
      const float mul = which == 0 ? scale : 1.f;
      const uint32_t src = which == 0 ? T2_TM_DK : T2_TM_DV;
      uint8_t* stage_tiles = smem + T2_OFF_DS + sbuf * 2 * T2_TILE;
      const uint32_t stage = smem_u32(stage_tiles + which * T2_TILE);   // atom 0 <- dK, atom 1 <- dV
#pragma unroll 1
      for (int c = col0; c < col0 + ncols; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(tmem_base + t_lane + src + c, o);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(tvec + c) + v);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o[8 * v + 2 * k] = __float_as_uint(fmaf(tcoef, bf16_lo(w4[k]), __uint_as_float(o[8 * v + 2 * k])));
              o[8 * v + 2 * k + 1] = __float_as_uint(fmaf(tcoef, bf16_hi(w4[k]), __uint_as_float(o[8 * v + 2 * k + 1])));
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sts128(stage + sw128_offset(r, (c >> 3) + k),
                 make_uint4(pack_bf16x2(__uint_as_float(o[8 * k]) * mul, __uint_as_float(o[8 * k + 1]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * k + 2]) * mul, __uint_as_float(o[8 * k + 3]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * k + 4]) * mul, __uint_as_float(o[8 * k + 5]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * k + 6]) * mul, __uint_as_float(o[8 * k + 7]) * mul)));
      }
      if (RB && col0 == 0) {
        uint32_t ob[16];
        tmem_ld_x16(tmem_base + t_lane + (which == 0 ? T2_TM_DKB : T2_TM_DVB), ob);
        tmem_ld_wait();
        if (tail) {
          const int nb = hd_g - 64 < 16 ? hd_g - 64 : 16;   // dims 64 .. hd (the columns past hd stay zero)
          for (int v = 0; v < nb / 8; ++v) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(tvec + 64) + v);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ob[8 * v + 2 * k] = __float_as_uint(fmaf(tcoef, bf16_lo(w4[k]), __uint_as_float(ob[8 * v + 2 * k])));
              ob[8 * v + 2 * k + 1] = __float_as_uint(fmaf(tcoef, bf16_hi(w4[k]), __uint_as_float(ob[8 * v + 2 * k + 1])));
            }
          }
        }
        const uint32_t dstb = smem_u32(smem + T2_OFF_NST + which * T2_BT);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          sts128(dstb + t2_sw32_offset(r, c),
                 make_uint4(pack_bf16x2(__uint_as_float(ob[8 * c]) * mul, __uint_as_float(ob[8 * c + 1]) * mul),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 2]) * mul, __uint_as_float(ob[8 * c + 3]) * mul),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 4]) * mul, __uint_as_float(ob[8 * c + 5]) * mul),
                            pack_bf16x2(__uint_as_float(ob[8 * c + 6]) * mul, __uint_as_float(ob[8 * c + 7]) * mul)));
      }
      warp_arrive(acc_free);   // dK / dV read: the next item's first accumulating MMA may overwrite them
      fence_proxy_async_smem();
      named_bar_sync(1, 32 * CW);
      if (threadIdx.x == 0) {
        tma_store_4d(&tmDQKV, stage_tiles, 0, H + h, t0, b);
        tma_store_4d(&tmDQKV, stage_tiles + T2_TILE, 0, 2 * H + h, t0, b);
        if (RB) {
          tma_store_4d(&tmDQKVb, smem + T2_OFF_NST, 64, H + h, t0, b);
          tma_store_4d(&tmDQKVb, smem + T2_OFF_NST + T2_BT, 64, 2 * H + h, t0, b);
        }
        tma_store_commit();
      }
    };

    bool staged = false;   // an item epilogue's output tiles sit in the dS buffer that the NEXT tile rewrites
    int n = 0, it = 0, t0 = 0, h = 0, b = 0, pt0 = 0, ph = 0, pb = 0;   // current item and the one before it
    for (int G = 0; G < total_tiles; ++G) {
      if (it == 0) {
        pt0 = t0, ph = h, pb = b;
        const int item = item_first + n;
        t0 = (item % nt) * T2_T, h = (item / nt) % H, b = item / (nt * H);
      }
      const int jr = t0 + r;               // this thread's key
      const bool row_ok = jr < Lm;
      const int nkv16 = (min(T2_T, Lm - t0) + 15) & ~15;   // key rows that any MMA result depends on
      const int s = G & 1;
      if (staged) {
        if (threadIdx.x == 0) tma_store_wait_read<0>();
        named_bar_sync(1, 32 * CW);
        staged = false;
      }
      // wide heads (one stationary set): this tile's scores cannot be issued before the previous item's last MMA, so its
      // epilogue goes first; hd = 64: between the two halves, when half a's dV / dK MMAs have nothing else to wait for
      if (RB && it == 0 && n > 0) {
        item_epilogue(pt0, ph, pb, G - 1, (G + 1) & 1);
        staged = true;
      }
      // this tile's statistics are in smem: the MMA thread saw sf[s] complete before it issued the scores s_full[] announces
      const uint32_t st_addr = smem_u32(stats + s * 256);
      const uint32_t ds_base = smem_u32(smem + T2_OFF_DS + (G & 1) * 2 * T2_TILE);
      for (int hh = 0; hh < 2; ++hh) {
        if (!RB && hh == 1 && it == 0 && n > 0) {
          item_epilogue(pt0, ph, pb, G - 1, (G + 1) & 1);
          staged = true;
        }
        t2_wait_t(&s_full[hh], G & 1, 52, spin, prof, hh ? w3 : w2);
        tc_fence_after();
        const int c0 = hh * 64 + csel * CG;     // first query column (inside the tile) of this thread's chunk
        const int q0c = it * T2_T + c0;
        // warp-uniform: chunks past the last query are read by no MMA; key rows past nkv16 only reach clipped dK / dV rows
        if (q0c < Lm && quad * 32 < nkv16 && !(dbg & 4)) {
          uint32_t sv[CG], dv[CG];
          t2_tmem_ld(tmem_base + t_lane + T2_TM_S + c0, sv);
          t2_tmem_ld(tmem_base + t_lane + T2_TM_DP + c0, dv);
          tmem_ld_wait();
          uint32_t pp[CG / 2], dd[CG / 2];
          if (!causal || q0c >= jr) {   // queries past the end: nl = -inf and zero-filled Q / dO rows give exact zeros
#pragma unroll
            for (int v = 0; v < CG / 4; ++v) {
              const float4 nl = lds_f32x4(st_addr + (c0 + 4 * v) * 4);
              const float4 nd = lds_f32x4(st_addr + (128 + c0 + 4 * v) * 4);
              float x0, x1, x2, x3, d0, d1, d2, d3;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[4 * v]), __uint_as_float(sv[4 * v + 1])), sc2, f2_pack(nl.x, nl.y)), x0, x1);
              f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[4 * v + 2]), __uint_as_float(sv[4 * v + 3])), sc2, f2_pack(nl.z, nl.w)), x2, x3);
              const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
              pp[2 * v] = pack_bf16x2(p0, p1);
              pp[2 * v + 1] = pack_bf16x2(p2, p3);
              f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(dv[4 * v]), __uint_as_float(dv[4 * v + 1])), f2_pack(nd.x, nd.y))), d0, d1);
              f2_unpack(f2_mul(f2_pack(p2, p3), f2_add(f2_pack(__uint_as_float(dv[4 * v + 2]), __uint_as_float(dv[4 * v + 3])), f2_pack(nd.z, nd.w))), d2, d3);
              dd[2 * v] = pack_bf16x2(d0, d1);
              dd[2 * v + 1] = pack_bf16x2(d2, d3);
            }
            // keys past the end of the sequence: K rows are zero-filled (a finite score), but nothing of them may survive
            if (!row_ok) {
#pragma unroll
              for (int j = 0; j < CG / 2; ++j) pp[j] = dd[j] = 0u;
            }
          } else {
            // the causal mask (transformer.py:757-763: query q sees keys <= q): exact zeros
#pragma unroll
            for (int j = 0; j < CG / 2; ++j) {
              const int q = q0c + 2 * j;
              const bool ok0 = row_ok && q < Lm && (!causal || q >= jr);
              const bool ok1 = row_ok && q + 1 < Lm && (!causal || q + 1 >= jr);
              const float2 nl = lds_f32x2(st_addr + (c0 + 2 * j) * 4);
              const float2 nd = lds_f32x2(st_addr + (128 + c0 + 2 * j) * 4);
              const float p0 = ok0 ? fast_exp2(fmaf(__uint_as_float(sv[2 * j]), s2, nl.x)) : 0.f;
              const float p1 = ok1 ? fast_exp2(fmaf(__uint_as_float(sv[2 * j + 1]), s2, nl.y)) : 0.f;
              const float d0 = ok0 ? p0 * (__uint_as_float(dv[2 * j]) + nd.x) : 0.f;
              const float d1 = ok1 ? p1 * (__uint_as_float(dv[2 * j + 1]) + nd.y) : 0.f;
              pp[j] = pack_bf16x2(p0, p1);
              dd[j] = pack_bf16x2(d0, d1);
            }
          }
          t2_tmem_st(tmem_base + t_lane + T2_TM_S + c0, pp);    // P^T: A operand of dV
          t2_tmem_st(tmem_base + t_lane + T2_TM_DP + c0, dd);   // dS^T: A operand of dK
#pragma unroll
          for (int k = 0; k < CG / 8; ++k)                        // dS^T row of this key: A operand of dQ
            sts128(ds_base + hh * T2_TILE + sw128_offset(r, csel * (CG / 8) + k),
                   make_uint4(dd[4 * k], dd[4 * k + 1], dd[4 * k + 2], dd[4 * k + 3]));
          tmem_st_wait();
          fence_proxy_async_smem();
        }
        warp_arrive(&p_ready[hh]);
      }
      // the previous tile's dQ partial (same item): its MMA went out a whole tile ago, and reading it here, behind this tile's
      // last half, lets dQ of THIS tile start as soon as its dS is complete
      if (it > 0) drain_dq(G - 1, (it - 1) * T2_T, h, b, t0 == 0);
      if (++it == nt) it = 0, ++n;
    }
    if (total_tiles > 0) item_epilogue(t0, h, b, total_tiles - 1, (total_tiles + 1) & 1);
    if (lane == 0) tma_store_wait_all<0>();
    if (prof && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32 * CW - 32))
      printf("[t2 compute warp %d] total %lld clk; waits: dq_full %lld sf %lld s_full0 %lld s_full1 %lld\n", warp, clock64() - t_start, w0, w1,
             w2, w3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CW + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Launch (called by attention_bwd.cu's host code, which owns the operand tensor maps, the delta / tail / conversion kernels).
// acc: the fp32 dQ accumulator [B, L, H, hd]; warps16: 16 compute warps (16 query columns of a half each) instead of 8.
template <int RB, int NW>
static int launch_t(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const CUtensorMap& tmDQKV, const CUtensorMap& tmQKVb,
                    const CUtensorMap& tmDOb, const CUtensorMap& tmDQKVb, float* acc, int B, const float* stats, int Lp, int L, int H,
                    float scale, int items, int grid, int per, int causal, int Lm, const float* ws, const __nv_bfloat16* qkv_g,
                    const __nv_bfloat16* dout_g, int hd, cudaStream_t s) {
  constexpr int CG = 64 / NW;
  CUtensorMap tmDQF, tmDQFb;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 4, (uint64_t)H * hd * 4, (uint64_t)L * H * hd * 4};
    const uint32_t box[4] = {CG, 1, 32, 1};
    const uint32_t boxb[4] = {16, 1, 32, 1};
    if ((rc = make_tmap_nd_f32(&tmDQF, acc, 4, dims, strides, box, CG == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap_nd_f32(&tmDQFb, acc, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_t_kernel<RB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention_bwd_t): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  attention_bwd_t_kernel<RB, NW><<<grid, t2_threads(NW), T2_SMEM_BYTES, s>>>(tmQKV, tmDO, tmDQKV, tmQKVb, tmDOb, tmDQKVb, tmDQF, tmDQFb, stats, Lp,
                                                                         L, H, scale, items, per, causal, Lm, ws, qkv_g, dout_g, hd);
  return check_launch("attention_bwd_t_kernel");
}

int launch_attention_bwd_t(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const CUtensorMap& tmDQKV,
                           const CUtensorMap& tmQKVb, const CUtensorMap& tmDOb, const CUtensorMap& tmDQKVb, float* acc, int B,
                           const float* stats, int Lp, int L, int H, float scale, int items, int grid, int causal, int Lm,
                           const float* ws, const __nv_bfloat16* qkv_g, const __nv_bfloat16* dout_g, int hd, bool warps16,
                           cudaStream_t s) {
  // whole (image, head)s per CTA: the first key tile's STORE into the dQ accumulator and the adds of the others stay in one CTA
  const int nt = (Lm + T2_T - 1) / T2_T;
  const int per = ((items + grid - 1) / grid + nt - 1) / nt * nt;
#define T2_LAUNCH(RB, NW) \
  launch_t<RB, NW>(tmQKV, tmDO, tmDQKV, tmQKVb, tmDOb, tmDQKVb, acc, B, stats, Lp, L, H, scale, items, grid, per, causal, Lm, ws, qkv_g, dout_g, hd, s)
  if (hd > T2_HD) return warps16 ? T2_LAUNCH(16, 4) : T2_LAUNCH(16, 2);
  return warps16 ? T2_LAUNCH(0, 4) : T2_LAUNCH(0, 2);
#undef T2_LAUNCH
}

}  // namespace ovk
