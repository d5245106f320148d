// Backward of the fused self-attention (what autograd derives from nn.MultiheadAttention's softmax(QK^T/sqrt(hd))V,
// open_clip/transformer.py:225,239-252) on tcgen05, flash-style: scores are recomputed tile by tile from Q, K and the
// saved log-sum-exp; nothing of size L x L is ever stored.
//   P = exp(S*scale - LSE),  dP = dO V^T,  D = rowsum(dO . O),  dS = P . (dP - D)
//   dV = P^T dO,  dK = scale * dS^T Q,  dQ = scale * dS K
// Two launches, no atomics, outputs written once as bf16 straight into the packed d(qkv) [B, L, 3, H, 64] layout that
// the in_proj backward GEMMs read.  Both are PERSISTENT: one CTA per SM walks work items (tile, head, batch) with stride
// gridDim.x, keeping its TMEM allocation, barriers and the TMA ring alive, and (hd = 64) prefetching the next item's
// stationary tiles into a second buffer while the current item computes:
//   MODE_DQ  : item = (128-query tile, head, batch); streams K/V tiles; also computes D and stores it.
//   MODE_DKV : item = (128-key tile, head, batch);   streams Q/dO tiles; reads D.
// Per streamed tile the MMA warp issues S / dP of the NEXT tile before the accumulating MMAs of the current one: both
// wait for the same event (the softmax warps are done with S / dP), and the softmax warps need the next S / dP first.
// The bf16 P / dS tile lives in shared memory as [q rows][kv cols] with 128-byte swizzled rows: the same bytes serve as
// a K-major A operand (dS K) and as an MN-major A operand (P^T dO, dS^T Q), so no transpose is ever materialised.
//
//   MODE_FUSED (default path): ONE pass over the score tiles instead of two.  The DKV walk above, plus per streamed query
//   tile the fifth MMA dQ_i(partial) = dS K_j into its own TMEM columns, which the softmax warps drain (TMEM -> fp32 rows
//   in shared memory, each warp its own 32 rows) and add into an fp32 [B, L, H, hd] scratch with TMA reduce-add
//   (cp.reduce.async.bulk.tensor .add): S, dP, P and dS are computed once (10 instead of 14 B H L^2 hd FLOPs, half the
//   exponentials).  A CTA walks a CONTIGUOUS run of items (key tile fastest), so the partial sums of one (image, head)
//   meet in L2.  D = rowsum(dO . O) comes from attention_bwd_delta_kernel, and attention_bwd_dq_convert_kernel turns the
//   scratch into the bf16 q slot of d(qkv) (x scale, + the remainder token's rank-1 term).
#include <algorithm>

#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

constexpr int AB_T = 128;                   // tile rows (queries or keys)
constexpr int AB_HD = 64;
constexpr int AB_TILE = AB_T * 128;         // 16 KB: [128 rows x 64 bf16]
constexpr int AB_SOFTMAX_WARPS = 8;
constexpr int AB_THREADS = 32 * (AB_SOFTMAX_WARPS + 2);  // + TMA producer warp + MMA warp
constexpr int AB_OFF_ST0 = 0;               // stationary tile 0 (DQ: Q_i   | DKV: K_j)
constexpr int AB_OFF_ST1 = AB_OFF_ST0 + AB_TILE;  //            1 (DQ: dO_i  | DKV: V_j)
constexpr int AB_OFF_SA = AB_OFF_ST1 + AB_TILE;   // streamed A x2 (DQ: K_j | DKV: Q_i)
constexpr int AB_OFF_SB = AB_OFF_SA + 2 * AB_TILE;  // streamed B x2 (DQ: V_j | DKV: dO_i)
constexpr int AB_OFF_P = AB_OFF_SB + 2 * AB_TILE;   // P  bf16 [128 x 128] as two 64-column atoms
constexpr int AB_OFF_DS = AB_OFF_P + 2 * AB_TILE;   // dS bf16, same layout
constexpr int AB_OFF_X = AB_OFF_DS + 2 * AB_TILE;   // extra tile (DQ: O_i for D = rowsum(dO . O))
constexpr int AB_OFF_ST2 = AB_OFF_X + AB_TILE;      // second set of stationary tiles (ST0, ST1, X) for the NEXT item, hd = 64 only
constexpr int AB_NUM_BARS = 13;
// hd = 64: two stationary sets; wider heads keep one (their narrow tiles need the room)
constexpr int ab_off_bar(int rb) { return rb ? AB_OFF_ST2 : AB_OFF_ST2 + 3 * AB_TILE; }
constexpr int AB_SMEM_BYTES = ab_off_bar(0) + AB_NUM_BARS * 8 + 16;
constexpr int AB_SMEM_BYTES_NARROW_BASE = ab_off_bar(16) + AB_NUM_BARS * 8 + 16;
// head widths 64 < hd <= 80: the dims past 64 ride along as narrow [128 rows x 32 B] SWIZZLE_32B tiles (see attention.cu)
constexpr int AB_BT = 128 * 32;
constexpr int AB_OFF_ST0B = (AB_SMEM_BYTES_NARROW_BASE + 1023) / 1024 * 1024;
constexpr int AB_OFF_ST1B = AB_OFF_ST0B + AB_BT;
constexpr int AB_OFF_SAB = AB_OFF_ST1B + AB_BT;       // x2
constexpr int AB_OFF_SBB = AB_OFF_SAB + 2 * AB_BT;    // x2
constexpr int AB_OFF_XB = AB_OFF_SBB + 2 * AB_BT;
constexpr int AB_OFF_OUTB = AB_OFF_XB + AB_BT;        // x2 output staging
constexpr int AB_SMEM_BYTES_RB = AB_OFF_OUTB + 2 * AB_BT;
constexpr uint32_t AB_TM_ACC0B = 448, AB_TM_ACC1B = 464;
constexpr uint32_t AB_TM_DQ = 384, AB_TM_DQB = 480;   // MODE_FUSED: per-iteration dQ partial (64 + 16 columns)
// MODE_FUSED: fp32 staging of the dQ partial, one [32 rows x 128 B] SWIZZLE_128B region per softmax warp.
//   hd = 64: tiles 12, 13 (the second stationary set moves to tiles 10, 11: X is not loaded in this mode)
//   wide heads: X (column half 0), XB + three regions behind the narrow tiles (column half 1); the 16 narrow columns are
//   staged in OUTB (dense 64-byte rows), which the item's epilogue reuses for its narrow output tiles.
constexpr int AB_DQ_REGION = 32 * 128;
constexpr int AB_OFF_DQST = AB_OFF_ST2 + AB_TILE;
constexpr int AB_OFF_DQX = AB_SMEM_BYTES_RB;
constexpr int AB_SMEM_BYTES_RB_FUSED = AB_OFF_DQX + 3 * AB_DQ_REGION;
static_assert(AB_SMEM_BYTES_RB_FUSED <= 232448 && AB_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(AB_OFF_DQX % 1024 == 0, "dQ staging regions must be 1024-byte aligned");
constexpr uint32_t AB_SW32 = 6;
__device__ __forceinline__ uint64_t ab_desc_sw32(uint32_t saddr) { return umma_desc(saddr, 16, 256, AB_SW32); }
__device__ __forceinline__ uint32_t ab_sw32_offset(uint32_t row, uint32_t chunk) {
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}
constexpr int AB_TMEM_COLS = 512;
constexpr uint32_t AB_TM_S = 0, AB_TM_DP = 128, AB_TM_ACC0 = 256, AB_TM_ACC1 = 320;
enum { MODE_DQ = 0, MODE_DKV = 1, MODE_FUSED = 2 };

template <int MODE, int RB>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQKV,
                     const __grid_constant__ CUtensorMap tmQKVb, const __grid_constant__ CUtensorMap tmOb,
                     const __grid_constant__ CUtensorMap tmDOb, const __grid_constant__ CUtensorMap tmDQKVb,
                     const float* __restrict__ lse, float* __restrict__ delta, int L, int H, float scale, int total_items,
                     int causal, int Lm, const float* __restrict__ ws, const __nv_bfloat16* __restrict__ qkv_g,
                     const __nv_bfloat16* __restrict__ dout_g, int hd_g, const __grid_constant__ CUtensorMap tmDQF,
                     const __grid_constant__ CUtensorMap tmDQFb, int per) {
  // Lm < L (= L - 1, remainder token of L = 128 k + 1): the tiles cover tokens [0, Lm) only; the remainder token's query
  // row and key / value row are computed by attention_bwd_tail_kernel, which also leaves, per (image, head), the three
  // vectors the epilogues below add as rank-1 terms: ws[0][i] = dS(i, t), ws[1][j] = P(t, j), ws[2][j] = dS(t, j).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[ovk] attention_bwd: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  constexpr int NST = RB ? 1 : 2;   // stationary tile sets
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ab_off_bar(RB));
  uint64_t* st_full = bars + 0;   // [2] stationary tiles of an item have landed
  uint64_t* st_empty = bars + 2;  // [2] ... and are no longer read (MMAs of the item done; DQ: D computed)
  uint64_t* sf = bars + 4;        // stream_full[2]
  uint64_t* se = bars + 6;        // stream_empty[2]
  uint64_t* s_full = bars + 8;    // S and dP of this iteration are in TMEM
  uint64_t* pds_ready = bars + 9; // softmax wrote P / dS (and finished reading S / dP)
  uint64_t* mma_done = bars + 10; // accumulating MMAs of this iteration finished (P / dS smem reusable)
  uint64_t* acc_free = bars + 11; // the epilogue has read the accumulators of the previous item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ab_off_bar(RB) + AB_NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int nt = (Lm + AB_T - 1) / AB_T;  // tiles per sequence: stationary tiles of an item, streamed tiles per item
  const float s2 = scale * 1.4426950408889634f;
  // stationary tile `which` (0: ST0, 1: ST1, 2: X) of set `buf`
  auto st_off = [](int buf, int which) {
    if (MODE == MODE_FUSED && buf == 1) return which == 0 ? AB_OFF_X : AB_OFF_ST2;   // tiles 12, 13: dQ staging
    return buf == 0 ? (which == 0 ? AB_OFF_ST0 : which == 1 ? AB_OFF_ST1 : AB_OFF_X) : AB_OFF_ST2 + which * AB_TILE;
  };
  // work items of this CTA: MODE_FUSED walks a contiguous run (key tiles of one (image, head) back to back: their dQ
  // partial sums meet in L2), the two-pass kernels stride by the grid
  const int item_first = (MODE == MODE_FUSED) ? blockIdx.x * per : blockIdx.x;
  const int item_last = (MODE == MODE_FUSED) ? min(total_items, item_first + per) : total_items;
  const int item_step = (MODE == MODE_FUSED) ? 1 : gridDim.x;

  if (warp == AB_SOFTMAX_WARPS && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
    if (MODE == MODE_FUSED) tma_prefetch_desc(&tmDQF);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1 + (MODE == MODE_DQ ? 32 * AB_SOFTMAX_WARPS : 0));
      mbar_init(&sf[i], 1);
      mbar_init(&se[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pds_ready, 32 * AB_SOFTMAX_WARPS);
    mbar_init(mma_done, 1);
    mbar_init(acc_free, 32 * AB_SOFTMAX_WARPS);
    fence_mbar_init();
  }
  if (warp == AB_SOFTMAX_WARPS + 1) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AB_SOFTMAX_WARPS) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      int n = 0, g = 0;
      for (int item = item_first; item < item_last; item += item_step, ++n) {
        const int t0 = (item % nt) * AB_T, h = (item / nt) % H, b = item / (nt * H);
        const int buf = n % NST;
        uint64_t* stf = &st_full[buf];
        mbar_wait(&st_empty[buf], ((n / NST) & 1) ^ 1, 19);
        if (MODE == MODE_DQ) {
          mbar_arrive_expect_tx(stf, 3 * AB_TILE + (RB ? 3 * AB_BT : 0));
          tma_load_4d(smem + st_off(buf, 0), &tmQKV, stf, 0, h, t0, b);          // Q_i
          tma_load_4d(smem + st_off(buf, 1), &tmDO, stf, 0, h, t0, b);           // dO_i
          tma_load_4d(smem + st_off(buf, 2), &tmO, stf, 0, h, t0, b);            // O_i
          if (RB) {
            tma_load_4d(smem + AB_OFF_ST0B, &tmQKVb, stf, 64, h, t0, b);
            tma_load_4d(smem + AB_OFF_ST1B, &tmDOb, stf, 64, h, t0, b);
            tma_load_4d(smem + AB_OFF_XB, &tmOb, stf, 64, h, t0, b);
          }
        } else {
          mbar_arrive_expect_tx(stf, 2 * AB_TILE + (RB ? 2 * AB_BT : 0));
          tma_load_4d(smem + st_off(buf, 0), &tmQKV, stf, 0, H + h, t0, b);      // K_j
          tma_load_4d(smem + st_off(buf, 1), &tmQKV, stf, 0, 2 * H + h, t0, b);  // V_j
          if (RB) {
            tma_load_4d(smem + AB_OFF_ST0B, &tmQKVb, stf, 64, H + h, t0, b);
            tma_load_4d(smem + AB_OFF_ST1B, &tmQKVb, stf, 64, 2 * H + h, t0, b);
          }
        }
        for (int it = 0; it < nt; ++it, ++g) {
          const int s = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          mbar_wait(&se[s], ph ^ 1, 20);
          mbar_arrive_expect_tx(&sf[s], 2 * AB_TILE + (RB ? 2 * AB_BT : 0));
          if (MODE == MODE_DQ) {
            tma_load_4d(smem + AB_OFF_SA + s * AB_TILE, &tmQKV, &sf[s], 0, H + h, it * AB_T, b);      // K_j
            tma_load_4d(smem + AB_OFF_SB + s * AB_TILE, &tmQKV, &sf[s], 0, 2 * H + h, it * AB_T, b);  // V_j
            if (RB) {
              tma_load_4d(smem + AB_OFF_SAB + s * AB_BT, &tmQKVb, &sf[s], 64, H + h, it * AB_T, b);
              tma_load_4d(smem + AB_OFF_SBB + s * AB_BT, &tmQKVb, &sf[s], 64, 2 * H + h, it * AB_T, b);
            }
          } else {
            tma_load_4d(smem + AB_OFF_SA + s * AB_TILE, &tmQKV, &sf[s], 0, h, it * AB_T, b);          // Q_i
            tma_load_4d(smem + AB_OFF_SB + s * AB_TILE, &tmDO, &sf[s], 0, h, it * AB_T, b);           // dO_i
            if (RB) {
              tma_load_4d(smem + AB_OFF_SAB + s * AB_BT, &tmQKVb, &sf[s], 64, h, it * AB_T, b);
              tma_load_4d(smem + AB_OFF_SBB + s * AB_BT, &tmDOb, &sf[s], 64, h, it * AB_T, b);
            }
          }
        }
      }
    }
  } else if (warp == AB_SOFTMAX_WARPS + 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t p_addr = smem_u32(smem + AB_OFF_P), ds_addr = smem_u32(smem + AB_OFF_DS);
      const uint32_t st0b = smem_u32(smem + AB_OFF_ST0B), st1b = smem_u32(smem + AB_OFF_ST1B);
      int n = 0, g = 0;
      for (int item = item_first; item < item_last; item += item_step, ++n) {
        const int t0 = (item % nt) * AB_T;
        const int buf = n % NST;
        const uint32_t st0 = smem_u32(smem + st_off(buf, 0)), st1 = smem_u32(smem + st_off(buf, 1));
        mbar_wait(&st_full[buf], (n / NST) & 1, 21);
        // S = Q K^T and dP = dO V^T of streamed tile `it` (global iteration gg) into TMEM
        auto issue_s = [&](int it, int gg) {
          const int s = gg & 1;
          const uint32_t sa = smem_u32(smem + AB_OFF_SA + s * AB_TILE), sb = smem_u32(smem + AB_OFF_SB + s * AB_TILE);
          const uint32_t sab = smem_u32(smem + AB_OFF_SAB + s * AB_BT), sbb = smem_u32(smem + AB_OFF_SBB + s * AB_BT);
          const int kv0 = (MODE == MODE_DQ) ? it * AB_T : t0;
          const int nkv = (min(AB_T, Lm - kv0) + 15) & ~15;
          const uint32_t q_addr = (MODE == MODE_DQ) ? st0 : sa;
          const uint32_t do_addr = (MODE == MODE_DQ) ? st1 : sb;
          const uint32_t k_addr = (MODE == MODE_DQ) ? sa : st0;
          const uint32_t v_addr = (MODE == MODE_DQ) ? sb : st1;
          const uint32_t qb_addr = (MODE == MODE_DQ) ? st0b : sab;
          const uint32_t dob_addr = (MODE == MODE_DQ) ? st1b : sbb;
          const uint32_t kb_addr = (MODE == MODE_DQ) ? sab : st0b;
          const uint32_t vb_addr = (MODE == MODE_DQ) ? sbb : st1b;
          mbar_wait(&sf[s], (gg >> 1) & 1, 22);
          tc_fence_after();
          const uint32_t idesc_s = umma_idesc_bf16(AB_T, nkv, 0, 0);
#pragma unroll
          for (int k = 0; k < AB_HD / 16; ++k)   // S = Q K^T
            umma_bf16_ss(tmem_base + AB_TM_S, umma_desc_kmajor_sw128(q_addr + k * 32), umma_desc_kmajor_sw128(k_addr + k * 32),
                         idesc_s, k != 0);
          if (RB) umma_bf16_ss(tmem_base + AB_TM_S, ab_desc_sw32(qb_addr), ab_desc_sw32(kb_addr), idesc_s, 1);
#pragma unroll
          for (int k = 0; k < AB_HD / 16; ++k)   // dP = dO V^T
            umma_bf16_ss(tmem_base + AB_TM_DP, umma_desc_kmajor_sw128(do_addr + k * 32), umma_desc_kmajor_sw128(v_addr + k * 32),
                         idesc_s, k != 0);
          if (RB) umma_bf16_ss(tmem_base + AB_TM_DP, ab_desc_sw32(dob_addr), ab_desc_sw32(vb_addr), idesc_s, 1);
          umma_commit(s_full);
        };
        issue_s(0, g);   // S / dP are free: pds_ready of the previous item's last iteration was waited on below
        for (int it = 0; it < nt; ++it, ++g) {
          const int s = g & 1;
          const uint32_t sa = smem_u32(smem + AB_OFF_SA + s * AB_TILE), sb = smem_u32(smem + AB_OFF_SB + s * AB_TILE);
          const uint32_t sab = smem_u32(smem + AB_OFF_SAB + s * AB_BT), sbb = smem_u32(smem + AB_OFF_SBB + s * AB_BT);
          const int kv0 = (MODE == MODE_DQ) ? it * AB_T : t0;
          const int q0 = (MODE == MODE_DQ) ? t0 : it * AB_T;
          const int nkv = (min(AB_T, Lm - kv0) + 15) & ~15;
          const int nq = (min(AB_T, Lm - q0) + 15) & ~15;
          const uint32_t q_addr = (MODE == MODE_DQ) ? st0 : sa;
          const uint32_t do_addr = (MODE == MODE_DQ) ? st1 : sb;
          const uint32_t k_addr = (MODE == MODE_DQ) ? sa : st0;
          const uint32_t qb_addr = (MODE == MODE_DQ) ? st0b : sab;
          const uint32_t dob_addr = (MODE == MODE_DQ) ? st1b : sbb;
          const uint32_t kb_addr = (MODE == MODE_DQ) ? sab : st0b;
          mbar_wait(pds_ready, g & 1, 23);   // the softmax warps are done with S / dP(it) and have written P / dS(it)
          if (it + 1 < nt) issue_s(it + 1, g + 1);
          if (it == 0) mbar_wait(acc_free, (n & 1) ^ 1, 28);   // the previous item's epilogue has read the accumulators
          tc_fence_after();
          if (MODE == MODE_DQ) {
            // dQ += dS K_j : A = dS K-major (two 64-column atoms), B = K_j MN-major (rows = keys)
            constexpr uint32_t idesc = umma_idesc_bf16(AB_T, AB_HD, 0, 1);
            for (int kk = 0; kk < nkv / 16; ++kk)
              umma_bf16_ss(tmem_base + AB_TM_ACC0, umma_desc_kmajor_sw128(ds_addr + (kk >> 2) * AB_TILE + (kk & 3) * 32),
                           umma_desc_mnmajor_sw128(k_addr + kk * 2048, AB_TILE), idesc, (it | kk) != 0);
            if (RB) {
              constexpr uint32_t idescb = umma_idesc_bf16(AB_T, 16, 0, 1);
              for (int kk = 0; kk < nkv / 16; ++kk)
                umma_bf16_ss(tmem_base + AB_TM_ACC0B, umma_desc_kmajor_sw128(ds_addr + (kk >> 2) * AB_TILE + (kk & 3) * 32),
                             ab_desc_sw32(kb_addr + kk * 512), idescb, (it | kk) != 0);
            }
          } else {
            // dV += P^T dO_i, dK += dS^T Q_i : A = (P | dS)^T MN-major (M = keys: two 64-key panels), B MN-major (rows = queries)
            constexpr uint32_t idesc = umma_idesc_bf16(AB_T, AB_HD, 1, 1);
            for (int kk = 0; kk < nq / 16; ++kk)
              umma_bf16_ss(tmem_base + AB_TM_ACC1, umma_desc_mnmajor_sw128(p_addr + kk * 2048, AB_TILE),
                           umma_desc_mnmajor_sw128(do_addr + kk * 2048, AB_TILE), idesc, (it | kk) != 0);
            for (int kk = 0; kk < nq / 16; ++kk)
              umma_bf16_ss(tmem_base + AB_TM_ACC0, umma_desc_mnmajor_sw128(ds_addr + kk * 2048, AB_TILE),
                           umma_desc_mnmajor_sw128(q_addr + kk * 2048, AB_TILE), idesc, (it | kk) != 0);
            if (RB) {
              constexpr uint32_t idescb = umma_idesc_bf16(AB_T, 16, 1, 1);
              for (int kk = 0; kk < nq / 16; ++kk)
                umma_bf16_ss(tmem_base + AB_TM_ACC1B, umma_desc_mnmajor_sw128(p_addr + kk * 2048, AB_TILE),
                             ab_desc_sw32(dob_addr + kk * 512), idescb, (it | kk) != 0);
              for (int kk = 0; kk < nq / 16; ++kk)
                umma_bf16_ss(tmem_base + AB_TM_ACC0B, umma_desc_mnmajor_sw128(ds_addr + kk * 2048, AB_TILE),
                             ab_desc_sw32(qb_addr + kk * 512), idescb, (it | kk) != 0);
            }
          }
          if (MODE == MODE_FUSED) {
            // dQ_i (partial over this key tile) = dS K_j, fresh every iteration: the softmax warps drained the previous one
            // before they arrived on pds_ready (it > 0) / acc_free (it = 0)
            constexpr uint32_t idq = umma_idesc_bf16(AB_T, AB_HD, 0, 1);
            for (int kk = 0; kk < nkv / 16; ++kk)
              umma_bf16_ss(tmem_base + AB_TM_DQ, umma_desc_kmajor_sw128(ds_addr + (kk >> 2) * AB_TILE + (kk & 3) * 32),
                           umma_desc_mnmajor_sw128(k_addr + kk * 2048, AB_TILE), idq, kk != 0);
            if (RB) {
              constexpr uint32_t idqb = umma_idesc_bf16(AB_T, 16, 0, 1);
              for (int kk = 0; kk < nkv / 16; ++kk)
                umma_bf16_ss(tmem_base + AB_TM_DQB, umma_desc_kmajor_sw128(ds_addr + (kk >> 2) * AB_TILE + (kk & 3) * 32),
                             ab_desc_sw32(kb_addr + kk * 512), idqb, kk != 0);
            }
          }
          umma_commit(&se[s]);
          umma_commit(mma_done);
          if (it == nt - 1) umma_commit(&st_empty[buf]);   // the item's stationary tiles are no longer read by any MMA
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax / dS warps
    const int quad = warp & 3;
    const int hsel = warp >> 2;           // which 64 of the 128 key columns this thread handles
    const int r = quad * 32 + lane;       // query row inside the tile = TMEM lane
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    int n = 0, g = 0;
    for (int item = item_first; item < item_last; item += item_step, ++n) {
    const int t0 = (item % nt) * AB_T, h = (item / nt) % H, b = item / (nt * H);
    const int buf = n % NST;
    const long long bh = static_cast<long long>(b) * H + h;
    float lse2 = INFINITY, dlt = 0.f;
    // MODE_FUSED: the dQ partial of the query tile at q0 leaves TMEM: fp32 rows into this warp's own staging region, then
    // one TMA reduce-add per warp (32 rows x 32 columns) into the fp32 scratch.  No CTA-wide synchronisation: the region
    // is private to the warp, whose lane 0 issued (and waits for) the previous reduce.
    auto drain_dq = [&](int q0) {
      if (q0 + quad * 32 >= Lm) return;   // warp-uniform: nothing but clipped rows
      tc_fence_after();                   // behind the mma_done wait of the caller
      uint32_t o[32];
      tmem_ld_x32(tmem_base + t_lane + AB_TM_DQ + hsel * 32, o);
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      uint32_t region;
      if (RB) region = hsel == 0 ? AB_OFF_X + quad * AB_DQ_REGION : (quad == 0 ? AB_OFF_XB : AB_OFF_DQX + (quad - 1) * AB_DQ_REGION);
      else region = AB_OFF_DQST + hsel * AB_TILE + quad * AB_DQ_REGION;
      const uint32_t dst = smem_u32(smem + region);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        sts128(dst + sw128_offset(lane, q), make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
      if (RB && hsel == 0) {
        uint32_t ob[16];
        tmem_ld_x16(tmem_base + t_lane + AB_TM_DQB, ob);
        tmem_ld_wait();
        const uint32_t dstb = smem_u32(smem + AB_OFF_OUTB + quad * (32 * 64)) + lane * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q) sts128(dstb + q * 16, make_uint4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_4d(&tmDQF, smem + region, hsel * 32, h, q0 + quad * 32, b);
        if (RB && hsel == 0) tma_reduce_add_4d(&tmDQFb, smem + AB_OFF_OUTB + quad * (32 * 64), 64, h, q0 + quad * 32, b);
        tma_store_commit();
      }
    };
    // the previous item's output tiles (staged in the P region) must have left shared memory before P / dS are rewritten
    if (threadIdx.x == 0) tma_store_wait_read<0>();
    named_bar_sync(1, 32 * AB_SOFTMAX_WARPS);
    if (MODE == MODE_DQ) {
      // D_i = rowsum(dO_i . O_i) from the stationary tiles (each of the two threads of a row computes all 64 terms)
      mbar_wait(&st_full[buf], (n / NST) & 1, 24);
      const uint32_t o_s = smem_u32(smem + st_off(buf, 2)), do_s = smem_u32(smem + st_off(buf, 1));
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = lds128(o_s + sw128_offset(r, c));
        const uint4 g = lds128(do_s + sw128_offset(r, c));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) d = fmaf(bf16_lo(aw[q]), bf16_lo(gw[q]), fmaf(bf16_hi(aw[q]), bf16_hi(gw[q]), d));
      }
      if (RB) {
        const uint32_t ob_s = smem_u32(smem + AB_OFF_XB), dob_s = smem_u32(smem + AB_OFF_ST1B);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 a = lds128(ob_s + ab_sw32_offset(r, c));
          const uint4 g = lds128(dob_s + ab_sw32_offset(r, c));
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) d = fmaf(bf16_lo(aw[q]), bf16_lo(gw[q]), fmaf(bf16_hi(aw[q]), bf16_hi(gw[q]), d));
        }
      }
      dlt = d;
      mbar_arrive(&st_empty[buf]);   // this thread is done reading the stationary tiles
      const int row = t0 + r;
      if (row < Lm) {
        lse2 = lse[bh * L + row] * 1.4426950408889634f;
        if (hsel == 0) delta[bh * L + row] = d;
      }
    }
    for (int it = 0; it < nt; ++it, ++g) {
      const int kv0 = (MODE == MODE_DQ) ? it * AB_T : t0;
      const int valid_kv = min(AB_T, Lm - kv0);
      // causal mask (transformer.py:757-763): query row qrow sees keys <= qrow; P and dS are zero above the diagonal, which
      // is the same per-element masking as the columns past the end of the sequence, with a per-thread column count
      const int qrow = ((MODE == MODE_DQ) ? t0 : it * AB_T) + r;
      const int kmax = causal ? min(valid_kv, qrow - kv0 + 1) : valid_kv;
      if (MODE != MODE_DQ) {
        const int row = it * AB_T + r;
        lse2 = row < Lm ? lse[bh * L + row] * 1.4426950408889634f : INFINITY;
        dlt = row < Lm ? delta[bh * L + row] : 0.f;
      }
      mbar_wait(s_full, g & 1, 25);
      tc_fence_after();
      if (g > 0) mbar_wait(mma_done, (g - 1) & 1, 26);  // previous accumulating MMAs no longer read P / dS
      if (MODE == MODE_FUSED && it > 0) drain_dq((it - 1) * AB_T);
      // Which parts of the P / dS tile the accumulating MMAs read (the rest may hold anything: it only reaches output
      // rows that the TMA store clips).  DQ: A = dS[q rows][keys < nkv].  DKV: A = (P | dS)^T, contraction over the first
      // nq query rows, key columns >= valid_kv only feed clipped dK / dV rows.  With L = 128 k + 1 most of the remainder
      // tile's work falls away here.
      const int rows_read = (MODE == MODE_DQ) ? min(AB_T, Lm - t0) : ((min(AB_T, Lm - it * AB_T) + 15) & ~15);
      const int cols_read = (MODE == MODE_DQ) ? ((valid_kv + 15) & ~15) : valid_kv;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        const int col = hsel * 64 + c;
        if (col >= cols_read || quad * 32 >= rows_read) continue;   // warp-uniform
        uint32_t sv[32], dv[32];
        tmem_ld_x32(tmem_base + t_lane + AB_TM_S + col, sv);
        tmem_ld_x32(tmem_base + t_lane + AB_TM_DP + col, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
        if (col + 32 <= kmax) {   // full chunk (the common case): no per-element masking
          // packed fp32 pairs (FFMA2 / FADD2 / FMUL2): 3 instead of 6 FMA-pipe instructions per pair of scores
          const uint64_t sc2 = f2_pack(s2, s2), nl2 = f2_pack(-lse2, -lse2), nd2 = f2_pack(-dlt, -dlt);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x0, x1, d0, d1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[2 * j]), __uint_as_float(sv[2 * j + 1])), sc2, nl2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            pp[j] = pack_bf16x2(p0, p1);
            f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(dv[2 * j]), __uint_as_float(dv[2 * j + 1])), nd2)), d0, d1);
            dd[j] = pack_bf16x2(d0, d1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * j]), s2, -lse2));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * j + 1]), s2, -lse2));
            p0 = (col + 2 * j < kmax) ? p0 : 0.f;
            p1 = (col + 2 * j + 1 < kmax) ? p1 : 0.f;
            const float d0 = (col + 2 * j < kmax) ? p0 * (__uint_as_float(dv[2 * j]) - dlt) : 0.f;
            const float d1 = (col + 2 * j + 1 < kmax) ? p1 * (__uint_as_float(dv[2 * j + 1]) - dlt) : 0.f;
            pp[j] = pack_bf16x2(p0, p1);
            dd[j] = pack_bf16x2(d0, d1);
          }
        }
        const uint32_t chunk0 = c >> 3;  // 16-byte chunk inside the 64-column atom
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (MODE != MODE_DQ)
            sts128(smem_u32(smem + AB_OFF_P + hsel * AB_TILE) + sw128_offset(r, chunk0 + q),
                   make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]));
          sts128(smem_u32(smem + AB_OFF_DS + hsel * AB_TILE) + sw128_offset(r, chunk0 + q),
                 make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
    }
    // -------------------------------------------------------------------- epilogue: accumulators -> bf16 -> TMA store
    mbar_wait(mma_done, (g - 1) & 1, 27);
    tc_fence_after();
    if (MODE == MODE_FUSED) {
      drain_dq((nt - 1) * AB_T);
      if (RB) {   // the narrow output tiles below reuse OUTB: every warp's narrow reduce must have read it
        if (lane == 0) tma_store_wait_read<0>();
        named_bar_sync(1, 32 * AB_SOFTMAX_WARPS);
      }
    }
    // staging reuses the P region: atom 0 <- ACC0 (dQ or dK), atom 1 <- ACC1 (dV)
    // rank-1 terms of the remainder token (see the kernel header): coefficient of this thread's row, and the row of the
    // packed qkv / dout tensors that multiplies it
    const bool tail = ws != nullptr;
    float tcoef = 0.f;
    const __nv_bfloat16* tvec = nullptr;
    if (tail) {
      const float* wsb = ws + bh * 3 * Lm;
      const long long trow = static_cast<long long>(b) * L + Lm;   // the remainder token
      if (MODE == MODE_DQ) {
        tcoef = wsb[t0 + r];                                                                  // dS(i, t)
        tvec = qkv_g + (trow * 3 * H + H + h) * hd_g;                                          // k_t
      } else if (hsel == 0) {
        tcoef = wsb[2 * Lm + t0 + r];                                                         // dS(t, j)
        tvec = qkv_g + (trow * 3 * H + h) * hd_g;                                              // q_t
      } else {
        tcoef = wsb[Lm + t0 + r];                                                             // P(t, j)
        tvec = dout_g + (trow * H + h) * hd_g;                                                 // dO_t
      }
    }
    // o[0..n) += tcoef * tvec[col0 .. col0 + n)   (n a multiple of 8)
    auto add_tail = [&](uint32_t* o, int col0, int n) {
      for (int v = 0; v < n / 8; ++v) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(tvec + col0) + v);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          o[8 * v + 2 * q] = __float_as_uint(fmaf(tcoef, bf16_lo(w4[q]), __uint_as_float(o[8 * v + 2 * q])));
          o[8 * v + 2 * q + 1] = __float_as_uint(fmaf(tcoef, bf16_hi(w4[q]), __uint_as_float(o[8 * v + 2 * q + 1])));
        }
      }
    };
    if (MODE == MODE_DQ) {
      uint32_t o[32];
      tmem_ld_x32(tmem_base + t_lane + AB_TM_ACC0 + hsel * 32, o);
      tmem_ld_wait();
      if (tail) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(tvec + hsel * 32) + v);
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[8 * v + 2 * q] = __float_as_uint(fmaf(tcoef, bf16_lo(w4[q]), __uint_as_float(o[8 * v + 2 * q])));
            o[8 * v + 2 * q + 1] = __float_as_uint(fmaf(tcoef, bf16_hi(w4[q]), __uint_as_float(o[8 * v + 2 * q + 1])));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sts128(smem_u32(smem + AB_OFF_P) + sw128_offset(r, hsel * 4 + q),
               make_uint4(pack_bf16x2(__uint_as_float(o[8 * q]) * scale, __uint_as_float(o[8 * q + 1]) * scale),
                          pack_bf16x2(__uint_as_float(o[8 * q + 2]) * scale, __uint_as_float(o[8 * q + 3]) * scale),
                          pack_bf16x2(__uint_as_float(o[8 * q + 4]) * scale, __uint_as_float(o[8 * q + 5]) * scale),
                          pack_bf16x2(__uint_as_float(o[8 * q + 6]) * scale, __uint_as_float(o[8 * q + 7]) * scale)));
    } else {
      const float mul = hsel == 0 ? scale : 1.f;   // hsel 0: dK (scaled), hsel 1: dV
      const uint32_t src = hsel == 0 ? AB_TM_ACC0 : AB_TM_ACC1;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(tmem_base + t_lane + src + c, o);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(tvec + c) + v);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              o[8 * v + 2 * q] = __float_as_uint(fmaf(tcoef, bf16_lo(w4[q]), __uint_as_float(o[8 * v + 2 * q])));
              o[8 * v + 2 * q + 1] = __float_as_uint(fmaf(tcoef, bf16_hi(w4[q]), __uint_as_float(o[8 * v + 2 * q + 1])));
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(smem_u32(smem + AB_OFF_P + hsel * AB_TILE) + sw128_offset(r, (c >> 3) + q),
                 make_uint4(pack_bf16x2(__uint_as_float(o[8 * q]) * mul, __uint_as_float(o[8 * q + 1]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * q + 2]) * mul, __uint_as_float(o[8 * q + 3]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * q + 4]) * mul, __uint_as_float(o[8 * q + 5]) * mul),
                            pack_bf16x2(__uint_as_float(o[8 * q + 6]) * mul, __uint_as_float(o[8 * q + 7]) * mul)));
      }
    }
    if (RB && (MODE != MODE_DQ || hsel == 0)) {
      // narrow accumulators: DQ: dQ_b (hsel 0); DKV: dK_b (hsel 0, scaled), dV_b (hsel 1)
      const float mulb = (MODE == MODE_DQ || hsel == 0) ? scale : 1.f;
      const uint32_t srcb = (MODE == MODE_DQ || hsel == 0) ? AB_TM_ACC0B : AB_TM_ACC1B;
      uint32_t ob[16];
      tmem_ld_x16(tmem_base + t_lane + srcb, ob);
      tmem_ld_wait();
      if (tail) add_tail(ob, 64, hd_g - 64 < 16 ? hd_g - 64 : 16);   // dims 64 .. hd (the columns past hd stay zero)
      const uint32_t dst = smem_u32(smem + AB_OFF_OUTB + hsel * AB_BT);
#pragma unroll
      for (int c = 0; c < 2; ++c)
        sts128(dst + ab_sw32_offset(r, c),
               make_uint4(pack_bf16x2(__uint_as_float(ob[8 * c]) * mulb, __uint_as_float(ob[8 * c + 1]) * mulb),
                          pack_bf16x2(__uint_as_float(ob[8 * c + 2]) * mulb, __uint_as_float(ob[8 * c + 3]) * mulb),
                          pack_bf16x2(__uint_as_float(ob[8 * c + 4]) * mulb, __uint_as_float(ob[8 * c + 5]) * mulb),
                          pack_bf16x2(__uint_as_float(ob[8 * c + 6]) * mulb, __uint_as_float(ob[8 * c + 7]) * mulb)));
    }
    tc_fence_before();
    mbar_arrive(acc_free);   // accumulators read: the next item's first accumulating MMA may overwrite them
    fence_proxy_async_smem();
    named_bar_sync(1, 32 * AB_SOFTMAX_WARPS);
    if (threadIdx.x == 0) {
      if (MODE == MODE_DQ) {
        tma_store_4d(&tmDQKV, smem + AB_OFF_P, 0, h, t0, b);
        if (RB) tma_store_4d(&tmDQKVb, smem + AB_OFF_OUTB, 64, h, t0, b);
      } else {
        tma_store_4d(&tmDQKV, smem + AB_OFF_P, 0, H + h, t0, b);
        tma_store_4d(&tmDQKV, smem + AB_OFF_P + AB_TILE, 0, 2 * H + h, t0, b);
        if (RB) {
          tma_store_4d(&tmDQKVb, smem + AB_OFF_OUTB, 64, H + h, t0, b);
          tma_store_4d(&tmDQKVb, smem + AB_OFF_OUTB + AB_BT, 64, 2 * H + h, t0, b);
        }
      }
      tma_store_commit();
    }
    }   // items
    if (threadIdx.x == 0 || (MODE == MODE_FUSED && lane == 0)) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AB_SOFTMAX_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<AB_TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ remainder token
// L = 128 k + 1 (class token + a power-of-two patch grid: every L/14 and H/14 training step): as a third row / column of
// 128-wide tiles the one remainder token t = L - 1 cost the tile kernels +63 % (B = 256, L = 257 vs 256).  Here it is a
// handful of dot products per (image, head) on the FMA pipe, one thread per token:
//   row t      : s_j = q_t . k_j, P(t, j), dS(t, j) for every key j  ->  dq_t = scale * sum_j dS(t, j) k_j
//   column t   : s_i = q_i . k_t, P(i, t), dS(i, t) for every query i ->  dk_t = scale * sum_i dS(i, t) q_i,  dv_t = sum_i P(i, t) dO_i
// and the vectors dS(., t), P(t, .), dS(t, .) over the tile tokens go to `ws` for the tile kernels' epilogues (rank-1 terms).
// One pass over q, k, v, dO, O of the (image, head): a row is read ONCE, 16 bytes per lane, LPR lanes per row (coalesced
// 128 - 160 byte segments); the five dot products of token j (as key of row t, as query of column t) are reduced over the
// row's lanes by shuffles, and the three weighted row sums accumulate in registers (8 dims per lane), reduced over rows at the
// end in a fixed order.  HBM-bound: 5 * B * L * H * hd * 2 bytes.  (The first version walked the rows once per dot product and
// once per weighted sum, one thread per row: 9 row passes with 16-byte accesses 7 680 bytes apart, 1.4 ms at batch 1024 x 16
// heads x 80 dims where the traffic bound is 0.5 ms.)
constexpr int ABT_THREADS = 256;
constexpr int ABT_WARPS = ABT_THREADS / 32;
constexpr int ABT_STAGES = 3;
constexpr int ABT_STAGE_BYTES = ABT_STAGES * ABT_THREADS * 80;   // 60 KB: three CTAs per SM

// 8 bf16 of a row times 8 fp32 of one of the remainder token's vectors (kept in shared memory: the four vectors as registers
// cost 32 of them and a third resident CTA per SM)
__device__ __forceinline__ float abt_dot8(const uint4& u, const float* v) {   // v in shared memory
  const float4 a = *reinterpret_cast<const float4*>(v), b = *reinterpret_cast<const float4*>(v + 4);
  return fmaf(bf16_lo(u.x), a.x, fmaf(bf16_hi(u.x), a.y, fmaf(bf16_lo(u.y), a.z, fmaf(bf16_hi(u.y), a.w,
         fmaf(bf16_lo(u.z), b.x, fmaf(bf16_hi(u.z), b.y, fmaf(bf16_lo(u.w), b.z, bf16_hi(u.w) * b.w)))))));
}
__device__ __forceinline__ float abt_dot8(const uint4& u, const float (&v)[8]) {
  return fmaf(bf16_lo(u.x), v[0], fmaf(bf16_hi(u.x), v[1], fmaf(bf16_lo(u.y), v[2], fmaf(bf16_hi(u.y), v[3],
         fmaf(bf16_lo(u.z), v[4], fmaf(bf16_hi(u.z), v[5], fmaf(bf16_lo(u.w), v[6], bf16_hi(u.w) * v[7])))))));
}
__device__ __forceinline__ float abt_dot8b(const uint4& u, const uint4& w) {
  return fmaf(bf16_lo(u.x), bf16_lo(w.x), fmaf(bf16_hi(u.x), bf16_hi(w.x), fmaf(bf16_lo(u.y), bf16_lo(w.y), fmaf(bf16_hi(u.y), bf16_hi(w.y),
         fmaf(bf16_lo(u.z), bf16_lo(w.z), fmaf(bf16_hi(u.z), bf16_hi(w.z), fmaf(bf16_lo(u.w), bf16_lo(w.w), bf16_hi(u.w) * bf16_hi(w.w))))))));
}
__device__ __forceinline__ void abt_axpy8(float (&a)[8], float c, const uint4& u) {
  a[0] = fmaf(c, bf16_lo(u.x), a[0]); a[1] = fmaf(c, bf16_hi(u.x), a[1]);
  a[2] = fmaf(c, bf16_lo(u.y), a[2]); a[3] = fmaf(c, bf16_hi(u.y), a[3]);
  a[4] = fmaf(c, bf16_lo(u.z), a[4]); a[5] = fmaf(c, bf16_hi(u.z), a[5]);
  a[6] = fmaf(c, bf16_lo(u.w), a[6]); a[7] = fmaf(c, bf16_hi(u.w), a[7]);
}

// Sum of v over the LPR lanes of a row group, delivered to all of them.  LPR = 8: xor butterfly.  LPR = 10 (head widths 72 / 80:
// ten 16-byte chunks per row, three rows per warp, lanes 30 / 31 idle): fold the upper five lanes onto the lower five, add those
// five up in lane 0 of the group, broadcast.  (The kernel is bound by its instruction stream: with 16 lanes per row, ten of them
// loading, every row cost 16 lane-passes and the wide-head kernel ran at 49 % issue utilisation and 32 % of the HBM rate.)
template <int LPR>
__device__ __forceinline__ float abt_group_sum(float v, int base) {
  if (LPR == 8) {
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  } else {
    v += __shfl_down_sync(0xffffffffu, v, 5);           // lanes 0..4 of the group: a_i + a_(i+5)
    float t = v + __shfl_down_sync(0xffffffffu, v, 1);  // lane 0: c0 + c1, lane 2: c2 + c3
    t += __shfl_down_sync(0xffffffffu, t, 2);           // lane 0: c0 + c1 + c2 + c3
    t += __shfl_down_sync(0xffffffffu, v, 4);           // lane 0: + c4
    return __shfl_sync(0xffffffffu, t, base);
  }
}

// LPR lanes per row: 8 (hd = 64) or 10 (hd = 72 / 80; three rows per warp).  PIPE: rows through the cp.async pipeline and the
// token's vectors in registers (two CTAs per SM) instead of plain loads and vectors in shared memory (three CTAs).  Measured
// (ncu, batch 1024 x 16 heads x 257 tokens): hd 64: plain 0.75 ms, pipeline 0.84-0.86; hd 80: plain 1.38, pipeline 1.08.
template <int LPR, bool PIPE>
__global__ void __launch_bounds__(ABT_THREADS, PIPE ? 2 : 3)
attention_bwd_tail_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                          const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv,
                          float* __restrict__ ws, int L, int H, int hd, float scale, float* __restrict__ delta_out,
                          float* __restrict__ stats, int Lp) {
  // delta_out / stats (optional): this kernel computes delta_j = dO_j . O_j of every token anyway, so for the one-pass tile
  // kernel it also leaves what attention_bwd_delta_kernel would (delta [B, H, L]; stats [B * H][2][Lp], Lp = L - 1)
  constexpr int RPW = 32 / LPR;               // rows per warp and trip
  __shared__ __align__(16) float vec[5][LPR * 8];   // q_t, k_t, v_t, dO_t, O_t (zero past hd)
  __shared__ float red[ABT_WARPS][3][LPR * 8];
  const int Lm = L - 1, t = L - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane / LPR;                 // LPR = 10: lanes 30, 31 form an idle fourth group
  const int sub = lane - grp * LPR;
  const bool active = grp < RPW && sub * 8 < hd;
  const int h = blockIdx.x;
  const long long b = blockIdx.y;
  const long long bh = b * H + h;
  const long long qtok = 3LL * H * hd;                 // elements per token of qkv / dqkv
  const long long otok = static_cast<long long>(H) * hd;
  const __nv_bfloat16* qb = qkv + b * L * qtok + h * hd;                 // q rows of this head (k: + H*hd, v: + 2*H*hd)
  const __nv_bfloat16* ob = out + b * L * otok + h * hd;
  const __nv_bfloat16* db = dout + b * L * otok + h * hd;
  const float s2 = scale * 1.4426950408889634f;
  __shared__ float part[ABT_WARPS];
  float pr = 0.f;   // this thread's share of delta_t = dO_t . O_t
  for (int d = tid; d < LPR * 8; d += ABT_THREADS) {
    const bool in = d < hd;
    const float dv = in ? __bfloat162float(db[t * otok + d]) : 0.f;
    const float ov = in ? __bfloat162float(ob[t * otok + d]) : 0.f;
    vec[0][d] = in ? __bfloat162float(qb[t * qtok + d]) : 0.f;
    vec[1][d] = in ? __bfloat162float(qb[t * qtok + H * hd + d]) : 0.f;
    vec[2][d] = in ? __bfloat162float(qb[t * qtok + 2 * H * hd + d]) : 0.f;
    vec[3][d] = dv;
    vec[4][d] = ov;
    pr = fmaf(dv, ov, pr);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, off);
  if (lane == 0) part[warp] = pr;
  __syncthreads();
  float delta_t = 0.f;   // (every thread walking the two vectors itself: 2 x hd shared-memory loads per thread and CTA)
#pragma unroll
  for (int w = 0; w < ABT_WARPS; ++w) delta_t += part[w];
  const float lse_t2 = lse[bh * L + t] * 1.4426950408889634f;
  // the token's own vectors in registers: re-reading them from shared memory for every row (8 LDS.128 per trip) kept the
  // L1 / shared-memory pipe, which also carries the cp.async stream below, at 75 %
  float qt[8], kt[8], vt[8], dt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    qt[e] = vec[0][sub * 8 + e];
    kt[e] = vec[1][sub * 8 + e];
    vt[e] = vec[2][sub * 8 + e];
    dt[e] = vec[3][sub * 8 + e];
  }
  const float *qs = &vec[0][sub * 8], *ks = &vec[1][sub * 8], *vs = &vec[2][sub * 8], *ds = &vec[3][sub * 8];
  float* wsb = ws + bh * 3 * Lm;
  float aq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // dq_t = scale * sum_j dS(t, j) k_j
  float ak[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // dk_t = scale * sum_i dS(i, t) q_i
  float av[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // dv_t = sum_i P(i, t) dO_i
  // PIPE: the rows come in through a three-deep cp.async pipeline, each lane fetching (and later reading back) its own five
  // 16-byte chunks (wide heads: neither fewer instructions per row nor contiguous rows nor a cheaper prologue moved the
  // plain-load kernel off 1.38 ms = 2.4 TB/s; two more trips in flight did).
  extern __shared__ __align__(16) uint8_t abt_stage[];   // [ABT_STAGES][ABT_THREADS][5] x 16 bytes, lane-private records
  const int ntrips = (L + ABT_WARPS * RPW - 1) / (ABT_WARPS * RPW);   // the same for every warp (rows past L are zero-filled)
  auto issue = [&](int trip) {
    if (trip < ntrips) {
      const int jj = trip * ABT_WARPS * RPW + warp * RPW + grp;
      const bool ldd = jj < L && active;
      const uint32_t nb = ldd ? 16u : 0u;
      const long long jr = ldd ? jj : 0;
      const uint32_t dst = smem_u32(abt_stage) + ((trip % ABT_STAGES) * ABT_THREADS + tid) * 80;
      cp_async_16(dst, reinterpret_cast<const uint4*>(qb + jr * qtok) + sub, nb);
      cp_async_16(dst + 16, reinterpret_cast<const uint4*>(qb + jr * qtok + H * hd) + sub, nb);
      cp_async_16(dst + 32, reinterpret_cast<const uint4*>(qb + jr * qtok + 2 * H * hd) + sub, nb);
      cp_async_16(dst + 48, reinterpret_cast<const uint4*>(db + jr * otok) + sub, nb);
      cp_async_16(dst + 64, reinterpret_cast<const uint4*>(ob + jr * otok) + sub, nb);
    }
    cp_async_commit();   // (an empty group past the last trip keeps the wait count uniform)
  };
  auto lse_of = [&](int trip) {
    const int jj = trip * ABT_WARPS * RPW + warp * RPW + grp;
    return (trip < ntrips && jj < L && grp < RPW) ? __ldg(lse + bh * L + jj) * 1.4426950408889634f : 0.f;
  };
  if (PIPE) {
    issue(0);
    issue(1);
  }
  float lse_a = lse_of(0), lse_b = lse_of(1);
#pragma unroll 1
  for (int trip = 0; trip < ntrips; ++trip) {
    if (PIPE) issue(trip + 2);
    const float lse_j2 = lse_a;
    lse_a = lse_b;
    lse_b = lse_of(trip + 2);
    const int j = trip * ABT_WARPS * RPW + warp * RPW + grp;
    const bool ok = j < L && grp < RPW;
    uint4 qq, kk, vv, dd, oo;
    if (PIPE) {
      cp_async_wait_group<2>();   // this trip's five chunks have landed (the two younger groups may still be in flight)
      const uint32_t src = smem_u32(abt_stage) + ((trip % ABT_STAGES) * ABT_THREADS + tid) * 80;
      qq = lds128(src), kk = lds128(src + 16), vv = lds128(src + 32), dd = lds128(src + 48), oo = lds128(src + 64);
    } else {
      const bool ld = ok && active;
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      qq = ld ? __ldg(reinterpret_cast<const uint4*>(qb + j * qtok) + sub) : z4;
      kk = ld ? __ldg(reinterpret_cast<const uint4*>(qb + j * qtok + H * hd) + sub) : z4;
      vv = ld ? __ldg(reinterpret_cast<const uint4*>(qb + j * qtok + 2 * H * hd) + sub) : z4;
      dd = ld ? __ldg(reinterpret_cast<const uint4*>(db + j * otok) + sub) : z4;
      oo = ld ? __ldg(reinterpret_cast<const uint4*>(ob + j * otok) + sub) : z4;
    }
    float s1 = PIPE ? abt_dot8(kk, qt) : abt_dot8(kk, qs);     // q_t . k_j      token j as a key of row t
    float dp1 = PIPE ? abt_dot8(vv, dt) : abt_dot8(vv, ds);    // dO_t . v_j
    float s2q = PIPE ? abt_dot8(qq, kt) : abt_dot8(qq, ks);    // q_j . k_t      token j as a query of column t
    float dp2 = PIPE ? abt_dot8(dd, vt) : abt_dot8(dd, vs);    // dO_j . v_t
    float dl = abt_dot8b(dd, oo);    // delta_j = dO_j . O_j
    s1 = abt_group_sum<LPR>(s1, grp * LPR);
    dp1 = abt_group_sum<LPR>(dp1, grp * LPR);
    s2q = abt_group_sum<LPR>(s2q, grp * LPR);
    dp2 = abt_group_sum<LPR>(dp2, grp * LPR);
    dl = abt_group_sum<LPR>(dl, grp * LPR);
    const float p1 = ok ? fast_exp2(fmaf(s1, s2, -lse_t2)) : 0.f;   // P(t, j)
    const float ds1 = p1 * (dp1 - delta_t);                          // dS(t, j)
    const float p2 = ok ? fast_exp2(fmaf(s2q, s2, -lse_j2)) : 0.f;  // P(j, t)
    const float ds2 = p2 * (dp2 - dl);                               // dS(j, t)
    abt_axpy8(aq, ds1, kk);
    abt_axpy8(ak, ds2, qq);
    abt_axpy8(av, p2, dd);
    if (sub == 0 && ok && delta_out != nullptr) {
      delta_out[bh * L + j] = dl;
      if (stats != nullptr && j < Lp) {
        stats[(bh * 2 + 0) * Lp + j] = -lse_j2;
        stats[(bh * 2 + 1) * Lp + j] = -dl;
      }
    }
    if (sub == 0 && ok && j < Lm) {
      wsb[j] = ds2;              // dS(i, t)  -> dQ_i += scale dS(i, t) k_t
      wsb[Lm + j] = p1;          // P(t, j)   -> dV_j += P(t, j) dO_t
      wsb[2 * Lm + j] = ds1;     // dS(t, j)  -> dK_j += scale dS(t, j) q_t
    }
  }
  // rows of the warp (lanes with the same dims), then warps through shared memory, both in a fixed order: deterministic
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (LPR == 8) {
#pragma unroll
      for (int off = 8; off < 32; off <<= 1) {
        aq[e] += __shfl_xor_sync(0xffffffffu, aq[e], off);
        ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], off);
        av[e] += __shfl_xor_sync(0xffffffffu, av[e], off);
      }
    } else {   // groups at lanes 0, 10, 20 (the accumulators of lanes 30, 31 are zero and are never read)
      aq[e] += __shfl_down_sync(0xffffffffu, aq[e], 10) + __shfl_down_sync(0xffffffffu, aq[e], 20);
      ak[e] += __shfl_down_sync(0xffffffffu, ak[e], 10) + __shfl_down_sync(0xffffffffu, ak[e], 20);
      av[e] += __shfl_down_sync(0xffffffffu, av[e], 10) + __shfl_down_sync(0xffffffffu, av[e], 20);
    }
  }
  if (grp == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][sub * 8 + e] = aq[e];
      red[warp][1][sub * 8 + e] = ak[e];
      red[warp][2][sub * 8 + e] = av[e];
    }
  }
  __syncthreads();
  for (int x = tid; x < 3 * hd; x += ABT_THREADS) {
    const int which = x / hd, d = x - which * hd;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < ABT_WARPS; ++w) sum += red[w][which][d];
    dqkv[(b * L + t) * qtok + which * H * hd + h * hd + d] = __float2bfloat16(which == 2 ? sum : sum * scale);
  }
}

// ------------------------------------------------------------------------------------------------ MODE_FUSED helpers
// delta[b, h, l] = sum_d dO[b, l, h, d] * O[b, l, h, d]: eight lanes per (token, head), one 16-byte vector each (two for the
// dims past 64), fully coalesced; HBM-bound (reads 2 * B * L * H * hd * 2 bytes).
__global__ void __launch_bounds__(256)
attention_bwd_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                           float* __restrict__ delta, long long rows, int L, int H, int hd, const float* __restrict__ lse,
                           float* __restrict__ stats, int Lp) {
  // stats (optional): f32 [B * H][2][Lp], what attention_bwd_t_kernel streams per query tile: [0] = -lse * log2(e), [1] = -delta,
  // tokens l < min(L, Lp); l in [L, Lp) padded with (-inf, 0) (a query past the end then has P = dS = 0)
  const int sub = threadIdx.x & 7;
  const int nv = hd >> 3;
  // four (token, head) rows per warp and trip; the trip count is warp-uniform (full-mask shuffles below)
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long row0 = warp0 * 4; row0 < rows; row0 += nwarps * 4) {
    const long long row = row0 + ((threadIdx.x & 31) >> 3);   // row = (b * L + l) * H + h
    const bool valid = row < rows;
    const uint4* po = reinterpret_cast<const uint4*>(out + row * hd);
    const uint4* pd = reinterpret_cast<const uint4*>(dout + row * hd);
    float a = 0.f;
    for (int v = sub; valid && v < nv; v += 8) {
      const uint4 x = __ldg(po + v), y = __ldg(pd + v);
      const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) a = fmaf(bf16_lo(xw[q]), bf16_lo(yw[q]), fmaf(bf16_hi(xw[q]), bf16_hi(yw[q]), a));
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    if (valid && sub == 0) {
      const long long tok = row / H;
      const int h = static_cast<int>(row - tok * H);
      const long long b = tok / L;
      const int l = static_cast<int>(tok - b * L);
      delta[(b * H + h) * L + l] = a;
      if (stats != nullptr && l < Lp) {
        stats[((b * H + h) * 2 + 0) * Lp + l] = -lse[(b * H + h) * L + l] * 1.4426950408889634f;
        stats[((b * H + h) * 2 + 1) * Lp + l] = -a;
      }
    }
  }
  if (stats != nullptr && Lp > L) {
    const int padn = Lp - L;
    const long long total = rows / L * padn;   // B * H * padn
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long bh = i / padn;
      const int l = L + static_cast<int>(i - bh * padn);
      stats[(bh * 2 + 0) * Lp + l] = -INFINITY;
      stats[(bh * 2 + 1) * Lp + l] = 0.f;
    }
  }
}

// d(qkv)[b, l, q slot, h, :] = bf16(scale * (acc[b, l, h, :] + dS(l, t) * k_t)) for the tokens l < Lm the tiles cover; the
// second term only when the remainder token t = L - 1 is handled outside the tiles (ws: see attention_bwd_tail_kernel).
// One thread per 8 elements: two 16-byte loads, one 16-byte store.
__global__ void __launch_bounds__(256)
attention_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv,
                                const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ ws, long long total8,
                                int L, int Lm, int H, int hd, float scale) {
  const int nv = hd >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nv);
    long long rest = i / nv;
    const int h = static_cast<int>(rest % H);
    rest /= H;
    const int l = static_cast<int>(rest % Lm);
    const long long b = rest / Lm;
    const float4* src = reinterpret_cast<const float4*>(acc + (((b * L + l) * H + h) * hd + v * 8));
    float4 x0 = __ldg(src), x1 = __ldg(src + 1);
    if (ws != nullptr) {
      const float c = ws[(b * H + h) * 3LL * Lm + l];                                             // dS(l, t)
      const uint4 k = __ldg(reinterpret_cast<const uint4*>(qkv + ((b * L + Lm) * 3LL * H + H + h) * hd + v * 8));   // k_t
      x0.x = fmaf(c, bf16_lo(k.x), x0.x); x0.y = fmaf(c, bf16_hi(k.x), x0.y);
      x0.z = fmaf(c, bf16_lo(k.y), x0.z); x0.w = fmaf(c, bf16_hi(k.y), x0.w);
      x1.x = fmaf(c, bf16_lo(k.z), x1.x); x1.y = fmaf(c, bf16_hi(k.z), x1.y);
      x1.z = fmaf(c, bf16_lo(k.w), x1.z); x1.w = fmaf(c, bf16_hi(k.w), x1.w);
    }
    const uint4 o = make_uint4(pack_bf16x2(x0.x * scale, x0.y * scale), pack_bf16x2(x0.z * scale, x0.w * scale),
                               pack_bf16x2(x1.x * scale, x1.y * scale), pack_bf16x2(x1.z * scale, x1.w * scale));
    *reinterpret_cast<uint4*>(dqkv + ((b * L + l) * 3LL * H + h) * hd + v * 8) = o;
  }
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                 float* delta, int B, int L, int H, int hd, float scale, void* stream) {
  return ovk_attention_bwd_ex(qkv, out, dout, lse, dqkv, delta, nullptr, B, L, H, hd, scale, 0, stream);
}

extern "C" long long ovk_attention_bwd_workspace_floats(int B, int L, int H, int flags) {
  // measured (tools/attn_bwd_ab.py, batch 1024 x 16 heads): L = 257: -4.5 % (head width 64), -8.4 % (80); L = 1025: +5 % (the
  // tail kernel's extra pass over q, k, v, dO, O outweighs one tile row / column in nine): offered up to three full tiles
  if (B <= 0 || L <= AB_T || H <= 0 || (L % AB_T) != 1 || L > 3 * AB_T + 1 || (flags & OVK_ATT_CAUSAL)) return 0;
  return 3LL * B * H * (L - 1);
}

extern "C" long long ovk_attention_bwd_fused_workspace_floats(int B, int L, int H, int hd, int flags) {
  if (B <= 0 || L <= 0 || H <= 0 || hd < 64 || hd > 80 || (hd % 8)) return 0;
  // [remainder-token vectors] [fp32 dQ accumulator B x L x H x hd] [per-query statistics B x H x 2 x Lp]
  return ovk_attention_bwd_workspace_floats(B, L, H, flags) + static_cast<long long>(B) * L * H * hd +
         2LL * B * H * ((L + AB_T - 1) / AB_T * AB_T);
}

static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                              float* delta, float* workspace, int B, int L, int H, int hd, float scale, int flags,
                              void* stream, bool fused);
namespace ovk {
int launch_attention_bwd_t(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const CUtensorMap& tmDQKV,
                           const CUtensorMap& tmQKVb, const CUtensorMap& tmDOb, const CUtensorMap& tmDQKVb, float* acc, int B,
                           const float* stats, int Lp, int L, int H, float scale, int items, int grid, int causal, int Lm,
                           const float* ws, const __nv_bfloat16* qkv_g, const __nv_bfloat16* dout_g, int hd, bool warps16,
                           cudaStream_t s);
}

extern "C" int ovk_attention_bwd_ex(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                    float* delta, float* workspace, int B, int L, int H, int hd, float scale, int flags,
                                    void* stream) {
  return attention_bwd_impl(qkv, out, dout, lse, dqkv, delta, workspace, B, L, H, hd, scale, flags, stream, false);
}

extern "C" int ovk_attention_bwd_fused(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                       float* delta, float* workspace, int B, int L, int H, int hd, float scale, int flags,
                                       void* stream) {
  if (!workspace) return set_error(OVK_ERR_SHAPE, "attention_bwd_fused: the workspace is required");
  return attention_bwd_impl(qkv, out, dout, lse, dqkv, delta, workspace, B, L, H, hd, scale, flags, stream, true);
}

static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                              float* delta, float* workspace, int B, int L, int H, int hd, float scale, int flags,
                              void* stream, bool fused) {
  if (B <= 0 || L <= 0 || H <= 0) return set_error(OVK_ERR_SHAPE, "attention_bwd: empty problem");
  if (flags & ~(OVK_ATT_CAUSAL | (fused ? (OVK_ATT_BWD_ONEPASS_V1 | OVK_ATT_BWD_8_WARPS | 0xff00) : 0)))
    return set_error(OVK_ERR_SHAPE, "attention_bwd: unknown flags 0x%x", flags);
  const int causal = (flags & OVK_ATT_CAUSAL) ? 1 : 0;
  if (hd < 64 || hd > 80 || (hd % 8))
    return set_error(OVK_ERR_SHAPE, "attention_bwd: head dim %d not supported (64, 72 or 80)", hd);
  const bool ext = hd > AB_HD;
  if (!lse || !delta) return set_error(OVK_ERR_SHAPE, "attention_bwd: lse and delta buffers are required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmQKV, tmDQKV, tmO, tmDO, tmQKVb, tmDQKVb, tmOb, tmDOb;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * H), (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)3 * H * hd * 2, (uint64_t)L * 3 * H * hd * 2};
    const uint32_t box[4] = {AB_HD, 1, AB_T, 1};
    const uint32_t boxb[4] = {16, 1, AB_T, 1};
    if ((rc = make_tmap_nd_bf16(&tmQKV, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmDQKV, dqkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmQKVb, qkv, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmDQKVb, dqkv, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hd * 2, (uint64_t)H * hd * 2, (uint64_t)L * H * hd * 2};
    const uint32_t box[4] = {AB_HD, 1, AB_T, 1};
    const uint32_t boxb[4] = {16, 1, AB_T, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmDO, dout, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmOb, out, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_tmap_nd_bf16(&tmDOb, dout, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_DQ, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_DKV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_DQ, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES_RB);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_DKV, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES_RB);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_FUSED, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<MODE_FUSED, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES_RB_FUSED);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention_bwd): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  // remainder token (L = 128 k + 1) outside the tiles when the caller supplied the workspace
  const bool tail = workspace != nullptr && ovk_attention_bwd_workspace_floats(B, L, H, flags) > 0 && B <= 65535 && H <= 65535;
  const int Lm = tail ? L - 1 : L;
  // one-pass path: fp32 dQ accumulator and per-query statistics behind the remainder-token vectors
  const bool v2 = fused && !(flags & OVK_ATT_BWD_ONEPASS_V1);
  float* acc = fused ? workspace + ovk_attention_bwd_workspace_floats(B, L, H, flags) : nullptr;
  float* stats = v2 ? acc + static_cast<long long>(B) * L * H * hd : nullptr;
  const int Lp = (Lm + AB_T - 1) / AB_T * AB_T;
  if (tail) {
    static PerDeviceOnce tail_once;
    if (tail_once.need()) {
      cudaError_t e = cudaFuncSetAttribute(attention_bwd_tail_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ABT_STAGE_BYTES);
      if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(attention_bwd_tail): %s", cudaGetErrorString(e));
      tail_once.done();
    }
    auto q_ = reinterpret_cast<const __nv_bfloat16*>(qkv);
    auto o_ = reinterpret_cast<const __nv_bfloat16*>(out);
    auto d_ = reinterpret_cast<const __nv_bfloat16*>(dout);
    float* dl_ = fused ? delta : nullptr;   // the two-pass dQ kernel computes delta itself
    if (hd == 64)
      attention_bwd_tail_kernel<8, false><<<dim3(H, B), ABT_THREADS, 0, s>>>(q_, o_, d_, lse, reinterpret_cast<__nv_bfloat16*>(dqkv), workspace, L, H, hd,
                                                                    scale, dl_, stats, Lp);
    else
      attention_bwd_tail_kernel<10, true><<<dim3(H, B), ABT_THREADS, ABT_STAGE_BYTES, s>>>(q_, o_, d_, lse, reinterpret_cast<__nv_bfloat16*>(dqkv), workspace, L, H, hd,
                                                                     scale, dl_, stats, Lp);
    if ((rc = check_launch("attention_bwd_tail_kernel"))) return rc;
  }
  const float* wsp = tail ? workspace : nullptr;
  auto qg = reinterpret_cast<const __nv_bfloat16*>(qkv);
  auto dg = reinterpret_cast<const __nv_bfloat16*>(dout);
  const long long items_ll = static_cast<long long>((Lm + AB_T - 1) / AB_T) * H * B;
  if (items_ll > 0x7fffffffLL) return set_error(OVK_ERR_SHAPE, "attention_bwd: too many work items");
  const int items = static_cast<int>(items_ll);
  const int grid = items < num_sms() ? items : num_sms();   // persistent: one CTA per SM (512 TMEM columns each)
  if (fused) {
    const long long nacc = static_cast<long long>(B) * L * H * hd;
    CUtensorMap tmDQF, tmDQFb;
    {
      const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)hd * 4, (uint64_t)H * hd * 4, (uint64_t)L * H * hd * 4};
      const uint32_t box[4] = {32, 1, 32, 1};
      const uint32_t boxb[4] = {16, 1, 32, 1};
      if ((rc = make_tmap_nd_f32(&tmDQF, acc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
      if ((rc = make_tmap_nd_f32(&tmDQFb, acc, 4, dims, strides, boxb, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    }
    if (!v2) {   // attention_bwd_t_kernel STORES the partial of an item's first key tile instead: no zero-fill
      cudaError_t e = cudaMemsetAsync(acc, 0, static_cast<size_t>(nacc) * sizeof(float), s);
      if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "attention_bwd_fused: cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    if (!tail) {   // with the remainder-token kernel in front, delta and the statistics are already there
      const long long drows = static_cast<long long>(B) * L * H;
      const int dgrid = static_cast<int>(std::min<long long>((drows * 8 + 255) / 256, 16LL * num_sms()));
      attention_bwd_delta_kernel<<<dgrid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(out), dg, delta, drows, L, H, hd, lse, stats, Lp);
      if ((rc = check_launch("attention_bwd_delta_kernel"))) return rc;
    }
    const int per = (items + grid - 1) / grid;
    if (v2) {
      // transposed score tiles, half-tile software pipeline (attention_bwd2.cu)
      if ((rc = launch_attention_bwd_t(tmQKV, tmDO, tmDQKV, tmQKVb, tmDOb, tmDQKVb, acc, B, stats, Lp, L, H, scale, items, grid,
                                       causal | (flags & 0xff00), Lm, wsp, qg, dg, hd, !(flags & OVK_ATT_BWD_8_WARPS), s)))
        return rc;
    } else if (ext)
      attention_bwd_kernel<MODE_FUSED, 16><<<grid, AB_THREADS, AB_SMEM_BYTES_RB_FUSED, s>>>(
          tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb, tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmDQF, tmDQFb, per);
    else
      attention_bwd_kernel<MODE_FUSED, 0><<<grid, AB_THREADS, AB_SMEM_BYTES, s>>>(
          tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb, tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmDQF, tmDQFb, per);
    if ((flags & OVK_ATT_BWD_ONEPASS_V1) && (rc = check_launch("attention_bwd_kernel<fused>"))) return rc;
    const long long total8 = static_cast<long long>(B) * Lm * H * (hd / 8);
    const int cgrid = static_cast<int>(std::min<long long>((total8 + 255) / 256, 16LL * num_sms()));
    attention_bwd_dq_convert_kernel<<<cgrid, 256, 0, s>>>(acc, reinterpret_cast<__nv_bfloat16*>(dqkv), qg, wsp, total8, L, Lm, H, hd, scale);
    return check_launch("attention_bwd_dq_convert_kernel");
  }
  if (ext) {
    attention_bwd_kernel<MODE_DQ, 16><<<grid, AB_THREADS, AB_SMEM_BYTES_RB, s>>>(tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb,
                                                                                 tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmQKV, tmQKVb, 0);
    if ((rc = check_launch("attention_bwd_kernel<dQ>"))) return rc;
    attention_bwd_kernel<MODE_DKV, 16><<<grid, AB_THREADS, AB_SMEM_BYTES_RB, s>>>(tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb,
                                                                                  tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmQKV, tmQKVb, 0);
  } else {
    attention_bwd_kernel<MODE_DQ, 0><<<grid, AB_THREADS, AB_SMEM_BYTES, s>>>(tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb,
                                                                             tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmQKV, tmQKVb, 0);
    if ((rc = check_launch("attention_bwd_kernel<dQ>"))) return rc;
    attention_bwd_kernel<MODE_DKV, 0><<<grid, AB_THREADS, AB_SMEM_BYTES, s>>>(tmQKV, tmO, tmDO, tmDQKV, tmQKVb, tmOb, tmDOb,
                                                                              tmDQKVb, lse, delta, L, H, scale, items, causal, Lm, wsp, qg, dg, hd, tmQKV, tmQKVb, 0);
  }
  return check_launch("attention_bwd_kernel<dKdV>");
}
