// Persistent warp-specialised tcgen05 GEMM main loop shared by every GEMM-shaped kernel in libovk:
//   D[M,N] (fp32, TMEM) = A[M,K] (bf16) * B[N,K]^T (bf16)
// Each operand may be K-major (rows of K, what nn.Linear weights and activations are) or MN-major (stored transposed:
// [K rows][M or N contiguous]) — the latter is what the backward GEMMs need (dX = dY*W reads W as [N(red), K(out)],
// dW = dY^T*X reads both operands with the token dimension outermost), so nothing is ever transposed in memory.
//
// Two flavours (template bool PAIR of the shared-memory layout):
//   PAIR = false: one CTA per SM, 128 x BN output tile, BK = 64 (one 128-byte swizzle atom), 4-stage TMA->smem ring.
//   PAIR = true : thread-block clusters of two CTAs (one SM pair) cooperate on a 256 x BN tile with
//                 tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the
//                 leader CTA's single MMA thread issues for both, each CTA's TMEM receives its 128 x BN accumulator.
//                 Per CTA and k-block that is 32 KB instead of 48 KB of L2->smem traffic and smem operand reads, and
//                 the ring is 6 stages deep.
// In both, two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle, warps 4..11 = epilogue:
// two groups of four warps, group g owns columns [g*BN/2, (g+1)*BN/2) of the tile, warp w reads TMEM lanes
// 32*(w%4)..+31 (the hardware's lane-quadrant rule for tcgen05.ld).
#pragma once
#include "ptx.cuh"

namespace ovk {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_CTRL_WARPS = 4;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (GEMM_CTRL_WARPS + GEMM_EPI_WARPS);  // 384
constexpr int GEMM_EPI_THREADS = 32 * GEMM_EPI_WARPS;                  // 256
constexpr int GEMM_GROUP_THREADS = 128;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_PANEL_BYTES = 64 * 128;                 // MN-major panel: 64 k-rows x 128 B

// C_BYTES: output staging area (TMA store), EPI_BYTES: epilogue scratch (bias tile, column statistics, ...).
template <int BN, int C_BYTES = 2 * GEMM_BM * 128, int EPI_BYTES = 2 * BN * 4, bool PAIR = false, int STAGES_ = 0>
struct GemmSmemLayout {
  static constexpr bool kPair = PAIR;
  static constexpr int STAGES = STAGES_ > 0 ? STAGES_ : (PAIR ? 6 : 4);   // ring depth (kernels with a big epilogue scratch ask for fewer)
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;  // rows of the B tile staged by THIS CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * GEMM_A_STAGE_BYTES;
  static constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;  // [128 rows x 128 B] staging buffers
  static constexpr int C_STAGE_BYTES = GEMM_BM * 128;
  static constexpr int OFF_EPI = OFF_C + C_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;
  // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2], aux[2], row-scale full / empty  + tmem base slot
  static constexpr int NUM_BARS = 2 * STAGES + 8;
  static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM_SLOT + 16;
  static constexpr int DYN_BYTES = TOTAL;  // the dynamic smem window starts 1024-byte aligned (checked in GemmCtx)
  static constexpr int TMEM_COLS = 2 * BN;        // 256 or 512 (power of two)
  static_assert(DYN_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");
};

struct GemmTileInfo {
  int m0, n0;
  int kb0, kb1;  // k-block range of this work item (split-K: several items share an output tile)
};

// Work items of one CTA (PAIR: of one CTA pair; `rank` = this CTA's rank in the pair selects its 128 rows).
struct GemmSched {
  int tiles_m, tiles_n, splits, kb_per, num_kb, total, bm, rank;
  __device__ __forceinline__ GemmSched(int M, int N, int BN, int K = GEMM_BK, int want_splits = 1, bool pair = false,
                                       int rank_ = 0) {
    bm = pair ? 2 * GEMM_BM : GEMM_BM;
    rank = rank_;
    tiles_m = (M + bm - 1) / bm;
    tiles_n = (N + BN - 1) / BN;
    num_kb = (K + GEMM_BK - 1) / GEMM_BK;
    kb_per = (num_kb + want_splits - 1) / want_splits;
    splits = (num_kb + kb_per - 1) / kb_per;
    total = tiles_m * tiles_n * splits;
  }
  // split fastest (the partial sums of one output tile are produced concurrently), then n (the CTAs resident at one
  // moment share A row-blocks through L2, B (weights) stays L2-resident).
  __device__ __forceinline__ GemmTileInfo tile(int t, int BN) const {
    GemmTileInfo ti;
    const int ks = t % splits;
    const int tt = t / splits;
    ti.n0 = (tt % tiles_n) * BN;
    ti.m0 = (tt / tiles_n) * bm + rank * GEMM_BM;
    ti.kb0 = ks * kb_per;
    ti.kb1 = min(num_kb, ti.kb0 + kb_per);
    return ti;
  }
};

template <int BN, class L_ = GemmSmemLayout<BN>>
struct GemmCtx {
  uint8_t* smem;  // 1024-aligned base
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  uint64_t* aux;  // two spare barriers for the epilogue groups (residual / auxiliary tile loads)
  uint64_t* rs;   // [0] full / [1] empty of the per-row scale vector (LayerNorm-folding GEMMs, filled by warp 3)
  uint32_t* tmem_slot;
  int rank;           // CTA rank inside the pair (0 when !PAIR)
  int first, stride;  // first work item and stride of this CTA (pair)
  using L = L_;
  __device__ __forceinline__ explicit GemmCtx(uint8_t* raw) {
    if ((smem_u32(raw) & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte aligned bases
      if (threadIdx.x == 0) printf("[ovk] gemm: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    smem = raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    full = bars;
    empty = bars + L::STAGES;
    tmem_full = bars + 2 * L::STAGES;
    tmem_empty = bars + 2 * L::STAGES + 2;
    aux = bars + 2 * L::STAGES + 4;
    rs = bars + 2 * L::STAGES + 6;
    tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_SLOT);
    if (L::kPair) {
      rank = static_cast<int>(cluster_ctarank());
      first = blockIdx.x >> 1;
      stride = gridDim.x >> 1;
    } else {
      rank = 0;
      first = blockIdx.x;
      stride = gridDim.x;
    }
  }
  __device__ __forceinline__ uint8_t* a_stage(int s) const { return smem + L::OFF_A + s * GEMM_A_STAGE_BYTES; }
  __device__ __forceinline__ uint8_t* b_stage(int s) const { return smem + L::OFF_B + s * L::B_STAGE_BYTES; }
  __device__ __forceinline__ uint8_t* c_stage(int s) const { return smem + L::OFF_C + s * L::C_STAGE_BYTES; }
  __device__ __forceinline__ uint8_t* epi_scratch() const { return smem + L::OFF_EPI; }
  // epilogue warps hand an accumulator buffer back to the (leader's) MMA warp
  __device__ __forceinline__ void release_accumulator(int acc) const {
    if (L::kPair) mbar_arrive_cluster(smem_u32(&tmem_empty[acc]) & PEER_BIT_MASK);
    else mbar_arrive(&tmem_empty[acc]);
  }
};

// Prologue executed by all threads: barrier init, TMEM alloc, descriptor prefetch. Returns TMEM base address.
template <int BN, class L>
__device__ __forceinline__ uint32_t gemm_prologue(const GemmCtx<BN, L>& cx, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                                  const CUtensorMap* tmC, const CUtensorMap* tmD = nullptr) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(tmA);
    tma_prefetch_desc(tmB);
    if (tmC) tma_prefetch_desc(tmC);
    if (tmD) tma_prefetch_desc(tmD);
  }
  if (warp == 1 && lane_id() == 0) {
    for (int i = 0; i < L::STAGES; ++i) {
      mbar_init(&cx.full[i], 1);
      mbar_init(&cx.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&cx.tmem_full[i], 1);
      mbar_init(&cx.tmem_empty[i], L::kPair ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);
      mbar_init(&cx.aux[i], 1);
      mbar_init(&cx.rs[i], i == 0 ? 32 : GEMM_EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (L::kPair) tmem_alloc_pair<L::TMEM_COLS>(cx.tmem_slot);
    else tmem_alloc<L::TMEM_COLS>(cx.tmem_slot);
  }
  tc_fence_before();
  if (L::kPair) cluster_sync();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  return *cx.tmem_slot;
}

template <int BN, class L>
__device__ __forceinline__ void gemm_teardown(const GemmCtx<BN, L>& cx, uint32_t tmem_base) {
  tc_fence_before();
  if (L::kPair) cluster_sync();  // neither CTA may exit (or free TMEM) while the other can still signal / read it
  else __syncthreads();
  if ((threadIdx.x >> 5) == 2) {
    tc_fence_after();
    if (L::kPair) tmem_dealloc_pair<L::TMEM_COLS>(tmem_base);
    else tmem_dealloc<L::TMEM_COLS>(tmem_base);
  }
}

// Warp 0, one elected lane. Streams A/B k-blocks of every tile this CTA owns through the smem ring.
// K-major operand: tensor map (inner = K, outer = rows), one box [rows x 64 k].
// MN-major operand: tensor map (inner = rows (M or N), outer = K), boxes of [64 k x 64 rows] = 8 KB panels.
// PAIR: each CTA loads its own A rows and its half of the B rows; every load signals the LEADER's full barrier.
template <int BN, bool A_MN, bool B_MN, class L>
__device__ __forceinline__ void gemm_producer(const GemmCtx<BN, L>& cx, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                              int M, int N, int K, int splits = 1) {
  constexpr bool PAIR = L::kPair;
  GemmSched sched(M, N, BN, K, splits, PAIR, cx.rank);
  int stage = 0;
  uint32_t phase = 0;
  for (int t = cx.first; t < sched.total; t += cx.stride) {
    GemmTileInfo ti = sched.tile(t, BN);
    const int bn0 = ti.n0 + (PAIR ? cx.rank * L::B_ROWS : 0);  // first B row staged by this CTA
    for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
      mbar_wait(&cx.empty[stage], phase ^ 1, 1);
      uint64_t* fb = &cx.full[stage];
      if (!PAIR) mbar_arrive_expect_tx(fb, L::STAGE_BYTES);
      else if (cx.rank == 0) mbar_arrive_expect_tx(fb, 2 * L::STAGE_BYTES);  // both CTAs' bytes land on the leader's barrier
      if constexpr (!A_MN) {
        tma_load_2d_x<PAIR>(cx.a_stage(stage), tmA, fb, kb * GEMM_BK, ti.m0);
      } else {
#pragma unroll
        for (int p = 0; p < GEMM_BM / 64; ++p)
          tma_load_2d_x<PAIR>(cx.a_stage(stage) + p * GEMM_PANEL_BYTES, tmA, fb, ti.m0 + 64 * p, kb * GEMM_BK);
      }
      if constexpr (!B_MN) {
        tma_load_2d_x<PAIR>(cx.b_stage(stage), tmB, fb, kb * GEMM_BK, bn0);
      } else {
#pragma unroll
        for (int p = 0; p < L::B_ROWS / 64; ++p)
          tma_load_2d_x<PAIR>(cx.b_stage(stage) + p * GEMM_PANEL_BYTES, tmB, fb, bn0 + 64 * p, kb * GEMM_BK);
      }
      if (++stage == L::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

// Warp 1, one elected lane (PAIR: of the leader CTA only). Issues BK/16 tcgen05.mma per k-block into the tile's TMEM
// accumulator buffer; completion is signalled to the smem ring / the epilogue with tcgen05.commit (PAIR: multicast to
// both CTAs' barriers).
template <int BN, bool A_MN, bool B_MN, class L>
__device__ __forceinline__ void gemm_mma_issuer(const GemmCtx<BN, L>& cx, uint32_t tmem_base, int M, int N, int K,
                                                int splits = 1) {
  constexpr bool PAIR = L::kPair;
  if (PAIR && cx.rank != 0) return;
  GemmSched sched(M, N, BN, K, splits, PAIR, cx.rank);
  constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  for (int t = cx.first; t < sched.total; t += cx.stride, ++it) {
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    mbar_wait(&cx.tmem_empty[acc], acc_phase ^ 1, 2);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * BN;
    const GemmTileInfo ti = sched.tile(t, BN);
    for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
      mbar_wait(&cx.full[stage], phase, 3);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(cx.a_stage(stage));
      const uint32_t b_addr = smem_u32(cx.b_stage(stage));
#pragma unroll
      for (int k = 0; k < GEMM_BK / 16; ++k) {
        const uint64_t ad = A_MN ? umma_desc_mnmajor_sw128(a_addr + k * 2048, GEMM_PANEL_BYTES)
                                 : umma_desc_kmajor_sw128(a_addr + k * 32);
        const uint64_t bd = B_MN ? umma_desc_mnmajor_sw128(b_addr + k * 2048, GEMM_PANEL_BYTES)
                                 : umma_desc_kmajor_sw128(b_addr + k * 32);
        umma_bf16_ss_x<PAIR>(d_tmem, ad, bd, idesc, ((kb - ti.kb0) | k) != 0);
      }
      umma_commit_x<PAIR>(&cx.empty[stage]);
      if (++stage == L::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit_x<PAIR>(&cx.tmem_full[acc]);
  }
}

}  // namespace ovk
