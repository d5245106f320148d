// ovk_gemm_*: C[M,N] = epilogue( alpha * op(A) * op(B) ) on tcgen05 tensor cores, bf16 operands, fp32 accumulation in TMEM.
// Replaces the reference's F.linear / addmm call sites (open_clip/transformer.py:225,233-235,250-252 via
// nn.MultiheadAttention in_proj/out_proj and mlp.c_fc/c_proj, :645-646 pooled @ proj) and, in the backward pass, the
// dgrad / wgrad matmuls autograd derives from them, with fused bias / GELU / residual / GELU' epilogues.
#include <stdlib.h>

#include "act.cuh"
#include "gemm_core.cuh"
#include "host_utils.h"

namespace ovk {

enum { EPI_LINEAR = 0, EPI_ACT = 1, EPI_DACT = 2 };

constexpr int GEMM_MAX_STAT_PARTS = 64;   // partial-sum slots (one per 128 columns of the LayerNorm width)

struct GemmEpi {
  const float* bias;  // f32[N] or null
  float alpha;
  const float* alpha_dev;  // optional DEVICE scalar multiplied into alpha (the temperature: no host sync to read it)
  int flags;          // OVK_EPI_BIAS | OVK_EPI_RESIDUAL | act id | OVK_EPI_SAVE_PREACT
  ActCoef act;
  // LayerNorm folded into the GEMM (FUSE kernels): y = rstd_i * acc + d_n with B = the row-centred W . gamma
  // (ovk_pack_ln_linear: sum_k B[n][k] = 0, which is what removes the mean), d = W beta + bias, and rstd from the per-row
  // (sum x, sum x^2) produced by the epilogue of the GEMM that wrote x.
  const float* row_stats_in; // f32[stats_parts_in][M][2] partial (sum, sum sq) of the A rows, or null
  float* row_stats_out;      // f32[ceil(N/128)][M][2] partial (sum, sum sq) of the output rows, or null
  int stats_parts_in;
  float inv_k, ln_eps;
  // FUSE kernels, linear epilogue: C[row] += row_add[row % row_period] (bf16 [row_period, N] table): the positional
  // embedding (+ class token in row 0) added to the patch-embedding GEMM's output, transformer.py:615-617
  const __nv_bfloat16* row_add;
  int row_period;
  // EPI_DACT: partial column sums of the OUTPUT (bf16-rounded, as autograd's bias gradient db = dU.sum(0) sees it):
  // f32 [2 * ceil(M / 128)][N], one row per 64 output rows, plain stores (deterministic); reduced by ovk_colsum_f32.
  float* colsum_ws;
};

// EPI_LINEAR: C = alpha*acc + bias (+ R)                   R  = residual tile [M,N] bf16 (tmR), may alias C
// EPI_ACT   : C = act(acc + bias), optionally D = acc+bias  D  = saved pre-activation (tmD) for the backward pass
// EPI_DACT  : C = alpha*acc * act'(R)                       R  = saved pre-activation
template <int BN, bool A_MN, bool B_MN, int EPI, bool OUT_F32, bool PAIR, bool FUSE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmD, const GemmEpi ep, int M, int N, int K, int splits) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using L = GemmSmemLayout<BN, 2 * GEMM_BM * 128, 2 * BN * 4 + (FUSE ? GEMM_BM * 4 : 0), PAIR>;
  GemmCtx<BN, L> cx(smem_raw);
  const uint32_t tmem_base = gemm_prologue(cx, &tmA, &tmB, &tmC, &tmR);
  const int warp = threadIdx.x >> 5;

  if (warp == 0) {
    if (elect_one()) gemm_producer<BN, A_MN, B_MN>(cx, &tmA, &tmB, M, N, K, splits);
  } else if (warp == 1) {
    if (elect_one()) gemm_mma_issuer<BN, A_MN, B_MN>(cx, tmem_base, M, N, K, splits);
  } else if (FUSE && warp == 3) {
    // ------------------------------------------------------------ row-scale producer (LayerNorm folded into the GEMM)
    // rstd of the 128 rows of the NEXT tile from the partial (sum, sum sq) slots, one tile ahead of the epilogue, so the
    // global-load latency never sits on the epilogue's critical path.  Slots are combined in slot order (deterministic).
    if (ep.row_stats_in != nullptr) {
      const uint32_t lane = lane_id();
      const uint32_t rs_base = smem_u32(cx.epi_scratch()) + 2 * BN * 4;
      GemmSched sched(M, N, BN, K, splits, PAIR, cx.rank);
      int it = 0;
      for (int t = cx.first; t < sched.total; t += cx.stride, ++it) {
        const GemmTileInfo ti = sched.tile(t, BN);
        float rstd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ti.m0 + static_cast<int>(lane) + 32 * i;
          float s1 = 0.f, s2 = 0.f;
          if (row < M) {
            const float2* src = reinterpret_cast<const float2*>(ep.row_stats_in) + row;
            for (int q = 0; q < ep.stats_parts_in; ++q) {
              const float2 v = __ldg(src + static_cast<long long>(q) * M);
              s1 += v.x;
              s2 += v.y;
            }
          }
          const float mu = s1 * ep.inv_k;
          rstd[i] = rsqrtf(fmaxf(s2 * ep.inv_k - mu * mu, 0.f) + ep.ln_eps);   // rows >= M: finite, never stored
        }
        mbar_wait(&cx.rs[1], (it & 1) ^ 1, 6);
#pragma unroll
        for (int i = 0; i < 4; ++i) sts_f32(rs_base + (lane + 32 * i) * 4, rstd[i]);
        mbar_arrive(&cx.rs[0]);
      }
    }
  } else if (warp >= GEMM_CTRL_WARPS) {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> smem -> TMA store
    constexpr int CW = OUT_F32 ? 32 : 64;         // columns per staged chunk (one 128-byte row)
    constexpr int GROUP_COLS = BN / 2;
    constexpr int NCHUNK = GROUP_COLS / CW;
    const int ew = warp - GEMM_CTRL_WARPS;
    const int grp = ew >> 2;                       // column half of the tile
    const int quad = ew & 3;                       // TMEM lane quadrant
    const uint32_t lane = lane_id();
    const int et = quad * 32 + lane;               // row inside the tile, thread index inside the group
    const bool leader = et == 0;
    const uint32_t bar_id = 1 + grp;
    const uint32_t sbuf = smem_u32(cx.c_stage(grp));
    const uint32_t bias_base = smem_u32(cx.epi_scratch());
    uint64_t* rbar = &cx.aux[grp];
    uint32_t rphase = 0;
    const bool has_bias = (ep.flags & OVK_EPI_BIAS) != 0;
    const bool has_res = !OUT_F32 && (EPI == EPI_DACT || (EPI == EPI_LINEAR && (ep.flags & OVK_EPI_RESIDUAL) != 0));
    const bool save_pre = EPI == EPI_ACT && (ep.flags & OVK_EPI_SAVE_PREACT) != 0;
    const float alpha = ep.alpha_dev != nullptr ? ep.alpha * __ldg(ep.alpha_dev) : ep.alpha;
    GemmSched sched(M, N, BN, K, splits, PAIR, cx.rank);
    const bool reduce_out = OUT_F32 && sched.splits > 1;   // split-K partial sums are ADDED into a zeroed C
    int it = 0;
    for (int t = cx.first; t < sched.total; t += cx.stride, ++it) {
      const GemmTileInfo ti = sched.tile(t, BN);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // per-column bias tile -> smem, double-buffered by tile parity (one barrier publishes it)
      const uint32_t bias_s = bias_base + (it & 1) * (BN * 4);
      {
        const int j = threadIdx.x - 32 * GEMM_CTRL_WARPS;
        if (j < BN) {
          const int n = ti.n0 + j;
          sts_f32(bias_s + j * 4, (has_bias && n < N) ? ep.bias[n] : 0.f);
        }
      }
      named_bar_sync(3, GEMM_EPI_THREADS);
      float a_mul = alpha, st1 = 0.f, st2 = 0.f;
      if constexpr (FUSE) {
        if (ep.row_stats_in != nullptr) {   // this row's rstd, prepared by warp 3
          mbar_wait(&cx.rs[0], it & 1, 7);
          a_mul = lds_f32(bias_base + 2 * BN * 4 + et * 4);
          mbar_arrive(&cx.rs[1]);
        }
      }
      mbar_wait(&cx.tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + grp * GROUP_COLS;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        const int col_in_tile = grp * GROUP_COLS + c * CW;
        const int ncol0 = ti.n0 + col_in_tile;
        const bool live = ncol0 < N;  // uniform across the group
        if (live) {
          // the staging buffer is free once the previous TMA store has finished reading it
          if (leader) tma_store_wait_read<0>();
          named_bar_sync(bar_id, GEMM_GROUP_THREADS);
          if (has_res && leader) {
            mbar_arrive_expect_tx(rbar, GEMM_BM * 128);
            tma_load_2d(cx.c_stage(grp), &tmR, rbar, ncol0, ti.m0);
          }
        }
        uint32_t v[CW];
        if (live) {
#pragma unroll
          for (int q = 0; q < CW / 32; ++q) {
            uint32_t(&vq)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32 * q]);
            tmem_ld_x32(taddr + c * CW + 32 * q, vq);
          }
          tmem_ld_wait();
        }
        if (c == NCHUNK - 1) {  // accumulator drained into registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) cx.release_accumulator(acc);
        }
        if (!live) continue;
        float x[CW];
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 b = lds_f32x4(bias_s + (col_in_tile + 4 * j) * 4);
          x[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), a_mul, b.x);
          x[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a_mul, b.y);
          x[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a_mul, b.z);
          x[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a_mul, b.w);
        }
        if constexpr (EPI == EPI_ACT) {
          if (save_pre) {  // pre-activation out first (same staging buffer), then the activation
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(sbuf + sw128_offset(et, j), make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                                            pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7])));
            fence_proxy_async_smem();
            named_bar_sync(bar_id, GEMM_GROUP_THREADS);
            if (leader) {
              tma_store_2d(&tmD, cx.c_stage(grp), ncol0, ti.m0);
              tma_store_commit();
            }
          }
#pragma unroll
          for (int j = 0; j < CW; ++j) x[j] = act_fwd(x[j], ep.act);
          if (save_pre) {
            if (leader) tma_store_wait_read<0>();
            named_bar_sync(bar_id, GEMM_GROUP_THREADS);
          }
        }
        if constexpr (!OUT_F32) if (has_res) {
          mbar_wait(rbar, rphase, 5);
          rphase ^= 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 r = lds128(sbuf + sw128_offset(et, j));
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float r0 = bf16_lo(rw[q]), r1 = bf16_hi(rw[q]);
              if constexpr (EPI == EPI_DACT) {
                x[8 * j + 2 * q] *= act_bwd(r0, ep.act);
                x[8 * j + 2 * q + 1] *= act_bwd(r1, ep.act);
              } else {
                x[8 * j + 2 * q] += r0;
                x[8 * j + 2 * q + 1] += r1;
              }
            }
          }
        }
        if constexpr (FUSE && EPI == EPI_LINEAR && !OUT_F32) {
          if (ep.row_add != nullptr) {   // token row -> its positional-embedding row (N % 64 == 0 checked by the host)
            const int row_g = ti.m0 + et;
            if (row_g < M) {
              const uint4* src = reinterpret_cast<const uint4*>(ep.row_add + static_cast<long long>(row_g % ep.row_period) * N + ncol0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 r = __ldg(src + j);
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  x[8 * j + 2 * q] += bf16_lo(rw[q]);
                  x[8 * j + 2 * q + 1] += bf16_hi(rw[q]);
                }
              }
            }
          }
        }
        if constexpr (OUT_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(sbuf + sw128_offset(et, j), make_uint4(__float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                                                          __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3])));
        } else {
          if constexpr (FUSE) {
            // row statistics of the tensor being written (for the NEXT LayerNorm); taken before the bf16 rounding, which
            // moves the mean and the variance by ~1e-5 relative
            if (ep.row_stats_out != nullptr) {
              float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                p1[j & 3] += x[j];
                p2[j & 3] = fmaf(x[j], x[j], p2[j & 3]);
              }
              st1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
              st2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
            }
          }
          uint32_t w[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) w[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 8; ++j) sts128(sbuf + sw128_offset(et, j), make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, GEMM_GROUP_THREADS);
        if (leader) {
          if (reduce_out) tma_reduce_add_2d(&tmC, cx.c_stage(grp), ncol0, ti.m0);
          else tma_store_2d(&tmC, cx.c_stage(grp), ncol0, ti.m0);
          tma_store_commit();
        }
        if constexpr (EPI == EPI_DACT && !OUT_F32) {
          // bias gradient of the layer whose pre-activation gradient this GEMM writes: column sums of the staged bf16 tile
          // (thread -> one column x 64 rows; the swizzle spreads a column's rows over the banks, lanes = consecutive columns:
          // conflict-free 2-byte loads).  Saves the stand-alone pass that re-read the whole [M, N] tensor from HBM.  The
          // staging buffer is rewritten only after the group's next barrier, which every thread reaches after these reads.
          if (ep.colsum_ws != nullptr && ti.m0 < M) {   // (CTA pairs: the second CTA of the last pair may sit past the last row)
            const int cc = et & 63, half = et >> 6;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 64; rr += 2) {
              const int r0 = half * 64 + rr;
              unsigned short a, b;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(a) : "r"(sbuf + sw128_offset(r0, cc >> 3) + (cc & 7) * 2));
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(b) : "r"(sbuf + sw128_offset(r0 + 1, cc >> 3) + (cc & 7) * 2));
              s0 += __uint_as_float(static_cast<uint32_t>(a) << 16);
              s1 += __uint_as_float(static_cast<uint32_t>(b) << 16);
            }
            if (ncol0 + cc < N) ep.colsum_ws[static_cast<long long>((ti.m0 >> 7) * 2 + half) * N + ncol0 + cc] = s0 + s1;
          }
        }
      }
      if constexpr (FUSE) {
        // one slot per (row, 128 output columns) = per epilogue group of a 256-wide tile: plain stores, no atomics
        if (ep.row_stats_out != nullptr && ti.m0 + et < M && ti.n0 + grp * GROUP_COLS < N) {
          const long long slot = (ti.n0 + grp * GROUP_COLS) >> 7;
          reinterpret_cast<float2*>(ep.row_stats_out)[slot * M + ti.m0 + et] = make_float2(st1, st2);
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
  }
  gemm_teardown(cx, tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
struct GemmArgs {
  const void* A; long long lda; bool a_mn;   // a_mn: A stored [K, M] (M contiguous)
  const void* B; long long ldb; bool b_mn;   // b_mn: B stored [K, N] (N contiguous); else [N, K]
  void* C; long long ldc; bool c_f32;
  const void* R; long long ldr;              // residual / saved pre-activation (bf16 [M, ldr]) or null
  void* D; long long ldd;                    // pre-activation output (bf16) or null
  int M, N, K;
  GemmEpi ep;
  int epi;
};

static bool pair_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OVK_GEMM_PAIR");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;   // default on; OVK_GEMM_PAIR=0 selects the single-CTA kernels
  }
  return v == 1;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool OUT_F32, bool PAIR, bool FUSE = false>
static int launch_gemm_t(const GemmArgs& g, cudaStream_t stream) {
  using L = GemmSmemLayout<BN, 2 * GEMM_BM * 128, 2 * BN * 4 + (FUSE ? GEMM_BM * 4 : 0), PAIR>;
  CUtensorMap tmA, tmB, tmC, tmR, tmD;
  int rc;
  if (A_MN) rc = make_tmap_2d_bf16(&tmA, g.A, g.M, g.K, g.lda, 64, 64);
  else rc = make_tmap_2d_bf16(&tmA, g.A, g.K, g.M, g.lda, GEMM_BK, GEMM_BM);
  if (rc) return rc;
  if (B_MN) rc = make_tmap_2d_bf16(&tmB, g.B, g.N, g.K, g.ldb, 64, 64);
  else rc = make_tmap_2d_bf16(&tmB, g.B, g.K, g.N, g.ldb, GEMM_BK, L::B_ROWS);
  if (rc) return rc;
  if (OUT_F32) rc = make_tmap_2d_f32(&tmC, g.C, g.N, g.M, g.ldc, 32, GEMM_BM);
  else rc = make_tmap_2d_bf16(&tmC, g.C, g.N, g.M, g.ldc, 64, GEMM_BM);
  if (rc) return rc;
  tmR = tmC;
  tmD = tmC;
  if (g.R && (rc = make_tmap_2d_bf16(&tmR, g.R, g.N, g.M, g.ldr, 64, GEMM_BM))) return rc;
  if (g.D && (rc = make_tmap_2d_bf16(&tmD, g.D, g.N, g.M, g.ldd, 64, GEMM_BM))) return rc;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI, OUT_F32, PAIR, FUSE>;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_once.done();
  }
  const int bm = PAIR ? 2 * GEMM_BM : GEMM_BM;
  const int tiles = ((g.M + bm - 1) / bm) * ((g.N + BN - 1) / BN);
  const int units = PAIR ? num_sms() / 2 : num_sms();   // CTA pairs or CTAs that can be resident
  // split-K (fp32 outputs only: the weight-gradient GEMMs have few output tiles and a huge K = tokens): partial sums
  // are added into a zeroed C by TMA reduce-add stores.  The split count is the one (<= 16, at least 8 k-blocks each)
  // that fills the waves best: e.g. 75 tiles on 74 CTA pairs are two waves at 51 % unsplit, nine waves at 90 % split 8
  // ways; every extra split costs one more fp32 reduce-add pass over the output, hence the penalty term.
  int splits = 1;
  const char* pol = getenv("OVK_SPLITK_SIMPLE");   // =1: the first-cut rule (fill one wave), kept for A/B measurements
  if (OUT_F32 && pol != nullptr && pol[0] == '1') {
    if (tiles * 10 < units * 7) {
      const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
      splits = units / tiles;
      if (splits > num_kb / 8) splits = num_kb / 8;
      if (splits < 1) splits = 1;
    }
  } else if (OUT_F32) {
    const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
    const int max_s = num_kb / 8 < 16 ? num_kb / 8 : 16;
    float best = -1.f;
    for (int sct = 1; sct <= max_s; ++sct) {
      const int kb_per = (num_kb + sct - 1) / sct;
      const int eff_s = (num_kb + kb_per - 1) / kb_per;   // splits the scheduler will really make
      if (eff_s != sct) continue;
      const long long items_s = static_cast<long long>(tiles) * sct;
      const long long waves = (items_s + units - 1) / units;
      // every split adds one fp32 reduce-add pass over the output: measured ~2.5 TB/s through L2, i.e. ~1040 / K of the
      // GEMM's own time per split
      const float eff = static_cast<float>(items_s) / static_cast<float>(waves * units) -
                        (1040.f / static_cast<float>(g.K)) * static_cast<float>(sct - 1);
      if (eff > best) {
        best = eff;
        splits = sct;
      }
    }
  }
  if (splits > 1) {
    cudaError_t e = cudaMemset2DAsync(g.C, static_cast<size_t>(g.ldc) * 4, 0, static_cast<size_t>(g.N) * 4, g.M, stream);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaMemset2DAsync(gemm split-K): %s", cudaGetErrorString(e));
  }
  const int items = tiles * splits;
  const int nunits = items < units ? items : units;
  if (!PAIR) {
    kern<<<nunits, GEMM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, tmC, tmR, tmD, g.ep, g.M, g.N, g.K, splits);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nunits, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = L::DYN_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, tmD, g.ep, g.M, g.N, g.K, splits);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaLaunchKernelEx(gemm pair): %s", cudaGetErrorString(e));
  }
  return check_launch("gemm_bf16_kernel");
}

template <bool A_MN, bool B_MN, int EPI, bool OUT_F32, bool FUSE = false>
static int launch_gemm_bn(const GemmArgs& g, cudaStream_t s) {
  if (g.N <= 128) return launch_gemm_t<128, A_MN, B_MN, EPI, OUT_F32, false, FUSE>(g, s);
  // CTA pairs (256-row tiles) once there are enough rows to fill them; small problems stay on single CTAs
  if (pair_mode_enabled() && g.M >= 512) return launch_gemm_t<256, A_MN, B_MN, EPI, OUT_F32, true, FUSE>(g, s);
  return launch_gemm_t<256, A_MN, B_MN, EPI, OUT_F32, false, FUSE>(g, s);
}

static int check_common(const GemmArgs& g, const char* who) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(OVK_ERR_SHAPE, "%s: empty problem M=%d N=%d K=%d", who, g.M, g.N, g.K);
  if ((g.lda % 8) || (g.ldb % 8) || (g.ldc % (g.c_f32 ? 4 : 8)))
    return set_error(OVK_ERR_ALIGN, "%s: leading dimensions must be multiples of 16 bytes (TMA strides)", who);
  // extents need no alignment (TMA zero-fills out-of-bounds box elements); only strides and base pointers do
  if (g.R && (g.ldr % 8)) return set_error(OVK_ERR_ALIGN, "%s: ldr must be a multiple of 8", who);
  if (g.D && (g.ldd % 8)) return set_error(OVK_ERR_ALIGN, "%s: ldd must be a multiple of 8", who);
  return OVK_OK;
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M,
                             int N, int K, const float* bias, const void* residual, long long ldr, int flags,
                             void* stream) {
  return ovk_gemm_bf16_ex(A, lda, B, ldb, C, ldc, M, N, K, bias, residual, ldr, nullptr, 0, flags, stream);
}

extern "C" int ovk_gemm_bf16_ex(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                int M, int N, int K, const float* bias, const void* residual, long long ldr,
                                void* preact, long long ldp, int flags, void* stream) {
  GemmArgs g{A, lda, false, B, ldb, false, C, ldc, false, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  const int act = flags & OVK_EPI_ACT_MASK;
  if ((flags & OVK_EPI_RESIDUAL) && residual == nullptr) return set_error(OVK_ERR_SHAPE, "gemm: residual flag without pointer");
  if ((flags & OVK_EPI_BIAS) && bias == nullptr) return set_error(OVK_ERR_SHAPE, "gemm: bias flag without pointer");
  if ((flags & OVK_EPI_SAVE_PREACT) && (preact == nullptr || act == 0))
    return set_error(OVK_ERR_SHAPE, "gemm: SAVE_PREACT needs an activation and an output pointer");
  if (act && (flags & OVK_EPI_RESIDUAL)) return set_error(OVK_ERR_SHAPE, "gemm: activation + residual in one epilogue is not supported");
  g.ep.bias = bias;
  g.ep.alpha = 1.f;
  g.ep.flags = flags;
  g.ep.act = act_coef(act);
  if (flags & OVK_EPI_RESIDUAL) g.R = residual, g.ldr = ldr;
  if (flags & OVK_EPI_SAVE_PREACT) g.D = preact, g.ldd = ldp;
  int rc = check_common(g, "gemm");
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (act) return launch_gemm_bn<false, false, EPI_ACT, false>(g, s);
  return launch_gemm_bn<false, false, EPI_LINEAR, false>(g, s);
}

extern "C" int ovk_gemm_bf16_nn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                int c_is_f32, int M, int N, int K, float alpha, const void* preact, long long ldp,
                                int act, void* stream) {
  GemmArgs g{A, lda, false, B, ldb, true, C, ldc, c_is_f32 != 0, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  g.ep.bias = nullptr;
  g.ep.alpha = alpha;
  g.ep.flags = 0;
  g.ep.act = act_coef(act & OVK_EPI_ACT_MASK);
  if (preact) {
    if (c_is_f32) return set_error(OVK_ERR_SHAPE, "gemm_nn: the act' epilogue writes bf16");
    if (!(act & OVK_EPI_ACT_MASK)) return set_error(OVK_ERR_SHAPE, "gemm_nn: preact given without an activation id");
    g.R = preact, g.ldr = ldp;
  }
  int rc = check_common(g, "gemm_nn");
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (preact) return launch_gemm_bn<false, true, EPI_DACT, false>(g, s);
  if (c_is_f32) return launch_gemm_bn<false, true, EPI_LINEAR, true>(g, s);
  return launch_gemm_bn<false, true, EPI_LINEAR, false>(g, s);
}

extern "C" long long ovk_gemm_colsum_rows(int M) { return 2LL * ((M + GEMM_BM - 1) / GEMM_BM); }

extern "C" int ovk_gemm_bf16_nn_dact_colsum(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                            int M, int N, int K, float alpha, const void* preact, long long ldp, int act,
                                            float* colsum_ws, void* stream) {
  GemmArgs g{A, lda, false, B, ldb, true, C, ldc, false, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  if (!preact || !(act & OVK_EPI_ACT_MASK)) return set_error(OVK_ERR_SHAPE, "gemm_nn_dact_colsum: preact and an activation id are required");
  if (!colsum_ws) return set_error(OVK_ERR_SHAPE, "gemm_nn_dact_colsum: workspace required (f32 [ovk_gemm_colsum_rows(M)][N])");
  g.ep.bias = nullptr;
  g.ep.alpha = alpha;
  g.ep.flags = 0;
  g.ep.act = act_coef(act & OVK_EPI_ACT_MASK);
  g.ep.colsum_ws = colsum_ws;
  g.R = preact, g.ldr = ldp;
  int rc = check_common(g, "gemm_nn_dact_colsum");
  if (rc) return rc;
  return launch_gemm_bn<false, true, EPI_DACT, false>(g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ovk_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                int c_is_f32, int M, int N, int K, float alpha, void* stream) {
  GemmArgs g{A, lda, true, B, ldb, true, C, ldc, c_is_f32 != 0, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  g.ep.bias = nullptr;
  g.ep.alpha = alpha;
  g.ep.flags = 0;
  g.ep.act = act_coef(0);
  int rc = check_common(g, "gemm_tn");
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (c_is_f32) return launch_gemm_bn<true, true, EPI_LINEAR, true>(g, s);
  return launch_gemm_bn<true, true, EPI_LINEAR, false>(g, s);
}

extern "C" int ovk_gemm_bf16_scaled(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* C,
                                    long long ldc, int c_is_f32, int M, int N, int K, float alpha, const float* alpha_dev,
                                    void* stream) {
  GemmArgs g{A, lda, a_mn != 0, B, ldb, b_mn != 0, C, ldc, c_is_f32 != 0, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  g.ep.bias = nullptr;
  g.ep.alpha = alpha;
  g.ep.alpha_dev = alpha_dev;
  g.ep.flags = 0;
  g.ep.act = act_coef(0);
  int rc = check_common(g, "gemm_scaled");
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a_mn && !b_mn) return set_error(OVK_ERR_SHAPE, "gemm_scaled: A^T * B^T is not provided (no call site needs it)");
  if (!a_mn && !b_mn) return c_is_f32 ? launch_gemm_bn<false, false, EPI_LINEAR, true>(g, s) : launch_gemm_bn<false, false, EPI_LINEAR, false>(g, s);
  if (!a_mn && b_mn) return c_is_f32 ? launch_gemm_bn<false, true, EPI_LINEAR, true>(g, s) : launch_gemm_bn<false, true, EPI_LINEAR, false>(g, s);
  return c_is_f32 ? launch_gemm_bn<true, true, EPI_LINEAR, true>(g, s) : launch_gemm_bn<true, true, EPI_LINEAR, false>(g, s);
}

extern "C" int ovk_gemm_bf16_rowadd(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M,
                                    int N, int K, const void* row_add, int row_period, void* stream) {
  GemmArgs g{A, lda, false, B, ldb, false, C, ldc, false, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  if (row_add == nullptr || row_period <= 0) return set_error(OVK_ERR_SHAPE, "gemm_rowadd: row_add table and period are required");
  if (N % 64) return set_error(OVK_ERR_SHAPE, "gemm_rowadd: N must be a multiple of 64");
  if (reinterpret_cast<uintptr_t>(row_add) & 15) return set_error(OVK_ERR_ALIGN, "gemm_rowadd: row_add must be 16-byte aligned");
  g.ep.alpha = 1.f;
  g.ep.act = act_coef(0);
  g.ep.row_add = reinterpret_cast<const __nv_bfloat16*>(row_add);
  g.ep.row_period = row_period;
  int rc = check_common(g, "gemm_rowadd");
  if (rc) return rc;
  return launch_gemm_bn<false, false, EPI_LINEAR, false, true>(g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ovk_gemm_bf16_ln(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                int M, int N, int K, const float* bias, const float* row_stats_in, int stats_parts_in,
                                float ln_eps, const void* residual, long long ldr, void* preact, long long ldp,
                                float* row_stats_out, int flags, void* stream) {
  GemmArgs g{A, lda, false, B, ldb, false, C, ldc, false, nullptr, 0, nullptr, 0, M, N, K, {}, EPI_LINEAR};
  const int act = flags & OVK_EPI_ACT_MASK;
  if ((flags & OVK_EPI_RESIDUAL) && residual == nullptr) return set_error(OVK_ERR_SHAPE, "gemm_ln: residual flag without pointer");
  if ((flags & OVK_EPI_BIAS) && bias == nullptr) return set_error(OVK_ERR_SHAPE, "gemm_ln: bias flag without pointer");
  if ((flags & OVK_EPI_SAVE_PREACT) && (preact == nullptr || act == 0))
    return set_error(OVK_ERR_SHAPE, "gemm_ln: SAVE_PREACT needs an activation and an output pointer");
  if (act && (flags & OVK_EPI_RESIDUAL)) return set_error(OVK_ERR_SHAPE, "gemm_ln: activation + residual in one epilogue is not supported");
  if (row_stats_in != nullptr && (stats_parts_in < 1 || stats_parts_in > GEMM_MAX_STAT_PARTS))
    return set_error(OVK_ERR_SHAPE, "gemm_ln: stats_parts_in must be in [1, %d]", GEMM_MAX_STAT_PARTS);
  if (row_stats_out != nullptr && ((N % 64) || N <= 128))
    return set_error(OVK_ERR_SHAPE, "gemm_ln: row statistics output needs N %% 64 == 0 and N > 128");
  if (row_stats_out != nullptr && act) return set_error(OVK_ERR_SHAPE, "gemm_ln: row statistics output is for the linear epilogue");
  g.ep.bias = bias;
  g.ep.alpha = 1.f;
  g.ep.flags = flags;
  g.ep.act = act_coef(act);
  g.ep.row_stats_in = row_stats_in;
  g.ep.row_stats_out = row_stats_out;
  g.ep.stats_parts_in = stats_parts_in;
  g.ep.inv_k = 1.f / static_cast<float>(K);
  g.ep.ln_eps = ln_eps;
  if (flags & OVK_EPI_RESIDUAL) g.R = residual, g.ldr = ldr;
  if (flags & OVK_EPI_SAVE_PREACT) g.D = preact, g.ldd = ldp;
  int rc = check_common(g, "gemm_ln");
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (act) return launch_gemm_bn<false, false, EPI_ACT, false, true>(g, s);
  return launch_gemm_bn<false, false, EPI_LINEAR, false, true>(g, s);
}
