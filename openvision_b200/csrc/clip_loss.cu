// Fused CLIP contrastive loss (open_clip/loss.py:102-131, ClipLoss.get_logits + 2 x F.cross_entropy; JAX twin
// src/losses/common.py:120-189).  The N x N logits z = s * I T^T are produced tile by tile on tcgen05 and consumed in the
// epilogue: they are never written to memory in the forward pass.
//
// forward  (clip_loss_fwd_kernel): per 128 x 256 tile, each epilogue thread owns one row x 128 columns of the fp32
//   accumulator.  One exponential per logit serves BOTH directions: e = 2^(z2 - c) with c the maximum of the warp's
//   32 x 32 sub-block; row sums are thread-local (online-softmax recurrence, src/models/bpt.py:105-124), column sums go
//   through a per-warp shared-memory transposition (warp_column_sums).  Rows / columns whose own maximum lies more than 2^100 below
//   c (so that e would lose them) take an exact slow path, so the result does not depend on how well-scaled the
//   features are.  Partials (max, sum) per (row, column-half of a tile) and per (row-tile, column) go to a workspace;
//   clip_loss_finalize_kernel merges them into row_lse, col_max / col_sum (a rank's partial column statistics).
// backward (clip_loss_grad_kernel): recomputes z on tcgen05 and writes G = dL/dz (bf16) with the epilogue
//   G_ij = w_row 2^(z2 - rlse_i) + w_col 2^(z2 - clse_j) - (w_row + w_col) [j == label(i)],  plus sum(G * z) / s.
//   dI = s G T and dT = s G^T I are then ovk_gemm_bf16_nn / ovk_gemm_bf16_tn.
#include "gemm_core.cuh"
#include "host_utils.h"

namespace ovk {

constexpr int CL_BN = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float NEG_INF = -INFINITY;

// Warp maximum with ONE redux instruction on an order-preserving integer key (negatives: all bits flipped, the rest: sign bit
// flipped) instead of five dependent shuffle + max steps: the reference offset of every 32 x 32 sub-block waits for it, and the
// epilogue of the forward kernel is latency-bound (profiles: wait / short-scoreboard stalls, tensor pipe 70-77 %).
__device__ __forceinline__ float warp_max_f(float v) {
  const uint32_t b = __float_as_uint(v);
  uint32_t k = b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
  k = __reduce_max_sync(0xffffffffu, k);
  return __uint_as_float(k ^ (((k >> 31) - 1u) | 0x80000000u));
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transposing reduction through shared memory: lane l holds e[0..31] (32 columns of its row); returns, in lane j, the sum
// over the warp's 32 rows of column j.  Column-major scratch with a pitch of 36 floats: the 32 scalar stores of a column
// are one conflict-free wavefront each (lanes = rows = consecutive words), and lane j reads its column back as eight
// 16-byte loads (144-byte pitch: the 8 lanes of a quarter-warp phase cover all 32 banks).  73 instructions per 32 x 32 block
// against 124 for the shuffle butterfly this replaces (31 SHFL + 31 FADD + 62 FSEL); the kernel is epilogue-bound.
constexpr int CL_TP = 36;                         // scratch pitch (floats)
constexpr int CL_T_BYTES = 32 * CL_TP * 4;        // per warp
__device__ __forceinline__ float warp_column_sums(const float (&e)[32], uint32_t lane, uint32_t scratch) {
#pragma unroll
  for (int c = 0; c < 32; ++c) sts_f32(scratch + (c * CL_TP + lane) * 4, e[c]);
  __syncwarp();
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = lds_f32x4(scratch + (lane * CL_TP + 4 * q) * 4);
    s[0] += v.x;
    s[1] += v.y;
    s[2] += v.z;
    s[3] += v.w;
  }
  __syncwarp();   // the scratch is rewritten by the next chunk
  return (s[0] + s[1]) + (s[2] + s[3]);
}

// ------------------------------------------------------------------------------------------------ forward
// epilogue scratch: [4 quads][256 cols] (ref, sum) + one transposition block per epilogue warp; the ring gives up a stage for it
template <bool PAIR>
using FwdLayout = GemmSmemLayout<CL_BN, 0, 4 * CL_BN * 8 + GEMM_EPI_WARPS * CL_T_BYTES, PAIR, PAIR ? 5 : 3>;

template <bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
clip_loss_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     float2* __restrict__ ws_row, float2* __restrict__ ws_col, float* __restrict__ diag, int n_loc,
                     int n_cols, int col_offset, int n_all, int E, int row_offset, const float* __restrict__ scale_ptr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using L = FwdLayout<PAIR>;
  GemmCtx<CL_BN, L> cx(smem_raw);
  const uint32_t tmem_base = gemm_prologue(cx, &tmA, &tmB, nullptr);
  const int warp = threadIdx.x >> 5;

  if (warp == 0) {
    if (elect_one()) gemm_producer<CL_BN, false, false>(cx, &tmA, &tmB, n_loc, n_cols, E);
  } else if (warp == 1) {
    if (elect_one()) gemm_mma_issuer<CL_BN, false, false>(cx, tmem_base, n_loc, n_cols, E);
  } else if (warp >= GEMM_CTRL_WARPS) {
    const int ew = warp - GEMM_CTRL_WARPS;
    const int grp = ew >> 2;
    const int quad = ew & 3;
    const uint32_t lane = lane_id();
    const int et = quad * 32 + lane;
    const uint32_t bar_id = 1 + grp;
    const uint32_t scol = smem_u32(cx.epi_scratch());  // float2 [4][256]
    const uint32_t tscr = scol + 4 * CL_BN * 8 + ew * CL_T_BYTES;   // this warp's transposition block
    const float scale = __ldg(scale_ptr);              // the temperature lives on the device (no host read-back)
    const float s2 = scale * LOG2E;
    const bool fast_ok = s2 > 0.f;   // (the temperature exp(logit_scale) is positive; anything else takes the general path)
    GemmSched sched(n_loc, n_cols, CL_BN, E, 1, PAIR, cx.rank);
    int it = 0;
    for (int t = cx.first; t < sched.total; t += cx.stride, ++it) {
      const GemmTileInfo ti = sched.tile(t, CL_BN);
      const int tile_n = (col_offset + ti.n0) / CL_BN;   // slot index over ALL columns (col_offset % 256 == 0)
      const int tile_m = ti.m0 / GEMM_BM;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = ti.m0 + et;
      const bool row_valid = row < n_loc;
      const int label = row_offset + row - col_offset;  // column (of this window) holding this row's positive pair
      float m_row = NEG_INF, l_row = 0.f;
      mbar_wait(&cx.tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * CL_BN + grp * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col_in_tile = grp * 128 + c * 32;
        const int col0 = ti.n0 + col_in_tile;
        const uint32_t sdst = scol + ((quad * CL_BN + col_in_tile + lane) << 3);
        if (col0 >= n_cols) {  // uniform: chunk entirely past the last column
          sts_f32x2(sdst, NEG_INF, 0.f);
          continue;
        }
        uint32_t v[32];
        tmem_ld_x32(taddr + c * 32, v);
        tmem_ld_wait();
        float z[32];
        float cmax;
        // interior chunk (every row and column valid, positive temperature — i.e. always, except at the edges of the
        // matrix): the maximum is taken on the raw accumulators (s2 > 0 commutes with max: 16 three-input maxima instead of
        // 32 multiplies + 32 maxima) and the scaled logits are never formed: the exponent argument is one packed FFMA2 per
        // pair further down.  The epilogue is what bounds this kernel (tensor pipe 70 % busy), so instructions count.
        const bool interior = fast_ok && ti.m0 + GEMM_BM <= n_loc && col0 + 32 <= n_cols;   // uniform
        if (interior) {
          float cm[4] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF};
#pragma unroll
          for (int j = 0; j < 16; ++j) cm[j & 3] = fmaxf(fmaxf(cm[j & 3], __uint_as_float(v[2 * j])), __uint_as_float(v[2 * j + 1]));
          cmax = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) * s2;
        } else {
          float cm[4] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF};   // four independent chains (the epilogue is latency-bound)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            z[j] = (row_valid && col0 + j < n_cols) ? __uint_as_float(v[j]) * s2 : NEG_INF;
            cm[j & 3] = fmaxf(cm[j & 3], z[j]);
          }
          cmax = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
        }
        // positive-pair logit z_ii (natural units) when this chunk crosses the label diagonal
        if (__any_sync(0xffffffffu, row_valid && label >= col0 && label < col0 + 32)) {
          const int dj = label - col0;
          float dv = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) dv = (j == dj) ? __uint_as_float(v[j]) : dv;
          if (row_valid && dj >= 0 && dj < 32) diag[row] = dv * scale;
        }
        const float c_ref = warp_max_f(cmax);
        if (c_ref == NEG_INF) {  // uniform: no valid row in this warp
          sts_f32x2(sdst, NEG_INF, 0.f);
          continue;
        }
        float e[32];
        float rs;
        if (interior) {
          const uint64_t sc2 = f2_pack(s2, s2), nc2 = f2_pack(-c_ref, -c_ref);
          uint64_t acc2[2] = {0ull, 0ull};
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float lo, hi;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nc2), lo, hi);
            e[2 * j] = fast_exp2(lo);
            e[2 * j + 1] = fast_exp2(hi);
            acc2[j & 1] = f2_add(acc2[j & 1], f2_pack(e[2 * j], e[2 * j + 1]));
          }
          float lo, hi;
          f2_unpack(f2_add(acc2[0], acc2[1]), lo, hi);
          rs = lo + hi;
        } else {
          float rsp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = fast_exp2(z[j] - c_ref);
            rsp[j & 3] += e[j];
          }
          rs = (rsp[0] + rsp[1]) + (rsp[2] + rsp[3]);
        }
        if (cmax > NEG_INF) {  // row update (online softmax over the chunks / tiles this thread sees)
          const float m_new = fmaxf(m_row, cmax);
          float add;
          if (c_ref - cmax > 100.f) {  // this row sits far below its neighbours: exact path
            add = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) add += fast_exp2((interior ? __uint_as_float(v[j]) * s2 : z[j]) - m_new);
          } else {
            add = rs * fast_exp2(c_ref - m_new);
          }
          l_row = fmaf(l_row, fast_exp2(m_row - m_new), add);
          m_row = m_new;
        }
        float cs = warp_column_sums(e, lane, tscr);  // lane j: sum over the warp's 32 rows of column col0 + j
        float cref_out = c_ref;
        unsigned need = __ballot_sync(0xffffffffu, cs < 7.8886e-31f /* 2^-100 */ && col0 + (int)lane < n_cols);
        while (need) {  // columns far below the sub-block maximum: exact path, one column at a time
          const int jj = __ffs(need) - 1;
          need &= need - 1;
          uint32_t r;
          tmem_ld_x1(taddr + c * 32 + jj, r);
          tmem_ld_wait();
          const float zz = row_valid ? __uint_as_float(r) * s2 : NEG_INF;
          const float mm = warp_max_f(zz);
          const float ss = warp_sum_f(mm > NEG_INF ? fast_exp2(zz - mm) : 0.f);
          if ((int)lane == jj) {
            cref_out = mm;
            cs = ss;
          }
        }
        sts_f32x2(sdst, cref_out, cs);
      }
      // accumulator fully consumed: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cx.release_accumulator(acc);
      if (row_valid) ws_row[static_cast<long long>(tile_n * 2 + grp) * n_loc + row] = make_float2(m_row, l_row);
      named_bar_sync(bar_id, GEMM_GROUP_THREADS);
      {  // merge the four row-quadrants of this group's 128 columns
        const int colt = grp * 128 + et;
        const int col = ti.n0 + colt;
        float2 p[4];
        float M = NEG_INF;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          p[q] = lds_f32x2(scol + ((q * CL_BN + colt) << 3));
          M = fmaxf(M, p[q].x);
        }
        float S = 0.f;
        if (M > NEG_INF) {
#pragma unroll
          for (int q = 0; q < 4; ++q) S = fmaf(p[q].y, fast_exp2(p[q].x - M), S);
        }
        // (PAIR: the second CTA of the last pair may sit entirely past the last row; it has no slot in ws_col)
        if (ti.m0 < n_loc && col < n_cols) ws_col[static_cast<long long>(tile_m) * n_all + col_offset + col] = make_float2(M, S);
      }
      named_bar_sync(bar_id, GEMM_GROUP_THREADS);
    }
  }
  gemm_teardown(cx, tmem_base);
}

// Merge per-tile partials. rows: 2*tiles_n partials each; columns: tiles_m partials each. log2 domain in, natural out.
__global__ void __launch_bounds__(256)
clip_loss_finalize_kernel(const float2* __restrict__ ws_row, const float2* __restrict__ ws_col, int n_loc, int n_all,
                          int row_parts, int col_parts, float* __restrict__ row_lse, float* __restrict__ col_max,
                          float* __restrict__ col_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_loc) {
    float M = NEG_INF;
    for (int p = 0; p < row_parts; ++p) M = fmaxf(M, ws_row[static_cast<long long>(p) * n_loc + i].x);
    float S = 0.f;
    for (int p = 0; p < row_parts; ++p) {
      const float2 v = ws_row[static_cast<long long>(p) * n_loc + i];
      if (v.x > NEG_INF) S = fmaf(v.y, exp2f(v.x - M), S);
    }
    row_lse[i] = (M + log2f(S)) * LN2;
  } else if (i - n_loc < n_all) {
    const int j = i - n_loc;
    float M = NEG_INF;
    for (int p = 0; p < col_parts; ++p) M = fmaxf(M, ws_col[static_cast<long long>(p) * n_all + j].x);
    float S = 0.f;
    for (int p = 0; p < col_parts; ++p) {
      const float2 v = ws_col[static_cast<long long>(p) * n_all + j];
      if (v.x > NEG_INF) S = fmaf(v.y, exp2f(v.x - M), S);
    }
    col_max[j] = M * LN2;
    col_sum[j] = S;
  }
}

// col_lse[j] = log sum_w exp(col_max[w][j]) * col_sum[w][j]  over the W ranks' partial statistics.
__global__ void __launch_bounds__(256)
clip_loss_combine_kernel(const float* __restrict__ col_max, const float* __restrict__ col_sum, int parts, int n_all,
                         float* __restrict__ col_lse) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_all) return;
  float M = NEG_INF;
  for (int w = 0; w < parts; ++w) M = fmaxf(M, col_max[static_cast<long long>(w) * n_all + j]);
  float S = 0.f;
  for (int w = 0; w < parts; ++w) {
    const float m = col_max[static_cast<long long>(w) * n_all + j];
    if (m > NEG_INF) S = fmaf(col_sum[static_cast<long long>(w) * n_all + j], expf(m - M), S);
  }
  col_lse[j] = M + logf(S);
}

// out[0] = 0.5/n * [ sum_i (row_lse_i - z_ii) + sum_i (col_lse_{label(i)} - z_ii) ]   (loss.py:126-129), out[1], out[2] the two sums
__global__ void __launch_bounds__(1024)
clip_loss_value_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse, const float* __restrict__ diag,
                       int n_loc, int row_offset, float* __restrict__ out) {
  __shared__ float red[2][32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
    const float d = diag[i];
    a += row_lse[i] - d;
    b += col_lse[row_offset + i] - d;
  }
  a = warp_sum_f(a);
  b = warp_sum_f(b);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = warp_sum_f(red[0][threadIdx.x]);
    b = warp_sum_f(red[1][threadIdx.x]);
    if (threadIdx.x == 0) {
      out[0] = 0.5f * (a + b) / static_cast<float>(n_loc);
      out[1] = a;
      out[2] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: G = dL/dz
template <bool PAIR>
using GradLayout = GemmSmemLayout<CL_BN, 2 * GEMM_BM * 128, 2 * CL_BN * 4, PAIR>;

template <bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
clip_loss_grad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmG, const float* __restrict__ row_lse,
                      const float* __restrict__ col_lse, float* __restrict__ d_scale, int n_loc, int n_all, int E,
                      int row_offset, const float* __restrict__ scale_ptr, const float* __restrict__ gout_ptr, float w_row,
                      float w_col) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using L = GradLayout<PAIR>;
  GemmCtx<CL_BN, L> cx(smem_raw);
  const uint32_t tmem_base = gemm_prologue(cx, &tmA, &tmB, &tmG);
  const int warp = threadIdx.x >> 5;

  if (warp == 0) {
    if (elect_one()) gemm_producer<CL_BN, false, false>(cx, &tmA, &tmB, n_loc, n_all, E);
  } else if (warp == 1) {
    if (elect_one()) gemm_mma_issuer<CL_BN, false, false>(cx, tmem_base, n_loc, n_all, E);
  } else if (warp >= GEMM_CTRL_WARPS) {
    const int ew = warp - GEMM_CTRL_WARPS;
    const int grp = ew >> 2;
    const int quad = ew & 3;
    const uint32_t lane = lane_id();
    const int et = quad * 32 + lane;
    const bool leader = et == 0;
    const uint32_t bar_id = 1 + grp;
    const uint32_t sbuf = smem_u32(cx.c_stage(grp));
    const uint32_t clse_base = smem_u32(cx.epi_scratch());
    const float s2 = __ldg(scale_ptr) * LOG2E;
    if (gout_ptr != nullptr) {   // upstream dL (a device scalar) folded into the weights
      const float go = __ldg(gout_ptr);
      w_row *= go;
      w_col *= go;
    }
    const float w_diag = w_row + w_col;
    float ds = 0.f;  // sum G_ij * acc_ij  (= sum G * z / scale)
    GemmSched sched(n_loc, n_all, CL_BN, E, 1, PAIR, cx.rank);
    int it = 0;
    for (int t = cx.first; t < sched.total; t += cx.stride, ++it) {
      const GemmTileInfo ti = sched.tile(t, CL_BN);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const uint32_t clse_s = clse_base + (it & 1) * (CL_BN * 4);
      {  // column LSEs of this tile (log2 units) -> smem; +inf for columns past the end (their G is 0)
        const int j = threadIdx.x - 32 * GEMM_CTRL_WARPS;
        const int n = ti.n0 + j;
        sts_f32(clse_s + j * 4, n < n_all ? col_lse[n] * LOG2E : INFINITY);
      }
      const int row = ti.m0 + et;
      const float rl2 = row < n_loc ? row_lse[row] * LOG2E : INFINITY;
      const int label = row_offset + row;
      named_bar_sync(3, GEMM_EPI_THREADS);
      mbar_wait(&cx.tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * CL_BN + grp * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col_in_tile = grp * 128 + c * 64;
        const int col0 = ti.n0 + col_in_tile;
        const bool live = col0 < n_all;
        uint32_t v[64];
        if (live) {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
          tmem_ld_x32(taddr + c * 64, v0);
          tmem_ld_x32(taddr + c * 64 + 32, v1);
          tmem_ld_wait();
        }
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) cx.release_accumulator(acc);
        }
        if (!live) continue;
        float g[64];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 cl = lds_f32x4(clse_s + (col_in_tile + 4 * j) * 4);
          const float cls[4] = {cl.x, cl.y, cl.z, cl.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = __uint_as_float(v[4 * j + q]);
            const float z2 = a * s2;
            g[4 * j + q] = fmaf(w_row, fast_exp2(z2 - rl2), w_col * fast_exp2(z2 - cls[q]));
          }
        }
        if (__any_sync(0xffffffffu, label >= col0 && label < col0 + 64)) {
          const int dj = label - col0;
#pragma unroll
          for (int j = 0; j < 64; ++j) g[j] = (j == dj) ? g[j] - w_diag : g[j];
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) ds = fmaf(g[j], __uint_as_float(v[j]), ds);
        // (packed fp32 pairs in this epilogue were measured: 168 registers + spills, 1.45 -> 1.75 ms)
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(bar_id, GEMM_GROUP_THREADS);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(sbuf + sw128_offset(et, j), make_uint4(pack_bf16x2(g[8 * j], g[8 * j + 1]), pack_bf16x2(g[8 * j + 2], g[8 * j + 3]),
                                                        pack_bf16x2(g[8 * j + 4], g[8 * j + 5]), pack_bf16x2(g[8 * j + 6], g[8 * j + 7])));
        fence_proxy_async_smem();
        named_bar_sync(bar_id, GEMM_GROUP_THREADS);
        if (leader) {
          tma_store_2d(&tmG, cx.c_stage(grp), col0, ti.m0);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
    ds = warp_sum_f(ds);
    if (lane == 0 && d_scale != nullptr) atomicAdd(d_scale, ds);
  }
  gemm_teardown(cx, tmem_base);
}

}  // namespace ovk

using namespace ovk;

// launch with (PAIR) a 2-CTA cluster per SM pair, or one CTA per SM
template <bool PAIR, class Kern, class... Args>
static int launch_loss_kernel(Kern kern, int smem_bytes, int tiles, cudaStream_t stream, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(clip_loss): %s", cudaGetErrorString(e));
  const int units = PAIR ? num_sms() / 2 : num_sms();
  const int n = tiles < units ? tiles : units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? 2 * n : n, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaLaunchKernelEx(clip_loss): %s", cudaGetErrorString(e));
  return check_launch("clip_loss kernel");
}

static inline int cl_tiles_m(int n_loc) { return (n_loc + GEMM_BM - 1) / GEMM_BM; }
static inline int cl_tiles_n(int n_all) { return (n_all + CL_BN - 1) / CL_BN; }

extern "C" long long ovk_clip_loss_workspace_floats(int n_loc, int n_all) {
  if (n_loc <= 0 || n_all <= 0) return 0;
  return 2LL * (2LL * cl_tiles_n(n_all) * n_loc + static_cast<long long>(cl_tiles_m(n_loc)) * n_all);
}

extern "C" int ovk_clip_loss_fwd(const void* a_loc, const void* b_rows, int n_loc, int n_cols, int col_offset, int n_all,
                                 int E, int row_offset, const float* scale_dev, float* diag, float* workspace, void* stream) {
  if (n_loc <= 0 || n_cols <= 0 || n_all <= 0 || E <= 0) return set_error(OVK_ERR_SHAPE, "clip_loss_fwd: empty problem");
  if (E % 8) return set_error(OVK_ERR_ALIGN, "clip_loss_fwd: E must be a multiple of 8");
  if (col_offset < 0 || col_offset + n_cols > n_all || (col_offset % CL_BN))
    return set_error(OVK_ERR_SHAPE, "clip_loss_fwd: column window [%d, %d) must start on a multiple of %d inside [0, %d)", col_offset,
                     col_offset + n_cols, CL_BN, n_all);
  if (col_offset + n_cols < n_all && (n_cols % CL_BN))
    return set_error(OVK_ERR_SHAPE, "clip_loss_fwd: only the last column window may end off a multiple of %d", CL_BN);
  if (row_offset < 0 || row_offset + n_loc > n_all)
    return set_error(OVK_ERR_SHAPE, "clip_loss_fwd: rows [%d, %d) have no matching columns in [0, %d)", row_offset, row_offset + n_loc, n_all);
  if (!workspace || !scale_dev) return set_error(OVK_ERR_SHAPE, "clip_loss_fwd: workspace and scale_dev are required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, a_loc, E, n_loc, E, GEMM_BK, GEMM_BM))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, b_rows, E, n_cols, E, GEMM_BK, n_loc >= 512 ? CL_BN / 2 : CL_BN))) return rc;
  const int tm = cl_tiles_m(n_loc), tn_all = cl_tiles_n(n_all), tn = cl_tiles_n(n_cols);
  float2* ws_row = reinterpret_cast<float2*>(workspace);
  float2* ws_col = ws_row + 2LL * tn_all * n_loc;
  if (n_loc >= 512)   // CTA pairs (256-row tiles)
    return launch_loss_kernel<true>(clip_loss_fwd_kernel<true>, FwdLayout<true>::DYN_BYTES, ((n_loc + 255) / 256) * tn, s, tmA, tmB,
                                    ws_row, ws_col, diag, n_loc, n_cols, col_offset, n_all, E, row_offset, scale_dev);
  return launch_loss_kernel<false>(clip_loss_fwd_kernel<false>, FwdLayout<false>::DYN_BYTES, tm * tn, s, tmA, tmB, ws_row, ws_col,
                                   diag, n_loc, n_cols, col_offset, n_all, E, row_offset, scale_dev);
}

extern "C" int ovk_clip_loss_finalize(const float* workspace, int n_loc, int n_all, float* row_lse, float* col_max,
                                      float* col_sum, void* stream) {
  if (n_loc <= 0 || n_all <= 0) return set_error(OVK_ERR_SHAPE, "clip_loss_finalize: empty problem");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int tm = cl_tiles_m(n_loc), tn = cl_tiles_n(n_all);
  const float2* ws_row = reinterpret_cast<const float2*>(workspace);
  const float2* ws_col = ws_row + 2LL * tn * n_loc;
  const int total = n_loc + n_all;
  clip_loss_finalize_kernel<<<(total + 255) / 256, 256, 0, s>>>(ws_row, ws_col, n_loc, n_all, 2 * tn, tm, row_lse, col_max, col_sum);
  return check_launch("clip_loss_finalize_kernel");
}

extern "C" int ovk_clip_loss_combine(const float* col_max_parts, const float* col_sum_parts, int parts, int n_all,
                                     float* col_lse, void* stream) {
  if (parts <= 0 || n_all <= 0) return set_error(OVK_ERR_SHAPE, "clip_loss_combine: empty problem");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  clip_loss_combine_kernel<<<(n_all + 255) / 256, 256, 0, s>>>(col_max_parts, col_sum_parts, parts, n_all, col_lse);
  return check_launch("clip_loss_combine_kernel");
}

extern "C" int ovk_clip_loss_value(const float* row_lse, const float* col_lse, const float* diag, int n_loc,
                                   int row_offset, float* out3, void* stream) {
  if (n_loc <= 0) return set_error(OVK_ERR_SHAPE, "clip_loss_value: empty problem");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  clip_loss_value_kernel<<<1, 1024, 0, s>>>(row_lse, col_lse, diag, n_loc, row_offset, out3);
  return check_launch("clip_loss_value_kernel");
}

extern "C" int ovk_clip_loss_grad_logits(const void* img_loc, const void* txt_all, int n_loc, int n_all, int E,
                                         int row_offset, const float* scale_dev, const float* row_lse, const float* col_lse,
                                         float w_row, float w_col, const float* grad_out_dev, void* G, long long ldg,
                                         float* d_scale_partial, void* stream) {
  if (n_loc <= 0 || n_all <= 0 || E <= 0) return set_error(OVK_ERR_SHAPE, "clip_loss_grad: empty problem");
  if ((E % 8) || (ldg % 8)) return set_error(OVK_ERR_ALIGN, "clip_loss_grad: E and ldg must be multiples of 8");
  if (!scale_dev) return set_error(OVK_ERR_SHAPE, "clip_loss_grad: scale_dev is required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB, tmG;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, img_loc, E, n_loc, E, GEMM_BK, GEMM_BM))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, txt_all, E, n_all, E, GEMM_BK, n_loc >= 512 ? CL_BN / 2 : CL_BN))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmG, G, n_all, n_loc, ldg, 64, GEMM_BM))) return rc;
  if (n_loc >= 512)
    return launch_loss_kernel<true>(clip_loss_grad_kernel<true>, GradLayout<true>::DYN_BYTES,
                                    ((n_loc + 255) / 256) * cl_tiles_n(n_all), s, tmA, tmB, tmG, row_lse, col_lse,
                                    d_scale_partial, n_loc, n_all, E, row_offset, scale_dev, grad_out_dev, w_row, w_col);
  return launch_loss_kernel<false>(clip_loss_grad_kernel<false>, GradLayout<false>::DYN_BYTES,
                                   cl_tiles_m(n_loc) * cl_tiles_n(n_all), s, tmA, tmB, tmG, row_lse, col_lse, d_scale_partial,
                                   n_loc, n_all, E, row_offset, scale_dev, grad_out_dev, w_row, w_col);
}
