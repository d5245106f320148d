// Bandwidth-bound kernels of the image tower: LayerNorm (fwd/bwd), im2col for the patch-embedding GEMM,
// cls/pos assembly, token pooling and L2 normalisation.  All use 128-bit vector loads/stores, one warp per row,
// warp-shuffle reductions and fp32 statistics.
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// transformer.py:15-30: y = (x - mean) / sqrt(var + eps) * gamma + beta, biased variance, statistics in fp32.
template <int MAXV>  // vectors (of 8 bf16) per lane; D <= MAXV * 256
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                            __nv_bfloat16* __restrict__ y, long long ldy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int rows, int D, float eps) {
  // Persistent: a warp walks rows with stride (warps in the grid) and keeps the gamma / beta of its own columns in
  // registers for all of them (re-reading them per row costs four L1 loads per 16 bytes of data, which is what limited
  // the first version).  MAXV <= 4: RPW = 2 rows are loaded before either is touched (more bytes in flight per SM).
  constexpr int RPW = MAXV >= 8 ? 1 : 2;   // (host: rows per block = 8 * RPW)
  constexpr bool GB_IN_REGS = MAXV <= 4;   // 16 * MAXV registers
  // rows of 1025..1536 elements (H/14: 1280): gamma / beta in registers would cost 80-96 registers and leave one CTA per SM
  // (measured 3.1 TB/s); they sit in shared memory instead (four conflict-free 16-byte loads per vector), wider rows read them
  // through L1 as before
  constexpr bool GB_IN_SMEM = MAXV == 5 || MAXV == 6;
  __shared__ float gb_s[GB_IN_SMEM ? 2 * MAXV * 256 : 1];
  if (GB_IN_SMEM) {
    for (int c = threadIdx.x; c < MAXV * 256; c += blockDim.x) {
      gb_s[c] = c < D ? gamma[c] : 0.f;
      gb_s[MAXV * 256 + c] = c < D ? beta[c] : 0.f;
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float4 g0[GB_IN_REGS ? MAXV : 1], g1[GB_IN_REGS ? MAXV : 1], b0[GB_IN_REGS ? MAXV : 1], b1[GB_IN_REGS ? MAXV : 1];
  if (GB_IN_REGS) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      const bool live = v < nvec;
      g0[i] = live ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      g1[i] = live ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      b0[i] = live ? __ldg(reinterpret_cast<const float4*>(beta) + 2 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      b1[i] = live ? __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int row0 = warp_g * RPW; row0 < rows; row0 += warps_total * RPW) {
    uint4 raw[RPW][MAXV];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row0 + rr) * ldx);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        raw[rr][i] = (v < nvec && row0 + rr < rows) ? xr[v] : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const int row = row0 + rr;
      if (row >= rows) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        float f[8];
        unpack8(raw[rr][i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
      }
      const float mean = warp_sum(s) / static_cast<float>(D);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (lane + i * 32 < nvec) {
          float f[8];
          unpack8(raw[rr][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - mean;
            q = fmaf(d, d, q);
          }
        }
      }
      const float var = warp_sum(q) / static_cast<float>(D);
      const float rstd = rsqrtf(var + eps);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
      }
      uint4* yr = reinterpret_cast<uint4*>(y + static_cast<long long>(row) * ldy);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
          float f[8];
          unpack8(raw[rr][i], f);
          float4 a0, a1, c0, c1;
          if (GB_IN_REGS) {
            a0 = g0[i]; a1 = g1[i]; c0 = b0[i]; c1 = b1[i];
          } else if (GB_IN_SMEM) {
            a0 = *reinterpret_cast<const float4*>(gb_s + 8 * v);
            a1 = *reinterpret_cast<const float4*>(gb_s + 8 * v + 4);
            c0 = *reinterpret_cast<const float4*>(gb_s + MAXV * 256 + 8 * v);
            c1 = *reinterpret_cast<const float4*>(gb_s + MAXV * 256 + 8 * v + 4);
          } else {
            a0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
            a1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
            c0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v);
            c1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + 1);
          }
          f[0] = fmaf((f[0] - mean) * rstd, a0.x, c0.x);
          f[1] = fmaf((f[1] - mean) * rstd, a0.y, c0.y);
          f[2] = fmaf((f[2] - mean) * rstd, a0.z, c0.z);
          f[3] = fmaf((f[3] - mean) * rstd, a0.w, c0.w);
          f[4] = fmaf((f[4] - mean) * rstd, a1.x, c1.x);
          f[5] = fmaf((f[5] - mean) * rstd, a1.y, c1.y);
          f[6] = fmaf((f[6] - mean) * rstd, a1.z, c1.z);
          f[7] = fmaf((f[7] - mean) * rstd, a1.w, c1.w);
          yr[v] = pack8(f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ row statistics
// (sum x, sum x^2) per row: the LayerNorm statistics in the form the LN-folding GEMM epilogue consumes
// (ovk_gemm_bf16_ln).  Inside the tower they come for free from the epilogue of the GEMM that writes x; this kernel
// seeds the first block.  One warp per row.
__global__ void __launch_bounds__(256) row_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                        float* __restrict__ stats, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
  float s1 = 0.f, s2 = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(xr[v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1 += f[j];
      s2 = fmaf(f[j], f[j], s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) *reinterpret_cast<float2*>(stats + 2 * static_cast<long long>(row)) = make_float2(s1, s2);
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// g = dy*gamma ; dx = rstd * (g - mean(g) - xhat * mean(g*xhat)) (+ dres) ; dgamma = sum_rows dy*xhat ; dbeta = sum_rows dy.
// Two streaming kernels instead of one register-heavy one (which ran at a single block per SM):
//   layernorm_bwd_dx_kernel   : one warp per row, like the forward (reads dy, x [, dres], writes dx)
//   layernorm_bwd_dgdb_kernel : column sums of dy*xhat and dy; block (32, 8): lane <-> one 16-byte column vector,
//                               8 row groups reduced through smem, one atomicAdd per column per block.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                               const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const __nv_bfloat16* dres, long long lddres,
                                                               __nv_bfloat16* dx, long long lddx, int rows, int D) {
  // persistent like the forward: a warp walks rows with a grid stride and keeps gamma of its columns in registers
  constexpr bool G_IN_REGS = MAXV <= 6;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  float4 gr0[G_IN_REGS ? MAXV : 1], gr1[G_IN_REGS ? MAXV : 1];
  if (G_IN_REGS) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      gr0[i] = v < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      gr1[i] = v < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<long long>(row) * lddy);
    const uint4* drr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<long long>(row) * lddres) : nullptr;
    uint4 rx[MAXV], rd[MAXV], rr[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      const bool ok = v < nvec;
      rx[i] = ok ? xr[v] : make_uint4(0, 0, 0, 0);
      rd[i] = ok ? dyr[v] : make_uint4(0, 0, 0, 0);
      rr[i] = (ok && drr) ? drr[v] : make_uint4(0, 0, 0, 0);
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float fx[8], fd[8];
        unpack8(rx[i], fx);
        unpack8(rd[i], fd);
        const float4 g0 = G_IN_REGS ? gr0[i] : __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
        const float4 g1 = G_IN_REGS ? gr1[i] : __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = fd[j] * gm[j];
          s1 += g;
          s2 = fmaf(g, (fx[j] - mu) * rs, s2);
        }
      }
    }
    const float m1 = warp_sum(s1) / static_cast<float>(D);
    const float m2 = warp_sum(s2) / static_cast<float>(D);
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<long long>(row) * lddx);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float fx[8], fd[8], fr[8], o[8];
        unpack8(rx[i], fx);
        unpack8(rd[i], fd);
        unpack8(rr[i], fr);
        const float4 g0 = G_IN_REGS ? gr0[i] : __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
        const float4 g1 = G_IN_REGS ? gr1[i] : __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(rs, fd[j] * gm[j] - m1 - (fx[j] - mu) * rs * m2, fr[j]);
        dxr[v] = pack8(o);
      }
    }
  }
}

// One-pass variant for D <= 1536 (MAXV <= 6): the same persistent row walk also accumulates this lane's columns of
// dgamma = sum dy * xhat and dbeta = sum dy in registers over all the rows of the warp, then the block reduces its eight
// warps through shared memory and issues one atomicAdd per column.  Saves the second pass over dy and x that the separate
// column-sum kernel needs (6 -> 4 passes over [rows, D] with the residual gradient).
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                  const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const __nv_bfloat16* dres, long long lddres,
                                                                  __nv_bfloat16* dx, long long lddx, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, int rows, int D) {
  extern __shared__ float red[];   // [8 warps][2][MAXV * 256] (+ [MAXV * 256] gamma when G_SMEM)
  // rows wider than 1024 (H/14: 1280): the 2 x 8 x MAXV column accumulators plus three row buffers leave no registers for
  // gamma, which then lives in shared memory (two 16-byte loads per vector and pass) instead
  constexpr bool G_SMEM = MAXV > 4;
  constexpr int DP = MAXV * 256;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nvec = D >> 3;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  // Packed fp32 pairs throughout (FFMA2 / FMUL2 / FADD2 on aligned register pairs): inside a training step the SMs run at
  // 1.2-1.3 GHz under the power cap and the 19 scalar instructions per element of the first version made this kernel
  // issue-bound there (0.62 of the HBM rate in the step against 0.87 alone); with xhat = x * rstd + (-mean * rstd) as one FMA
  // and the output as three, it is 11.
  uint64_t gm[G_SMEM ? 1 : MAXV][4], dg[MAXV][4], db[MAXV][4];
  float* gs = red + 16 * DP;
  if (G_SMEM) {
    for (int c = threadIdx.x; c < DP; c += blockDim.x) gs[c] = c < D ? gamma[c] : 0.f;
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (!G_SMEM) {
      const int v = lane + i * 32;
      const float4 g0 = v < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g1 = v < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      gm[i][0] = f2_pack(g0.x, g0.y); gm[i][1] = f2_pack(g0.z, g0.w);
      gm[i][2] = f2_pack(g1.x, g1.y); gm[i][3] = f2_pack(g1.z, g1.w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) dg[i][j] = db[i][j] = f2_pack(0.f, 0.f);
  }
  auto gamma_of = [&](int i, uint64_t (&g)[4]) {
    if (G_SMEM) {
      const float4 a = *reinterpret_cast<const float4*>(gs + (lane + i * 32) * 8);
      const float4 b = *reinterpret_cast<const float4*>(gs + (lane + i * 32) * 8 + 4);
      g[0] = f2_pack(a.x, a.y); g[1] = f2_pack(a.z, a.w); g[2] = f2_pack(b.x, b.y); g[3] = f2_pack(b.z, b.w);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = gm[G_SMEM ? 0 : i][j];
    }
  };
  // the four 32-bit words of a 16-byte vector as fp32 pairs (elements 2j, 2j + 1)
  auto unpack4 = [](const uint4& u, uint64_t (&f)[4]) {
    f[0] = f2_pack(bf16_lo(u.x), bf16_hi(u.x));
    f[1] = f2_pack(bf16_lo(u.y), bf16_hi(u.y));
    f[2] = f2_pack(bf16_lo(u.z), bf16_hi(u.z));
    f[3] = f2_pack(bf16_lo(u.w), bf16_hi(u.w));
  };
  for (int row = blockIdx.x * (blockDim.x >> 5) + wib; row < rows; row += warps_total) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<long long>(row) * lddy);
    const uint4* drr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<long long>(row) * lddres) : nullptr;
    uint4 rx[MAXV], rd[MAXV], rr[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      const bool ok = v < nvec;
      rx[i] = ok ? xr[v] : make_uint4(0, 0, 0, 0);
      rd[i] = ok ? dyr[v] : make_uint4(0, 0, 0, 0);
      rr[i] = (ok && drr) ? drr[v] : make_uint4(0, 0, 0, 0);
    }
    const float mu = mean[row], rs = rstd[row];
    const uint64_t a2 = f2_pack(rs, rs), b2 = f2_pack(-mu * rs, -mu * rs);   // xhat = x * a + b
    uint64_t s1 = f2_pack(0.f, 0.f), s2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      uint64_t fx[4], fd[4], gv[4];
      unpack4(rx[i], fx);
      unpack4(rd[i], fd);
      gamma_of(i, gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t xh = f2_fma(fx[j], a2, b2);
        const uint64_t g = f2_mul(fd[j], gv[j]);
        s1 = f2_add(s1, g);
        s2 = f2_fma(g, xh, s2);
        dg[i][j] = f2_fma(fd[j], xh, dg[i][j]);
        db[i][j] = f2_add(db[i][j], fd[j]);
      }
    }
    float s1a, s1b, s2a, s2b;
    f2_unpack(s1, s1a, s1b);
    f2_unpack(s2, s2a, s2b);
    const float m1 = warp_sum(s1a + s1b) / static_cast<float>(D);
    const float m2 = warp_sum(s2a + s2b) / static_cast<float>(D);
    // dx = rstd * (dy * gamma - m1 - xhat * m2) + dres  =  (dy * gamma) * rstd + dres  +  xhat * (-rstd * m2)  +  (-rstd * m1)
    const uint64_t c1 = f2_pack(-rs * m1, -rs * m1), c2 = f2_pack(-rs * m2, -rs * m2);
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<long long>(row) * lddx);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        uint64_t fx[4], fd[4], fr[4], gv[4];
        unpack4(rx[i], fx);
        unpack4(rd[i], fd);
        unpack4(rr[i], fr);
        gamma_of(i, gv);
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t xh = f2_fma(fx[j], a2, b2);
          const uint64_t t = f2_fma(f2_mul(fd[j], gv[j]), a2, fr[j]);
          float o0, o1;
          f2_unpack(f2_add(f2_fma(xh, c2, t), c1), o0, o1);
          ow[j] = pack_bf16x2(o0, o1);
        }
        dxr[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
  // block reduction of the column sums, then one atomicAdd per column and block
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (lane + i * 32) * 8 + 2 * j;
      float g0, g1, b0, b1;
      f2_unpack(dg[i][j], g0, g1);
      f2_unpack(db[i][j], b0, b1);
      red[(wib * 2 + 0) * DP + c] = g0;
      red[(wib * 2 + 0) * DP + c + 1] = g1;
      red[(wib * 2 + 1) * DP + c] = b0;
      red[(wib * 2 + 1) * DP + c + 1] = b1;
    }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[(w * 2 + 0) * DP + c];
      b += red[(w * 2 + 1) * DP + c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

constexpr int LNB_TY = 8;
__global__ void __launch_bounds__(32 * LNB_TY) layernorm_bwd_dgdb_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                          const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                          const float* __restrict__ mean,
                                                                          const float* __restrict__ rstd,
                                                                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                          int rows, int D, int rows_per_block) {
  __shared__ float red[2][LNB_TY][32][8];
  const int v = blockIdx.y * 32 + threadIdx.x;
  const int nvec = D >> 3;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v < nvec) {
    for (int r = r0 + threadIdx.y; r < r1; r += LNB_TY) {
      float fx[8], fd[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ldx + v * 8), fx);
      unpack8(*reinterpret_cast<const uint4*>(dy + static_cast<long long>(r) * lddy + v * 8), fd);
      const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dg[j] = fmaf(fd[j], (fx[j] - mu) * rs, dg[j]);
        db[j] += fd[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.y][threadIdx.x][j] = dg[j];
    red[1][threadIdx.y][threadIdx.x][j] = db[j];
  }
  __syncthreads();
  if (threadIdx.y < 2 && v < nvec) {
    float* dst = threadIdx.y == 0 ? dgamma : dbeta;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < LNB_TY; ++t) s += red[threadIdx.y][t][threadIdx.x][j];
      atomicAdd(&dst[v * 8 + j], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ im2col
// Row r = (b, gy, gx) of `cols` holds conv1's receptive field in the order of conv1.weight.reshape(D, 3*P*P):
// k = (c*P + ph)*P + pw  <->  images[b, c, gy*P + ph, gx*P + pw].   transformer.py:469,610-612.
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ cols,
                                                     long long ldc, int B, int H, int W, int P, int lead) {
  const int gh = H / P, gw = W / P;
  const int K = 3 * P * P;
  const int vec_per_row = static_cast<int>(ldc >> 3);
  const int rpi = gh * gw + lead;  // rows per image; the `lead` rows in front of each image's patches are zero
  const long long total = static_cast<long long>(B) * rpi * vec_per_row;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec_per_row);
    const long long r = idx / vec_per_row;
    const int rr = static_cast<int>(r % rpi) - lead;
    const int b = static_cast<int>(r / rpi);
    const int gx = rr % gw;
    const int gy = rr / gw;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      float val = 0.f;
      if (k < K && rr >= 0) {
        const int pw = k % P;
        const int ph = (k / P) % P;
        const int c = k / (P * P);
        const long long src = ((static_cast<long long>(b) * 3 + c) * H + (gy * P + ph)) * W + (gx * P + pw);
        val = static_cast<float>(img[src]);
      }
      f[j] = val;
    }
    *reinterpret_cast<uint4*>(cols + r * ldc + v * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ cls / pos assembly
// transformer.py:615-617: x = cat([class_embedding, patches]) + positional_embedding.
__global__ void __launch_bounds__(256) embed_assemble_kernel(const __nv_bfloat16* patch, const float* __restrict__ cls,
                                                             const float* __restrict__ pos, __nv_bfloat16* tokens,
                                                             int B, int N, int D, int lead) {
  const int L = N + 1;
  const int dv = D >> 3;
  const long long total = static_cast<long long>(B) * L * dv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % dv);
    const long long bl = idx / dv;
    const int l = static_cast<int>(bl % L);
    const long long b = bl / L;
    float f[8];
    if (l == 0) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls) + 2 * v);
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(cls) + 2 * v + 1);
      f[0] = c0.x; f[1] = c0.y; f[2] = c0.z; f[3] = c0.w; f[4] = c1.x; f[5] = c1.y; f[6] = c1.z; f[7] = c1.w;
    } else {
      unpack8(*reinterpret_cast<const uint4*>(patch + (b * (N + lead) + (l - 1 + lead)) * D + v * 8), f);
    }
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(l) * D) + 2 * v);
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(l) * D) + 2 * v + 1);
    f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
    *reinterpret_cast<uint4*>(tokens + bl * D + v * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ token pooling
// transformer.py:599-607: 'avg' -> mean over tokens 1..L-1 (cls excluded); 'tok' -> token 0.
// grid (B, ceil(D/8/32)); block (32, TY): lane <-> one 16-byte column vector, TY row groups reduced through smem.
constexpr int POOL_TY = 8;
__global__ void __launch_bounds__(32 * POOL_TY) pool_tokens_kernel(const __nv_bfloat16* __restrict__ x,
                                                                    __nv_bfloat16* __restrict__ pooled, int L, int D,
                                                                    int mode) {
  __shared__ float red[POOL_TY][32][8];
  const int b = blockIdx.x;
  const int v = blockIdx.y * 32 + threadIdx.x;
  const int dv = D >> 3;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * L * D;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v < dv) {
    if (mode == 1) {
      if (threadIdx.y == 0) unpack8(*reinterpret_cast<const uint4*>(xb + v * 8), acc);
    } else {
      for (int l = 1 + threadIdx.y; l < L; l += POOL_TY) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xb + static_cast<long long>(l) * D + v * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && v < dv) {
    float o[8];
    const float inv = mode == 1 ? 1.f : 1.f / static_cast<float>(L - 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < POOL_TY; ++t) s += red[t][threadIdx.x][j];
      o[j] = s * inv;
    }
    *reinterpret_cast<uint4*>(pooled + static_cast<long long>(b) * D + v * 8) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------ L2 normalise
// model.py:267,284: F.normalize(x, dim=-1) = x / max(||x||_2, eps).  One warp per row.
template <bool OUT_F32>
__global__ void __launch_bounds__(256) l2_normalize_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ y,
                                                           float* __restrict__ norms, int rows, int E, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = E >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * E);
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(xr[v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(f[j], f[j], s);
  }
  const float nrm = sqrtf(warp_sum(s));
  if (lane == 0 && norms) norms[row] = nrm;
  const float inv = 1.f / fmaxf(nrm, eps);
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(xr[v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    if (OUT_F32) {
      float4* yr = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + static_cast<long long>(row) * E);
      yr[2 * v] = make_float4(f[0], f[1], f[2], f[3]);
      yr[2 * v + 1] = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + static_cast<long long>(row) * E)[v] = pack8(f);
    }
  }
}

static int grid_for(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ovk

using namespace ovk;

extern "C" int ovk_layernorm_fwd(const void* x, long long ldx, void* y, long long ldy, const float* gamma,
                                 const float* beta, float* mean, float* rstd, int rows, int D, float eps, void* stream) {
  if (rows <= 0 || D <= 0) return set_error(OVK_ERR_SHAPE, "layernorm: empty input");
  if ((D % 8) || (ldx % 8) || (ldy % 8)) return set_error(OVK_ERR_ALIGN, "layernorm: D, ldx, ldy must be multiples of 8");
  if (D > 2048) return set_error(OVK_ERR_SHAPE, "layernorm: D=%d > 2048 not supported", D);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int rpb = 8 * (D > 1536 ? 1 : 2);  // 8 warps x rows per warp (layernorm_fwd_kernel::RPW)
  const int need = (rows + rpb - 1) / rpb;
  const int cap = 4 * num_sms();           // persistent: CTAs walk the rows with a grid stride
  const int grid = need < cap ? need : cap;
  auto xp = reinterpret_cast<const __nv_bfloat16*>(x);
  auto yp = reinterpret_cast<__nv_bfloat16*>(y);
  if (D <= 256) layernorm_fwd_kernel<1><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);
  else if (D <= 512) layernorm_fwd_kernel<2><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);
  else if (D <= 768) layernorm_fwd_kernel<3><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);   // B/16
  else if (D <= 1024) layernorm_fwd_kernel<4><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);
  else if (D <= 1280) layernorm_fwd_kernel<5><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);   // H/14
  else if (D <= 1536) layernorm_fwd_kernel<6><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);
  else layernorm_fwd_kernel<8><<<grid, 256, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, mean, rstd, rows, D, eps);
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int ovk_row_stats(const void* x, long long ldx, float* stats, int rows, int D, void* stream) {
  if (rows <= 0 || D <= 0) return set_error(OVK_ERR_SHAPE, "row_stats: empty input");
  if ((D % 8) || (ldx % 8)) return set_error(OVK_ERR_ALIGN, "row_stats: D and ldx must be multiples of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  row_stats_kernel<<<(rows + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, stats, rows, D);
  return check_launch("row_stats_kernel");
}

extern "C" int ovk_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* gamma,
                                 const float* mean, const float* rstd, const void* dres, long long lddres, void* dx,
                                 long long lddx, float* dgamma, float* dbeta, int rows, int D, void* stream) {
  if (rows <= 0 || D <= 0) return set_error(OVK_ERR_SHAPE, "layernorm_bwd: empty input");
  if ((D % 8) || (ldx % 8) || (lddy % 8) || (lddx % 8) || (dres && (lddres % 8)))
    return set_error(OVK_ERR_ALIGN, "layernorm_bwd: D and leading dimensions must be multiples of 8");
  if (D > 2048) return set_error(OVK_ERR_SHAPE, "layernorm_bwd: D=%d > 2048 not supported", D);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto xp = reinterpret_cast<const __nv_bfloat16*>(x);
  auto dxp = reinterpret_cast<__nv_bfloat16*>(dx);
  auto drp = reinterpret_cast<const __nv_bfloat16*>(dres);
  if (dgamma != nullptr && dbeta != nullptr && D <= 1536) {   // one pass: dx and the parameter gradients together
    const int need = (rows + 7) / 8;
    const int cap = (D <= 512 ? 2 : 1) * num_sms();   // D > 512: 213 registers per thread, one CTA per SM
    const int grid = need < cap ? need : cap;
    int rc;
#define OVK_LNB_FUSED(MV)                                                                                                     \
    {                                                                                                                         \
      const size_t sm = (8 * 2 + ((MV) > 4 ? 1 : 0)) * (MV) * 256 * sizeof(float);                                            \
      static PerDeviceOnce attr_##MV;                                                                                         \
      if (attr_##MV.need()) {                                                                                                 \
        cudaFuncSetAttribute(layernorm_bwd_fused_kernel<MV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)); \
        attr_##MV.done();                                                                                                     \
      }                                                                                                                       \
      layernorm_bwd_fused_kernel<MV><<<grid, 256, sm, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx,      \
                                                           dgamma, dbeta, rows, D);                                           \
    }
    if (D <= 256) OVK_LNB_FUSED(1)
    else if (D <= 512) OVK_LNB_FUSED(2)
    else if (D <= 768) OVK_LNB_FUSED(3)
    else if (D <= 1024) OVK_LNB_FUSED(4)
    else if (D <= 1280) OVK_LNB_FUSED(5)   // H/14: gamma in shared memory (see the kernel)
    else OVK_LNB_FUSED(6)
#undef OVK_LNB_FUSED
    rc = check_launch("layernorm_bwd_fused_kernel");
    return rc;
  }
  if (dgamma != nullptr && dbeta != nullptr) {   // parameter gradients first: dx may alias dy
    const int gy = (D / 8 + 31) / 32;
    int gx = (num_sms() * 8 + gy - 1) / gy;
    int rpb = (rows + gx - 1) / gx;
    if (rpb < LNB_TY) rpb = LNB_TY;
    gx = (rows + rpb - 1) / rpb;
    dim3 grid(gx, gy), block(32, LNB_TY);
    layernorm_bwd_dgdb_kernel<<<grid, block, 0, s>>>(dyp, lddy, xp, ldx, mean, rstd, dgamma, dbeta, rows, D, rpb);
    int rc = check_launch("layernorm_bwd_dgdb_kernel");
    if (rc) return rc;
  }
  const int need_dx = (rows + 7) / 8;
  const int grid = need_dx < 4 * num_sms() ? need_dx : 4 * num_sms();   // persistent: warps walk the rows with a grid stride
  if (D <= 256) layernorm_bwd_dx_kernel<1><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  else if (D <= 512) layernorm_bwd_dx_kernel<2><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  else if (D <= 768) layernorm_bwd_dx_kernel<3><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  else if (D <= 1024) layernorm_bwd_dx_kernel<4><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  else if (D <= 1536) layernorm_bwd_dx_kernel<6><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  else layernorm_bwd_dx_kernel<8><<<grid, 256, 0, s>>>(dyp, lddy, xp, ldx, gamma, mean, rstd, drp, lddres, dxp, lddx, rows, D);
  return check_launch("layernorm_bwd_dx_kernel");
}

extern "C" int ovk_im2col_patches(const void* images, int img_is_f32, void* cols, long long ldc, int B, int H, int W,
                                  int P, int lead_rows, void* stream) {
  if (B <= 0 || P <= 0 || H % P || W % P) return set_error(OVK_ERR_SHAPE, "im2col: H=%d W=%d not divisible by P=%d", H, W, P);
  if (ldc % 8 || ldc < 3LL * P * P) return set_error(OVK_ERR_ALIGN, "im2col: ldc must be a multiple of 8 and >= 3*P*P");
  if (lead_rows < 0 || lead_rows > 1) return set_error(OVK_ERR_SHAPE, "im2col: lead_rows must be 0 or 1");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * ((H / P) * (W / P) + lead_rows) * (ldc / 8);
  const int grid = grid_for(total, 256);
  if (img_is_f32)
    im2col_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(images),
                                              reinterpret_cast<__nv_bfloat16*>(cols), ldc, B, H, W, P, lead_rows);
  else
    im2col_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(images),
                                                      reinterpret_cast<__nv_bfloat16*>(cols), ldc, B, H, W, P, lead_rows);
  return check_launch("im2col_kernel");
}

extern "C" int ovk_embed_assemble(const void* patch, int patch_lead_rows, const float* cls, const float* pos,
                                  void* tokens, int B, int N, int D, void* stream) {
  if (B <= 0 || N <= 0 || D <= 0 || D % 8) return set_error(OVK_ERR_SHAPE, "embed_assemble: bad shape B=%d N=%d D=%d", B, N, D);
  if (patch_lead_rows < 0 || patch_lead_rows > 1) return set_error(OVK_ERR_SHAPE, "embed_assemble: lead rows must be 0 or 1");
  if (patch == tokens && patch_lead_rows != 1)
    return set_error(OVK_ERR_SHAPE, "embed_assemble: in-place operation needs the [B, N+1, D] patch layout");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * (N + 1) * (D / 8);
  embed_assemble_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(patch), cls, pos,
                                                             reinterpret_cast<__nv_bfloat16*>(tokens), B, N, D,
                                                             patch_lead_rows);
  return check_launch("embed_assemble_kernel");
}

extern "C" int ovk_pool_tokens(const void* x, void* pooled, int B, int L, int D, int mode, void* stream) {
  if (B <= 0 || L <= 0 || D <= 0 || D % 8) return set_error(OVK_ERR_SHAPE, "pool_tokens: bad shape");
  if (mode == 0 && L < 2) return set_error(OVK_ERR_SHAPE, "pool_tokens: avg pooling needs L >= 2");
  if (mode != 0 && mode != 1) return set_error(OVK_ERR_SHAPE, "pool_tokens: mode must be 0 (avg) or 1 (tok)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(B, (D / 8 + 31) / 32), block(32, POOL_TY);
  pool_tokens_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                            reinterpret_cast<__nv_bfloat16*>(pooled), L, D, mode);
  return check_launch("pool_tokens_kernel");
}

extern "C" int ovk_l2_normalize(const void* x, void* y, int y_is_f32, float* norms, int rows, int E, float eps,
                                void* stream) {
  if (rows <= 0 || E <= 0 || E % 8) return set_error(OVK_ERR_SHAPE, "l2_normalize: bad shape");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  if (y_is_f32)
    l2_normalize_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, norms, rows, E, eps);
  else
    l2_normalize_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, norms, rows, E, eps);
  return check_launch("l2_normalize_kernel");
}
