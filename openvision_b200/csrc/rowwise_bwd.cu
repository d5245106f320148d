// Bandwidth-bound backward kernels of the image tower: column sums (bias gradients, positional-embedding gradient),
// L2-normalise backward, pooling backward, col2im (gradient w.r.t. the input image through the patch embedding).
#include "act.cuh"
#include "host_utils.h"
#include "ptx.cuh"

namespace ovk {

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8b(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] += sum_r x[r, c].  Bias gradients of F.linear (db = sum over tokens of dY) and, with rows = batch and
// cols = L*D, the positional-embedding gradient (sum over images of dX).  block (32, 8): lane <-> one 16-byte column
// vector, 8 row groups reduced through smem, one atomicAdd per column per block.
constexpr int CS_TY = 8;
__global__ void __launch_bounds__(32 * CS_TY) colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int rows,
                                                             int cols, int rows_per_block, float* __restrict__ out) {
  __shared__ float red[CS_TY][32][8];
  const int v = blockIdx.y * 32 + threadIdx.x;
  const int nvec = cols >> 3;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v < nvec) {
    for (int r = r0 + threadIdx.y; r < r1; r += CS_TY) {
      float f[8];
      unpack8b(*reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ldx + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && v < nvec) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < CS_TY; ++t) s += red[t][threadIdx.x][j];
      atomicAdd(&out[v * 8 + j], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ L2-normalise bwd
// y = x / max(||x||, eps)  ->  dx = (dy - y (y . dy)) / max(||x||, eps)   (for ||x|| > eps).  One warp per row.
template <bool DY_F32>
__global__ void __launch_bounds__(256) l2_normalize_bwd_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ dy,
                                                               __nv_bfloat16* __restrict__ dx, int rows, int E, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = E >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * E);
  float ss = 0.f, dot = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8], g[8];
    unpack8b(xr[v], f);
    if (DY_F32) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + static_cast<long long>(row) * E) + 2 * v;
      const float4 a = p[0], b = p[1];
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
    } else {
      unpack8b(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dy) + static_cast<long long>(row) * E)[v], g);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss = fmaf(f[j], f[j], ss);
      dot = fmaf(f[j], g[j], dot);
    }
  }
  ss = warp_sum_b(ss);
  dot = warp_sum_b(dot);
  const float nrm = sqrtf(ss);
  const bool clamped = nrm < eps;
  const float inv = 1.f / fmaxf(nrm, eps);
  const float k = clamped ? 0.f : dot * inv * inv * inv;  // (x . dy) / ||x||^3
  for (int v = lane; v < nvec; v += 32) {
    float f[8], g[8], o[8];
    unpack8b(xr[v], f);
    if (DY_F32) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + static_cast<long long>(row) * E) + 2 * v;
      const float4 a = p[0], b = p[1];
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
    } else {
      unpack8b(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dy) + static_cast<long long>(row) * E)[v], g);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(g[j], inv, -f[j] * k);
    reinterpret_cast<uint4*>(dx + static_cast<long long>(row) * E)[v] = pack8b(o);
  }
}

// ------------------------------------------------------------------------------------------------ pooling bwd
// 'avg': dx[b, 0, :] = 0, dx[b, l >= 1, :] = dpooled[b, :] / (L - 1);  'tok': dx[b, 0, :] = dpooled[b, :], rest 0.
__global__ void __launch_bounds__(256) pool_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dpooled,
                                                              __nv_bfloat16* __restrict__ dx, int B, int L, int D, int mode) {
  const int dv = D >> 3;
  const long long total = static_cast<long long>(B) * L * dv;
  const float inv = 1.f / static_cast<float>(L > 1 ? L - 1 : 1);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % dv);
    const long long bl = idx / dv;
    const int l = static_cast<int>(bl % L);
    const long long b = bl / L;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (mode == 1 ? (l == 0) : (l != 0)) {
      float f[8];
      unpack8b(*reinterpret_cast<const uint4*>(dpooled + b * D + v * 8), f);
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= inv;
      }
      o = pack8b(f);
    }
    *reinterpret_cast<uint4*>(dx + bl * D + v * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------ col2im
// Inverse of im2col for kernel == stride (every pixel belongs to exactly one patch column):
// dimages[b, c, y, x] = dcols[b*(N+lead) + lead + (y/P)*gw + x/P, (c*P + y%P)*P + x%P].
template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dcols, long long ldc,
                                                     T* __restrict__ dimg, int B, int H, int W, int P, int lead) {
  const int gw = W / P, gh = H / P;
  const long long total = static_cast<long long>(B) * 3 * H * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>((idx / (static_cast<long long>(W) * H)) % 3);
    const long long b = idx / (3LL * W * H);
    const long long r = b * (gh * gw + lead) + lead + (y / P) * gw + x / P;
    const int k = (c * P + y % P) * P + x % P;
    dimg[idx] = static_cast<T>(__bfloat162float(dcols[r * ldc + k]));
  }
}

// ------------------------------------------------------------------------------------------------ activation modules
// nn.GELU / QuickGELU called as a MODULE (the ov-* scripts hook its output, cliptoolsoptimized.py:1149-1164): same
// functional form as the fused GEMM epilogue.  mode 0: y = act(x);  mode 1: y = g * act'(x).
template <int MODE>
__global__ void __launch_bounds__(256) act_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g,
                                                  __nv_bfloat16* __restrict__ y, long long nvec, ActCoef k) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], o[8];
    unpack8b(reinterpret_cast<const uint4*>(x)[idx], f);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_fwd(f[j], k);
    } else {
      float gg[8];
      unpack8b(reinterpret_cast<const uint4*>(g)[idx], gg);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = gg[j] * act_bwd(f[j], k);
    }
    reinterpret_cast<uint4*>(y)[idx] = pack8b(o);
  }
}

// y = a + b (bf16): the residual add of transformer.py:263-264 when a hooked sub-module produced the branch (its output may
// have been replaced by the hook, so the add cannot ride in a GEMM epilogue).
__global__ void __launch_bounds__(256) add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                  __nv_bfloat16* __restrict__ y, long long nvec) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float fa[8], fb[8], o[8];
    unpack8b(reinterpret_cast<const uint4*>(a)[idx], fa);
    unpack8b(reinterpret_cast<const uint4*>(b)[idx], fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fa[j] + fb[j];
    reinterpret_cast<uint4*>(y)[idx] = pack8b(o);
  }
}

static int grid_for_b(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ovk

using namespace ovk;

// out[c] (+)= sum_r x[r][c] for an fp32 [rows, cols] matrix (the partial column sums a GEMM epilogue left behind)
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ x, int rows, int cols, int rows_per_block,
                                                         float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += x[static_cast<long long>(r) * cols + c];
    s1 += x[static_cast<long long>(r + 1) * cols + c];
    s2 += x[static_cast<long long>(r + 2) * cols + c];
    s3 += x[static_cast<long long>(r + 3) * cols + c];
  }
  for (; r < r1; ++r) s0 += x[static_cast<long long>(r) * cols + c];
  atomicAdd(out + c, (s0 + s1) + (s2 + s3));
}

extern "C" int ovk_colsum_f32(const float* x, int rows, int cols, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return set_error(OVK_ERR_SHAPE, "colsum_f32: empty input");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int gx = (cols + 255) / 256;
  int gy = (4 * num_sms() + gx - 1) / gx;
  if (gy > rows) gy = rows;
  const int rpb = (rows + gy - 1) / gy;
  gy = (rows + rpb - 1) / rpb;
  colsum_f32_kernel<<<dim3(gx, gy), 256, 0, s>>>(x, rows, cols, rpb, out);
  return check_launch("colsum_f32_kernel");
}

extern "C" int ovk_colsum_bf16(const void* x, long long ldx, int rows, int cols, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return set_error(OVK_ERR_SHAPE, "colsum: empty input");
  if ((cols % 8) || (ldx % 8)) return set_error(OVK_ERR_ALIGN, "colsum: cols and ldx must be multiples of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int gy = (cols / 8 + 31) / 32;
  int gx = (num_sms() * 8 + gy - 1) / gy;  // ~8 blocks per SM overall
  if (gx < 1) gx = 1;
  int rpb = (rows + gx - 1) / gx;
  if (rpb < CS_TY) rpb = CS_TY;
  gx = (rows + rpb - 1) / rpb;
  dim3 grid(gx, gy), block(32, CS_TY);
  colsum_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, cols, rpb, out);
  return check_launch("colsum_kernel");
}

extern "C" int ovk_l2_normalize_bwd(const void* x, const void* dy, int dy_is_f32, void* dx, int rows, int E, float eps,
                                    void* stream) {
  if (rows <= 0 || E <= 0 || E % 8) return set_error(OVK_ERR_SHAPE, "l2_normalize_bwd: bad shape");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  if (dy_is_f32)
    l2_normalize_bwd_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), dy,
                                                       reinterpret_cast<__nv_bfloat16*>(dx), rows, E, eps);
  else
    l2_normalize_bwd_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), dy,
                                                        reinterpret_cast<__nv_bfloat16*>(dx), rows, E, eps);
  return check_launch("l2_normalize_bwd_kernel");
}

extern "C" int ovk_pool_tokens_bwd(const void* dpooled, void* dx, int B, int L, int D, int mode, void* stream) {
  if (B <= 0 || L <= 0 || D <= 0 || D % 8) return set_error(OVK_ERR_SHAPE, "pool_tokens_bwd: bad shape");
  if (mode != 0 && mode != 1) return set_error(OVK_ERR_SHAPE, "pool_tokens_bwd: mode must be 0 (avg) or 1 (tok)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * L * (D / 8);
  pool_tokens_bwd_kernel<<<grid_for_b(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dpooled),
                                                                reinterpret_cast<__nv_bfloat16*>(dx), B, L, D, mode);
  return check_launch("pool_tokens_bwd_kernel");
}

extern "C" int ovk_col2im_patches(const void* dcols, long long ldc, void* dimages, int img_is_f32, int B, int H, int W,
                                  int P, int lead_rows, void* stream) {
  if (B <= 0 || P <= 0 || H % P || W % P) return set_error(OVK_ERR_SHAPE, "col2im: H=%d W=%d not divisible by P=%d", H, W, P);
  if (ldc < 3LL * P * P) return set_error(OVK_ERR_SHAPE, "col2im: ldc must be >= 3*P*P");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * 3 * H * W;
  if (img_is_f32)
    col2im_kernel<float><<<grid_for_b(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dcols), ldc,
                                                                reinterpret_cast<float*>(dimages), B, H, W, P, lead_rows);
  else
    col2im_kernel<__nv_bfloat16><<<grid_for_b(total, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dcols), ldc,
                                                                        reinterpret_cast<__nv_bfloat16*>(dimages), B, H, W, P,
                                                                        lead_rows);
  return check_launch("col2im_kernel");
}

extern "C" int ovk_act_fwd(const void* x, void* y, long long n, int act, void* stream) {
  if (n <= 0 || (n % 8)) return set_error(OVK_ERR_SHAPE, "act_fwd: element count must be a positive multiple of 8");
  if (!(act & OVK_EPI_ACT_MASK)) return set_error(OVK_ERR_SHAPE, "act_fwd: unknown activation id %d", act);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  act_kernel<0><<<grid_for_b(n / 8, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), nullptr,
                                                       reinterpret_cast<__nv_bfloat16*>(y), n / 8, act_coef(act & OVK_EPI_ACT_MASK));
  return check_launch("act_kernel<fwd>");
}

extern "C" int ovk_act_bwd(const void* x, const void* dy, void* dx, long long n, int act, void* stream) {
  if (n <= 0 || (n % 8)) return set_error(OVK_ERR_SHAPE, "act_bwd: element count must be a positive multiple of 8");
  if (!(act & OVK_EPI_ACT_MASK)) return set_error(OVK_ERR_SHAPE, "act_bwd: unknown activation id %d", act);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  act_kernel<1><<<grid_for_b(n / 8, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                       reinterpret_cast<const __nv_bfloat16*>(dy),
                                                       reinterpret_cast<__nv_bfloat16*>(dx), n / 8, act_coef(act & OVK_EPI_ACT_MASK));
  return check_launch("act_kernel<bwd>");
}

extern "C" int ovk_add_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
  if (n <= 0 || (n % 8)) return set_error(OVK_ERR_SHAPE, "add: element count must be a positive multiple of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  add_kernel<<<grid_for_b(n / 8, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
                                                   reinterpret_cast<__nv_bfloat16*>(y), n / 8);
  return check_launch("add_kernel");
}
