// ovk_patch_embed: conv1 of the image tower (open_clip/transformer.py:469,610-612: Conv2d(3 -> D, kernel = stride = P,
// bias=False)), the reshape / permute to tokens and `cat([class_embedding, x]) + positional_embedding` (:615-617) as ONE
// kernel: an im2col GEMM whose A operand is never materialised in memory.
//
//   warp 0   producer : TMA boxes of RAW pixels (fp32 or bf16, straight out of the NCHW image: for every patch row of the
//                       tile a [R image rows x W pixels] box of one channel) into a 2-4 slot staging ring (88 KB cut to the tile's size), and the weight
//                       k-blocks (bf16, K-major, SWIZZLE_128B) into a 2-stage B ring
//   warps 4-11 transform: two threads per patch of the 128-patch tile (one per half of the k-block's R image rows): read
//                       their R/2 x P pixels of the staged rows, convert to bf16 and write their part of row m of the K-major
//                       SWIZZLE_128B A tile (what a TMA load of an im2col matrix would have produced), zero-filling the
//                       padding columns.  (Four transform warps — one per scheduler — were the bottleneck of the first
//                       version: 1 240 clocks per k-block against 512 for its MMAs, ncu source view of profiles/r02 run d.)
//   warp 1   MMA      : tcgen05.mma 128 x 256 x 64 per k-block into one of two TMEM accumulators
//   warps 12-19 epilogue: TMEM -> registers, + positional-embedding row (bf16 table, class token folded into row 0), bf16,
//                       swizzled smem, 3-D TMA store into tokens[b, 1 + p, :]; the first tile of an image also writes the
//                       class-token row tokens[b, 0, :]
//
// K ordering (weights are packed to match, openvision_b200/transformer.py PatchEmbedConv.packed_weight_fused):
//   k' = ((c * PG + phg) * R + phl) * PW + pw,   PW = 16 (P <= 16) or 32, R = 64 / PW image rows per k-block,
//   PG = ceil(P / R);  columns with pw >= P or phg * R + phl >= P are zero in both operands.
// P = 14 and P = 16 both give K' = 768 = 12 k-blocks.
//
// HBM traffic: the image is read once (the four 256-column tiles of one 128-patch row block run on neighbouring CTAs and
// share it through L2), tokens are written once: no im2col buffer (the round-1 path wrote and re-read B * N * 3P^2 bf16).
#include <stdlib.h>

#include "gemm_core.cuh"
#include "host_utils.h"

namespace ovk {

constexpr int PE_BM = 128;
constexpr int PE_BN = 256;
constexpr int PE_THREADS = 640;
constexpr int PE_TR_WARPS = 8;        // transform warps (4 .. 11)
constexpr int PE_EPI_WARP0 = 4 + PE_TR_WARPS;
constexpr int PE_RAW_BYTES = 90112;   // staging ring: 88 KB, cut into 2-4 slots of the tile's size (largest tile: 43 KB)
constexpr int PE_RAW_MAX_STAGES = 4;
constexpr int PE_A_BYTES = PE_BM * 128;
constexpr int PE_B_BYTES = PE_BN * 128;
constexpr int PE_C_BYTES = PE_BM * 128;
constexpr int PE_OFF_A = 0;
constexpr int PE_OFF_C = PE_OFF_A + 2 * PE_A_BYTES;
constexpr int PE_OFF_B = PE_OFF_C + 2 * PE_C_BYTES;     // b_nst weight stages, then the pixel staging ring (both sized by the host)
constexpr int PE_DYN_BYTES = 2 * PE_B_BYTES + PE_RAW_BYTES;   // B ring + staging ring together: 152 KB
constexpr int PE_B_MAX_STAGES = 3;
constexpr int PE_OFF_BAR = PE_OFF_B + PE_DYN_BYTES;
constexpr int PE_NUM_BARS = 8 + 2 * PE_B_MAX_STAGES + 2 * PE_RAW_MAX_STAGES;
constexpr int PE_OFF_SLOT = PE_OFF_BAR + PE_NUM_BARS * 8;
constexpr int PE_SMEM = PE_OFF_SLOT + 16;
static_assert(PE_SMEM <= 232448, "exceeds 227 KB of dynamic shared memory");

struct PatchEmbedArgs {
  int B, N, gw, D, L;     // images, patches per image, patches per image row, width, tokens per image (N + 1)
  int nx, xbox;           // an image row is fetched as nx boxes of xbox pixels (TMA boxes are <= 256 wide)
  int row_bytes;          // bytes of one patch row inside a staged box: R image rows x xbox pixels
  int box_bytes;          // bytes between the staged boxes = npy_box patch rows (rounded up to 128)
  int box_bytes_tx;       // bytes one box transfers (unrounded)
  int raw_stage, raw_nst; // bytes per staging slot (multiple of 1024) and number of slots (2..4)
  int b_nst;              // weight k-block stages (2 or 3); the staging ring starts behind them
  int tiles_pi, tiles_n;  // 128-patch tiles per image, 256-column tiles
  const __nv_bfloat16* table;   // [L, D] positional embedding (+ class token in row 0) or null (plain conv tokens)
  __nv_bfloat16* out;           // [B, L, D]
};

template <int BYTES>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t* w) {
  if constexpr (BYTES == 16) {
    const uint4 v = lds128(addr);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else if constexpr (BYTES == 8) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
  } else {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(addr));
  }
}

template <int P, typename T>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmI, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const PatchEmbedArgs a) {
  constexpr int PW = P <= 16 ? 16 : 32;
  constexpr int R = 64 / PW;
  constexpr int PG = (P + R - 1) / R;
  constexpr int NKB = 3 * PG;
  constexpr int ES = sizeof(T);
  constexpr int PBYTES = P * ES;                                           // bytes of one patch row in the staged image row
  constexpr int VB = (PBYTES % 16 == 0) ? 16 : (PBYTES % 8 == 0) ? 8 : 4;  // widest aligned shared load
  constexpr int NV = PBYTES / VB;
  static_assert(P % 2 == 0 && P <= 32, "patch sizes: even, <= 32");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PE_OFF_BAR);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* tmem_full = bars + 4;     // [2]
  uint64_t* tmem_empty = bars + 6;    // [2]
  uint64_t* b_full = bars + 8;                            // [b_nst]
  uint64_t* b_empty = bars + 8 + PE_B_MAX_STAGES;         // [b_nst]
  uint64_t* raw_full = bars + 8 + 2 * PE_B_MAX_STAGES;                         // [raw_nst]
  uint64_t* raw_empty = bars + 8 + 2 * PE_B_MAX_STAGES + PE_RAW_MAX_STAGES;    // [raw_nst]
  uint8_t* const raw_base = smem + PE_OFF_B + a.b_nst * PE_B_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PE_OFF_SLOT);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmI);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < PE_RAW_MAX_STAGES; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], PE_TR_WARPS);
    }
    for (int i = 0; i < PE_B_MAX_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], PE_TR_WARPS);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = a.B * a.tiles_pi * a.tiles_n;
  // tile t -> (image b, 128-patch block mt, 256-column block nt); nt fastest: the column blocks of one patch block run at
  // the same time on neighbouring CTAs and read the same pixels out of L2
  auto decode = [&](int t, int& b, int& p0, int& n0) {
    const int nt = t % a.tiles_n;
    const int r = t / a.tiles_n;
    p0 = (r % a.tiles_pi) * PE_BM;
    b = r / a.tiles_pi;
    n0 = nt * PE_BN;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------------------------------ TMA producer

      int rs = 0, bs = 0;   // staging ring / weight ring: slot and phase
      uint32_t rph = 0, bph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int b, p0, n0;
        decode(t, b, p0, n0);
        const int py0 = p0 / a.gw;
        for (int kb = 0; kb < NKB; ++kb) {
          const int c = kb / PG, phg = kb % PG;
          mbar_wait(&raw_empty[rs], rph ^ 1, 11);
          // ONE 4-D box per k-block (two when the image is wider than 256 pixels): [npy_box patch rows][R image rows][xbox
          // pixels] of channel c, image rows phg*R.. of every patch row of the tile.  Rows past the patch (ph >= P) and
          // patch rows past the image are zero-filled by the map.  (The first version issued one 3-D box per patch row: the
          // single producer thread could not issue 8-14 of them per k-block fast enough and the transform warps spent 57 %
          // of their time waiting for pixels.)
          mbar_arrive_expect_tx(&raw_full[rs], static_cast<uint32_t>(a.nx * a.box_bytes_tx));
          uint8_t* dst = raw_base + rs * a.raw_stage;
          for (int ix = 0; ix < a.nx; ++ix)
            tma_load_4d(dst + ix * a.box_bytes, &tmI, &raw_full[rs], ix * a.xbox, phg * R, py0, b * 3 + c);
          if (++rs == a.raw_nst) {
            rs = 0;
            rph ^= 1;
          }
          mbar_wait(&b_empty[bs], bph ^ 1, 12);
          mbar_arrive_expect_tx(&b_full[bs], PE_B_BYTES);
          tma_load_2d(smem + PE_OFF_B + bs * PE_B_BYTES, &tmB, &b_full[bs], kb * 64, n0);
          if (++bs == a.b_nst) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(PE_BM, PE_BN, 0, 0);
      uint32_t kbc = 0;
      int bs = 0;
      uint32_t bph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 13);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PE_BN;
        for (int kb = 0; kb < NKB; ++kb, ++kbc) {
          const int s = kbc & 1;
          const uint32_t ph = (kbc >> 1) & 1;
          mbar_wait(&a_full[s], ph, 14);
          mbar_wait(&b_full[bs], bph, 15);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + PE_OFF_A + s * PE_A_BYTES);
          const uint32_t b_addr = smem_u32(smem + PE_OFF_B + bs * PE_B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(d_tmem, umma_desc_kmajor_sw128(a_addr + k * 32), umma_desc_kmajor_sw128(b_addr + k * 32), idesc,
                         (kb | k) != 0);
          umma_commit(&a_empty[s]);
          umma_commit(&b_empty[bs]);
          if (++bs == a.b_nst) {
            bs = 0;
            bph ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= 4 && warp < PE_EPI_WARP0) {
    // ------------------------------------------------------------------------------------------ pixel -> A-tile transform
    constexpr int RH = R / 2;                // image rows of the k-block per thread
    const int tt = threadIdx.x - 128;
    const int m = tt & 127;                  // patch inside the tile = row of the A tile
    const int r0 = (tt >> 7) * RH;           // first of this thread's rows of the k-block
    uint32_t kbc = 0;
    int rs = 0;
    uint32_t rph = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int b, p0, n0;
      decode(t, b, p0, n0);
      const int p = p0 + m;
      const bool valid = p < a.N;
      const int py = p / a.gw, px = p - py * a.gw;
      const int pyl = py - p0 / a.gw;
      const int xpix = px * P;
      const int ix = xpix / a.xbox;
      const uint32_t src_off = static_cast<uint32_t>(ix * a.box_bytes + pyl * a.row_bytes + (xpix - ix * a.xbox) * ES);
      for (int kb = 0; kb < NKB; ++kb, ++kbc) {
        const int s = kbc & 1;
        const uint32_t ph = (kbc >> 1) & 1;
        const int phg = kb % PG;
        uint32_t w[RH][PW / 2];   // bf16 pairs of this thread's RH x PW slice of the k-block
        mbar_wait(&raw_full[rs], rph, 16);
        const uint32_t src = smem_u32(raw_base + rs * a.raw_stage) + src_off;
#pragma unroll
        for (int rr = 0; rr < RH; ++rr) {
          const int r = r0 + rr;
          const bool row_ok = valid && (phg * R + r < P);
          if (row_ok) {
            uint32_t raw[PBYTES / 4];
#pragma unroll
            for (int v = 0; v < NV; ++v) lds_words<VB>(src + r * a.xbox * ES + v * VB, &raw[v * (VB / 4)]);
            if constexpr (ES == 4) {
#pragma unroll
              for (int j = 0; j < P / 2; ++j) w[rr][j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
            } else {
#pragma unroll
              for (int j = 0; j < P / 2; ++j) w[rr][j] = raw[j];
            }
#pragma unroll
            for (int j = P / 2; j < PW / 2; ++j) w[rr][j] = 0u;
          } else {
#pragma unroll
            for (int j = 0; j < PW / 2; ++j) w[rr][j] = 0u;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);   // pixels are in registers: the staging slot may be refilled
        if (++rs == a.raw_nst) {
          rs = 0;
          rph ^= 1;
        }
        mbar_wait(&a_empty[s], ph ^ 1, 17);
        const uint32_t dst = smem_u32(smem + PE_OFF_A + s * PE_A_BYTES);
#pragma unroll
        for (int rr = 0; rr < RH; ++rr)
#pragma unroll
          for (int j = 0; j < PW / 8; ++j)
            sts128(dst + sw128_offset(m, (r0 + rr) * (PW / 8) + j),
                   make_uint4(w[rr][4 * j], w[rr][4 * j + 1], w[rr][4 * j + 2], w[rr][4 * j + 3]));
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
      }
    }
  } else if (warp >= PE_EPI_WARP0) {
    // ------------------------------------------------------------------------------------------ epilogue
    const int ew = warp - PE_EPI_WARP0;
    const int grp = ew >> 2;
    const int quad = ew & 3;
    const int et = quad * 32 + lane;
    const bool leader = et == 0;
    const uint32_t bar_id = 1 + grp;
    uint8_t* cbuf = smem + PE_OFF_C + grp * PE_C_BYTES;
    const uint32_t sbuf = smem_u32(cbuf);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      int b, p0, n0;
      decode(t, b, p0, n0);
      const int acc = it & 1;
      const int p = p0 + et;
      const bool valid = p < a.N;
      if (p0 == 0) {   // class-token row of this image, columns of this tile: table row 0 (class_embedding + pos[0]) or zeros
        const int n = n0 + (ew * 32 + static_cast<int>(lane));
        if (n < a.D) a.out[static_cast<long long>(b) * a.L * a.D + n] = a.table != nullptr ? a.table[n] : __float2bfloat16(0.f);
      }
      mbar_wait(&tmem_full[acc], (it >> 1) & 1, 18);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * PE_BN + grp * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int ncol0 = n0 + grp * 128 + c * 64;
        const bool live = ncol0 < a.D;
        if (live) {
          if (leader) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
        }
        // two passes of 32 columns (the 640-thread CTA leaves 96 registers per thread)
#pragma unroll 1
        for (int hcol = 0; hcol < 2; ++hcol) {
          uint32_t v[32];
          if (live) {
            tmem_ld_x32(taddr + c * 64 + hcol * 32, v);
            tmem_ld_wait();
          }
          if (c == 1 && hcol == 1) {   // accumulator drained: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (!live) continue;
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          const int ncol = ncol0 + hcol * 32;
          if (a.table != nullptr && valid && ncol < a.D) {
            const uint4* src = reinterpret_cast<const uint4*>(a.table + static_cast<long long>(1 + p) * a.D + ncol);
            const int nvec = min(4, (a.D - ncol) >> 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < nvec) {
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
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(sbuf + sw128_offset(et, hcol * 4 + j),
                   make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                              pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7])));
        }
        if (!live) continue;
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_3d(&tmO, cbuf, ncol0, 1 + p0, b);   // token rows past the image's last patch are clipped by the map
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int P, typename T>
static int launch_patch_embed(const CUtensorMap& tmI, const CUtensorMap& tmB, const CUtensorMap& tmO, const PatchEmbedArgs& a,
                              cudaStream_t s) {
  auto kern = patch_embed_kernel<P, T>;
  static PerDeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM);
    if (e != cudaSuccess) return set_error(OVK_ERR_CUDA, "cudaFuncSetAttribute(patch_embed): %s", cudaGetErrorString(e));
    once.done();
  }
  const int total = a.B * a.tiles_pi * a.tiles_n;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, PE_THREADS, PE_SMEM, s>>>(tmI, tmB, tmO, a);
  return check_launch("patch_embed_kernel");
}

}  // namespace ovk

using namespace ovk;

// most patch rows a 128-patch tile can touch
static int pe_npy_box(int gw) {
  return gw % PE_BM == 0 || PE_BM % gw == 0 ? (PE_BM + gw - 1) / gw : (gw - 1 + PE_BM - 1) / gw + 1;
}
// bytes of one staged box ([npy_box][R][xbox] pixels, rounded up to 128) and of one staging slot (nx boxes, rounded up to 1 KB)
static int pe_box_bytes(int gw, int R, int xbox, int es) { return (pe_npy_box(gw) * R * xbox * es + 127) / 128 * 128; }
static int pe_stage_bytes(int gw, int nx, int R, int xbox, int es) { return (nx * pe_box_bytes(gw, R, xbox, es) + 1023) / 1024 * 1024; }

extern "C" int ovk_patch_embed_kdim(int P) {
  if (P != 14 && P != 16 && P != 32) return 0;
  const int PW = P <= 16 ? 16 : 32, R = 64 / PW, PG = (P + R - 1) / R;
  return 3 * PG * 64;
}

extern "C" int ovk_patch_embed_supported(int img_is_f32, int H, int W, int P, int D) {
  if (ovk_patch_embed_kdim(P) == 0 || H <= 0 || W <= 0 || H % P || W % P || D <= 0 || D % 8) return 0;
  const int es = img_is_f32 ? 4 : 2;
  const int gw = W / P;
  const int nx = (W + 255) / 256;
  if (gw % nx || ((W / nx) * es) % 16 || (W * es) % 16) return 0;
  const int R = 64 / (P <= 16 ? 16 : 32);
  if (pe_npy_box(gw) > 256) return 0;
  return 2 * pe_stage_bytes(gw, nx, R, W / nx, es) <= PE_DYN_BYTES - 2 * PE_B_BYTES ? 1 : 0;   // at least double buffers
}

extern "C" int ovk_patch_embed(const void* images, int img_is_f32, const void* w_packed, const void* pos_table, void* tokens,
                               int B, int H, int W, int P, int D, void* stream) {
  if (B <= 0) return set_error(OVK_ERR_SHAPE, "patch_embed: empty batch");
  if (!ovk_patch_embed_supported(img_is_f32, H, W, P, D))
    return set_error(OVK_ERR_SHAPE, "patch_embed: unsupported geometry H=%d W=%d P=%d D=%d (see ovk_patch_embed_supported)", H, W, P, D);
  if (pos_table != nullptr && (reinterpret_cast<uintptr_t>(pos_table) & 15))
    return set_error(OVK_ERR_ALIGN, "patch_embed: pos_table must be 16-byte aligned");
  const int es = img_is_f32 ? 4 : 2;
  PatchEmbedArgs a;
  a.B = B;
  a.gw = W / P;
  a.N = (H / P) * a.gw;
  a.D = D;
  a.L = a.N + 1;
  a.nx = (W + 255) / 256;
  a.xbox = W / a.nx;
  const int R = 64 / (P <= 16 ? 16 : 32);
  const int npy_box = pe_npy_box(a.gw);
  a.row_bytes = R * a.xbox * es;
  a.box_bytes_tx = npy_box * a.row_bytes;
  a.box_bytes = pe_box_bytes(a.gw, R, a.xbox, es);
  a.raw_stage = pe_stage_bytes(a.gw, a.nx, R, a.xbox, es);
  // 152 KB for the weight ring and the pixel staging ring: a third weight stage when two staging slots still fit behind it
  // (bytes in flight per SM are what bounds this kernel: both rings are fed from L2 with ~1 us latency)
  a.b_nst = (PE_DYN_BYTES - 3 * PE_B_BYTES) / a.raw_stage >= 2 ? 3 : 2;
  {
    const char* e = getenv("OVK_PE_BSTAGES");   // A/B
    if (e != nullptr && e[0] == '2') a.b_nst = 2;
  }
  const int raw_room = (PE_DYN_BYTES - a.b_nst * PE_B_BYTES) / a.raw_stage;
  a.raw_nst = raw_room < PE_RAW_MAX_STAGES ? raw_room : PE_RAW_MAX_STAGES;
  a.tiles_pi = (a.N + PE_BM - 1) / PE_BM;
  a.tiles_n = (D + PE_BN - 1) / PE_BN;
  a.table = reinterpret_cast<const __nv_bfloat16*>(pos_table);
  a.out = reinterpret_cast<__nv_bfloat16*>(tokens);
  const int Kp = ovk_patch_embed_kdim(P);
  CUtensorMap tmI, tmB, tmO;
  int rc;
  {
    // the image as (x, row inside the patch, patch row, channel x image): one box = R image rows of every patch row of a tile
    uint64_t dims[4] = {static_cast<uint64_t>(W), static_cast<uint64_t>(P), static_cast<uint64_t>(H / P), static_cast<uint64_t>(3) * B};
    uint64_t strides[3] = {static_cast<uint64_t>(W) * es, static_cast<uint64_t>(W) * P * es, static_cast<uint64_t>(W) * H * es};
    uint32_t box[4] = {static_cast<uint32_t>(a.xbox), static_cast<uint32_t>(R), static_cast<uint32_t>(npy_box), 1};
    rc = img_is_f32 ? make_tmap_nd_f32(&tmI, images, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)
                    : make_tmap_nd_bf16(&tmI, images, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  if ((rc = make_tmap_2d_bf16(&tmB, w_packed, Kp, D, Kp, 64, PE_BN))) return rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(a.L), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(D) * a.L * 2};
    uint32_t box[3] = {64, PE_BM, 1};
    if ((rc = make_tmap_nd_bf16(&tmO, tokens, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define OVK_PE_CASE(PP)                                                                    \
  if (P == PP) return img_is_f32 ? launch_patch_embed<PP, float>(tmI, tmB, tmO, a, s)      \
                                 : launch_patch_embed<PP, __nv_bfloat16>(tmI, tmB, tmO, a, s);
  OVK_PE_CASE(14)
  OVK_PE_CASE(16)
  OVK_PE_CASE(32)
#undef OVK_PE_CASE
  return set_error(OVK_ERR_SHAPE, "patch_embed: unsupported patch size %d", P);
}
