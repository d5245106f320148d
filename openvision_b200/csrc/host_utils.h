// Host-side helpers shared by the C-ABI translation units: error reporting, device queries, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ovk.h"

namespace ovk {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int num_sms();

// Per-device latch for function attributes (cudaFuncSetAttribute is per device, a process may drive several GPUs):
//   static PerDeviceOnce once;  if (once.need()) { ...cudaFuncSetAttribute...; once.done(); }
// Two threads racing through need() both set the attribute, which is harmless.
struct PerDeviceOnce {
  unsigned long long mask = 0;
  static int slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= 64) ? -1 : dev;   // devices past 63: never latched, attribute set on every launch
  }
  bool need() const {
    const int s = slot();
    return s < 0 || ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> s) & 1ull) == 0;
  }
  void done() {
    const int s = slot();
    if (s >= 0) __atomic_fetch_or(&mask, 1ull << s, __ATOMIC_RELEASE);
  }
};

// 2-D bf16 row-major tensor [outer, inner] with row stride ld (elements); box = [box_outer, box_inner];
// SWIZZLE_128B (box_inner * 2 bytes must be <= 128).
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);
// Same for an fp32 tensor (box_inner * 4 bytes <= 128).
int make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                     uint32_t box_inner, uint32_t box_outer);
// Generic N-d (<=5) bf16 map; dims / strides innermost first (strides in bytes for dims 1..rank-1).
int make_tmap_nd_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapSwizzle swizzle);

// Generic N-d (<=5) fp32 map, same conventions.
int make_tmap_nd_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace ovk
