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

}  // namespace ovk
