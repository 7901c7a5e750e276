// Shared device-side types for the RPA hot path (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "types.h"

namespace trpa {

typedef uint32_t u32;
typedef uint64_t u64;

#ifdef __CUDACC__
// Column codes of a 32-base word for the edit-distance kernel's table lookup: base j (bit j of the two
// bit-planes) as a 2-bit code (p1:p0) at bits [2(j%16)+1 : 2(j%16)]: .x = bases 0..15, .y = bases 16..31.
// (N positions carry code 0: pairs with N use the planes.)
__device__ __forceinline__ u32 spread16(u32 v) {   // bit i -> bit 2i
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ uint2 nt_codes(u32 p0, u32 p1) {
  return make_uint2(spread16(p0 & 0xffffu) | (spread16(p1 & 0xffffu) << 1), spread16(p0 >> 16) | (spread16(p1 >> 16) << 1));
}
#endif


}  // namespace trpa
