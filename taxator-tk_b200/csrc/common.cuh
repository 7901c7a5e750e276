// Shared device-side types for the RPA hot path (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "types.h"

namespace trpa {

typedef uint32_t u32;
typedef uint64_t u64;


}  // namespace trpa
