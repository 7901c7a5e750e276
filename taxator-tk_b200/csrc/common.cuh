// Shared device-side types for the RPA hot path (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "types.h"

namespace trpa {

typedef uint32_t u32;
typedef uint64_t u64;

#define TRPA_CUDA_OK(expr)                                                   \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      trpa::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));  \
      return -1;                                                             \
    }                                                                        \
  } while (0)

}  // namespace trpa
