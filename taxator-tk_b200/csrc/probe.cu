// INT32 ALU-pipe throughput probe: the roofline denominator for the bit-vector edit-distance kernel.
// Each thread runs 8 independent chains of LOP3 / IADD3 / SHF (the instruction classes the Myers
// column update is made of, all on the "alu" pipe) so the pipe, not latency, is the limit.
#include "common.cuh"
#include "launch.h"

namespace trpa {

constexpr int kProbeOpsPerIter = 8 * 6;  // counted in SASS: 8 chains x (3 LOP3 + 2 IADD3 + 1 SHF)

__global__ void __launch_bounds__(256) alu_probe_kernel(u32* sink, int iters) {
  u32 x[8], y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 2654435761u + k; y[k] = blockIdx.x * 40503u + 7 * k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      u32 a = (x[k] & y[k]) ^ (u32)it;          // LOP3
      u32 b = x[k] + y[k] + a;                  // IADD3
      u32 c = (a | b) & ~x[k];                  // LOP3
      u32 d = __funnelshift_l(b, c, 1);         // SHF
      x[k] = (c ^ d) | a;                       // LOP3
      y[k] = b + d + c;                         // IADD3
    }
  }
  u32 r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r ^= x[k] + y[k];
  if (r == 0x12345678u) sink[0] = r;  // practically never; keeps the chains alive
}

int alu_probe_ops_per_iter() { return kProbeOpsPerIter; }

cudaError_t launch_alu_probe(u32* sink, int iters, cudaStream_t stream, int* blocks, int* threads) {
  *blocks = 148 * 8;
  *threads = 256;
  alu_probe_kernel<<<*blocks, *threads, 0, stream>>>(sink, iters);
  return cudaGetLastError();
}

}  // namespace trpa
