// Consensus binning of a sample's prediction records on the GPU: what the reference's `binner` does after parsing
//   core/binner.cpp:213-282   sample support of every taxon (running maximum of a record's support from its lower
//                             node up to the root, summed over all records) + optional noise pruning
//   core/src/predictionranges.hh:122-266   per sequence (group of records): majority walk down the taxonomy
//   core/binner.cpp:296-329   identity constraints per rank, the (taxon, support, length) that is written
// Three kernels: one thread per record accumulates the sample support with atomics (a segmented tree-walk
// reduction: <= depth hops per record), one thread per record prunes, one thread per group runs the walk down
// (bodies in binner_core.h).  The reference's integer widths are kept: per-record supports are uint32, everything
// inside the combination is uint16 (medium_unsigned_int, core/src/types.hh:35) and wraps exactly like there; the
// majority vote sums floats in record order.  HBM-bound streaming over the record table; the walk down is latency
// bound (<= depth levels).
#include <algorithm>

#include "common.cuh"
#include "launch.h"
#include "binner_core.h"

namespace trpa {

__global__ void bin_support_kernel(BinTables T, u32* __restrict__ node_support, u32* __restrict__ node_seen, u32* __restrict__ min_found) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < T.n_records) bin_support_record(T, r, node_support, node_seen, min_found);
}
__global__ void bin_prune_kernel(BinTables T, const u32* __restrict__ node_support, u32 min_support, u32* __restrict__ node_pruned) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < T.n_records) bin_prune_record(T, r, node_support, min_support, node_pruned);
}
__global__ void bin_combine_kernel(BinTables T, const u32* __restrict__ group_begin, u32 n_groups, trpa_bin_params pp,
                                   const uint8_t* __restrict__ rank_of_node, const float* __restrict__ pid_per_rank, BinWork W,
                                   trpa_bin_result* __restrict__ out) {
  const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) bin_combine_group(T, g, group_begin, pp, rank_of_node, pid_per_rank, W, out);
}
__global__ void bin_count_kernel(const u32* __restrict__ flags_a, const u32* __restrict__ flags_b, u32 n, u32* __restrict__ out2) {
  u32 a = 0, b = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a += flags_a[i] ? 1u : 0u; b += flags_b[i] ? 1u : 0u; }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(&out2[0], a); if (b) atomicAdd(&out2[1], b); }
}

__global__ void bin_init_kernel(BinTables T) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= T.n_records) return;
  T.lower[r] = T.recs[r].lower_node;
  T.alive[r] = 1;
}

cudaError_t launch_binner(const trpa_bin_record* recs, u32 n_records, const u32* supports, const u32* group_begin, u32 n_groups,
                          const Taxonomy& tax, u32 n_nodes, u32 max_depth, const trpa_bin_params& pp, const uint8_t* rank_of_node,
                          const float* pid_per_rank, BinScratch& S, trpa_bin_result* out, u32* stats4, cudaStream_t stream) {
  BinTables T{recs, supports, tax.parent, tax.depth, tax.root, n_records, S.lower, S.alive};
  const u32 rb = (n_records + 255) / 256;
  cudaError_t e;
  if ((e = cudaMemsetAsync(S.node_support, 0, sizeof(u32) * n_nodes, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(S.node_seen, 0, sizeof(u32) * n_nodes, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(S.node_pruned, 0, sizeof(u32) * n_nodes, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(stats4, 0, sizeof(u32) * 4, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(stats4 + 2, 0xff, sizeof(u32), stream)) != cudaSuccess) return e;   // minimum support found
  if (n_records) {
    bin_init_kernel<<<rb, 256, 0, stream>>>(T);
    bin_support_kernel<<<rb, 256, 0, stream>>>(T, S.node_support, S.node_seen, stats4 + 2);
  }
  // the thresholds of the noise filter need the root's support: one small read-back
  u32 h[2] = {0, 0xffffffffu};
  if ((e = cudaMemcpyAsync(&h[0], S.node_support + tax.root, sizeof(u32), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(&h[1], stats4 + 2, sizeof(u32), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  u32 min_support = pp.min_support_in_sample;
  if (pp.min_support_in_sample_fraction) min_support = h[0] * pp.min_support_in_sample_fraction;   // binner.cpp:256
  if (n_records && h[1] < min_support) bin_prune_kernel<<<rb, 256, 0, stream>>>(T, S.node_support, min_support, S.node_pruned);
  // the root always has an entry in the reference's map (binner.cpp:220)
  if ((e = cudaMemsetAsync(S.node_seen + tax.root, 1, 1, stream)) != cudaSuccess) return e;
  bin_count_kernel<<<64, 256, 0, stream>>>(S.node_seen, S.node_pruned, n_nodes, stats4);
  if (n_groups)
  {
    const BinWork W{S.state, S.curnode, S.maj_node, S.maj_sum, S.tot, max_depth + 1, S.path_node, S.path_direct, S.path_total, S.path_branch};
    bin_combine_kernel<<<(n_groups + 63) / 64, 64, 0, stream>>>(T, group_begin, n_groups, pp, rank_of_node, pid_per_rank, W, out);
  }
  return cudaGetLastError();
}

}  // namespace trpa
