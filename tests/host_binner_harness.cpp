// TEST INFRASTRUCTURE.  Runs the product's per-record / per-group binning bodies (taxator-tk_b200/csrc/binner_core.h,
// the code the CUDA kernels execute) on the host, one record / group after the other, in the order of the launcher
// (csrc/binner.cu launch_binner), so that the CPU suite can compare them with the oracle without a GPU.
#include <cstring>
#include <vector>

#include "../taxator-tk_b200/csrc/binner_core.h"

using namespace trpa;

extern "C" int hb_binner(const uint32_t* parent, const uint8_t* depth, uint32_t n_nodes, uint32_t root,
                         const trpa_bin_params* pp, const trpa_bin_record* records, uint32_t n_records,
                         const uint32_t* supports, const uint32_t* group_begin, uint32_t n_groups,
                         const uint8_t* rank_of_node, const float* pid_per_rank, trpa_bin_result* out, trpa_bin_stats* stats) {
  uint32_t max_depth = 0;
  for (uint32_t i = 0; i < n_nodes; ++i) max_depth = depth[i] > max_depth ? depth[i] : max_depth;
  const uint32_t D1 = max_depth + 1;
  std::vector<uint32_t> node_support(n_nodes, 0), node_seen(n_nodes, 0), node_pruned(n_nodes, 0), lower(n_records + 1), curnode(n_records + 1),
      maj_node(n_records + 1), path_node((size_t)(n_groups + 1) * D1);
  std::vector<uint8_t> alive(n_records + 1, 1), state(n_records + 1), path_branch((size_t)(n_groups + 1) * D1);
  std::vector<float> maj_sum(n_records + 1);
  std::vector<u16> tot((size_t)(n_records + 1) * D1), path_direct((size_t)(n_groups + 1) * D1), path_total((size_t)(n_groups + 1) * D1);
  BinTables T{records, supports, parent, depth, root, n_records, lower.data(), alive.data()};
  uint32_t min_found = 0xffffffffu;
  for (uint32_t r = 0; r < n_records; ++r) { lower[r] = records[r].lower_node; alive[r] = 1; }
  for (uint32_t r = 0; r < n_records; ++r) bin_support_record(T, r, node_support.data(), node_seen.data(), &min_found);
  uint32_t min_support = pp->min_support_in_sample;
  if (pp->min_support_in_sample_fraction) min_support = node_support[root] * pp->min_support_in_sample_fraction;
  if (n_records && min_found < min_support)
    for (uint32_t r = 0; r < n_records; ++r) bin_prune_record(T, r, node_support.data(), min_support, node_pruned.data());
  node_seen[root] = 1;
  const BinWork W{state.data(), curnode.data(), maj_node.data(), maj_sum.data(), tot.data(), D1, path_node.data(), path_direct.data(),
                  path_total.data(), path_branch.data()};
  for (uint32_t g = 0; g < n_groups; ++g) bin_combine_group(T, g, group_begin, *pp, rank_of_node, pid_per_rank, W, out);
  if (stats) {
    stats->nested_taxa = stats->pruned_taxa = 0;
    for (uint32_t i = 0; i < n_nodes; ++i) { stats->nested_taxa += node_seen[i] ? 1 : 0; stats->pruned_taxa += node_pruned[i] ? 1 : 0; }
    stats->root_support = node_support[root];
    stats->min_support_found = min_found;
  }
  return 0;
}
