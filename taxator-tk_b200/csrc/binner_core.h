// Per-record and per-group bodies of the consensus binning kernels (binner.cu), written so that the same code
// compiles for the host: tests single-step them against the oracle without a GPU; the product only ever runs them
// inside the kernels.  Reference: core/binner.cpp:213-282, 296-329; core/src/predictionranges.hh:122-266.
#pragma once
#include <cstdint>
#include "../../include/taxator_rpa_b200.h"
#include "machine.h"   // TRPA_HD, atomic_add_u32

namespace trpa {

typedef uint32_t u32;

TRPA_HD void atomic_min_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}

typedef unsigned short u16;

struct BinTables {
  const trpa_bin_record* recs;
  const u32* supports;
  const u32* parent;
  const uint8_t* depth;
  u32 root;
  u32 n_records;
  // per record, after pruning
  u32* lower;       // current lower node
  uint8_t* alive;   // 0: erased by the noise filter
};

// PredictionRecordBase::getSupportAt(depth) (predictionrecord.hh:76-88) with the record's possibly pruned lower node:
// pruneLowerNode() only shortens taxon_support_ (:185-189)
TRPA_HD u32 bin_support_at(const BinTables& T, u32 r, u32 lower, int d) {
  const trpa_bin_record b = T.recs[r];
  const int up = (int)T.depth[b.upper_node];
  const int index = d - up;
  if (index < 0) return 0u;
  const int size = (int)T.depth[lower] - up + 1;
  return T.supports[b.support_begin + (index < size ? index : size - 1)];
}

// STEP 1, binner.cpp:219-252
TRPA_HD void bin_support_record(const BinTables& T, u32 r, u32* node_support, u32* node_seen, u32* min_found) {
  const trpa_bin_record b = T.recs[r];
  u32 pit = b.lower_node;
  u32 total = bin_support_at(T, r, b.lower_node, (int)T.depth[pit]);
  atomic_min_u32(min_found, total);
  atomic_add_u32(&node_support[pit], total);
  node_seen[pit] = 1u;
  while (pit != T.root) {
    pit = T.parent[pit];
    const u32 s = bin_support_at(T, r, b.lower_node, (int)T.depth[pit]);
    total = total > s ? total : s;
    atomic_add_u32(&node_support[pit], total);
    node_seen[pit] = 1u;
  }
}

// noise removal, binner.cpp:259-281
TRPA_HD void bin_prune_record(const BinTables& T, u32 r, const u32* node_support, u32 min_support, u32* node_pruned) {
  const trpa_bin_record b = T.recs[r];
  u32 pit = b.lower_node;
  while (pit != b.upper_node && node_support[pit] < min_support) { node_pruned[pit] = 1u; pit = T.parent[pit]; }
  if (pit == b.upper_node && node_support[pit] < min_support) {
    node_pruned[pit] = 1u;
    T.alive[r] = 0;
    return;
  }
  T.lower[r] = pit;
}

TRPA_HD u32 bin_ancestor_at(const BinTables& T, u32 node, int d) {
  while ((int)T.depth[node] > d) node = T.parent[node];
  return node;
}

// STEP 2: combinePredictionRanges (predictionranges.hh:122-266) + binner.cpp:296-329, one thread per group.
// Scratch per record (the group's slice): state (0 gone, 1 in the list), current node, majority table.
// D1 = taxonomy depth + 1; tot[r * D1 + d] = total_support[d] of record r (uint16).
struct BinWork {   // scratch of the walk down, see launch.h BinScratch
  uint8_t* state; u32* curnode; u32* maj_node; float* maj_sum; u16* tot; u32 D1;
  u32* path_node; u16* path_direct; u16* path_total; uint8_t* path_branch;
};
TRPA_HD void bin_combine_group(const BinTables& T, u32 g, const u32* group_begin, const trpa_bin_params& pp,
                               const uint8_t* rank_of_node, const float* pid_per_rank, const BinWork& W, trpa_bin_result* out) {
  uint8_t* state = W.state; u32* curnode = W.curnode; u32* maj_node = W.maj_node; float* maj_sum = W.maj_sum;
  u16* tot = W.tot; const u32 D1 = W.D1;
  u32* path_node = W.path_node; u16* path_direct = W.path_direct; u16* path_total = W.path_total; uint8_t* path_branch = W.path_branch;
  const u32 gb = group_begin[g], ge = group_begin[g + 1];
  trpa_bin_result o;
  o.node = o.support = o.length = 0; o.mode = TRPA_BIN_EMPTY;
  o.lower_node = o.upper_node = T.root; o.lower_support = o.upper_support = 0;
  u32 n_alive = 0, first = gb;
  for (u32 r = ge; r-- > gb;) if (T.alive[r]) { ++n_alive; first = r; }
  if (n_alive == 0) { out[g] = o; return; }
  const u32 root_depth = T.depth[T.root];
  if (n_alive == 1) {   // pass-through (binner.cpp:301-303)
    const trpa_bin_record b = T.recs[first];
    const u32 lower = T.lower[first];
    o.mode = TRPA_BIN_SINGLE;
    o.lower_node = lower; o.upper_node = b.upper_node;
    o.lower_support = bin_support_at(T, first, lower, (int)T.depth[lower]);
    o.upper_support = bin_support_at(T, first, lower, (int)T.depth[b.upper_node]);
    o.length = b.query_length;
  } else {
    // initialise the tuples (predictionranges.hh:134-160); query lengths are summed once per sequence identifier
    u16 summed_support = 0;
    u32 summed_length = 0;
    for (u32 r = gb; r < ge; ++r) {
      state[r] = T.alive[r];
      if (!T.alive[r]) continue;
      const u32 lower = T.lower[r];
      int i = (int)T.depth[lower];
      const u16 support = (u16)bin_support_at(T, r, lower, i);
      summed_support = (u16)(summed_support + support);
      const u32 qid = T.recs[r].query_id;
      bool seen = false;
      for (u32 q = gb; q < r && !seen; ++q) seen = T.alive[q] && T.recs[q].query_id == qid;
      if (!seen) summed_length += T.recs[r].query_length;
      u16* tr = tot + (size_t)r * D1;
      tr[i] = support;
      while (--i >= 0) {
        const u16 d = (u16)bin_support_at(T, r, lower, i);
        tr[i] = tr[i + 1] > d ? tr[i + 1] : d;
      }
      curnode[r] = T.root;
    }
    const u16 frac = static_cast<u16>(pp.signal_majority * summed_support);
    const u16 thresh = frac > (u16)pp.min_support_per_sequence ? frac : (u16)pp.min_support_per_sequence;
    u32* pn = path_node + (size_t)g * D1; u16* pd = path_direct + (size_t)g * D1; u16* pt = path_total + (size_t)g * D1;
    uint8_t* pb = path_branch + (size_t)g * D1;
    int cur = (int)root_depth, n_path = 0, lower_direct = -1;
    u32 n_list = n_alive;
    auto get_support = [&](u16& direct, u16& total) {
      u16 d = 0, t = 0;
      for (u32 r = gb; r < ge; ++r) if (state[r]) {
        d = (u16)(d + (u16)bin_support_at(T, r, T.lower[r], cur));
        t = (u16)(t + tot[(size_t)r * D1 + cur]);
      }
      direct = d; total = t;
    };
    u16 direct, total;
    get_support(direct, total);
    while (n_list) {
      u32 node = T.root;
      for (u32 r = gb; r < ge; ++r) if (state[r]) { node = curnode[r]; break; }
      if (direct >= thresh) lower_direct = n_path;
      pn[n_path] = node; pd[n_path] = direct; pt[n_path] = total; pb[n_path] = 0;
      // removeIf<3>: paths that ended at this level; stepDown
      for (u32 r = gb; r < ge; ++r) if (state[r]) {
        if ((int)T.depth[T.lower[r]] == cur) { state[r] = 0; --n_list; }
      }
      ++cur;
      for (u32 r = gb; r < ge; ++r) if (state[r]) curnode[r] = bin_ancestor_at(T, T.lower[r], cur);
      // reduceToMajority (predictionranges.hh:78-110)
      bool branched = false;
      if (n_list >= 2) {
        u32 n_maj = 0, max_node = 0xffffffffu;
        float max_support = .0f;
        for (u32 r = gb; r < ge; ++r) if (state[r]) {
          const u32 nd = curnode[r];
          u32 k = 0;
          while (k < n_maj && maj_node[gb + k] != nd) ++k;
          const float v = (float)tot[(size_t)r * D1 + cur];
          if (k == n_maj) { maj_node[gb + k] = nd; maj_sum[gb + k] = v; ++n_maj; }
          else maj_sum[gb + k] += v;
          if (maj_sum[gb + k] > max_support) { max_support = maj_sum[gb + k]; max_node = nd; }
        }
        if (n_maj != 1) {
          for (u32 r = gb; r < ge; ++r) if (state[r] && curnode[r] != max_node) { state[r] = 0; --n_list; }
          branched = true;
        }
      }
      pb[n_path] = branched ? 1 : 0;
      ++n_path;
      get_support(direct, total);
    }
    o.length = summed_length;
    if (lower_direct >= 0) {
      o.mode = TRPA_BIN_DIRECT;
      o.lower_node = pn[lower_direct]; o.lower_support = pt[lower_direct];
      o.upper_node = o.lower_node; o.upper_support = o.lower_support;
      for (int j = lower_direct; j >= 0; --j) {
        if (pd[j] >= thresh) {
          o.upper_support = pt[j]; o.upper_node = pn[j];
          if (pb[j]) break;
        }
      }
    } else {
      o.mode = TRPA_BIN_FALLBACK;
      int i = n_path - 1;
      while (i >= 0 && !(pt[i] >= thresh)) --i;
      if (i < 0) i = 0;
      o.lower_node = o.upper_node = pn[i];
      o.lower_support = o.upper_support = pt[i];
    }
  }
  // the record that is written (binner.cpp:305-329)
  auto prec_support_at = [&](u32 node) -> u32 {
    const int index = (int)T.depth[node] - (int)T.depth[o.upper_node];
    if (index < 0) return 0u;
    if (o.mode == TRPA_BIN_SINGLE) return bin_support_at(T, first, T.lower[first], (int)T.depth[node]);
    return index == 0 ? o.upper_support : o.lower_support;
  };
  if (o.upper_node != T.root && pp.n_ranks) {
    const double seqlen = static_cast<double>(o.length);
    float min_pid = 0.f;
    u32 predict = T.root;
    const u32 target = o.upper_node;
    const float rank_pid = prec_support_at(target) / seqlen;
    int d = (int)root_depth;
    u32 pit;
    do {
      ++d;
      pit = bin_ancestor_at(T, target, d);
      const float c = pid_per_rank[rank_of_node[pit]];
      if (c >= 0.f) min_pid = min_pid > c ? min_pid : c;
      if (rank_pid < min_pid) break;
      predict = pit;
    } while (pit != target);
    o.node = predict;
    o.support = prec_support_at(predict);
  } else {
    o.node = o.upper_node;
    o.support = prec_support_at(o.upper_node);
  }
  out[g] = o;
}

}  // namespace trpa
