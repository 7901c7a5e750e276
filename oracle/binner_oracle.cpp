// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's `binner` (the consumer of taxator's GFF3) over
// flat arrays, with the reference's own containers and integer widths.  Nothing in the product path
// (taxator-tk_b200/) may include, link or call this file.
//
// Parity status: PINNED against the real reference binner built in this container (oracle/_ref/binner,
// unmodified core/binner.cpp + Boost shims): tests/golden/binner_*.tsv, tests/test_binner.py.
//
// What each part follows (paths relative to /root/reference/core):
//   sample support + noise pruning   binner.cpp:213-282
//   combine()                        src/predictionranges.hh:29-110 (details::*), :122-266 (combinePredictionRanges)
//   support_at()                     src/predictionrecord.hh:76-88 (getSupportAt), :185-189 (pruneLowerNode)
//   identity constraints, output     binner.cpp:296-329
//   integer widths                   src/types.hh:34-36 (small = uint8, medium = uint16, large = uint32)
#include <algorithm>
#include <cstdint>
#include <list>
#include <map>
#include <set>
#include <vector>

#include "../include/taxator_rpa_b200.h"

namespace {

typedef uint8_t small_uint;
typedef uint16_t medium_uint;
typedef uint32_t large_uint;

struct Tax {
  const uint32_t* parent;
  const uint8_t* depth;
  uint32_t root;
};

// one prediction record with its own taxon_support_ vector (upper ... lower)
struct Rec {
  uint32_t lower, upper;
  std::vector<large_uint> support;
  large_uint qlen;
  uint32_t qid;
};

large_uint support_at(const Tax& t, const Rec& r, int depth) {   // predictionrecord.hh:76-88
  const int index = depth - (int)t.depth[r.upper];
  if (index >= 0) {
    if (index < (int)r.support.size()) return r.support.at(index);
    return r.support.back();
  }
  return 0;
}

// node on the path root -> `lower` at depth d
uint32_t ancestor_at(const Tax& t, uint32_t lower, int d) {
  uint32_t x = lower;
  while ((int)t.depth[x] > d) x = t.parent[x];
  return x;
}

struct Tuple {   // details::TupleRangeCombine: (path iterator, direct support, total support, path ended)
  uint32_t lower;
  int cur;       // depth of the iterator's current node
  std::vector<medium_uint> direct, total;
  bool ended;
  uint32_t node(const Tax& t) const { return ancestor_at(t, lower, cur); }
};

struct PathEntry { uint32_t node; medium_uint direct, total; bool branching; };

void get_support(const std::list<Tuple>& tl, medium_uint& direct, medium_uint& total) {   // predictionranges.hh:58-68
  medium_uint d = 0, i = 0;
  for (const Tuple& x : tl) { d += x.direct[x.cur]; i += x.total[x.cur]; }
  direct = d; total = i;
}

bool reduce_to_majority(const Tax& t, std::list<Tuple>& tl) {   // predictionranges.hh:78-110
  if (tl.size() < 2) return false;
  std::map<uint32_t, float> supports;
  uint32_t max_node = 0xffffffffu;
  float max_support = .0;
  for (const Tuple& x : tl) {
    const uint32_t node = x.node(t);
    std::map<uint32_t, float>::iterator f = supports.find(node);
    if (f != supports.end()) f->second += x.total[x.cur];
    else f = supports.insert(std::make_pair(node, (float)x.total[x.cur])).first;
    if (f->second > max_support) { max_support = f->second; max_node = node; }
  }
  if (supports.size() == 1) return false;
  for (std::list<Tuple>::iterator it = tl.begin(); it != tl.end();) {
    if (it->node(t) != max_node) it = tl.erase(it);
    else ++it;
  }
  return true;
}

// combinePredictionRanges (predictionranges.hh:122-266) for the records of one group (size > 1)
void combine(const Tax& t, const std::list<Rec>& preds, float min_signal_percentage, medium_uint min_support,
             trpa_bin_result& out) {
  std::list<Tuple> tlist;
  medium_uint summed_support = 0;
  large_uint summed_length = 0;
  {
    std::set<uint32_t> processed;
    for (const Rec& r : preds) {
      int i = t.depth[r.lower];
      const medium_uint support = support_at(t, r, i);
      summed_support += support;
      if (processed.insert(r.qid).second) summed_length += r.qlen;
      const size_t depth = i + 1;
      Tuple tp;
      tp.lower = r.lower; tp.cur = 0; tp.ended = false;
      tp.direct.assign(depth, 0); tp.total.assign(depth, 0);
      tp.total[i] = tp.direct[i] = support;
      while (--i >= 0) {
        tp.direct[i] = support_at(t, r, i);
        tp.total[i] = std::max(tp.total[i + 1], tp.direct[i]);
      }
      tlist.push_back(tp);
    }
  }
  const medium_uint thresh = std::max(static_cast<medium_uint>(min_signal_percentage * summed_support), min_support);
  medium_uint direct_support, total_support;
  std::vector<PathEntry> path;
  auto set_path_end_state = [&]() { for (Tuple& x : tlist) x.ended = x.cur == (int)(x.direct.size() - 1); };
  set_path_end_state();
  get_support(tlist, direct_support, total_support);
  int lower_direct_node_index = -1, running_index = 0;
  while (!tlist.empty()) {
    const uint32_t node = tlist.front().node(t);
    if (direct_support >= thresh) lower_direct_node_index = running_index;
    path.push_back(PathEntry{node, direct_support, total_support, false});
    for (std::list<Tuple>::iterator it = tlist.begin(); it != tlist.end();) {   // removeIf<3>
      if (it->ended) it = tlist.erase(it);
      else ++it;
    }
    for (Tuple& x : tlist) ++x.cur;   // stepDown
    ++running_index;
    set_path_end_state();
    path.back().branching = reduce_to_majority(t, tlist);
    get_support(tlist, direct_support, total_support);
  }
  out.length = summed_length;
  if (lower_direct_node_index >= 0) {   // direct mode
    out.mode = TRPA_BIN_DIRECT;
    const uint32_t lower_node = path[lower_direct_node_index].node;
    const medium_uint lower_support = path[lower_direct_node_index].total;
    medium_uint upper_support = lower_support;
    uint32_t upper_node = lower_node;
    for (int j = lower_direct_node_index; j >= 0; --j) {
      if (path[j].direct >= thresh) {
        upper_support = path[j].total;
        upper_node = path[j].node;
        if (path[j].branching) break;
      }
    }
    out.lower_node = lower_node; out.lower_support = lower_support;
    out.upper_node = upper_node; out.upper_support = upper_support;
    return;
  }
  out.mode = TRPA_BIN_FALLBACK;
  for (int i = (int)path.size() - 1; i >= 0; --i) {
    if (path[i].total >= thresh) {
      out.lower_node = out.upper_node = path[i].node;
      out.lower_support = out.upper_support = path[i].total;
      return;
    }
  }
  out.lower_node = out.upper_node = path[0].node;
  out.lower_support = out.upper_support = path[0].total;
}

}  // namespace

extern "C" int orc_binner(const uint32_t* parent, const uint8_t* depth, uint32_t n_nodes, uint32_t root,
                          const trpa_bin_params* pp, const trpa_bin_record* records, uint32_t n_records,
                          const uint32_t* supports, const uint32_t* group_begin, uint32_t n_groups,
                          const uint8_t* rank_of_node, const float* pid_per_rank, trpa_bin_result* out,
                          trpa_bin_stats* stats) {
  (void)n_nodes;
  const Tax t{parent, depth, root};
  std::vector<std::list<Rec>> groups(n_groups);
  for (uint32_t g = 0; g < n_groups; ++g)
    for (uint32_t k = group_begin[g]; k < group_begin[g + 1]; ++k) {
      const trpa_bin_record& b = records[k];
      Rec r;
      r.lower = b.lower_node; r.upper = b.upper_node; r.qlen = b.query_length; r.qid = b.query_id;
      const uint32_t n = (uint32_t)depth[b.lower_node] - depth[b.upper_node] + 1;
      r.support.assign(supports + b.support_begin, supports + b.support_begin + n);
      groups[g].push_back(r);
    }
  (void)n_records;

  // STEP 1 (binner.cpp:213-254): sample support of every taxon
  large_uint minimum_support_found = 0xffffffffu;
  std::map<uint32_t, large_uint> support;   // FastNodeMap: one map per depth, same content
  support[root];
  for (const auto& recs : groups)
    for (const Rec& r : recs) {
      uint32_t pit = r.lower;
      large_uint total = support_at(t, r, depth[pit]);
      minimum_support_found = std::min(minimum_support_found, total);
      support[pit] += total;
      if (pit != root) {
        for (pit = parent[pit]; pit != root; pit = parent[pit]) {
          total = std::max(total, support_at(t, r, depth[pit]));
          support[pit] += total;
        }
        total = std::max(total, support_at(t, r, depth[root]));
        support[root] += total;
      }
    }
  large_uint min_support_in_sample = pp->min_support_in_sample;
  if (pp->min_support_in_sample_fraction) min_support_in_sample = support[root] * pp->min_support_in_sample_fraction;
  // noise removal (binner.cpp:259-282)
  std::set<uint32_t> pruned;
  if (minimum_support_found < min_support_in_sample) {
    for (auto& recs : groups)
      for (std::list<Rec>::iterator it = recs.begin(); it != recs.end();) {
        uint32_t pit = it->lower;
        while (pit != it->upper && support[pit] < min_support_in_sample) { pruned.insert(pit); pit = parent[pit]; }
        if (pit == it->upper && support[pit] < min_support_in_sample) {
          pruned.insert(pit);
          it = recs.erase(it);
          continue;
        }
        if (pit != it->lower) {   // pruneLowerNode
          it->support.resize(depth[pit] - depth[it->upper] + 1);
          it->lower = pit;
        }
        ++it;
      }
  }
  if (stats) {
    stats->nested_taxa = support.size();
    stats->root_support = support[root];
    stats->pruned_taxa = pruned.size();
    stats->min_support_found = minimum_support_found;
  }

  // STEP 2 (binner.cpp:284-329)
  for (uint32_t g = 0; g < n_groups; ++g) {
    trpa_bin_result& o = out[g];
    o = trpa_bin_result();
    const std::list<Rec>& recs = groups[g];
    if (recs.empty()) { o.mode = TRPA_BIN_EMPTY; continue; }
    if (recs.size() > 1) combine(t, recs, pp->signal_majority, (medium_uint)pp->min_support_per_sequence, o);
    else {
      const Rec& r = recs.front();
      o.mode = TRPA_BIN_SINGLE;
      o.lower_node = r.lower; o.upper_node = r.upper;
      o.lower_support = r.support.back(); o.upper_support = r.support.front();
      o.length = r.qlen;
    }
    // the combined record: taxon_support_[0] = upper support (setNodeRange, predictionrecord.hh:152-158)
    auto prec_support_at = [&](uint32_t node) -> large_uint {
      const int index = (int)depth[node] - (int)depth[o.upper_node];
      if (index < 0) return 0;
      if (recs.size() == 1) return support_at(t, recs.front(), depth[node]);
      // direct mode: upper support at index 0, in-between values are path direct supports (only the upper node and
      // nodes above it are ever asked for here)
      return index == 0 ? o.upper_support : o.lower_support;
    };
    if (o.upper_node != root && pp->n_ranks) {   // binner.cpp:305-325
      const double seqlen = static_cast<double>(o.length);
      float min_pid = 0.;
      uint32_t predict_node = root;
      const uint32_t target = o.upper_node;
      const float rank_pid = prec_support_at(target) / seqlen;
      int d = depth[root];
      uint32_t pit;
      do {
        ++d;
        pit = ancestor_at(t, target, d);
        const float c = pid_per_rank[rank_of_node[pit]];
        if (c >= 0.f) min_pid = std::max(min_pid, c);
        if (rank_pid < min_pid) break;
        predict_node = pit;
      } while (pit != target);
      o.node = predict_node;
      o.support = prec_support_at(predict_node);
    } else {
      o.node = o.upper_node;
      o.support = prec_support_at(o.upper_node);
    }
  }
  return 0;
}
