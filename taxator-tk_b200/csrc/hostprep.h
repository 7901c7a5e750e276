// Host-side batch preparation shared by the C-ABI glue (api.cu) and the CPU test harness.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "../../include/taxator_rpa_b200.h"

namespace trpa {

// SortFilter (core/src/alignmentsfilter.hh:171-190): list::sort (stable) with "second < first" on
// AlignmentRecord::operator< (score, then identities; core/src/alignmentrecord.hh:89-93).
inline void sort_candidates(const trpa_segment* segs, uint32_t n_segs, trpa_candidate* cands) {
  for (uint32_t s = 0; s < n_segs; ++s) {
    trpa_candidate* b = cands + segs[s].cand_begin;
    std::stable_sort(b, b + segs[s].cand_count, [](const trpa_candidate& x, const trpa_candidate& y) {
      if (y.score < x.score) return true;
      if (y.score > x.score) return false;
      return y.identities < x.identities;
    });
  }
}

// Upper bound of the staging arena units (32-base words for NT, 4-byte-rounded bytes for AA) that
// segment s can ever request: its query range plus every candidate's extended range before clipping
// (core/src/taxonpredictionmodelsequence.hh:523, :856-880).
inline uint64_t segment_arena_bound(const trpa_segment& sg, const trpa_candidate* cands, bool protein) {
  if (sg.cand_count < 2) return 0;
  const trpa_candidate* c = cands + sg.cand_begin;
  uint32_t qs = c[0].qstart, qe = c[0].qstop;
  for (uint32_t i = 1; i < sg.cand_count; ++i) { qs = std::min(qs, c[i].qstart); qe = std::max(qe, c[i].qstop); }
  auto units = [&](uint64_t len) -> uint64_t { return protein ? ((len + 3) & ~3ull) : ((len + 31) >> 5); };
  uint64_t total = units((uint64_t)qe - qs + 1);
  for (uint32_t i = 0; i < sg.cand_count; ++i) {
    const uint64_t span = (c[i].rstart <= c[i].rstop ? (uint64_t)c[i].rstop - c[i].rstart : (uint64_t)c[i].rstart - c[i].rstop) + 1;
    const uint64_t ext = (uint64_t)(c[i].qstart - qs) + (uint64_t)(qe - c[i].qstop);
    total += units(span + ext);
  }
  return total;
}

}  // namespace trpa
