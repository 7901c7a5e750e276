// The alignment-free placement models of the reference behind the same predict() interface
// (SURVEY.md 8 f4): DummyPredictionModel, LCASimplePredictionModel, MeganLCAPredictionModel,
// NBestLCAPredictionModel (core/src/taxonpredictionmodel.hh:57-259; taxator.cpp:346-361 instantiates all of
// them with treat_unclassified = false, i.e. getLCA, never getLCC).  One sub-warp GROUP of 8 lanes per segment
// (4 segments per warp in flight: the kernel is a chain of dependent loads per segment -- segment row, records,
// tree walks -- so memory-level parallelism, not bandwidth, limits it): the records of a segment are contiguous
// 36-byte rows, lane i of the group reads record i, i + 8, ...; every filter of the reference is a
// threshold on (score, e-value, node flag), so "is record i still unmasked" is recomputed from a few
// per-segment scalars instead of being stored.  HBM-bound: 36 B (+ 8 B e-value) read per record, 48 B
// written per segment.
#include "common.cuh"
#include "launch.h"
#include "machine.h"
#include <algorithm>
#include <cstdlib>

namespace trpa {

template <int GL>
__device__ __forceinline__ float group_max_f(float v) {
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int GL>   // lanes per segment: 4, 8, 16 or 32
__global__ void __launch_bounds__(256)
lca_models_kernel(const trpa_segment* __restrict__ segs, u32 n_segs, const trpa_candidate* __restrict__ cands,
                  const double* __restrict__ evalue, const uint8_t* __restrict__ uncl, const Taxonomy tax,
                  const trpa_lca_params pp, trpa_result* __restrict__ out) {
  constexpr u32 G = GL;
  const u32 lane = threadIdx.x & (G - 1u);                 // lane inside the group
  const u32 gshift = (threadIdx.x & 31u) & ~(G - 1u);      // first warp lane of the group
  const u32 gmask = G == 32u ? 0xffffffffu : ((1u << (G & 31u)) - 1u) << gshift;
  const u32 ngroups = (gridDim.x * blockDim.x) / G;
  // all groups of a warp run the same number of iterations (the shuffles below are full-warp); a group past the
  // end works on an empty dummy segment and does not store
  const u32 iters = (n_segs + ngroups - 1u) / ngroups;
  for (u32 it = 0; it < iters; ++it) {
    const u32 s = it * ngroups + (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool real = s < n_segs;
    trpa_segment sg;
    sg.query_seq = 0; sg.cand_begin = 0; sg.cand_count = 0; sg.reserved = 0;
    if (real) sg = segs[s];
    const trpa_candidate* rec = cands + sg.cand_begin;
    const double* ev = evalue ? evalue + sg.cand_begin : nullptr;
    const u32 n = sg.cand_count;
    bool classified = pp.model != TRPA_MODEL_DUMMY;
    // thresholds that define the unmasked records
    float min_score = -INFINITY;     // record masked if score < min_score   (false for NaN, like the reference)
    bool use_ms_me = false;
    if (pp.model == TRPA_MODEL_MEGAN_LCA) {
      // MinScoreMaxEvalueTopPercentFilter::filter (alignmentsfilter.hh:353-376): running maximum of the
      // records that pass (minscore, maxevalue), starting from 0; support = how often the maximum rose
      use_ms_me = true;
      float run_max = 0.f;
      u32 support = 0;
      u32 nmax = n;   // longest segment of the warp's groups
#pragma unroll
      for (int o = 16; o >= (int)G; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
      for (u32 base = 0; base < nmax; base += G) {
        const u32 i = base + lane;
        float sc = -INFINITY;
        if (i < n) {
          sc = rec[i].score;
          if (sc < pp.minscore || (ev && ev[i] > (double)pp.maxevalue)) sc = -INFINITY;
          if (sc != sc) sc = -INFINITY;   // NaN never raises the maximum
        }
        // exclusive prefix maximum over the lanes, seeded with the running maximum
        float pm = sc;
#pragma unroll
        for (int o = 1; o < (int)G; o <<= 1) {
          const float v = __shfl_up_sync(0xffffffffu, pm, o, G);
          if ((int)lane >= o) pm = fmaxf(pm, v);
        }
        float ex = __shfl_up_sync(0xffffffffu, pm, 1, G);
        if (lane == 0) ex = -INFINITY;
        ex = fmaxf(ex, run_max);
        support += __popc(__ballot_sync(0xffffffffu, sc > ex) & gmask);
        run_max = fmaxf(run_max, __shfl_sync(0xffffffffu, pm, G - 1, G));
      }
      min_score = (float)((1.0 - (double)pp.toppercent) * (double)run_max);   // alignmentsfilter.hh:369
      if (support < pp.minsupport) classified = false;                       // taxonpredictionmodel.hh:136
    } else if (pp.model == TRPA_MODEL_NBEST_LCA) {
      // NumBestBitscoreFilter::filter (alignmentsfilter.hh:497-527): keep the records of the nbest largest
      // distinct scores; `--count <= 0` on an unsigned: nbest == 0 masks nothing
      if (pp.nbest > 0) {
        float cur = INFINITY;
        // (the loop is warp-uniform: a group that has run out of distinct values idles until all have)
        bool done = false;
        for (u32 k = 0; k < pp.nbest; ++k) {
          float best = -INFINITY;
          if (!done) for (u32 i = lane; i < n; i += G) { const float sc = rec[i].score; if (sc < cur) best = fmaxf(best, sc); }
          best = group_max_f<GL>(best);
          if (best == -INFINITY) done = true;   // fewer distinct values than nbest
          else cur = best;
          if (__all_sync(0xffffffffu, done)) break;
        }
        if (cur != INFINITY) min_score = cur;
      }
    }
    // LCASimplePredictionModel::predict (taxonpredictionmodel.hh:73-125) over the unmasked records
    u32 qlo = 0xffffffffu, qhi = 0, node_all = TRPA_NO_NODE;
    float maxscore = -INFINITY;
    u32 any = 0;
    auto alive = [&](u32 i, const trpa_candidate& r) -> bool {
      if (r.score < min_score) return false;
      if (use_ms_me) {
        if (r.score < pp.minscore || (ev && ev[i] > (double)pp.maxevalue)) return false;
        if (pp.ignore_unclassified && uncl && uncl[r.node]) return false;
      }
      return true;
    };
    if (classified) {
      for (u32 i = lane; i < n; i += G) {
        const trpa_candidate r = rec[i];
        if (!alive(i, r)) continue;
        any = 1;
        qlo = min(qlo, min(r.qstart, r.qstop));
        qhi = max(qhi, max(r.qstart, r.qstop));
        maxscore = fmaxf(maxscore, r.score);
        node_all = node_all == TRPA_NO_NODE ? r.node : tx_lca(tax, node_all, r.node);
      }
    }
    any = (__ballot_sync(0xffffffffu, any) & gmask) != 0u;
    maxscore = group_max_f<GL>(maxscore);
    trpa_result res;
    res.ival = -1.f; res.signal = 0.f; res.n_pass0 = res.n_pass1 = res.n_pass2 = 0; res.cells = 0;
    const bool placed = classified && any;
    u32 node_best = TRPA_NO_NODE;
    if (placed)
      for (u32 i = lane; i < n; i += G) {
        const trpa_candidate r = rec[i];
        if (alive(i, r) && r.score == maxscore) node_best = node_best == TRPA_NO_NODE ? r.node : tx_lca(tax, node_best, r.node);
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {   // full-warp shuffles, group-local partners (o < G)
      qlo = min(qlo, __shfl_xor_sync(0xffffffffu, qlo, o));
      qhi = max(qhi, __shfl_xor_sync(0xffffffffu, qhi, o));
      const u32 oa = __shfl_xor_sync(0xffffffffu, node_all, o);
      const u32 ob = __shfl_xor_sync(0xffffffffu, node_best, o);
      if (oa != TRPA_NO_NODE) node_all = node_all == TRPA_NO_NODE ? oa : tx_lca(tax, node_all, oa);
      if (ob != TRPA_NO_NODE) node_best = node_best == TRPA_NO_NODE ? ob : tx_lca(tax, node_best, ob);
    }
    if (!placed) {
      // setUnclassified after initPredictionRecord: root, support 0; the feature range stays the whole query
      // (1..query length of the record, which the host knows: like TRPA_KIND_NONE of the RPA path)
      res.qrstart = 1; res.qrstop = 0;
      res.lower_node = res.upper_node = res.rtax_node = tax.root;
      res.support = 0; res.kind = TRPA_KIND_NONE;
    } else {
      res.qrstart = qlo; res.qrstop = qhi;
      res.lower_node = res.upper_node = node_all;
      res.rtax_node = node_best == TRPA_NO_NODE ? node_all : node_best;   // NaN maximum: no best set, like size()==size()
      res.support = qhi - qlo + 1u;                                      // setNodePoint(node): getQueryFeatureWidth()
      res.kind = TRPA_KIND_LCA;
    }
    if (lane == 0 && real) out[s] = res;
  }
}

cudaError_t launch_lca_models(const trpa_segment* segs, u32 n_segs, const trpa_candidate* cands, const double* evalue,
                              const uint8_t* uncl, const Taxonomy& tax, const trpa_lca_params& pp, trpa_result* out,
                              cudaStream_t stream) {
  if (n_segs == 0) return cudaSuccess;
  // one group per segment, as many CTAs as that takes: the hardware scheduler streams the small CTAs (no tail of a
  // persistent loop); TRPA_LCA_GROUP picks the lanes per segment (tuning hook, default 8)
  static const int gl = [] { const char* e = getenv("TRPA_LCA_GROUP"); const int v = e ? atoi(e) : 8; return (v == 4 || v == 16 || v == 32) ? v : 8; }();
  const u64 threads = (u64)n_segs * (u64)gl;
  const u32 blocks = (u32)std::min<u64>((threads + 255) / 256, 1u << 30);
  switch (gl) {
    case 4: lca_models_kernel<4><<<blocks, 256, 0, stream>>>(segs, n_segs, cands, evalue, uncl, tax, pp, out); break;
    case 16: lca_models_kernel<16><<<blocks, 256, 0, stream>>>(segs, n_segs, cands, evalue, uncl, tax, pp, out); break;
    case 32: lca_models_kernel<32><<<blocks, 256, 0, stream>>>(segs, n_segs, cands, evalue, uncl, tax, pp, out); break;
    default: lca_models_kernel<8><<<blocks, 256, 0, stream>>>(segs, n_segs, cands, evalue, uncl, tax, pp, out); break;
  }
  return cudaGetLastError();
}

}  // namespace trpa
