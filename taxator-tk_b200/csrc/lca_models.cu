// The alignment-free placement models of the reference behind the same predict() interface
// (SURVEY.md 8 f4): DummyPredictionModel, LCASimplePredictionModel, MeganLCAPredictionModel,
// NBestLCAPredictionModel (core/src/taxonpredictionmodel.hh:57-259; taxator.cpp:346-361 instantiates all of
// them with treat_unclassified = false, i.e. getLCA, never getLCC).  One WARP per segment: the records of a
// segment are contiguous 36-byte rows, lane i reads record i, i + 32, ...; every filter of the reference is a
// threshold on (score, e-value, node flag), so "is record i still unmasked" is recomputed from a few
// per-segment scalars instead of being stored.  HBM-bound: 36 B (+ 8 B e-value) read per record, 48 B
// written per segment.
#include "common.cuh"
#include "launch.h"
#include "machine.h"

namespace trpa {

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
lca_models_kernel(const trpa_segment* __restrict__ segs, u32 n_segs, const trpa_candidate* __restrict__ cands,
                  const double* __restrict__ evalue, const uint8_t* __restrict__ uncl, const Taxonomy tax,
                  const trpa_lca_params pp, trpa_result* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_segs; s += nwarps) {
    const trpa_segment sg = segs[s];
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
      for (u32 base = 0; base < n; base += 32) {
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
        for (int o = 1; o < 32; o <<= 1) {
          const float v = __shfl_up_sync(0xffffffffu, pm, o);
          if ((int)lane >= o) pm = fmaxf(pm, v);
        }
        float ex = __shfl_up_sync(0xffffffffu, pm, 1);
        if (lane == 0) ex = -INFINITY;
        ex = fmaxf(ex, run_max);
        support += __popc(__ballot_sync(0xffffffffu, sc > ex));
        run_max = fmaxf(run_max, __shfl_sync(0xffffffffu, pm, 31));
      }
      min_score = (float)((1.0 - (double)pp.toppercent) * (double)run_max);   // alignmentsfilter.hh:369
      if (support < pp.minsupport) classified = false;                       // taxonpredictionmodel.hh:136
    } else if (pp.model == TRPA_MODEL_NBEST_LCA) {
      // NumBestBitscoreFilter::filter (alignmentsfilter.hh:497-527): keep the records of the nbest largest
      // distinct scores; `--count <= 0` on an unsigned: nbest == 0 masks nothing
      if (pp.nbest > 0) {
        float cur = INFINITY;
        for (u32 k = 0; k < pp.nbest; ++k) {
          float best = -INFINITY;
          for (u32 i = lane; i < n; i += 32) { const float sc = rec[i].score; if (sc < cur) best = fmaxf(best, sc); }
          best = warp_max_f(best);
          if (best == -INFINITY) break;   // fewer distinct values than nbest
          cur = best;
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
      for (u32 i = lane; i < n; i += 32) {
        const trpa_candidate r = rec[i];
        if (!alive(i, r)) continue;
        any = 1;
        qlo = min(qlo, min(r.qstart, r.qstop));
        qhi = max(qhi, max(r.qstart, r.qstop));
        maxscore = fmaxf(maxscore, r.score);
        node_all = node_all == TRPA_NO_NODE ? r.node : tx_lca(tax, node_all, r.node);
      }
      any = __any_sync(0xffffffffu, any);
      maxscore = warp_max_f(maxscore);
    }
    trpa_result res;
    res.ival = -1.f; res.signal = 0.f; res.n_pass0 = res.n_pass1 = res.n_pass2 = 0; res.cells = 0;
    if (!classified || !any) {
      // setUnclassified after initPredictionRecord: root, support 0; the feature range stays the whole query
      // (1..query length of the record, which the host knows: like TRPA_KIND_NONE of the RPA path)
      res.qrstart = 1; res.qrstop = 0;
      res.lower_node = res.upper_node = res.rtax_node = tax.root;
      res.support = 0; res.kind = TRPA_KIND_NONE;
    } else {
      u32 node_best = TRPA_NO_NODE;
      for (u32 i = lane; i < n; i += 32) {
        const trpa_candidate r = rec[i];
        if (alive(i, r) && r.score == maxscore) node_best = node_best == TRPA_NO_NODE ? r.node : tx_lca(tax, node_best, r.node);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        qlo = min(qlo, __shfl_xor_sync(0xffffffffu, qlo, o));
        qhi = max(qhi, __shfl_xor_sync(0xffffffffu, qhi, o));
        const u32 oa = __shfl_xor_sync(0xffffffffu, node_all, o);
        const u32 ob = __shfl_xor_sync(0xffffffffu, node_best, o);
        if (oa != TRPA_NO_NODE) node_all = node_all == TRPA_NO_NODE ? oa : tx_lca(tax, node_all, oa);
        if (ob != TRPA_NO_NODE) node_best = node_best == TRPA_NO_NODE ? ob : tx_lca(tax, node_best, ob);
      }
      res.qrstart = qlo; res.qrstop = qhi;
      res.lower_node = res.upper_node = node_all;
      res.rtax_node = node_best == TRPA_NO_NODE ? node_all : node_best;   // NaN maximum: no best set, like size()==size()
      res.support = qhi - qlo + 1u;                                      // setNodePoint(node): getQueryFeatureWidth()
      res.kind = TRPA_KIND_LCA;
    }
    if (lane == 0) out[s] = res;
  }
}

cudaError_t launch_lca_models(const trpa_segment* segs, u32 n_segs, const trpa_candidate* cands, const double* evalue,
                              const uint8_t* uncl, const Taxonomy& tax, const trpa_lca_params& pp, trpa_result* out,
                              cudaStream_t stream) {
  if (n_segs == 0) return cudaSuccess;
  u32 blocks = (n_segs + 7) / 8;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  lca_models_kernel<<<blocks, 256, 0, stream>>>(segs, n_segs, cands, evalue, uncl, tax, pp, out);
  return cudaGetLastError();
}

}  // namespace trpa
