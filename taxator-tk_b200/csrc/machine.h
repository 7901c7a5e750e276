// Resumable per-segment placement state machine: the control flow of
//   RPAPredictionModel::predict()   core/src/taxonpredictionmodelsequence.hh:341-838
// re-expressed so that one GPU thread owns one query segment and *yields* whenever it needs a
// pairwise alignment that has not been computed yet.  A whole batch of segments advances in
// lock-step rounds: decide kernel (this file) -> stage kernel -> alignment kernels -> decide ...
// All floating-point expressions keep the reference's operand types (float vs double vs int) so
// that every comparison resolves identically; compile device code with --fmad=false.
//
// The same header compiles for the host (plain C++) so that tests can single-step the machine
// against the oracle without a GPU; the product only ever runs it inside decide_kernel.
#pragma once
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include "../../include/taxator_rpa_b200.h"
#include "shapes.h"

#include "types.h"

namespace trpa {

TRPA_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  const uint32_t o = *p; *p += v; return o;  // host build: single-threaded test harness
#endif
}
#define TRPA_ATOMIC_ADD_U32(p, v) atomic_add_u32((p), (v))

enum Phase : uint32_t {
  PH_INIT = 0,
  PH_P0_WAIT,      // pass-0 alignments in flight
  PH_P1_WAIT,      // pass-1 alignment (i <-> anchor) in flight
  PH_P2_WAIT_SEG,  // pass-2 alignment (i <-> anchor) in flight
  PH_P2_WAIT_QRY,  // pass-2 alignment (anchor <-> query) in flight
  PH_DONE
};

// candidate flag bits
constexpr uint8_t CF_QGROUP = 1, CF_OUTGROUP = 2, CF_STAGED = 4, CF_P0_ALIGNED = 8;

struct SegState {
  uint32_t phase;
  uint32_t n, cbeg, query_seq;
  uint32_t qrstart, qrstop, qrlength;
  uint32_t anchors_support, rtax, lca_all;
  uint32_t lnode_g, unode_g;
  float ival_global, bandfactor_max;
  uint32_t lca_root_dist_min;
  // pass-1 / pass-2 loop state
  uint32_t anchor, i;
  float qdist, ldist, udist;
  uint32_t rnode, lnode;
  double qpid_upper, thr_guarantee, thr_heur;
  int32_t score_thr_i;   // pass 1: int threshold (hh:602)
  float score_thr_f;     // pass 2: float threshold (hh:754)
  double qpid_thresh2;   // pass 2
  float pend_dist;       // pass 2: distance i<->anchor kept while anchor<->query is in flight
  uint32_t og_n, bf_n;
  uint32_t c0, c1, c2;
  uint64_t cells;
};

struct Taxonomy {
  const uint32_t* parent;
  const uint32_t* left;
  const uint32_t* right;
  const uint8_t* depth;
  uint32_t root;
};

// Work queues of one round (device memory); counters[] is what the host reads back.
enum Counter : uint32_t { CN_PAIRS = 0, CN_STAGE, CN_ACTIVE, CN_ARENA, CN_OVERFLOW, CN_TRACE, CN_END };
constexpr uint32_t kNumCounters = 8;

struct Batch {
  // inputs
  const trpa_segment* segs;
  const trpa_candidate* cands;  // stably sorted per segment by descending (score, identities)
  uint32_t n_segs, n_cands;
  const uint32_t* q_len;  // sequence lengths of the stores
  const uint32_t* r_len;
  Taxonomy tax;
  int protein;
  float exclude_factor;
  float reeval_bandwidth_factor;
  // per-segment / per-candidate state
  SegState* st;
  float* qd;
  float* qsim;
  uint8_t* cflags;
  uint32_t* tag;     // per candidate: anchor index + 1 whose distance its result slot holds (0: none)
  uint32_t spec_k;   // look-ahead budget: extra pass-1/2 alignments a segment may request per round
  uint32_t* og_i;
  int32_t* og_d;
  float* bf_d;       // n+1 per segment, base cbeg + segment index
  uint32_t* bf_node;
  // alignment results: slot < n_cands -> candidate slot (pass 0); n_cands + s -> segment slot
  const int32_t* res_nt;     // edit distance
  const int32_t* res_aa;     // int2 {mutual, #diag}
  // staged sequence table: [s] query segment of s; [n_segs + c] candidate c
  SeqDesc* descs;
  uint32_t arena_base;       // first unit of this pipeline's arena region
  uint32_t arena_capacity;   // units of the region: words (NT) / bytes (AA)
  // queues
  PairDesc* pairs;
  StageReq* stage;
  uint32_t* counters;
  trpa_result* results;
  // optional alignment trace (verbose log): entries appended at counters[CN_TRACE], nullptr = off
  trpa_trace_entry* trace;
  uint32_t trace_capacity;
};

TRPA_HD uint32_t tx_lca(const Taxonomy& t, uint32_t A, uint32_t B) {
  // taxonomyinterface.cpp:67-77, incl. min(A.left, B.right)
  const uint32_t left_min = t.left[A] < t.right[B] ? t.left[A] : t.right[B];
  const uint32_t right_max = t.right[A] > t.right[B] ? t.right[A] : t.right[B];
  uint32_t x = A;
  while (t.left[x] > left_min || t.right[x] < right_max) x = t.parent[x];
  return x;
}
TRPA_HD bool tx_is_parent_of(const Taxonomy& t, uint32_t A, uint32_t B) {
  return t.right[A] > t.left[B] && t.left[A] < t.left[B];  // taxonomyinterface.cpp:52-55
}

struct Machine {
  const Batch& B;
  uint32_t s;        // segment index
  SegState& S;
  const trpa_candidate* rec;
  float* qd; float* qsim; uint8_t* fl;

  TRPA_HD Machine(const Batch& b, uint32_t seg)
      : B(b), s(seg), S(b.st[seg]), rec(b.cands + b.segs[seg].cand_begin), qd(b.qd + b.segs[seg].cand_begin),
        qsim(b.qsim + b.segs[seg].cand_begin), fl(b.cflags + b.segs[seg].cand_begin) {}
  // the segment's state in a caller-owned copy (the decide kernel keeps it in registers / local memory for the round
  // and writes it back once)
  TRPA_HD Machine(const Batch& b, uint32_t seg, SegState& state)
      : B(b), s(seg), S(state), rec(b.cands + b.segs[seg].cand_begin), qd(b.qd + b.segs[seg].cand_begin),
        qsim(b.qsim + b.segs[seg].cand_begin), fl(b.cflags + b.segs[seg].cand_begin) {}

  // ---- segment coordinates: hh:856-880 + store clipping (sequencestorage.hh:353, faidx.h:325-331)
  TRPA_HD void stage_candidate(uint32_t i) {
    if (fl[i] & CF_STAGED) return;
    fl[i] |= CF_STAGED;
    const trpa_candidate& r = rec[i];
    const uint32_t left_ext = r.qstart - S.qrstart, right_ext = S.qrstop - r.qstop;
    uint64_t ns, ne; uint32_t rev = 0;
    if (r.rstart <= r.rstop) { ns = left_ext < r.rstart ? r.rstart - left_ext : 1; ne = (uint64_t)r.rstop + right_ext; }
    else { ns = right_ext < r.rstop ? r.rstop - right_ext : 1; ne = (uint64_t)r.rstart + left_ext; rev = B.protein ? 0u : 1u; }
    emit_stage(B.n_segs + S.cbeg + i, 1, r.ref_seq, ns, ne, B.r_len[r.ref_seq], rev);
  }
  TRPA_HD void emit_stage(uint32_t desc, uint32_t store, uint32_t seq, uint64_t start1, uint64_t stop1, uint32_t seqlen,
                          uint32_t rev) {
    if (stop1 > seqlen) stop1 = seqlen;
    uint64_t b = start1 - 1; if (b > seqlen) b = seqlen;
    uint64_t e = stop1 > b ? stop1 : b; if (e > seqlen) e = seqlen;
    const uint32_t len = (uint32_t)(e - b);
    const uint32_t units = B.protein ? ((len + 3u) & ~3u) : ((len + 31u) >> 5);
    const uint32_t woff = TRPA_ATOMIC_ADD_U32(&B.counters[CN_ARENA], units);
    if ((uint64_t)woff + units > B.arena_capacity) { TRPA_ATOMIC_ADD_U32(&B.counters[CN_OVERFLOW], 1u); }
    B.descs[desc] = SeqDesc{B.arena_base + woff, len, 0u, 0u};
    const uint32_t q = TRPA_ATOMIC_ADD_U32(&B.counters[CN_STAGE], 1u);
    B.stage[q] = StageReq{desc, store, seq, (uint32_t)b, rev};
  }
  // a = A (row 0 / horizontal), b = B (row 1 / vertical) as passed to getAlignment(A, B)
  // hint: estimate of the distance (0 = none).  It only seeds the band of the edit-distance kernel,
  // which verifies it and widens the band when it was too small: results never depend on it.
  TRPA_HD void emit_pair(uint32_t da, uint32_t db, uint32_t slot, uint32_t hint = 0u) {
    const uint32_t q = TRPA_ATOMIC_ADD_U32(&B.counters[CN_PAIRS], 1u);
    B.pairs[q] = PairDesc{da, db, slot, hint, 0u};
  }
  // the aligner's own alignment of record i (alnlen columns, `identities` matches) is an edit script
  // of its query range; scaled to the whole query range of the segment
  TRPA_HD uint32_t hint_query(uint32_t i) const {
    const trpa_candidate& r = rec[i];
    if (r.alnlen == 0 || r.identities > r.alnlen) return 0u;
    const uint64_t e = (uint64_t)(r.alnlen - r.identities) * S.qrlength / r.alnlen;
    return (uint32_t)(e < 0x7fffffffu ? e + 1u : 0x7fffffffu);
  }
  // segment i <-> segment j: triangle inequality over the query, exact query distances where known.
  // Bit 31 (kHintIsBound): both query distances are exact edit distances of the very strings the pair
  // aligns, so the sum is a true upper bound and the planner adds no safety margin.  CF_P0_ALIGNED marks
  // exactly the records whose qd[] is such a distance: pass 2 clears it when it overwrites qd[i] with the
  // distance to an outgroup anchor (hh:785) and sets it when it realigns an anchor against the query.
  TRPA_HD uint32_t hint_pair(uint32_t i, uint32_t j) const {
    if (B.protein) return 0u;
    const bool xi = qd[i] != FLT_MAX && (fl[i] & CF_P0_ALIGNED), xj = qd[j] != FLT_MAX && (fl[j] & CF_P0_ALIGNED);
    const float di = xi ? qd[i] : (float)hint_query(i);
    const float dj = xj ? qd[j] : (float)hint_query(j);
    const float e = di + dj;
    if (!(e < 1.0e9f)) return 0x7fffffffu;
    return ((uint32_t)e + 1u) | (xi && xj ? kHintIsBound : 0u);
  }
  // cells are counted when an alignment is consumed, i.e. exactly for the alignments the reference
  // performs; look-ahead results that are never consumed do not count
  TRPA_HD void count_cells(uint32_t da, uint32_t db) { S.cells += (uint64_t)B.descs[da].len * B.descs[db].len; }

  // Look-ahead (pass 1): while the GPU has idle capacity, also request the alignments of the next
  // records that would qualify under the CURRENT thresholds.  Decisions are still replayed strictly
  // in order from exact distances, so results are identical; a look-ahead result is simply found in
  // the record's slot (tag == anchor + 1) when the loop gets there.
  TRPA_HD void speculate_p1(uint32_t j0) {
    uint32_t budget = B.spec_k;
    const double thr = S.thr_guarantee > S.thr_heur ? S.thr_guarantee : S.thr_heur;
    for (uint32_t j = j0; budget && j < S.n && rec[j].score >= S.score_thr_i; ++j) {
      if (j == S.anchor || qd[j] == .0f) continue;
      if (!(static_cast<double>(qsim[j]) / S.qrlength >= thr)) continue;
      if (pair_slot(j, S.anchor) != 0xffffffffu) continue;
      stage_candidate(j);
      emit_pair(desc_cand(j), desc_cand(S.anchor), S.cbeg + j, hint_pair(j, S.anchor));
      B.tag[S.cbeg + j] = S.anchor + 1u;
      --budget;
    }
  }
  TRPA_HD void speculate_p2(uint32_t j0) {
    uint32_t budget = B.spec_k;
    for (uint32_t j = j0; budget && j < S.n && rec[j].score >= S.score_thr_f; ++j) {
      if (j == S.anchor) continue;
      if (!(static_cast<double>(qsim[j]) / S.qrlength >= S.qpid_thresh2)) continue;
      const uint32_t cnode = rec[j].node;
      if (tx_is_parent_of(B.tax, S.unode_g, cnode) || cnode == S.unode_g) continue;
      if (pair_slot(j, S.anchor) != 0xffffffffu) continue;
      stage_candidate(j);
      emit_pair(desc_cand(j), desc_cand(S.anchor), S.cbeg + j, hint_pair(j, S.anchor));
      B.tag[S.cbeg + j] = S.anchor + 1u;
      --budget;
    }
  }
  TRPA_HD uint32_t desc_cand(uint32_t i) const { return B.n_segs + S.cbeg + i; }
  // Result slot that already holds the alignment of segment i with segment `anchor`, or ~0: record
  // i's own slot (tag = anchor + 1: computed for this anchor by pass 1, pass 2 or look-ahead) or --
  // nucleotide only, the edit distance and the derived match count are symmetric -- the slot of
  // `anchor` when that record was itself aligned against i while i was the anchor.
  TRPA_HD uint32_t pair_slot(uint32_t i, uint32_t anchor) const {
    if (B.tag[S.cbeg + i] == anchor + 1u) return S.cbeg + i;
    if (!B.protein && B.tag[S.cbeg + anchor] == i + 1u) return S.cbeg + anchor;
    return 0xffffffffu;
  }

  // distance/similarity of a finished alignment (hh:133-171 / hh:173-242)
  TRPA_HD void read_alignment(uint32_t slot, uint32_t da, uint32_t db, float& dist, float& sim) const {
    const uint32_t la = B.descs[da].len, lb = B.descs[db].len;
    if (B.trace) {   // every call is one alignment the reference computes at this point of its control flow
      const uint32_t k = TRPA_ATOMIC_ADD_U32(&B.counters[CN_TRACE], 1u);
      if (k < B.trace_capacity) {
        trpa_trace_entry e;
        e.seg = s;
        e.a = da < B.n_segs ? TRPA_TRACE_QUERY : da - B.n_segs - S.cbeg;
        e.b = db < B.n_segs ? TRPA_TRACE_QUERY : db - B.n_segs - S.cbeg;
        e.r0 = B.protein ? B.res_aa[2 * slot] : B.res_nt[slot];
        e.r1 = B.protein ? B.res_aa[2 * slot + 1] : 0;
        e.len_a = la; e.len_b = lb;
        e.self = B.protein ? B.descs[da].pad + B.descs[db].pad : 0u;
        B.trace[k] = e;
      }
    }
    if (!B.protein) {
      const int d = B.res_nt[slot];
      const int llong = (int)(la > lb ? la : lb), lshort = (int)(la > lb ? lb : la);
      const int lendiff = llong - lshort;
      const int mismatch = d - lendiff;
      const int match = lshort - mismatch;
      dist = (float)d;
      sim = (float)match;
    } else {
      const int mutual = B.res_aa[2 * slot], nd = B.res_aa[2 * slot + 1];
      const int self = (int)B.descs[da].pad + (int)B.descs[db].pad;
      const unsigned int len = la + lb - (unsigned int)nd;
      const float norm = len / static_cast<float>(self);
      dist = (self - 2 * mutual) * norm;
      sim = (2 * mutual) * norm;
    }
  }

  TRPA_HD void finish(uint32_t kind, uint32_t lower, uint32_t upper, uint32_t support, float ival, uint32_t rtax) {
    trpa_result& R = B.results[s];
    R.qrstart = S.qrstart; R.qrstop = S.qrstop;
    R.lower_node = lower; R.upper_node = upper; R.rtax_node = rtax; R.support = support;
    R.ival = ival; R.signal = 0.f;
    R.n_pass0 = S.c0; R.n_pass1 = S.c1; R.n_pass2 = S.c2; R.kind = kind; R.cells = S.cells;
    S.phase = PH_DONE;
  }

  // ---- BandFactor (hh:259-323) over bf_d/bf_node[0..bf_n)
  TRPA_HD float band_factor() {
    float* d = B.bf_d + S.cbeg + s;
    uint32_t* nd = B.bf_node + S.cbeg + s;
    const uint32_t cnt = S.bf_n;
    // ascending by distance from element 1 on (insertion sort; the result of setBandFactor does not
    // depend on the order inside a run of equal distances)
    for (uint32_t a = 2; a < cnt; ++a) {
      const float kd = d[a]; const uint32_t kn = nd[a];
      uint32_t b = a;
      while (b > 1 && d[b - 1] > kd) { d[b] = d[b - 1]; nd[b] = nd[b - 1]; --b; }
      d[b] = kd; nd[b] = kn;
    }
    float bf = 1.f;
    const uint32_t anchor = nd[0];
    uint32_t last_rank = B.tax.depth[anchor];
    float worst[64];
    uint64_t have = 0;
    worst[last_rank] = d[0]; have |= 1ull << last_rank;
    for (uint32_t a = 1; a < cnt; ++a) {
      const float sc = d[a];
      const uint32_t rank = B.tax.depth[tx_lca(B.tax, nd[a], anchor)];
      if (rank == last_rank) continue;
      if (rank < last_rank) { worst[rank] = sc; have |= 1ull << rank; last_rank = rank; continue; }
      for (int r = (int)rank - 1; r >= 0; --r) {
        if ((have >> r) & 1ull) {
          const float ref = worst[r];
          if (ref) { const float q = sc / ref; if (q > bf) bf = q; }
        }
      }
    }
    if (bf > FLT_MAX) bf = FLT_MAX;
    return sqrtf(bf);
  }
  TRPA_HD void bf_add(float dist, uint32_t node) {
    B.bf_d[S.cbeg + s + S.bf_n] = dist;
    B.bf_node[S.cbeg + s + S.bf_n] = node;
    ++S.bf_n;
  }

  TRPA_HD uint32_t first_flag(uint8_t bit) const {
    for (uint32_t i = 0; i < S.n; ++i) if (fl[i] & bit) return i;
    return 0xffffffffu;
  }

  // ================================================================================ main entry
  // Runs until the segment needs an alignment result (returns with phase == *_WAIT) or is done.
  TRPA_HD void advance() {
    const Taxonomy& T = B.tax;
    const uint32_t root = T.root;
    const uint32_t n = B.segs[s].cand_count;
    const uint32_t slot_seg = B.n_cands + s;

    if (S.phase == PH_DONE) return;

    if (S.phase == PH_INIT) {
      const trpa_segment sg = B.segs[s];
      S.n = sg.cand_count; S.cbeg = sg.cand_begin; S.query_seq = sg.query_seq;
      S.c0 = S.c1 = S.c2 = 0; S.cells = 0;
      S.qrstart = S.qrstop = 0;
      const uint32_t nn = sg.cand_count;
      if (nn == 0) { finish(TRPA_KIND_NONE, root, root, 0, -2.f, root); return; }          // hh:359-368
      if (nn == 1) {                                                                         // hh:371-388
        S.qrstart = rec[0].qstart; S.qrstop = rec[0].qstop;
        finish(TRPA_KIND_SINGLE, rec[0].node, root, rec[0].identities, 1.f, rec[0].node);
        return;
      }
      uint32_t qs = rec[0].qstart, qe = rec[0].qstop;                                        // hh:391-404
      for (uint32_t i = 1; i < nn; ++i) {
        if (rec[i].qstart < qs) qs = rec[i].qstart;
        if (rec[i].qstop > qe) qe = rec[i].qstop;
      }
      S.qrstart = qs; S.qrstop = qe; S.qrlength = qe - qs + 1;
      const uint32_t qrlength = S.qrlength;
      if (rec[0].alnlen == qrlength && rec[0].identities == qrlength) {                      // hh:431-472
        const float best = rec[0].score;
        uint32_t lnode = rec[0].node, unode = TRPA_NO_NODE, i = 1;
        while (true) {
          if (i == nn) { unode = root; break; }
          const float sc = rec[i].score;
          if (sc == best) lnode = tx_lca(T, lnode, rec[i].node);
          else {
            const float us = sc;
            unode = lnode;
            do { unode = tx_lca(T, unode, rec[i].node); } while (++i < nn && rec[i].score == us);
            break;
          }
          ++i;
        }
        finish(TRPA_KIND_IDENTICAL, lnode, unode, qrlength, 0.f, lnode);
        return;
      }
      // pass 0 set-up (hh:497-539): which records get realigned against the query does not depend
      // on any alignment result, so all of them are issued in one round.
      const float thr = B.reeval_bandwidth_factor * rec[0].score;
      bool query_staged = false;
      for (uint32_t i = 0; i < nn; ++i) {
        fl[i] = 0;
        B.tag[S.cbeg + i] = 0;
        if (rec[i].alnlen == qrlength && rec[i].identities == qrlength) {
          fl[i] = CF_QGROUP;
        } else if (rec[i].score >= thr) {
          fl[i] = CF_QGROUP | CF_P0_ALIGNED;
          if (!query_staged) {   // hh:415 (query store: in-memory, sequencestorage.hh:105-120)
            emit_stage(s, 0, S.query_seq, qs, qe, B.q_len[S.query_seq], 0);
            query_staged = true;
          }
          stage_candidate(i);
          emit_pair(desc_cand(i), s, S.cbeg + i, B.protein ? 0u : hint_query(i));
          count_cells(desc_cand(i), s);
          ++S.c0;
        }
      }
      if (!query_staged) emit_stage(s, 0, S.query_seq, qs, qe, B.q_len[S.query_seq], 0);
      S.phase = PH_P0_WAIT;
      if (S.c0 > 0) return;
      // no alignment needed in pass 0: fall through
    }

    // every later phase re-enters one of the loops below
    bool resume_p1 = false, resume_p2 = false;
    float dist = 0.f;

    if (S.phase == PH_P0_WAIT) {
      uint32_t ibest = 0;
      uint32_t support = 0;
      uint32_t lca_all = rec[0].node;
      for (uint32_t i = 0; i < n; ++i) {                                                     // hh:502-549
        float dd, sim;
        if (fl[i] & CF_P0_ALIGNED) {
          float asim;
          read_alignment(S.cbeg + i, desc_cand(i), s, dd, asim);
          const float idf = static_cast<float>(rec[i].identities);
          sim = asim < idf ? idf : asim;   // std::max(a, b): b only if a < b
        } else if (fl[i] & CF_QGROUP) { dd = 0; sim = rec[i].identities; }
        else { dd = FLT_MAX; sim = rec[i].identities; }
        qd[i] = dd; qsim[i] = sim;
        if (dd < qd[ibest]) ibest = i;
        else if (dd == qd[ibest]) {
          if (sim > qsim[ibest]) ibest = i;
          else if (sim == qsim[ibest] && rec[i].score > rec[ibest].score) ibest = i;
        }
        const uint32_t simu = static_cast<uint32_t>(sim);
        if (simu > support) support = simu;
        lca_all = tx_lca(T, lca_all, rec[i].node);
      }
      S.anchors_support = support; S.lca_all = lca_all;
      uint32_t rtax = rec[ibest].node;                                                       // hh:553-562
      for (uint32_t i = 0; i < n; ++i) {
        if (!(fl[i] & CF_QGROUP)) continue;
        if (qd[i] != qd[ibest] || qsim[i] != qsim[ibest] || rec[i].score != rec[ibest].score) fl[i] &= ~CF_QGROUP;
        else rtax = tx_lca(T, rtax, rec[i].node);
      }
      S.rtax = rtax;
      S.ival_global = 0.f; S.lnode_g = rtax; S.unode_g = rtax; S.bandfactor_max = 1.f;
      S.lca_root_dist_min = 255;
      goto p1_anchor_begin;
    }
    if (S.phase == PH_P1_WAIT) {
      float sim;
      read_alignment(S.cbeg + S.i, desc_cand(S.i), desc_cand(S.anchor), dist, sim);
      count_cells(desc_cand(S.i), desc_cand(S.anchor));
      ++S.c1;
      resume_p1 = true;
      goto p1_loop;
    }
    if (S.phase == PH_P2_WAIT_SEG) {
      float sim;
      read_alignment(S.cbeg + S.i, desc_cand(S.i), desc_cand(S.anchor), dist, sim);
      count_cells(desc_cand(S.i), desc_cand(S.anchor));
      ++S.c2;
      qd[S.i] = dist;              // hh:785: no longer a distance to the query
      fl[S.i] &= ~CF_P0_ALIGNED;   // ... so hint_pair must not treat it as one
      resume_p2 = true;
      goto p2_loop;
    }
    if (S.phase == PH_P2_WAIT_QRY) {
      resume_p2 = true;
      goto p2_loop;
    }
    return;

  // ------------------------------------------------------------------------- pass 1, hh:576-733
  p1_anchor_begin: {
      const uint32_t anchor = first_flag(CF_QGROUP);
      fl[anchor] &= ~CF_QGROUP;
      S.anchor = anchor;
      S.qdist = qd[anchor];
      S.rnode = rec[anchor].node;
      S.bf_n = 0;
      bf_add(0.f, S.rnode);
      S.lnode = S.rtax;
      S.ldist = 0.f; S.udist = FLT_MAX;
      S.og_n = 0;
      S.qpid_upper = 0.; S.thr_guarantee = 0.; S.thr_heur = 0.;
      S.score_thr_i = 0;
      S.i = 0;
    }
  p1_loop:
    for (;;) {
      uint32_t i = S.i;
      if (!resume_p1) {
        if (!(S.lnode != root && i < n && rec[i].score >= S.score_thr_i)) break;
      }
      {
        const uint32_t cnode = rec[i].node;
        const double qsearchpid = static_cast<double>(rec[i].identities) / S.qrlength;
        bool take = true;
        if (!resume_p1) {
          const double qpid = static_cast<double>(qsim[i]) / S.qrlength;
          const double qpid_thresh = S.thr_guarantee > S.thr_heur ? S.thr_guarantee : S.thr_heur;  // std::max(a,b)
          take = qpid >= qpid_thresh;
          if (take) {
            if (i == S.anchor) dist = .0f;
            else if (qd[i] == .0f) dist = qd[S.anchor];
            else if (pair_slot(i, S.anchor) != 0xffffffffu) {  // already computed (look-ahead, or as anchor <-> i)
              float sim;
              read_alignment(pair_slot(i, S.anchor), desc_cand(i), desc_cand(S.anchor), dist, sim);
              count_cells(desc_cand(i), desc_cand(S.anchor));
              ++S.c1;
            } else {
              stage_candidate(S.anchor);
              stage_candidate(i);
              emit_pair(desc_cand(i), desc_cand(S.anchor), S.cbeg + i, hint_pair(i, S.anchor));
              B.tag[S.cbeg + i] = S.anchor + 1u;
              speculate_p1(i + 1);
              S.phase = PH_P1_WAIT;
              return;
            }
          }
        }
        resume_p1 = false;
        if (take) {
          bf_add(dist, cnode);
          if (dist == .0f) fl[i] &= ~CF_QGROUP;
          else if (dist <= S.qdist) {
            S.lnode = tx_lca(T, S.lnode, cnode);
            if (dist > S.ldist) S.ldist = dist;
          } else {
            if (dist < S.udist) {
              S.udist = dist;
              if (qsearchpid > S.qpid_upper) {
                S.qpid_upper = qsearchpid;
                S.thr_guarantee = qsearchpid * 2. - 1.;
                S.thr_heur = qsearchpid * B.exclude_factor;
              }
              if (!S.score_thr_i) S.score_thr_i = (int32_t)(rec[i].score * B.exclude_factor);
            }
            B.og_i[S.cbeg + S.og_n] = i;
            B.og_d[S.cbeg + S.og_n] = (int32_t)dist;  // tuple<uint,int>: truncation (hh:592,661)
            ++S.og_n;
          }
        }
      }
      S.i = i + 1;
    }
    {  // anchor epilogue, hh:667-727
      const float bandfactor = band_factor();
      if (bandfactor > S.bandfactor_max) S.bandfactor_max = bandfactor;
      const float qdist = S.qdist;
      const float qdist_ex = qdist * bandfactor;
      float min_upper = (float)INT_MAX;
      uint32_t* ogi = B.og_i + S.cbeg;
      int32_t* ogd = B.og_d + S.cbeg;
      uint32_t keep = 0;
      for (uint32_t k = 0; k < S.og_n; ++k) {
        const float dd = (float)ogd[k];
        bool erase = false;
        if (dd > qdist_ex) {
          if (dd > min_upper) erase = true;
          else if (dd < min_upper) min_upper = dd;
        } else {
          if (min_upper > qdist_ex) min_upper = dd;
          else min_upper = min_upper < dd ? dd : min_upper;
        }
        if (!erase) { ogi[keep] = ogi[k]; ogd[keep] = ogd[k]; ++keep; }
      }
      uint32_t unode = S.lnode;  // min_upper != FLT_MAX always holds (hh:670 vs :690)
      for (uint32_t k = 0; k < keep; ++k) {
        const float dd = (float)ogd[k];
        const uint32_t ci = ogi[k];
        const uint32_t cnode = rec[ci].node;
        if (dd > min_upper) continue;
        unode = tx_lca(T, cnode, unode);
        const uint32_t lrd = T.depth[tx_lca(T, cnode, S.rtax)];
        if (lrd > S.lca_root_dist_min) continue;
        else if (lrd < S.lca_root_dist_min) {
          S.lca_root_dist_min = lrd;
          for (uint32_t q = 0; q < n; ++q) fl[q] &= ~CF_OUTGROUP;
        }
        fl[ci] |= CF_OUTGROUP;
      }
      float ival = 0.f;
      if (unode != S.lnode && S.ldist < qdist) ival = (qdist - S.ldist) / (S.udist - S.ldist);
      if (ival > S.ival_global) S.ival_global = ival;  // std::max(ival, ival_global)
      S.unode_g = tx_lca(T, S.unode_g, unode);
      S.lnode_g = tx_lca(T, S.lnode_g, S.lnode);
      if (first_flag(CF_QGROUP) != 0xffffffffu && S.lnode_g != root) goto p1_anchor_begin;
    }

  // ------------------------------------------------------------------------- pass 2, hh:737-822
  p2_anchor_begin:
    for (;;) {
      const uint32_t anchor = first_flag(CF_OUTGROUP);
      if (anchor == 0xffffffffu) goto done;
      fl[anchor] &= ~CF_OUTGROUP;
      if (S.unode_g == S.lca_all) continue;
      S.anchor = anchor;
      const double qpid_anchor = static_cast<double>(qsim[anchor]) / S.qrlength;
      const double tg = qpid_anchor * 2. - 1.;
      const double th = qpid_anchor * B.exclude_factor;
      S.qpid_thresh2 = tg > th ? tg : th;
      S.score_thr_f = rec[anchor].score * B.exclude_factor;
      S.i = 0;
      break;
    }
  p2_loop:
    for (;;) {
      const uint32_t i = S.i;
      const uint32_t anchor = S.anchor;
      if (!resume_p2) {
        if (!(i < n && rec[i].score >= S.score_thr_f)) break;
      }
      {
        bool take = true;
        bool second = false;  // resuming after anchor<->query
        if (resume_p2) {
          if (S.phase == PH_P2_WAIT_QRY) { second = true; dist = S.pend_dist; }
        } else {
          const double qpid = static_cast<double>(qsim[i]) / S.qrlength;
          take = qpid >= S.qpid_thresh2;
          if (take) {
            const uint32_t cnode = rec[i].node;
            if (i == anchor) dist = .0f;
            else if (tx_is_parent_of(T, S.unode_g, cnode) || cnode == S.unode_g) take = false;  // "continue"
            else if (pair_slot(i, anchor) != 0xffffffffu) {  // already computed (look-ahead, pass 1, or as anchor <-> i)
              float sim;
              read_alignment(pair_slot(i, anchor), desc_cand(i), desc_cand(anchor), dist, sim);
              count_cells(desc_cand(i), desc_cand(anchor));
              ++S.c2;
              qd[i] = dist;
              fl[i] &= ~CF_P0_ALIGNED;
            } else {
              stage_candidate(anchor);
              stage_candidate(i);
              emit_pair(desc_cand(i), desc_cand(anchor), S.cbeg + i, hint_pair(i, anchor));
              B.tag[S.cbeg + i] = anchor + 1u;
              speculate_p2(i + 1);
              S.phase = PH_P2_WAIT_SEG;
              return;
            }
          }
        }
        resume_p2 = false;
        if (take) {
          const uint32_t cnode = rec[i].node;
          if (dist == .0f) fl[i] &= ~CF_OUTGROUP;
          else {
            float qdist_ex;
            if (second) {
              float d2, s2;
              read_alignment(slot_seg, desc_cand(anchor), s, d2, s2);
              count_cells(desc_cand(anchor), s);
              const float sim = s2 < qsim[anchor] ? qsim[anchor] : s2;  // std::max(s2, qsim[anchor])
              qd[anchor] = d2; qsim[anchor] = sim;
              fl[anchor] |= CF_P0_ALIGNED;   // an exact query distance again
              qdist_ex = d2 * S.bandfactor_max;
              ++S.c2;
            } else if (qd[anchor] == FLT_MAX) {
              stage_candidate(anchor);
              emit_pair(desc_cand(anchor), s, slot_seg, B.protein ? 0u : hint_query(anchor));
              S.pend_dist = dist;
              S.phase = PH_P2_WAIT_QRY;
              return;
            } else qdist_ex = qd[anchor] * S.bandfactor_max;
            if (dist <= qdist_ex) S.unode_g = tx_lca(T, S.unode_g, cnode);
          }
        }
      }
      S.i = i + 1;
    }
    goto p2_anchor_begin;

  done:
    if (S.unode_g == S.lnode_g) S.ival_global = 1.f;
    finish(TRPA_KIND_PLACED, S.lnode_g, S.unode_g, S.anchors_support, S.ival_global, S.rtax);
  }
};

}  // namespace trpa
