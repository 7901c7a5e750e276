// TEST HARNESS (CPU): single-steps the product's resumable placement state machine
// (taxator-tk_b200/csrc/machine.h, compiled for the host) in the same decide -> stage -> align
// rounds the GPU runs, but serves staging and alignment requests from the oracle
// (oracle/rpa_oracle.cpp).  Lets `pytest -m "not gpu"` check the host logic of the device decision
// kernel against the oracle's straight-line restatement of predict() without a GPU.
#include <cstring>
#include <vector>
#include "../taxator-tk_b200/csrc/machine.h"
#include "../taxator-tk_b200/csrc/hostprep.h"

extern "C" {
int orc_edit_distance(const uint8_t* a, int la, const uint8_t* b, int lb);
void orc_protein_align(const uint8_t* a, int la, const uint8_t* b, int lb, int* out6);
uint32_t orc_fetch_segment(const uint8_t* codes, const uint64_t* off, const uint32_t* len, uint32_t nseq, int protein,
                           uint32_t id, uint32_t start, uint32_t stop, uint32_t left_ext, uint32_t right_ext,
                           uint8_t* out);
}

using namespace trpa;

extern "C" int hm_predict_batch(
    const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth, uint32_t n_nodes, uint32_t root,
    const uint8_t* q_codes, const uint64_t* q_off, const uint32_t* q_len, uint32_t q_n,
    const uint8_t* r_codes, const uint64_t* r_off, const uint32_t* r_len, uint32_t r_n,
    int protein, float exclude_factor, float toppercent,
    const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands_in, uint32_t n_cands,
    trpa_result* results, uint32_t* rounds_out, uint32_t spec_k, trpa_trace_entry* trace, uint32_t trace_cap, uint32_t* trace_n) {
  std::vector<trpa_candidate> cands(cands_in, cands_in + n_cands);
  sort_candidates(segs, n_segs, cands.data());

  std::vector<SegState> st(n_segs);
  memset(st.data(), 0, sizeof(SegState) * n_segs);
  std::vector<float> qd(n_cands), qsim(n_cands), bf_d(n_cands + n_segs);
  std::vector<uint8_t> cflags(n_cands, 0);
  std::vector<uint32_t> tag(n_cands, 0);
  std::vector<uint32_t> og_i(n_cands), bf_node(n_cands + n_segs);
  std::vector<int32_t> og_d(n_cands);
  std::vector<int32_t> res_nt(n_cands + n_segs, 0), res_aa(2 * (size_t)(n_cands + n_segs), 0);
  std::vector<SeqDesc> descs(n_segs + n_cands);
  std::vector<PairDesc> pairs(n_cands + n_segs);
  std::vector<StageReq> stage(n_cands + n_segs);
  std::vector<uint32_t> counters(kNumCounters, 0);
  std::vector<std::vector<uint8_t>> staged(n_segs + n_cands);

  Batch B;
  B.segs = segs; B.cands = cands.data(); B.n_segs = n_segs; B.n_cands = n_cands;
  B.q_len = q_len; B.r_len = r_len;
  B.tax = Taxonomy{parent, left, right, depth, root};
  B.protein = protein; B.exclude_factor = exclude_factor;
  B.reeval_bandwidth_factor = 1. - toppercent;
  B.st = st.data(); B.qd = qd.data(); B.qsim = qsim.data(); B.cflags = cflags.data();
  B.tag = tag.data(); B.spec_k = spec_k;
  B.og_i = og_i.data(); B.og_d = og_d.data(); B.bf_d = bf_d.data(); B.bf_node = bf_node.data();
  B.res_nt = res_nt.data(); B.res_aa = res_aa.data();
  B.descs = descs.data(); B.arena_capacity = 0xffffffffu; B.arena_base = 0;
  B.pairs = pairs.data(); B.stage = stage.data(); B.counters = counters.data(); B.results = results;
  B.trace = trace; B.trace_capacity = trace_cap;

  uint32_t rounds = 0;
  for (;;) {
    counters[CN_PAIRS] = 0; counters[CN_STAGE] = 0; counters[CN_ACTIVE] = 0;
    for (uint32_t s = 0; s < n_segs; ++s) {
      if (st[s].phase == PH_DONE) continue;
      Machine M(B, s);
      M.advance();
      if (st[s].phase != PH_DONE) ++counters[CN_ACTIVE];
    }
    ++rounds;
    for (uint32_t k = 0; k < counters[CN_STAGE]; ++k) {
      const StageReq& rq = stage[k];
      const SeqDesc& d = descs[rq.desc];
      std::vector<uint8_t>& out = staged[rq.desc];
      out.resize(d.len);
      const uint8_t* src = (rq.store ? r_codes + r_off[rq.seq] : q_codes + q_off[rq.seq]) + rq.begin;
      for (uint32_t j = 0; j < d.len; ++j) {
        if (!rq.rev) out[j] = src[j];
        else { uint8_t c = src[d.len - 1 - j]; out[j] = c < 4 ? 3 - c : 4; }
      }
      if (protein) {
        // self score through the oracle's alignment of X with X
        int o6[6];
        orc_protein_align(out.data(), (int)out.size(), out.data(), (int)out.size(), o6);
        descs[rq.desc].pad = (uint32_t)(o6[1] / 2);
      }
    }
    if (counters[CN_PAIRS] == 0) {
      if (counters[CN_ACTIVE] != 0) return -1;  // stuck
      break;
    }
    for (uint32_t k = 0; k < counters[CN_PAIRS]; ++k) {
      const PairDesc& p = pairs[k];
      const std::vector<uint8_t>& A = staged[p.a];
      const std::vector<uint8_t>& Bq = staged[p.b];
      if (!protein) res_nt[p.out] = orc_edit_distance(A.data(), (int)A.size(), Bq.data(), (int)Bq.size());
      else {
        int o6[6];
        orc_protein_align(A.data(), (int)A.size(), Bq.data(), (int)Bq.size(), o6);
        res_aa[2 * p.out] = o6[0];
        res_aa[2 * p.out + 1] = (int)A.size() + (int)Bq.size() - o6[2];
      }
    }
    if (rounds > 1000000) return -2;
  }
  if (rounds_out) *rounds_out = rounds;
  if (trace_n) *trace_n = counters[CN_TRACE];
  (void)n_nodes; (void)q_n; (void)r_n;
  return 0;
}
