// TEST INFRASTRUCTURE.  Calls the product's verbose-log writer (taxator-tk_b200/host/verbose_log.cpp) on flat arrays
// so that the CPU suite can write the log from a trace produced by the host-compiled state machine
// (tests/host_machine_harness.cpp) and compare it with the log of the real reference (`taxator -l`).
#include <fstream>
#include <string>
#include <vector>

#include "../taxator-tk_b200/host/verbose_log.h"

using namespace taxator_b200;

extern "C" int hl_write_log(const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth, uint32_t n_nodes,
                            const char* names_blob, const uint32_t* names_off,
                            const char* q_chars, const uint64_t* q_off, const uint32_t* q_len, uint32_t q_n,
                            const char* r_chars, const uint64_t* r_off, const uint32_t* r_len, uint32_t r_n,
                            int protein, float exclude_factor, float toppercent,
                            const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, const trpa_result* res,
                            const char* qid_blob, const uint32_t* qid_off,
                            const trpa_trace_entry* trace, uint32_t n_trace, const char* out_path, char* err, uint32_t err_cap) {
  try {
    FlatTaxonomy tax;
    tax.parent.assign(parent, parent + n_nodes); tax.left.assign(left, left + n_nodes); tax.right.assign(right, right + n_nodes);
    tax.depth.assign(depth, depth + n_nodes);
    tax.root = 0;
    for (uint32_t i = 0; i < n_nodes; ++i) tax.name.push_back(std::string(names_blob + names_off[i], names_blob + names_off[i + 1]));
    SeqStore q, r;
    auto fill = [](SeqStore& s, const char* chars, const uint64_t* off, const uint32_t* len, uint32_t n) {
      s.off.assign(off, off + n); s.len.assign(len, len + n);
      s.chars.assign(chars, n ? off[n - 1] + len[n - 1] : 0);
    };
    fill(q, q_chars, q_off, q_len, q_n);
    fill(r, r_chars, r_off, r_len, r_n);
    VerboseLogContext lc;
    lc.tax = &tax; lc.q_store = &q; lc.db_store = &r; lc.protein = protein != 0;
    lc.exclude_factor = exclude_factor;
    lc.reeval_bandwidth_factor = 1. - toppercent;
    std::ofstream out(out_path);
    size_t t = 0;
    for (uint32_t s = 0; s < n_segs; ++s) {
      size_t e = t;
      while (e < n_trace && trace[e].seg == s) ++e;
      write_segment_log(lc, std::string(qid_blob + qid_off[s], qid_blob + qid_off[s + 1]), segs[s].query_seq, cands + segs[s].cand_begin,
                        segs[s].cand_count, res[s], trace + t, e - t, out);
      t = e;
    }
    return 0;
  } catch (std::exception& e) {
    snprintf(err, err_cap, "%s", e.what());
    return -1;
  }
}
