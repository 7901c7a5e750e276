// TEST HARNESS (CPU): the product's host code (taxator-tk_b200/host: taxonomy loader, FASTA/.fai
// stores, alignment parser, record-set segmentation, candidate flattening, GFF3 printing, driver
// loop) with the GPU batch predictor swapped for the oracle (oracle/rpa_oracle.cpp).  Run on the
// same files as the real reference, its GFF3 must be identical -- checks the host logic without a
// GPU.  Usage: host_cli_harness <nucleotide|protein> mapping query.fna ref.fna ref.fna.fai [batch] [split] [sorted]
//        [fast_block_bytes]   (> 0: the fast ingest path of the CLI, ingest.h, with that block size)
#include <cstring>
#include <fstream>
#include <iostream>

#include "../taxator-tk_b200/host/driver.h"
#include "../taxator-tk_b200/host/ingest.h"

extern "C" {
struct OrcResult {
  uint32_t qrstart, qrstop, lower, upper, rtax, support;
  float ival, signal;
  uint32_t p0, p1, p2, kind;
  uint64_t cells;
};
int orc_char2dna5(int c);
int orc_char2aa(int c);
int orc_predict_segment(const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth,
                        uint32_t n_nodes, uint32_t root, const uint8_t* q_codes, const uint64_t* q_off,
                        const uint32_t* q_len, uint32_t q_n, const uint8_t* r_codes, const uint64_t* r_off,
                        const uint32_t* r_len, uint32_t r_n, int protein, float exclude_factor, float toppercent,
                        uint32_t query_seq, const void* cands, uint32_t n, OrcResult* res, void* plog, uint32_t plog_cap,
                        uint32_t* plog_n);
}

using namespace taxator_b200;

int main(int argc, char** argv) {
  if (argc < 6) { std::cerr << "usage\n"; return 2; }
  try {
    const bool protein = std::string(argv[1]) == "protein";
    const size_t batch = argc > 6 ? std::stoul(argv[6]) : 100000;
    const bool split = argc > 7 ? std::string(argv[7]) == "1" : true;
    const bool sorted = argc > 8 ? std::string(argv[8]) == "1" : false;
    FlatTaxonomy tax = load_taxonomy_from_environment(kDefaultRanks, true);
    SeqIdMapping mapping = load_mapping(argv[2]);
    SeqStore q = load_fasta_inmemory(argv[3]);
    SeqStore r = load_fasta_indexed(argv[4], argv[5]);
    std::vector<uint8_t> qc(q.chars.size()), rc(r.chars.size());
    for (size_t i = 0; i < qc.size(); ++i) qc[i] = (uint8_t)(protein ? orc_char2aa((unsigned char)q.chars[i]) : orc_char2dna5((unsigned char)q.chars[i]));
    for (size_t i = 0; i < rc.size(); ++i) rc[i] = (uint8_t)(protein ? orc_char2aa((unsigned char)r.chars[i]) : orc_char2dna5((unsigned char)r.chars[i]));
    const size_t fast_bytes = argc > 9 ? std::stoul(argv[9]) : 0;
    auto oracle_segment = [&](uint32_t query_seq, const trpa_candidate* c, uint32_t n, trpa_result* out) {
      OrcResult o;
      orc_predict_segment(tax.parent.data(), tax.left.data(), tax.right.data(), tax.depth.data(), (uint32_t)tax.size(),
                          tax.root, qc.data(), q.off.data(), q.len.data(), (uint32_t)q.size(), rc.data(), r.off.data(),
                          r.len.data(), (uint32_t)r.size(), protein ? 1 : 0, 0.5f, 0.05f, query_seq, c, n, &o, nullptr, 0,
                          nullptr);
      memcpy(out, &o, sizeof(o));
    };
    if (fast_bytes) {
      IngestOptions io;
      io.split = split;
      io.block_bytes = fast_bytes;
      io.threads = 4;
      io.min_parallel_bytes = 2048;   // exercise the chunked parallel parser on small fixtures
      run_prediction_fast(
          stdin, mapping, tax, q, r, io,
          [&](const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t, trpa_result* res) {
            for (uint32_t i = 0; i < n_segs; ++i) oracle_segment(segs[i].query_seq, cands + segs[i].cand_begin, segs[i].cand_count, &res[i]);
          },
          std::cout, nullptr, nullptr);
      return 0;
    }
    RecordSetReader reader(std::cin, mapping, tax, split, sorted);
    std::ofstream nolog("/dev/null");
    run_prediction_stream(
        reader, tax, batch,
        [&](std::vector<RecordSet>& sets, std::vector<PredictionRecord>& precs, std::ostream& log) {
          std::vector<trpa_segment> segs;
          std::vector<trpa_candidate> cands;
          flatten_record_sets(sets, q, r, precs, segs, cands);
          std::vector<trpa_result> res(sets.size());
          static_assert(sizeof(OrcResult) == sizeof(trpa_result), "result layouts must agree");
          for (size_t i = 0; i < sets.size(); ++i) {
            OrcResult o;
            orc_predict_segment(tax.parent.data(), tax.left.data(), tax.right.data(), tax.depth.data(), (uint32_t)tax.size(),
                                tax.root, qc.data(), q.off.data(), q.len.data(), (uint32_t)q.size(), rc.data(),
                                r.off.data(), r.len.data(), (uint32_t)r.size(), protein ? 1 : 0, 0.5f, 0.05f,
                                segs[i].query_seq, cands.data() + segs[i].cand_begin, segs[i].cand_count, &o, nullptr, 0,
                                nullptr);
            memcpy(&res[i], &o, sizeof(o));
          }
          apply_results(tax, segs, res, precs, log, nullptr);
        },
        std::cout, nolog);
    return 0;
  } catch (std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
}
