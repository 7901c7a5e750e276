// TEST HARNESS (CPU): the product's host code for the alignment-free models (taxonomy loader incl. the
// "unclassified" flags, alignment parser, block-parallel ingest in its store-less mode with the e-value column,
// GFF3 formatting) with the GPU predictor swapped for the oracle (oracle/rpa_oracle.cpp orc_predict_lca_model).
// Run on the files the real reference reads, its GFF3 must be identical -- checks the host logic without a GPU.
// Usage: host_lca_harness <model 0..3> <toppercent> <minscore> <maxevalue> <minsupport> <nbest> <ignore_unclassified>
//        mapping [block_bytes]        (alignments on stdin, taxonomy from $TAXATORTK_TAXONOMY_NCBI)
#include <cstring>
#include <iostream>

#include "../taxator-tk_b200/host/ingest.h"

extern "C" {
struct OrcResult {
  uint32_t qrstart, qrstop, lower, upper, rtax, support;
  float ival, signal;
  uint32_t p0, p1, p2, kind;
  uint64_t cells;
};
int orc_predict_lca_model(const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth,
                          uint32_t n_nodes, uint32_t root, uint32_t model, float toppercent, float minscore,
                          double maxevalue_arg, uint32_t minsupport, uint32_t nbest, int ignore_unclassified,
                          const void* cands, const double* evalue, uint32_t n, const uint8_t* unclassified, OrcResult* res);
}

using namespace taxator_b200;

int main(int argc, char** argv) {
  if (argc < 9) { std::cerr << "usage\n"; return 2; }
  try {
    const uint32_t model = (uint32_t)std::stoul(argv[1]);
    const float toppercent = std::stof(argv[2]), minscore = std::stof(argv[3]);
    const double maxevalue = std::stod(argv[4]);
    const uint32_t minsupport = (uint32_t)std::stoul(argv[5]), nbest = (uint32_t)std::stoul(argv[6]);
    const int ignore_unclassified = std::stoi(argv[7]);
    FlatTaxonomy tax = load_taxonomy_from_environment(kDefaultRanks, true);
    SeqIdMapping mapping = load_mapping(argv[8]);
    IngestOptions io;
    io.block_bytes = argc > 9 ? std::stoul(argv[9]) : (128u << 20);
    io.threads = 4;
    io.min_parallel_bytes = 1;
    io.need_stores = false;
    io.want_evalue = true;
    const SeqStore no_store;
    std::ios::sync_with_stdio(false);
    run_prediction_fast_blocks(
        stdin, mapping, tax, no_store, no_store, io,
        [&](FlatBlock& b) {
          for (size_t s = 0; s < b.segs.size(); ++s) {
            OrcResult o;
            const trpa_segment& sg = b.segs[s];
            orc_predict_lca_model(tax.parent.data(), tax.left.data(), tax.right.data(), tax.depth.data(), (uint32_t)tax.size(),
                                  tax.root, model, toppercent, minscore, maxevalue, minsupport, nbest, ignore_unclassified,
                                  b.cands.data() + sg.cand_begin, b.evalue.data() + sg.cand_begin, sg.cand_count,
                                  tax.unclassified.data(), &o);
            static_assert(sizeof(OrcResult) == sizeof(trpa_result), "result layouts differ");
            memcpy(&b.res[s], &o, sizeof(o));
          }
        },
        std::cout, nullptr, nullptr);
    return 0;
  } catch (std::exception& e) {
    std::cerr << "An unrecoverable error occurred: " << e.what() << std::endl;
    return 1;
  }
}
