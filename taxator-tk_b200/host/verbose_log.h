// Verbose per-alignment log of `taxator -l` for segments placed on the GPU; see verbose_log.cpp.
#pragma once
#include <ostream>
#include <string>

#include "seqstore.h"
#include "taxonomy.h"
#include "taxator_rpa_b200.h"

namespace taxator_b200 {

struct VerboseLogContext {
  const FlatTaxonomy* tax = nullptr;
  const SeqStore* q_store = nullptr;    // only read to render protein alignments (nullable)
  const SeqStore* db_store = nullptr;
  bool protein = false;
  float exclude_factor = 0.5f;               // -x
  float reeval_bandwidth_factor = 0.95f;     // float(1. - toppercent), hh:334
};

// One segment: its record set as handed to the GPU (unmasked records, record-set order), the GPU's result record
// and the slice of the alignment trace that belongs to it (trpa_batch_trace).  Throws TaxatorError if the replay of
// the reference's control flow does not arrive at `res`.
void write_segment_log(const VerboseLogContext& ctx, const std::string& qid, uint32_t query_seq, const trpa_candidate* cands, uint32_t n,
                       const trpa_result& res, const trpa_trace_entry* trace, size_t n_trace, std::ostream& logsink);

}  // namespace taxator_b200
