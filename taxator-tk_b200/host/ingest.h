// Fast ingest for the batched GPU path: the caller side of the drop-in boundary at GPU speed.
//
// The reference parses alignments on ONE producer thread, one heap-allocated record per line
// (FileParser / AlignmentRecordFactory / RecordSetGenerator, core/src/fileparser.hh:28-76,
// core/src/alignmentrecord.hh:95-158, :413-504, core/taxator.cpp:80-124): fine next to a CPU that
// places ~300 segments/s, a 10x bottleneck next to a B200 that places 230 000.  Here the alignment
// stream is read in large text blocks cut at query boundaries; the lines of a block are parsed and
// segmented by all host cores in parallel straight into the flat trpa_segment / trpa_candidate
// tables of the C ABI (no per-record objects), the GPU batch runs while the next block is parsed,
// and GFF3 lines are formatted in parallel while the GPU works on the following block.
//
// Semantics are those of the record-at-a-time path (records.cpp), which stays the reference
// implementation on the host: any line that is not a plain well-formed record is handed to
// parse_alignment_line() so that every accepted oddity and every error message is identical; the
// segmentation is RecordSetGeneratorUnsorted<split> (alignmentrecord.hh:457-504).
#pragma once
#include <cstdio>
#include <functional>
#include <memory>
#include <ostream>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "records.h"
#include "rpa_model.h"
#include "seqstore.h"
#include "taxonomy.h"

namespace taxator_b200 {

// reference id -> (ordinal in the reference store, taxon node); resolved once at start-up
class RefResolver {
 public:
  static constexpr uint32_t kNone = 0xffffffffu;
  struct Entry { uint32_t ordinal, node; };
  RefResolver(const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& db_store);
  // nullptr: not in the mapping (the record-at-a-time parser then throws TaxonMappingNotFound)
  const Entry* find(std::string_view id) const {
    auto it = map_.find(id);
    return it == map_.end() ? nullptr : &it->second;
  }

 private:
  std::unordered_map<std::string_view, Entry> map_;  // keys live in the SeqIdMapping
};

struct SegMeta {
  const char* qid;       // into the block's text
  uint32_t qid_len;
  uint32_t qlen;         // query length column of the set's first record (initPredictionRecord)
  uint32_t any_active;   // 0: every record of the set is masked (n == 0)
};

// one block of the alignment stream, flattened
struct FlatBlock {
  std::vector<char> text;            // owns what SegMeta::qid points into
  // Plain pageable vectors.  Page-locked blocks (trpa_host_alloc) were measured and dropped: cudaHostAlloc / FreeHost of
  // the ~60 MB tables cost 0.3-0.7 s per run and stall the GPU stage while they hold the driver lock, against ~16 ms of
  // staged copy saved per 100 k-segment block (scripts/cli_timing_probe.py: 100 k segments 2.4-2.6 s vs 1.73 s wall).
  std::vector<trpa_segment> segs;
  std::vector<trpa_candidate> cands;
  std::vector<double> evalue;        // one per candidate (IngestOptions::want_evalue)
  std::vector<SegMeta> meta;
  std::vector<trpa_result> res;
  uint64_t first_line = 0;           // line number (1-based) of the block's first line
};

struct IngestOptions {
  bool split = true;                 // -s
  size_t block_bytes = 128u << 20;   // text per block (a block always ends at a query boundary)
  unsigned threads = 0;              // 0: hardware concurrency (max 32)
  size_t min_parallel_bytes = 1u << 20;  // smaller blocks are parsed by one thread
  // the alignment-free models read no sequence stores (ordinals stay 0) but need the e-value column
  bool need_stores = true;
  bool want_evalue = false;
};

class FastIngest {
 public:
  FastIngest(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const RefResolver& refs,
             const SeqStore& q_store, const IngestOptions& opt);
  // false: end of input, nothing produced
  bool next(FlatBlock& out);

 private:
  void parse_block(FlatBlock& out, size_t len);
  FILE* in_;
  const SeqIdMapping& mapping_;
  const FlatTaxonomy& tax_;
  const RefResolver& refs_;
  const SeqStore& q_store_;
  IngestOptions opt_;
  std::vector<char> carry_;          // trailing query group of the previous block
  bool eof_ = false;
  uint64_t lines_done_ = 0;
};

// results of a flat batch (any device, any shard order) -> results array
typedef std::function<void(const trpa_segment*, uint32_t, const trpa_candidate*, uint32_t, trpa_result*)> FlatPredictor;
// same, with the whole block (e-values): fills b.res
typedef std::function<void(FlatBlock&)> BlockPredictor;

// GFF3 lines of a block, formatted in parallel; carry = (ival, signal) of the previous record
// (n == 0 sets inherit them: hh:359-368, taxator.cpp:66); updated to the block's last record
void format_block(const FlatBlock& b, const FlatTaxonomy& tax, float& carry_ival, float& carry_signal, unsigned threads,
                  std::string& out, std::ostream* statslog);

// The whole CLI loop: ingest -> predictor -> GFF3, the three stages overlapped.  Returns segments.
struct StageTimes { double ingest_s = 0, predict_s = 0, output_s = 0; uint64_t blocks = 0; };   // busy time per stage
uint64_t run_prediction_fast(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& q_store,
                             const SeqStore& db_store, const IngestOptions& opt, const FlatPredictor& predict,
                             std::ostream& out, std::ostream* statslog, PredictStats* stats, StageTimes* times = nullptr);
uint64_t run_prediction_fast_blocks(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& q_store,
                                    const SeqStore& db_store, const IngestOptions& opt, const BlockPredictor& predict,
                                    std::ostream& out, std::ostream* statslog, PredictStats* stats, StageTimes* times = nullptr);

}  // namespace taxator_b200
