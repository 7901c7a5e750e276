// RPAPredictionModelGPU: the host-side drop-in for the reference's RPAPredictionModel
// (core/src/taxonpredictionmodelsequence.hh:326-881) behind the same plugin interface
// TaxonPredictionModel<ContainerT>::predict(recordset, prec, logsink)
// (core/src/taxonpredictionmodel.hh:36-52).  All arithmetic happens on the GPU through the C ABI
// (include/taxator_rpa_b200.h); this class only flattens record sets into candidate tables and turns
// result records back into PredictionRecords.  predict() handles one segment per call like the
// reference; predictBatch() is what a driver loop should use to keep a B200 busy.
#pragma once
#include <memory>
#include <mutex>
#include <ostream>
#include <string>
#include <vector>

#include "records.h"
#include "seqstore.h"
#include "taxonomy.h"
#include "taxator_rpa_b200.h"

namespace taxator_b200 {

// == PredictionRecord (core/src/predictionrecord.hh:38-435), the fields predict() sets + GFF3 print
class PredictionRecord {
 public:
  explicit PredictionRecord(const FlatTaxonomy* tax) : tax_(tax) {}
  void initialize(const std::string& qid, uint32_t qlen) {
    query_identifier_ = qid; query_length_ = qlen; query_feature_begin_ = 1; query_feature_end_ = qlen;
  }
  const std::string& getQueryIdentifier() const { return query_identifier_; }
  uint32_t getQueryLength() const { return query_length_; }
  uint32_t getQueryFeatureBegin() const { return query_feature_begin_; }
  uint32_t getQueryFeatureEnd() const { return query_feature_end_; }
  uint32_t getLowerNode() const { return lower_node_; }
  uint32_t getUpperNode() const { return upper_node_; }
  uint32_t getBestReferenceTaxon() const { return rtax_; }
  float getInterpolationValue() const { return interpolation_value_; }
  float getSignalStrength() const { return signal_strength_; }
  uint32_t getSupport() const { return support_; }
  void setQueryFeatureBegin(uint32_t v) { query_feature_begin_ = v; }
  void setQueryFeatureEnd(uint32_t v) { query_feature_end_ = v; }
  void setInterpolationValue(float v) { interpolation_value_ = v; }
  void setSignalStrength(float v) { signal_strength_ = v; }
  void setBestReferenceTaxon(uint32_t n) { rtax_ = n; }
  void setNodeRange(uint32_t lower, uint32_t upper, uint32_t support) { lower_node_ = lower; upper_node_ = upper; support_ = support; }
  void setNodePoint(uint32_t node, uint32_t support) { setNodeRange(node, node, support); }
  void print(std::ostream& os) const;  // GFF3 line, predictionrecord.hh:248-308

 private:
  const FlatTaxonomy* tax_;
  std::string query_identifier_;
  uint32_t query_length_ = 0, query_feature_begin_ = 0, query_feature_end_ = 0;
  uint32_t lower_node_ = 0, upper_node_ = 0, rtax_ = 0, support_ = 0;
  float interpolation_value_ = -1.f, signal_strength_ = 0.f;
};
inline std::ostream& operator<<(std::ostream& os, const PredictionRecord& p) { p.print(os); return os; }
extern const char* const kGFF3Header;  // "##gff-version 3\n" (core/src/predictionrecord.cpp:22-25)

// plugin interface of the reference (taxonpredictionmodel.hh:36-52)
template <typename ContainerT>
class TaxonPredictionModel {
 public:
  explicit TaxonPredictionModel(const FlatTaxonomy* tax) : tax_(tax) {}
  virtual ~TaxonPredictionModel() {}
  virtual void predict(ContainerT& recordset, PredictionRecord& prec, std::ostream& logsink) = 0;
 protected:
  const FlatTaxonomy* tax_;
};

struct PredictStats { uint64_t segments = 0, alignments = 0, cells = 0; };

// record sets -> flat segment / candidate tables of the C ABI (unmasked records only, hh:350-356);
// also runs initPredictionRecord on precs[i]
void flatten_record_sets(std::vector<RecordSet>& recordsets, const SeqStore& q_store, const SeqStore& db_store,
                         std::vector<PredictionRecord>& precs, std::vector<trpa_segment>& segs,
                         std::vector<trpa_candidate>& cands);
// result records -> PredictionRecords (+ the STATS log line of hh:834-837)
void apply_results(const FlatTaxonomy& tax, const std::vector<trpa_segment>& segs, const std::vector<trpa_result>& res,
                   std::vector<PredictionRecord>& precs, std::ostream& logsink, PredictStats* stats);

class RPAPredictionModelGPU : public TaxonPredictionModel<RecordSet> {
 public:
  // same arguments as RPAPredictionModel(tax, q_storage, db_storage, exclude_factor, reeval_bandwidth)
  // (hh:329) plus the alphabet and the GPUs to use (one context per device)
  RPAPredictionModelGPU(const FlatTaxonomy* tax, const SeqStore& q_storage, const SeqStore& db_storage,
                        float exclude_factor, float reeval_bandwidth, bool protein, const std::vector<int>& devices);
  ~RPAPredictionModelGPU() override;

  // one segment per call, re-entrant (serialised on the first device)
  void predict(RecordSet& recordset, PredictionRecord& prec, std::ostream& logsink) override;
  // many segments per call; precs[i] must arrive carrying the state a reused record would have
  // (only matters for n==0 sets, whose ival the reference leaves untouched)
  void predictBatch(std::vector<RecordSet>& recordsets, std::vector<PredictionRecord>& precs, std::ostream& logsink);

  // the flat tables of the C ABI in, result records out: segments sharded over the model's GPUs
  // (what the fast ingest path of the CLI calls; ingest.h)
  void predictFlat(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands,
                   trpa_result* res);

  // same on the first GPU only, recording every consumed alignment (trpa_set_trace / trpa_batch_trace): what the
  // verbose log (-l) is written from; trace entries come back ordered by segment
  void predictFlatTraced(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands,
                         trpa_result* res, std::vector<trpa_trace_entry>& trace);

  // a store as the first GPU packed it (what write_refpack() puts into a .trpk file); which = 0 query, 1 reference
  void exportStore(int which, std::vector<uint64_t>& woff, std::vector<uint32_t>& len, std::vector<char>& payload, int& alphabet);

  typedef PredictStats Stats;
  Stats stats() const { return stats_; }
  Stats& mutable_stats() { return stats_; }

 private:
  void run_shard(size_t dev_slot, const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                 uint32_t cand_begin, uint32_t n_cands, trpa_result* out, std::string* error);
  const SeqStore& q_store_;
  const SeqStore& db_store_;
  bool protein_;
  std::vector<trpa_ctx*> ctx_;
  std::mutex single_mutex_;
  Stats stats_;
};

}  // namespace taxator_b200
