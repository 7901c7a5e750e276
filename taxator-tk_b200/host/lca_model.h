// The alignment-free placement models of the reference behind the same plugin interface
// TaxonPredictionModel<ContainerT>::predict(recordset, prec, logsink) (core/src/taxonpredictionmodel.hh:36-52):
// DummyPredictionModel (:57-67), LCASimplePredictionModel (:71-125), MeganLCAPredictionModel (:129-156),
// NBestLCAPredictionModel (:235-252), constructed like taxator.cpp:346-361 does.  The placement itself runs
// on the GPU through trpa_predict_lca_batch (include/taxator_rpa_b200.h); this class flattens record sets
// and turns result records back into PredictionRecords.
#pragma once
#include <limits>
#include <mutex>
#include <vector>

#include "rpa_model.h"

namespace taxator_b200 {

class LCAPredictionModelGPU : public TaxonPredictionModel<RecordSet> {
 public:
  // DummyPredictionModel(tax) / LCASimplePredictionModel(tax)
  static trpa_lca_params dummy() { trpa_lca_params p{}; p.model = TRPA_MODEL_DUMMY; return p; }
  static trpa_lca_params simple() { trpa_lca_params p{}; p.model = TRPA_MODEL_SIMPLE_LCA; return p; }
  // MeganLCAPredictionModel(tax, ignore_unclassified, toppercent, minscore, minsupport, maxevalue)
  static trpa_lca_params megan(bool ignore_unclassified = false, float toppercent = 1.0f, float minscore = 0.0f,
                               int minsupport = 1, double maxevalue = std::numeric_limits<double>::max()) {
    trpa_lca_params p{};
    p.model = TRPA_MODEL_MEGAN_LCA; p.ignore_unclassified = ignore_unclassified; p.toppercent = toppercent;
    p.minscore = minscore; p.minsupport = (uint32_t)minsupport;
    p.maxevalue = (float)maxevalue;   // the filter's constructor takes a float (alignmentsfilter.hh:351)
    return p;
  }
  // NBestLCAPredictionModel(tax, n)
  static trpa_lca_params nbest(int n = 1) { trpa_lca_params p{}; p.model = TRPA_MODEL_NBEST_LCA; p.nbest = (uint32_t)n; return p; }

  LCAPredictionModelGPU(const FlatTaxonomy* tax, const trpa_lca_params& params, int device = 0);
  ~LCAPredictionModelGPU() override;
  void predict(RecordSet& recordset, PredictionRecord& prec, std::ostream& logsink) override;
  void predictBatch(std::vector<RecordSet>& recordsets, std::vector<PredictionRecord>& precs, std::ostream& logsink);
  // the flat tables of the C ABI in, result records out (what the fast ingest path of the CLI calls; ingest.h)
  void predictFlat(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands,
                   const double* evalue, trpa_result* res);

 private:
  trpa_lca_params params_;
  trpa_ctx* ctx_ = nullptr;
  std::mutex mutex_;
};

}  // namespace taxator_b200
