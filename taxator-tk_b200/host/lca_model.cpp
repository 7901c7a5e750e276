#include "lca_model.h"

namespace taxator_b200 {

LCAPredictionModelGPU::LCAPredictionModelGPU(const FlatTaxonomy* tax, const trpa_lca_params& params, int device)
    : TaxonPredictionModel<RecordSet>(tax), params_(params) {
  ctx_ = trpa_create(device, nullptr);
  if (!ctx_) throw TaxatorError(std::string("GPU context: ") + trpa_last_error());
  if (trpa_load_taxonomy(ctx_, tax->parent.data(), tax->left.data(), tax->right.data(), tax->depth.data(),
                         (uint32_t)tax->size(), tax->root))
    throw TaxatorError(std::string("GPU set-up: ") + trpa_last_error());
}

LCAPredictionModelGPU::~LCAPredictionModelGPU() { if (ctx_) trpa_destroy(ctx_); }

void LCAPredictionModelGPU::predictBatch(std::vector<RecordSet>& recordsets, std::vector<PredictionRecord>& precs,
                                         std::ostream&) {
  const size_t n = recordsets.size();
  if (precs.size() != n) throw TaxatorError("predictBatch: precs and recordsets differ in size");
  if (!n) return;
  std::vector<trpa_segment> segs(n);
  std::vector<trpa_candidate> cands;
  std::vector<double> evalue;
  for (size_t i = 0; i < n; ++i) {
    RecordSet& rs = recordsets[i];
    if (rs.empty()) throw TaxatorError("predictBatch: empty record set");
    // initPredictionRecord (taxonpredictionmodel.hh:41-43): the first record, masked or not
    precs[i].initialize(rs.front()->getQueryIdentifier(), rs.front()->getQueryLength());
    segs[i].query_seq = 0; segs[i].reserved = 0;
    segs[i].cand_begin = (uint32_t)cands.size();
    for (AlignmentRecord* r : rs) {
      if (r->isFiltered()) continue;   // every filter and loop of these models skips masked records
      trpa_candidate c{};
      c.qstart = r->getQueryStart(); c.qstop = r->getQueryStop();
      c.rstart = r->getReferenceStart(); c.rstop = r->getReferenceStop();
      c.score = r->getScore(); c.identities = r->getIdentities(); c.alnlen = r->getAlignmentLength();
      c.node = r->getReferenceNode();
      cands.push_back(c);
      evalue.push_back(r->evalue);
    }
    segs[i].cand_count = (uint32_t)cands.size() - segs[i].cand_begin;
  }
  std::vector<trpa_result> res(n);
  {
    std::lock_guard<std::mutex> lock(mutex_);
    if (trpa_predict_lca_batch(ctx_, &params_, segs.data(), (uint32_t)n, cands.data(), (uint32_t)cands.size(), evalue.data(),
                               tax_->unclassified.empty() ? nullptr : tax_->unclassified.data(), res.data(), 1, nullptr))
      throw TaxatorError(std::string("GPU prediction: ") + trpa_last_error());
  }
  for (size_t i = 0; i < n; ++i) {
    const trpa_result& r = res[i];
    PredictionRecord& p = precs[i];
    if (r.kind == TRPA_KIND_NONE) {   // setUnclassified: the feature range stays as initialised
      p.setNodePoint(tax_->root, 0);
      p.setBestReferenceTaxon(tax_->root);
    } else {
      p.setQueryFeatureBegin(r.qrstart);
      p.setQueryFeatureEnd(r.qrstop);
      p.setNodePoint(r.lower_node, r.support);
      p.setBestReferenceTaxon(r.rtax_node);
    }
  }
}

void LCAPredictionModelGPU::predictFlat(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                                        uint32_t n_cands, const double* evalue, trpa_result* res) {
  if (!n_segs) return;
  std::lock_guard<std::mutex> lock(mutex_);
  if (trpa_predict_lca_batch(ctx_, &params_, segs, n_segs, cands, n_cands, evalue,
                             tax_->unclassified.empty() ? nullptr : tax_->unclassified.data(), res, 1, nullptr))
    throw TaxatorError(std::string("GPU prediction: ") + trpa_last_error());
}

void LCAPredictionModelGPU::predict(RecordSet& recordset, PredictionRecord& prec, std::ostream& logsink) {
  std::vector<RecordSet> one(1);
  one[0].swap(recordset);
  std::vector<PredictionRecord> precs(1, prec);
  try {
    predictBatch(one, precs, logsink);
  } catch (...) {
    one[0].swap(recordset);
    throw;
  }
  one[0].swap(recordset);
  prec = precs[0];
}

}  // namespace taxator_b200
