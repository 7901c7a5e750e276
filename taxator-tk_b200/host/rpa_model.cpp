#include "rpa_model.h"

#include <cmath>
#include <thread>

namespace taxator_b200 {

const char* const kGFF3Header = "##gff-version 3\n";

// predictionrecord.hh:248-308.  setNodeRange(l, u, s) stores the same support at every rank between
// lower and upper, so printFeatureTax reduces to "LOWER:SUP-UPPER" / "NODE:SUP" (and drops ":0").
void PredictionRecord::print(std::ostream& os) const {
  os << query_identifier_ << "\ttaxator-tk\tsequence_feature\t" << query_feature_begin_ << '\t' << query_feature_end_ << '\t';
  if (std::isnan(signal_strength_)) os << '.';
  else os << signal_strength_;
  os << "\t.\t.\tseqlen=" << query_length_ << ";tax=";
  uint32_t last_support = 0;
  uint32_t node = lower_node_;
  while (node != upper_node_) {
    if (support_ != last_support) { os << tax_->taxid[node] << ':' << support_ << '-'; last_support = support_; }
    node = tax_->parent[node];
  }
  os << tax_->taxid[node];
  if (support_ != last_support) os << ':' << support_;
  os << ";rtax=" << tax_->taxid[rtax_];
  if (interpolation_value_ >= 0. && interpolation_value_ < 1.) os << ";ival=" << interpolation_value_;
  os << '\n';
}

RPAPredictionModelGPU::RPAPredictionModelGPU(const FlatTaxonomy* tax, const SeqStore& q_storage,
                                             const SeqStore& db_storage, float exclude_factor, float reeval_bandwidth,
                                             bool protein, const std::vector<int>& devices)
    : TaxonPredictionModel<RecordSet>(tax), q_store_(q_storage), db_store_(db_storage), protein_(protein) {
  const int alpha = protein ? TRPA_ALPHA_AA : TRPA_ALPHA_NT;
  for (int dev : devices) {
    trpa_ctx* c = trpa_create(dev, nullptr);
    if (!c) throw TaxatorError(std::string("GPU context: ") + trpa_last_error());
    ctx_.push_back(c);
    auto load = [&](int which, const SeqStore& st) -> int {
      if (st.packed) {   // refpack file: already in the HBM layout
        if (st.alphabet != alpha) { throw TaxatorError("refpack file holds the other alphabet (-b nucleotide / protein)"); }
        return trpa_load_store_packed(c, which, alpha, st.woff.data(), st.len.data(), (uint32_t)st.size(), st.payload.data(),
                                      st.woff.empty() ? 0 : st.woff.back());
      }
      return trpa_load_store(c, which, alpha, st.chars.data(), st.off.data(), st.len.data(), (uint32_t)st.size());
    };
    if (trpa_set_params(c, exclude_factor, reeval_bandwidth) ||
        trpa_load_taxonomy(c, tax->parent.data(), tax->left.data(), tax->right.data(), tax->depth.data(),
                           (uint32_t)tax->size(), tax->root) ||
        load(TRPA_STORE_QUERY, q_storage) || load(TRPA_STORE_REF, db_storage))
      throw TaxatorError(std::string("GPU set-up: ") + trpa_last_error());
  }
  if (ctx_.empty()) throw TaxatorError("no GPU given");
}

void RPAPredictionModelGPU::exportStore(int which, std::vector<uint64_t>& woff, std::vector<uint32_t>& len,
                                        std::vector<char>& payload, int& alphabet) {
  uint32_t n_seq = 0;
  uint64_t n_words = 0;
  if (trpa_store_info(ctx_[0], which, &alphabet, &n_seq, &n_words)) throw TaxatorError(std::string("store export: ") + trpa_last_error());
  woff.assign((size_t)n_seq + 1, 0);
  len.assign(n_seq, 0);
  payload.assign((size_t)(n_words * (alphabet == TRPA_ALPHA_NT ? 12 : 4)), 0);
  std::vector<uint32_t> len1(std::max<uint32_t>(n_seq, 1));
  std::vector<char> pay1(std::max<size_t>(payload.size(), 1));
  if (trpa_export_store(ctx_[0], which, woff.data(), len1.data(), pay1.data())) throw TaxatorError(std::string("store export: ") + trpa_last_error());
  std::copy(len1.begin(), len1.begin() + n_seq, len.begin());
  std::copy(pay1.begin(), pay1.begin() + payload.size(), payload.begin());
}

RPAPredictionModelGPU::~RPAPredictionModelGPU() {
  for (trpa_ctx* c : ctx_) trpa_destroy(c);
}

void RPAPredictionModelGPU::run_shard(size_t dev_slot, const trpa_segment* segs, uint32_t n_segs,
                                      const trpa_candidate* cands, uint32_t cand_begin, uint32_t n_cands,
                                      trpa_result* out, std::string* error) {
  // rebase the shard's candidate ranges to its own table
  std::vector<trpa_segment> local(segs, segs + n_segs);
  for (auto& s : local) s.cand_begin -= cand_begin;
  if (trpa_predict_batch(ctx_[dev_slot], local.data(), n_segs, cands + cand_begin, n_cands, out)) *error = trpa_last_error();
}

void flatten_record_sets(std::vector<RecordSet>& recordsets, const SeqStore& q_store, const SeqStore& db_store,
                         std::vector<PredictionRecord>& precs, std::vector<trpa_segment>& segs,
                         std::vector<trpa_candidate>& cands) {
  const size_t n = recordsets.size();
  if (precs.size() != n) throw TaxatorError("predictBatch: precs and recordsets differ in size");
  segs.assign(n, trpa_segment());
  cands.clear();
  for (size_t i = 0; i < n; ++i) {
    RecordSet& rs = recordsets[i];
    if (rs.empty()) throw TaxatorError("predictBatch: empty record set");
    // initPredictionRecord (taxonpredictionmodel.hh:41-43): the first record, masked or not
    precs[i].initialize(rs.front()->getQueryIdentifier(), rs.front()->getQueryLength());
    segs[i].cand_begin = (uint32_t)cands.size();
    segs[i].reserved = 0;
    uint32_t cnt = 0;
    for (AlignmentRecord* r : rs) {
      if (r->isFiltered()) continue;  // active_records, hh:350-356
      trpa_candidate c;
      c.ref_seq = db_store.ordinal(r->getReferenceIdentifier());
      c.rstart = r->getReferenceStart(); c.rstop = r->getReferenceStop();
      c.qstart = r->getQueryStart(); c.qstop = r->getQueryStop();
      c.score = r->getScore(); c.identities = r->getIdentities(); c.alnlen = r->getAlignmentLength();
      c.node = r->getReferenceNode();
      cands.push_back(c);
      ++cnt;
    }
    segs[i].cand_count = cnt;
    // the reference looks the query up only when it realigns (n >= 2, hh:415)
    segs[i].query_seq = cnt >= 2 ? q_store.ordinal(rs.front()->getQueryIdentifier()) : 0;
  }
}

void apply_results(const FlatTaxonomy& tax, const std::vector<trpa_segment>& segs, const std::vector<trpa_result>& res,
                   std::vector<PredictionRecord>& precs, std::ostream& logsink, PredictStats* stats) {
  for (size_t i = 0; i < res.size(); ++i) {
    const trpa_result& r = res[i];
    PredictionRecord& p = precs[i];
    if (r.kind == TRPA_KIND_NONE) {          // setUnclassified; ival and feature range stay as initialised
      p.setNodePoint(tax.root, 0);
      p.setBestReferenceTaxon(tax.root);
    } else {
      p.setQueryFeatureBegin(r.qrstart);
      p.setQueryFeatureEnd(r.qrstop);
      p.setInterpolationValue(r.ival);
      p.setNodeRange(r.lower_node, r.upper_node, r.support);
      p.setBestReferenceTaxon(r.rtax_node);
      if (r.kind == TRPA_KIND_PLACED) p.setSignalStrength(r.signal);
    }
    if (stats) {
      stats->segments++;
      stats->alignments += (uint64_t)r.n_pass0 + r.n_pass1 + r.n_pass2;
      stats->cells += r.cells;
    }
    // the per-segment STATS line of the reference's log (hh:834-837), without its timers
    logsink << "STATS\t" << r.qrstart << ':' << r.qrstop << '@' << p.getQueryIdentifier() << '\t' << segs[i].cand_count
            << '\t' << r.n_pass0 << '\t' << r.n_pass1 << '\t' << r.n_pass2 << '\t'
            << (r.n_pass0 + r.n_pass1 + r.n_pass2) << '\n';
  }
}

void RPAPredictionModelGPU::predictFlat(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                                        uint32_t n_cands, trpa_result* res) {
  const size_t n = n_segs;
  if (!n) return;
  // shard contiguous segment ranges over the GPUs, balanced by the estimated DP work (trpa_shard_bounds:
  // 1 + sum of span^2 per segment -- the same cuts python/shard.py and bench.py make); no collective
  const size_t G = std::min(ctx_.size(), std::max<size_t>(1, n));
  std::vector<uint32_t> cut(G + 1, n_segs);
  if (trpa_shard_bounds(segs, n_segs, cands, n_cands, (uint32_t)G, cut.data()))
    throw TaxatorError(std::string("sharding failed: ") + trpa_last_error());
  std::vector<std::string> errors(G);
  if (G == 1) {
    if (trpa_predict_batch(ctx_[0], segs, n_segs, cands, n_cands, res)) errors[0] = trpa_last_error();
  } else {
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
      const size_t b = cut[g], e = cut[g + 1];
      if (e <= b) continue;
      const uint32_t cb = segs[b].cand_begin;
      const uint32_t ce = e < n ? segs[e].cand_begin : n_cands;
      th.emplace_back(&RPAPredictionModelGPU::run_shard, this, g, segs + b, (uint32_t)(e - b), cands, cb, ce - cb, res + b,
                      &errors[g]);
    }
    for (auto& t : th) t.join();
  }
  for (const auto& e : errors) if (!e.empty()) throw TaxatorError("GPU prediction failed: " + e);
}

void RPAPredictionModelGPU::predictFlatTraced(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                                              uint32_t n_cands, trpa_result* res, std::vector<trpa_trace_entry>& trace) {
  trace.clear();
  if (!n_segs) return;
  trpa_ctx* c = ctx_[0];
  uint64_t n = 0;
  int rc = trpa_set_trace(c, 1);
  if (!rc) rc = trpa_predict_batch(c, segs, n_segs, cands, n_cands, res);
  if (!rc) rc = trpa_batch_trace(c, nullptr, 0, &n);
  if (!rc && n) {
    trace.resize(n);
    rc = trpa_batch_trace(c, trace.data(), n, &n);
  }
  const std::string msg = rc ? trpa_last_error() : "";
  trpa_set_trace(c, 0);
  if (rc) throw TaxatorError("GPU prediction failed: " + msg);
}

void RPAPredictionModelGPU::predictBatch(std::vector<RecordSet>& recordsets, std::vector<PredictionRecord>& precs,
                                         std::ostream& logsink) {
  const size_t n = recordsets.size();
  std::vector<trpa_segment> segs;
  std::vector<trpa_candidate> cands;
  flatten_record_sets(recordsets, q_store_, db_store_, precs, segs, cands);
  std::vector<trpa_result> res(n);
  predictFlat(segs.data(), (uint32_t)n, cands.data(), (uint32_t)cands.size(), res.data());
  apply_results(*tax_, segs, res, precs, logsink, &stats_);
}

void RPAPredictionModelGPU::predict(RecordSet& recordset, PredictionRecord& prec, std::ostream& logsink) {
  std::lock_guard<std::mutex> lock(single_mutex_);
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
