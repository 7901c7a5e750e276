// Reference-side binding of the B200 RPA path: a TaxonPredictionModel that a taxator-tk maintainer drops next to
// RPAPredictionModel (core/src/taxonpredictionmodelsequence.hh:326-881).  It compiles against the UNMODIFIED
// reference headers (core/src/taxonpredictionmodel.hh:36-52, sequencestorage.hh:41-52, taxontree.hh:46-73,
// predictionrecord.hh) and talks to the GPU only through the C ABI of include/taxator_rpa_b200.h.
//
// Same template parameters and constructor as RPAPredictionModel (hh:326-339, call site core/taxator.cpp:252), so
// it is selected by changing one identifier there; integration/taxator_b200_main.cpp does exactly that for the
// parity test without touching the reference tree.
//
// The reference's store interface (RandomSeqStoreROInterface) can only be asked for ranges of named sequences, it
// cannot be enumerated.  The binding therefore mirrors the stores lazily: the first time a record names a
// sequence, the whole sequence is fetched through the interface and the HBM-resident store of the context is
// extended.  The reference store is complete after every genome has been seen once; the query store holds the
// last few queries (bounded, see kMaxQueries / kMaxQueryChars).  One segment per call is latency bound on a GPU -- this class
// is the minimal drop-in; a driver that hands over whole batches (RPAPredictionModelGPU::predictBatch in
// taxator-tk_b200/host/rpa_model.h) is what keeps a B200 busy.
#ifndef taxonpredictionmodelgpu_hh_
#define taxonpredictionmodelgpu_hh_

#include <limits>
#include <map>
#include <string>
#include <vector>
#include <boost/thread/mutex.hpp>
#include "src/taxonpredictionmodel.hh"
#include "src/sequencestorage.hh"
#include "src/exception.hh"
#include <taxator_rpa_b200.h>

template< typename ContainerT, typename QStorType, typename DBStorType, typename StringType >
class RPAPredictionModelB200 : public TaxonPredictionModel< ContainerT > {
  public:
    RPAPredictionModelB200(const Taxonomy* tax, QStorType& q_storage, const DBStorType& db_storage, float exclude_factor, float reeval_bandwidth = .1, int device = 0)
      : TaxonPredictionModel< ContainerT >(tax), query_sequences_(q_storage), db_sequences_(db_storage),
        alphabet_(IsProtein< StringType >::value ? TRPA_ALPHA_AA : TRPA_ALPHA_NT) {
      ctx_ = trpa_create(device, NULL);
      if(!ctx_) fail("trpa_create");
      // flat taxonomy: dense pre-order indices over the (already pruned) tree; taxontree.hh:62-68
      std::vector<uint32_t> parent, left, right;
      std::vector<uint8_t> depth;
      std::vector<const TaxonNode*> stack(1, this->root_);
      while(!stack.empty()) {
        const TaxonNode* node = stack.back();
        stack.pop_back();
        node_index_[node] = static_cast<uint32_t>(nodes_.size());
        nodes_.push_back(node);
        for(const TaxonNode* child = node->last_child; child; child = child->prev_sibling) stack.push_back(child);
      }
      for(std::size_t i = 0; i < nodes_.size(); ++i) {
        const TaxonNode* node = nodes_[i];
        parent.push_back(node == this->root_ ? node_index_[node] : node_index_[node->parent]);
        left.push_back(node->data->leftvalue);
        right.push_back(node->data->rightvalue);
        depth.push_back(node->data->root_pathlength);
      }
      if(trpa_set_params(ctx_, exclude_factor, reeval_bandwidth) ||                                  // taxator.cpp:252 (-x, -t)
         trpa_load_taxonomy(ctx_, parent.data(), left.data(), right.data(), depth.data(), static_cast<uint32_t>(nodes_.size()), 0))
        fail("taxonomy");
    }

    ~RPAPredictionModelB200() { trpa_destroy(ctx_); }

    // the reference's interface: one record set per call, called concurrently by the -p N consumers (taxator.cpp:168)
    void predict(ContainerT& recordset, PredictionRecord& prec, std::ostream& logsink) {
      this->initPredictionRecord(recordset, prec);                                                  // taxonpredictionmodel.hh:41-43
      const std::string& qid = prec.getQueryIdentifier();
      boost::mutex::scoped_lock lock(mutex_);                                                         // a context is single-threaded
      std::vector<trpa_candidate> cands;
      for(typename ContainerT::iterator rec_it = recordset.begin(); rec_it != recordset.end(); ++rec_it) {
        if((*rec_it)->isFiltered()) continue;                                                          // hh:350-356
        trpa_candidate c;
        c.ref_seq = ordinal(db_, db_sequences_, (*rec_it)->getReferenceIdentifier());
        c.rstart = (*rec_it)->getReferenceStart(); c.rstop = (*rec_it)->getReferenceStop();
        c.qstart = (*rec_it)->getQueryStart(); c.qstop = (*rec_it)->getQueryStop();
        c.score = (*rec_it)->getScore(); c.identities = (*rec_it)->getIdentities();
        c.alnlen = (*rec_it)->getAlignmentLength();
        c.node = node_index_.at((*rec_it)->getReferenceNode());
        cands.push_back(c);
      }
      trpa_segment seg = {0, 0, static_cast<uint32_t>(cands.size()), 0};
      if(cands.size() >= 2) {                                                                          // the query is only fetched for n >= 2 (hh:415)
        if(!q_.id2ord.count(qid) && (q_.len.size() >= kMaxQueries || q_.chars.size() > kMaxQueryChars)) q_ = Mirror();   // bounded mirror of the query store
        seg.query_seq = ordinal(q_, query_sequences_, qid);
      }
      upload(q_, TRPA_STORE_QUERY);
      upload(db_, TRPA_STORE_REF);
      trpa_result r;
      if(trpa_predict_batch(ctx_, &seg, 1, cands.data(), static_cast<uint32_t>(cands.size()), &r)) fail("trpa_predict_batch");
      lock.unlock();

      if(r.kind == TRPA_KIND_NONE) { this->setUnclassified(prec); return; }                          // hh:359-368
      prec.setQueryFeatureBegin(r.qrstart);                                                          // hh:829-833
      prec.setQueryFeatureEnd(r.qrstop);
      prec.setInterpolationValue(r.ival);
      prec.setNodeRange(nodes_[r.lower_node], nodes_[r.upper_node], r.support);
      prec.setBestReferenceTaxon(nodes_[r.rtax_node]);
      if(r.kind == TRPA_KIND_PLACED) prec.setSignalStrength(r.signal);                               // hh:828
      logsink << "STATS\t" << r.qrstart << ':' << r.qrstop << '@' << qid << '\t' << cands.size() << '\t' << r.n_pass0 << '\t'
              << r.n_pass1 << '\t' << r.n_pass2 << '\t' << (r.n_pass0 + r.n_pass1 + r.n_pass2) << std::endl;   // hh:834-837 without the timers
    }

  private:
    template< typename T > struct IsProtein { static const bool value = false; };
    static const std::size_t kMaxQueryChars = 256u << 20;
    static const std::size_t kMaxQueries = 64;   // record sets arrive grouped by query; -p N consumers work on neighbouring queries

    // host mirror of a store of the reference (sequence characters as the store's alphabet prints them)
    struct Mirror {
      std::vector<char> chars;
      std::vector<uint64_t> off;
      std::vector<uint32_t> len;
      std::map<std::string, uint32_t> id2ord;
      bool dirty;
      Mirror() : dirty(true) {}
    };

    template< typename StorType >
    uint32_t ordinal(Mirror& m, const StorType& store, const std::string& id) {
      std::map<std::string, uint32_t>::const_iterator it = m.id2ord.find(id);
      if(it != m.id2ord.end()) return it->second;
      // whole sequence through the reference's own accessor: both stores clip `stop` to the sequence length
      // (sequencestorage.hh:115, :353); SequenceNotFound propagates like in the reference
      const StringType seq = store.getSequence(id, 1, std::numeric_limits<int>::max());
      const uint32_t ord = static_cast<uint32_t>(m.len.size());
      m.off.push_back(m.chars.size());
      m.len.push_back(static_cast<uint32_t>(seqan::length(seq)));
      for(std::size_t i = 0; i < seqan::length(seq); ++i) m.chars.push_back(static_cast<char>(seq[i]));
      m.id2ord[id] = ord;
      m.dirty = true;
      return ord;
    }

    void upload(Mirror& m, int which) {
      if(!m.dirty) return;
      const char none = 0;
      if(trpa_load_store(ctx_, which, alphabet_, m.chars.empty() ? &none : m.chars.data(), m.off.data(), m.len.data(), static_cast<uint32_t>(m.len.size())))
        fail("trpa_load_store");
      m.dirty = false;
    }

    void fail(const char* what) {
      BOOST_THROW_EXCEPTION(GeneralError() << general_info(std::string(what) + ": " + trpa_last_error()));
    }

    QStorType& query_sequences_;
    const DBStorType& db_sequences_;
    const int alphabet_;
    trpa_ctx* ctx_;
    std::vector<const TaxonNode*> nodes_;
    std::map<const TaxonNode*, uint32_t> node_index_;
    Mirror q_, db_;
    boost::mutex mutex_;
};

template< typename ContainerT, typename QStorType, typename DBStorType, typename StringType >
template< typename TSpec >
struct RPAPredictionModelB200< ContainerT, QStorType, DBStorType, StringType >::IsProtein< seqan::String< seqan::AminoAcid, TSpec > > { static const bool value = true; };

#endif // taxonpredictionmodelgpu_hh_
