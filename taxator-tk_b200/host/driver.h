// The driver loop of the CLI (== doPredictionsSerial, core/taxator.cpp:50-77, but in batches):
// gathers record sets, calls a batch predictor, prints GFF3 in input order and replays the
// reference's reuse of one PredictionRecord (taxator.cpp:66).  The predictor is injected so that the
// host logic can be exercised without a GPU in the CPU test-suite.
#pragma once
#include <functional>
#include <ostream>
#include <sstream>
#include <vector>

#include "records.h"
#include "rpa_model.h"

namespace taxator_b200 {

typedef std::function<void(std::vector<RecordSet>&, std::vector<PredictionRecord>&, std::ostream&)> BatchPredictor;

inline uint64_t run_prediction_stream(RecordSetReader& reader, const FlatTaxonomy& tax, size_t batch_segments,
                                      const BatchPredictor& predict, std::ostream& out, std::ostream& logsink) {
  out << kGFF3Header;
  PredictionRecord carry(&tax);
  std::vector<RecordSet> sets;
  std::vector<PredictionRecord> precs;
  uint64_t total = 0;
  while (reader.notEmpty()) {
    sets.clear();
    while (reader.notEmpty() && sets.size() < batch_segments) {
      sets.emplace_back();
      reader.getNext(sets.back());
      if (sets.back().empty()) sets.pop_back();
    }
    precs.assign(sets.size(), PredictionRecord(&tax));
    predict(sets, precs, logsink);
    std::ostringstream os;
    for (size_t i = 0; i < sets.size(); ++i) {
      bool any_active = false;
      for (AlignmentRecord* r : sets[i]) any_active |= !r->isFiltered();
      if (!any_active) {  // n==0: ival and signal are whatever the reused record held (hh:359-368)
        precs[i].setInterpolationValue(carry.getInterpolationValue());
        precs[i].setSignalStrength(carry.getSignalStrength());
      }
      carry = precs[i];
      os << precs[i];
      delete_records(sets[i]);
    }
    out << os.str();
    total += sets.size();
  }
  out.flush();
  return total;
}

}  // namespace taxator_b200
