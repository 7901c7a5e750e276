// Alignment records and record sets (= query segments): host-side input of the drop-in boundary.
// Mirrors AlignmentRecordTaxonomy (core/src/alignmentrecord.hh:39-238), FileParser
// (core/src/fileparser.hh:28-76) and the RecordSetGenerator family (alignmentrecord.hh:413-631).
#pragma once
#include <cstdint>
#include <istream>
#include <list>
#include <string>
#include <unordered_map>
#include <vector>

#include "taxonomy.h"

namespace taxator_b200 {

struct AlignmentRecord {
  std::string query_id, ref_id, alignment_code;
  uint32_t qstart = 0, qstop = 0, qlen = 0, rstart = 0, rstop = 0;
  float score = 0;
  double evalue = 0;
  uint32_t identities = 0, alnlen = 0;
  bool masked = false;      // line started with '*' (isFiltered)
  uint32_t node = 0;        // taxon node of the reference (getReferenceNode)

  const std::string& getQueryIdentifier() const { return query_id; }
  uint32_t getQueryStart() const { return qstart; }
  uint32_t getQueryStop() const { return qstop; }
  uint32_t getQueryLength() const { return qlen; }
  const std::string& getReferenceIdentifier() const { return ref_id; }
  uint32_t getReferenceStart() const { return rstart; }
  uint32_t getReferenceStop() const { return rstop; }
  float getScore() const { return score; }
  uint32_t getIdentities() const { return identities; }
  uint32_t getAlignmentLength() const { return alnlen; }
  bool isFiltered() const { return masked; }
  uint32_t getReferenceNode() const { return node; }
};

typedef std::list<AlignmentRecord*> RecordSet;  // taxator.cpp:48

// seqid -> taxid mapping (core/src/accessconv.hh:50-99)
struct SeqIdMapping {
  std::unordered_map<std::string, std::string> map;
  std::string filename;
  const std::string& operator[](const std::string& acc) const {
    auto it = map.find(acc);
    if (it == map.end()) throw TaxonMappingNotFound("bad taxon identifier mapping: " + acc + " (" + filename + ")");
    return it->second;
  }
};
SeqIdMapping load_mapping(const std::string& filename);

// one 12-column TAB line -> record (throws ParsingError / TaxonMappingNotFound / TaxonNotFound)
AlignmentRecord* parse_alignment_line(const std::string& line, const SeqIdMapping& mapping, const FlatTaxonomy& tax);

// Reads records from a stream and hands out record sets like RecordSetGenerator{Unsorted,Sorted}.
class RecordSetReader {
 public:
  RecordSetReader(std::istream& in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, bool split_alignments,
                  bool alignments_sorted);
  ~RecordSetReader();
  bool notEmpty() const;
  void getNext(RecordSet& rset);  // caller owns the records (delete them after predict)

 private:
  AlignmentRecord* next_record();
  std::istream& in_;
  const SeqIdMapping& mapping_;
  const FlatTaxonomy& tax_;
  bool split_, sorted_;
  unsigned line_num_ = 0;
  AlignmentRecord* pending_ = nullptr;       // first record of the next query
  std::vector<AlignmentRecord*> ranges_;     // current query, ordered by (qstart, qstop, arrival)
  size_t tmpindex_ = 0;
  uint32_t rstop_ = 0;
};

void delete_records(RecordSet& rset);

}  // namespace taxator_b200
