#include "records.h"

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <fstream>

namespace taxator_b200 {

static void split_fields(const std::string& s, size_t begin, std::vector<std::string>& out, int maxfields) {
  // tokenizeSingleCharDelim(line, fields, "\t", maxfields, false) (core/src/utils.hh:58-78)
  out.clear();
  size_t last = begin;
  const size_t n = s.size();
  while (maxfields && last < n) {
    size_t pos = s.find('\t', last);
    if (pos == std::string::npos) { out.push_back(s.substr(last)); last = n; break; }
    out.push_back(s.substr(last, pos - last));
    --maxfields;
    last = pos + 1;
  }
  out.push_back(s.substr(std::min(last, n)));
}

static uint32_t to_u32(const std::string& s, const char* what) {
  // boost::lexical_cast<unsigned>: digits only (no sign handled here, no leading blanks)
  if (s.empty() || s[0] < '0' || s[0] > '9') throw ParsingError(std::string("bad record: ") + what);
  errno = 0;
  char* end = nullptr;
  unsigned long long v = strtoull(s.c_str(), &end, 10);
  if (errno || *end != '\0' || v > 0xffffffffull) throw ParsingError(std::string("bad record: ") + what);
  return (uint32_t)v;
}
// boost::lexical_cast<float/double> accepts neither leading blanks nor hexadecimal floats (strtof would)
static bool plain_decimal(const std::string& s) {
  if (s.empty() || isspace((unsigned char)s[0])) return false;
  const size_t k = (s[0] == '-' || s[0] == '+') ? 1 : 0;
  return !(s.size() > k + 1 && s[k] == '0' && (s[k + 1] == 'x' || s[k + 1] == 'X'));
}
static float to_float(const std::string& s, const char* what) {
  if (!plain_decimal(s)) throw ParsingError(std::string("bad record: ") + what);
  errno = 0;
  char* end = nullptr;
  float v = strtof(s.c_str(), &end);
  if (s.empty() || *end != '\0') throw ParsingError(std::string("bad record: ") + what);
  return v;
}
static double to_double(const std::string& s, const char* what) {
  if (!plain_decimal(s)) throw ParsingError(std::string("bad record: ") + what);
  char* end = nullptr;
  double v = strtod(s.c_str(), &end);
  if (s.empty() || *end != '\0') throw ParsingError(std::string("bad record: ") + what);
  return v;
}

SeqIdMapping load_mapping(const std::string& filename) {
  std::ifstream in(filename.c_str());
  if (!in.good()) throw FileNotFound("could not find file: " + filename);
  SeqIdMapping m;
  m.filename = filename;
  std::string line;
  std::vector<std::string> f;
  while (std::getline(in, line)) {
    if (!line.empty() && line[0] == '#') continue;
    if (line.empty()) continue;
    split_fields(line, 0, f, 2);
    if (f.size() < 2) throw ParsingError("Could not parse line: " + line);
    m.map[f[0]] = f[1];
  }
  return m;
}

AlignmentRecord* parse_alignment_line(const std::string& line, const SeqIdMapping& mapping, const FlatTaxonomy& tax) {
  if (line.size() <= 1) throw ParsingError("bad record: alignment line too short");
  std::vector<std::string> f;
  AlignmentRecord* r = new AlignmentRecord();
  try {
    r->masked = line[0] == '*';
    split_fields(line, r->masked ? 1 : 0, f, 12);
    if (f.size() < 11) throw ParsingError("bad record: bad number of fields in alignment line");
    r->qstart = to_u32(f[1], "bad position number or query length");
    r->qstop = to_u32(f[2], "bad position number or query length");
    if (r->qstart > r->qstop) throw ParsingError("bad record: reverse query positions not allowed (only reference positions can be swapped to indicate the reverse complement, adjust input");
    r->qlen = to_u32(f[3], "bad position number or query length");
    r->rstart = to_u32(f[5], "bad position number or query length");
    r->rstop = to_u32(f[6], "bad position number or query length");
    r->score = to_float(f[7], "bad score");
    r->evalue = to_double(f[8], "bad E-value");
    r->identities = to_u32(f[9], "bad identity value");
    r->alnlen = to_u32(f[10], "bad alignment length");
    if (f.size() >= 12) r->alignment_code = f[11];
    r->query_id = f[0];
    r->ref_id = f[4];
    r->node = tax.node_of(mapping[r->ref_id]);
  } catch (...) {
    delete r;
    throw;
  }
  return r;
}

RecordSetReader::RecordSetReader(std::istream& in, const SeqIdMapping& mapping, const FlatTaxonomy& tax,
                                 bool split_alignments, bool alignments_sorted)
    : in_(in), mapping_(mapping), tax_(tax), split_(split_alignments), sorted_(alignments_sorted) {
  pending_ = next_record();
  if (pending_) rstop_ = pending_->qstop;
}

RecordSetReader::~RecordSetReader() {
  delete pending_;
  for (size_t i = tmpindex_; i < ranges_.size(); ++i) delete ranges_[i];
}

AlignmentRecord* RecordSetReader::next_record() {
  std::string line;
  while (std::getline(in_, line)) {
    ++line_num_;
    if (!line.empty() && line[0] == '#') continue;  // ignoreLine
    try {
      return parse_alignment_line(line, mapping_, tax_);
    } catch (TaxatorError& e) {
      throw ParsingError(std::string(e.what()) + " (line " + std::to_string(line_num_) + ")");
    }
  }
  return nullptr;
}

bool RecordSetReader::notEmpty() const { return pending_ != nullptr || ranges_.size() > tmpindex_; }

void RecordSetReader::getNext(RecordSet& rset) {
  if (sorted_ && split_) {  // RecordSetGeneratorSorted<true>: alignmentrecord.hh:561-600
    while (pending_) {
      AlignmentRecord* rec = pending_;
      if (!rset.empty() && rset.front()->query_id != rec->query_id) { rstop_ = rec->qstop; return; }
      if (!rset.empty() && rec->qstart > rstop_) { rstop_ = rec->qstop; return; }
      rstop_ = rset.empty() ? rec->qstop : std::max(rec->qstop, rstop_);
      rset.push_back(rec);
      pending_ = next_record();
    }
    return;
  }
  if (ranges_.size() <= tmpindex_) {  // read all records of the next query
    ranges_.clear();
    tmpindex_ = 0;
    if (!pending_) throw TaxatorError("alignment recordset parser: read from empty input");
    const std::string qid = pending_->query_id;
    ranges_.push_back(pending_);
    pending_ = nullptr;
    for (;;) {
      AlignmentRecord* rec = next_record();
      if (!rec) break;
      if (rec->query_id == qid) ranges_.push_back(rec);
      else { pending_ = rec; break; }
    }
    if (split_) {
      // std::sort on (qstart, qstop, pointer) in the reference (alignmentrecord.hh:480); arrival order
      // stands in for the pointer order (only matters for records tied in score AND identities)
      std::stable_sort(ranges_.begin(), ranges_.end(), [](const AlignmentRecord* a, const AlignmentRecord* b) {
        if (a->qstart != b->qstart) return a->qstart < b->qstart;
        return a->qstop < b->qstop;
      });
    }
  }
  if (!split_) {
    for (AlignmentRecord* r : ranges_) rset.push_back(r);
    ranges_.clear();
    tmpindex_ = 0;
    return;
  }
  uint32_t run_stop = ranges_[tmpindex_]->qstop;
  rset.push_back(ranges_[tmpindex_]);
  for (size_t i = tmpindex_ + 1; i < ranges_.size(); ++i) {
    if (ranges_[i]->qstart > run_stop) { tmpindex_ = i; return; }  // split point
    run_stop = std::max(run_stop, ranges_[i]->qstop);
    rset.push_back(ranges_[i]);
  }
  ranges_.clear();
  tmpindex_ = 0;
}

void delete_records(RecordSet& rset) {
  for (AlignmentRecord* r : rset) delete r;
  rset.clear();
}

}  // namespace taxator_b200
