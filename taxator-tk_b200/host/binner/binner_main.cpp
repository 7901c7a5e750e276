// binner-b200: command-line drop-in for the reference's `binner` (core/binner.cpp), the consumer of taxator's GFF3:
// one taxon per (globbed) sequence identifier from all its segment predictions.  Same options, same input (GFF3 on
// stdin or -f files, $TAXATORTK_TAXONOMY_NCBI) and the same Bioboxes binning file on stdout.  The host parses and
// groups; the reductions run on the GPU through trpa_bin_batch (include/taxator_rpa_b200.h, csrc/binner.cu).
//
// Grouping order: the reference keeps the groups in a std::unordered_map<std::string, ...> and writes them in its
// iteration order (binner.cpp:186, 296).  The same container with the same insertion sequence is used here, so with
// the same standard library the lines come out in the same order.
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <regex>
#include <set>
#include <sstream>
#include <unordered_map>

#include "../taxonomy.h"
#include "taxator_rpa_b200.h"

using namespace taxator_b200;

namespace {

const char* kVersion = "1.5.0";   // program_version (core/src/constants.hh:46): written into the header

// tokenizeSingleCharDelim (core/src/utils.hh:58-78)
void tokenize(const std::string& str, std::vector<std::string>& tokens, char delim, int fieldnum, bool trimempty) {
  const size_t n = str.size();
  if (!fieldnum) fieldnum = (int)n;
  size_t pos, last = 0;
  while (fieldnum && last < n) {
    pos = str.find(delim, last);
    if (pos == std::string::npos) {
      pos = n;
      if (pos != last || !trimempty) tokens.push_back(str.substr(last, pos - last));
      last = pos;
      break;
    }
    if (pos != last || !trimempty) { tokens.push_back(str.substr(last, pos - last)); --fieldnum; }
    last = pos + 1;
  }
  tokens.push_back(str.substr(std::min(last, n)));
}

// boost::lexical_cast<large_unsigned_int>: digits only here (no sign, no blanks)
bool to_u32(const std::string& s, uint32_t& v) {
  if (s.empty()) return false;
  uint64_t x = 0;
  for (char ch : s) {
    if (ch < '0' || ch > '9') return false;
    x = x * 10 + (uint64_t)(ch - '0');
    if (x > 0xffffffffull) return false;
  }
  v = (uint32_t)x;
  return true;
}
bool to_float(const std::string& s, float& v) {
  if (s.empty() || isspace((unsigned char)s[0])) return false;
  char* end = nullptr;
  v = strtof(s.c_str(), &end);
  return *end == '\0';
}

struct Record {   // the fields of PredictionRecordBase the binner reads (core/src/predictionrecord.hh:262-272)
  std::string qid;
  uint32_t qlen = 0, begin = 0, end = 0;
  uint32_t lower = 0, upper = 0;
  std::vector<uint32_t> support;   // taxon_support_: upper ... lower
};

// PredictionRecordBase::parse (core/src/predictionrecord.hh:199-245) + parseKeyValue (:310-378)
void parse_record(const std::string& line, const FlatTaxonomy& tax, Record& r) {
  if (line.empty()) throw ParsingError("empty GFF3 line");
  std::vector<std::string> f;
  tokenize(line, f, '\t', 9, false);
  if (f.size() < 9) throw ParsingError("too few GFF3 fields in line");
  if (f[1].size() < 10 || f[1].compare(0, 10, "taxator-tk") != 0) std::cerr << "warning: gff3 produced by unknown algorithm" << std::endl;
  if (!to_u32(f[3], r.begin) || !to_u32(f[4], r.end)) throw ParsingError("bad GFF3 feature position");
  if (r.begin > r.end) throw ParsingError("GFF3 reverse query positions");
  float signal;
  if (f[5] != "." && !to_float(f[5], signal)) throw ParsingError("bad GFF3 taxonomic signal score");
  std::vector<std::string> kvs, kv;
  tokenize(f[8], kvs, ';', 0, true);
  for (const std::string& item : kvs) {
    kv.clear();
    tokenize(item, kv, '=', 2, false);
    const std::string& key = kv[0];
    const std::string value = kv.size() > 1 ? kv[1] : std::string();
    if (key == "seqlen") {
      if (!to_u32(value, r.qlen)) throw ParsingError("bad GFF3 key value: seqlen");
    } else if (key == "ival") {
      float v;
      if (!to_float(value, v)) throw ParsingError("bad GFF3 key value: ival");
    } else if (key == "tax") {
      std::vector<std::string> path, ts;
      tokenize(value, path, '-', 0, false);
      size_t it = 0;
      tokenize(path[it], ts, ':', 2, false);
      uint32_t support;
      if (ts.size() < 2 || ts[1].empty()) support = r.end - r.begin + 1;
      else if (!to_u32(ts[1], support)) throw ParsingError("bad GFF3 key value: tax");
      uint32_t last = tax.node_of(ts[0]);
      r.lower = last;
      std::vector<uint32_t> rev;   // supports from the lower node upwards
      while (++it < path.size() && !path[it].empty()) {
        ts.clear();
        tokenize(path[it], ts, ':', 2, false);
        const uint32_t node = tax.node_of(ts[0]);
        if (!tax.is_parent_of(node, last)) throw ParsingError("bad taxon path: " + tax.taxid[node] + " is not parent of " + tax.taxid[last]);
        for (uint32_t x = last; x != node; x = tax.parent[x]) rev.push_back(support);
        if (ts.size() > 1 && !ts[1].empty() && !to_u32(ts[1], support)) throw ParsingError("bad GFF3 key value: tax");
        last = node;
      }
      rev.push_back(support);
      r.upper = last;
      r.support.assign(rev.rbegin(), rev.rend());
    } else if (key == "rtax") {
      tax.node_of(value);   // TaxonNotFound like the reference
    }
  }
  r.qid = f[0];
}

struct Options {
  std::string sample_identifier, glob_regex = "(.+)", logfile = "binning.log", min_support_in_sample = "0";
  uint32_t min_support_per_sequence = 50;
  float signal_majority = .7f;
  std::vector<std::string> identity_constrain, files, ranks;
  bool delete_unmarked = true;
  int device = 0;
};

void usage(std::ostream& os) {
  os << "Allowed options:\n"
        "  -h [ --help ]                          show help message\n"
        "  -V [ --version ]                       show program version\n"
        "  -n [ --sample-identifier ] arg         unique sample identifier\n"
        "  -g [ --glob-identifier ] arg (=(.+))   grouping regex for substring matching to glob sequence identifiers\n"
        "  -s [ --sequence-min-support ] arg (=50)  minimum number of positions supporting a taxonomic signal for any single sequence\n"
        "  -j [ --signal-majority ] arg (=0.7)    minimum combined fraction of support for any single sequence\n"
        "  -i [ --identity-constrain ] arg        minimum required identity for this rank (e.g. -i species:0.8 -i genus:0.7)\n"
        "  -f [ --files ] arg                     prediction files (\"-\" = standard input)\n"
        "  -l [ --logfile ] arg (=binning.log)    log file\n"
        "  -r [ --ranks ] arg                     ranks at which to do predictions\n"
        "  -m [ --sample-min-support ] arg (=0)   minimum support in positions (>=1) or fraction of total support (<1) for any taxon\n"
        "  -d [ --delete-notranks ] arg (=1)      delete all nodes that don't have any of the given ranks\n"
        "  --gpu arg (=0)                         CUDA device\n";
}

int parse_args(int argc, char** argv, Options& o) {
  struct Spec { const char* lng; char shrt; bool multi; };
  static const Spec specs[] = {{"help", 'h', false}, {"version", 'V', false}, {"sample-identifier", 'n', false}, {"glob-identifier", 'g', false},
    {"sequence-min-support", 's', false}, {"signal-majority", 'j', false}, {"identity-constrain", 'i', false}, {"files", 'f', true},
    {"logfile", 'l', false}, {"ranks", 'r', true}, {"sample-min-support", 'm', false}, {"delete-notranks", 'd', false},
    {"preallocate-num-queries", 'P', false}, {"gpu", 'G', false}, {"citation", 'C', false}, {"advanced-options", 'A', false}};
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    const Spec* sp = nullptr;
    std::string val; bool has_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string name = a.substr(2);
      const size_t eq = name.find('=');
      if (eq != std::string::npos) { val = name.substr(eq + 1); has_val = true; name.resize(eq); }
      for (const auto& s : specs) if (name == s.lng) sp = &s;
    } else if (a.size() >= 2 && a[0] == '-') {
      for (const auto& s : specs) if (a[1] == s.shrt && s.shrt != 'P' && s.shrt != 'G' && s.shrt != 'C' && s.shrt != 'A') sp = &s;
      if (a.size() > 2) { val = a.substr(2); has_val = true; }
    }
    if (!sp) throw TaxatorError("unrecognised option '" + a + "'");
    auto need = [&]() -> std::string {
      if (has_val) return val;
      if (i + 1 >= argc) throw TaxatorError("option '" + a + "' needs a value");
      return argv[++i];
    };
    auto many = [&](std::vector<std::string>& out) {
      if (has_val) out.push_back(val);
      while (i + 1 < argc && (argv[i + 1][0] != '-' || !strcmp(argv[i + 1], "-"))) out.push_back(argv[++i]);
    };
    switch (sp->shrt) {
      case 'h': usage(std::cout); return 1;
      case 'V': std::cout << kVersion << std::endl; return 1;
      case 'C': case 'A': return 1;
      case 'n': o.sample_identifier = need(); break;
      case 'g': o.glob_regex = need(); break;
      case 's': o.min_support_per_sequence = (uint32_t)std::stoul(need()); break;
      case 'j': o.signal_majority = std::stof(need()); break;
      case 'i': o.identity_constrain.push_back(need()); break;
      case 'f': many(o.files); break;
      case 'l': o.logfile = need(); break;
      case 'r': many(o.ranks); break;
      case 'm': o.min_support_in_sample = need(); break;
      case 'd': { const std::string v = need(); o.delete_unmarked = !(v == "0" || v == "false" || v == "off" || v == "no"); break; }
      case 'G': o.device = std::stoi(need()); break;
      default: need(); break;
    }
  }
  if (o.sample_identifier.empty()) throw TaxatorError("the option '--sample-identifier' is required but missing");
  return 0;
}

// extractRegex (binner.cpp:45-55)
std::string extract(const std::string& text, const std::regex& re, bool empty_regex) {
  if (empty_regex) return "consensus_sequence";
  std::cmatch m;
  if (!std::regex_match(text.c_str(), m, re) || m.size() < 2 || m[1].first == m[1].second)
    throw TaxatorError("sequence identifier '" + text + "' does not match the glob regex with a non-empty group");
  return std::string(m[1].first, m[1].second);
}

}  // namespace

int main(int argc, char** argv) {
  Options opt;
  try {
    if (parse_args(argc, argv, opt)) return EXIT_SUCCESS;
    if (opt.ranks.empty()) opt.ranks = kDefaultRanks;
    const std::regex glob(opt.glob_regex);
    const bool empty_regex = opt.glob_regex.empty();
    trpa_bin_params pp;
    memset(&pp, 0, sizeof(pp));
    pp.signal_majority = opt.signal_majority;
    pp.min_support_per_sequence = opt.min_support_per_sequence;
    if (opt.min_support_in_sample.find('.') == std::string::npos) {   // binner.cpp:125-126
      if (!to_u32(opt.min_support_in_sample, pp.min_support_in_sample)) throw TaxatorError("bad --sample-min-support");
    } else if (!to_float(opt.min_support_in_sample, pp.min_support_in_sample_fraction)) throw TaxatorError("bad --sample-min-support");

    FlatTaxonomy tax = load_taxonomy_from_environment(opt.ranks, opt.delete_unmarked);
    std::string tax_version;   // optional $TAXATORTK_TAXONOMY_NCBI/version.txt (core/src/ncbidata.cpp:183-207)
    if (const char* root = getenv("TAXATORTK_TAXONOMY_NCBI")) {
      std::ifstream vf((std::string(root) + "/version.txt").c_str());
      if (vf.good()) std::getline(vf, tax_version);
    }

    // identity constraints: -i rank:pid (binner.cpp:137-156); rank ids = distinct rank names of the taxonomy
    std::vector<uint8_t> rank_of_node;
    std::vector<float> pid_per_rank;
    if (!opt.identity_constrain.empty()) {
      std::map<std::string, uint8_t> rank_id;
      rank_of_node.resize(tax.size());
      for (size_t i = 0; i < tax.size(); ++i) {
        auto it = rank_id.find(tax.rank[i]);
        if (it == rank_id.end()) {
          if (rank_id.size() >= 255) throw TaxatorError("too many distinct ranks");
          it = rank_id.insert(std::make_pair(tax.rank[i], (uint8_t)rank_id.size())).first;
        }
        rank_of_node[i] = it->second;
      }
      pid_per_rank.assign(rank_id.size() + 1, -1.f);
      for (const std::string& v : opt.identity_constrain) {
        std::vector<std::string> f;
        tokenize(v, f, ':', 1, false);
        if (f[0].empty()) {
          std::cerr << "Could not read identity constrain: rank cannot be empty string, use e.g. \"-i species:0.8\"" << std::endl;
          return EXIT_FAILURE;
        }
        float pid;
        if (f.size() < 2 || !to_float(f[1], pid)) {
          std::cerr << "Could not read identity constrain: \"" << (f.size() > 1 ? f[1] : "") << "\" for rank \"" << f[0]
                    << "\" as float, use e.g. \"-i species:0.8\"" << std::endl;
          return EXIT_FAILURE;
        }
        auto it = rank_id.find(f[0]);
        if (it != rank_id.end()) pid_per_rank[it->second] = pid;   // a rank no node carries can never be met on a path
      }
      pp.n_ranks = (uint32_t)pid_per_rank.size();
    }

    // STEP 0: parse all inputs, group by glob identifier (binner.cpp:162-210)
    typedef std::vector<Record> RecordGroup;
    std::unordered_map<std::string, RecordGroup> grouped;
    {
      std::string new_name, old_name;
      RecordGroup* records = nullptr;
      auto consume = [&](std::istream& in) {
        std::string line;
        while (std::getline(in, line)) {
          if (line.empty() || line[0] == '#') continue;
          Record r;
          parse_record(line, tax, r);
          new_name = extract(r.qid, glob, empty_regex);
          if (new_name != old_name) {
            records = &grouped.emplace(new_name, RecordGroup()).first->second;
            old_name.swap(new_name);
          }
          records->push_back(std::move(r));
        }
      };
      if (opt.files.empty()) consume(std::cin);
      else
        for (const std::string& fn : opt.files) {
          if (fn == "-") { consume(std::cin); continue; }
          std::ifstream in(fn.c_str());
          if (!in.good()) { std::cerr << "Could not read file \"" << fn << "\"" << std::endl; return EXIT_FAILURE; }
          consume(in);
        }
    }

    // flat tables in the map's iteration order (= output order)
    std::vector<trpa_bin_record> recs;
    std::vector<uint32_t> supports, group_begin(1, 0);
    std::vector<const std::string*> names;
    std::unordered_map<std::string, uint32_t> qids;
    for (auto& kv : grouped) {
      names.push_back(&kv.first);
      for (const Record& r : kv.second) {
        trpa_bin_record b;
        b.lower_node = r.lower; b.upper_node = r.upper; b.support_begin = (uint32_t)supports.size();
        b.query_length = r.qlen; b.reserved = 0;
        b.query_id = qids.emplace(r.qid, (uint32_t)qids.size()).first->second;
        supports.insert(supports.end(), r.support.begin(), r.support.end());
        recs.push_back(b);
      }
      group_begin.push_back((uint32_t)recs.size());
    }

    trpa_ctx* ctx = trpa_create(opt.device, nullptr);
    if (!ctx) throw TaxatorError(std::string("GPU context: ") + trpa_last_error());
    std::vector<trpa_bin_result> res(names.size());
    trpa_bin_stats stats;
    memset(&stats, 0, sizeof(stats));
    if (trpa_load_taxonomy(ctx, tax.parent.data(), tax.left.data(), tax.right.data(), tax.depth.data(), (uint32_t)tax.size(), tax.root) ||
        trpa_bin_batch(ctx, &pp, recs.data(), (uint32_t)recs.size(), supports.data(), (uint32_t)supports.size(), group_begin.data(),
                       (uint32_t)names.size(), pp.n_ranks ? rank_of_node.data() : nullptr, pp.n_ranks ? pid_per_rank.data() : nullptr,
                       res.data(), &stats)) {
      const std::string msg = trpa_last_error();
      trpa_destroy(ctx);
      throw TaxatorError("GPU binning failed: " + msg);
    }
    trpa_destroy(ctx);
    std::cerr << "Analyzing sample composition: " << stats.nested_taxa << " nested taxa with total support of " << stats.root_support
              << " positions" << std::endl;
    std::cerr << "Noise removal: " << stats.pruned_taxa << " taxa removed" << std::endl;
    std::cerr << "Consensus taxonomy assignment: ";
    std::ofstream log(opt.logfile.c_str());   // the reference's per-group debug trace is not reproduced

    // Bioboxes binning format (core/src/bioboxes.cpp:20-63)
    std::ostringstream os;
    os << "# This is the bioboxes.org binning output format at\n# https://github.com/bioboxes/rfc/tree/master/data-format\n\n";
    os << "@Version:0.9.1\n@SampleID:" << opt.sample_identifier << '\n';
    if (!tax_version.empty()) os << "@TaxonomyID:" << tax_version << '\n';
    os << "@_TaxatorTK_Version:" << kVersion << "\n\n";
    os << "@@SequenceID\tTaxID\t_TaxatorTK_Support\t_TaxatorTK_Length\n";
    for (size_t g = 0; g < names.size(); ++g) {
      if (res[g].mode == TRPA_BIN_EMPTY) continue;
      os << *names[g] << '\t' << tax.taxid[res[g].node] << '\t' << res[g].support << '\t' << res[g].length << '\n';
      log << *names[g] << '\t' << (res[g].mode == TRPA_BIN_SINGLE ? "single" : res[g].mode == TRPA_BIN_DIRECT ? "direct" : "fallback")
          << '\t' << tax.taxid[res[g].lower_node] << ':' << res[g].lower_support << '-' << tax.taxid[res[g].upper_node] << ':'
          << res[g].upper_support << '\n';
    }
    std::cout << os.str() << std::flush;
    std::cerr << " done" << std::endl;
    return EXIT_SUCCESS;
  } catch (std::exception& e) {
    std::cerr << "An unrecoverable error occurred." << std::endl << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}
