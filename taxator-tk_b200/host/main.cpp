// taxator-b200: command-line drop-in for `taxator -a rpa` (core/taxator.cpp) with the RPA hot path on
// B200 GPUs.  Same inputs (alignments on stdin, FASTA (+.fai) stores, seqid->taxid mapping,
// $TAXATORTK_TAXONOMY_NCBI) and the same GFF3 on stdout.  Record sets are gathered into batches and
// handed to RPAPredictionModelGPU::predictBatch; output keeps the input order.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "driver.h"
#include "ingest.h"
#include "records.h"
#include "lca_model.h"
#include "rpa_model.h"
#include "seqstore.h"
#include "taxonomy.h"
#include "verbose_log.h"

using namespace taxator_b200;

static const char* kVersion = "1.5.0-b200";

struct Options {
  std::string algorithm = "rpa", mapping, query, query_index, ref, ref_index, logfile = "/dev/null", dataformat = "nucleotide";
  std::vector<std::string> ranks;
  unsigned processors = 1;
  bool split_alignments = true, alignments_sorted = false, delete_unmarked = true;
  float filterout = 0.5f, toppercent = 0.05f;
  // the alignment-free models' options (taxator.cpp:287-292)
  double maxevalue = 1000.0;
  unsigned minsupport = 1, nbest = 1;
  float minscore = 0.0f;
  bool ignore_unclassified = false;
  std::vector<int> gpus = {0};
  size_t batch_segments = 200000;
  size_t batch_bytes = 128u << 20;
  bool legacy_ingest = false;
  std::string write_refpack, write_querypack;
  bool timing = false;
};

static void usage(std::ostream& os) {
  os << "Allowed options:\n"
        "  -h [ --help ]                     show help message\n"
        "  -V [ --version ]                  show program version\n"
        "  -a [ --algorithm ] arg (=rpa)     rpa, simple-lca, megan-lca, ic-megan-lca, n-best-lca, dummy\n"
        "  -g [ --seqid-taxid-mapping ] arg  filename of seqid->taxid mapping for reference\n"
        "  -q [ --query-sequences ] arg      query sequences FASTA\n"
        "  -v [ --query-sequences-index ] arg  query sequences FASTA index\n"
        "  -f [ --ref-sequences ] arg        reference sequences FASTA\n"
        "  -i [ --ref-sequences-index ] arg  reference FASTA index (.fai)\n"
        "  -p [ --processors ] arg (=1)      accepted for compatibility (the work runs on the GPUs)\n"
        "  -l [ --logfile ] arg (=/dev/null) verbose per-alignment log (ID/PASS/+ALN/EXT/SCORE/RANGE/STATS lines)\n"
        "  -b [ --dataformat ] arg (=nucleotide)  nucleotide or protein\n"
        "  -r [ --ranks ] arg...             node ranks at which to do predictions\n"
        "  -s [ --split-alignments ] arg (=1)\n"
        "  -o [ --alignments-sorted ] arg (=0)\n"
        "  -d [ --delete-notranks ] arg (=1)\n"
        "  -x [ --heuristic-cutoff ] arg (=0.5)\n"
        "  -t [ --toppercent ] arg (=0.05)   RPA re-evaluation band or top percent parameter for LCA methods\n"
        "  -e [ --max-evalue ] arg (=1000)   maximum evalue (megan-lca)\n"
        "  -c [ --min-support ] arg (=1)     minimum support (megan-lca)\n"
        "  -m [ --minscore ] arg (=0)        minimum score (megan-lca)\n"
        "  -n [ --nbest ] arg (=1)           n-best-lca parameter\n"
        "  -u [ --ignore-unclassified ]      ignore alignments to (partly) unclassified taxa (megan-lca)\n"
        "  --gpus arg (=0)                   comma separated CUDA device indices to shard segments over\n"
        "  --batch-segments arg (=200000)    record sets per GPU batch (record-at-a-time ingest)\n"
        "  --batch-bytes arg (=134217728)    alignment text per GPU batch (fast ingest)\n"
        "  --write-refpack arg               write the reference store (-f/-i) as a packed .trpk file and exit;\n"
        "                                    a .trpk file is accepted wherever a FASTA is (-f, -q)\n"
        "  --write-querypack arg             same for the query store (-q)\n"
        "  --legacy-ingest                   parse records one at a time like the reference (always used with -o 1)\n"
        "  --timing                          print load/predict timing to stderr\n";
}

static bool parse_bool(const std::string& s) {
  if (s == "1" || s == "true" || s == "on" || s == "yes") return true;
  if (s == "0" || s == "false" || s == "off" || s == "no") return false;
  throw TaxatorError("bad boolean option value: " + s);
}

static int parse_args(int argc, char** argv, Options& o) {
  struct Spec { const char* lng; char shrt; };
  static const Spec specs[] = {{"help", 'h'}, {"version", 'V'}, {"algorithm", 'a'}, {"seqid-taxid-mapping", 'g'},
    {"query-sequences", 'q'}, {"query-sequences-index", 'v'}, {"ref-sequences", 'f'}, {"ref-sequences-index", 'i'},
    {"processors", 'p'}, {"logfile", 'l'}, {"dataformat", 'b'}, {"ranks", 'r'}, {"split-alignments", 's'},
    {"alignments-sorted", 'o'}, {"delete-notranks", 'd'}, {"heuristic-cutoff", 'x'}, {"toppercent", 't'},
    {"gpus", 'G'}, {"batch-segments", 'B'}, {"timing", 'T'}, {"batch-bytes", 'Y'}, {"legacy-ingest", 'L'}, {"write-refpack", 'W'}, {"write-querypack", 'Q'},
    // the alignment-free models' knobs; -w is accepted and ignored
    {"max-evalue", 'e'}, {"min-support", 'c'}, {"minscore", 'm'}, {"nbest", 'n'}, {"db-whitelist", 'w'},
    {"ignore-unclassified", 'u'}, {"citation", 'C'}, {"advanced-options", 'A'}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    char key = 0;
    std::string val; bool has_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string name = a.substr(2);
      size_t eq = name.find('=');
      if (eq != std::string::npos) { val = name.substr(eq + 1); has_val = true; name.resize(eq); }
      for (const auto& s : specs) if (name == s.lng) key = s.shrt;
    } else if (a.size() >= 2 && a[0] == '-') {
      for (const auto& s : specs) if (a[1] == s.shrt && s.shrt != 'G' && s.shrt != 'B' && s.shrt != 'T' && s.shrt != 'C' && s.shrt != 'A' && s.shrt != 'Y' && s.shrt != 'L' && s.shrt != 'W' && s.shrt != 'Q') key = s.shrt;
      if (a.size() > 2) { val = a.substr(2); has_val = true; }
    }
    if (!key) throw TaxatorError("unrecognised option '" + a + "'");
    auto need = [&]() -> std::string {
      if (has_val) return val;
      if (i + 1 >= argc) throw TaxatorError("option '" + a + "' needs a value");
      return argv[++i];
    };
    switch (key) {
      case 'h': usage(std::cout); return 1;
      case 'V': std::cout << kVersion << std::endl; return 1;
      case 'C': case 'A': return 1;
      case 'a': o.algorithm = need(); break;
      case 'g': o.mapping = need(); break;
      case 'q': o.query = need(); break;
      case 'v': o.query_index = need(); break;
      case 'f': o.ref = need(); break;
      case 'i': o.ref_index = need(); break;
      case 'p': o.processors = (unsigned)std::stoul(need()); break;
      case 'l': o.logfile = need(); break;
      case 'b': o.dataformat = need(); break;
      case 'r':
        if (has_val) o.ranks.push_back(val);
        while (i + 1 < argc && argv[i + 1][0] != '-') o.ranks.push_back(argv[++i]);
        break;
      case 's': o.split_alignments = parse_bool(need()); break;
      case 'o': o.alignments_sorted = parse_bool(need()); break;
      case 'd': o.delete_unmarked = parse_bool(need()); break;
      case 'x': o.filterout = std::stof(need()); break;
      case 't': o.toppercent = std::stof(need()); break;
      case 'G': {
        o.gpus.clear();
        std::stringstream ss(need());
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) o.gpus.push_back(std::stoi(tok));
        break;
      }
      case 'B': o.batch_segments = std::stoul(need()); break;
      case 'Y': o.batch_bytes = std::stoul(need()); break;
      case 'L': o.legacy_ingest = true; break;
      case 'W': o.write_refpack = need(); break;
      case 'Q': o.write_querypack = need(); break;
      case 'T': o.timing = true; break;
      case 'u': o.ignore_unclassified = true; break;
      case 'e': o.maxevalue = std::stod(need()); break;
      case 'c': o.minsupport = (unsigned)std::stoul(need()); break;
      case 'm': o.minscore = std::stof(need()); break;
      case 'n': o.nbest = (unsigned)std::stoul(need()); break;
      default: need(); break;  // ignored options with a value
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  Options opt;
  try {
    if (parse_args(argc, argv, opt)) return EXIT_SUCCESS;
    if (opt.ranks.empty()) opt.ranks = kDefaultRanks;
    if (opt.mapping.empty()) {
      std::cout << "Specify a taxonomy mapping file for the reference sequence identifiers" << std::endl;
      usage(std::cout);
      return EXIT_FAILURE;
    }
    if (opt.algorithm != "rpa") {
      // the alignment-free models (taxator.cpp:346-361): no sequence stores, records one set at a time
      trpa_lca_params params;
      if (opt.algorithm == "dummy") params = LCAPredictionModelGPU::dummy();
      else if (opt.algorithm == "simple-lca") params = LCAPredictionModelGPU::simple();
      else if (opt.algorithm == "megan-lca" || opt.algorithm == "ic-megan-lca") {
        if (opt.minsupport >= 2)
          // the filter counts rises of the running maximum in record-set order, and the reference orders records of
          // equal (qstart, qstop) by heap address (alignmentrecord.hh:479): its own -c >= 2 output is allocator
          // dependent.  This build uses the order of the input file, which is one of the orders the reference can take.
          std::cerr << "taxator-b200: note: with -c >= 2 the reference's result depends on its allocator for records with "
                       "equal query ranges; using input-file order" << std::endl;
        params = LCAPredictionModelGPU::megan(opt.ignore_unclassified, opt.toppercent, opt.minscore, (int)opt.minsupport, opt.maxevalue);
      } else if (opt.algorithm == "n-best-lca") params = LCAPredictionModelGPU::nbest((int)opt.nbest);
      else {
        std::cout << "classification algorithm can either be: rpa (default), simple-lca, megan-lca, ic-megan-lca, n-best-lca" << std::endl;
        return EXIT_FAILURE;
      }
      // CUDA context creation (about a second) overlaps the loading of the taxonomy and the mapping
      const int device = opt.gpus.empty() ? 0 : opt.gpus[0];
      std::thread warm([device]() { trpa_ctx* c = trpa_create(device, nullptr); if (c) trpa_destroy(c); });
      struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_joiner{warm};
      FlatTaxonomy tax = load_taxonomy_from_environment(opt.ranks, opt.delete_unmarked);
      SeqIdMapping mapping = load_mapping(opt.mapping);
      std::ofstream logsink(opt.logfile.c_str(), std::ios_base::app);
      warm.join();
      LCAPredictionModelGPU model(&tax, params, device);
      std::ios::sync_with_stdio(false);
      if (opt.legacy_ingest || opt.alignments_sorted) {
        RecordSetReader reader(std::cin, mapping, tax, opt.split_alignments, opt.alignments_sorted);
        run_prediction_stream(
            reader, tax, opt.batch_segments,
            [&](std::vector<RecordSet>& sets, std::vector<PredictionRecord>& precs, std::ostream& log) { model.predictBatch(sets, precs, log); },
            std::cout, logsink);
      } else {
        // block-parallel parser -> flat tables (+ e-values) -> GPU -> parallel GFF3 formatter; no sequence stores
        IngestOptions io;
        io.split = opt.split_alignments;
        io.block_bytes = opt.batch_bytes;
        io.need_stores = false;
        io.want_evalue = true;
        const SeqStore no_store;
        run_prediction_fast_blocks(
            stdin, mapping, tax, no_store, no_store, io,
            [&](FlatBlock& b) {
              model.predictFlat(b.segs.data(), (uint32_t)b.segs.size(), b.cands.data(), (uint32_t)b.cands.size(),
                                b.evalue.data(), b.res.data());
            },
            std::cout, nullptr, nullptr);
      }
      return EXIT_SUCCESS;
    }
    if (opt.dataformat != "nucleotide" && opt.dataformat != "protein") {
      std::cout << "data format can either be nucleotide or protein" << std::endl;
      return EXIT_FAILURE;
    }
    const bool protein = opt.dataformat == "protein";
    auto t0 = std::chrono::steady_clock::now();
    // CUDA context creation (about a second) overlaps the loading of the input files
    std::thread warm([&opt]() {
      for (int dev : opt.gpus) { trpa_ctx* c = trpa_create(dev, nullptr); if (c) trpa_destroy(c); }
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_joiner{warm};
    FlatTaxonomy tax = load_taxonomy_from_environment(opt.ranks, opt.delete_unmarked);
    SeqIdMapping mapping = load_mapping(opt.mapping);
    std::ofstream logsink(opt.logfile.c_str(), std::ios_base::app);
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    auto ta = std::chrono::steady_clock::now();
    const double load_tax_s = secs(t0, ta);

    SeqStore q_store;
    if (is_refpack_file(opt.query)) q_store = load_refpack(opt.query);
    else if (opt.query_index.empty()) {
      std::cerr << "Loading '" << opt.query;
      q_store = load_fasta_inmemory(opt.query);
      std::cerr << "' (total=" << q_store.size() << ")" << std::endl;
    } else q_store = load_fasta_indexed(opt.query, opt.query_index);
    auto tb = std::chrono::steady_clock::now();
    const double load_q_s = secs(ta, tb);
    SeqStore db_store = is_refpack_file(opt.ref) ? load_refpack(opt.ref)
                                                 : load_fasta_indexed(opt.ref, opt.ref_index.empty() ? opt.ref + ".fai" : opt.ref_index);
    auto tc = std::chrono::steady_clock::now();
    const double load_r_s = secs(tb, tc);

    warm.join();
    RPAPredictionModelGPU model(&tax, q_store, db_store, opt.filterout, opt.toppercent, protein, opt.gpus);
    auto t1 = std::chrono::steady_clock::now();
    const double load_gpu_s = secs(tc, t1);
    if (!opt.write_refpack.empty() || !opt.write_querypack.empty()) {
      // the GPU has packed the stores: write them out for later runs and stop
      for (int which = 0; which < 2; ++which) {
        const std::string& path = which ? opt.write_refpack : opt.write_querypack;
        if (path.empty()) continue;
        std::vector<uint64_t> woff; std::vector<uint32_t> len; std::vector<char> payload; int alphabet = 0;
        model.exportStore(which, woff, len, payload, alphabet);
        write_refpack(path, alphabet, which ? db_store.ids : q_store.ids, woff, len, payload);
        std::cerr << "taxator-b200: wrote " << path << " (" << len.size() << " sequences, " << payload.size() << " packed bytes)" << std::endl;
      }
      return EXIT_SUCCESS;
    }

    std::ios::sync_with_stdio(false);
    uint64_t total_sets = 0;
    StageTimes stage_times;
    if (opt.legacy_ingest || opt.alignments_sorted) {
      RecordSetReader reader(std::cin, mapping, tax, opt.split_alignments, opt.alignments_sorted);
      total_sets = run_prediction_stream(
          reader, tax, opt.batch_segments,
          [&](std::vector<RecordSet>& sets, std::vector<PredictionRecord>& precs, std::ostream& log) {
            model.predictBatch(sets, precs, log);
          },
          std::cout, logsink);
    } else {
      // parallel block parser -> flat tables -> GPU -> parallel GFF3 formatter, the stages overlapped
      IngestOptions io;
      io.split = opt.split_alignments;
      io.block_bytes = opt.batch_bytes;
      const bool want_log = opt.logfile != "/dev/null";
      if (want_log) {
        // -l: the reference's verbose per-alignment log (hh:341-838).  The first GPU places the block and records the
        // alignments it consumed; the log is written from that trace, segment by segment in input order.
        VerboseLogContext lc;
        lc.tax = &tax; lc.q_store = &q_store; lc.db_store = &db_store; lc.protein = protein;
        lc.exclude_factor = opt.filterout;
        lc.reeval_bandwidth_factor = 1. - opt.toppercent;
        std::vector<trpa_trace_entry> trace;
        total_sets = run_prediction_fast_blocks(
            stdin, mapping, tax, q_store, db_store, io,
            [&](FlatBlock& b) {
              model.predictFlatTraced(b.segs.data(), (uint32_t)b.segs.size(), b.cands.data(), (uint32_t)b.cands.size(), b.res.data(), trace);
              size_t t = 0;
              for (size_t i = 0; i < b.segs.size(); ++i) {
                size_t e = t;
                while (e < trace.size() && trace[e].seg == i) ++e;
                write_segment_log(lc, std::string(b.meta[i].qid, b.meta[i].qid_len), b.segs[i].query_seq, b.cands.data() + b.segs[i].cand_begin,
                                  b.segs[i].cand_count, b.res[i], trace.data() + t, e - t, logsink);
                t = e;
              }
            },
            std::cout, nullptr, &model.mutable_stats(), &stage_times);
      } else
      total_sets = run_prediction_fast(
          stdin, mapping, tax, q_store, db_store, io,
          [&](const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands, trpa_result* res) {
            model.predictFlat(segs, n_segs, cands, n_cands, res);
          },
          std::cout, nullptr, &model.mutable_stats(), &stage_times);
    }
    auto t2 = std::chrono::steady_clock::now();
    if (opt.timing) {
      auto st = model.stats();
      const double load_s = std::chrono::duration<double>(t1 - t0).count();
      const double run_s = std::chrono::duration<double>(t2 - t1).count();
      std::cerr << "taxator-b200: load " << load_s << " s, predict " << run_s << " s, " << total_sets << " segments, "
                << st.alignments << " alignments, " << st.cells << " cells, " << (run_s > 0 ? st.cells / run_s / 1e9 : 0)
                << " GCUPS" << std::endl;
      if (stage_times.blocks)
        std::cerr << "taxator-b200: stages busy: ingest " << stage_times.ingest_s << " s, gpu " << stage_times.predict_s
                  << " s, output " << stage_times.output_s << " s over " << stage_times.blocks << " blocks" << std::endl;
      std::cerr << "taxator-b200: load: taxonomy+mapping " << load_tax_s << " s, query store " << load_q_s << " s, reference store "
                << load_r_s << " s, GPU set-up " << load_gpu_s << " s" << std::endl;
    }
    return EXIT_SUCCESS;
  } catch (std::exception& e) {
    std::cerr << "An unrecoverable error occurred: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}
