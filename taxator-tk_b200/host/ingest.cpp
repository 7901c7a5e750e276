#include "ingest.h"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>

namespace taxator_b200 {

// ------------------------------------------------------------------------------------ RefResolver
RefResolver::RefResolver(const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& db_store) {
  map_.reserve(mapping.map.size() * 2);
  for (const auto& kv : mapping.map) {
    Entry e;
    auto n = tax.index.find(kv.second);
    e.node = n == tax.index.end() ? kNone : n->second;
    auto o = db_store.index.find(kv.first);
    e.ordinal = o == db_store.index.end() ? kNone : o->second;
    map_.emplace(std::string_view(kv.first), e);
  }
}

// ------------------------------------------------------------------------------------ line helpers
namespace {

unsigned pick_threads(unsigned want) {
  unsigned t = want ? want : std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  if (t > 32) t = 32;
  return t;
}

template <class F>
void parallel_for(unsigned T, const F& f) {
  if (T <= 1) { f(0u); return; }
  std::vector<std::thread> th;
  th.reserve(T);
  for (unsigned t = 0; t < T; ++t) th.emplace_back([&f, t]() { f(t); });
  for (auto& x : th) x.join();
}

inline const char* line_end(const char* p, const char* end) {
  const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
  return e ? e : end;
}
// start of the line that contains position pos (pos > begin), i.e. one past the previous '\n'
inline const char* line_start(const char* begin, const char* pos) {
  const char* p = pos;
  while (p > begin && p[-1] != '\n') --p;
  return p;
}
inline bool is_comment(const char* p, const char* e) { return p < e && *p == '#'; }   // ignoreLine, utils.hh:46-53
// query identifier of a data line: first TAB field after the optional mask character
inline std::string_view line_qid(const char* p, const char* e) {
  if (p < e && *p == '*') ++p;
  const char* t = (const char*)memchr(p, '\t', (size_t)(e - p));
  return std::string_view(p, (size_t)((t ? t : e) - p));
}

inline bool parse_u32(const char* b, const char* e, uint32_t& v) {
  if (b == e) return false;
  uint64_t x = 0;
  for (const char* p = b; p < e; ++p) {
    const unsigned d = (unsigned)(*p - '0');
    if (d > 9u) return false;
    x = x * 10u + d;
    if (x > 0xffffffffull) return false;
  }
  v = (uint32_t)x;
  return true;
}
template <class T>
inline bool parse_fp(const char* b, const char* e, T& v) {
  if (b == e) return false;
  if (*b == '+' && e - b > 1 && b[1] != '-' && b[1] != '+') ++b;   // lexical_cast accepts an explicit plus sign, from_chars does not
  auto r = std::from_chars(b, e, v);
  return r.ec == std::errc() && r.ptr == e;
}

struct Rec {
  const char* qid; const char* rid;
  uint32_t qid_len, rid_len;
  uint32_t qstart, qstop, qlen, ord, node, rstart, rstop, ident, alnlen;
  float score;
  uint32_t masked;
  double evalue;
};

struct Piece {
  std::vector<trpa_segment> segs;
  std::vector<trpa_candidate> cands;
  std::vector<double> evalue;
  std::vector<SegMeta> meta;
  std::exception_ptr err;
  size_t err_off = 0;
  bool err_has_line = false;
};

}  // namespace

// ------------------------------------------------------------------------------------ FastIngest
FastIngest::FastIngest(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const RefResolver& refs,
                       const SeqStore& q_store, const IngestOptions& opt)
    : in_(in), mapping_(mapping), tax_(tax), refs_(refs), q_store_(q_store), opt_(opt) {
  if (opt_.block_bytes < 4096) opt_.block_bytes = 4096;
}

bool FastIngest::next(FlatBlock& out) {
  out.segs.clear(); out.cands.clear(); out.meta.clear(); out.res.clear();
  std::vector<char> buf;
  buf.swap(carry_);
  size_t target = std::max(opt_.block_bytes, buf.size() * 2);
  size_t len = 0;
  for (;;) {
    while (!eof_ && buf.size() < target) {
      const size_t old = buf.size();
      const size_t want = std::min<size_t>(target - old, 64u << 20);
      buf.resize(old + want);
      const size_t got = fread(buf.data() + old, 1, want, in_);
      buf.resize(old + got);
      if (got < want) eof_ = true;
    }
    if (buf.empty()) return false;
    const char* b = buf.data();
    const char* end = b + buf.size();
    if (eof_) { len = buf.size(); break; }
    // keep the trailing partial line and the (possibly unfinished) last query group for the next block
    const char* last_nl = end;
    while (last_nl > b && last_nl[-1] != '\n') --last_nl;        // one past the last '\n'
    if (last_nl == b) { target *= 2; continue; }
    const char* ls = line_start(b, last_nl - 1);                 // last complete line
    while (ls > b && is_comment(ls, last_nl)) ls = line_start(b, ls - 1);
    const char* cut = ls;
    if (!is_comment(ls, last_nl)) {
      const std::string_view q = line_qid(ls, line_end(ls, end));
      while (cut > b) {
        const char* ps = line_start(b, cut - 1);
        const char* pe = cut - 1;
        if (is_comment(ps, pe) || line_qid(ps, pe) == q) cut = ps;
        else break;
      }
    } else cut = last_nl;                                        // only comments: nothing to keep together
    if (cut == b) { target *= 2; continue; }                     // one query fills the block: read on
    len = (size_t)(cut - b);
    carry_.assign(cut, end);
    break;
  }
  out.text.swap(buf);
  out.first_line = lines_done_ + 1;
  size_t nl = 0;
  for (const char* p = out.text.data(), *e = p + len; p < e;) {
    const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
    if (!q) { ++nl; break; }
    ++nl;
    p = q + 1;
  }
  lines_done_ += nl;
  parse_block(out, len);
  return true;
}

void FastIngest::parse_block(FlatBlock& out, size_t len) {
  const char* const base = out.text.data();
  const char* const end = base + len;
  const unsigned T = len < opt_.min_parallel_bytes ? 1u : pick_threads(opt_.threads);
  // chunk boundaries at query-group starts
  std::vector<const char*> cutp(T + 1, end);
  cutp[0] = base;
  for (unsigned t = 1; t < T; ++t) {
    const char* p = base + (len / T) * t;
    if (p <= cutp[t - 1]) { cutp[t] = cutp[t - 1]; continue; }
    p = line_start(base, p);
    // previous data line
    const char* ps = p;
    std::string_view prevq;
    bool have_prev = false;
    while (ps > base) {
      const char* s = line_start(base, ps - 1);
      if (!is_comment(s, ps - 1)) { prevq = line_qid(s, ps - 1); have_prev = true; break; }
      ps = s;
    }
    if (have_prev) {
      while (p < end) {
        const char* e = line_end(p, end);
        if (!is_comment(p, e) && line_qid(p, e) != prevq) break;
        p = e < end ? e + 1 : end;
      }
    }
    cutp[t] = std::max(p, cutp[t - 1]);
  }
  std::vector<Piece> pieces(T);
  const bool split = opt_.split, need_stores = opt_.need_stores, want_evalue = opt_.want_evalue;

  parallel_for(T, [&](unsigned t) {
    Piece& P = pieces[t];
    const char* p = cutp[t];
    const char* const pend = cutp[t + 1];
    std::vector<Rec> grp;
    std::vector<uint32_t> order;
    const char* cur_line = p;
    auto emit_group = [&]() {
      const uint32_t n = (uint32_t)grp.size();
      if (!n) return;
      order.resize(n);
      for (uint32_t i = 0; i < n; ++i) order[i] = i;
      if (split)   // (qstart, qstop, arrival): alignmentrecord.hh:480, see records.cpp
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
          if (grp[a].qstart != grp[b].qstart) return grp[a].qstart < grp[b].qstart;
          return grp[a].qstop < grp[b].qstop;
        });
      uint32_t i = 0;
      while (i < n) {
        uint32_t j = i + 1;
        if (split) {
          uint32_t run_stop = grp[order[i]].qstop;
          while (j < n && grp[order[j]].qstart <= run_stop) { run_stop = std::max(run_stop, grp[order[j]].qstop); ++j; }
        } else j = n;
        trpa_segment sg;
        sg.cand_begin = (uint32_t)P.cands.size();
        sg.reserved = 0;
        uint32_t cnt = 0;
        for (uint32_t k = i; k < j; ++k) {
          const Rec& r = grp[order[k]];
          if (r.masked) continue;   // active_records, hh:350-356
          if (need_stores && r.ord == RefResolver::kNone) throw SequenceNotFound("bad sequence identifier: " + std::string(r.rid, r.rid_len));
          trpa_candidate c;
          c.ref_seq = need_stores ? r.ord : 0u; c.rstart = r.rstart; c.rstop = r.rstop; c.qstart = r.qstart; c.qstop = r.qstop;
          c.score = r.score; c.identities = r.ident; c.alnlen = r.alnlen; c.node = r.node;
          P.cands.push_back(c);
          if (want_evalue) P.evalue.push_back(r.evalue);
          ++cnt;
        }
        sg.cand_count = cnt;
        const Rec& first = grp[order[i]];
        // the reference looks the query up only when it realigns (n >= 2, hh:415)
        sg.query_seq = need_stores && cnt >= 2 ? q_store_.ordinal(std::string(first.qid, first.qid_len)) : 0u;
        P.segs.push_back(sg);
        P.meta.push_back(SegMeta{first.qid, first.qid_len, first.qlen, cnt ? 1u : 0u});
        i = j;
      }
      grp.clear();
    };
    try {
      while (p < pend) {
        const char* e = line_end(p, pend);
        cur_line = p;
        const char* next = e < pend ? e + 1 : pend;
        if (is_comment(p, e)) { p = next; continue; }
        Rec r;
        bool ok = e - p > 1;
        const char* f[12];
        const char* fe[12];
        int nf = 0;
        if (ok) {
          r.masked = *p == '*' ? 1u : 0u;
          const char* q = p + r.masked;
          while (nf < 11) {
            const char* tpos = (const char*)memchr(q, '\t', (size_t)(e - q));
            f[nf] = q; fe[nf] = tpos ? tpos : e;
            ++nf;
            if (!tpos) break;
            q = tpos + 1;
          }
          ok = nf == 11;
        }
        double evalue;
        ok = ok && parse_u32(f[1], fe[1], r.qstart) && parse_u32(f[2], fe[2], r.qstop) && parse_u32(f[3], fe[3], r.qlen) &&
             parse_u32(f[5], fe[5], r.rstart) && parse_u32(f[6], fe[6], r.rstop) && parse_fp(f[7], fe[7], r.score) &&
             parse_fp(f[8], fe[8], evalue) && parse_u32(f[9], fe[9], r.ident) && parse_u32(f[10], fe[10], r.alnlen) &&
             r.qstart <= r.qstop;
        const RefResolver::Entry* ent = nullptr;
        if (ok) {
          ent = refs_.find(std::string_view(f[4], (size_t)(fe[4] - f[4])));
          ok = ent && ent->node != RefResolver::kNone;
        }
        if (ok) {
          r.qid = f[0]; r.qid_len = (uint32_t)(fe[0] - f[0]);
          r.rid = f[4]; r.rid_len = (uint32_t)(fe[4] - f[4]);
          r.ord = ent->ordinal; r.node = ent->node;
          r.evalue = evalue;
        } else {
          // anything unusual: the record-at-a-time parser decides (same acceptance, same errors)
          AlignmentRecord* a = parse_alignment_line(std::string(p, (size_t)(e - p)), mapping_, tax_);
          r.masked = a->masked ? 1u : 0u;
          r.qstart = a->qstart; r.qstop = a->qstop; r.qlen = a->qlen; r.rstart = a->rstart; r.rstop = a->rstop;
          r.score = a->score; r.ident = a->identities; r.alnlen = a->alnlen; r.node = a->node; r.evalue = a->evalue;
          const std::string_view qv = line_qid(p, e);
          r.qid = qv.data(); r.qid_len = (uint32_t)qv.size();
          // the reference id is the 5th field of the line however odd the rest was
          const char* q = p + r.masked;
          for (int k = 0; k < 4; ++k) { const char* tpos = (const char*)memchr(q, '\t', (size_t)(e - q)); q = tpos ? tpos + 1 : e; }
          const char* tpos = (const char*)memchr(q, '\t', (size_t)(e - q));
          r.rid = q; r.rid_len = (uint32_t)((tpos ? tpos : e) - q);
          const RefResolver::Entry* en = refs_.find(std::string_view(r.rid, r.rid_len));
          r.ord = en ? en->ordinal : RefResolver::kNone;
          delete a;
        }
        if (!grp.empty() && (grp.back().qid_len != r.qid_len || memcmp(grp.back().qid, r.qid, r.qid_len) != 0)) emit_group();
        grp.push_back(r);
        p = next;
      }
      cur_line = pend;
      emit_group();
    } catch (ParsingError&) {
      P.err = std::current_exception(); P.err_off = (size_t)(cur_line - base); P.err_has_line = true;
    } catch (TaxonMappingNotFound&) {
      P.err = std::current_exception(); P.err_off = (size_t)(cur_line - base); P.err_has_line = true;
    } catch (TaxonNotFound&) {
      P.err = std::current_exception(); P.err_off = (size_t)(cur_line - base); P.err_has_line = true;
    } catch (...) {
      P.err = std::current_exception(); P.err_off = (size_t)(cur_line - base); P.err_has_line = false;
    }
  });

  // the first failure in file order wins (the record-at-a-time path stops there)
  for (unsigned t = 0; t < T; ++t) {
    if (!pieces[t].err) continue;
    if (!pieces[t].err_has_line) std::rethrow_exception(pieces[t].err);
    uint64_t line = out.first_line;
    for (const char* q = base; q < base + pieces[t].err_off;) {
      const char* nlp = (const char*)memchr(q, '\n', (size_t)(base + pieces[t].err_off - q));
      if (!nlp) break;
      ++line;
      q = nlp + 1;
    }
    try { std::rethrow_exception(pieces[t].err); }
    catch (TaxatorError& e) { throw ParsingError(std::string(e.what()) + " (line " + std::to_string(line) + ")"); }
  }
  // concatenate the pieces
  size_t ns = 0, nc = 0;
  std::vector<size_t> soff(T), coff(T);
  for (unsigned t = 0; t < T; ++t) { soff[t] = ns; coff[t] = nc; ns += pieces[t].segs.size(); nc += pieces[t].cands.size(); }
  if (nc >= 0xfffffff0ull || ns >= 0xfffffff0ull) throw TaxatorError("alignment block too large; lower --batch-bytes");
  out.segs.resize(ns); out.cands.resize(nc); out.meta.resize(ns);
  out.evalue.resize(opt_.want_evalue ? nc : 0);
  parallel_for(T, [&](unsigned t) {
    Piece& P = pieces[t];
    for (size_t i = 0; i < P.segs.size(); ++i) {
      trpa_segment s = P.segs[i];
      s.cand_begin += (uint32_t)coff[t];
      out.segs[soff[t] + i] = s;
    }
    if (!P.meta.empty()) memcpy(&out.meta[soff[t]], P.meta.data(), P.meta.size() * sizeof(SegMeta));
    if (!P.cands.empty()) memcpy(&out.cands[coff[t]], P.cands.data(), P.cands.size() * sizeof(trpa_candidate));
    if (!P.evalue.empty()) memcpy(&out.evalue[coff[t]], P.evalue.data(), P.evalue.size() * sizeof(double));
  });
}

// ------------------------------------------------------------------------------------ GFF3 output
namespace {

inline void put_u32(std::string& s, uint32_t v) {
  char tmp[12];
  auto r = std::to_chars(tmp, tmp + sizeof(tmp), v);
  s.append(tmp, (size_t)(r.ptr - tmp));
}
inline void put_float(std::string& s, float v) {   // operator<<(float): %g, 6 significant digits
  char tmp[32];
  const int n = snprintf(tmp, sizeof(tmp), "%g", (double)v);
  s.append(tmp, (size_t)n);
}

}  // namespace

void format_block(const FlatBlock& b, const FlatTaxonomy& tax, float& carry_ival, float& carry_signal, unsigned threads,
                  std::string& out, std::ostream* statslog) {
  const size_t n = b.segs.size();
  // what a reused PredictionRecord would carry into n == 0 sets (taxator.cpp:66, hh:359-368)
  std::vector<float> ival(n), signal(n);
  for (size_t i = 0; i < n; ++i) {
    const trpa_result& r = b.res[i];
    if (r.kind == TRPA_KIND_NONE) { ival[i] = carry_ival; signal[i] = carry_signal; }
    else { ival[i] = r.ival; signal[i] = r.kind == TRPA_KIND_PLACED ? r.signal : 0.f; }
    carry_ival = ival[i]; carry_signal = signal[i];
  }
  const unsigned T = n < 4096 ? 1u : pick_threads(threads);
  std::vector<std::string> parts(T);
  parallel_for(T, [&](unsigned t) {
    std::string& s = parts[t];
    const size_t lo = n * t / T, hi = n * (t + 1) / T;
    s.reserve((hi - lo) * 96);
    for (size_t i = lo; i < hi; ++i) {
      const trpa_result& r = b.res[i];
      const SegMeta& m = b.meta[i];
      const bool none = r.kind == TRPA_KIND_NONE;
      const uint32_t lower = none ? tax.root : r.lower_node, upper = none ? tax.root : r.upper_node;
      const uint32_t rtax = none ? tax.root : r.rtax_node, support = none ? 0u : r.support;
      s.append(m.qid, m.qid_len);
      s += "\ttaxator-tk\tsequence_feature\t";
      put_u32(s, none ? 1u : r.qrstart); s += '\t';
      put_u32(s, none ? m.qlen : r.qrstop); s += '\t';
      if (signal[i] != signal[i]) s += '.';
      else put_float(s, signal[i]);
      s += "\t.\t.\tseqlen=";
      put_u32(s, m.qlen);
      s += ";tax=";
      uint32_t last_support = 0, node = lower;   // predictionrecord.hh:248-308, see PredictionRecord::print
      while (node != upper) {
        if (support != last_support) { s += tax.taxid[node]; s += ':'; put_u32(s, support); s += '-'; last_support = support; }
        node = tax.parent[node];
      }
      s += tax.taxid[node];
      if (support != last_support) { s += ':'; put_u32(s, support); }
      s += ";rtax=";
      s += tax.taxid[rtax];
      if (ival[i] >= 0. && ival[i] < 1.) { s += ";ival="; put_float(s, ival[i]); }
      s += '\n';
    }
  });
  size_t total = 0;
  for (auto& p : parts) total += p.size();
  out.clear();
  out.reserve(total);
  for (auto& p : parts) out += p;
  if (statslog) {   // the per-segment STATS line of the reference's log (hh:834-837), without its timers
    for (size_t i = 0; i < n; ++i) {
      const trpa_result& r = b.res[i];
      *statslog << "STATS\t" << r.qrstart << ':' << r.qrstop << '@' << std::string_view(b.meta[i].qid, b.meta[i].qid_len)
                << '\t' << b.segs[i].cand_count << '\t' << r.n_pass0 << '\t' << r.n_pass1 << '\t' << r.n_pass2 << '\t'
                << (r.n_pass0 + r.n_pass1 + r.n_pass2) << '\n';
    }
  }
}

// ------------------------------------------------------------------------------------ pipeline
namespace {

template <class T>
class BoundedQueue {
 public:
  explicit BoundedQueue(size_t cap) : cap_(cap) {}
  bool push(T v) {   // false: the consumer went away
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return q_.size() < cap_ || closed_; });
    if (closed_) return false;
    q_.push_back(std::move(v));
    cv_.notify_all();
    return true;
  }
  bool pop(T& v) {   // false: closed and drained
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty() || done_ || closed_; });
    if (q_.empty()) return false;
    v = std::move(q_.front());
    q_.pop_front();
    cv_.notify_all();
    return true;
  }
  void finish() { std::lock_guard<std::mutex> l(m_); done_ = true; cv_.notify_all(); }    // producer is done
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; q_.clear(); cv_.notify_all(); }  // abort

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
  size_t cap_;
  bool done_ = false, closed_ = false;
};

}  // namespace

uint64_t run_prediction_fast(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& q_store,
                             const SeqStore& db_store, const IngestOptions& opt, const FlatPredictor& predict,
                             std::ostream& out, std::ostream* statslog, PredictStats* stats, StageTimes* times) {
  return run_prediction_fast_blocks(
      in, mapping, tax, q_store, db_store, opt,
      [&](FlatBlock& b) { predict(b.segs.data(), (uint32_t)b.segs.size(), b.cands.data(), (uint32_t)b.cands.size(), b.res.data()); },
      out, statslog, stats, times);
}

uint64_t run_prediction_fast_blocks(FILE* in, const SeqIdMapping& mapping, const FlatTaxonomy& tax, const SeqStore& q_store,
                                    const SeqStore& db_store, const IngestOptions& opt, const BlockPredictor& predict,
                                    std::ostream& out, std::ostream* statslog, PredictStats* stats, StageTimes* times) {
  out << kGFF3Header;
  typedef std::chrono::steady_clock Clock;
  auto secs = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  StageTimes st;
  RefResolver refs(mapping, tax, db_store);
  FastIngest ingest(in, mapping, tax, refs, q_store, opt);
  typedef std::unique_ptr<FlatBlock> BlockPtr;
  BoundedQueue<BlockPtr> parsed(2), placed(2);
  // written blocks go back to the producer: their tables keep their capacity
  std::mutex free_m;
  std::vector<BlockPtr> free_blocks;
  std::mutex err_m;
  std::exception_ptr err;
  auto fail = [&](std::exception_ptr e) {
    { std::lock_guard<std::mutex> l(err_m); if (!err) err = e; }
    parsed.close(); placed.close();
  };
  uint64_t total = 0;

  std::thread producer([&]() {
    try {
      for (;;) {
        BlockPtr b;
        {
          std::lock_guard<std::mutex> l(free_m);
          if (!free_blocks.empty()) { b = std::move(free_blocks.back()); free_blocks.pop_back(); }
        }
        if (!b) b.reset(new FlatBlock());
        const auto t0 = Clock::now();
        const bool more = ingest.next(*b);
        st.ingest_s += secs(t0, Clock::now());
        if (!more) break;
        if (b->segs.empty()) continue;
        if (!parsed.push(std::move(b))) return;
      }
      parsed.finish();
    } catch (...) { fail(std::current_exception()); }
  });
  std::thread writer([&]() {
    try {
      float carry_ival = -1.f, carry_signal = 0.f;   // a fresh PredictionRecord
      std::string text;
      BlockPtr b;
      while (placed.pop(b)) {
        const auto t0 = Clock::now();
        format_block(*b, tax, carry_ival, carry_signal, opt.threads, text, statslog);
        out.write(text.data(), (std::streamsize)text.size());
        st.output_s += secs(t0, Clock::now());
        total += b->segs.size();
        if (stats) {
          for (const trpa_result& r : b->res) {
            stats->segments++;
            stats->alignments += (uint64_t)r.n_pass0 + r.n_pass1 + r.n_pass2;
            stats->cells += r.cells;
          }
        }
        {
          std::lock_guard<std::mutex> l(free_m);
          if (free_blocks.size() < 8) free_blocks.push_back(std::move(b));
        }
        b.reset();
      }
    } catch (...) { fail(std::current_exception()); }
  });
  try {
    BlockPtr b;
    while (parsed.pop(b)) {
      b->res.resize(b->segs.size());
      const auto t0 = Clock::now();
      predict(*b);
      st.predict_s += secs(t0, Clock::now());
      st.blocks++;
      if (!placed.push(std::move(b))) break;
    }
    placed.finish();
  } catch (...) { fail(std::current_exception()); }
  producer.join();
  writer.join();
  out.flush();
  if (times) *times = st;
  if (err) std::rethrow_exception(err);
  return total;
}

}  // namespace taxator_b200
