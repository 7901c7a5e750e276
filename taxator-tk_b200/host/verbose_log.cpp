// The verbose per-alignment log of `taxator -l` (ID / NUMREF / PASS / +ALN / current ... node / EXT / SCORE / NUMALN /
// NUMOUTGRP / RANGE / STATS lines of core/src/taxonpredictionmodelsequence.hh:341-838) for segments that were placed on
// the GPU.  The GPU records every alignment it consumed, in the reference's order (trpa_set_trace /
// trpa_batch_trace); this file walks predict()'s control flow once more on the host, takes the distances from that
// trace instead of aligning, and writes the reference's lines at the reference's places.  It is a log writer, not a
// second predictor: at the end it checks that the walk arrived at the very result record the GPU produced and that
// it consumed exactly the trace, and throws otherwise.
// The rendering of protein alignments (hh:534, 637, 783, 803 print the SeqAn Align object) needs the alignment
// itself, which the score-only kernels never build: the logged pairs are re-traced here on the host with SeqAn's
// tie order (diagonal >= vertical >= horizontal) and printed in SeqAn's block format
// (includes-external/seqan/align/align_base.h:477-558).  Nucleotide alignments print as an empty line like in the
// reference (its edit-distance path builds no Align object, hh:133-171).
#include "verbose_log.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <list>
#include <set>
#include <sstream>

#include "../csrc/blosum62_table.h"

namespace taxator_b200 {

namespace {

const char kAaOrder[] = "ABCDEFGHIJKLMNOPQRSTUVWYZX*";   // SeqAn AminoAcid ordinals
int aa_ordinal(int c) {
  if (c >= 'a' && c <= 'z') c -= 32;
  for (int i = 0; i < 27; ++i) if (kAaOrder[i] == c) return i;
  return 25;
}

// getSequence(id, start, stop, left_ext, right_ext) (hh:856-880) on the host store, protein only (forward)
std::string fetch_aa(const SeqStore& st, uint32_t seq, uint64_t start, uint64_t stop, uint64_t left_ext, uint64_t right_ext) {
  uint64_t ns, ne;
  if (start <= stop) { ns = left_ext < start ? start - left_ext : 1; ne = stop + right_ext; }
  else { ns = right_ext < stop ? stop - right_ext : 1; ne = start + left_ext; }
  const uint64_t L = st.len[seq];
  if (ne > L) ne = L;
  uint64_t b = ns - 1; if (b > L) b = L;
  uint64_t e = ne > b ? ne : b; if (e > L) e = L;
  std::string s = st.chars.substr(st.off[seq] + b, e - b);
  for (char& c : s) c = kAaOrder[aa_ordinal((unsigned char)c)];
  return s;
}

// globalAlignment(align, Blosum62(), LinearGaps()) (hh:199-213): A = row 0 (horizontal), B = row 1 (vertical)
void render_protein_alignment(const std::string& A, const std::string& B, std::ostream& os) {
  const int la = (int)A.size(), lb = (int)B.size();
  std::vector<uint8_t> dir((size_t)(la + 1) * (lb + 1), 0);   // 0 diagonal, 1 vertical, 2 horizontal
  std::vector<int> prev(la + 1), cur(la + 1);
  for (int j = 0; j <= la; ++j) { prev[j] = -j; dir[j] = 2; }
  for (int i = 1; i <= lb; ++i) {
    cur[0] = -i;
    dir[(size_t)i * (la + 1)] = 1;
    const int bi = aa_ordinal((unsigned char)B[i - 1]);
    for (int j = 1; j <= la; ++j) {
      int best = prev[j - 1] + TRPA_BLOSUM62[aa_ordinal((unsigned char)A[j - 1])][bi];
      uint8_t d = 0;
      if (prev[j] - 1 > best) { best = prev[j] - 1; d = 1; }
      if (cur[j - 1] - 1 > best) { best = cur[j - 1] - 1; d = 2; }
      cur[j] = best;
      dir[(size_t)i * (la + 1) + j] = d;
    }
    prev.swap(cur);
  }
  std::string r0, r1;
  for (int i = lb, j = la; i > 0 || j > 0;) {
    const uint8_t d = dir[(size_t)i * (la + 1) + j];
    if (d == 0) { r0.push_back(A[--j]); r1.push_back(B[--i]); }
    else if (d == 1) { r0.push_back('-'); r1.push_back(B[--i]); }
    else { r0.push_back(A[--j]); r1.push_back('-'); }
  }
  std::reverse(r0.begin(), r0.end());
  std::reverse(r1.begin(), r1.end());
  // align_base.h:477-558
  const size_t end = std::min(r0.size(), r1.size());
  unsigned base = 0;
  for (size_t begin = 0; begin < end;) {
    const size_t w = std::min<size_t>(50, end - begin);
    char buf[20];
    snprintf(buf, sizeof(buf), "%7u", base);
    os << buf << ' ';
    base += (unsigned)w;
    for (size_t i = 1; i <= w; ++i) os << ((i % 10) == 0 ? ':' : (i % 5) == 0 ? '.' : ' ');
    os << " \n";
    os << "        " << r0.substr(begin, w) << '\n';
    os << "        ";
    for (size_t k = 0; k < w; ++k) os << ((r0[begin + k] != '-' && r1[begin + k] != '-' && r0[begin + k] == r1[begin + k]) ? '|' : ' ');
    os << '\n';
    os << "        " << r1.substr(begin, w) << '\n';
    os << '\n';
    begin += w;
  }
  os << '\n';
}

struct Aln { float dist, sim; };

}  // namespace

void write_segment_log(const VerboseLogContext& ctx, const std::string& qid, uint32_t query_seq, const trpa_candidate* cands_in, uint32_t n,
                       const trpa_result& res, const trpa_trace_entry* trace, size_t n_trace, std::ostream& logsink) {
  const FlatTaxonomy& T = *ctx.tax;
  const uint32_t root = T.root;
  auto name = [&](uint32_t node) -> const std::string& { return T.name[node]; };
  auto seqname = [&](long long a, long long b) { std::ostringstream o; o << a << ':' << b << '@' << qid; return o.str(); };
  logsink << std::fixed << std::setprecision(2);   // hh:347

  if (n == 0) {   // hh:359-368
    const std::string q = seqname(-1, -1);
    logsink << "ID\t" << q << std::endl;
    logsink << "  NUMREF\t" << n << std::endl << std::endl;
    logsink << "    RANGE\t" << name(root) << '\t' << name(root) << '\t' << name(root) << std::endl << std::endl;
    logsink << "STATS\t" << q << '\t' << n << "\t0\t0\t0\t0\t0\t0\t0\t.0" << std::endl << std::endl;
    return;
  }
  if (n == 1) {   // hh:371-388
    const std::string q = seqname(cands_in[0].qstart, cands_in[0].qstop);
    logsink << "ID\t" << q << std::endl;
    logsink << "  NUMREF\t" << n << std::endl;
    logsink << "  RANGE\t" << name(cands_in[0].node) << '\t' << name(cands_in[0].node) << '\t' << name(root) << std::endl << std::endl;
    logsink << "STATS\t" << q << '\t' << n << "\t0\t0\t0\t0\t0\t0\t0\t.0" << std::endl << std::endl;
    return;
  }
  // SortFilter (alignmentsfilter.hh:171-190)
  std::vector<trpa_candidate> rec(cands_in, cands_in + n);
  std::stable_sort(rec.begin(), rec.end(), [](const trpa_candidate& x, const trpa_candidate& y) {
    if (y.score < x.score) return true;
    if (y.score > x.score) return false;
    return y.identities < x.identities;
  });
  uint32_t qrstart = rec[0].qstart, qrstop = rec[0].qstop;
  for (uint32_t i = 1; i < n; ++i) { qrstart = std::min(qrstart, rec[i].qstart); qrstop = std::max(qrstop, rec[i].qstop); }
  const uint32_t qrlength = qrstop - qrstart + 1;
  const std::string qrseqname = seqname(qrstart, qrstop);
  logsink << "ID\t" << qrseqname << std::endl;
  logsink << "  NUMREF\t" << n << std::endl;

  auto fail = [&](const char* what) { throw TaxatorError(std::string("verbose log: replay of segment ") + qrseqname + " diverged from the GPU (" + what + ")"); };

  if (rec[0].alnlen == qrlength && rec[0].identities == qrlength) {   // hh:431-472
    const float best = rec[0].score;
    uint32_t lnode = rec[0].node, unode = root, i = 1;
    while (true) {
      if (i == n) { unode = root; break; }
      const float searchscore = rec[i].score;
      if (searchscore == best) {
        lnode = T.lca(lnode, rec[i].node);
        logsink << "    current ref/lower node: (" << searchscore << ") " << name(lnode) << " (+ " << name(rec[i].node) << " )" << std::endl;
      } else {
        const float uscore = searchscore;
        unode = lnode;
        do {
          const uint32_t cnode = rec[i].node;
          unode = T.lca(unode, cnode);
          logsink << "    current upper node: (" << uscore << ") " << name(unode) << " (+ " << name(cnode) << " at "
                  << static_cast<int>(T.depth[T.lca(cnode, lnode)]) << " )" << std::endl;
        } while (++i < n && rec[i].score == uscore);
        break;
      }
      ++i;
    }
    logsink << "  RANGE\t" << name(lnode) << '\t' << name(lnode) << '\t' << name(unode) << std::endl << std::endl;
    logsink << "STATS\t" << qrseqname << '\t' << n << "\t0\t0\t0\t0\t" << 0 << "\t0\t0\t.0" << std::endl << std::endl;
    if (res.kind != TRPA_KIND_IDENTICAL || res.lower_node != lnode || res.upper_node != unode) fail("identical-hit shortcut");
    return;
  }

  // alignments come from the GPU's trace, in order
  size_t tpos = 0;
  auto next_alignment = [&](uint32_t a, uint32_t b) -> Aln {
    if (tpos >= n_trace) fail("trace exhausted");
    const trpa_trace_entry& e = trace[tpos++];
    if (e.a != a || e.b != b) fail("unexpected pair in the trace");
    Aln r;
    if (!ctx.protein) {   // hh:133-171
      const int d = e.r0;
      const int llong = (int)std::max(e.len_a, e.len_b), lshort = (int)std::min(e.len_a, e.len_b);
      const int mismatch = d - (llong - lshort);
      r.dist = (float)d;
      r.sim = (float)(lshort - mismatch);
    } else {              // hh:173-242
      const int mutual = e.r0, self = (int)e.self;
      const unsigned int len = e.len_a + e.len_b - (unsigned int)e.r1;
      const float norm = len / static_cast<float>(self);
      r.dist = (self - 2 * mutual) * norm;
      r.sim = (2 * mutual) * norm;
    }
    return r;
  };
  // what `logsink << alignment << std::endl` prints (hh:58-63)
  auto print_alignment = [&](uint32_t a, uint32_t b) {
    if (ctx.protein && ctx.db_store && !ctx.db_store->packed && ctx.q_store && !ctx.q_store->packed) {
      auto seg = [&](uint32_t k) -> std::string {
        if (k == TRPA_TRACE_QUERY) return fetch_aa(*ctx.q_store, query_seq, qrstart, qrstop, 0, 0);
        return fetch_aa(*ctx.db_store, rec[k].ref_seq, rec[k].rstart, rec[k].rstop, rec[k].qstart - qrstart, qrstop - rec[k].qstop);
      };
      render_protein_alignment(seg(a), seg(b), logsink);
    }
    logsink << std::endl;
  };

  std::vector<float> querydistance(n, FLT_MAX), querysimilarity(n, .0f);
  unsigned pass_0_counter = 0, pass_0_counter_naive = 0, pass_1_counter = 0, pass_1_counter_naive = 0, pass_2_counter = 0,
           pass_2_counter_naive = 0;
  std::set<unsigned> qgroup;
  uint32_t anchors_support = 0, rtax = root, lca_allnodes = rec[0].node;
  {   // pass 0, hh:497-566
    logsink << std::endl << "  PASS\t0" << std::endl;
    const float threshold = ctx.reeval_bandwidth_factor * rec[0].score;
    unsigned index_best = 0;
    for (unsigned i = 0; i < n; ++i) {
      float dist, sim;
      const float qsearchscore = rec[i].score;
      const uint32_t qsearchmatch = rec[i].identities;
      const double qsearchpid = static_cast<double>(qsearchmatch) / qrlength;
      if (rec[i].alnlen == qrlength && rec[i].identities == qrlength) {
        qgroup.insert(i);
        dist = 0;
        sim = rec[i].identities;
        logsink << "    *ALN " << i << " <=> query\tdist=" << dist << "; sim=" << sim << "; qsearchscore=" << qsearchscore
                << "; qsearchmatch=" << qsearchmatch << "; qpid=1.0" << std::endl;
        ++pass_0_counter_naive;
      } else if (rec[i].score >= threshold) {
        qgroup.insert(i);
        const Aln a = next_alignment(i, TRPA_TRACE_QUERY);
        dist = a.dist;
        ++pass_0_counter; ++pass_0_counter_naive;
        sim = std::max(a.sim, static_cast<float>(rec[i].identities));
        const double qpid = static_cast<double>(sim) / qrlength;
        logsink << "    +ALN " << i << " <=> query\tdist=" << dist << "; sim=" << sim << "; qsearchscore=" << qsearchscore
                << "; qsearchmatch=" << qsearchmatch << "; qsearchpid=" << qsearchpid << "; qpid=" << qpid << std::endl;
        print_alignment(i, TRPA_TRACE_QUERY);
      } else {
        dist = FLT_MAX;
        sim = rec[i].identities;
      }
      querydistance[i] = dist; querysimilarity[i] = sim;
      if (dist < querydistance[index_best]) index_best = i;
      else if (dist == querydistance[index_best]) {
        if (sim > querysimilarity[index_best]) index_best = i;
        else if (sim == querysimilarity[index_best] && qsearchscore > rec[index_best].score) index_best = i;
      }
      anchors_support = std::max(anchors_support, static_cast<uint32_t>(sim));
      lca_allnodes = T.lca(lca_allnodes, rec[i].node);
    }
    rtax = rec[index_best].node;
    for (std::set<unsigned>::iterator it = qgroup.begin(); it != qgroup.end();) {
      if (querydistance[*it] != querydistance[index_best] || querysimilarity[*it] != querysimilarity[index_best] ||
          rec[*it].score != rec[index_best].score) qgroup.erase(it++);
      else {
        const uint32_t cnode = rec[*it].node;
        rtax = T.lca(rtax, cnode);
        logsink << "      current ref node: (" << querydistance[*it] << ") " << name(rtax) << " (+ " << name(cnode) << " )" << std::endl;
        ++it;
      }
    }
    logsink << "    NUMALN\t" << pass_0_counter << '\t' << pass_0_counter_naive - pass_0_counter << std::endl << std::endl;
    if (qgroup.empty()) fail("empty query group");
  }

  float ival_global = 0.f, bandfactor_max = 1.f;
  uint32_t lnode_global = rtax, unode_global = rtax;
  std::set<unsigned> outgroup;
  {   // pass 1, hh:576-733
    logsink << "  PASS\t1" << std::endl;
    unsigned lca_root_dist_min = 255;
    do {
      std::vector<std::pair<float, uint32_t>> bf;   // BandFactor (hh:259-323)
      const unsigned index_anchor = *qgroup.begin();
      qgroup.erase(qgroup.begin());
      const float qdist = querydistance[index_anchor];
      const uint32_t rnode = rec[index_anchor].node;
      bf.push_back(std::make_pair(0.f, rnode));
      uint32_t lnode = rtax;
      uint32_t unode = root;
      bool have_unode = false;
      float ldist = 0, udist = FLT_MAX;
      std::list<std::pair<unsigned, int>> outgroup_tmp;
      logsink << "      query: (" << qdist << ") unknown" << std::endl;
      pass_1_counter_naive += n - 1;
      double qpid_upper = 0., qpid_thresh_guarantee = 0., qpid_thresh_heuristic = 0.;
      int qsearchscore_thresh_heuristic = 0;
      for (unsigned i = 0; lnode != root && i < n && rec[i].score >= qsearchscore_thresh_heuristic; ++i) {
        const uint32_t cnode = rec[i].node;
        const uint32_t qsearchmatch = rec[i].identities;
        const double qsearchpid = static_cast<double>(qsearchmatch) / qrlength;
        const double qpid = static_cast<double>(querysimilarity[i]) / qrlength;
        const float qsearchscore = rec[i].score;
        const double qpid_thresh = std::max(qpid_thresh_guarantee, qpid_thresh_heuristic);
        if (qpid >= qpid_thresh) {
          float dist;
          if (i == index_anchor) dist = .0f;
          else if (querydistance[i] == .0f) dist = querydistance[index_anchor];
          else {
            const Aln a = next_alignment(i, index_anchor);
            dist = a.dist;
            ++pass_1_counter;
            logsink << "    +ALN " << i << " <=> " << index_anchor << "\tdist=" << dist << "; sim=" << a.sim << "; qsearchscore=" << qsearchscore
                    << "; qsearchmatch=" << qsearchmatch << "; qsearchpid=" << qsearchpid << "; qpid=" << qpid
                    << "; qsearchscore_cut=" << qsearchscore_thresh_heuristic << "; qpid_cutg=" << qpid_thresh_guarantee
                    << "; qpid_cut_h=" << qpid_thresh_heuristic << std::endl;
            print_alignment(i, index_anchor);
          }
          bf.push_back(std::make_pair(dist, cnode));
          if (dist == .0f) qgroup.erase(i);
          else if (dist <= qdist) {
            lnode = T.lca(lnode, cnode);
            if (dist > ldist) ldist = dist;
            logsink << "      current lower node: (" << dist << ") " << name(lnode) << " (+ " << name(cnode) << " at "
                    << static_cast<int>(T.depth[T.lca(cnode, rnode)]) << " )" << std::endl;
          } else {
            if (dist < udist) {
              udist = dist;
              if (qsearchpid > qpid_upper) {
                qpid_upper = qsearchpid;
                qpid_thresh_guarantee = qsearchpid * 2. - 1.;
                qpid_thresh_heuristic = qsearchpid * ctx.exclude_factor;
              }
              if (!qsearchscore_thresh_heuristic) qsearchscore_thresh_heuristic = rec[i].score * ctx.exclude_factor;
            }
            outgroup_tmp.push_back(std::make_pair(i, (int)dist));
          }
        }
      }
      // BandFactor::getFactor (hh:270-323)
      float bandfactor = 1.f;
      {
        std::sort(bf.begin() + 1, bf.end(), [](const std::pair<float, uint32_t>& x, const std::pair<float, uint32_t>& y) { return x.first < y.first; });
        const uint32_t anchor = bf[0].second;
        unsigned last_rank = T.depth[anchor];
        std::vector<float> worst(256, 0.f);
        std::vector<char> have(256, 0);
        worst[last_rank] = bf[0].first; have[last_rank] = 1;
        for (size_t a = 1; a < bf.size(); ++a) {
          const float sc = bf[a].first;
          const unsigned rank = T.depth[T.lca(bf[a].second, anchor)];
          if (rank == last_rank) continue;
          if (rank < last_rank) { worst[rank] = sc; have[rank] = 1; last_rank = rank; continue; }
          for (int r = (int)rank - 1; r >= 0; --r)
            if (have[r] && worst[r]) { const float q = sc / worst[r]; if (q > bandfactor) bandfactor = q; }
        }
        if (bandfactor > FLT_MAX) bandfactor = FLT_MAX;
        bandfactor = sqrtf(bandfactor);
      }
      bandfactor_max = std::max(bandfactor_max, bandfactor);
      const float qdist_ex = qdist * bandfactor;
      float min_upper_dist = (float)INT_MAX;
      logsink << std::endl << "    EXT\tquerydist = " << qdist << "; threshold = " << qdist_ex << "; bandfactor = " << bandfactor << std::endl;
      for (auto it = outgroup_tmp.begin(); it != outgroup_tmp.end();) {
        const float dist = it->second;
        if (dist > qdist_ex) {
          if (dist > min_upper_dist) it = outgroup_tmp.erase(it);
          else { if (dist < min_upper_dist) min_upper_dist = dist; ++it; }
        } else {
          if (min_upper_dist > qdist_ex) min_upper_dist = dist;
          else min_upper_dist = std::max(min_upper_dist, dist);
          ++it;
        }
      }
      if (min_upper_dist != FLT_MAX) { unode = lnode; have_unode = true; }
      for (auto it = outgroup_tmp.begin(); it != outgroup_tmp.end(); ++it) {
        const unsigned i = it->first;
        const float dist = it->second;
        const uint32_t cnode = rec[i].node;
        if (dist > min_upper_dist) continue;
        unode = T.lca(cnode, unode);
        logsink << "      current upper node: (" << dist << ") " << name(unode) << " (+ " << name(cnode) << " at "
                << static_cast<int>(T.depth[T.lca(cnode, rnode)]) << " )" << std::endl;
        const unsigned lca_root_dist = T.depth[T.lca(cnode, rtax)];
        if (lca_root_dist > lca_root_dist_min) continue;
        else if (lca_root_dist < lca_root_dist_min) { lca_root_dist_min = lca_root_dist; outgroup.clear(); }
        outgroup.insert(i);
      }
      float ival = 0.f;
      if (!have_unode) { unode = root; udist = -1; ival = 1.f; }
      else if (unode != lnode && ldist < qdist) ival = (qdist - ldist) / (udist - ldist);
      logsink << std::endl << "    SCORE\tldist = " << ldist << "; udist = " << udist << "; querydist = " << qdist << "; querydist_ex = " << qdist_ex
              << "; ival = " << ival << std::endl << std::endl;
      ival_global = std::max(ival, ival_global);
      unode_global = T.lca(unode_global, unode);
      lnode_global = T.lca(lnode_global, lnode);
    } while (!qgroup.empty() && lnode_global != root);
    logsink << "    NUMALN\t" << pass_1_counter << '\t' << pass_1_counter_naive - pass_1_counter << std::endl;
    logsink << "    NUMOUTGRP\t" << outgroup.size() << std::endl;
  }
  logsink << "    RANGE\t" << name(rtax) << '\t' << name(lnode_global) << '\t' << name(unode_global) << std::endl << std::endl;
  {   // pass 2, hh:737-822
    logsink << "  PASS\t2" << std::endl;
    while (!outgroup.empty()) {
      const unsigned index_anchor = *outgroup.begin();
      outgroup.erase(outgroup.begin());
      if (unode_global == lca_allnodes) {
        if (querydistance[index_anchor] == FLT_MAX) pass_2_counter_naive += n;
        else pass_2_counter_naive += n - 1;
        continue;
      }
      const double qpid_anchor = static_cast<double>(querysimilarity[index_anchor]) / qrlength;
      const double qpid_thresh_guarantee = qpid_anchor * 2. - 1.;
      const double qpid_thresh_heuristic = qpid_anchor * ctx.exclude_factor;
      const double qpid_thresh = std::max(qpid_thresh_guarantee, qpid_thresh_heuristic);
      const float qsearchscore_thresh_heuristic = rec[index_anchor].score * ctx.exclude_factor;
      ++pass_2_counter_naive;
      for (unsigned i = 0; i < n && rec[i].score >= qsearchscore_thresh_heuristic; ++i) {
        const double qpid = static_cast<double>(querysimilarity[i]) / qrlength;
        if (qpid >= qpid_thresh) {
          const uint32_t cnode = rec[i].node;
          const float qsearchscore = rec[i].score;
          const uint32_t qsearchmatch = rec[i].identities;
          float dist;
          if (i == index_anchor) dist = .0f;
          else {
            ++pass_2_counter_naive;
            if (T.is_parent_of(unode_global, cnode) || cnode == unode_global) continue;
            const Aln a = next_alignment(i, index_anchor);
            dist = a.dist;
            logsink << "    +ALN " << i << " <=> " << index_anchor << "\tdist=" << dist << "; sim=" << a.sim << "; qsearchscore=" << qsearchscore
                    << "; qsearchmatch=" << qsearchmatch << "; qpid=" << qpid << std::endl;
            print_alignment(i, index_anchor);
            ++pass_2_counter;
            querydistance[i] = dist;
          }
          if (dist == .0f) outgroup.erase(i);
          else {
            float qdist_ex;
            if (querydistance[index_anchor] == FLT_MAX) {
              const Aln a = next_alignment(index_anchor, TRPA_TRACE_QUERY);
              const float dist2 = a.dist;
              const float sim2 = std::max(a.sim, querysimilarity[index_anchor]);
              const double qpid2 = static_cast<double>(sim2) / qrlength;
              logsink << "    +ALN query <=> " << index_anchor << "\tdist=" << dist2 << "; sim=" << sim2 << "; qsearchscore=" << rec[index_anchor].score
                      << "; qsearchmatch=" << qsearchmatch << "; qpid=" << qpid2 << std::endl;
              print_alignment(index_anchor, TRPA_TRACE_QUERY);
              querydistance[index_anchor] = dist2;
              querysimilarity[index_anchor] = sim2;
              qdist_ex = dist2 * bandfactor_max;
              logsink << "      query: (" << qdist_ex << ") unknown" << std::endl;
              ++pass_2_counter;
            } else qdist_ex = querydistance[index_anchor] * bandfactor_max;
            if (dist <= qdist_ex) {
              const uint32_t rnode = rec[index_anchor].node;
              unode_global = T.lca(unode_global, cnode);
              logsink << "      current upper node: (" << dist << ") " << name(unode_global) << " (+ " << name(cnode) << " at "
                      << static_cast<int>(T.depth[T.lca(cnode, rnode)]) << " )" << std::endl;
            }
          }
        }
      }
      logsink << std::endl;
    }
    logsink << "    NUMALN\t" << pass_2_counter << '\t' << pass_2_counter_naive - pass_2_counter << std::endl;
  }
  if (unode_global == lnode_global) ival_global = 1.f;
  logsink << "    RANGE\t" << name(rtax) << '\t' << name(lnode_global) << '\t' << name(unode_global) << std::endl << std::endl;
  const unsigned gcounter = pass_0_counter + pass_1_counter + pass_2_counter;
  const float normalised_rt = (float)gcounter / (float)n;
  // the three stopwatch columns (CPU time of init / sequence retrieval / processing of THIS record on a host
  // thread, hh:345, 488-490) have no counterpart when a whole batch is placed at once: written as 0
  logsink << "STATS\t" << qrseqname << '\t' << n << '\t' << pass_0_counter << '\t' << pass_1_counter << '\t' << pass_2_counter << '\t' << gcounter
          << '\t' << 0 << '\t' << 0 << '\t' << 0 << '\t' << normalised_rt << std::endl << std::endl;
  // a log writer, not a second predictor: it must have arrived where the GPU did
  if (tpos != n_trace) fail("trace not fully consumed");
  if (res.kind != TRPA_KIND_PLACED || res.lower_node != lnode_global || res.upper_node != unode_global || res.rtax_node != rtax ||
      res.support != anchors_support || res.n_pass0 != pass_0_counter || res.n_pass1 != pass_1_counter || res.n_pass2 != pass_2_counter ||
      !(res.ival == ival_global))
    fail("final record");
}

}  // namespace taxator_b200
