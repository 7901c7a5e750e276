// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's RPA hot path over flat arrays.
// Nothing in the product path (taxator-tk_b200/) may include, link or call this file; it is the
// checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
//
// Parity status: PINNED against the real reference built in this container
//   - kernel level: oracle/_ref/libseqan_ref.so (unmodified vendored SeqAn) -- tests/test_oracle.py
//   - path level:   oracle/_ref/taxator (unmodified reference sources + Boost shims) GFF3 --
//                   tests/golden/ + tests/test_oracle_golden.py
// (upstream ships no test for this path).
//
// What each function follows (all paths relative to /root/reference/core):
//   orc_char2dna5 / orc_char2aa   includes-external/seqan/basic/alphabet_residue_tabs.h (char -> ordinal)
//   orc_edit_distance*            src/taxonpredictionmodelsequence.hh:133-171 +
//                                 includes-external/seqan/align/global_alignment_myers_impl.h:62-197
//   orc_protein_align             src/taxonpredictionmodelsequence.hh:173-242 +
//                                 includes-external/seqan/align/dp_formula_linear.h:62-105,
//                                 dp_formula.h:152-163, dp_traceback_impl.h:379-418
//   fetch_segment                 src/taxonpredictionmodelsequence.hh:856-880,
//                                 src/sequencestorage.hh:341-369,430-457, src/faidx.h:315-350
//   lca / is_parent_of            src/taxonomyinterface.cpp:52-77
//   band_factor                   src/taxonpredictionmodelsequence.hh:259-323
//   orc_predict_segment           src/taxonpredictionmodelsequence.hh:341-838,
//                                 src/alignmentsfilter.hh:171-190, src/alignmentrecord.hh:89-93
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <set>
#include <vector>

#include "blosum62_table.h"

namespace {

typedef std::vector<uint8_t> Seq;

// ---------------------------------------------------------------------------------- alphabets
inline int dna5_of(int c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;  // everything else is N
  }
}
inline int aa_of(int c) {
  static const char order[] = "ABCDEFGHIJKLMNOPQRSTUVWYZX*";
  if (c >= 'a' && c <= 'z') c -= 32;
  for (int i = 0; i < 27; ++i)
    if (order[i] == c) return i;
  return 25;  // X
}

// ------------------------------------------------------------------------------ edit distance
// Textbook Levenshtein, two rows.  N (4) equals N.
int edit_distance_dp(const uint8_t* a, int la, const uint8_t* b, int lb) {
  std::vector<int> prev(lb + 1), cur(lb + 1);
  for (int j = 0; j <= lb; ++j) prev[j] = j;
  for (int i = 1; i <= la; ++i) {
    cur[0] = i;
    for (int j = 1; j <= lb; ++j) {
      int sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
      int del = prev[j] + 1;
      int ins = cur[j - 1] + 1;
      cur[j] = std::min(sub, std::min(del, ins));
    }
    prev.swap(cur);
  }
  return prev[lb];
}

// Bit-vector formulation with 64-bit blocks (same integer, used where the DP would be slow).
// Pattern = shorter sequence, as in the reference (hh:142-147, myers_impl.h:70-71).
int edit_distance_bitvector(const uint8_t* a, int la, const uint8_t* b, int lb) {
  const uint8_t* pat = a; int m = la;
  const uint8_t* txt = b; int n = lb;
  if (la > lb) { pat = b; m = lb; txt = a; n = la; }
  if (m == 0) return n;
  const int nb = (m + 63) / 64;
  std::vector<uint64_t> peq(5 * (size_t)nb, 0), vp(nb, ~0ull), vn(nb, 0);
  for (int j = 0; j < m; ++j) peq[(size_t)pat[j] * nb + j / 64] |= 1ull << (j % 64);
  int score = m;
  const uint64_t last = 1ull << ((m - 1) % 64);
  for (int pos = 0; pos < n; ++pos) {
    const uint64_t* eqrow = &peq[(size_t)txt[pos] * nb];
    uint64_t add_carry = 0, hp_carry = 1, hn_carry = 0;  // first row grows by +1 per column: global
    uint64_t hp = 0, hn = 0;
    for (int k = 0; k < nb; ++k) {
      uint64_t x = eqrow[k] | vn[k];
      uint64_t t = x & vp[k];
      uint64_t s1 = vp[k] + t;
      uint64_t c1 = s1 < t;
      uint64_t s2 = s1 + add_carry;
      uint64_t c2 = s2 < s1;
      add_carry = c1 | c2;
      uint64_t d0 = (s2 ^ vp[k]) | x;
      hn = vp[k] & d0;
      hp = vn[k] | ~(vp[k] | d0);
      uint64_t xs = (hp << 1) | hp_carry;
      uint64_t hs = (hn << 1) | hn_carry;
      hp_carry = hp >> 63;
      hn_carry = hn >> 63;
      vn[k] = xs & d0;
      vp[k] = hs | ~(xs | d0);
    }
    if (hp & last) ++score;
    else if (hn & last) --score;
  }
  return score;
}

// ------------------------------------------------------------------------- protein alignment
// Needleman-Wunsch, BLOSUM62, linear gap -1 per column (Blosum62() default ctor: open=extend=-1).
// SeqAn's linear recurrence takes max in the order diagonal, then vertical, then horizontal, and
// keeps the earlier candidate on ties; its single-trace walker follows exactly that stored
// direction.  Carrying (#diagonal steps) forward with the same priority therefore reproduces the
// traced alignment length |A|+|B|-#diag and, with (#matching diagonals), the match count.
// H = A (row 0), V = B (row 1) as at hh:196-197.
struct ProtAln { int mutual, self, len, match, mismatch, gap; };

ProtAln protein_align(const uint8_t* A, int la, const uint8_t* B, int lb) {
  // cell (i over V=B rows, j over H=A columns)
  struct Cell { int s, nd, nm; };
  std::vector<Cell> prev(la + 1), cur(la + 1);
  for (int j = 0; j <= la; ++j) prev[j] = Cell{-j, 0, 0};
  for (int i = 1; i <= lb; ++i) {
    cur[0] = Cell{-i, 0, 0};
    for (int j = 1; j <= la; ++j) {
      const int sub = ORC_BLOSUM62[A[j - 1]][B[i - 1]];
      Cell best{prev[j - 1].s + sub, prev[j - 1].nd + 1, prev[j - 1].nm + (A[j - 1] == B[i - 1])};
      const int up = prev[j].s - 1;       // vertical: consumes B[i-1] against a gap
      if (up > best.s) best = Cell{up, prev[j].nd, prev[j].nm};
      const int left = cur[j - 1].s - 1;  // horizontal: consumes A[j-1] against a gap
      if (left > best.s) best = Cell{left, cur[j - 1].nd, cur[j - 1].nm};
      cur[j] = best;
    }
    prev.swap(cur);
  }
  ProtAln r;
  r.mutual = prev[la].s;
  int self = 0;  // NW(X,X) with this matrix is the diagonal sum (cross-checked against SeqAn in tests)
  for (int j = 0; j < la; ++j) self += ORC_BLOSUM62[A[j]][A[j]];
  for (int i = 0; i < lb; ++i) self += ORC_BLOSUM62[B[i]][B[i]];
  r.self = self;
  r.len = la + lb - prev[la].nd;
  r.match = prev[la].nm;
  r.mismatch = prev[la].nd - prev[la].nm;
  r.gap = r.len - prev[la].nd;
  return r;
}

// ------------------------------------------------------------------------------------- stores
struct Store {
  const uint8_t* codes;   // ordinals, all sequences concatenated
  const uint64_t* off;
  const uint32_t* len;
  uint32_t n;
};

// 1-based inclusive [start, stop]; stop clipped to the sequence (sequencestorage.hh:353, :112);
// a start past the end yields an empty string (faidx.h:325-331).
Seq store_get(const Store& st, uint32_t id, uint64_t start, uint64_t stop) {
  const uint64_t L = st.len[id];
  stop = std::min<uint64_t>(stop, L);
  uint64_t b = std::min<uint64_t>(start - 1, L);
  uint64_t e = std::min<uint64_t>(std::max<uint64_t>(b, stop), L);
  const uint8_t* p = st.codes + st.off[id];
  return Seq(p + b, p + e);
}

// hh:870-880 (+ :856-860: the protein build also runs through this routine; the amino-acid
// store's "reverse complement" returns the forward string, sequencestorage.hh:452-457)
Seq fetch_segment(const Store& st, bool protein, uint32_t id, uint32_t start, uint32_t stop,
                  uint32_t left_ext, uint32_t right_ext) {
  if (start <= stop) {
    uint32_t ns = left_ext < start ? start - left_ext : 1;
    uint32_t ne = stop + right_ext;
    return store_get(st, id, ns, ne);
  }
  uint32_t ns = right_ext < stop ? stop - right_ext : 1;
  uint32_t ne = start + left_ext;
  Seq s = store_get(st, id, ns, ne);
  if (!protein) {
    std::reverse(s.begin(), s.end());
    for (auto& c : s) c = c < 4 ? 3 - c : 4;
  }
  return s;
}

// ----------------------------------------------------------------------------------- taxonomy
struct Tax {
  const uint32_t* parent;
  const uint32_t* left;
  const uint32_t* right;
  const uint8_t* depth;
  uint32_t n, root;
};

uint32_t lca(const Tax& t, uint32_t A, uint32_t B) {
  // literal restatement incl. the min(A.left, B.right) of taxonomyinterface.cpp:68
  uint32_t left_min = std::min(t.left[A], t.right[B]);
  uint32_t right_max = std::max(t.right[A], t.right[B]);
  uint32_t x = A;
  while (t.left[x] > left_min || t.right[x] < right_max) x = t.parent[x];
  return x;
}
bool is_parent_of(const Tax& t, uint32_t A, uint32_t B) {
  return t.right[A] > t.left[B] && t.left[A] < t.left[B];
}

// --------------------------------------------------------------------------------- BandFactor
struct BandFactor {
  const Tax& tax;
  std::vector<std::pair<float, uint32_t>> data;
  explicit BandFactor(const Tax& t) : tax(t) {}
  void add(float d, uint32_t node) { data.emplace_back(d, node); }
  float factor() {
    // same library sort on the same input order as the reference (unstable above 16 elements)
    std::sort(data.begin() + 1, data.end(),
              [](const std::pair<float, uint32_t>& a, const std::pair<float, uint32_t>& b) { return a.first < b.first; });
    float bf = 1.f;
    const uint32_t anchor = data[0].second;
    uint8_t last_rank = tax.depth[anchor];
    std::map<uint8_t, float> worst;
    worst[last_rank] = data[0].first;
    for (size_t i = 1; i < data.size(); ++i) {
      const float sc = data[i].first;
      const uint8_t rank = tax.depth[lca(tax, data[i].second, anchor)];
      if (rank == last_rank) continue;
      if (rank < last_rank) { worst[rank] = sc; last_rank = rank; continue; }
      uint8_t r = rank - 1;
      do {
        auto it = worst.find(r);
        if (it != worst.end() && it->second) bf = std::max(bf, sc / it->second);
      } while (r--);
    }
    bf = std::min(bf, FLT_MAX);
    return (float)sqrt(bf);  // ::sqrt(double) on a float argument, rounded back (hh:276)
  }
};

}  // namespace

extern "C" {

struct OrcCand {
  uint32_t ref_seq, rstart, rstop, qstart, qstop;
  float score;
  uint32_t identities, alnlen, node;
};

struct OrcResult {
  uint32_t qrstart, qrstop;
  uint32_t lower, upper, rtax, support;
  float ival, signal;
  uint32_t p0, p1, p2;
  uint32_t kind;  // 0: no record, 1: single record, 2: identical-hit shortcut, 3: three-pass placement
  uint64_t cells; // sum |A|*|B| over performed alignments
};

struct OrcPairLog {  // optional trace of performed alignments
  uint32_t pass, i, j;  // j == 0xffffffff: query
  uint32_t la, lb;
  float dist, sim;
};

int orc_char2dna5(int c) { return dna5_of(c); }
int orc_char2aa(int c) { return aa_of(c); }

int orc_edit_distance_dp(const uint8_t* a, int la, const uint8_t* b, int lb) { return edit_distance_dp(a, la, b, lb); }
int orc_edit_distance(const uint8_t* a, int la, const uint8_t* b, int lb) { return edit_distance_bitvector(a, la, b, lb); }

void orc_protein_align(const uint8_t* a, int la, const uint8_t* b, int lb, int* out6) {
  ProtAln r = protein_align(a, la, b, lb);
  out6[0] = r.mutual; out6[1] = r.self; out6[2] = r.len; out6[3] = r.match; out6[4] = r.mismatch; out6[5] = r.gap;
}

// returns length; out must hold (stop-start+1+left_ext+right_ext) bytes
uint32_t orc_fetch_segment(const uint8_t* codes, const uint64_t* off, const uint32_t* len, uint32_t nseq, int protein,
                           uint32_t id, uint32_t start, uint32_t stop, uint32_t left_ext, uint32_t right_ext,
                           uint8_t* out) {
  Store st{codes, off, len, nseq};
  Seq s = fetch_segment(st, protein != 0, id, start, stop, left_ext, right_ext);
  if (!s.empty()) memcpy(out, s.data(), s.size());
  return (uint32_t)s.size();
}

uint32_t orc_lca(const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth, uint32_t n,
                 uint32_t root, uint32_t a, uint32_t b) {
  Tax t{parent, left, right, depth, n, root};
  return lca(t, a, b);
}

struct OrcAln { float distance, similarity; };

static OrcAln align_pair(bool protein, const Seq& A, const Seq& B, uint64_t& cells) {
  cells += (uint64_t)A.size() * (uint64_t)B.size();
  OrcAln r;
  if (!protein) {  // hh:133-171
    int d = edit_distance_bitvector(A.data(), (int)A.size(), B.data(), (int)B.size());
    int llong = (int)std::max(A.size(), B.size()), lshort = (int)std::min(A.size(), B.size());
    int lendiff = llong - lshort;
    int mismatch = d - lendiff;
    int match = lshort - mismatch;
    r.distance = (float)d;
    r.similarity = (float)match;
  } else {  // hh:173-242
    ProtAln p = protein_align(A.data(), (int)A.size(), B.data(), (int)B.size());
    unsigned int len = (unsigned int)p.len;
    float norm = len / static_cast<float>(p.self);
    r.distance = (p.self - 2 * p.mutual) * norm;
    r.similarity = (2 * p.mutual) * norm;
  }
  return r;
}

// One query segment.  cands are the unmasked records in record-set order; they are stably sorted
// here by descending (score, identities) as SortFilter does.  Returns 0.
int orc_predict_segment(
    // taxonomy
    const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth, uint32_t n_nodes, uint32_t root,
    // stores (ordinals)
    const uint8_t* q_codes, const uint64_t* q_off, const uint32_t* q_len, uint32_t q_n,
    const uint8_t* r_codes, const uint64_t* r_off, const uint32_t* r_len, uint32_t r_n,
    int protein, float exclude_factor, float toppercent,
    uint32_t query_seq, const OrcCand* cands_in, uint32_t n,
    OrcResult* res, OrcPairLog* plog, uint32_t plog_cap, uint32_t* plog_n) {
  Tax tax{parent, left, right, depth, n_nodes, root};
  Store qs{q_codes, q_off, q_len, q_n}, rs{r_codes, r_off, r_len, r_n};
  const bool prot = protein != 0;
  const float reeval_bandwidth_factor = 1. - toppercent;  // hh:334
  memset(res, 0, sizeof(*res));
  uint32_t nlog = 0;
  auto log_pair = [&](uint32_t pass, uint32_t i, uint32_t j, const Seq& A, const Seq& B, OrcAln a) {
    if (plog && nlog < plog_cap) plog[nlog] = OrcPairLog{pass, i, j, (uint32_t)A.size(), (uint32_t)B.size(), a.distance, a.similarity};
    ++nlog;
  };
  struct Fin { uint32_t* p; uint32_t& v; ~Fin() { if (p) *p = v; } } fin{plog_n, nlog};

  if (n == 0) {  // hh:359-368
    res->kind = 0; res->lower = res->upper = res->rtax = root; res->support = 0; res->ival = -2.f;
    return 0;
  }
  if (n == 1) {  // hh:371-388
    const OrcCand& r = cands_in[0];
    res->kind = 1; res->qrstart = r.qstart; res->qrstop = r.qstop; res->ival = 1.f;
    res->lower = r.node; res->upper = root; res->support = r.identities; res->rtax = r.node;
    return 0;
  }
  uint32_t qrstart = cands_in[0].qstart, qrstop = cands_in[0].qstop;
  for (uint32_t i = 1; i < n; ++i) {
    qrstart = std::min(cands_in[i].qstart, qrstart);
    qrstop = std::max(cands_in[i].qstop, qrstop);
  }
  const uint32_t qrlength = qrstop - qrstart + 1;
  std::vector<OrcCand> rec(cands_in, cands_in + n);
  std::stable_sort(rec.begin(), rec.end(), [](const OrcCand& a, const OrcCand& b) {
    // "b < a" with operator< on (score, identities), alignmentrecord.hh:89-93
    if (b.score < a.score) return true;
    if (b.score > a.score) return false;
    return b.identities < a.identities;
  });
  const Seq qrseq = store_get(qs, query_seq, qrstart, qrstop);  // hh:415, sequencestorage.hh:105-120
  const float qmax_searchscore = rec[0].score;
  res->qrstart = qrstart; res->qrstop = qrstop;

  if (rec[0].alnlen == qrlength && rec[0].identities == qrlength) {  // hh:431-472
    float best = rec[0].score;
    uint32_t lnode = rec[0].node, unode = UINT32_MAX, i = 1;
    while (true) {
      if (i == n) { unode = root; break; }
      float s = rec[i].score;
      if (s == best) lnode = lca(tax, lnode, rec[i].node);
      else {
        float us = s;
        unode = lnode;
        do { unode = lca(tax, unode, rec[i].node); } while (++i < n && rec[i].score == us);
        break;
      }
      ++i;
    }
    res->kind = 2; res->ival = 0.f; res->lower = lnode; res->upper = unode; res->support = qrlength; res->rtax = lnode;
    return 0;
  }

  std::vector<Seq> seg(n);
  std::vector<char> have(n, 0);
  auto need = [&](uint32_t i) {
    if (!have[i] || seg[i].empty()) {
      seg[i] = fetch_segment(rs, prot, rec[i].ref_seq, rec[i].rstart, rec[i].rstop, rec[i].qstart - qrstart, qrstop - rec[i].qstop);
      have[i] = 1;
    }
  };
  std::vector<float> qd(n, FLT_MAX), qsim(n, 0.f);
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  uint64_t cells = 0;
  std::set<uint32_t> qgroup;
  uint32_t anchors_support = 0;
  uint32_t rtax, lca_all = rec[0].node;

  {  // pass 0, hh:497-566
    const float thr = reeval_bandwidth_factor * qmax_searchscore;
    uint32_t ibest = 0;
    for (uint32_t i = 0; i < n; ++i) {
      float dist, sim;
      const float sc = rec[i].score;
      if (rec[i].alnlen == qrlength && rec[i].identities == qrlength) {
        qgroup.insert(i); dist = 0; sim = rec[i].identities;
      } else if (rec[i].score >= thr) {
        qgroup.insert(i);
        need(i);
        OrcAln a = align_pair(prot, seg[i], qrseq, cells);
        log_pair(0, i, UINT32_MAX, seg[i], qrseq, a);
        dist = a.distance; ++c0;
        sim = std::max(a.similarity, static_cast<float>(rec[i].identities));
      } else { dist = FLT_MAX; sim = rec[i].identities; }
      qd[i] = dist; qsim[i] = sim;
      if (dist < qd[ibest]) ibest = i;
      else if (dist == qd[ibest]) {
        if (sim > qsim[ibest]) ibest = i;
        else if (sim == qsim[ibest] && sc > rec[ibest].score) ibest = i;
      }
      anchors_support = std::max(anchors_support, static_cast<uint32_t>(sim));
      lca_all = lca(tax, lca_all, rec[i].node);
    }
    rtax = rec[ibest].node;
    for (auto it = qgroup.begin(); it != qgroup.end();) {
      if (qd[*it] != qd[ibest] || qsim[*it] != qsim[ibest] || rec[*it].score != rec[ibest].score) qgroup.erase(it++);
      else { rtax = lca(tax, rtax, rec[*it].node); ++it; }
    }
  }

  float ival_global = 0.f;
  uint32_t lnode_g = rtax, unode_g = rtax;
  std::set<uint32_t> outgroup;
  float bandfactor_max = 1.f;

  {  // pass 1, hh:576-733
    uint8_t lca_root_dist_min = 255;
    do {
      BandFactor bf(tax);
      const uint32_t anchor = *qgroup.begin();
      qgroup.erase(qgroup.begin());
      const float qdist = qd[anchor];
      const uint32_t rnode = rec[anchor].node;
      bf.add(0, rnode);
      uint32_t lnode = rtax;
      bool have_unode = false; uint32_t unode = 0;
      float ldist = 0, udist = FLT_MAX;
      std::list<std::pair<uint32_t, int>> og_tmp;
      double qpid_upper = 0., thr_guarantee = 0., thr_heur = 0.;
      int score_thr_heur = 0.;
      for (uint32_t i = 0; lnode != root && i < n && rec[i].score >= score_thr_heur; ++i) {
        const uint32_t cnode = rec[i].node;
        const double qsearchpid = static_cast<double>(rec[i].identities) / qrlength;
        const double qpid = static_cast<double>(qsim[i]) / qrlength;
        const double qpid_thresh = std::max(thr_guarantee, thr_heur);
        if (qpid >= qpid_thresh) {
          float dist;
          if (i == anchor) dist = .0;
          else if (qd[i] == .0) dist = qd[anchor];
          else {
            need(anchor); need(i);
            OrcAln a = align_pair(prot, seg[i], seg[anchor], cells);
            log_pair(1, i, anchor, seg[i], seg[anchor], a);
            dist = a.distance; ++c1;
          }
          bf.add(dist, cnode);
          if (dist == .0) qgroup.erase(i);
          else if (dist <= qdist) {
            lnode = lca(tax, lnode, cnode);
            if (dist > ldist) ldist = dist;
          } else {
            if (dist < udist) {
              udist = dist;
              if (qsearchpid > qpid_upper) {
                qpid_upper = qsearchpid;
                thr_guarantee = qsearchpid * 2. - 1.;
                thr_heur = qsearchpid * exclude_factor;
              }
              if (!score_thr_heur) score_thr_heur = rec[i].score * exclude_factor;
            }
            og_tmp.emplace_back(i, (int)dist);  // tuple<uint,int>: distance truncated
          }
        }
      }
      const float bandfactor = bf.factor();
      bandfactor_max = std::max(bandfactor_max, bandfactor);
      const float qdist_ex = qdist * bandfactor;
      float min_upper = std::numeric_limits<int>::max();
      for (auto it = og_tmp.begin(); it != og_tmp.end();) {
        float dist = it->second;
        if (dist > qdist_ex) {
          if (dist > min_upper) it = og_tmp.erase(it);
          else { if (dist < min_upper) min_upper = dist; ++it; }
        } else {
          if (min_upper > qdist_ex) min_upper = dist;
          else min_upper = std::max(min_upper, dist);
          ++it;
        }
      }
      if (min_upper != FLT_MAX) { unode = lnode; have_unode = true; }  // always taken (int max vs float max)
      for (auto& e : og_tmp) {
        float dist = e.second;
        const uint32_t cnode = rec[e.first].node;
        if (dist > min_upper) continue;
        unode = lca(tax, cnode, unode);
        const uint8_t lrd = depth[lca(tax, cnode, rtax)];
        if (lrd > lca_root_dist_min) continue;
        else if (lrd < lca_root_dist_min) { lca_root_dist_min = lrd; outgroup.clear(); }
        outgroup.insert(e.first);
      }
      float ival = 0.;
      if (!have_unode) { unode = root; udist = -1; ival = 1.; }
      else if (unode != lnode && ldist < qdist) ival = (qdist - ldist) / (udist - ldist);
      ival_global = std::max(ival, ival_global);
      unode_g = lca(tax, unode_g, unode);
      lnode_g = lca(tax, lnode_g, lnode);
    } while (!qgroup.empty() && lnode_g != root);
  }

  {  // pass 2, hh:737-822
    while (!outgroup.empty()) {
      const uint32_t anchor = *outgroup.begin();
      outgroup.erase(outgroup.begin());
      if (unode_g == lca_all) continue;
      const double qpid_anchor = static_cast<double>(qsim[anchor]) / qrlength;
      const double thr_guarantee = qpid_anchor * 2. - 1.;
      const double thr_heur = qpid_anchor * exclude_factor;
      const double qpid_thresh = std::max(thr_guarantee, thr_heur);
      const float score_thr = rec[anchor].score * exclude_factor;
      for (uint32_t i = 0; i < n && rec[i].score >= score_thr; ++i) {
        const double qpid = static_cast<double>(qsim[i]) / qrlength;
        if (qpid >= qpid_thresh) {
          const uint32_t cnode = rec[i].node;
          float dist;
          if (i == anchor) dist = .0;
          else {
            if (is_parent_of(tax, unode_g, cnode) || cnode == unode_g) continue;
            need(anchor); need(i);
            OrcAln a = align_pair(prot, seg[i], seg[anchor], cells);
            log_pair(2, i, anchor, seg[i], seg[anchor], a);
            dist = a.distance; ++c2;
            qd[i] = dist;
          }
          if (dist == .0) outgroup.erase(i);
          else {
            float qdist_ex;
            if (qd[anchor] == FLT_MAX) {
              need(anchor);
              OrcAln a = align_pair(prot, seg[anchor], qrseq, cells);
              log_pair(2, anchor, UINT32_MAX, seg[anchor], qrseq, a);
              float d2 = a.distance;
              float s2 = std::max(a.similarity, qsim[anchor]);
              qd[anchor] = d2; qsim[anchor] = s2;
              qdist_ex = d2 * bandfactor_max; ++c2;
            } else qdist_ex = qd[anchor] * bandfactor_max;
            if (dist <= qdist_ex) unode_g = lca(tax, unode_g, cnode);
          }
        }
      }
    }
  }

  if (unode_g == lnode_g) ival_global = 1.;
  res->kind = 3; res->ival = ival_global; res->signal = 0.f;
  res->lower = lnode_g; res->upper = unode_g; res->support = anchors_support; res->rtax = rtax;
  res->p0 = c0; res->p1 = c1; res->p2 = c2; res->cells = cells;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The alignment-free placement models (core/src/taxonpredictionmodel.hh:57-259) with their filters
// (core/src/alignmentsfilter.hh:349-386, :493-534, :612-623), restated record by record on a std::list
// with a `filtered` flag like AlignmentRecord::filterOut(); node sets are std::set like the reference's
// std::set<const TaxonNode*> (the LCA of a set does not depend on its order).
// model: 0 dummy, 1 simple-lca, 2 megan-lca (== ic-megan-lca, taxator.cpp:352-357), 3 n-best-lca.
// evalue / unclassified may be null.  The result's kind is 4 (point estimate) or 0 (setUnclassified).
int orc_predict_lca_model(const uint32_t* parent, const uint32_t* left, const uint32_t* right, const uint8_t* depth,
                          uint32_t n_nodes, uint32_t root, uint32_t model, float toppercent, float minscore,
                          double maxevalue_arg, uint32_t minsupport, uint32_t nbest, int ignore_unclassified,
                          const OrcCand* cands, const double* evalue, uint32_t n, const uint8_t* unclassified,
                          OrcResult* res) {
  Tax tax{parent, left, right, depth, n_nodes, root};
  struct Rec { float score; double evalue; uint32_t qstart, qstop, node; bool filtered; };
  std::list<Rec> recordset;
  for (uint32_t i = 0; i < n; ++i)
    recordset.push_back(Rec{cands[i].score, evalue ? evalue[i] : 0.0, cands[i].qstart, cands[i].qstop, cands[i].node, false});
  memset(res, 0, sizeof(*res));
  res->ival = -1.f;   // never set by these models: PredictionRecordBase starts at -1 (predictionrecord.hh:40)
  auto set_unclassified = [&]() {   // taxonpredictionmodel.hh:46-49 after initPredictionRecord (:42-44)
    res->kind = 0; res->lower = res->upper = res->rtax = root; res->support = 0; res->qrstart = 1; res->qrstop = 0;
  };
  auto lca_set = [&](const std::set<uint32_t>& nodes) -> uint32_t {   // taxonomyinterface.hh:61-74
    auto it = nodes.begin();
    uint32_t tmp = *it++;
    while (it != nodes.end()) { tmp = lca(tax, tmp, *it); ++it; }
    return tmp;
  };
  auto lca_simple = [&]() {   // LCASimplePredictionModel::predict, taxonpredictionmodel.hh:73-125
    std::set<uint32_t> refnodes, refnodes_best;
    auto rec_it = recordset.begin();
    while (rec_it != recordset.end() && rec_it->filtered) ++rec_it;   // firstUnmaskedIter
    if (rec_it == recordset.end()) { set_unclassified(); return; }
    uint32_t qrstart = rec_it->qstart, qrstop = rec_it->qstop;
    float maxscore = rec_it->score;
    if (qrstart > qrstop) std::swap(qrstart, qrstop);
    refnodes.insert(rec_it->node);
    ++rec_it;
    for (; rec_it != recordset.end(); ++rec_it) {
      if (!rec_it->filtered) {
        if (rec_it->qstart <= rec_it->qstop) { qrstart = std::min(rec_it->qstart, qrstart); qrstop = std::max(rec_it->qstop, qrstop); }
        else { qrstart = std::min(rec_it->qstop, qrstart); qrstop = std::max(rec_it->qstart, qrstop); }
        if (rec_it->score > maxscore) maxscore = rec_it->score;
        refnodes.insert(rec_it->node);
      }
    }
    for (auto& r : recordset) if (!r.filtered && r.score == maxscore) refnodes_best.insert(r.node);
    const uint32_t node = lca_set(refnodes);
    res->kind = 4; res->qrstart = qrstart; res->qrstop = qrstop;
    res->lower = res->upper = node; res->support = qrstop - qrstart + 1;   // setNodePoint(node): feature width
    res->rtax = refnodes.size() != refnodes_best.size() ? lca_set(refnodes_best) : node;
  };
  if (model == 0) { set_unclassified(); return 0; }
  if (model == 1) { lca_simple(); return 0; }
  if (model == 2) {
    // MinScoreMaxEvalueTopPercentFilter (alignmentsfilter.hh:351: the constructor takes maxevalue as float)
    const float maxevalue = (float)maxevalue_arg;
    float max_bitscore = .0;
    unsigned int support = 0;
    for (auto& r : recordset) {
      if (!r.filtered) {
        if (r.score < minscore || r.evalue > maxevalue) r.filtered = true;
        else if (r.score > max_bitscore) { max_bitscore = r.score; support++; }
      }
    }
    max_bitscore = (1.0 - toppercent) * max_bitscore;
    for (auto& r : recordset) if (r.score < max_bitscore) r.filtered = true;
    if (ignore_unclassified) for (auto& r : recordset) if (unclassified && unclassified[r.node]) r.filtered = true;
    if (support >= minsupport) { lca_simple(); return 0; }
    set_unclassified();
    return 0;
  }
  if (model == 3) {
    // NumBestBitscoreFilter (alignmentsfilter.hh:497-527)
    std::multimap<float, Rec*, std::greater<float>> sorted_bitscores;
    for (auto& r : recordset) if (!r.filtered) sorted_bitscores.insert(std::make_pair(r.score, &r));
    if (!sorted_bitscores.empty()) {
      unsigned int count = nbest;
      auto sb_it = sorted_bitscores.begin();
      float lastvalue = sb_it++->first;
      for (; sb_it != sorted_bitscores.end(); ++sb_it) {
        if (sb_it->first != lastvalue) {
          if (--count <= 0) break;
          lastvalue = sb_it->first;
        }
      }
      for (; sb_it != sorted_bitscores.end(); ++sb_it) sb_it->second->filtered = true;
    }
    lca_simple();
    return 0;
  }
  return -1;
}

}  // extern "C"
