// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Thin C-callable harness around the *unmodified* vendored SeqAn 2.4.0 headers of the
// reference (/root/reference/core/includes-external/seqan).  It exposes exactly the SeqAn
// calls the reference's hot path makes, so that oracle/rpa_oracle.cpp (the restatement) and
// the CUDA kernels can be pinned against the real arithmetic:
//
//   * nucleotide: seqan::globalAlignmentScore(short, long, MyersBitVector())
//       as called at core/src/taxonpredictionmodelsequence.hh:150
//   * protein:    seqan::globalAlignmentScore(X, X, Blosum62(), LinearGaps()) twice and
//                 seqan::globalAlignment(align, Blosum62(), LinearGaps()) + row walk
//       as called at core/src/taxonpredictionmodelsequence.hh:190-227
//   * alphabet conversion char -> Dna5 / AminoAcid ordinal and the Blosum62 table
//       (seqan/basic/alphabet_residue*.h, seqan/score/score_matrix_data.h)
//   * seqan::reverseComplement on Dna5 (core/src/sequencestorage.hh:364-369)
//
// Built by oracle/Makefile into oracle/_ref/libseqan_ref.so (git-ignored).  Only this file is
// ours; the SeqAn sources are compiled where they lie and are not copied.
#include <seqan/basic.h>
#include <seqan/sequence.h>
#include <seqan/align.h>
#include <seqan/score.h>
#include <seqan/modifier.h>

using namespace seqan;

extern "C" {

int ref_char2dna5(int c) { return (int)ordValue(Dna5((char)c)); }
int ref_char2aa(int c) { return (int)ordValue(AminoAcid((char)c)); }
int ref_aa2char(int o) { AminoAcid a; a.value = (unsigned char)o; return (int)(char)a; }

int ref_blosum62(int a, int b) {
  Blosum62 sc;
  AminoAcid x, y;
  x.value = (unsigned char)a;
  y.value = (unsigned char)b;
  return score(sc, x, y);
}
int ref_blosum62_gap_open() { Blosum62 sc; return scoreGapOpen(sc); }
int ref_blosum62_gap_extend() { Blosum62 sc; return scoreGapExtend(sc); }

// hh:133-171 getAlignmentDNA -> positive edit distance
int ref_edit_distance(const char* a, int la, const char* b, int lb) {
  String<Dna5> A, B;
  resize(A, la);
  resize(B, lb);
  for (int i = 0; i < la; ++i) A[i] = Dna5(a[i]);
  for (int i = 0; i < lb; ++i) B[i] = Dna5(b[i]);
  const String<Dna5>* long_seq = &A;
  const String<Dna5>* short_seq = &B;
  if (length(A) < length(B)) { long_seq = &B; short_seq = &A; }
  return -globalAlignmentScore(*short_seq, *long_seq, MyersBitVector());
}

// hh:173-242 getAlignmentProtein -> out = {mutualscore, selfscore, len, match, mismatch, gap}
void ref_protein_align(const char* a, int la, const char* b, int lb, int* out) {
  typedef String<AminoAcid> S;
  typedef Align<S, ArrayGaps> TAlign;
  typedef Row<TAlign>::Type TRow;
  typedef Iterator<TRow>::Type TRowIterator;
  S A, B;
  resize(A, la);
  resize(B, lb);
  for (int i = 0; i < la; ++i) A[i] = AminoAcid(a[i]);
  for (int i = 0; i < lb; ++i) B[i] = AminoAcid(b[i]);
  Blosum62 sc;
  LinearGaps algo;
  int selfscore = globalAlignmentScore(A, A, sc, algo) + globalAlignmentScore(B, B, sc, algo);
  TAlign aln;
  resize(rows(aln), 2);
  assignSource(row(aln, 0), A);
  assignSource(row(aln, 1), B);
  int mutual = globalAlignment(aln, sc, algo);
  TRow& r1 = row(aln, 0);
  TRow& r2 = row(aln, 1);
  int gap = 0, match = 0, mismatch = 0;
  TRowIterator it1 = begin(r1), e1 = end(r1), it2 = begin(r2);
  for (; it1 != e1; ++it1, ++it2) {
    if (isGap(it1) || isGap(it2)) gap++;
    else if (*it1 == *it2) match++;
    else mismatch++;
  }
  out[0] = mutual;
  out[1] = selfscore;
  out[2] = gap + match + mismatch;
  out[3] = match;
  out[4] = mismatch;
  out[5] = gap;
}

// sequencestorage.hh:364-369; in/out are ASCII, out gets the Dna5->char rendering
void ref_revcomp_dna5(const char* a, int la, char* out) {
  String<Dna5> A;
  resize(A, la);
  for (int i = 0; i < la; ++i) A[i] = Dna5(a[i]);
  reverseComplement(A);
  for (int i = 0; i < la; ++i) out[i] = (char)A[i];
}

}  // extern "C"
