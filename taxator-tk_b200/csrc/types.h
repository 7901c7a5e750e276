// Plain (host + device) record types shared by the kernels, the state machine and the C-ABI glue.
#pragma once
#include <cstdint>

namespace trpa {

// A staged sequence.  Nucleotide: bit-plane words in HBM -- base j lives at bit (j & 31) of word
// woff + (j >> 5); planes.x = low bit of the 2-bit code (A=0,C=1,G=2,T=3), planes.y = high bit,
// nplane = "is N" (planes are 0 where N); bits past `len` are 0.  Protein: one ordinal byte per
// residue at byte offset woff; `pad` then holds the self score sum BLOSUM62(a_i,a_i).
struct SeqDesc {
  uint32_t woff;
  uint32_t len;
  uint32_t flags;  // bit0: contains N
  uint32_t pad;
};

// PairDesc.pad on entry: distance hint; bit 31 set = the value is a true upper bound (no margin needed)
constexpr uint32_t kHintIsBound = 0x80000000u;

// One pairwise alignment request: getAlignment(A = descs[a], B = descs[b]) -> result slot `out`.
// pad: distance hint on entry (0 = none); after planning bits 0..7 = kernel shape, bits 8..31 = band
// threshold k0.  aux: after planning the wedge request (shapes.h wedge_pack: half width at the last
// pattern row | rows of full width), 0 = the plain band.
struct PairDesc {
  uint32_t a, b, out, pad, aux;
  uint32_t cls;   // after planning: duration class of the pair (log2 of the planner's time estimate); the
                  // pairs of a shape bucket are launched longest class first
};

// One segment-staging request: cut [begin, begin+descs[desc].len) (0-based) out of sequence `seq`
// of store `store` (0 query, 1 reference), reverse-complemented if rev, into descs[desc].woff.
struct StageReq {
  uint32_t desc, store, seq, begin, rev;
};

}  // namespace trpa
