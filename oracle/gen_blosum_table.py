"""Regenerates oracle/blosum62_table.h and taxator-tk_b200/csrc/blosum62_table.h from the real
SeqAn score matrix through oracle/_ref/libseqan_ref.so (build container only)."""
import ctypes, os
here = os.path.dirname(os.path.abspath(__file__))
L = ctypes.CDLL(os.path.join(here, "_ref", "libseqan_ref.so"))
letters = "".join(chr(L.ref_aa2char(i)) for i in range(27))
rows = [[L.ref_blosum62(a, b) for b in range(27)] for a in range(27)]
HDR = """// BLOSUM62 substitution scores in SeqAn's 27-letter AminoAcid ordinal order
//   %s
// (core/includes-external/seqan/basic/alphabet_residue.h:552-556; values as returned by
// seqan::score(Blosum62(), a, b), score/score_matrix_data.h:324-352).  The matrix itself is the
// public NCBI BLOSUM62; this table was dumped through oracle/seqan_harness.cpp by
// oracle/gen_blosum_table.py and is plain data.  Rows/cols padded to 32 for cheap indexing.
#pragma once
static const signed char %s[27][32] = {
"""
def emit(name):
    s = HDR % (" ".join(letters), name)
    for a in range(27):
        s += "  /* %s */ {" % letters[a] + ",".join("%3d" % v for v in rows[a] + [0] * 5) + "},\n"
    return s + "};\n"
open(os.path.join(here, "blosum62_table.h"), "w").write(emit("ORC_BLOSUM62"))
open(os.path.join(here, "..", "taxator-tk_b200", "csrc", "blosum62_table.h"), "w").write(emit("TRPA_BLOSUM62"))
