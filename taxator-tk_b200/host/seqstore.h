// Host side of the sequence stores: reads FASTA (+ .fai) into the flat (chars, offset, length) table
// that trpa_load_store() packs into HBM, and keeps the id -> ordinal map.  Mirrors
// RandomInmemorySeqStoreRO (ids = full FASTA header, core/src/sequencestorage.hh:56-140) and
// RandomIndexedSeqstoreRO (ids = name column of the .fai, :318-406, core/src/faidx.h:438-524).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace taxator_b200 {

struct SeqStore {
  std::string chars;               // residues of all sequences, concatenated, no line breaks (empty if packed)
  std::vector<uint64_t> off;
  std::vector<uint32_t> len;
  std::vector<std::string> ids;
  std::unordered_map<std::string, uint32_t> index;
  // refpack file (load_refpack): the packed HBM layout instead of characters
  bool packed = false;
  int alphabet = -1;               // TRPA_ALPHA_NT / TRPA_ALPHA_AA of the packed payload
  std::vector<uint64_t> woff;      // n + 1 word offsets
  std::vector<char> payload;

  uint32_t ordinal(const std::string& id) const;  // throws SequenceNotFound
  size_t size() const { return len.size(); }
};

// whole FASTA in memory; id = complete header line after '>'
SeqStore load_fasta_inmemory(const std::string& fasta);
// FASTA + samtools-style .fai (name, length, offset, linebases, linebytes); id = .fai name column.
// A missing .fai is built in memory (name = header up to the first whitespace, faidx.h:590).
SeqStore load_fasta_indexed(const std::string& fasta, const std::string& fai);

// Refpack file (.trpk): a sequence store in the packed layout the GPU keeps in HBM (2 bit planes + N
// plane per 32 bases / six 5-bit residues per word) with lengths, word offsets and identifiers, so
// that loading a reference collection is one read and one copy instead of a FASTA parse and a pack.
// Written from a store that a context has packed (trpa_export_store); see INTEGRATION.md.
bool is_refpack_file(const std::string& path);
void write_refpack(const std::string& path, int alphabet, const std::vector<std::string>& ids,
                   const std::vector<uint64_t>& woff, const std::vector<uint32_t>& len, const std::vector<char>& payload);
SeqStore load_refpack(const std::string& path);

}  // namespace taxator_b200
