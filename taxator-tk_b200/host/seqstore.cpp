#include "seqstore.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

#include "taxonomy.h"

namespace taxator_b200 {

static bool file_exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }

static std::string read_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw FileNotFound("could not find file: " + path);
  std::string out;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n > 0) {
    out.resize((size_t)n);
    if (fread(&out[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); throw FileError("could not access file: " + path); }
  }
  fclose(f);
  return out;
}

uint32_t SeqStore::ordinal(const std::string& id) const {
  auto it = index.find(id);
  if (it == index.end()) throw SequenceNotFound("bad sequence identifier: " + id);
  return it->second;
}

static void add_record(SeqStore& s, const std::string& id, uint64_t off, uint32_t len) {
  s.index[id] = (uint32_t)s.ids.size();
  s.ids.push_back(id);
  s.off.push_back(off);
  s.len.push_back(len);
}

static SeqStore parse_fasta(const std::string& data, bool id_is_first_word) {
  SeqStore s;
  s.chars.reserve(data.size());
  size_t p = 0;
  const size_t n = data.size();
  while (p < n) {
    if (data[p] != '>') {  // skip anything before the first header
      size_t e = data.find('\n', p);
      p = e == std::string::npos ? n : e + 1;
      continue;
    }
    size_t e = data.find('\n', p);
    if (e == std::string::npos) e = n;
    std::string hdr = data.substr(p + 1, e - p - 1);
    if (!hdr.empty() && hdr.back() == '\r') hdr.pop_back();
    if (id_is_first_word) {
      size_t w = hdr.find_first_of(" \t");
      if (w != std::string::npos) hdr.resize(w);
    }
    p = e == n ? n : e + 1;
    const uint64_t start = s.chars.size();
    while (p < n && data[p] != '>') {
      size_t le = data.find('\n', p);
      if (le == std::string::npos) le = n;
      // bulk copy of the line; blanks inside a sequence line are rare and removed afterwards
      const size_t before = s.chars.size();
      s.chars.append(data, p, le - p);
      if (memchr(s.chars.data() + before, ' ', le - p) || memchr(s.chars.data() + before, '\t', le - p) ||
          memchr(s.chars.data() + before, '\r', le - p)) {
        size_t w = before;
        for (size_t k = before; k < s.chars.size(); ++k) {
          const char c = s.chars[k];
          if (c != ' ' && c != '\t' && c != '\r') s.chars[w++] = c;
        }
        s.chars.resize(w);
      }
      p = le == n ? n : le + 1;
    }
    add_record(s, hdr, start, (uint32_t)(s.chars.size() - start));
  }
  return s;
}

SeqStore load_fasta_inmemory(const std::string& fasta) {
  return parse_fasta(read_file(fasta), false);
}

SeqStore load_fasta_indexed(const std::string& fasta, const std::string& fai) {
  const std::string data = read_file(fasta);
  if (!file_exists(fai)) return parse_fasta(data, true);
  std::ifstream in(fai.c_str());
  if (!in.good()) throw FileError("could not access file: " + fai);
  SeqStore s;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    std::vector<std::string> f;
    size_t last = 0;
    for (;;) {
      size_t t = line.find('\t', last);
      if (t == std::string::npos) { f.push_back(line.substr(last)); break; }
      f.push_back(line.substr(last, t - last));
      last = t + 1;
    }
    if (f.size() < 5) throw ParsingError("bad FASTA index line: " + line);
    const uint64_t seqlen = std::stoull(f[1]), offset = std::stoull(f[2]);
    const uint64_t linebases = std::stoull(f[3]), linebytes = std::stoull(f[4]);
    if (seqlen > 0xffffffffull) throw ParsingError("sequence longer than 2^32 is not supported: " + f[0]);
    const uint64_t start = s.chars.size();
    // faidx.h:315-350: base b lives at offset + (b / linebases) * linebytes + (b % linebases)
    uint64_t b = 0;
    while (b < seqlen) {
      const uint64_t take = std::min<uint64_t>(linebases ? linebases : seqlen, seqlen - b);
      const uint64_t pos = offset + (linebases ? (b / linebases) * linebytes : 0);
      if (pos + take > data.size()) throw ParsingError("FASTA index points past the end of " + fasta + " for " + f[0]);
      s.chars.append(data, pos, take);
      b += take;
    }
    add_record(s, f[0], start, (uint32_t)seqlen);
  }
  return s;
}

// ------------------------------------------------------------------------------------ refpack file
namespace {
const char kRefpackMagic[8] = {'T', 'R', 'P', 'K', 1, 0, 0, 0};
struct RefpackHeader {
  char magic[8];
  uint32_t alphabet, n_seq;
  uint64_t n_words, ids_bytes, payload_bytes;
};
size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }
}  // namespace

bool is_refpack_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char m[8];
  const bool ok = fread(m, 1, 8, f) == 8 && memcmp(m, kRefpackMagic, 8) == 0;
  fclose(f);
  return ok;
}

void write_refpack(const std::string& path, int alphabet, const std::vector<std::string>& ids,
                   const std::vector<uint64_t>& woff, const std::vector<uint32_t>& len, const std::vector<char>& payload) {
  if (woff.size() != len.size() + 1 || ids.size() != len.size()) throw TaxatorError("write_refpack: inconsistent tables");
  std::string idblob;
  for (const std::string& id : ids) { idblob += id; idblob.push_back('\0'); }
  RefpackHeader h;
  memcpy(h.magic, kRefpackMagic, 8);
  h.alphabet = (uint32_t)alphabet; h.n_seq = (uint32_t)len.size();
  h.n_words = woff.back(); h.ids_bytes = idblob.size(); h.payload_bytes = payload.size();
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw FileError("could not write file: " + path);
  const char zeros[8] = {0};
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  ok = ok && fwrite(woff.data(), 8, woff.size(), f) == woff.size();
  ok = ok && (len.empty() || fwrite(len.data(), 4, len.size(), f) == len.size());
  ok = ok && fwrite(zeros, 1, pad8(4 * len.size()) - 4 * len.size(), f) == pad8(4 * len.size()) - 4 * len.size();
  ok = ok && (idblob.empty() || fwrite(idblob.data(), 1, idblob.size(), f) == idblob.size());
  ok = ok && fwrite(zeros, 1, pad8(idblob.size()) - idblob.size(), f) == pad8(idblob.size()) - idblob.size();
  ok = ok && (payload.empty() || fwrite(payload.data(), 1, payload.size(), f) == payload.size());
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw FileError("could not write file: " + path);
}

SeqStore load_refpack(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw FileNotFound("could not find file: " + path);
  SeqStore s;
  try {
    RefpackHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kRefpackMagic, 8) != 0) throw ParsingError("not a refpack file: " + path);
    if (h.alphabet > 1) throw ParsingError("refpack file with unknown alphabet: " + path);
    const uint64_t per_word = h.alphabet == 0 ? 12 : 4;
    if (h.payload_bytes != h.n_words * per_word) throw ParsingError("refpack file is inconsistent: " + path);
    s.packed = true;
    s.alphabet = (int)h.alphabet;
    s.woff.resize((size_t)h.n_seq + 1);
    std::vector<uint32_t> lens(h.n_seq);
    std::string idblob((size_t)h.ids_bytes, '\0');
    s.payload.resize((size_t)h.payload_bytes);
    char skip[8];
    bool ok = fread(s.woff.data(), 8, s.woff.size(), f) == s.woff.size();
    ok = ok && (lens.empty() || fread(lens.data(), 4, lens.size(), f) == lens.size());
    const size_t p1 = pad8(4 * lens.size()) - 4 * lens.size();
    ok = ok && (p1 == 0 || fread(skip, 1, p1, f) == p1);
    ok = ok && (idblob.empty() || fread(&idblob[0], 1, idblob.size(), f) == idblob.size());
    const size_t p2 = pad8(idblob.size()) - idblob.size();
    ok = ok && (p2 == 0 || fread(skip, 1, p2, f) == p2);
    ok = ok && (s.payload.empty() || fread(s.payload.data(), 1, s.payload.size(), f) == s.payload.size());
    if (!ok || s.woff.back() != h.n_words) throw ParsingError("refpack file is truncated or inconsistent: " + path);
    size_t p = 0;
    uint64_t off = 0;
    for (uint32_t i = 0; i < h.n_seq; ++i) {
      const size_t e = idblob.find('\0', p);
      if (e == std::string::npos) throw ParsingError("refpack file has too few identifiers: " + path);
      add_record(s, idblob.substr(p, e - p), off, lens[i]);   // off: position in the (virtual) residue stream
      off += lens[i];
      p = e + 1;
    }
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
  return s;
}

}  // namespace taxator_b200
