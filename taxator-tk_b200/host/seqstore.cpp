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

}  // namespace taxator_b200
