#include "taxonomy.h"

#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <set>
#include <sys/stat.h>

namespace taxator_b200 {

const std::vector<std::string> kDefaultRanks = {"superkingdom", "phylum", "class", "order", "family", "genus", "species"};

static bool file_exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }

// whole (possibly gzip-compressed) file into memory; gzread passes plain files through unchanged
static std::string slurp(const std::string& path) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw FileNotFound("could not find file: " + path);
  std::string out;
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof(buf))) > 0) out.append(buf, n);
  gzclose(f);
  return out;
}

// split on the NCBI "\t|\t" separator, at most `maxfields` leading fields + rest
static void split_ncbi(const std::string& line, std::vector<std::string>& out, int maxfields) {
  out.clear();
  size_t last = 0;
  while (maxfields && last < line.size()) {
    size_t pos = line.find("\t|\t", last);
    if (pos == std::string::npos) { out.push_back(line.substr(last)); return; }
    out.push_back(line.substr(last, pos - last));
    --maxfields;
    last = pos + 3;
  }
  out.push_back(line.substr(std::min(last, line.size())));
}

uint32_t FlatTaxonomy::lca(uint32_t a, uint32_t b) const {
  const uint32_t left_min = std::min(left[a], right[b]);
  const uint32_t right_max = std::max(right[a], right[b]);
  uint32_t x = a;
  while (left[x] > left_min || right[x] < right_max) x = parent[x];
  return x;
}

FlatTaxonomy load_ncbi_taxonomy(const std::string& nodes_file, const std::string& names_file,
                                const std::vector<std::string>& ranks, bool delete_unmarked) {
  struct Raw { std::string taxid, parent, rank, name; };
  std::vector<Raw> raw;
  std::unordered_map<std::string, uint32_t> raw_index;
  std::vector<std::string> f;
  {
    const std::string data = slurp(nodes_file);
    size_t p = 0;
    while (p < data.size()) {
      size_t e = data.find('\n', p);
      if (e == std::string::npos) e = data.size();
      std::string line = data.substr(p, e - p);
      p = e + 1;
      if (line.empty()) continue;
      split_ncbi(line, f, 3);
      if (f.size() < 3) throw ParsingError("bad line in " + nodes_file + ": " + line);
      raw_index[f[0]] = (uint32_t)raw.size();
      raw.push_back(Raw{f[0], f[1], f[2], ""});
    }
  }
  {
    const std::string data = slurp(names_file);
    size_t p = 0;
    while (p < data.size()) {
      size_t e = data.find('\n', p);
      if (e == std::string::npos) e = data.size();
      std::string line = data.substr(p, e - p);
      p = e + 1;
      if (line.empty()) continue;
      split_ncbi(line, f, 4);
      if (f.size() >= 4 && f[3] == "scientific name\t|") {
        auto it = raw_index.find(f[0]);
        if (it != raw_index.end()) raw[it->second].name = f[1];
      }
    }
  }
  auto rit = raw_index.find("1");
  if (rit == raw_index.end()) throw ParsingError("taxonomy has no root node with taxid 1");
  const uint32_t raw_root = rit->second;
  const std::set<std::string> keep_ranks(ranks.begin(), ranks.end());

  // children lists of the raw tree (only nodes reachable from the root matter)
  std::vector<std::vector<uint32_t>> kids(raw.size());
  for (uint32_t i = 0; i < raw.size(); ++i) {
    if (i == raw_root) continue;
    auto pit = raw_index.find(raw[i].parent);
    if (pit != raw_index.end() && pit->second != i) kids[pit->second].push_back(i);
  }
  FlatTaxonomy T;
  // iterative DFS from the root; `kept_parent` = nearest kept ancestor in the output index space
  struct Frame { uint32_t raw; uint32_t out; size_t next; bool kept; bool uncl; };
  auto add_node = [&](uint32_t r, uint32_t parent_out, uint8_t d, bool uncl) {
    const uint32_t id = (uint32_t)T.parent.size();
    T.parent.push_back(parent_out == UINT32_MAX ? id : parent_out);
    T.left.push_back(0); T.right.push_back(0); T.depth.push_back(d); T.unclassified.push_back(uncl ? 1 : 0);
    T.taxid.push_back(raw[r].taxid); T.name.push_back(raw[r].name); T.rank.push_back(raw[r].rank);
    T.index[raw[r].taxid] = id;
    return id;
  };
  uint32_t counter = 0;
  std::vector<Frame> st;
  const uint32_t root_out = add_node(raw_root, UINT32_MAX, 0, false);
  T.root = root_out;
  T.left[root_out] = ++counter;
  st.push_back(Frame{raw_root, root_out, 0, true, false});
  while (!st.empty()) {
    Frame& fr = st.back();
    if (fr.next < kids[fr.raw].size()) {
      const uint32_t c = kids[fr.raw][fr.next++];
      const bool keep = !delete_unmarked || keep_ranks.count(raw[c].rank) > 0;
      const bool uncl = fr.uncl || raw[c].name.find("unclassified") != std::string::npos;
      const uint32_t parent_frame_out = fr.out;
      if (keep) {
        const uint32_t parent_out = fr.out;
        if (T.depth[parent_out] >= 62) throw ParsingError("taxonomy deeper than 62 levels is not supported");
        const uint32_t id = add_node(c, parent_out, (uint8_t)(T.depth[parent_out] + 1), uncl);
        T.left[id] = ++counter;
        st.push_back(Frame{c, id, 0, true, uncl});
      } else {
        st.push_back(Frame{c, parent_frame_out, 0, false, uncl});  // dropped: its children attach to the same kept ancestor
      }
    } else {
      if (fr.kept) T.right[fr.out] = ++counter;
      st.pop_back();
    }
  }
  return T;
}

FlatTaxonomy load_taxonomy_from_environment(const std::vector<std::string>& ranks, bool delete_unmarked) {
  const char* env = getenv("TAXATORTK_TAXONOMY_NCBI");
  if (!env) throw FileNotFound("Specify the folder containing the NCBI taxonomy dump files as TAXATORTK_TAXONOMY_NCBI environment variable");
  const std::string dir = env;
  std::string nodes = dir + "/nodes.dmp", names = dir + "/names.dmp";
  if (file_exists(nodes + ".gz")) nodes += ".gz";
  else if (!file_exists(nodes)) throw FileNotFound("\"" + nodes + "\" not found");
  if (file_exists(names + ".gz")) names += ".gz";
  else if (!file_exists(names)) throw FileNotFound("\"" + names + "\" not found");
  return load_ncbi_taxonomy(nodes, names, ranks, delete_unmarked);
}

}  // namespace taxator_b200
