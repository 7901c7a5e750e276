// Flat NCBI taxonomy for the GPU path.  Host-side mirror of the reference's TaxonTree /
// TaxonomyInterface (core/src/taxontree.hh:46-224, core/src/ncbidata.cpp:17-209,
// core/src/taxontree.cpp:55-70) as structure-of-arrays: parent / nested-set left,right / depth.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace taxator_b200 {

struct TaxatorError : std::runtime_error {
  explicit TaxatorError(const std::string& what) : std::runtime_error(what) {}
};
// same failure classes as core/src/exception.hh
struct TaxonNotFound : TaxatorError { using TaxatorError::TaxatorError; };
struct TaxonMappingNotFound : TaxatorError { using TaxatorError::TaxatorError; };
struct SequenceNotFound : TaxatorError { using TaxatorError::TaxatorError; };
struct FileNotFound : TaxatorError { using TaxatorError::TaxatorError; };
struct FileError : TaxatorError { using TaxatorError::TaxatorError; };
struct ParsingError : TaxatorError { using TaxatorError::TaxatorError; };

extern const std::vector<std::string> kDefaultRanks;  // core/src/constants.hh:32

struct FlatTaxonomy {
  std::vector<uint32_t> parent, left, right;
  std::vector<uint8_t> depth;
  std::vector<uint8_t> unclassified;   // Taxon::is_unclassified (ncbidata.cpp:119-126): the name of the node or of any
                                       // ancestor in the FULL tree (the root excepted) contains "unclassified"
  std::vector<std::string> taxid, name, rank;
  std::unordered_map<std::string, uint32_t> index;  // taxid -> node
  uint32_t root = 0;

  uint32_t node_of(const std::string& id) const {  // TaxonomyInterface::getNode
    auto it = index.find(id);
    if (it == index.end()) throw TaxonNotFound("bad taxon: " + id);
    return it->second;
  }
  uint32_t lca(uint32_t a, uint32_t b) const;        // TaxonomyInterface::getLCA
  bool is_parent_of(uint32_t a, uint32_t b) const {  // TaxonomyInterface::isParentOf
    return right[a] > left[b] && left[a] < left[b];
  }
  size_t size() const { return parent.size(); }
};

// parseNCBIFlatFiles + (optionally) deleteUnmarkedNodes: keep the root and every node whose rank is in
// `ranks`, re-parent the rest to the nearest kept ancestor, depth = parent depth + 1.
FlatTaxonomy load_ncbi_taxonomy(const std::string& nodes_file, const std::string& names_file,
                                const std::vector<std::string>& ranks, bool delete_unmarked);
// loadTaxonomyFromEnvironment: $TAXATORTK_TAXONOMY_NCBI/{nodes,names}.dmp[.gz]
FlatTaxonomy load_taxonomy_from_environment(const std::vector<std::string>& ranks, bool delete_unmarked);

}  // namespace taxator_b200
