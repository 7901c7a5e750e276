// Host-side launch entry points of the CUDA translation units.
#pragma once
#include "common.cuh"
#include "shapes.h"
#include "taxator_rpa_b200.h"

namespace trpa {

// banded edit distance (myers3.cuh).  launch_plan: initial threshold k0 + shape of every pair of the
// round (PairDesc.pad: hint on entry; bits 0..7 shape, 8..31 k0 on exit) and the per-shape histogram.
// band: 0 = full matrix, 1 = band, > 1 = band with that forced initial threshold (test hook).
cudaError_t launch_plan(PairDesc* pairs, u32 n_pairs, const SeqDesc* descs, const uint2* planes, const u32* nplane,
                        u32* hist, u32 lanes_total, int band, int force_shape, int wedge, const PlanParams& pp, cudaStream_t stream);
// pairs[0..count) all of `shape`; scratch: scratch_stride uint4 per group slot (slots_out != nullptr:
// only report how many slots the launch would use); stats: {word-blocks, retries, pairs, failed wedges} accumulators
cudaError_t launch_myers3(int shape, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                          const u32* nplane, const uint2* codes, int* out, uint4* scratch, u32 scratch_stride, u32* cursor,
                          unsigned long long* stats, int force_full, u32* slots_out, cudaStream_t stream);

// alignment-free placement models (lca_models.cu); Taxonomy: machine.h
struct Taxonomy;
cudaError_t launch_lca_models(const trpa_segment* segs, u32 n_segs, const trpa_candidate* cands, const double* evalue,
                              const uint8_t* uncl, const Taxonomy& tax, const trpa_lca_params& pp, trpa_result* out,
                              cudaStream_t stream);

// consensus binning (binner.cu): scratch of one trpa_bin_batch call, all device pointers
struct BinScratch {
  u32* node_support; u32* node_seen; u32* node_pruned;       // n_nodes each
  u32* lower; uint8_t* alive; uint8_t* state; u32* curnode;  // n_records each
  u32* maj_node; float* maj_sum;                             // n_records each
  unsigned short* tot;                                       // n_records x (max_depth + 1)
  u32* path_node; unsigned short* path_direct; unsigned short* path_total; uint8_t* path_branch;   // n_groups x (max_depth + 1)
};
cudaError_t launch_binner(const trpa_bin_record* recs, u32 n_records, const u32* supports, const u32* group_begin, u32 n_groups,
                          const Taxonomy& tax, u32 n_nodes, u32 max_depth, const trpa_bin_params& pp, const uint8_t* rank_of_node,
                          const float* pid_per_rank, BinScratch& S, trpa_bin_result* out, u32* stats4, cudaStream_t stream);

// BLOSUM62 linear-gap NW with traced length; out2[pair.out] = {mutual, #diagonal steps}
// scratch: scratch_stride int2 per pair, needed only when a pair's A is longer than 512 residues
// aa_mask: bit r set = residue ordinal r may occur in the staged sequences (0 = unknown: all 27); the short-pair kernel
// keeps profile rows only for those residues (more resident pairs per SM) and traps on a residue outside the mask
cudaError_t launch_protein(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues,
                           int2* out2, int2* scratch, u32 scratch_stride, u32 max_len, u32 aa_mask, cudaStream_t stream);

// ASCII -> packed stores
cudaError_t launch_pack_nt(const uint8_t* chars, const u64* off, const SeqDesc* seqs, u32 n_seq, u64 total_words,
                           uint2* planes, u32* nplane, u32* seq_flags, cudaStream_t stream);

cudaError_t ensure_blosum_constant(int device);
cudaError_t launch_aa_codes(const uint8_t* chars, uint8_t* out, u64 n, cudaStream_t stream);
cudaError_t launch_aa_mask(const u32* packed, u64 n_words, u32* mask, cudaStream_t stream);
cudaError_t launch_pack_aa(const uint8_t* chars, const u64* off, const u64* woff, const u32* len, u32 n_seq,
                           u64 total_words, u32* packed, cudaStream_t stream);
cudaError_t launch_stage_nt(const StageReq* reqs, u32 n_req, const uint2* q_planes, const u32* q_n, const u64* q_woff,
                            const uint2* r_planes, const u32* r_n, const u64* r_woff, SeqDesc* descs,
                            uint2* out_planes, u32* out_n, uint2* out_codes, cudaStream_t stream);
cudaError_t launch_codes_from_planes(const uint2* planes, uint2* codes, u64 n_words, cudaStream_t stream);
cudaError_t launch_stage_aa(const StageReq* reqs, u32 n_req, const u32* q_packed, const u64* q_woff,
                            const u32* r_packed, const u64* r_woff, SeqDesc* descs, uint8_t* out, cudaStream_t stream);
cudaError_t launch_selfscore(SeqDesc* descs, u32 n, const uint8_t* residues, cudaStream_t stream);
cudaError_t launch_unstage_nt(const SeqDesc* descs, u32 n, const uint2* planes, const u32* nplane,
                              const u64* out_off, uint8_t* out, cudaStream_t stream);
cudaError_t launch_alu_probe(u32* sink, int iters, cudaStream_t stream, int* blocks, int* threads);

}  // namespace trpa
