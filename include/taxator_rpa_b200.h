/*
 * taxator_rpa_b200.h -- C ABI of the B200-native RPA (realignment placement algorithm) hot path.
 *
 * Drop-in boundary for fungs/taxator-tk's `taxator -a rpa`: everything below replaces, for a whole
 * batch of query segments at once, what the reference does one segment at a time on a CPU thread in
 *     RPAPredictionModel::predict()          core/src/taxonpredictionmodelsequence.hh:341-838
 * i.e. the SeqAn pairwise alignments (hh:133-256), the reference/query segment fetch
 * (hh:856-880, core/src/sequencestorage.hh:341-369,430-457, core/src/faidx.h:315-350) and the
 * distance -> taxon-range reduction with LCA (core/src/taxonomyinterface.cpp:52-77).
 * The reference has no FFI today (single C++ process); INTEGRATION.md shows the ~60-line
 * RPAPredictionModel subclass a maintainer would add on the reference side to bind these symbols.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers unless a name says "dev";
 * every function returns 0 on success and a negative code on failure, with a message available
 * from trpa_last_error() (thread-local).  No C++ exceptions cross this boundary.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with TRPA_ERR_CUDA.
 * A context is bound to one GPU; use one context (one process or thread) per GPU.
 */
#ifndef TAXATOR_RPA_B200_H_
#define TAXATOR_RPA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRPA_ABI_VERSION 9

#define TRPA_OK 0
#define TRPA_ERR_CUDA (-1)
#define TRPA_ERR_ARG (-2)
#define TRPA_ERR_STATE (-3)
#define TRPA_ERR_NOMEM (-4)

#define TRPA_NO_NODE 0xffffffffu

typedef struct trpa_ctx trpa_ctx;

/* Sequence alphabets.  NT: SeqAn Dna5 (A,C,G,T/U -> 0..3, anything else -> N, N==N matches),
 * stored as 2-bit planes + N plane.  AA: SeqAn AminoAcid 27 letters (unknown -> X), 5 bit, 6/word. */
#define TRPA_ALPHA_NT 0
#define TRPA_ALPHA_AA 1

/* Which store a sequence table is loaded into (taxator.cpp:230-249: query store / db store). */
#define TRPA_STORE_QUERY 0
#define TRPA_STORE_REF 1

/* One alignment record of a segment == the fields of AlignmentRecordTaxonomy that predict() reads
 * (core/src/alignmentrecord.hh:39-93,207-238).  Coordinates are 1-based inclusive; rstart > rstop
 * means reverse strand (hh:872-879).  `node` indexes the arrays given to trpa_load_taxonomy. */
typedef struct trpa_candidate {
  uint32_t ref_seq;     /* ordinal in the reference store (refid2position_, sequencestorage.hh:347) */
  uint32_t rstart, rstop;
  uint32_t qstart, qstop;
  float score;
  uint32_t identities;
  uint32_t alnlen;
  uint32_t node;
} trpa_candidate;

/* One query segment == one record set handed to predict() (taxator.cpp:66-72,163-175), unmasked
 * records only, in record-set order (the library applies SortFilter's stable sort itself).
 * CONTRACT: the record sets of a table are disjoint and in ascending order -- segs[s+1].cand_begin >=
 * segs[s].cand_begin + segs[s].cand_count (trpa_predict_lca_batch: ==, no gaps) -- every entry point that takes a
 * segment table rejects anything else with TRPA_ERR_ARG (the per-candidate work arrays are laid out by it). */
typedef struct trpa_segment {
  uint32_t query_seq;   /* ordinal in the query store */
  uint32_t cand_begin;  /* first candidate of this segment in the candidate table */
  uint32_t cand_count;
  uint32_t reserved;
} trpa_segment;

#define TRPA_KIND_NONE 0      /* n==0: unclassified, hh:359-368 (ival left untouched by the reference) */
#define TRPA_KIND_SINGLE 1    /* n==1: hh:371-388 */
#define TRPA_KIND_IDENTICAL 2 /* full-length 100% hit shortcut: hh:431-472 */
#define TRPA_KIND_PLACED 3    /* three-pass placement: hh:474-837 */
#define TRPA_KIND_LCA 4       /* point estimate of an alignment-free model (trpa_predict_lca_batch) */

/* == the PredictionRecord fields predict() sets (core/src/predictionrecord.hh:38-167) plus the
 * STATS counters of hh:834-837. */
typedef struct trpa_result {
  uint32_t qrstart, qrstop;           /* query feature begin/end */
  uint32_t lower_node, upper_node;    /* setNodeRange(lower, upper, support) */
  uint32_t rtax_node;                 /* setBestReferenceTaxon */
  uint32_t support;
  float ival;                         /* setInterpolationValue; undefined for TRPA_KIND_NONE */
  float signal;                       /* setSignalStrength (always 0, hh:722-725) */
  uint32_t n_pass0, n_pass1, n_pass2; /* alignments computed per pass */
  uint32_t kind;
  uint64_t cells;                     /* sum |A|*|B| over those alignments (GCUPS numerator) */
} trpa_result;

/* One consumed pairwise alignment of a segment, in the order the reference would compute them (what the verbose
 * per-alignment log of `taxator -l` is written from, hh:516-837).  a / b: record index inside the segment's record
 * set in SortFilter order (the `i` of the reference's log), TRPA_TRACE_QUERY for the query range.  r0 / r1: the raw
 * integers of the alignment -- nucleotide: r0 = edit distance; protein: r0 = mutual BLOSUM62 score, r1 = number of
 * diagonal steps of the traced path, self = sum of the two self scores (hh:190-191). */
#define TRPA_TRACE_QUERY 0xffffffffu
typedef struct trpa_trace_entry {
  uint32_t seg;
  uint32_t a, b;
  int32_t r0, r1;
  uint32_t len_a, len_b;
  uint32_t self;
} trpa_trace_entry;

/* Per-kernel device time accumulated since the last reset (CUDA events on the context's stream). */
typedef struct trpa_profile {
  double ms_edit_distance;  uint64_t launches_edit_distance;
  uint64_t cells_edit_distance;   /* DP cells the edit-distance kernel EXECUTED (band only; <= algorithmic cells) */
  double ms_protein;        uint64_t launches_protein;        uint64_t cells_protein;
  double ms_stage;          uint64_t launches_stage;          uint64_t bytes_stage;
  double ms_decide;         uint64_t launches_decide;
  double ms_other;          uint64_t launches_other;
  uint64_t rounds;
  uint64_t pairs;
  uint64_t band_retries;          /* pairs re-run (threshold too small, or a wedge whose certificate failed) */
  uint64_t wedge_failures;        /* of those: wedges whose certificate failed */
} trpa_profile;

/* ---- lifecycle --------------------------------------------------------------------------- */
int trpa_abi_version(void);
const char* trpa_last_error(void);
/* cuda_stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL for an own stream. */
trpa_ctx* trpa_create(int device, void* cuda_stream);
void trpa_destroy(trpa_ctx* ctx);
/* exclude_factor = taxator -x (default 0.5), toppercent = taxator -t (default 0.05); taxator.cpp:252 */
int trpa_set_params(trpa_ctx* ctx, float exclude_factor, float toppercent);
/* upper bound for the per-chunk staging arena in bytes (default: 1/4 of free HBM) */
int trpa_set_arena_bytes(trpa_ctx* ctx, uint64_t bytes);
/* look-ahead budget for passes 1/2 (extra alignments a segment may request per round to shorten the
 * chain of dependent rounds; results are identical for every value).  -1 (default): automatic, only
 * when the GPU has idle capacity; 0: off. */
int trpa_set_lookahead(trpa_ctx* ctx, int k);
/* Edit-distance band: 1 (default) = compute only the cells inside an Ukkonen band whose threshold is
 * an upper bound of the distance (verified, widened and re-run if it was not): same integers, fewer
 * cells.  0 = the full DP matrix like the reference's bit-vector loop (A/B runs). */
int trpa_set_band(trpa_ctx* ctx, int on);
/* Test / tuning hooks (results never depend on them): "band_k0" = forced initial band threshold
 * (exercises the verify-and-widen loop), "wedge" = 0 keeps the band at constant width (1: let it narrow
 * where the kernel's certificate proves that exact), "plan_lanes" = weight of a pair's latency in the shape
 * planner (0: the resident lanes), "tail_log2" = its convex term time^2 / 2^tail_log2, "hint_mul64" /
 * "hint_add" = safety margin on distance estimates, "cost_word10" / "cost_col10" / "cost_step" / "cost_setup" /
 * "cost_setup_w" = the planner's instruction-cost model, "la_cap" (0 = scaled with the batch) / "la_max" = look-ahead budget per round /
 * per segment, "force_shape" = one kernel shape for every pair. */
int trpa_set_tuning(trpa_ctx* ctx, const char* key, int64_t value);
int trpa_profile_reset(trpa_ctx* ctx);
int trpa_profile_get(trpa_ctx* ctx, trpa_profile* out);

/* ---- taxonomy: replaces TaxonomyInterface over TaxonTree (taxontree.hh:46-69) ------------- */
/* parent[root] == root.  left/right are the nested-set values, depth = root_pathlength (< 64). */
int trpa_load_taxonomy(trpa_ctx* ctx, const uint32_t* parent, const uint32_t* left, const uint32_t* right,
                       const uint8_t* depth, uint32_t n_nodes, uint32_t root);

/* ---- sequence stores: replace RandomInmemorySeqStoreRO / RandomIndexedSeqstoreRO ---------- */
/* chars: the concatenated residue characters (FASTA payload without line breaks), off[i]/len[i] the
 * start and length of sequence i.  Packed on the GPU into the HBM-resident store. */
int trpa_load_store(trpa_ctx* ctx, int store, int alphabet, const char* chars, const uint64_t* off,
                    const uint32_t* len, uint32_t n_seq);

/* Packed stores (the refpack file format of the host): the HBM layout that
 * trpa_load_store builds -- NT: n_words x {uint32 low-bit plane, uint32 high-bit plane} followed by
 * n_words x uint32 "is N" plane, every sequence starting on a 32-base word; AA: n_words x uint32 with
 * six 5-bit ordinals each -- read back to the host once and loaded again without the FASTA / ASCII
 * detour.  woff has n_seq + 1 entries (word offset of every sequence, then n_words). */
int trpa_store_info(trpa_ctx* ctx, int store, int* alphabet, uint32_t* n_seq, uint64_t* n_words);
int trpa_export_store(trpa_ctx* ctx, int store, uint64_t* woff, uint32_t* len, void* payload);
int trpa_load_store_packed(trpa_ctx* ctx, int store, int alphabet, const uint64_t* woff, const uint32_t* len,
                           uint32_t n_seq, const void* payload, uint64_t n_words);

/* ---- the hot path ----------------------------------------------------------------------- */
/* All segments of a batch, host buffers in / host buffers out; == predict() per segment. */
int trpa_predict_batch(trpa_ctx* ctx, const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                       uint32_t n_cands, trpa_result* out);
/* Same, split so that a caller can keep a batch resident in HBM (benchmarking, pipelining). */
int trpa_batch_upload(trpa_ctx* ctx, const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                      uint32_t n_cands);
int trpa_batch_run(trpa_ctx* ctx);
int trpa_batch_download(trpa_ctx* ctx, trpa_result* out);
/* Alignment trace (verbose log, `taxator -l`): trpa_set_trace(ctx, 1) makes the following batches record every
 * consumed alignment; trpa_batch_trace copies the entries of the last trpa_batch_run to `out` (capacity `cap`
 * entries, *n = entries available; call with cap = 0 to size the buffer), ordered by segment and, inside a segment,
 * in the reference's order.  Results never depend on tracing. */
int trpa_set_trace(trpa_ctx* ctx, int on);
int trpa_batch_trace(trpa_ctx* ctx, trpa_trace_entry* out, uint64_t cap, uint64_t* n);

/* Page-locked host memory for the tables a caller hands to trpa_predict_batch (DMA straight from the caller's buffers
 * instead of a staged copy).  Portable across contexts / devices; allocated under the device of the process's first
 * context.  Before any context exists, or without a CUDA device, the memory is ordinary host memory (only the transfer
 * speed differs; compute entry points still fail).  Thread-safe. */
void* trpa_host_alloc(uint64_t bytes);
void trpa_host_free(void* p);

/* Multi-GPU (taxator.cpp:181-210 parallelises over record sets only): segments are independent, so a batch is cut
 * into `world` contiguous shards, one per GPU / context, and the fixed-size result records are concatenated in
 * shard order.  Host-only helper, no GPU needed: bounds[0..world] = segment indices of the cuts, balanced by the
 * estimated DP work (1 + sum of span^2 over a segment's records).  A shard's segment table must be rebased
 * (cand_begin -= segs[bounds[r]].cand_begin) before it is handed to trpa_predict_batch. */
int trpa_shard_bounds(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands,
                      uint32_t world, uint32_t* bounds);
/* Device address of the result table of the last trpa_batch_run (n_segs records, valid until the next upload):
 * lets a multi-process driver gather shards GPU to GPU (NVLink) instead of through host memory. */
int trpa_batch_results_dev(trpa_ctx* ctx, void** dev_ptr, uint32_t* n_segs);

/* ---- the alignment-free placement models behind the same predict() interface ------------------ */
/* DummyPredictionModel, LCASimplePredictionModel, MeganLCAPredictionModel, NBestLCAPredictionModel
 * (core/src/taxonpredictionmodel.hh:57-259) as taxator.cpp:346-361 builds them (treat_unclassified = false;
 * "ic-megan-lca" is the same model as "megan-lca" there).  Fields = the command line options they read. */
#define TRPA_MODEL_DUMMY 0
#define TRPA_MODEL_SIMPLE_LCA 1
#define TRPA_MODEL_MEGAN_LCA 2
#define TRPA_MODEL_NBEST_LCA 3
typedef struct trpa_lca_params {
  uint32_t model;
  float toppercent;               /* -t: MinScoreMaxEvalueTopPercentFilter (alignmentsfilter.hh:349-386) */
  float minscore;                 /* -m */
  float maxevalue;                /* -e, narrowed to float like the filter's constructor argument (:351) */
  uint32_t minsupport;            /* -c */
  uint32_t nbest;                 /* -n: NumBestBitscoreFilter (alignmentsfilter.hh:493-534) */
  uint32_t ignore_unclassified;   /* -u: RemoveUnclassifiedFilter (alignmentsfilter.hh:612-623) */
  uint32_t reserved;
} trpa_lca_params;
/* Same segment / candidate tables as trpa_predict_batch (unmasked records, record-set order; only qstart,
 * qstop, score and node are read).  evalue: one double per candidate (nullable: 0); node_unclassified: one
 * byte per taxonomy node (TaxonAnnotation is_unclassified, ncbidata.cpp:119-126; nullable).  Results:
 * kind TRPA_KIND_LCA (qrstart/qrstop, lower == upper node, rtax, support = feature width; ival = -1: these
 * models never set it) or TRPA_KIND_NONE (setUnclassified; the feature range stays the whole query).
 * repeat > 1 re-runs the kernel for timing; kernel_ms (nullable) = device ms per run. */
int trpa_predict_lca_batch(trpa_ctx* ctx, const trpa_lca_params* params, const trpa_segment* segs, uint32_t n_segs,
                           const trpa_candidate* cands, uint32_t n_cands, const double* evalue,
                           const uint8_t* node_unclassified, trpa_result* out, int repeat, double* kernel_ms);

/* ---- downstream consensus: `binner` (core/binner.cpp:186-333, core/src/predictionranges.hh:122-266) -------------
 * One call = one sample: all GFF3 prediction records of the sample, grouped by (globbed) sequence identifier.
 * Step 1 (binner.cpp:213-281): per-taxon sample support = sum over records of the running maximum of the record's
 * support from its lower node up to the root (uint32 sums), optional noise pruning of range ends whose sample
 * support is below a minimum.  Step 2 (predictionranges.hh): per group, walk down from the root, keep the majority
 * branch, place the group at the deepest node whose direct support reaches the threshold ("direct") or at the
 * deepest node whose total support does ("fallback") -- with the reference's uint16 arithmetic
 * (medium_unsigned_int, core/src/types.hh:35).  Optional per-rank identity constraints (binner.cpp:305-325).
 * A record's supports[] slice is its taxon_support_ vector: upper node first, lower node last,
 * depth[lower] - depth[upper] + 1 entries (predictionrecord.hh:152-158, 330-372). */
typedef struct trpa_bin_record {
  uint32_t lower_node, upper_node;
  uint32_t support_begin;   /* first entry of this record in supports[] */
  uint32_t query_length;    /* seqlen= */
  uint32_t query_id;        /* dense id of the record's own sequence identifier (a group sums each length once) */
  uint32_t reserved;
} trpa_bin_record;
typedef struct trpa_bin_params {
  float signal_majority;              /* -j (default 0.7) */
  uint32_t min_support_per_sequence;  /* -s (default 50) */
  uint32_t min_support_in_sample;     /* -m as a position count (>= 1), or 0 */
  float min_support_in_sample_fraction; /* -m as a fraction of the root's support (value containing '.'), or 0 */
  uint32_t n_ranks;                   /* identity constraints: pid_per_rank[rank_of_node[n]] (< 0: none); 0 = no -i given */
  uint32_t reserved;
} trpa_bin_params;
#define TRPA_BIN_EMPTY 0      /* every record of the group was removed by the noise filter: no output line */
#define TRPA_BIN_SINGLE 1     /* one record: passed through (binner.cpp:301-303) */
#define TRPA_BIN_DIRECT 2
#define TRPA_BIN_FALLBACK 3
typedef struct trpa_bin_result {
  uint32_t node;            /* the taxon written to the binning file */
  uint32_t support;         /* _TaxatorTK_Support column */
  uint32_t length;          /* _TaxatorTK_Length column */
  uint32_t mode;            /* TRPA_BIN_* */
  uint32_t lower_node, upper_node;   /* the combined record's range (before the identity constraints) */
  uint32_t lower_support, upper_support;
} trpa_bin_result;
typedef struct trpa_bin_stats {
  uint64_t nested_taxa;     /* taxa with sample support ("N nested taxa", binner.cpp:253) */
  uint64_t root_support;    /* total support of the root */
  uint64_t pruned_taxa;     /* "N taxa removed" (binner.cpp:282) */
  uint64_t min_support_found;
} trpa_bin_stats;
/* group_begin has n_groups + 1 entries: group g owns records[group_begin[g] .. group_begin[g+1]) in input order.
 * rank_of_node / pid_per_rank are only read when params->n_ranks > 0 (nullable otherwise).  Uses the taxonomy
 * loaded with trpa_load_taxonomy. */
int trpa_bin_batch(trpa_ctx* ctx, const trpa_bin_params* params, const trpa_bin_record* records, uint32_t n_records,
                   const uint32_t* supports, uint32_t n_supports, const uint32_t* group_begin, uint32_t n_groups,
                   const uint8_t* rank_of_node, const float* pid_per_rank, trpa_bin_result* out, trpa_bin_stats* stats);

/* ---- lower-level entry points (unit tests, micro-benchmarks) ------------------------------ */
/* edit distance of n_pairs pairs over a private ASCII sequence table; == getAlignmentDNA distance
 * (hh:133-171).  repeat > 1 re-runs the kernels for timing; kernel_ms (nullable) = device ms/run. */
int trpa_edit_distance_batch(trpa_ctx* ctx, const char* chars, const uint64_t* off, const uint32_t* len,
                             uint32_t n_seq, const uint32_t* pair_a, const uint32_t* pair_b, uint32_t n_pairs,
                             int32_t* out_dist, int repeat, double* kernel_ms);
/* BLOSUM62 linear-gap global alignment; out3[3*i..] = {mutual score, self score, traced length};
 * == getAlignmentProtein (hh:173-242); a = horizontal (row 0), b = vertical (row 1). */
int trpa_protein_align_batch(trpa_ctx* ctx, const char* chars, const uint64_t* off, const uint32_t* len,
                             uint32_t n_seq, const uint32_t* pair_a, const uint32_t* pair_b, uint32_t n_pairs,
                             int32_t* out3, int repeat, double* kernel_ms);
/* Fetch candidate reference segments exactly as getSequence(id,start,stop,left_ext,right_ext)
 * (hh:856-880) from the loaded reference store; out_chars receives ordinals (0..4 / 0..26),
 * out_off[i] the start of segment i in out_chars, out_len[i] its length. */
int trpa_fetch_segments(trpa_ctx* ctx, const uint32_t* ref_seq, const uint32_t* start, const uint32_t* stop,
                        const uint32_t* left_ext, const uint32_t* right_ext, uint32_t n, uint8_t* out_codes,
                        uint64_t out_capacity, uint64_t* out_off, uint32_t* out_len);
/* LCA of node pairs over the loaded taxonomy; == TaxonomyInterface::getLCA. */
int trpa_lca_batch(trpa_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* out);
/* INT32 ALU-pipe probe: sustained LOP3+IADD3 lane-ops per second on this GPU (roofline denominator). */
int trpa_int_alu_peak(trpa_ctx* ctx, double* lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TAXATOR_RPA_B200_H_ */
