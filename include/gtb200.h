/* gtb200.h -- C ABI of libgtb200: the B200-native read -> pangenome-graph genotyping path.
 *
 * This is the drop-in boundary.  The reference (graphtyper v2.7.7) has no
 * FFI layer; the seams this ABI replaces are C++ call sites:
 *
 *   gtb_region_begin()   replaces  PHIndex index_graph(Graph const&)                include/graphtyper/index/indexer.hpp:16
 *                                  (called at src/utilities/genotype.cpp:294,550, genotype_sv.cpp:97)
 *                                  and uploads the flattened gyper::graph            include/graphtyper/graph/graph.hpp:39-171
 *   gtb_pool_begin()     replaces  VcfWriter writer; writer.set_samples(...)         src/utilities/hts_parallel_reader.cpp:493-494
 *   gtb_submit_reads()   replaces  the per-record body of the pool loop: get_sequence + align_read +
 *                                  update_paths + get_better_paths + VcfWriter::update_haplotype_scores_geno
 *                                  (src/utilities/hts_parallel_reader.cpp:245-338,655-708; src/typer/alignment.cpp:331,482,557;
 *                                  src/typer/vcf_writer.cpp:88-250,503-676; src/graph/haplotype.cpp:180-585)
 *   gtb_submit_bam_records()  the same with the per-record derivations (sequence, AS-XS, duplicate shortcut, read-name
 *                                  maps) on the device; the caller still reads the records with HtsParallelReader
 *   gtb_submit_bgzf()    replaces  that reader as well: HtsReader::get_next_read_in_order + HtsParallelReader::read_record +
 *                                  the pool loop's filters (src/utilities/hts_reader.cpp:166-303, hts_parallel_reader.cpp:66-136,
 *                                  528-568,655-663) and, below them, htslib's bgzf_read_block / bam_read1 / hts_itr_next -- the
 *                                  caller hands over the files' COMPRESSED bytes
 *   gtb_pool_finish()    replaces  reading VcfWriter::haplotypes[*].hap_samples[*]   include/graphtyper/typer/vcf_writer.hpp:51-52
 *                                  (input of Vcf::add_haplotype, src/typer/vcf.cpp:1507)
 *   gtb_calls_from_accumulators()  = get_haplotype_phred + SampleCall ctor/get_gt_call/get_gq
 *                                  (src/typer/vcf.cpp:47-81, src/typer/sample_call.cpp:34-131)
 *
 * Conventions: every function returns 0 on success or a negative gtb_status; gtb_last_error() gives the
 * message (the C++ shim on the reference side turns that into print_log(error)+exit(1), the reference's
 * own error convention).  All pointers are plain host pointers; sizes are element counts.  No torch or
 * CUDA types appear here.  All positions are the reference's 1-based ABSOLUTE positions
 * (include/graphtyper/graph/absolute_position.hpp:13-39); special positions are >= GTB_SPECIAL_START.
 */
#ifndef GTB200_H
#define GTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTB_K 32u                       /* include/graphtyper/constants.hpp.in:20 */
#define GTB_INVALID_ID 0xFFFFFFFFu      /* constants.hpp.in:21 */
#define GTB_SPECIAL_START 0xD0000000u   /* constants.hpp.in:33 */
#define GTB_MAX_READ_LENGTH 151u        /* constants.hpp.in:27 */
#define GTB_SEQ_STRIDE 76u              /* bytes of 4-bit bases per read: ceil(151/2) */

typedef enum gtb_status {
  GTB_OK = 0,
  GTB_ERR_ARG = -1,        /* bad argument / inconsistent view */
  GTB_ERR_CUDA = -2,       /* CUDA runtime failure (or no device: the product path never falls back to the CPU) */
  GTB_ERR_STATE = -3,      /* call order violated */
  GTB_ERR_CAPACITY = -4,   /* a device-side capacity was exceeded (paths per read, alleles per bubble ...) */
  GTB_ERR_INPUT = -5,      /* input the reference aborts on (e.g. two mates both first-in-pair) */
  GTB_ERR_NCCL = -6
} gtb_status;

/* Flattened gyper::Graph (ref_nodes / var_nodes, graph.hpp:45-52): ref node r is followed by the bubble
 * var nodes [ref_var_off[r], ref_var_off[r+1]) (allele 0 = reference allele, graph.cpp:341-345), which all
 * lead to ref node r+1.  Sequence bytes are ASCII. */
typedef struct gtb_graph_view {
  uint32_t n_ref;
  uint32_t n_var;
  int32_t is_sv_graph;
  int32_t reserved;
  const uint32_t *ref_order;    /* [n_ref]    Label::order of each ref node */
  const uint64_t *ref_seq_off;  /* [n_ref+1]  into seq */
  const uint32_t *ref_var_off;  /* [n_ref+1]  CSR of out_var_ids */
  const uint32_t *var_order;    /* [n_var] */
  const uint64_t *var_seq_off;  /* [n_var+1]  into seq */
  const uint32_t *var_out_ref;  /* [n_var]    VarNode::out_ref_id */
  const uint8_t *seq;           /* [seq_len] */
  uint64_t seq_len;
  const uint32_t *var_ev_off;   /* [n_var+1]  VarNode::events (sorted), may be NULL when no events */
  const int64_t *var_ev;
  const uint32_t *var_aev_off;  /* [n_var+1]  VarNode::anti_events */
  const int64_t *var_aev;
  uint32_t n_special;           /* Graph::actual_poses / ref_reach_poses (graph.cpp:1759-1803) */
  const uint32_t *actual_poses;
  const uint32_t *ref_reach_poses;
  uint32_t n_sp_keys;           /* Graph::ref_reach_to_special_pos, keys ascending */
  const uint32_t *sp_keys;      /* [n_sp_keys] */
  const uint32_t *sp_off;       /* [n_sp_keys+1] */
  const uint32_t *sp_list;      /* special position codes */
} gtb_graph_view;

/* One KmerLabel (include/graphtyper/index/kmer_label.hpp:13-38). */
typedef struct gtb_label {
  uint32_t start;
  uint32_t end;
  uint32_t var_id;
} gtb_label;

/* Records of one pool in the order the reference's k-way merge delivers them
 * (HtsParallelReader::read_record, hts_parallel_reader.cpp:98-136), after the flag filter. */
typedef struct gtb_read_batch {
  uint32_t n_reads;
  uint32_t seq_stride;        /* bytes per read in seq4 (>= (max lseq+1)/2); GTB_SEQ_STRIDE typical */
  const uint8_t *seq4;        /* BAM 4-bit bases, high nibble first (bam_get_seq) */
  const uint16_t *lseq;       /* core.l_qseq */
  const uint16_t *flag;       /* core.flag */
  const uint8_t *mapq;        /* core.qual */
  const int32_t *isize;       /* core.isize clamped to int32 */
  const uint8_t *same_tid;    /* core.tid == core.mtid */
  const uint8_t *score_diff;  /* AS-XS as get_score_diff() computes it (alignment.cpp:140-325) */
  const uint8_t *clipped;     /* clipped_count(rec) > 3 (alignment.cpp:105-138; always 0 in the reference) */
  const int32_t *sample;      /* sample index within the pool */
  const int32_t *mate;        /* index (in this batch) of the earlier record this one pairs with, else -1
                                 (read-name map semantics of genotype_only, hts_parallel_reader.cpp:270-337) */
  const int32_t *dup_of;      /* index of the record whose alignment is re-used (equal_pos_seq shortcut,
                                 hts_parallel_reader.cpp:666-684), else -1 */
  const uint8_t *leftover;    /* SV graphs only, may be NULL: 1 = paired record whose mate never arrived in this pool
                                 (processed alone at pool end, hts_parallel_reader.cpp:719-772) */
} gtb_read_batch;

/* Per-pool accumulators = VcfWriter::haplotypes[b].hap_samples[s] (+ var_stats), bubble-major:
 * for bubble b (cnum = n_alleles[b], tri = cnum(cnum+1)/2) and sample s
 *   log_score     [ score_off[b] * n_samples + s * tri .. +tri )
 *   gt_coverage   [ cov_off[b]   * n_samples + s * cnum .. +cnum )
 *   per-sample scalars at [b * n_samples + s]
 *   per-allele stats at [cov_off[b] + a]
 * Buffers are owned by the caller and sized with gtb_accumulator_sizes(). */
typedef struct gtb_accumulators {
  uint32_t n_bubbles;
  uint32_t n_samples;
  uint32_t *bubble_id;        /* [n_bubbles]  Genotype::id = order of the bubble's var nodes */
  uint32_t *n_alleles;        /* [n_bubbles] */
  uint64_t *score_off;        /* [n_bubbles+1] prefix sum of tri */
  uint64_t *cov_off;          /* [n_bubbles+1] prefix sum of cnum */
  uint16_t *log_score;        /* [score_off[n_bubbles] * n_samples] */
  uint16_t *gt_coverage;      /* [cov_off[n_bubbles] * n_samples] */
  uint16_t *max_log_score;    /* [n_bubbles * n_samples] */
  uint8_t *ambiguous_depth;   /* [n_bubbles * n_samples] */
  uint8_t *ambiguous_depth_alt;
  uint8_t *alt_proper_pair_depth;
  uint32_t *saturated;        /* [n_bubbles * n_samples] 1 when the max_log_score guard (haplotype.cpp:561) fired:
                                 result is order dependent beyond ~6000x depth, out of contract */
  uint64_t *vs_clipped_reads; /* [n_bubbles]  VarStats::clipped_reads */
  uint64_t *vs_mapq_squared;  /* [n_bubbles] */
  uint64_t *pa_clipped_bp;    /* [cov_off[n_bubbles]] VarStatsPerAllele */
  uint64_t *pa_mapq_squared;
  uint64_t *pa_score_diff;
  uint64_t *pa_mismatches;
  uint32_t *read_strand;      /* [cov_off[n_bubbles] * 4]  r1_forward r1_reverse r2_forward r2_reverse */
  /* SV graphs only (ReferenceDepth, src/graph/reference_depth.cpp:114-203): per-sample per-base depth of the
   * graph-aligned reads over [reference_offset, reference_offset + depth_size); ref_depth may be NULL. */
  uint32_t depth_size;
  uint32_t reference_offset;
  uint16_t *ref_depth;        /* [n_samples * depth_size] */
} gtb_accumulators;

/* Counters of one gtb_submit_reads call (diagnostics + bench bookkeeping). */
typedef struct gtb_submit_stats {
  uint64_t n_records;         /* records in the batch */
  uint64_t n_alignments;      /* distinct (record) alignments run (non-duplicates) */
  uint64_t n_oriented;        /* read orientations aligned (1 or 2 per alignment) */
  uint64_t n_pairs_scored;    /* pairs passed to the accumulator */
  uint64_t n_singles_scored;
  uint64_t n_capacity_overflow; /* reads that exceeded a device capacity (must be 0; else GTB_ERR_CAPACITY) */
  uint64_t kernel_launches;   /* kernels launched by this call */
} gtb_submit_stats;

typedef struct gtb_ctx gtb_ctx;

const char *gtb_last_error(void);
const char *gtb_version(void);

/* device_id < 0: host-only context (index build, host finalisation); compute entry points then fail. */
int gtb_create(int device_id, gtb_ctx **out);
void gtb_destroy(gtb_ctx *ctx);

/* Region: graph upload + index build (device by default, gtb_set_index_build). Regions are identified by a small integer so
 * several regions can be resident and processed by ONE batched launch (region-batched mode). */
int gtb_region_begin(gtb_ctx *ctx, int region_id, const gtb_graph_view *graph);
/* Several regions at once: ONE launch sequence builds all their indexes on the device (or, with the host builder,
 * the builds run in parallel threads and the uploads follow). */
int gtb_region_begin_multi(gtb_ctx *ctx, int n, const int *region_ids, const gtb_graph_view *graphs);
int gtb_region_end(gtb_ctx *ctx, int region_id);
/* Threading: a context serves ONE host thread at a time; different contexts may be used from different threads concurrently
 * (own CUDA streams and buffers each).  This is the reference's model -- one pool per worker thread, no locks on the hot
 * path (paw::Station in src/typer/caller.cpp:272-391; SURVEY.md section 8b "Threading") -- and the index is shared by all
 * pool threads by const pointer (PHIndex const *, include/graphtyper/utilities/hts_parallel_reader.hpp:69-82):
 * gtb_region_attach makes region `owner_region_id` of `owner` (same device) usable in `ctx` as `region_id` WITHOUT copying
 * graph or index; pools (gtb_pool_begin ... gtb_pool_finish) are per context.  The owner must keep the region alive and must
 * not end it while it is attached elsewhere; gtb_region_end(ctx, region_id) detaches. */
int gtb_region_attach(gtb_ctx *ctx, int region_id, gtb_ctx *owner, int owner_region_id);

/* Index inspection (parity with PHIndex contents): keys ascending, labels in bucket order. */
int gtb_index_size(gtb_ctx *ctx, int region_id, uint64_t *n_keys, uint64_t *n_labels);
int gtb_index_export(gtb_ctx *ctx, int region_id, uint64_t *keys, uint32_t *label_off, gtb_label *labels);

/* Pool of samples for one region.
 * A submit that returns an error (bad links, a read longer than 2 * GTB_SEQ_STRIDE bases, a device capacity, two mates with the
 * same first-in-pair flag ...) may already have added the batch's other records to the pool's accumulators: the pool is then
 * undefined until gtb_pool_reset (or a new gtb_pool_begin), and the library enforces it: further submits on that pool return
 * GTB_ERR_STATE until it is reset.  Callers that want to recover rather than exit must check what
 * they can BEFORE submitting -- the reference-side reader (integration/gtb_pool_reader.cpp) checks read lengths while it
 * collects the records, and bubbles beyond the allele capacity are refused by gtb_region_begin before anything runs. */
int gtb_pool_begin(gtb_ctx *ctx, int region_id, int n_samples);
int gtb_submit_reads(gtb_ctx *ctx, int region_id, const gtb_read_batch *batch, gtb_submit_stats *stats);
int gtb_accumulator_sizes(gtb_ctx *ctx, int region_id, uint32_t *n_bubbles, uint64_t *n_scores, uint64_t *n_cov);
/* SV graphs: length and first absolute position of the per-base reference depth track (0 for non-SV graphs). */
int gtb_ref_depth_size(gtb_ctx *ctx, int region_id, uint32_t *depth_size, uint32_t *reference_offset);
int gtb_pool_finish(gtb_ctx *ctx, int region_id, gtb_accumulators *out);
/* Several regions with one stream synchronisation (outs[i] belongs to region_ids[i]); zero several pools. */
int gtb_pool_finish_multi(gtb_ctx *ctx, int n, const int *region_ids, gtb_accumulators *outs);
int gtb_pool_reset_multi(gtb_ctx *ctx, int n, const int *region_ids);

/* Region-batched submission: regions[i] / batches[i] for i < n; one fused launch sequence for all. */
int gtb_submit_reads_multi(gtb_ctx *ctx, int n, const int *region_ids, const gtb_read_batch *batches,
                           gtb_submit_stats *stats);

/* Debug/parity taps of the last gtb_submit_reads on a region (sizes via the *_sizes call first).
 * Seeds: for alignment a (= non-duplicate record, in batch order), orientation o, ham h (0 exact, 1 Hamming-1),
 * slot i: labels in PHIndex::multi_get order (ph_index.cpp:66-107). */
int gtb_debug_enable(gtb_ctx *ctx, int on);
int gtb_debug_seed_sizes(gtb_ctx *ctx, int region_id, uint64_t *n_units, uint64_t *n_slots, uint64_t *n_labels);
int gtb_debug_seeds(gtb_ctx *ctx, int region_id, uint32_t *unit_record, uint32_t *nslots /*[n_units*2*2]*/,
                    uint32_t *nlabels, gtb_label *labels);
/* Paths: GenotypePaths of align_read (alignment.cpp:331) per (alignment, orientation). */
int gtb_debug_path_sizes(gtb_ctx *ctx, int region_id, uint64_t *n_units, uint64_t *n_paths, uint64_t *n_vars,
                         uint64_t *n_nums);
int gtb_debug_paths(gtb_ctx *ctx, int region_id, uint32_t *gp_npaths /*[n_units*2]*/, uint32_t *gp_longest,
                    uint32_t *p_fields /*[n_paths*6]: start end rs re mm nvar*/, uint32_t *v_order,
                    uint32_t *v_nnum, uint16_t *v_nums);

/* Host finalisation (pure function of the accumulators): PL (uint8), GT pair, GQ per bubble x sample. */
int gtb_calls_from_accumulators(const gtb_accumulators *acc, uint8_t *phred /*[n_scores*n_samples]*/,
                                uint16_t *gt /*[n_bubbles*n_samples*2]*/, uint8_t *gq /*[n_bubbles*n_samples]*/);

/* The remaining SampleCall fields (SampleCall ctor, src/typer/sample_call.cpp:34-61): per bubble x sample
 *   ref_total_depth = min(0xFFFF, coverage[0] + ambiguous_depth - ambiguous_depth_alt)
 *   alt_total_depth = min(0xFFFF, sum(coverage[1..]) + ambiguous_depth)
 * (coverage = gt_coverage, AD; ambiguous_depth = MD; alt_proper_pair_depth = PP are the accumulators themselves). */
int gtb_sample_depths(const gtb_accumulators *acc, uint16_t *ref_total_depth /*[n_bubbles*n_samples]*/,
                      uint16_t *alt_total_depth /*[n_bubbles*n_samples]*/);

/* Bench/ops helpers: re-run the kernels on the batch already resident in HBM; device times (ms) of the last
 * submit/replay from CUDA events on the library's stream; zero a region's accumulators. */
int gtb_replay_last(gtb_ctx *ctx, gtb_submit_stats *stats);
int gtb_last_timing(gtb_ctx *ctx, float *h2d_ms, float *align_ms, float *score_ms, float *d2h_ms);
/* align_ms = batch preparation + probe_kernel + chain_kernel; score_ms = everything after chain_kernel measured as ONE span
 * (chain_general_kernel and the first slow_kernel launch run beside the first score pass, then slow_kernel / huge_kernel for
 * what they re-queued and the second score pass), so align_ms + score_ms is the device time of the launch sequence when the
 * submit was a single chunk.  gtb_last_kernel_timing: chain_ms = both chain tiers, slow_ms = both slow launches + huge_kernel,
 * score_ms = both score passes, each summed over the chunks although they overlap. */
int gtb_last_kernel_timing(gtb_ctx *ctx, float *probe_ms, float *chain_ms, float *slow_ms, float *score_ms,
                           uint64_t *n_slow);
/* The two tiers of the chaining / end-extension stage in the last submit/replay: device time (ms) of chain_kernel (fast tier,
 * one thread per read orientation, registers + shared memory) and of chain_general_kernel (the tasks the fast tier queued),
 * the number of read orientations aligned, how many of them went to the general tier and why (reasons16: SV graph, > 8 seed
 * labels, > 4 bubbles on a path, path splits, > 2 full chains, no full chain, chain end inside a bubble, walk capacity, walk
 * splits, special position, path pool; the rest unused).  Any pointer may be NULL. */
int gtb_last_chain_timing(gtb_ctx *ctx, float *fast_ms, float *general_ms, uint64_t *n_tasks, uint64_t *n_general,
                          uint64_t *reasons16);
/* Device time (ms) of the batch-preparation kernels of the last submit/replay: alignment units of the duplicate-read
 * shortcut (hts_parallel_reader.cpp:666-684), the orientations align_read aligns (alignment.cpp:343-360), link checks. */
int gtb_last_prep_timing(gtb_ctx *ctx, float *prep_ms);
int gtb_pool_reset(gtb_ctx *ctx, int region_id);
/* Pipeline chunks of gtb_submit_reads_multi: staging of chunk k+1 overlaps copy + kernels of chunk k (0 = automatic). */
int gtb_set_chunks(gtb_ctx *ctx, int n_chunks);
/* Where gtb_region_begin(_multi) builds the k-mer index (replaces index_graph, src/index/indexer.cpp:246-291):
 * 1 (default with a CUDA device) = on the device, all regions of a call in one launch sequence; 0 = host builder.
 * Both produce the same index (keys, labels, bucket order).  Environment override: GTB_INDEX_BUILD=host|device. */
int gtb_set_index_build(gtb_ctx *ctx, int on_device);
/* Page-locked host memory: batch columns (seq4 above all) placed here are DMA-ed without a staging copy. */
int gtb_host_alloc(size_t bytes, void **out);
/* Diagnostics: out24[0..11] why chain_kernel re-queued tasks for slow_kernel, out24[12..23] why slow_kernel re-queued
 * tasks for huge_kernel (codes: refs vars paths locs labels candv cands keys tap pool len probe-flag). */
int gtb_debug_counters(gtb_ctx *ctx, uint64_t *out24);
int gtb_host_free(void *p);

/* NCCL bootstrap (libnccl is bound lazily with dlopen): rank 0 creates the 128-byte unique id, the caller
 * distributes it (e.g. torch.distributed.broadcast), every rank calls gtb_nccl_init. */
int gtb_nccl_unique_id(uint8_t *id128);
int gtb_nccl_init(gtb_ctx *ctx, int n_ranks, int rank, const uint8_t *id128);

/* Per-pool variant summary = Variant::scan_calls (src/typer/variant.cpp:230-428) on the calls of one pool, and its
 * cross-pool merge VarStats::add_stats (src/typer/var_stats.cpp:141-189: sums, except max for maximum_alt_support and
 * its ratio; n_max_alt_proper_pairs is SUMMED there as a uint8, reproduced).  Layouts (row-major):
 *   var[b][9]     n_genotyped n_calls n_passed_calls n_max_alt_proper_pairs seqdepth het_ad0 het_ad1 hom_ad0 hom_ad1
 *   allele[a][13] qd_qual qd_depth total_depth ac pass_ac n_ref_ref n_ref_alt n_alt_alt maximum_alt_support
 *                 het_multi0 het_multi1 hom_multi0 hom_multi1          (a runs over cov_off[b] + allele)
 *   ratio[a]      maximum_alt_support_ratio */
int gtb_scan_calls(const gtb_accumulators *acc, const uint8_t *phred, uint64_t *var, uint64_t *allele, double *ratio);
/* gtb_calls_from_accumulators + gtb_scan_calls for n regions in one call; the rows of the regions follow each other in
 * var / allele / ratio (the layout gtb_allreduce_varstats and gtb_merge_varstats take). */
int gtb_scan_calls_multi(int n, const gtb_accumulators *accs, uint64_t *var, uint64_t *allele, double *ratio);
int gtb_merge_varstats(uint32_t n_bubbles, uint64_t n_alleles_total, uint64_t *var, uint64_t *allele, double *ratio,
                       const uint64_t *var_src, const uint64_t *allele_src, const double *ratio_src);

/* Record-parsing entry (SURVEY.md section 8f, N3 -- first step: everything genotype_only's caller derives per record moves
 * to the device; for THIS entry BGZF inflate and the k-way merge stay in htslib -- gtb_submit_bgzf below takes those as well).  The shim hands over the records of one pool in merge
 * order (after the flag filter, hts_parallel_reader.cpp:655-663, and for SV graphs is_good_read, :528-568) as htslib holds
 * them: the core fields below (bam1_core_t, htslib/sam.h) + bam1_t::data (qname | cigar | seq | qual | aux), concatenated.
 * On the device: 4-bit bases, lengths, flags, MAPQ, insert size, AS-XS exactly as get_score_diff walks the aux block
 * (src/typer/alignment.cpp:140-325), the duplicate shortcut equal_pos_seq (include/graphtyper/utilities/hts_utils.hpp:110-128,
 * hts_parallel_reader.cpp:666-684), mate pairing with the read-name map semantics of genotype_only (:270-337: a record whose
 * name waits in its read group's map pairs with it -- paired flag or not --, otherwise a paired record waits, an unpaired one
 * is scored alone) by sorting 64-bit (read group, name) hashes and verifying the names, and for SV graphs the leftover mates
 * (:719-772).  Then the same kernels as gtb_submit_reads.  All records of a pool must come in ONE call (mates pair within it). */
typedef struct gtb_bam_core {
  int64_t pos, mpos, isize;   /* core.pos, core.mpos, core.isize */
  int32_t tid, mtid;
  int32_t l_qseq;
  uint32_t n_cigar;
  uint16_t flag;
  uint16_t l_qname;           /* core.l_qname: includes the NUL terminator and htslib's padding NULs */
  uint8_t mapq;               /* core.qual */
  uint8_t reserved[3];
} gtb_bam_core;
typedef struct gtb_bam_batch {
  uint32_t n_reads;
  uint32_t reserved;
  const gtb_bam_core *core;   /* [n_reads] */
  const uint8_t *data;        /* bam1_t::data of every record, back to back */
  const uint64_t *data_off;   /* [n_reads + 1] */
  const int32_t *sample;      /* [n_reads] sample index within the pool (HtsParallelReader::get_sample_and_rg_index) */
  const int32_t *rg;          /* [n_reads] read-group index: one read-name map per read group */
} gtb_bam_batch;
int gtb_submit_bam_records(gtb_ctx *ctx, int region_id, const gtb_bam_batch *batch, gtb_submit_stats *stats);
/* BGZF entry (SURVEY.md section 8f, N3 -- second step: the decode itself).  The shim hands over, for every BAM file of the pool,
 * the file's COMPRESSED bytes that cover the region -- the chunks of the region iterator (hts_itr_t::off, htslib hts.c:4046-4098;
 * adjacent chunks merged), each starting at a BGZF block boundary and running far enough behind the chunk's end for the last
 * record that is read (or to the end of the file).  On the device: DEFLATE + CRC-32 of every block (one warp per block),
 * record boundaries, the iterator's rules (a chunk is read while the offset behind the last record is below its end; the first
 * record on another contig or at / beyond the region's end finishes the file; records that do not overlap the region are
 * skipped), the pool loop's flag filter (hts_parallel_reader.cpp:655-663) and, for SV calling, is_good_read (:528-568), the
 * merge order of HtsReader::get_next_read_in_order + the heap of HtsParallelReader (hts_reader.cpp:166-303,
 * hts_parallel_reader.cpp:66-136: ascending position, sequence length, packed sequence bytes; exact ties in file order), and
 * from there everything gtb_submit_bam_records does.  Limits (the shim keeps the reference's host reader for these): CRAM,
 * files with more than one read group (sample / read-group index are per file here), the coverage-bin cap, an empty BGZF block
 * before the end of the file, a file that is not in coordinate order (the reference's heap merge is the sorted order only for
 * sorted files; unmapped reads behind the mapped ones count as unsorted), more than 65 534 contigs.  A stream that does not decode, a malformed record or bytes that end before the chunk does return
 * GTB_ERR_INPUT, a read beyond the length capacity GTB_ERR_CAPACITY -- both BEFORE anything is added to the pool's
 * accumulators, so the caller can fall back to its own reader for this pool. */
typedef struct gtb_bgzf_segment {
  const uint8_t *comp;      /* compressed bytes; comp[0] is the first byte of a BGZF block */
  uint64_t comp_bytes;
  uint64_t file_offset;     /* of comp[0] in the file (virtual offsets are relative to the file) */
  uint64_t v_end;           /* virtual offset at which the chunk ends (hts_pair64_t::v) */
  uint32_t first_offset;    /* of the first record inside the first block's inflated bytes (hts_pair64_t::u & 0xFFFF) */
  uint32_t to_eof;          /* the bytes run to the end of the file */
} gtb_bgzf_segment;
typedef struct gtb_bgzf_file {
  uint32_t n_segments;
  uint32_t reserved;
  const gtb_bgzf_segment *segments; /* in file order */
  int32_t sample;           /* sample index within the pool */
  int32_t rg;               /* read-group index within the pool */
} gtb_bgzf_file;
typedef struct gtb_bgzf_query {
  int32_t tid;              /* contig of the region */
  uint32_t flag_filter;     /* Options::sam_flag_filter */
  int64_t beg, end;         /* 0-based, half open (hts_itr_t::beg / end) */
  uint32_t sv_read_filter;  /* 1: is_good_read (SV calling) */
  uint32_t check_crc;       /* 1: verify the CRC-32 of every block as htslib does */
  uint32_t whole_file;      /* 1: no region -- every record from the segment's first record to the end of the file, as
                             * HtsReader reads with region "." (src/utilities/hts_reader.cpp:94-97; `genotype` reads its
                             * pools' bamshrink output this way); tid / beg / end are ignored */
  uint32_t reserved;
} gtb_bgzf_query;
int gtb_submit_bgzf(gtb_ctx *ctx, int region_id, int n_files, const gtb_bgzf_file *files, const gtb_bgzf_query *query,
                    gtb_submit_stats *stats);
/* Parity taps.  gtb_debug_bgzf_records: the record batch the last gtb_submit_bgzf built on the device, in the layout of
 * gtb_bam_batch; sizes first: pass NULL arrays.  gtb_debug_bgzf_host: the same decode + selection + order computed serially
 * on the CPU from the same source functions (test infrastructure of the CPU suite; no device needed, never used by a submit). */
int gtb_debug_bgzf_records(gtb_ctx *ctx, uint32_t *n_reads, uint64_t *n_data, gtb_bam_core *core, uint8_t *data,
                           uint64_t *data_off, int32_t *sample, int32_t *rg);
/* Files whose record boundaries were found by the per-block walks (every guessed block entry confirmed) rather than by the
 * serial walk: of the last gtb_submit_bgzf of ctx, or of the last gtb_debug_bgzf_host when ctx is NULL. */
int gtb_debug_bgzf_stitched(gtb_ctx *ctx, uint32_t *n_files_stitched);
int gtb_debug_bgzf_host(int n_files, const gtb_bgzf_file *files, const gtb_bgzf_query *query, uint32_t *n_reads,
                        uint64_t *n_data, gtb_bam_core *core, uint8_t *data, uint64_t *data_off, int32_t *sample, int32_t *rg,
                        uint64_t *n_inflated, uint8_t *inflated);

/* Several regions' pools in one launch sequence (region_ids[i] / batches[i]); pairing and the duplicate shortcut never cross
 * a pool.  gtb_debug_bam_columns then addresses the concatenation of the batches. */
int gtb_submit_bam_records_multi(gtb_ctx *ctx, int n, const int *region_ids, const gtb_bam_batch *batches,
                                 gtb_submit_stats *stats);
/* Debug/parity tap: the per-record columns the device derived in the last gtb_submit_bam_records call (any pointer may be NULL). */
int gtb_debug_bam_columns(gtb_ctx *ctx, uint32_t n_reads, uint8_t *seq4 /*[n*GTB_SEQ_STRIDE]*/, uint16_t *lseq, uint16_t *flag,
                          uint8_t *mapq, int32_t *isize, uint8_t *same_tid, uint8_t *score_diff, int32_t *mate, int32_t *dup_of,
                          uint8_t *leftover);

/* Phasing connections = HapSample::connections (include/graphtyper/graph/haplotype.hpp:42): for every read / read pair that
 * explains alleles of several bubbles, support counts between (bubble hap1, allele1) and a LATER bubble's (hap2, allele2),
 * as VcfWriter::push_to_haplotype_scores builds them (src/typer/vcf_writer.cpp:587-637: weight 6/(n1*n2) inside one read)
 * and update_haplotype_scores_geno merges them over the two mates (vcf_writer.cpp:186-249: +1 per cross-mate pair).
 * The reference only consumes them when is_writing_hap (discovery iterations, hts_parallel_reader.cpp:782), so they are OFF
 * by default; gtb_set_connections(ctx, 1) applies to pools begun afterwards.  Counts are uint16 in the reference (wrap
 * around at 65536, reproduced).  hap = bubble index within the region (the reference's uint16 haplotype index). */
typedef struct gtb_connection {
  uint32_t sample;
  uint16_t hap1, allele1, hap2, allele2;
  uint32_t count;
} gtb_connection;
int gtb_set_connections(gtb_ctx *ctx, int on);
/* After gtb_pool_finish: number of non-zero entries, then the entries sorted by (sample, hap1, allele1, hap2, allele2). */
int gtb_connections_size(gtb_ctx *ctx, int region_id, uint64_t *n);
int gtb_connections(gtb_ctx *ctx, int region_id, gtb_connection *out);
/* The phase-support map `ph` the pool derives from connections + allele coverage when is_writing_hap
 * (src/utilities/hts_parallel_reader.cpp:782-893): for bubbles less than 100 bp apart, per pair of ALT alleles, the OR over
 * samples of IS_ANY_HAP_SUPPORT (1) / IS_ANY_ANTI_HAP_SUPPORT (2) (constants.hpp.in:56-57).  Pure host function.
 * Call with out == NULL to get the count; entries sorted by (hap1, allele1, hap2, allele2). */
typedef struct gtb_phase_support_entry {
  uint16_t hap1, allele1, hap2, allele2;
  int8_t flags;
  uint8_t pad[7];
} gtb_phase_support_entry;
int gtb_phase_support(const gtb_accumulators *acc, uint64_t n_conn, const gtb_connection *conn, uint64_t *n_out,
                      gtb_phase_support_entry *out);
/* Read-sharded multi-GPU runs (one sample's reads split over ranks): the ranks' connection lists of the same pool are
 * additive (every entry comes from one read or read pair, and mates stay together).  Merges two sorted lists into one
 * sorted list, counts added modulo 2^16 like the reference's uint16 counters; out == NULL returns the size in *n_out. */
int gtb_merge_connections(uint64_t n_a, const gtb_connection *a, uint64_t n_b, const gtb_connection *b, uint64_t *n_out,
                          gtb_connection *out);

/* Multi-GPU: sum-reduce widened accumulators of a region over an NCCL communicator supplied by the host
 * (ncclComm_t passed as void*; all ranks call; result valid on every rank).  Used when ONE sample's reads
 * are split over GPUs; with sample sharding no collective is needed.
 * The sum is taken in place: afterwards the pool holds the total of all ranks, so another submit or another reduce would
 * count the other ranks' reads again -- both return GTB_ERR_STATE until gtb_pool_reset.  Connection lists are not reduced
 * here: merge them with gtb_merge_connections. */
int gtb_allreduce_accumulators(gtb_ctx *ctx, int region_id, void *nccl_comm);
/* Several regions in ONE NCCL group and one stream synchronisation. */
int gtb_allreduce_accumulators_multi(gtb_ctx *ctx, int n, const int *region_ids, void *nccl_comm);

/* Sample-sharded runs (the reference's own pool split src/typer/caller.cpp:272-391): every rank owns a
 * block of samples, accumulators and calls stay local, and the ONE exchange at the end is the cross-pool merge of the
 * per-bubble summaries, VarStats::add_stats (src/typer/var_stats.cpp:141-189): every counter summed, maximum_alt_support and
 * maximum_alt_support_ratio max-ed, n_max_alt_proper_pairs summed as a uint8 -- three collectives in one NCCL group.
 * var / allele / ratio: the gtb_scan_calls layouts, any number of regions back to back (n_var_rows bubbles,
 * n_allele_rows alleles in total); in place, result valid on every rank. */
int gtb_allreduce_varstats(gtb_ctx *ctx, uint64_t n_var_rows, uint64_t n_allele_rows, uint64_t *var, uint64_t *allele,
                           double *ratio, void *nccl_comm);
/* The same with the cross-pool merge of this rank's own pools folded in: n_src summaries of the same regions (one per pool
 * thread) are merged (VarStats::add_stats) while they are packed for the collective; the result lands in var[0] / allele[0] /
 * ratio[0]. */
int gtb_allreduce_varstats_multi(gtb_ctx *ctx, int n_src, uint64_t n_var_rows, uint64_t n_allele_rows, uint64_t *const *var,
                                 uint64_t *const *allele, double *const *ratio, void *nccl_comm);

/* Discovery re-alignment.  Replaces paw::pairwise_alignment(read, haplotype window, opts)
 * + AlignmentResults::get_database_begin_end + apply_clipping as realign_to_indels calls them
 * (src/typer/caller.cpp:1864-1870 options, :2007 the call; paw/include/paw/align/pairwise_alignment.hpp:146-388):
 * query global, database ends free, match 1 / mismatch 4 / gap open 7 / extend 1, clip 5.
 * Pair k aligns query[q_off[k]..q_off[k+1]) (<= GTB_SW_MAX_QUERY bases, raw ASCII) against
 * database[d_off[k]..d_off[k+1]) (<= GTB_SW_MAX_DATABASE).  Results are identical to paw's
 * AlignmentResults::{score, database_begin, database_end, clip_begin, clip_end}. */
#define GTB_SW_MAX_QUERY 160
#define GTB_SW_MAX_DATABASE 2048
typedef struct gtb_sw_result
{
  int32_t score;          /* after clipping (alignment_results.hpp:451-669) */
  int32_t database_begin; /* alignment_results.hpp:392-447 */
  int32_t database_end;
  int32_t clip_begin;     /* query bases [clip_begin, clip_end) are kept */
  int32_t clip_end;
} gtb_sw_result;
int gtb_sw_align_batch(gtb_ctx *ctx, int n_pairs, const uint8_t *query, const int32_t *q_off, const uint8_t *database,
                       const int32_t *d_off, gtb_sw_result *out);
/* Device time (ms, CUDA events) of the last gtb_sw_align_batch kernel, and of its copies. */
int gtb_sw_last_timing(gtb_ctx *ctx, float *kernel_ms, float *h2d_ms, float *d2h_ms);
/* Bench helper: re-runs the kernel of the last gtb_sw_align_batch on its device-resident inputs. */
int gtb_sw_replay_last(gtb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GTB200_H */
