// gtb_device.cuh -- device-side data layout shared by the kernels and the host orchestration (product code).
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/gtb200.h"
#include "gtb_index_host.hpp"
#include "gtb_bamscan.cuh"

namespace gtb
{
// ---- capacities of the per-warp shared-memory working set (exceeding one flags the task, never silently wrong)
constexpr int MAXP = 24;      // paths per read orientation held during chaining/extension
constexpr int MAXV = 16;      // bubbles per path
constexpr int REF_CAP = 160;  // staged index bucket references per read orientation
constexpr int MAX_SLOTS = 4;  // seed slots: 1 + (151-32)/31
constexpr int NLISTS = MAX_SLOTS * 2;
constexpr int CAND_CAP = 32;  // partial sequences during bubble expansion (reference limit: 128)
constexpr int CAND_SPILL = 480; // further candidates per warp in global scratch (32 + 480 = 512)
constexpr int CAND_V = 16;    // var nodes per candidate
constexpr int MAXLOC = 8;     // graph locations of a position (reference limit: 256)
constexpr int WL_CAP = 96;    // labels staged by one walk_read_starts/ends call
constexpr int MAX_SEQ = 152;  // bases per read (GTB_SEQ_STRIDE * 2)
// huge_kernel (third tier, one warp per task, working set in a per-warp global-memory slab): the reference's own limits.
// 512 paths (walks stop above 256, constants.hpp.in:43), 48 bubbles per path, 640 open candidates (reference: 128 + one
// round of growth), 40 locations of a position (<= 32 alleles per bubble + the reference node), 2048 staged labels.
constexpr int HUGE_P = 512, HUGE_V = 48, HUGE_C = 640, HUGE_CV = 48, HUGE_WL = 2048, HUGE_LOC = 40, HUGE_REFS = 2048;
// chain_kernel (one thread per read orientation, working set in local memory) -- small capacities, overflow -> slow_kernel
constexpr int FAST_P = 6, FAST_V = 8, FAST_C = 6, FAST_CV = 8, FAST_WL = 12, FAST_LOC = 4;
constexpr int SEED_INLINE = 10;     // index bucket references handed from probe_kernel to chain_kernel per task
constexpr int SEED_REC_BYTES = 16 + 8 * SEED_INLINE;
// Label record handed from probe_kernel to chain_kernel (fast tier): 16-byte header {labels per list (4 bits each), number of
// labels (0xFF: more than FT_LAB, the general tier's business) | nslots << 8 | needs-slow-kernel << 16, 0, 0}, then the seed
// labels of all lists in PHIndex::multi_get order, 16 bytes each: start, end, bubble order, allele number | list << 8 |
// has-variant << 16.
constexpr int FT_LAB = 14;
constexpr int LAB_REC_BYTES = 16 + 16 * FT_LAB;
#ifndef GTB_PROBE_WARPS
#define GTB_PROBE_WARPS 8
#endif
constexpr int PROBE_WARPS = GTB_PROBE_WARPS; // warps per block of probe_kernel
#ifndef GTB_CHAIN_THREADS
#define GTB_CHAIN_THREADS 128
#endif
#ifndef GTB_CHAIN_MIN_BLOCKS
#define GTB_CHAIN_MIN_BLOCKS 5
#endif
constexpr int CHAIN_THREADS = GTB_CHAIN_THREADS;     // threads per block of chain_kernel
constexpr int CHAIN_MIN_BLOCKS = GTB_CHAIN_MIN_BLOCKS; // chain_kernel (fast tier): 4 x 128 threads x 364 B of shared memory per SM
constexpr int MAX_TOUCH = 48; // bubbles touched by one read in the accumulate kernel (= HUGE_V)
using allele_mask_t = uint32_t; // allele set of one bubble on a path: bit a = allele a  (<= 32 alleles per bubble)
constexpr int MAX_ALLELES = 32;

// ---- hash of the k-mer table: GF(2)-linear, h(x) = XOR of basis[i] over the set bits i of x (32-bit result; table slot =
// top log2(capacity) bits, presence-bitmap bit = top log2(capacity) + 2 bits).  Linearity is what the probe kernel lives on:
// h(key ^ m) = h(key) ^ h(m), so the 96 Hamming-1 neighbours of a seed cost one XOR each with per-lane constants h(m), and
// h(key) itself is the XOR-reduction of one per-lane word per base.  With a pseudo-random basis the family is universal,
// which is all open addressing at load <= 0.25 asks for.  Full keys are hashed by byte tables (8 lookups).
struct HashTables
{
  uint32_t basis[64];
  uint32_t tab[8 * 256]; // tab[j * 256 + b] = hash of byte value b at byte position j
};
inline const HashTables & hash_tables()
{
  static const HashTables T = []() {
    HashTables t{};
    uint64_t x = 0x6A09E667F3BCC909ull; // splitmix64 stream
    for (int i = 0; i < 64; ++i)
    {
      x += 0x9E3779B97F4A7C15ull;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      t.basis[i] = (uint32_t)(z >> 32);
    }
    for (int j = 0; j < 8; ++j)
      for (int b = 0; b < 256; ++b)
      {
        uint32_t h = 0;
        for (int i = 0; i < 8; ++i)
          if ((b >> i) & 1)
            h ^= t.basis[8 * j + i];
        t.tab[j * 256 + b] = h;
      }
    return t;
  }();
  return T;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hash32_tab(const uint32_t * T, uint64_t k)
{
  uint32_t const lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
  return T[lo & 255u] ^ T[256 + ((lo >> 8) & 255u)] ^ T[512 + ((lo >> 16) & 255u)] ^ T[768 + (lo >> 24)] ^ T[1024 + (hi & 255u)] ^
         T[1280 + ((hi >> 8) & 255u)] ^ T[1536 + ((hi >> 16) & 255u)] ^ T[1792 + (hi >> 24)];
}
#endif
// every translation unit with kernels keeps its own device copy of the tables (no relocatable device code); called by gtb_create
int upload_hash_tables_kernels();
int upload_hash_tables_index();

struct DevLabel
{
  uint32_t start, end, var;
};

// One region resident on the device: flat graph + k-mer table + widened accumulators.  All pointers point into
// one arena allocation (a single H2D copy per region).
struct DevRegion
{
  // graph (include/gtb200.h: gtb_graph_view)
  uint32_t n_ref, n_var, n_special, n_sp_keys;
  uint32_t is_sv, n_bubbles, n_samples, table_mask;
  int table_shift, pad0; // table_shift = 32 - log2(table capacity): slot = hash32 >> table_shift
  const uint32_t * ref_order;
  const uint32_t * ref_seq_off; // [n_ref+1]
  const uint32_t * ref_var_off; // [n_ref+1]
  const uint32_t * var_order;
  const uint32_t * var_seq_off; // [n_var+1]
  const uint32_t * var_out_ref;
  const uint8_t * seq;
  const uint32_t * actual_poses;
  const uint32_t * ref_reach_poses;
  const uint32_t * sp_keys;
  const uint32_t * sp_off;
  const uint32_t * sp_list;
  const uint32_t * bubble_order; // [n_bubbles] order of bubble b's var nodes (Genotype::id)
  const uint16_t * hap_of_order; // [hap_span] bubble index by (order - hap_base), 0xFFFF = none; nullptr = search bubble_order
  uint32_t hap_base, hap_span;
  const uint8_t * var_num;       // [n_var] allele number of a var node within its bubble (graph.cpp:341-345)
  const uint32_t * pos_bucket;   // [n_pos_bucket] last ref node whose order <= pos_base + 16 * bucket (0 when none): replaces the
  uint32_t pos_base, n_pos_bucket; // binary search of Graph::get_locations_of_a_position by one load + a short scan
  const uint32_t * score_off;    // [n_bubbles+1]
  const uint32_t * cov_off;      // [n_bubbles+1]
  // index
  const IndexSlot * table;
  const uint32_t * bitmap;       // presence bits, 4 per table slot: bit (hash >> (table_shift - 2))
  const DevLabel * labels;
  // accumulators (widened; clamped on download)
  uint32_t * log_score;     // [score_off[NB] * NS]
  uint32_t * gt_cov;        // [cov_off[NB] * NS]
  uint32_t * max_log_score; // [NB * NS]
  uint32_t * amb;           // [NB * NS]
  uint32_t * amb_alt;
  uint32_t * alt_pp;
  unsigned long long * vs_clipped_reads; // [NB]
  unsigned long long * vs_mapq_squared;
  unsigned long long * pa_clipped_bp;    // [cov_off[NB]]
  unsigned long long * pa_mapq_squared;
  unsigned long long * pa_score_diff;
  unsigned long long * pa_mismatches;
  uint32_t * read_strand;                // [cov_off[NB] * 4]
  // SV graphs: ReferenceDepth as a difference array per sample (prefix-summed on download), (depth_size + 1) ints each
  int * ref_depth_delta;
  uint32_t depth_size, reference_offset;
  // phasing connections (HapSample::connections, haplotype.hpp:42) of all samples of the pool: one open-addressing table,
  // key = CONN_OCCUPIED | sample << 42 | hap1 << 26 | allele1 << 21 | hap2 << 5 | allele2 (so key order = the order the
  // entries are reported in), value = widened support count.  conn_mask == 0: connections are off.
  unsigned long long * conn_keys; // [conn_mask + 1], 0 = empty
  uint32_t * conn_vals;           // [conn_mask + 1]
  uint32_t * conn_state;          // [0] occupied slots, [1] set when an insert found the table full
  uint32_t conn_mask, pad1;
};
constexpr unsigned long long CONN_OCCUPIED = 1ull << 63;
constexpr uint32_t CONN_MAX_SAMPLES = 1u << 21, CONN_MAX_BUBBLES = 1u << 16;

// Records of one submit (possibly several regions concatenated), SoA on the device.
struct DevBatch
{
  uint32_t n_records, n_units; // n_units: upper bound (= n_records) used for sizing; exact count in counters->n_units
  const uint8_t * seq4;     // [n * GTB_SEQ_STRIDE]
  const uint16_t * lseq;
  const uint16_t * flag;
  const uint8_t * mapq;
  const int32_t * isize;
  const uint8_t * same_tid;
  const uint8_t * score_diff;
  const uint8_t * clipped;
  const int32_t * sample;
  const int32_t * mate;        // batch-global index or -1
  const int32_t * unit;        // alignment unit of each record
  const uint16_t * region;     // region slot of each record
  const int32_t * unit_record; // [n_units] record that defines the unit
  const uint8_t * leftover;    // SV only: paired record whose mate never arrived
};

// Result of aligning one read orientation (GenotypePaths summary) -- 16 bytes.
struct TaskSummary
{
  uint16_t npaths;
  uint16_t longest;
  uint16_t mm0;       // paths[0].mismatches
  uint16_t altcalls;  // alternative_call_count (genotype_paths.cpp:1043-1057)
  uint16_t bits;      // TS_* flags
  uint16_t pad;
  uint32_t path_off;  // word offset into the path pool
};
constexpr uint16_t TS_ALL_UNIQUE = 1, TS_OVERFLOW = 2, TS_COMPUTED = 4;

// Path pool record (uint32 words): start, end, rs | re << 16, mm | nvar << 16, then nvar x (order, mask)
constexpr int PATH_HDR_WORDS = 4;
constexpr int INLINE_WORDS = 24; // inline path-record words per task before spilling to the overflow pool

struct DevCounters
{
  unsigned long long path_words;     // bump cursor of the path pool
  unsigned long long n_oriented;
  unsigned long long n_pairs_scored;
  unsigned long long n_singles_scored;
  unsigned long long n_overflow;
  unsigned long long n_input_error;  // mates with equal IS_FIRST_IN_PAIR
  unsigned long long dbg_label_words; // bump cursor of the debug seed pool
  unsigned long long n_slow;         // tasks queued for slow_kernel
  unsigned long long fast_reasons[12]; // why chain_kernel handed a task to slow_kernel (same codes; [11] = probe flag)
  unsigned long long reasons[12];    // why slow_kernel handed a task to huge_kernel: refs vars paths locs labels candv cands keys tap pool len -
  unsigned long long n_huge;         // tasks queued for huge_kernel
  unsigned long long final_reasons[12]; // capacity overflows nothing could hold (reported as GTB_ERR_CAPACITY)
  // written by the batch-preparation kernels (prep_*): alignment units and aligned read orientations of this chunk, and
  // what is wrong with the input (PREP_ERR_* bits)
  uint32_t n_units, n_active, input_bits, n_deferred;
  // chain_kernel (fast tier) -> chain_general_kernel: queue length, why (T0_* codes), and how many it finished itself
  uint32_t n_gen, n_slow2; // n_slow2: tasks chain_general_kernel queued for the second slow_kernel launch
  unsigned long long t0_reasons[16];
};
// why the fast tier handed a task to the general tier
constexpr int T0_SV = 0, T0_LABELS = 1, T0_VARS = 2, T0_MULTI = 3, T0_SURVIVORS = 4, T0_NO_CHAIN = 5, T0_END_IN_BUBBLE = 6,
              T0_WALK_CAP = 7, T0_WALK_MULTI = 8, T0_SPECIAL = 9, T0_POOL = 10;
constexpr uint32_t PREP_ERR_SAMPLE = 1, PREP_ERR_DUP = 2, PREP_ERR_MATE = 4, PREP_ERR_LEN = 8, PREP_ERR_RECORD = 16;

// Record parsing on the device (gtb_bam.cu): raw htslib records of one pool -> the chunk's record columns
struct BamParams
{
  uint32_t n;                          // records of all regions of the call, concatenated
  uint32_t n_regions;
  const uint32_t * rec_begin;          // [n_regions + 1] first record of each region
  const uint16_t * slots;              // [n_regions] device region slot
  const unsigned long long * data_base; // [n_regions] where each region's data block starts in `data`
  const DevRegion * regions;
  const gtb_bam_core * core;           // [n]
  const uint8_t * data;
  const unsigned long long * data_off; // region r: its (n_r + 1) region-relative offsets start at rec_begin[r] + r
  const int32_t * rg;
  uint8_t * seq4;
  uint16_t * lseq;
  uint16_t * flag;
  uint16_t * region;
  uint8_t * mapq;
  uint8_t * same_tid;
  uint8_t * score_diff;
  uint8_t * clipped;
  uint8_t * leftover;
  int32_t * isize;
  int32_t * mate;
  int32_t * dup_of;
  unsigned long long hash_mask;          // kept bits of the name hash (all ones; fewer only in tests)
  unsigned long long * name_hash;        // [n] hash of (read group, read name)
  unsigned long long * name_hash_sorted; // [n]
  uint32_t * idx;                        // [n] 0 .. n-1
  uint32_t * idx_sorted;                 // [n]
  DevCounters * counters;
};
size_t bam_sort_temp_bytes(uint32_t n);
int launch_bam_parse(const BamParams & p, void * sort_temp, size_t sort_temp_bytes, void * stream);

// BGZF blocks -> the pool's records in merge order (gtb_bgzf.cu): everything gtb_submit_bam_records expects from the host
// (core, data, data_off, rg, sample) is produced on the device.
struct BgzfParams
{
  const uint8_t * comp; // compressed bytes of all segments
  uint8_t * out;        // inflated bytes (a segment's blocks are contiguous)
  const BgzfBlock * blocks;
  const BgzfSegment * segs;
  const BgzfFile * files;
  uint32_t n_blocks, n_files, n_slots, check_crc;
  BamQuery q;
  // one 32-byte block, read back by the host between the parts of the pipeline
  int * status;          // first error (INF_ERR_* / SCAN_ERR_*), 0 = none
  uint32_t * n_too_long; // records beyond the read-length capacity
  uint32_t * n_kept;     // m: records the iterators return (the ordering set)
  uint32_t * n_final;    // n: of those, the records the pool loop keeps
  uint32_t * need_host;  // bit 0: exact duplicates in different files, bit 1: more than 16 records of one position
  uint32_t * n_serial_files; // files whose per-block walks did not stitch (records straddling blocks): walked serially
  unsigned long long * rec_start; // [n_slots] scanned records, file-major
  unsigned long long * block_slot; // [n_block_slots] records found by the per-block walks
  BlockWalk * walks;              // [n_blocks]
  uint32_t * block_take;          // [n_blocks] records of the block the iterator reads
  uint32_t * block_dst;           // [n_blocks] their place in rec_start
  uint32_t n_block_slots, serial_walk; // serial_walk: skip the per-block walks (GTB_BGZF_SERIAL_WALK, measurements)
  uint32_t * file_nrec;           // [n_files]
  uint32_t * keep;                // [n_slots]
  uint32_t * keep_pos;            // [n_slots]
  uint8_t * filtered;             // [n_slots]
  unsigned long long * sel_start; // [m] the ordering set in (file, file order)
  uint32_t * sel_file;            // file | 0x80000000 when the pool loop filters the record
  unsigned long long * key;       // (position, length)
  unsigned long long * key_sorted;
  uint32_t * idx;
  uint32_t * idx_sorted;
  uint32_t * new_group;           // [m] sorted position j opens a new (position, length, sequence) group
  uint32_t * rank;                // [m] dense rank
  uint32_t * file_sorted;         // [m]
  uint32_t * idx_final;           // [m] ordering-set index at final position t
  uint32_t * keep2;               // [m]
  uint32_t * keep2_pos;           // [m]
  uint32_t * out_idx;             // [n]
  // the record batch in gtb_submit_bam_records' layout
  gtb_bam_core * core;
  uint8_t * data;
  unsigned long long * data_off;
  int32_t * rg;
  int32_t * sample;
};
size_t bgzf_temp_bytes(uint32_t n_slots);
// trace_events: four cudaEvent_t recorded before the inflate, after it, after the record walks, after the compaction (or null)
int launch_bgzf_front(const BgzfParams & p, void * temp, size_t temp_bytes, void * stream, void * const * trace_events);
int launch_bgzf_order(const BgzfParams & p, uint32_t m, void * temp, size_t temp_bytes, void * stream);
int launch_bgzf_back(const BgzfParams & p, uint32_t m, uint32_t n, const uint32_t * perm, void * temp, size_t temp_bytes, void * stream);
void reference_merge_order(uint32_t m, uint32_t n_files, const unsigned long long * key_sorted, const uint32_t * rank,
                           const uint32_t * idx_sorted, const uint32_t * file_sorted, std::vector<uint32_t> & perm);
// the same pipeline serially on the CPU (gtb_debug_bgzf_host)
int bgzf_host_pipeline(const uint8_t * comp, const std::vector<BgzfBlock> & blocks, const std::vector<BgzfSegment> & segs,
                       const std::vector<BgzfFile> & files, const BamQuery & q, bool check_crc, std::vector<uint8_t> & inflated,
                       std::vector<gtb_bam_core> & core, std::vector<uint8_t> & data, std::vector<unsigned long long> & data_off,
                       std::vector<int32_t> & sample, std::vector<int32_t> & rg, uint32_t * n_too_long, bool force_merge,
                       uint32_t * n_stitched);

// Batch preparation on the device (the per-record part of what genotype_only's caller does, hts_parallel_reader.cpp:655-708):
// alignment units (records that are not duplicates of an earlier one), the list of read orientations align_read aligns at
// all (alignment.cpp:331-363), link validation.
struct PrepParams
{
  uint32_t n_records;
  const int32_t * dup_of;       // batch-global index of the record whose alignment is re-used, or -1
  int32_t * mate;               // in/out: a link that does not point to an earlier record is cut (and reported)
  int32_t * sample;             // in/out: out of range -> 0 (and reported)
  const uint16_t * lseq;
  const uint16_t * flag;
  const uint8_t * same_tid;
  const int32_t * isize;
  const uint16_t * region;
  const DevRegion * regions;
  unsigned long long * scan;    // [n_records] (is_unit << 32 | orientations) -> exclusive prefix sums, in place
  int32_t * unit;               // out [n_records]
  int32_t * unit_record;        // out [n_units]
  uint32_t * active;            // out [n_active]
  DevCounters * counters;
};
void launch_prep_flags(const PrepParams & p, void * stream);
void launch_prep_fill(const PrepParams & p, void * stream);
size_t scan64_temp_bytes(uint32_t n);
int exclusive_scan64(void * temp, size_t temp_bytes, const unsigned long long * in, unsigned long long * out, uint32_t n,
                     void * stream);

// Debug tap of the seed stage: per task, per list (slot*2 + ham): count and offset into a label pool
struct DevSeedTap
{
  uint32_t * list_count; // [n_tasks * NLISTS]
  uint32_t * list_off;   // [n_tasks * NLISTS]
  uint32_t * nslots;     // [n_tasks]
  DevLabel * pool;
  unsigned long long pool_cap;
};

struct LaunchParams
{
  const DevRegion * regions;
  DevBatch batch;
  TaskSummary * summaries; // [n_units * 2]
  uint32_t * path_pool;
  unsigned long long path_pool_cap; // words
  DevCounters * counters;
  DevSeedTap tap; // tap.list_count == nullptr when disabled
  void * cand_spill; // per-resident-warp global extension of the bubble-expansion candidate list
  uint32_t n_active;            // UPPER BOUND (grid / buffer sizing) of the read orientations that are actually aligned
                                // (align_read, alignment.cpp:331-363); the exact count is counters->n_active (prep kernels)
  const uint32_t * active_tasks; // [counters->n_active] task id = unit * 2 + orientation
  void * seed_recs;             // [n_active] SeedRec
  void * lab_recs;              // [n_active] label records (LAB_REC_BYTES each)
  uint32_t * slow_tasks;        // [n_active] queue filled by chain_kernel (tasks probe_kernel marked: IUPAC/N seeds, many references)
  uint32_t * slow2_tasks;       // [n_active] queue filled by chain_general_kernel (capacity overflows)
  uint32_t * huge_tasks;        // [n_active] queue filled by slow_kernel
  void * huge_states;           // [SM count] HugeState slabs
  uint8_t * pending;            // [n_units * 2] set by chain_kernel for tasks it hands to slow_kernel; never cleared by the
                                // slower tiers, so the first score pass can read it while they run
  uint32_t * deferred;          // [n_records] records the first score pass left for the second (some task pending)
  uint32_t * gen_tasks;         // [n_active] positions (in active_tasks) chain_kernel left for chain_general_kernel
  uint32_t gen_lanes;           // chain_general_kernel: tasks per warp (1..32; fewer = less divergence serialisation)
  int filter_max_fold;          // probe_kernel: fold the presence bitmap into shared memory up to 2^this times (default
                                // FILTER_MAX_FOLD; < 0 = always probe the global bitmap) -- GTB_PROBE_FILTER_FOLD, tests
  uint32_t defer;               // 1: score_kernel runs before slow_kernel and defers; 0: it runs after and scores everything
  unsigned long long * task_times; // profiling aid (GTB_TASK_TIMES=file): [n_active][2] globaltimer ns at start / end of
                                   // each chain_kernel task; nullptr normally
};

// The chunks of one submit (gtb_submit_reads_multi pipelines copy and compute chunk by chunk); slow_kernel / huge_kernel
// serve the queues of all of them in one launch.
constexpr int MAX_CHUNKS = 8;
struct MultiLaunch
{
  int n;
  LaunchParams p[MAX_CHUNKS];
};
static_assert(sizeof(MultiLaunch) <= 4000, "MultiLaunch is passed as a kernel parameter");

// host launchers (gtb_kernels.cu)
void launch_build_table(const IndexSlot * uniq, uint32_t n, IndexSlot * table, uint32_t mask, int shift, uint32_t * bitmap,
                        void * stream);
void launch_probe(const LaunchParams & p, void * stream);
void launch_chain(const LaunchParams & p, void * stream);         // fast tier, one thread per task
void launch_chain_general(const LaunchParams & p, void * stream); // general tier over the queue chain_kernel filled
// which = 0: the queues chain_kernel filled (slow_kernel only, beside chain_general_kernel); 1: the queues chain_general_kernel
// filled, then huge_kernel over everything both slow_kernel launches re-queued
void launch_slow(const MultiLaunch & m, int which, void * stream);
// first pass: every record of one chunk whose tasks are all computed by chain_kernel; records with a task still queued for
// slow_kernel are listed in p.deferred.  Second pass (after slow_kernel / huge_kernel): the deferred records of all chunks.
void launch_score(const LaunchParams & p, bool with_connections, void * stream);
void launch_score_deferred(const MultiLaunch & m, const bool * with_connections, void * stream);
// Many small buffers in one launch: copy every segment to dst + dst_off (gather before ONE D2H) or zero it.
// Pointers and sizes are multiples of 16 bytes.
struct Segment
{
  void * ptr;
  unsigned long long dst_off;
  unsigned long long bytes;
};
// The segment table travels in the kernel's parameter space (no staging buffer, no copy to wait for): seg is a HOST array,
// launches take SEGMENTS_PER_LAUNCH entries each.
constexpr int SEGMENTS_PER_LAUNCH = 160;
struct SegmentTable
{
  Segment s[SEGMENTS_PER_LAUNCH];
};
void launch_gather_segments(const Segment * seg, int n, unsigned long long max_bytes, void * dst, void * stream);
void launch_zero_segments(const Segment * seg, int n, unsigned long long max_bytes, void * stream);

// Small record columns of a read batch, read by the device straight from the caller's page-locked (mapped) host buffers:
// one job per region of the chunk, pointers are DEVICE aliases of the caller's arrays (cudaPointerGetAttributes).  Replaces the
// host-side gather into a staging buffer + one H2D copy; mate / duplicate links are rebased to chunk-global record indices and
// the region slot column is filled on the way.  The table travels in the kernel's parameter space.
constexpr int GATHER_JOBS_PER_LAUNCH = 28;
struct ColumnJob
{
  const uint16_t *lseq, *flag;
  const uint8_t *mapq, *same_tid, *score_diff, *clipped, *leftover; // clipped / leftover may be null (zeros)
  const int32_t *isize, *sample, *mate, *dup_of;                    // mate / dup_of may be null (-1)
  uint32_t n, rec_base, slot, tile_begin;                           // tile_begin: first work item of this job
};
struct ColumnGather
{
  uint8_t * dst; // the chunk's device block (ChunkLayout offsets below)
  unsigned long long o_lseq, o_flag, o_region, o_mapq, o_same, o_sd, o_clip, o_left, o_isize, o_sample, o_mate, o_dup;
  uint32_t n_jobs, n_tiles;
  ColumnJob job[GATHER_JOBS_PER_LAUNCH];
};
constexpr uint32_t GATHER_TILE = 4096; // records per work item
void launch_gather_columns(const ColumnGather & g, void * stream);
// connection table maintenance: re-insert every entry of (keys, vals)[0..n_slots) into R's (larger, zeroed) table;
// compact the non-empty entries of R's table into (out_keys, out_vals), count in *out_n
void launch_conn_rehash(const unsigned long long * keys, const uint32_t * vals, uint32_t n_slots, const DevRegion & R, void * stream);
void launch_conn_compact(const DevRegion & R, unsigned long long * out_keys, uint32_t * out_vals, uint32_t * out_n, void * stream);
int align_kernel_blocks_per_sm();
size_t align_spill_bytes();
size_t huge_state_bytes();

// device-side index build (gtb_index_dev.cu): one descriptor per region of a gtb_region_begin_multi call
struct IdxRegion
{
  uint32_t n_ref, n_var, n_sp_keys, n_sweep;
  const uint32_t * ref_order;
  const uint32_t * ref_seq_off;
  const uint32_t * ref_var_off;
  const uint32_t * var_order;
  const uint32_t * var_seq_off;
  const uint32_t * var_out_ref;
  const uint8_t * seq;
  const uint32_t * sp_keys;
  const uint32_t * sp_off;
  const uint32_t * sp_list;
  const uint32_t * var_ev_off;  // null when the graph has no events
  const int64_t * var_ev;
  const uint32_t * var_aev_off;
  const int64_t * var_aev;
  const uint32_t * sweep_node;    // [n_sweep] (is_var << 31) | node id, in the reference's sweep order
  const uint32_t * sweep_job_off; // [n_sweep + 1] first END position (= base) of each node
  // outputs, in the region's index arena
  DevLabel * labels;
  IndexSlot * uniq;
  IndexSlot * table;
  uint32_t * bitmap;
  uint32_t table_mask;
  int table_shift;
};
void idx_launch_count(const IdxRegion * regions, uint32_t n_regions, const uint32_t * region_job_off, uint32_t total_jobs,
                      uint32_t * job_cnt, uint32_t * err, void * stream);
size_t idx_scan_temp_bytes(uint32_t n);
int idx_exclusive_scan(void * temp, size_t temp_bytes, const uint32_t * in, uint32_t * out, uint32_t n, void * stream);
void idx_launch_region_totals(const uint32_t * job_off, const uint32_t * region_job_off, uint32_t n_regions,
                              uint32_t * region_tuple_off, void * stream);
void idx_launch_emit(const IdxRegion * regions, uint32_t n_regions, const uint32_t * region_job_off, uint32_t total_jobs,
                     const uint32_t * job_off, uint64_t * keys, DevLabel * labels, uint32_t * tuple_idx, uint32_t * err, void * stream);
size_t idx_sort_temp_bytes(uint32_t total, uint32_t n_regions);
int idx_sort(void * temp, size_t temp_bytes, uint64_t * keys_a, uint64_t * keys_b, uint32_t * idx_a, uint32_t * idx_b, uint32_t * aux,
             uint32_t total, uint32_t n_regions, const uint32_t * region_tuple_off, void * stream);
void idx_launch_group(const IdxRegion * regions, const uint64_t * skeys, const uint32_t * sidx, const DevLabel * labels_emit,
                      uint32_t * head, uint32_t * head_incl, void * scan_temp, size_t scan_temp_bytes, const uint32_t * region_tuple_off,
                      uint32_t n_regions, uint32_t total, uint32_t * region_n_uniq, void * stream);

// discovery re-alignment (gtb_sw.cu)
struct SwParams
{
  int n_pairs;
  int max_db;               // rows of backtrack scratch per warp
  const uint8_t * q;
  const int32_t * q_off;    // [n_pairs + 1]
  const uint8_t * d;
  const int32_t * d_off;    // [n_pairs + 1]
  gtb_sw_result * out;
  uint32_t * bt;            // [resident warps][max_db + 5][32]
};
int sw_resident_warps(int max_db);
void launch_sw(const SwParams & p, int resident_warps, void * stream);

} // namespace gtb
