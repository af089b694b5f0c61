// gtb_bgzf.cu -- BGZF blocks -> the pool's BAM records in merge order, on the device (product code; SURVEY.md section 8f, N3:
// the decode half of src/utilities/hts_reader.cpp:166-303 + hts_parallel_reader.cpp:66-136, which the reference leaves to
// htslib / libdeflate and a heap over the files).
//
//   bgzf_inflate_kernel   one warp per BGZF block: header, DEFLATE, ISIZE and CRC-32 (gtb_inflate.cuh)
//   bam_walk_kernel       one thread per file: record boundaries in file order, as far as the file's region iterator reads
//   bam_classify_kernel   one thread per scanned record: iterator overlap rule (-> the ORDERING set), flag filter, SV read
//                         filter (-> dropped after the merge, like the pool loop does), length capacity
//   cub exclusive scan + bam_compact_kernel: the ordering set in (file, file order), first sort key (position, length)
//   cub stable radix sort, bam_tie_kernel (one thread per run of equal keys: heap sort by the packed sequence bytes),
//   bam_rank_kernel + cub inclusive scan: dense ranks (exact duplicates share one), detection of what the device order cannot
//                         decide: exact duplicates in different files, more than 16 records of one position
//   [only then] reference_merge_order() on the host: the reference's per-file std::sort + std::push_heap / pop_heap merge replayed
//                         on (rank, file) integers -- the standard library's own algorithms, so the order of ties is the
//                         reference binary's; bam_permute_kernel applies it
//   bam_final_kernel + scan + bam_select_kernel: the pool loop's filters; bam_length_kernel + scan + bam_gather_kernel (one warp
//                         per record): core fields, data blocks, sample and read-group columns in the layout
//                         gtb_submit_bam_records takes -- resident, no record byte goes back to the host
// The per-record rules are the host/device functions of gtb_bamscan.cuh; bgzf_host_pipeline() below runs the same functions
// serially on the CPU (gtb_debug_bgzf_host: test infrastructure for the CPU suite, never called by the product path).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gtb_device.cuh"

namespace gtb
{
namespace
{
constexpr int INFLATE_WARPS = 8;

__global__ void __launch_bounds__(INFLATE_WARPS * 32) bgzf_inflate_kernel(BgzfParams p)
{
  __shared__ InflateTables tables[INFLATE_WARPS];
  uint32_t const b = blockIdx.x * INFLATE_WARPS + (threadIdx.x >> 5);
  if (b >= p.n_blocks)
    return;
  BgzfBlock const blk = p.blocks[b];
  int const rc = bgzf_inflate_block(p.comp + blk.comp_off, blk.comp_bytes, p.out + blk.out_off, tables[threadIdx.x >> 5], p.check_crc != 0);
  if ((threadIdx.x & 31) == 0 && rc != INF_OK)
    atomicCAS(p.status, 0, rc);
}

__global__ void __launch_bounds__(64) bam_walk_blocks_kernel(BgzfParams p)
{
  uint32_t const b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.n_blocks || *p.status != 0)
    return;
  BgzfBlock const & blk = p.blocks[b];
  BgzfSegment const & s = p.segs[blk.segment];
  p.walks[b] = bam_walk_block(p.out, blk, s, b == s.block_begin, p.q, p.block_slot + blk.slot_base);
}

__global__ void __launch_bounds__(32) bam_walk_kernel(BgzfParams p)
{
  uint32_t const f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.n_files || *p.status != 0)
    return;
  int st = SCAN_OK;
  uint32_t n = 0;
  BgzfFile const & file = p.files[f];
  if (p.serial_walk || !bam_stitch_file(p.blocks, p.segs, file, p.walks, p.block_take, p.block_dst, &n, &st))
  {
    // a record straddles two blocks where the per-block walks guessed a boundary: the serial walk decides
    for (uint32_t si = file.seg_begin; si < file.seg_end; ++si)
      for (uint32_t b = p.segs[si].block_begin; b < p.segs[si].block_end; ++b)
        p.block_take[b] = 0;
    n = bam_walk_file(p.out, p.blocks, p.segs, file, p.q, p.rec_start, &st);
    atomicAdd(p.n_serial_files, 1u);
  }
  p.file_nrec[f] = n;
  if (st != SCAN_OK)
    atomicCAS(p.status, 0, st);
}

__global__ void __launch_bounds__(256) bam_walk_scatter_kernel(BgzfParams p)
{
  uint32_t const slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= p.n_block_slots || *p.status != 0)
    return;
  uint32_t lo = 0, hi = p.n_blocks; // last block with slot_base <= slot
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (p.blocks[mid].slot_base <= slot)
      lo = mid;
    else
      hi = mid;
  }
  uint32_t const i = slot - p.blocks[lo].slot_base;
  if (i < p.block_take[lo])
    p.rec_start[p.block_dst[lo] + i] = p.block_slot[slot];
}

// slot -> file (slots of a file are contiguous)
__device__ __forceinline__ uint32_t file_of_slot(const BgzfParams & p, uint32_t slot)
{
  uint32_t lo = 0, hi = p.n_files;
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (p.files[mid].rec_base <= slot)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) bam_classify_kernel(BgzfParams p)
{
  uint32_t const slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= p.n_slots)
    return;
  uint32_t keep = 0;
  uint8_t filtered = 0;
  if (*p.status == 0)
  {
    uint32_t const f = file_of_slot(p, slot);
    if (slot - p.files[f].rec_base < p.file_nrec[f])
    {
      const uint8_t * rec = p.out + p.rec_start[slot];
      int const c = bam_classify(rec, bam_fixed(rec), p.q);
      keep = c == BAM_KEEP || c == BAM_FILTERED;
      filtered = c == BAM_FILTERED;
      if (c == BAM_TOO_LONG)
        atomicAdd(p.n_too_long, 1u);
      if (c == BAM_KEEP)
        atomicAdd(p.n_final, 1u);
    }
  }
  p.keep[slot] = keep;
  p.filtered[slot] = filtered;
}

__global__ void __launch_bounds__(256) bam_compact_kernel(BgzfParams p)
{
  uint32_t const slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= p.n_slots)
    return;
  uint32_t const at = p.keep_pos[slot];
  if (slot + 1 == p.n_slots)
    *p.n_kept = at + p.keep[slot];
  if (!p.keep[slot])
    return;
  unsigned long long const g = p.rec_start[slot];
  p.sel_start[at] = g;
  p.sel_file[at] = file_of_slot(p, slot) | (p.filtered[slot] ? 0x80000000u : 0u);
  BamFixed const f = bam_fixed(p.out + g);
  if (!bam_key_fits(f))
    atomicCAS(p.status, 0, SCAN_ERR_KEY);
  p.key[at] = bam_order_key(f);
  p.idx[at] = at;
}

// heap sort of idx_sorted[q0 .. q1) by (packed sequence bytes, index): the records of one position and length
__global__ void __launch_bounds__(128) bam_tie_kernel(BgzfParams p, uint32_t n)
{
  uint32_t const q0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (q0 >= n)
    return;
  unsigned long long const key = p.key_sorted[q0];
  if (q0 > 0 && p.key_sorted[q0 - 1] == key)
    return;
  uint32_t q1 = q0 + 1;
  while (q1 < n && p.key_sorted[q1] == key)
    ++q1;
  uint32_t const m = q1 - q0;
  if (m < 2)
    return;
  uint32_t * a = p.idx_sorted + q0;
  auto less = [&](uint32_t x, uint32_t y) {
    const uint8_t * rx = p.out + p.sel_start[x];
    const uint8_t * ry = p.out + p.sel_start[y];
    return bam_tie_less(rx, bam_fixed(rx), p.sel_file[x] & 0x7FFFFFFFu, x, ry, bam_fixed(ry), p.sel_file[y] & 0x7FFFFFFFu, y);
  };
  auto sift = [&](uint32_t root, uint32_t end) {
    for (;;)
    {
      uint32_t child = 2 * root + 1;
      if (child >= end)
        return;
      if (child + 1 < end && less(a[child], a[child + 1]))
        ++child;
      if (!less(a[root], a[child]))
        return;
      uint32_t const t = a[root];
      a[root] = a[child];
      a[child] = t;
      root = child;
    }
  };
  for (uint32_t i = m / 2; i-- > 0;)
    sift(i, m);
  for (uint32_t end = m - 1; end > 0; --end)
  {
    uint32_t const t = a[0];
    a[0] = a[end];
    a[end] = t;
    sift(0, end);
  }
}

// The merge of the files is the sorted order only if every file is in coordinate order itself (and the reader's "records of one
// position" groups do not mix contigs): checked on the ordering set, which is in (file, file order).
__global__ void __launch_bounds__(256) bam_sorted_check_kernel(BgzfParams p, uint32_t m)
{
  uint32_t const at = blockIdx.x * blockDim.x + threadIdx.x;
  if (at == 0 || at >= m || (p.sel_file[at] & 0x7FFFFFFFu) != (p.sel_file[at - 1] & 0x7FFFFFFFu))
    return;
  unsigned long long const a = p.key[at - 1], b = p.key[at];
  if (key_place(b) < key_place(a) || (key_pos(a) == key_pos(b) && key_place(a) != key_place(b)))
    atomicCAS(p.status, 0, SCAN_ERR_UNSORTED);
}

// new_group[j] = record j of the sorted ordering set differs from j - 1 in (position, length, sequence); what only the
// reference's own algorithms can order is flagged in *p.need_host
__global__ void __launch_bounds__(256) bam_rank_kernel(BgzfParams p, uint32_t m)
{
  uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m)
    return;
  uint32_t const x = p.idx_sorted[j];
  uint32_t const fx = p.sel_file[x] & 0x7FFFFFFFu;
  p.file_sorted[j] = fx;
  uint32_t fresh = 1;
  if (j > 0 && p.key_sorted[j - 1] == p.key_sorted[j])
  {
    uint32_t const y = p.idx_sorted[j - 1];
    const uint8_t * rx = p.out + p.sel_start[x];
    const uint8_t * ry = p.out + p.sel_start[y];
    if (bam_seq_compare(ry, bam_fixed(ry), rx, bam_fixed(rx)) == 0)
    {
      fresh = 0;
      if ((p.sel_file[y] & 0x7FFFFFFFu) != fx)
        atomicOr(p.need_host, 1u); // exact duplicates in different files: the heap's history decides
    }
  }
  p.new_group[j] = fresh;
  // more than 16 records of one position: std::sort leaves insertion sort for introsort
  if (j + 16 < m && (j == 0 || key_place(p.key_sorted[j - 1]) != key_place(p.key_sorted[j])) &&
      key_place(p.key_sorted[j + 16]) == key_place(p.key_sorted[j]))
    atomicOr(p.need_host, 2u);
}

// final order t -> index into the ordering set; perm == nullptr: the device order stands
__global__ void __launch_bounds__(256) bam_final_kernel(BgzfParams p, uint32_t m, const uint32_t * perm)
{
  uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m)
    return;
  uint32_t const x = p.idx_sorted[perm ? perm[t] : t];
  p.idx_final[t] = x;
  p.keep2[t] = (p.sel_file[x] & 0x80000000u) ? 0u : 1u; // the pool loop's filters, after the merge
}

__global__ void __launch_bounds__(256) bam_select_kernel(BgzfParams p, uint32_t m)
{
  uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m || !p.keep2[t])
    return;
  p.out_idx[p.keep2_pos[t]] = p.idx_final[t];
}

__global__ void __launch_bounds__(256) bam_length_kernel(BgzfParams p, uint32_t n)
{
  uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  p.data_off[j] = (unsigned long long)(le32(p.out + p.sel_start[p.out_idx[j]]) - 32u);
  if (j + 1 == n)
    p.data_off[n] = 0; // the exclusive scan over n + 1 entries leaves the total here
}

__global__ void __launch_bounds__(256) bam_gather_kernel(BgzfParams p, uint32_t n)
{
  uint32_t const j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned const lane = threadIdx.x & 31u;
  if (j >= n)
    return;
  uint32_t const src_i = p.out_idx[j];
  const uint8_t * rec = p.out + p.sel_start[src_i];
  BamFixed const f = bam_fixed(rec);
  if (lane == 0)
  {
    gtb_bam_core c;
    c.pos = f.pos;
    c.mpos = f.mpos;
    c.isize = f.tlen;
    c.tid = f.tid;
    c.mtid = f.mtid;
    c.l_qseq = f.l_seq;
    c.n_cigar = f.n_cigar;
    c.flag = (uint16_t)f.flag;
    c.l_qname = (uint16_t)f.l_read_name; // the file's length: no in-memory padding NULs here
    c.mapq = (uint8_t)f.mapq;
    c.reserved[0] = c.reserved[1] = c.reserved[2] = 0;
    p.core[j] = c;
    BgzfFile const & file = p.files[p.sel_file[src_i] & 0x7FFFFFFFu];
    p.rg[j] = file.rg;
    p.sample[j] = file.sample;
  }
  uint32_t const len = (uint32_t)f.block_size - 32u;
  uint8_t * dst = p.data + p.data_off[j];
  const uint8_t * src = rec + 36;
  for (uint32_t i = lane; i < len; i += 32)
    dst[i] = src[i];
}
} // namespace

size_t bgzf_temp_bytes(uint32_t n_slots)
{
  size_t a = 0, b = 0, c = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_slots);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int)n_slots);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n_slots + 1);
  return std::max({a, b, c}) + 256;
}

// first part: up to the sizes (the host needs them for what follows): ordering set m = *n_kept, final batch n = *n_final
int launch_bgzf_front(const BgzfParams & p, void * temp, size_t temp_bytes, void * stream, void * const * trace_events)
{
  cudaStream_t const s = (cudaStream_t)stream;
  auto mark = [&](int i) {
    if (trace_events)
      cudaEventRecord((cudaEvent_t)trace_events[i], s);
  };
  mark(0);
  if (p.n_blocks)
    bgzf_inflate_kernel<<<(p.n_blocks + INFLATE_WARPS - 1) / INFLATE_WARPS, INFLATE_WARPS * 32, 0, s>>>(p);
  mark(1);
  if (p.n_blocks && !p.serial_walk)
    bam_walk_blocks_kernel<<<(p.n_blocks + 63) / 64, 64, 0, s>>>(p);
  if (p.n_files)
    bam_walk_kernel<<<(p.n_files + 31) / 32, 32, 0, s>>>(p);
  if (p.n_block_slots && !p.serial_walk)
    bam_walk_scatter_kernel<<<(p.n_block_slots + 255) / 256, 256, 0, s>>>(p);
  mark(2);
  if (p.n_slots)
  {
    bam_classify_kernel<<<(p.n_slots + 255) / 256, 256, 0, s>>>(p);
    if (cub::DeviceScan::ExclusiveSum(temp, temp_bytes, p.keep, p.keep_pos, (int)p.n_slots, s) != cudaSuccess)
      return -1;
    bam_compact_kernel<<<(p.n_slots + 255) / 256, 256, 0, s>>>(p);
  }
  mark(3);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// second part: the ordering set sorted by (position, length, sequence), dense ranks, *need_host
int launch_bgzf_order(const BgzfParams & p, uint32_t m, void * temp, size_t temp_bytes, void * stream)
{
  cudaStream_t const s = (cudaStream_t)stream;
  if (m == 0)
    return 0;
  bam_sorted_check_kernel<<<(m + 255) / 256, 256, 0, s>>>(p, m);
  if (cub::DeviceRadixSort::SortPairs(temp, temp_bytes, p.key, p.key_sorted, p.idx, p.idx_sorted, (int)m, 0, 64, s) != cudaSuccess)
    return -1;
  bam_tie_kernel<<<(m + 127) / 128, 128, 0, s>>>(p, m);
  bam_rank_kernel<<<(m + 255) / 256, 256, 0, s>>>(p, m);
  if (cub::DeviceScan::InclusiveSum(temp, temp_bytes, p.new_group, p.rank, (int)m, s) != cudaSuccess)
    return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// third part: final order (perm: device array from reference_merge_order, or nullptr), the pool loop's filters, the n records
// gathered into (core, data, data_off, rg, sample)
int launch_bgzf_back(const BgzfParams & p, uint32_t m, uint32_t n, const uint32_t * perm, void * temp, size_t temp_bytes, void * stream)
{
  cudaStream_t const s = (cudaStream_t)stream;
  if (m == 0 || n == 0)
    return 0;
  bam_final_kernel<<<(m + 255) / 256, 256, 0, s>>>(p, m, perm);
  if (cub::DeviceScan::ExclusiveSum(temp, temp_bytes, p.keep2, p.keep2_pos, (int)m, s) != cudaSuccess)
    return -1;
  bam_select_kernel<<<(m + 255) / 256, 256, 0, s>>>(p, m);
  bam_length_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n);
  if (cub::DeviceScan::ExclusiveSum(temp, temp_bytes, p.data_off, p.data_off, (int)n + 1, s) != cudaSuccess)
    return -1;
  bam_gather_kernel<<<(n + 7) / 8, 256, 0, s>>>(p, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// The reference's order of a pool's records, replayed on integers.  Input: the ordering set sorted by the device, per sorted
// position j its key (position in the high word), dense rank (exact duplicates share one), index in (file, file order) and
// file.  Per file (hts_reader.cpp:196-233): the records of one position, in file order, are sorted with std::sort by
// "greater" and handed out from the back.  Over the files (hts_parallel_reader.cpp:82-136): a std::push_heap / pop_heap heap
// of one head per file under "greater".  Both comparators see (position, length, sequence) only = the rank.  perm[t] = j.
void reference_merge_order(uint32_t m, uint32_t n_files, const unsigned long long * key_sorted, const uint32_t * rank,
                           const uint32_t * idx_sorted, const uint32_t * file_sorted, std::vector<uint32_t> & perm)
{
  struct Rec
  {
    uint32_t rank, j, order;
  };
  std::vector<std::vector<Rec>> per_file(n_files);
  for (uint32_t j = 0; j < m; ++j)
    per_file[file_sorted[j]].push_back(Rec{rank[j], j, idx_sorted[j]});
  auto greater = [](Rec const & a, Rec const & b) { return a.rank > b.rank; };
  std::vector<std::vector<Rec>> streams(n_files);
  for (uint32_t f = 0; f < n_files; ++f)
  {
    std::vector<Rec> & v = per_file[f];
    std::sort(v.begin(), v.end(), [](Rec const & a, Rec const & b) { return a.order < b.order; }); // file order (orders are distinct)
    std::vector<Rec> & out = streams[f];
    out.reserve(v.size());
    std::vector<Rec> group;
    for (size_t i = 0; i < v.size();)
    {
      size_t e = i + 1;
      while (e < v.size() && key_pos(key_sorted[v[e].j]) == key_pos(key_sorted[v[i].j]))
        ++e;
      group.assign(v.begin() + i, v.begin() + e);
      std::sort(group.begin(), group.end(), greater);
      for (size_t k = group.size(); k-- > 0;)
        out.push_back(group[k]);
      i = e;
    }
  }
  struct Head
  {
    uint32_t rank, file;
  };
  auto head_greater = [](Head const & a, Head const & b) { return a.rank > b.rank; };
  std::vector<size_t> next(n_files, 0);
  std::vector<Head> heap;
  for (uint32_t f = 0; f < n_files; ++f)
    if (!streams[f].empty())
    {
      heap.push_back(Head{streams[f][0].rank, f});
      std::push_heap(heap.begin(), heap.end(), head_greater);
    }
  perm.clear();
  perm.reserve(m);
  while (!heap.empty())
  {
    uint32_t const f = heap[0].file;
    perm.push_back(streams[f][next[f]].j);
    ++next[f];
    std::pop_heap(heap.begin(), heap.end(), head_greater);
    if (next[f] < streams[f].size())
    {
      heap.back() = Head{streams[f][next[f]].rank, f};
      std::push_heap(heap.begin(), heap.end(), head_greater);
    }
    else
      heap.pop_back();
  }
}

// ---- the same pipeline, serially, on the CPU (test infrastructure: gtb_debug_bgzf_host)
int bgzf_host_pipeline(const uint8_t * comp, const std::vector<BgzfBlock> & blocks, const std::vector<BgzfSegment> & segs,
                       const std::vector<BgzfFile> & files, const BamQuery & q, bool check_crc, std::vector<uint8_t> & inflated,
                       std::vector<gtb_bam_core> & core, std::vector<uint8_t> & data, std::vector<unsigned long long> & data_off,
                       std::vector<int32_t> & sample, std::vector<int32_t> & rg, uint32_t * n_too_long, bool force_merge,
                       uint32_t * n_stitched)
{
  *n_stitched = 0;
  unsigned long long out_bytes = 0;
  for (auto const & b : blocks)
    out_bytes = std::max(out_bytes, b.out_off + b.isize);
  inflated.assign(out_bytes + 64, 0);
  InflateTables T;
  for (auto const & b : blocks)
    if (int rc = bgzf_inflate_block(comp + b.comp_off, b.comp_bytes, inflated.data() + b.out_off, T, check_crc))
      return rc;
  uint32_t n_slots = 0;
  for (auto const & f : files)
    n_slots = std::max(n_slots, f.rec_base + f.rec_cap);
  std::vector<unsigned long long> rec_start(n_slots, 0), sel_start;
  std::vector<uint32_t> sel_file;
  std::vector<uint8_t> sel_filtered;
  *n_too_long = 0;
  for (uint32_t fi = 0; fi < files.size(); ++fi)
  {
    int st = SCAN_OK;
    uint32_t const nr = bam_walk_file(inflated.data(), blocks.data(), segs.data(), files[fi], q, rec_start.data(), &st);
    {
      // the per-block walks + stitch the device uses must agree with the serial walk whenever the stitch accepts them
      std::vector<BlockWalk> walks(blocks.size());
      std::vector<uint32_t> take(blocks.size(), 0), dst(blocks.size(), 0);
      uint32_t n_block_slots = 0;
      for (auto const & b : blocks)
        n_block_slots = std::max(n_block_slots, b.slot_base + b.isize / 36 + 1);
      std::vector<unsigned long long> slot(n_block_slots, 0);
      for (uint32_t si = files[fi].seg_begin; si < files[fi].seg_end; ++si)
        for (uint32_t b = segs[si].block_begin; b < segs[si].block_end; ++b)
          walks[b] = bam_walk_block(inflated.data(), blocks[b], segs[si], b == segs[si].block_begin, q, slot.data() + blocks[b].slot_base);
      uint32_t n2 = 0;
      int st2 = SCAN_OK;
      if (bam_stitch_file(blocks.data(), segs.data(), files[fi], walks.data(), take.data(), dst.data(), &n2, &st2))
      {
        if (n2 != nr || st2 != st)
          return -30;
        for (uint32_t si = files[fi].seg_begin; si < files[fi].seg_end; ++si)
          for (uint32_t b = segs[si].block_begin; b < segs[si].block_end; ++b)
            for (uint32_t i = 0; i < take[b]; ++i)
              if (slot[blocks[b].slot_base + i] != rec_start[dst[b] + i])
                return -30;
        ++*n_stitched;
      }
    }
    if (st != SCAN_OK)
      return st;
    for (uint32_t i = 0; i < nr; ++i)
    {
      const uint8_t * rec = inflated.data() + rec_start[files[fi].rec_base + i];
      int const c = bam_classify(rec, bam_fixed(rec), q);
      if (c == BAM_TOO_LONG)
        ++*n_too_long;
      if (c == BAM_KEEP || c == BAM_FILTERED)
      {
        if (!bam_key_fits(bam_fixed(rec)))
          return SCAN_ERR_KEY;
        sel_start.push_back(rec_start[files[fi].rec_base + i]);
        sel_file.push_back(fi);
        sel_filtered.push_back(c == BAM_FILTERED);
      }
    }
  }
  uint32_t const m = (uint32_t)sel_start.size();
  for (uint32_t at = 1; at < m; ++at)
    if (sel_file[at] == sel_file[at - 1])
    {
      unsigned long long const a = bam_order_key(bam_fixed(inflated.data() + sel_start[at - 1])),
                               b = bam_order_key(bam_fixed(inflated.data() + sel_start[at]));
      if (key_place(b) < key_place(a) || (key_pos(a) == key_pos(b) && key_place(a) != key_place(b)))
        return SCAN_ERR_UNSORTED;
    }
  std::vector<uint32_t> sorted(m);
  for (uint32_t i = 0; i < m; ++i)
    sorted[i] = i;
  std::sort(sorted.begin(), sorted.end(), [&](uint32_t x, uint32_t y) {
    const uint8_t * rx = inflated.data() + sel_start[x];
    const uint8_t * ry = inflated.data() + sel_start[y];
    BamFixed const fx = bam_fixed(rx), fy = bam_fixed(ry);
    unsigned long long const kx = bam_order_key(fx), ky = bam_order_key(fy);
    if (kx != ky)
      return kx < ky;
    return bam_tie_less(rx, fx, sel_file[x], x, ry, fy, sel_file[y], y);
  });
  // dense ranks + the two conditions under which the reference's own algorithms decide, as bam_rank_kernel finds them
  std::vector<unsigned long long> key_sorted(m);
  std::vector<uint32_t> rank(m), file_sorted(m);
  uint32_t need_host = 0;
  for (uint32_t j = 0; j < m; ++j)
  {
    const uint8_t * rx = inflated.data() + sel_start[sorted[j]];
    key_sorted[j] = bam_order_key(bam_fixed(rx));
    file_sorted[j] = sel_file[sorted[j]];
  }
  for (uint32_t j = 0; j < m; ++j)
  {
    bool fresh = true;
    if (j > 0 && key_sorted[j - 1] == key_sorted[j])
    {
      const uint8_t * rx = inflated.data() + sel_start[sorted[j]];
      const uint8_t * ry = inflated.data() + sel_start[sorted[j - 1]];
      if (bam_seq_compare(ry, bam_fixed(ry), rx, bam_fixed(rx)) == 0)
      {
        fresh = false;
        if (file_sorted[j - 1] != file_sorted[j])
          need_host |= 1u;
      }
    }
    rank[j] = (j ? rank[j - 1] : 0u) + (fresh ? 1u : 0u);
    if (j + 16 < m && (j == 0 || key_place(key_sorted[j - 1]) != key_place(key_sorted[j])) &&
        key_place(key_sorted[j + 16]) == key_place(key_sorted[j]))
      need_host |= 2u;
  }
  std::vector<uint32_t> perm;
  if (need_host || force_merge)
    reference_merge_order(m, (uint32_t)files.size(), key_sorted.data(), rank.data(), sorted.data(), file_sorted.data(), perm);
  std::vector<uint32_t> order;
  for (uint32_t t = 0; t < m; ++t)
  {
    uint32_t const x = sorted[perm.empty() ? t : perm[t]];
    if (!sel_filtered[x])
      order.push_back(x);
  }
  uint32_t const n = (uint32_t)order.size();
  core.resize(n);
  sample.resize(n);
  rg.resize(n);
  data_off.assign(n + 1, 0);
  data.clear();
  for (uint32_t j = 0; j < n; ++j)
  {
    const uint8_t * rec = inflated.data() + sel_start[order[j]];
    BamFixed const f = bam_fixed(rec);
    gtb_bam_core c{};
    c.pos = f.pos;
    c.mpos = f.mpos;
    c.isize = f.tlen;
    c.tid = f.tid;
    c.mtid = f.mtid;
    c.l_qseq = f.l_seq;
    c.n_cigar = f.n_cigar;
    c.flag = (uint16_t)f.flag;
    c.l_qname = (uint16_t)f.l_read_name;
    c.mapq = (uint8_t)f.mapq;
    core[j] = c;
    sample[j] = files[sel_file[order[j]]].sample;
    rg[j] = files[sel_file[order[j]]].rg;
    data.insert(data.end(), rec + 36, rec + 4 + f.block_size);
    data_off[j + 1] = data.size();
  }
  return 0;
}
} // namespace gtb
