// gtb_bam.cu -- record parsing on the device (product code; SURVEY.md section 8f, N3, first step).
//
// Input: the records of one pool in merge order as htslib holds them (gtb_bam_batch: core fields + bam1_t::data).
// Output: the record columns of a chunk (gtb_device.cuh: DevBatch) exactly as the reference's per-record code derives them:
//   bam_parse_kernel   1 thread / record: lengths, flags, MAPQ, insert size, same-contig flag, AS-XS by get_score_diff's walk of
//                      the aux block (src/typer/alignment.cpp:140-325), 64-bit hash of (read group, read name)
//   bam_seq_kernel     1 warp / record: the 4-bit bases into the 76-byte-stride seq4 column
//   bam_dup_kernel     1 thread / record: equal_pos_seq against the previous record (hts_utils.hpp:110-128); equality is
//                      transitive, so "equals the last non-duplicate" (hts_parallel_reader.cpp:666-684) = "equals its
//                      predecessor", and the batch-preparation kernels resolve the chains to their roots
//   cub::DeviceRadixSort (stable) of (hash, record index), then
//   bam_mate_kernel    1 thread / run of equal hashes: genotype_only's read-name map (hts_parallel_reader.cpp:270-337) replayed
//                      over the run in record order -- a record whose name waits pairs with it (paired flag or not), else a
//                      paired record waits, an unpaired one stays alone; names are compared byte by byte, and a run that holds several
//                      distinct names (a 64-bit hash collision) is replayed once per name; SV graphs: what still waits is a leftover mate (:719-772)
#include <cstdint>

#include <cub/device/device_radix_sort.cuh>

#include "gtb_device.cuh"

namespace gtb
{
namespace
{
// get_score_diff (alignment.cpp:140-325)
__device__ uint8_t score_diff_of(const uint8_t * it, long l_aux)
{
  long i = 0;
  long long as = -1, xs = -1;
  while (i < l_aux)
  {
    i += 3;
    if (i > l_aux)
      break; // truncated tag header: the reference would read past the record; nothing of value there
    uint8_t const type = it[i - 1];
    bool const s_tag = it[i - 2] == 'S';
    bool const is_as = s_tag && it[i - 3] == 'A', is_xs = s_tag && it[i - 3] == 'X';
    long long num = 0;
    bool have = false;
    switch (type)
    {
    case 'A':
      ++i;
      break;
    case 'Z':
      while (i < l_aux && it[i] != '\0' && it[i] != '\n')
        ++i;
      ++i;
      break;
    case 'c':
      num = (int8_t)it[i];
      have = true;
      i += 1;
      break;
    case 'C':
      num = it[i];
      have = true;
      i += 1;
      break;
    case 's':
      num = (int16_t)((uint16_t)it[i] | ((uint16_t)it[i + 1] << 8));
      have = true;
      i += 2;
      break;
    case 'S':
      num = (uint16_t)((uint16_t)it[i] | ((uint16_t)it[i + 1] << 8));
      have = true;
      i += 2;
      break;
    case 'i':
      num = (int32_t)((uint32_t)it[i] | ((uint32_t)it[i + 1] << 8) | ((uint32_t)it[i + 2] << 16) | ((uint32_t)it[i + 3] << 24));
      have = true;
      i += 4;
      break;
    case 'I':
      num = (uint32_t)((uint32_t)it[i] | ((uint32_t)it[i + 1] << 8) | ((uint32_t)it[i + 2] << 16) | ((uint32_t)it[i + 3] << 24));
      have = true;
      i += 4;
      break;
    case 'f':
      i += 4;
      break;
    default:
      i = l_aux; // unknown type: the walk stops
      break;
    }
    if (have && is_as)
      as = num;
    else if (have && is_xs)
      xs = num;
  }
  if (as == -1 || as < xs)
    return 0;
  if (xs == -1)
    xs = 0;
  long long const diff = as - xs;
  return diff < 255 ? (uint8_t)diff : (uint8_t)255;
}

// region of record k (records of a region are contiguous) and the bounds of its data block
struct RecLoc
{
  uint32_t reg;
  unsigned long long o0, o1;
};
__device__ __forceinline__ RecLoc locate(const BamParams & p, uint32_t k)
{
  uint32_t lo = 0, hi = p.n_regions; // last region with rec_begin <= k
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (p.rec_begin[mid] <= k)
      lo = mid;
    else
      hi = mid;
  }
  const unsigned long long * off = p.data_off + p.rec_begin[lo] + lo + (k - p.rec_begin[lo]);
  unsigned long long const base = p.data_base[lo];
  return RecLoc{lo, base + off[0], base + off[1]};
}

__device__ __forceinline__ bool record_ok(const gtb_bam_core & c, unsigned long long l_data)
{
  if (c.l_qseq < 0 || c.l_qname == 0)
    return false;
  unsigned long long const need = (unsigned long long)c.l_qname + 4ull * c.n_cigar + (unsigned long long)(c.l_qseq + 1) / 2 + (unsigned long long)c.l_qseq;
  return need <= l_data;
}

__global__ void __launch_bounds__(256) bam_parse_kernel(BamParams p)
{
  uint32_t const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n)
    return;
  gtb_bam_core const c = p.core[k];
  RecLoc const loc = locate(p, k);
  unsigned long long const o0 = loc.o0, o1 = loc.o1;
  bool ok = o1 >= o0 && record_ok(c, o1 - o0);
  if (ok && c.l_qseq > MAX_SEQ)
  {
    atomicOr(&p.counters->input_bits, PREP_ERR_LEN);
    ok = false;
  }
  else if (!ok)
    atomicOr(&p.counters->input_bits, PREP_ERR_RECORD);
  p.region[k] = p.slots[loc.reg];
  p.clipped[k] = 0; // clipped_count() returns a bool, so "> 3" is never true in the reference (alignment.cpp:105-138)
  p.leftover[k] = 0;
  p.mate[k] = -1;
  p.idx[k] = k;
  if (!ok)
  {
    // neutral record: never aligned, never paired
    p.lseq[k] = 0;
    p.flag[k] = 0;
    p.mapq[k] = 0;
    p.isize[k] = 0;
    p.same_tid[k] = 0;
    p.score_diff[k] = 0;
    p.name_hash[k] = 0xFFFFFFFFFFFFFFFFull - k; // a run of its own
    return;
  }
  const uint8_t * d = p.data + o0;
  p.lseq[k] = (uint16_t)c.l_qseq;
  p.flag[k] = c.flag;
  p.mapq[k] = c.mapq;
  long long const isz = c.isize;
  p.isize[k] = (int32_t)(isz > 2147483647ll ? 2147483647ll : isz < -2147483648ll ? -2147483648ll : isz);
  p.same_tid[k] = c.tid == c.mtid;
  unsigned long long const o_aux = (unsigned long long)c.l_qname + 4ull * c.n_cigar + (unsigned long long)(c.l_qseq + 1) / 2 + (unsigned long long)c.l_qseq;
  p.score_diff[k] = score_diff_of(d + o_aux, (long)((o1 - o0) - o_aux));
  // FNV-1a over (region, read group, name up to the NUL), finished with a 64-bit mix
  unsigned long long h = 0xCBF29CE484222325ull ^ (unsigned long long)(uint32_t)p.rg[k] ^ ((unsigned long long)loc.reg << 32);
  h *= 0x100000001B3ull;
  for (uint32_t j = 0; j < c.l_qname && d[j] != 0; ++j)
  {
    h ^= d[j];
    h *= 0x100000001B3ull;
  }
  h ^= h >> 33;
  h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33;
  p.name_hash[k] = h & p.hash_mask; // all ones; tests narrow it (GTB_BAM_HASH_BITS) to force colliding names into one run
}

__global__ void __launch_bounds__(256) bam_seq_kernel(BamParams p)
{
  uint32_t const k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int const lane = threadIdx.x & 31;
  if (k >= p.n)
    return;
  uint32_t const L = p.lseq[k]; // 0 for a rejected record
  uint32_t const nb = (L + 1) / 2;
  const uint8_t * src = p.data + locate(p, k).o0 + p.core[k].l_qname + 4ull * p.core[k].n_cigar;
  uint8_t * dst = p.seq4 + (size_t)k * GTB_SEQ_STRIDE;
  for (uint32_t j = lane; j < GTB_SEQ_STRIDE; j += 32)
    dst[j] = j < nb ? src[j] : (uint8_t)0;
}

__global__ void __launch_bounds__(256) bam_dup_kernel(BamParams p)
{
  uint32_t const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n)
    return;
  int32_t dup = -1;
  if (k > 0 && p.lseq[k] != 0)
  {
    gtb_bam_core const a = p.core[k - 1], b = p.core[k];
    if (a.tid == b.tid && a.pos == b.pos && a.l_qseq == b.l_qseq && p.lseq[k - 1] != 0 && p.region[k - 1] == p.region[k] &&
        locate(p, k).reg == locate(p, k - 1).reg) // the shortcut never crosses a pool
    {
      // rows are zero-padded to the stride, so whole-row equality = equality of the (l_qseq + 1) / 2 sequence bytes;
      // a row is 76 bytes = 19 words and starts 4-byte aligned
      const uint32_t * x = reinterpret_cast<const uint32_t *>(p.seq4 + (size_t)(k - 1) * GTB_SEQ_STRIDE);
      const uint32_t * y = reinterpret_cast<const uint32_t *>(p.seq4 + (size_t)k * GTB_SEQ_STRIDE);
      bool same = true;
      for (int j = 0; j < (int)(GTB_SEQ_STRIDE / 4) && same; ++j)
        same = x[j] == y[j];
      if (same)
        dup = (int32_t)(k - 1);
    }
  }
  p.dup_of[k] = dup;
}

__global__ void __launch_bounds__(128) bam_mate_kernel(BamParams p)
{
  uint32_t const q0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (q0 >= p.n)
    return;
  unsigned long long const key = p.name_hash_sorted[q0];
  if (q0 > 0 && p.name_hash_sorted[q0 - 1] == key)
    return; // not the head of its run
  uint32_t q1 = q0 + 1;
  while (q1 < p.n && p.name_hash_sorted[q1] == key)
    ++q1;
  // same read group, same pool and the same name, byte by byte
  auto same_name = [&](uint32_t ra, uint32_t rb) {
    RecLoc const la = locate(p, ra), lb = locate(p, rb);
    if (p.rg[ra] != p.rg[rb] || la.reg != lb.reg)
      return false;
    const uint8_t * na = p.data + la.o0;
    const uint8_t * nb = p.data + lb.o0;
    for (uint32_t j = 0;; ++j)
    {
      if (na[j] != nb[j])
        return false;
      if (na[j] == 0)
        return true;
    }
  };
  // A run of equal hashes is one read name -- or, once in ~2^64 / n^2 pools, several names that collide.  Every distinct name
  // of the run gets its own replay of the read-name map (a run has 2 records almost always, so the quadratic scan is nothing).
  for (uint32_t g = q0; g < q1; ++g)
  {
    uint32_t const head = p.idx_sorted[g]; // ascending within the run: the sort is stable
    bool first_of_its_name = true;
    for (uint32_t e = q0; e < g && first_of_its_name; ++e)
      first_of_its_name = !same_name(p.idx_sorted[e], head);
    if (!first_of_its_name)
      continue;
    int32_t waiting = -1;
    for (uint32_t q = g; q < q1; ++q)
    {
      uint32_t const r = p.idx_sorted[q];
      if (q != g && !same_name(head, r))
        continue;
      if (waiting >= 0)
      {
        p.mate[r] = waiting;
        waiting = -1;
      }
      else if (p.flag[r] & 1u)
        waiting = (int32_t)r;
    }
    if (waiting >= 0 && p.regions[p.slots[locate(p, head).reg]].is_sv)
      p.leftover[waiting] = 1;
  }
}
} // namespace

size_t bam_sort_temp_bytes(uint32_t n)
{
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  return bytes;
}

int launch_bam_parse(const BamParams & p, void * sort_temp, size_t sort_temp_bytes, void * stream)
{
  if (p.n == 0)
    return 0;
  cudaStream_t const s = (cudaStream_t)stream;
  bam_parse_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  bam_seq_kernel<<<(p.n + 7) / 8, 256, 0, s>>>(p);
  bam_dup_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  if (cub::DeviceRadixSort::SortPairs(sort_temp, sort_temp_bytes, p.name_hash, p.name_hash_sorted, p.idx, p.idx_sorted, (int)p.n, 0, 64,
                                      s) != cudaSuccess)
    return -1;
  bam_mate_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
  return 0;
}
} // namespace gtb
