// gtb_index_dev.cu -- device-side k-mer index construction: replaces
// PHIndex index_graph(Graph const&) (reference: src/index/indexer.cpp:246-291, 26-81, 83-178, 213-244) for all regions of
// one gtb_region_begin_multi call in ONE launch sequence.
//
// Formulation (the same set property as the host builder, gtb_index_host.hpp): the index holds one entry per 32-base
// walk through the bubble graph that is pure A/C/G/T, respects the allele-combination limit (indexer.cpp:13-20,198-211)
// and the anti-event rule (indexer.cpp:114-121); the labels of one k-mer are ordered by END position in sweep order and,
// within one END position, by a depth-first backward enumeration with ascending allele loops.  Every END position is an
// independent job:
//   1. idx_enum_kernel<false>   one thread per END position (all regions): number of labels it emits
//   2. exclusive scan           label offsets = the reference's insertion order, region-major
//   3. idx_enum_kernel<true>    same walk, writes (k-mer, label) at its offset
//   4. stable radix sorts       by k-mer, then by region (cub::DeviceRadixSort, LSD) -- equal k-mers of a region keep
//                               their emission order, which is the bucket order the aligner depends on
//   5. idx_group_kernel         run heads -> distinct k-mer slots {key, first label, count}, labels gathered in order
//   6. idx_table_kernel         CAS-insert of the distinct k-mers into each region's open-addressing table + bitmap
// Integer/byte work only; bound by launch latency and the sort at these sizes (73 k labels per 50 kb region).

#include <cstdint>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "gtb_device.cuh"

namespace gtb
{
__device__ uint32_t g_hash_tab_idx[8 * 256]; // this translation unit's copy of HashTables::tab (gtb_device.cuh)

int upload_hash_tables_index()
{
  return (int)cudaMemcpyToSymbol(g_hash_tab_idx, hash_tables().tab, sizeof(uint32_t) * 8 * 256);
}

namespace
{
constexpr int IDX_MAX_FRAMES = 64;
constexpr uint32_t IDX_ERR_SPECIAL = 1, IDX_ERR_FRAMES = 2;

__device__ __forceinline__ int code_of(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

struct Walk // state of the backward enumeration (gtb_index_host.hpp: IndexBuilder::Walk)
{
  uint64_t key;
  uint64_t alt_prod;  // product of the allele counts of the bubbles entered through a non-reference allele
  int depth;          // bases collected so far (from the end)
  int nvars;
  uint32_t alt_count; // non-reference alleles on the walk so far
  uint32_t vars[32];  // var nodes touched, in backward order
  uint8_t vbases[32]; // bases of that var node on the walk (>= 2 enables the self anti-event corner)
};

__device__ __forceinline__ uint32_t ref_len(const IdxRegion & G, uint32_t r) { return G.ref_seq_off[r + 1] - G.ref_seq_off[r]; }
__device__ __forceinline__ uint32_t var_len(const IdxRegion & G, uint32_t v) { return G.var_seq_off[v + 1] - G.var_seq_off[v]; }

// consume bases dna[hi], dna[hi-1], ... dna[0] (as many as still needed); false on a non-ACGT base
__device__ __forceinline__ bool take_backward(Walk & w, const uint8_t * dna, int hi, int & taken)
{
  taken = 0;
  for (int i = hi; i >= 0 && w.depth < 32; --i)
  {
    int const c = code_of(dna[i]);
    if (c < 0)
      return false;
    w.key |= (uint64_t)c << (2 * w.depth);
    ++w.depth;
    ++taken;
  }
  return true;
}

__device__ uint32_t special_pos(const IdxRegion & G, uint32_t pos, uint32_t ref_reach, uint32_t * err)
{
  uint32_t lo = 0, hi = G.n_sp_keys;
  while (lo < hi)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (G.sp_keys[mid] < ref_reach)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo >= G.n_sp_keys || G.sp_keys[lo] != ref_reach)
  {
    atomicOr(err, IDX_ERR_SPECIAL);
    return 0xFFFFFFFFu;
  }
  uint32_t const idx = pos - ref_reach - 1;
  if (G.sp_off[lo] + idx >= G.sp_off[lo + 1])
  {
    atomicOr(err, IDX_ERR_SPECIAL);
    return 0xFFFFFFFFu;
  }
  return G.sp_list[G.sp_off[lo] + idx];
}

// position of base d of var node v; beyond the reference allele's reach it is special-position encoded (indexer.cpp:146-147)
__device__ uint32_t encode_var_pos(const IdxRegion & G, uint32_t v, uint32_t d, uint32_t * err)
{
  uint32_t pos = G.var_order[v] + d;
  uint32_t const v0 = G.ref_var_off[G.var_out_ref[v] - 1];
  uint32_t const rr = G.var_order[v0] + var_len(G, v0) - 1;
  if (pos > rr)
    pos = special_pos(G, pos, rr, err);
  return pos;
}

__device__ bool has_event(const int64_t * b, const int64_t * e, int64_t x)
{
  while (b < e)
  {
    const int64_t * mid = b + (e - b) / 2;
    if (*mid < x)
      b = mid + 1;
    else if (*mid > x)
      e = mid;
    else
      return true;
  }
  return false;
}

// a walk may not enter a var node whose events intersect the anti_events of an earlier var node (indexer.cpp:114-121);
// a node's own anti_events only apply from its second base on
__device__ bool events_ok(const IdxRegion & G, const Walk & w)
{
  if (!G.var_ev || !G.var_aev)
    return true;
  for (int i = w.nvars - 1; i >= 0; --i)
  {
    uint32_t const v = w.vars[i];
    const int64_t * eb = G.var_ev + G.var_ev_off[v];
    const int64_t * ee = G.var_ev + G.var_ev_off[v + 1];
    if (eb == ee)
      continue;
    for (int m = w.nvars - 1; m >= i; --m)
    {
      if (m == i && w.vbases[i] < 2)
        continue;
      uint32_t const u = w.vars[m];
      for (uint32_t a = G.var_aev_off[u]; a < G.var_aev_off[u + 1]; ++a)
        if (has_event(eb, ee, G.var_aev[a]))
          return false;
    }
  }
  return true;
}

template <bool EMIT>
__device__ __forceinline__ void emit(const IdxRegion & G, const Walk & w, uint32_t start_pos, uint32_t end_pos, uint32_t & n,
                                     uint64_t * keys, DevLabel * labels, uint32_t out_base)
{
  if (!events_ok(G, w))
    return;
  if (w.nvars == 0)
  {
    if (EMIT)
    {
      keys[out_base + n] = w.key;
      labels[out_base + n] = DevLabel{start_pos, end_pos, 0xFFFFFFFFu};
    }
    ++n;
    return;
  }
  for (int i = w.nvars - 1; i >= 0; --i) // ascending var id = forward order
  {
    if (EMIT)
    {
      keys[out_base + n] = w.key;
      labels[out_base + n] = DevLabel{start_pos, end_pos, w.vars[i]};
    }
    ++n;
  }
}

// All walks ending at base d of ref node `node` (in_var = false) or var node `node` (in_var = true), in the reference's
// emission order.  Returns the number of labels; EMIT writes them at out_base.
template <bool EMIT>
__device__ uint32_t enum_job(const IdxRegion & G, bool in_var, uint32_t node, uint32_t d, uint64_t * keys, DevLabel * labels,
                             uint32_t out_base, uint32_t * err)
{
  uint32_t n = 0;
  Walk w;
  w.key = 0;
  w.alt_prod = 1;
  w.depth = 0;
  w.nvars = 0;
  w.alt_count = 0;
  int taken;
  uint32_t end_pos;
  uint32_t b; // bubble to walk back into next
  bool into_ref; // the walk continues into ref node b first (var start), else straight into bubble b
  if (!in_var)
  {
    end_pos = G.ref_order[node] + d;
    if (!take_backward(w, G.seq + G.ref_seq_off[node], (int)d, taken))
      return 0;
    if (w.depth == 32)
    {
      emit<EMIT>(G, w, G.ref_order[node] + d + 1 - (uint32_t)taken, end_pos, n, keys, labels, out_base);
      return n;
    }
    if (node == 0)
      return 0;
    b = node - 1;
    into_ref = false;
  }
  else
  {
    uint32_t const v = node;
    b = G.var_out_ref[v] - 1;
    uint32_t const vb = G.ref_var_off[b];
    end_pos = encode_var_pos(G, v, d, err);
    if (v != vb)
    {
      w.alt_count = 1;
      w.alt_prod = G.ref_var_off[b + 1] - vb;
    }
    if (!take_backward(w, G.seq + G.var_seq_off[v], (int)d, taken))
      return 0;
    w.vars[0] = v;
    w.vbases[0] = (uint8_t)min(taken, 255);
    w.nvars = 1;
    if (w.depth == 32)
    {
      emit<EMIT>(G, w, encode_var_pos(G, v, d + 1 - (uint32_t)taken, err), end_pos, n, keys, labels, out_base);
      return n;
    }
    into_ref = true;
  }

  // explicit stack of the bubbles being iterated: the walk state when the bubble was entered + the next allele to try
  uint64_t f_key[IDX_MAX_FRAMES], f_prod[IDX_MAX_FRAMES];
  uint32_t f_next[IDX_MAX_FRAMES], f_bubble[IDX_MAX_FRAMES];
  uint8_t f_depth[IDX_MAX_FRAMES], f_nvars[IDX_MAX_FRAMES], f_altc[IDX_MAX_FRAMES];
  int sp = 0;

  // continue backwards from the END of ref node r; pushes the bubble before it unless the walk completes or dies
  auto enter_ref = [&](uint32_t r) {
    int const len = (int)ref_len(G, r);
    int tk;
    if (!take_backward(w, G.seq + G.ref_seq_off[r], len - 1, tk))
      return;
    if (w.depth == 32)
    {
      emit<EMIT>(G, w, G.ref_order[r] + (uint32_t)(len - tk), end_pos, n, keys, labels, out_base);
      return;
    }
    if (r == 0)
      return; // ran out of graph
    if (sp >= IDX_MAX_FRAMES)
    {
      atomicOr(err, IDX_ERR_FRAMES);
      return;
    }
    f_key[sp] = w.key;
    f_prod[sp] = w.alt_prod;
    f_depth[sp] = (uint8_t)w.depth;
    f_nvars[sp] = (uint8_t)w.nvars;
    f_altc[sp] = (uint8_t)w.alt_count;
    f_bubble[sp] = r - 1;
    f_next[sp] = G.ref_var_off[r - 1];
    ++sp;
  };

  if (into_ref)
    enter_ref(b);
  else
  {
    f_key[0] = w.key;
    f_prod[0] = w.alt_prod;
    f_depth[0] = (uint8_t)w.depth;
    f_nvars[0] = (uint8_t)w.nvars;
    f_altc[0] = (uint8_t)w.alt_count;
    f_bubble[0] = b;
    f_next[0] = G.ref_var_off[b];
    sp = 1;
  }

  while (sp > 0)
  {
    int const t = sp - 1;
    uint32_t const bb = f_bubble[t];
    uint32_t const vb = G.ref_var_off[bb], ve = G.ref_var_off[bb + 1];
    uint32_t const v = f_next[t];
    if (v >= ve)
    {
      --sp;
      continue;
    }
    f_next[t] = v + 1;
    w.key = f_key[t];
    w.alt_prod = f_prod[t];
    w.depth = f_depth[t];
    w.nvars = f_nvars[t];
    w.alt_count = f_altc[t];
    if (v != vb)
    {
      ++w.alt_count;
      w.alt_prod *= (uint64_t)(ve - vb);
      if (w.alt_prod > 0xFFFFFFFFull)
        w.alt_prod = 0xFFFFFFFFull;
      if (w.alt_count > 1 && (w.alt_prod > 181 || w.alt_count > 4))
        continue;
    }
    int const m = (int)var_len(G, v);
    int tk;
    if (!take_backward(w, G.seq + G.var_seq_off[v], m - 1, tk))
      continue;
    if (tk > 0)
    {
      if (w.nvars >= 32)
        continue; // unreachable: every recorded var node contributed a base
      w.vars[w.nvars] = v;
      w.vbases[w.nvars] = (uint8_t)min(tk, 255);
      ++w.nvars;
    }
    if (w.depth == 32)
    {
      emit<EMIT>(G, w, encode_var_pos(G, v, (uint32_t)(m - tk), err), end_pos, n, keys, labels, out_base);
      continue;
    }
    enter_ref(bb);
  }
  return n;
}

__device__ __forceinline__ uint32_t upper_region(const uint32_t * off, uint32_t n, uint32_t x) // last r with off[r] <= x
{
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (off[mid] <= x)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) idx_enum_kernel(const IdxRegion * regions, uint32_t n_regions, const uint32_t * region_job_off,
                                                       uint32_t total_jobs, uint32_t * job_cnt, const uint32_t * job_off,
                                                       uint64_t * keys, DevLabel * labels, uint32_t * tuple_idx, uint32_t * err)
{
  uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total_jobs)
    return;
  uint32_t const r = upper_region(region_job_off, n_regions, j);
  const IdxRegion & G = regions[r];
  uint32_t const lj = j - region_job_off[r];
  uint32_t const k = upper_region(G.sweep_job_off, G.n_sweep, lj);
  uint32_t const d = lj - G.sweep_job_off[k];
  uint32_t const node = G.sweep_node[k];
  bool const in_var = (node >> 31) != 0;
  uint32_t const id = node & 0x7FFFFFFFu;
  uint32_t const base = EMIT ? job_off[j] : 0u;
  uint32_t n;
  if (!in_var && d >= 31)
  {
    // the window lies inside this ref node: the only walk ending here
    const uint8_t * dna = G.seq + G.ref_seq_off[id] + d - 31;
    uint64_t key = 0;
    bool ok = true;
#pragma unroll 8
    for (int i = 0; i < 32; ++i)
    {
      int const c = code_of(dna[i]);
      ok = ok && c >= 0;
      key = (key << 2) | (uint64_t)(c & 3);
    }
    n = ok ? 1u : 0u;
    if (EMIT && ok)
    {
      keys[base] = key;
      labels[base] = DevLabel{G.ref_order[id] + d - 31, G.ref_order[id] + d, 0xFFFFFFFFu};
    }
  }
  else
    n = enum_job<EMIT>(G, in_var, id, d, keys, labels, base, err);
  if (!EMIT)
    job_cnt[j] = n;
  else
    for (uint32_t i = 0; i < n; ++i)
      tuple_idx[base + i] = base + i;
}

// region_tuple_off[r] = first label of region r in the global emission order (r = n_regions: total)
__global__ void idx_region_totals_kernel(const uint32_t * job_off, const uint32_t * region_job_off, uint32_t n_regions,
                                         uint32_t * region_tuple_off)
{
  uint32_t const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_regions)
    region_tuple_off[r] = job_off[region_job_off[r]];
}

// sorted position i -> 1 if it starts a new k-mer within its region
__global__ void __launch_bounds__(256) idx_heads_kernel(const uint64_t * skeys, const uint32_t * region_tuple_off, uint32_t n_regions,
                                                        uint32_t total, uint32_t * head)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total)
    return;
  uint32_t const r = upper_region(region_tuple_off, n_regions, i);
  head[i] = (i == region_tuple_off[r] || skeys[i] != skeys[i - 1]) ? 1u : 0u;
}

// gathers the labels into each region's arena in sorted order and fills the distinct k-mer slots:
// uniq[u] = {key, first label, END label (exclusive; turned into a count by idx_table_kernel)}
__global__ void __launch_bounds__(256) idx_group_kernel(const IdxRegion * regions, const uint64_t * skeys, const uint32_t * sidx,
                                                        const DevLabel * labels_emit, const uint32_t * head_incl,
                                                        const uint32_t * region_tuple_off, uint32_t n_regions, uint32_t total,
                                                        uint32_t * region_n_uniq)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total)
    return;
  uint32_t const r = upper_region(region_tuple_off, n_regions, i);
  const IdxRegion & G = regions[r];
  uint32_t const first = region_tuple_off[r], last = region_tuple_off[r + 1];
  uint32_t const li = i - first;
  uint32_t const heads_before = first ? head_incl[first - 1] : 0u;
  uint32_t const u = head_incl[i] - heads_before - 1; // distinct k-mer index within the region (keys ascending)
  G.labels[li] = labels_emit[sidx[i]];
  uint64_t const key = skeys[i];
  if (i == first || skeys[i - 1] != key)
  {
    G.uniq[u].key = key;
    G.uniq[u].off = li;
  }
  if (i + 1 == last || skeys[i + 1] != key)
    G.uniq[u].cnt = li + 1;
  if (i + 1 == last)
    region_n_uniq[r] = u + 1;
}

// one thread per (region, distinct k-mer): finish the slot and insert it into the region's table (see build_table_kernel)
__global__ void __launch_bounds__(256) idx_table_kernel(const IdxRegion * regions, const uint32_t * head_incl,
                                                        const uint32_t * region_tuple_off, uint32_t n_regions, uint32_t total)
{
  uint32_t const g = blockIdx.x * blockDim.x + threadIdx.x; // global distinct k-mer number
  if (total == 0 || g >= head_incl[total - 1])
    return;
  // region of the g-th head: heads_before[r] <= g
  uint32_t lo = 0, hi = n_regions;
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    uint32_t const f = region_tuple_off[mid];
    uint32_t const hb = f ? head_incl[f - 1] : 0u;
    if (hb <= g)
      lo = mid;
    else
      hi = mid;
  }
  // (regions without labels share their successor's head count; "last region with heads_before <= g" skips them)
  const IdxRegion & G = regions[lo];
  uint32_t const f = region_tuple_off[lo];
  uint32_t const u = g - (f ? head_incl[f - 1] : 0u);
  IndexSlot s = G.uniq[u];
  s.cnt -= s.off;
  G.uniq[u].cnt = s.cnt;
  unsigned long long const val = (unsigned long long)s.off | ((unsigned long long)s.cnt << 32);
  uint32_t const hs = hash32_tab(g_hash_tab_idx, s.key);
  uint32_t const bi = hs >> (G.table_shift - 34); // presence bitmap: 4 bits per table slot
  atomicOr(&G.bitmap[bi >> 5], 1u << (bi & 31u));
  uint32_t h = hs >> (G.table_shift - 32);
  while (true)
  {
    unsigned long long * w = reinterpret_cast<unsigned long long *>(&G.table[h]) + 1;
    if (atomicCAS(w, 0ull, val) == 0ull)
    {
      G.table[h].key = s.key;
      return;
    }
    h = (h + 1) & G.table_mask;
  }
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host launchers
void idx_launch_count(const IdxRegion * regions, uint32_t n_regions, const uint32_t * region_job_off, uint32_t total_jobs,
                      uint32_t * job_cnt, uint32_t * err, void * stream)
{
  if (total_jobs == 0)
    return;
  idx_enum_kernel<false><<<(total_jobs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(regions, n_regions, region_job_off, total_jobs,
                                                                                  job_cnt, nullptr, nullptr, nullptr, nullptr, err);
}

size_t idx_scan_temp_bytes(uint32_t n)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  size_t b2 = 0;
  cub::DeviceScan::InclusiveSum(nullptr, b2, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  return bytes > b2 ? bytes : b2;
}

int idx_exclusive_scan(void * temp, size_t temp_bytes, const uint32_t * in, uint32_t * out, uint32_t n, void * stream)
{
  return (int)cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, (cudaStream_t)stream);
}

size_t scan64_temp_bytes(uint32_t n)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
  return bytes;
}

int exclusive_scan64(void * temp, size_t temp_bytes, const unsigned long long * in, unsigned long long * out, uint32_t n,
                     void * stream)
{
  return (int)cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, (cudaStream_t)stream);
}

int idx_inclusive_scan(void * temp, size_t temp_bytes, const uint32_t * in, uint32_t * out, uint32_t n, void * stream)
{
  return (int)cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, (int)n, (cudaStream_t)stream);
}

void idx_launch_region_totals(const uint32_t * job_off, const uint32_t * region_job_off, uint32_t n_regions,
                              uint32_t * region_tuple_off, void * stream)
{
  idx_region_totals_kernel<<<(n_regions + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(job_off, region_job_off, n_regions,
                                                                                       region_tuple_off);
}

void idx_launch_emit(const IdxRegion * regions, uint32_t n_regions, const uint32_t * region_job_off, uint32_t total_jobs,
                     const uint32_t * job_off, uint64_t * keys, DevLabel * labels, uint32_t * tuple_idx, uint32_t * err, void * stream)
{
  if (total_jobs == 0)
    return;
  idx_enum_kernel<true><<<(total_jobs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(regions, n_regions, region_job_off, total_jobs,
                                                                                 nullptr, job_off, keys, labels, tuple_idx, err);
}

// Stable sort of the emitted (k-mer, label index) pairs by k-mer WITHIN each region, as two LSD radix sorts: all pairs by
// the 64-bit k-mer (8 onesweep passes), then -- for more than one region -- by region number (1-2 passes); both sorts
// are stable, so equal k-mers of a region keep their emission order.  (cub::DeviceSegmentedSort was measured first: with
// 20 segments of 73 k pairs it falls back to one block per segment, 1.3 ms per call against 0.25 ms for this.)
namespace
{
__global__ void __launch_bounds__(256) idx_region_ids_kernel(const uint32_t * sidx, const uint32_t * region_tuple_off, uint32_t n_regions,
                                                             uint32_t total, uint32_t * rid, uint32_t * pos)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total)
    return;
  rid[i] = upper_region(region_tuple_off, n_regions, sidx[i]); // emission index -> region (emission order is region-major)
  pos[i] = i;
}

__global__ void __launch_bounds__(256) idx_gather_kernel(const uint64_t * k_in, const uint32_t * i_in, const uint32_t * pos, uint32_t total,
                                                         uint64_t * k_out, uint32_t * i_out)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total)
    return;
  uint32_t const p = pos[i];
  k_out[i] = k_in[p];
  i_out[i] = i_in[p];
}

int region_bits(uint32_t n_regions)
{
  int bits = 1;
  while ((1u << bits) < n_regions)
    ++bits;
  return bits;
}
} // namespace

size_t idx_sort_temp_bytes(uint32_t total, uint32_t n_regions)
{
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int)total, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int)total, 0, region_bits(n_regions));
  return a > b ? a : b;
}

// In: keys_a / idx_a in emission order.  Out: sorted pairs in keys_b / idx_b when the function returns 0 (b buffers), the
// a buffers and aux (4 x total u32) are scratch.
int idx_sort(void * temp, size_t temp_bytes, uint64_t * keys_a, uint64_t * keys_b, uint32_t * idx_a, uint32_t * idx_b, uint32_t * aux,
             uint32_t total, uint32_t n_regions, const uint32_t * region_tuple_off, void * stream)
{
  cudaStream_t const st = (cudaStream_t)stream;
  if (n_regions <= 1)
    return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, idx_a, idx_b, (int)total, 0, 64, st);
  // 1. by k-mer: a -> b
  if (int rc = (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, idx_a, idx_b, (int)total, 0, 64, st))
    return rc;
  // 2. by region: (rid, pos) -> (rid2, pos2)
  uint32_t * rid = aux;
  uint32_t * pos = aux + total;
  uint32_t * rid2 = aux + 2 * (size_t)total;
  uint32_t * pos2 = aux + 3 * (size_t)total;
  idx_region_ids_kernel<<<(total + 255) / 256, 256, 0, st>>>(idx_b, region_tuple_off, n_regions, total, rid, pos);
  if (int rc = (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, rid, rid2, pos, pos2, (int)total, 0, region_bits(n_regions), st))
    return rc;
  // 3. apply the permutation: b -> a, then hand a back as b
  idx_gather_kernel<<<(total + 255) / 256, 256, 0, st>>>(keys_b, idx_b, pos2, total, keys_a, idx_a);
  cudaMemcpyAsync(keys_b, keys_a, (size_t)total * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(idx_b, idx_a, (size_t)total * 4, cudaMemcpyDeviceToDevice, st);
  return 0;
}

void idx_launch_group(const IdxRegion * regions, const uint64_t * skeys, const uint32_t * sidx, const DevLabel * labels_emit,
                      uint32_t * head, uint32_t * head_incl, void * scan_temp, size_t scan_temp_bytes, const uint32_t * region_tuple_off,
                      uint32_t n_regions, uint32_t total, uint32_t * region_n_uniq, void * stream)
{
  if (total == 0)
    return;
  cudaStream_t const st = (cudaStream_t)stream;
  idx_heads_kernel<<<(total + 255) / 256, 256, 0, st>>>(skeys, region_tuple_off, n_regions, total, head);
  cub::DeviceScan::InclusiveSum(scan_temp, scan_temp_bytes, head, head_incl, (int)total, st);
  idx_group_kernel<<<(total + 255) / 256, 256, 0, st>>>(regions, skeys, sidx, labels_emit, head_incl, region_tuple_off, n_regions, total,
                                                        region_n_uniq);
  idx_table_kernel<<<(total + 255) / 256, 256, 0, st>>>(regions, head_incl, region_tuple_off, n_regions, total);
}

} // namespace gtb
