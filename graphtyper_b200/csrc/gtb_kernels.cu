// gtb_kernels.cu -- sm_100a kernels of the read -> graph genotyping path (product code).
//
// One launch sequence per chunk of a submit:
//   prep_flags_kernel / prep_fill_kernel (+ a 64-bit scan): alignment units of the duplicate-read shortcut, the list of read
//       orientations align_read aligns at all (alignment.cpp:343-360), link checks -- the per-record part of the pool loop
//       (hts_parallel_reader.cpp:655-708).
//   probe_kernel : phase A, one persistent block per SM, one warp per (alignment unit, read orientation) at a time: the
//       32-mer seed keys and their GF(2)-linear hashes (warp reduces), the 96 Hamming-1 neighbours per seed against the
//       region's presence filter in SHARED memory, the few keys that pass against the open-addressed k-mer table
//       (replaces to_uint64_vec / query_index / PHIndex::multi_get, src/utilities/type_conversions.cpp:207-288,
//       src/utilities/kmer_help_functions.cpp:51-119, src/index/ph_index.cpp:66-107).
//   chain_kernel (one thread per task, working set in local memory) / slow_kernel (one warp per task, shared memory) /
//   huge_kernel (one warp per task, global slab) -- three capacity tiers of the same templated code:
//       B. chain the seed labels into paths (GenotypePaths::add_next_kmer_labels, src/typer/genotype_paths.cpp:294-352,
//          Path merge src/typer/path.cpp:38-82);
//       C. extend both read ends through the bubble graph under the shrinking mismatch budget
//          (walk_read_starts/ends genotype_paths.cpp:483-621, Graph::get_locations_of_a_position graph.cpp:931-1185,
//          get_labels_forward/backward graph.cpp:1187-1701, count_mismatches graph_utils.hpp:7-69);
//       D. apply the path filters (alignment.cpp:68-87) and write a compact GenotypePaths record.
//   score_kernel / score_deferred_kernel : one thread per record: mate/orientation selection (get_better_paths
//       alignment.cpp:557-622, compare_pair_of_genotype_paths genotype_paths.cpp:943-1169), read acceptance
//       (vcf_writer.cpp:28-60) and the integer likelihood / depth / stat accumulation (vcf_writer.cpp:503-676,
//       haplotype.cpp:180-585) as atomics into widened per-bubble per-sample accumulators; optionally the phasing
//       connections (vcf_writer.cpp:88-250,587-637) into the pool's open-addressing table.
//   gather_segments_kernel / zero_segments_kernel, conn_rehash_kernel / conn_compact_kernel, build_table_kernel: upkeep.
//
// Order-sensitive semantics of the reference (bucket order, path order, candidate order) are preserved.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cuda_runtime.h>
#include <type_traits>

#include "gtb_device.cuh"

namespace gtb
{
__device__ uint32_t g_hash_tab[8 * 256]; // this translation unit's copies of HashTables (gtb_device.cuh)
__constant__ uint32_t c_hash_basis[64];

int upload_hash_tables_kernels()
{
  cudaError_t e = cudaMemcpyToSymbol(g_hash_tab, hash_tables().tab, sizeof(uint32_t) * 8 * 256);
  if (e == cudaSuccess)
    e = cudaMemcpyToSymbol(c_hash_basis, hash_tables().basis, sizeof(uint32_t) * 64);
  return (int)e;
}

namespace
{
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WARPS_PER_BLOCK = 4;
constexpr uint32_t OV_REFS = 1, OV_VARS = 2, OV_PATHS = 4, OV_LOCS = 8, OV_LABELS = 16, OV_CANDV = 32, OV_CANDS = 64,
                   OV_KEYS = 128, OV_TAP = 256, OV_POOL = 512, OV_LEN = 1024;
constexpr uint32_t BIGMM = 0x4000u; // "rejected" mismatch count ('<' / '>' in the graph sequence)
constexpr uint32_t INVALID = 0xFFFFFFFFu;
constexpr uint32_t SPECIAL_START = 0xD0000000u;
constexpr uint16_t F_PAIRED = 1, F_PROPER = 2, F_REV = 16, F_MREV = 32, F_FIRST = 64, F_MAPQ_BAD = 4096;

struct Loc
{
  uint32_t type; // 'R' or 'V'
  uint32_t node, order, offset;
};

__device__ __forceinline__ uint8_t comp4(uint8_t c) // seqan TranslateTableIupacToIupacComplement_: 4-bit reversal
{
  return (uint8_t)(((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3));
}

// seqan Iupac value -> char "UACMGRSVTWYHKDBN" (alphabet_residue_tabs.h:222-240) without a memory lookup
__device__ __forceinline__ uint8_t iupac_char(uint8_t c)
{
  uint64_t const lo = 0x565352474D434155ull; // "UACMGRSV" little endian
  uint64_t const hi = 0x4E42444B48595754ull; // "TWYHKDBN"
  return (uint8_t)(((c < 8 ? lo : hi) >> (8 * (c & 7))) & 0xFF);
}

// Working set of one read orientation during chaining / extension / filtering.  Two instantiations:
//   SlowState : big capacities, lives in shared memory, one warp per task (lane 0 runs the scalar logic)
//   FastState : small capacities, lives in per-thread local memory, one THREAD per task (32 tasks per warp);
//               a task that exceeds a small capacity is re-run by the slow kernel.
// 4 consecutive bytes from an arbitrary address: two aligned 32-bit loads + funnel shift.  May touch up to 7 bytes past
// p -- the graph sequence array and the read buffers are padded for that.
__device__ __forceinline__ uint32_t ld4_global(const uint8_t * p)
{
  uintptr_t const a = reinterpret_cast<uintptr_t>(p);
  const uint32_t * w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
  return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3u) * 8u);
}

__device__ __forceinline__ uint32_t ld4_state(const uint8_t * base, int j) // base is 4-byte aligned (state member)
{
  const uint32_t * w = reinterpret_cast<const uint32_t *>(base) + (j >> 2);
  return __funnelshift_r(w[0], w[1], (uint32_t)(j & 3) * 8u);
}

template <int P_, int V_, int C_, int CV_, int WL_, int LOC_>
struct StateT
{
  static constexpr int MAXP = P_, MAXV = V_, CANDS = C_, CANDV = CV_, WLCAP = WL_, MAXLOC = LOC_;
  struct Path
  {
    uint32_t start, end;
    uint16_t rs, re, mm, nvar;
    uint32_t order[V_];
    allele_mask_t mask[V_];
  };
  struct Cand
  {
    uint32_t len;
    uint32_t pos;
    uint32_t mm;
    uint32_t nvar;
    uint32_t vars[CV_];
  };
  Path paths[P_];
  Path pp[P_];
  Path op, np;
  Cand cands[C_];
  DevLabel wl[WL_];
  Loc locs[LOC_];
  uint16_t wl_list_start[P_ + 1];
  uint16_t wl_list_idx[P_];
  uint8_t matched[P_];
  int npaths, npp;
  uint32_t longest;
  uint32_t overflow;
};

// Warp-per-task working set: SlowState lives in shared memory (13 KB), HugeState in a per-warp global-memory slab with
// the reference's own limits as capacities (third tier, for reads the shared-memory state cannot hold).
template <int P_, int V_, int C_, int CV_, int WL_, int LOC_, int SPILL_, int REFS_>
struct WarpStateT : StateT<P_, V_, C_, CV_, WL_, LOC_>
{
  using Base = StateT<P_, V_, C_, CV_, WL_, LOC_>;
  using Cand = typename Base::Cand;
  static constexpr int CAND_TOTAL = C_ + SPILL_;
  static constexpr int REFCAP = REFS_;
  uint2 refs[REFS_];
  uint16_t list_start[NLISTS + 1];
  alignas(4) uint8_t seq[MAX_SEQ + 8]; // 4-bit codes in phase A, IUPAC characters afterwards
  Cand * cand_spill;        // this warp's global-memory extension of cands[] (SPILL_ entries)
  __device__ __forceinline__ uint8_t rd(int j) const { return seq[j]; }
  __device__ __forceinline__ uint32_t rd4(int j) const { return ld4_state(seq, j); }
  __device__ __forceinline__ void prepare(int, int) {} // the whole read is already decoded in seq[]
  // candidate i of the bubble expansion: the first C_ live in the state itself, the (rare) rest in a per-warp
  // global scratch area, so that the reference's limit of 128 open candidates (+ one round of growth) fits
  __device__ __forceinline__ Cand & cand_at(int i) { return (SPILL_ == 0 || i < C_) ? this->cands[i] : cand_spill[i - C_]; }
};
using SlowState = WarpStateT<MAXP, MAXV, CAND_CAP, CAND_V, WL_CAP, MAXLOC, CAND_SPILL, REF_CAP>;
using HugeState = WarpStateT<HUGE_P, HUGE_V, HUGE_C, HUGE_CV, HUGE_WL, HUGE_LOC, 0, HUGE_REFS>;

struct FastState : StateT<FAST_P, FAST_V, FAST_C, FAST_CV, FAST_WL, FAST_LOC>
{
  static constexpr int CAND_TOTAL = FAST_C;
  const uint8_t * s4; // packed 4-bit read
  int L, orient;
  alignas(4) uint8_t buf[MAX_SEQ + 8]; // IUPAC characters of the read window the current walk compares against
  // decodes read bases [from, to) once; the comparison loops of a walk touch the same ~26-base tail several times
  __device__ void prepare(int from, int to)
  {
    for (int j = from; j < to; ++j)
    {
      int const src = orient ? (L - 1 - j) : j;
      uint8_t const b = __ldg(s4 + (src >> 1));
      uint8_t c = (src & 1) ? (b & 15) : (b >> 4);
      if (orient)
        c = comp4(c);
      buf[j] = iupac_char(c);
    }
  }
  __device__ __forceinline__ uint8_t rd(int j) const { return buf[j]; }
  __device__ __forceinline__ uint32_t rd4(int j) const { return ld4_state(buf, j); }
  __device__ __forceinline__ Cand & cand_at(int i) { return cands[i]; }
};


__device__ __constant__ char IUPAC_CHAR[17] = "UACMGRSVTWYHKDBN";

// ------------------------------------------------------------------------------------------------ graph accessors
struct GR
{
  const DevRegion & R;
  __device__ GR(const DevRegion & r) : R(r) {}
  __device__ uint32_t ref_len(uint32_t r) const { return R.ref_seq_off[r + 1] - R.ref_seq_off[r]; }
  __device__ uint32_t var_len(uint32_t v) const { return R.var_seq_off[v + 1] - R.var_seq_off[v]; }
  __device__ const uint8_t * ref_dna(uint32_t r) const { return R.seq + R.ref_seq_off[r]; }
  __device__ const uint8_t * var_dna(uint32_t v) const { return R.seq + R.var_seq_off[v]; }
  __device__ uint32_t ref_reach(uint32_t r) const { return R.ref_order[r] + ref_len(r) - 1; }
  __device__ uint32_t var_reach(uint32_t v) const { return R.var_order[v] + var_len(v) - 1; }
  __device__ uint32_t variant_num(uint32_t v) const { return R.var_num[v]; }
  __device__ uint32_t bubble_ref_reach(uint32_t v) const { return var_reach(R.ref_var_off[R.var_out_ref[v] - 1]); }
  __device__ bool is_special(uint32_t p) const { return p >= SPECIAL_START && (p - SPECIAL_START) < R.n_special; }
  __device__ uint32_t ref_reach_pos(uint32_t p) const { return is_special(p) ? R.ref_reach_poses[p - SPECIAL_START] : p; }
  __device__ uint32_t actual_pos(uint32_t p) const { return is_special(p) ? R.actual_poses[p - SPECIAL_START] : p; }
  __device__ uint32_t special(uint32_t pos, uint32_t ref_reach) const // graph.cpp:1775-1782
  {
    int lo = 0, hi = (int)R.n_sp_keys;
    while (lo < hi)
    {
      int mid = (lo + hi) >> 1;
      if (R.sp_keys[mid] < ref_reach)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo >= (int)R.n_sp_keys || R.sp_keys[lo] != ref_reach)
      return INVALID;
    uint32_t const idx = pos - ref_reach - 1;
    if (R.sp_off[lo] + idx >= R.sp_off[lo + 1])
      return INVALID;
    return R.sp_list[R.sp_off[lo] + idx];
  }
};

// last ref node whose order <= pos (callers make sure pos >= ref_order[0])
__device__ __forceinline__ uint32_t last_ref_le(const DevRegion & R, uint32_t pos)
{
  if (R.pos_bucket)
  {
    uint32_t b = (pos - R.pos_base) >> 4;
    if (b >= R.n_pos_bucket)
      b = R.n_pos_bucket - 1;
    uint32_t r = R.pos_bucket[b];
    while (r + 1 < R.n_ref && R.ref_order[r + 1] <= pos)
      ++r;
    return r;
  }
  int lo = 1, hi = (int)R.n_ref;
  while (lo < hi)
  {
    int const mid = (lo + hi) >> 1;
    if (R.ref_order[mid] <= pos)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (uint32_t)(lo - 1);
}

// ------------------------------------------------------------------------------------------------ phase A: seeds
__device__ __forceinline__ uint32_t slot_of(const DevRegion & R, uint64_t key)
{
  return hash32_tab(g_hash_tab, key) >> (R.table_shift - 32);
}

// continues a linear-probing lookup whose first slot `s` (at index h) has already been loaded
__device__ __forceinline__ bool probe_resolve(const DevRegion & R, uint64_t key, uint32_t h, uint4 s, uint32_t & off, uint32_t & cnt)
{
  const uint4 * tab = reinterpret_cast<const uint4 *>(R.table);
  while (true)
  {
    if (s.w == 0)
      return false;
    uint64_t const k = (uint64_t)s.x | ((uint64_t)s.y << 32);
    if (k == key)
    {
      off = s.z;
      cnt = s.w;
      return true;
    }
    h = (h + 1) & R.table_mask;
    s = __ldg(tab + h);
  }
}

__device__ __forceinline__ bool probe(const DevRegion & R, uint64_t key, uint32_t & off, uint32_t & cnt)
{
  uint32_t const h = slot_of(R, key);
  uint4 const s = __ldg(reinterpret_cast<const uint4 *>(R.table) + h);
  return probe_resolve(R, key, h, s, off, cnt);
}

// Appends the bucket references of `nk` keys (key k produced by keyfn(k)) to S.refs in key order, applying the
// "more than 75 labels while probing several keys => give the slot up" rule of PHIndex::multi_get.
template <class WST, typename KeyFn>
__device__ void query_list(WST & S, const DevRegion & R, KeyFn keyfn, int nk, bool multi, int lane, int & nrefs)
{
  int const start = nrefs;
  uint32_t total = 0;
  bool dropped = false;
  for (int base = 0; base < nk; base += 32)
  {
    int const k = base + lane;
    bool found = false;
    uint32_t off = 0, cnt = 0;
    if (k < nk)
      found = probe(R, keyfn(k), off, cnt);
    uint32_t inc = found ? cnt : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      uint32_t const t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d)
        inc += t;
    }
    bool const exceed = multi && found && (total + inc) > 75u;
    if (__any_sync(FULL, exceed))
    {
      dropped = true;
      break;
    }
    unsigned const fm = __ballot_sync(FULL, found);
    int const pos = nrefs + __popc(fm & ((1u << lane) - 1u));
    if (found && pos < WST::REFCAP)
      S.refs[pos] = make_uint2(off, cnt);
    nrefs += __popc(fm);
    total += __shfl_sync(FULL, inc, 31);
  }
  if (dropped)
    nrefs = start;
  if (nrefs > WST::REFCAP)
  {
    if (lane == 0)
      S.overflow |= OV_REFS;
    nrefs = start;
  }
}

// to_uint64_vec for a slot with IUPAC / N bases (type_conversions.cpp:207-266); lane 0 only; returns key count
__device__ int expand_keys(const uint8_t * codes, uint64_t * keys, int cap)
{
  int n = 1;
  keys[0] = 0;
  for (int i = 0; i < 32; ++i)
  {
    int const origin = n;
    if (origin > 97)
      return 0;
    uint8_t const c = codes[i];
    for (int k = 0; k < origin; ++k)
    {
      if (c == 15 || c == 0)
      {
        if (n + 3 > cap)
          return -1;
        keys[n++] = keys[k] * 4 + 0;
        keys[n++] = keys[k] * 4 + 1;
        keys[n++] = keys[k] * 4 + 2;
        keys[k] = keys[k] * 4 + 3;
      }
      else
      {
        int set_count = __popc((unsigned)c);
        for (int b = 0; b < 4; ++b)
          if (c & (1 << b))
          {
            if (set_count == 1)
              keys[k] = keys[k] * 4 + b;
            else
            {
              if (n + 1 > cap)
                return -1;
              keys[n++] = keys[k] * 4 + b;
            }
            --set_count;
          }
      }
    }
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ phase B: chaining
template <class PathT>
__device__ __forceinline__ uint32_t psize(const PathT & p) { return (uint32_t)p.re - p.rs + 1u; }

// Path assignment that moves the header and only the nvar live (order, mask) entries: the working sets sit in local /
// shared / global memory, and a whole-struct copy moves all MAXV slots (typical paths cross 0-2 bubbles)
template <class PathT>
__device__ __forceinline__ void copy_path(PathT & dst, const PathT & src)
{
  dst.start = src.start;
  dst.end = src.end;
  dst.rs = src.rs;
  dst.re = src.re;
  dst.mm = src.mm;
  int const n = src.nvar;
  dst.nvar = (uint16_t)n;
  for (int i = 0; i < n; ++i)
  {
    dst.order[i] = src.order[i];
    dst.mask[i] = src.mask[i];
  }
}

template <class W>
__device__ void merge_with_current(W & S, const GR & g, typename W::Path & p, const DevLabel & l) // path.cpp:105-129
{
  if (l.var == INVALID)
    return;
  uint32_t const vo = g.R.var_order[l.var];
  allele_mask_t const bit = (allele_mask_t)1 << g.variant_num(l.var);
  for (int i = 0; i < p.nvar; ++i)
    if (p.order[i] == vo)
    {
      p.mask[i] |= bit;
      return;
    }
  if (p.nvar >= W::MAXV)
  {
    S.overflow |= OV_VARS;
    return;
  }
  p.order[p.nvar] = vo;
  p.mask[p.nvar] = bit;
  ++p.nvar;
}

// find_all_nonduplicated_paths, one label at a time (genotype_paths.cpp:32-66)
template <class W>
__device__ void pp_add_label(W & S, const GR & g, const DevLabel & l, uint16_t rs, uint16_t re, uint16_t mm)
{
  for (int d = 0; d < S.npp; ++d)
    if (S.pp[d].start == l.start && S.pp[d].end == l.end)
    {
      merge_with_current(S, g, S.pp[d], l);
      return;
    }
  if (S.npp >= W::MAXP)
  {
    S.overflow |= OV_PATHS;
    return;
  }
  typename W::Path & p = S.pp[S.npp++];
  p.start = l.start;
  p.end = l.end;
  p.rs = rs;
  p.re = re;
  p.mm = mm;
  p.nvar = 0;
  if (l.var != INVALID)
  {
    p.order[0] = g.R.var_order[l.var];
    p.mask[0] = (allele_mask_t)1 << g.variant_num(l.var);
    p.nvar = 1;
  }
}

// Path(p1, p2) (path.cpp:38-82); false = empty allele intersection (the caller drops the merge)
template <class W>
__device__ bool merge_paths(W & S, const typename W::Path & p1, const typename W::Path & p2, typename W::Path & out)
{
  copy_path(out, p2);
  for (int i = 0; i < p1.nvar; ++i)
  {
    bool found = false;
    for (int j = 0; j < out.nvar; ++j)
      if (p1.order[i] == out.order[j])
      {
        out.mask[j] &= p1.mask[i];
        if (out.mask[j] == 0)
          return false;
        found = true;
        break;
      }
    if (!found)
    {
      if (out.nvar >= W::MAXV)
      {
        S.overflow |= OV_VARS;
        return false;
      }
      out.order[out.nvar] = p1.order[i];
      out.mask[out.nvar] = p1.mask[i];
      ++out.nvar;
    }
  }
  out.rs = p1.rs;
  out.start = p1.start;
  out.mm = (uint16_t)(out.mm + p1.mm);
  return true;
}

template <class W>
__device__ void push_path(W & S, const typename W::Path & p)
{
  if (S.npaths >= W::MAXP)
  {
    S.overflow |= OV_PATHS;
    return;
  }
  copy_path(S.paths[S.npaths++], p);
}

// second half of add_next_kmer_labels (genotype_paths.cpp:308-351): S.pp holds the grouped new labels
template <class W>
__device__ void add_next(W & S, uint32_t rs)
{
  int const orig = S.npaths;
  for (int j = 0; j < S.npp; ++j)
    S.matched[j] = 0;
  for (int i = 0; i < orig; ++i)
  {
    if (S.paths[i].re != rs)
      continue;
    bool once = false;
    copy_path(S.op, S.paths[i]);
    for (int j = 0; j < S.npp; ++j)
    {
      if (S.op.end != S.pp[j].start)
        continue;
      if (!merge_paths(S, S.op, S.pp[j], S.np))
        continue;
      S.matched[j] = 1;
      if (once)
        push_path(S, S.np);
      else
      {
        S.longest = max(psize(S.np), S.longest);
        copy_path(S.paths[i], S.np);
        once = true;
      }
    }
  }
  for (int j = 0; j < S.npp; ++j)
    if (!S.matched[j])
    {
      S.longest = max(psize(S.pp[j]), S.longest);
      push_path(S, S.pp[j]);
    }
}

// second half of add_prev_kmer_labels (genotype_paths.cpp:247-291)
template <class W>
__device__ void add_prev(W & S, uint32_t re)
{
  int const orig = S.npaths;
  for (int j = 0; j < S.npp; ++j)
    S.matched[j] = 0;
  for (int i = 0; i < orig; ++i)
  {
    if (S.paths[i].rs != re)
      continue;
    bool once = false;
    copy_path(S.op, S.paths[i]);
    for (int j = 0; j < S.npp; ++j)
    {
      if (S.pp[j].end != S.op.start)
        continue;
      if (!merge_paths(S, S.pp[j], S.op, S.np))
        continue;
      S.matched[j] = 1;
      if (once)
        push_path(S, S.np);
      else
      {
        S.longest = max(psize(S.np), S.longest);
        copy_path(S.paths[i], S.np);
        once = true;
      }
    }
  }
  for (int j = 0; j < S.npp; ++j)
    if (!S.matched[j])
    {
      S.longest = max(psize(S.pp[j]), S.longest);
      push_path(S, S.pp[j]);
    }
}

// ------------------------------------------------------------------------------------------------ filters
template <class W>
__device__ void remove_short_paths(W & S) // genotype_paths.cpp:824-834
{
  if (S.longest <= 1)
    return;
  int w = 0;
  for (int i = 0; i < S.npaths; ++i)
    if (psize(S.paths[i]) >= S.longest)
    {
      if (w != i)
        copy_path(S.paths[w], S.paths[i]);
      ++w;
    }
  S.npaths = w;
}

template <class W>
__device__ void update_longest(W & S)
{
  uint32_t L = 0;
  for (int i = 0; i < S.npaths; ++i)
    L = max(L, psize(S.paths[i]));
  S.longest = L;
}

template <class W>
__device__ bool all_paths_unique(const W & S, const GR & g) // genotype_paths.cpp:219-231
{
  for (int i = 1; i < S.npaths; ++i)
    if (g.ref_reach_pos(S.paths[0].start) != g.ref_reach_pos(S.paths[i].start) &&
        g.ref_reach_pos(S.paths[0].end) != g.ref_reach_pos(S.paths[i].end))
      return false;
  return true;
}

template <class PathT>
__device__ bool path_is_reference(const PathT & p)
{
  for (int i = 0; i < p.nvar; ++i)
    if ((p.mask[i] & 1u) == 0)
      return false;
  return true;
}

// ------------------------------------------------------------------------------------------------ phase C: graph walk
// mismatches between read[r0 + i] and dna[i], i < m, by the reference's per-base rule (graph_utils.hpp:7-69: 'N' on either
// side matches, '<' / '>' in the graph rejects = BIGMM).  Four bases per step: equal words need no further look (a read
// never holds '<' or '>', so a graph tag always differs); a word that differs somewhere goes through the per-base rule.
template <class W>
__device__ uint32_t cmp_range(const W & S, int r0, const uint8_t * dna, uint32_t m)
{
  uint32_t mm = 0;
  for (uint32_t i = 0; i < m; i += 4)
  {
    uint32_t const g = ld4_global(dna + i), r = S.rd4(r0 + (int)i);
    uint32_t const rem = m - i;
    uint32_t x = g ^ r;
    if (rem < 4)
      x &= (1u << (8 * rem)) - 1u;
    if (x == 0)
      continue;
    uint32_t const nb = rem < 4 ? rem : 4u;
    for (uint32_t k = 0; k < nb; ++k)
    {
      uint32_t const gc = (g >> (8 * k)) & 0xFFu, rc = (r >> (8 * k)) & 0xFFu;
      if (gc == '>' || gc == '<')
        return BIGMM;
      mm += (gc != rc && rc != 'N' && gc != 'N');
    }
  }
  return mm;
}

// mismatches between read[koff + have + i] and dna[i] (graph_utils.hpp:7-37); BIGMM when the graph has '<' or '>'
template <class W>
__device__ uint32_t cmp_fwd(const W & S, int koff, uint32_t RL, uint32_t have, const uint8_t * dna, uint32_t n)
{
  if (have >= RL)
    return 0;
  return cmp_range(S, koff + (int)have, dna, min(n, RL - have));
}

// backward: the k-mer is read[0 .. RL), the candidate already covers its last `have` bases (graph_utils.hpp:39-69):
// dna[n-1-i] against read[RL-1-have-i] for i < m, i.e. the last m bases of dna against read[RL-have-m .. RL-have)
template <class W>
__device__ uint32_t cmp_bwd(const W & S, uint32_t RL, uint32_t have, const uint8_t * dna, uint32_t n)
{
  if (have >= RL)
    return 0;
  uint32_t const m = min(n, RL - have);
  return cmp_range(S, (int)(RL - have - m), dna + (n - m), m);
}

// Graph::get_locations_of_a_position (graph.cpp:931-1029,1154-1185) -> S.locs; returns count
template <class W>
__device__ int get_locations(W & S, const GR & g, uint32_t pos, const typename W::Path & path)
{
  const DevRegion & R = g.R;
  bool const sp = g.is_special(pos);
  if (sp)
    pos = R.actual_poses[pos - SPECIAL_START];
  int n = 0;
  if (pos < R.ref_order[0])
    return 0;
  if (R.n_ref == 1)
  {
    S.locs[0] = Loc{'R', 0u, R.ref_order[0], pos - R.ref_order[0]};
    return 1;
  }
  int rr = (int)last_ref_le(R, pos); // last ref node whose order <= pos
  if (pos < R.ref_order[rr] + g.ref_len(rr))
  {
    if (!sp)
    {
      S.locs[0] = Loc{'R', (uint32_t)rr, R.ref_order[rr], pos - R.ref_order[rr]};
      return 1;
    }
    --rr;
  }
  long long const PADDING = R.is_sv ? 1000000 : 1000;
  bool const path_empty = path.start == path.end;
  while (rr >= 0 && (long long)g.ref_reach(rr) + PADDING > (long long)pos)
  {
    uint32_t const vb = R.ref_var_off[rr], ve = R.ref_var_off[rr + 1];
    for (uint32_t v = vb; v < ve; ++v)
    {
      uint32_t const vo = R.var_order[v];
      if (pos >= vo && pos <= g.var_reach(v))
      {
        int j = -1;
        for (int k = 0; k < path.nvar; ++k)
          if (path.order[k] == vo)
          {
            j = k;
            break;
          }
        if (j < 0)
          continue;
        if (path_empty || ((path.mask[j] >> (v - vb)) & 1u))
        {
          if (n >= W::MAXLOC)
          {
            S.overflow |= OV_LOCS;
            return n;
          }
          S.locs[n++] = Loc{'V', v, vo, pos - vo};
        }
      }
    }
    --rr;
  }
  return n;
}

template <class W>
__device__ void emit_labels(W & S, const typename W::Cand & c, uint32_t a, uint32_t b, bool cand_is_end, int & wn)
{
  // forward walk: (start = a fixed, end = cand.pos); backward: (start = cand.pos, end = b fixed)
  uint32_t const st = cand_is_end ? a : c.pos;
  uint32_t const en = cand_is_end ? c.pos : b;
  if (c.nvar == 0)
  {
    if (wn >= W::WLCAP)
    {
      S.overflow |= OV_LABELS;
      return;
    }
    S.wl[wn++] = DevLabel{st, en, INVALID};
    return;
  }
  for (uint32_t k = 0; k < c.nvar; ++k)
  {
    if (wn >= W::WLCAP)
    {
      S.overflow |= OV_LABELS;
      return;
    }
    S.wl[wn++] = DevLabel{st, en, c.vars[k]};
  }
}

template <class W>
__device__ __forceinline__ bool cand_add_var(W & S, typename W::Cand & c, uint32_t v)
{
  if (c.nvar >= W::CANDV)
  {
    S.overflow |= OV_CANDV;
    return false;
  }
  c.vars[c.nvar++] = v;
  return true;
}

// Graph::get_labels_forward (graph.cpp:1187-1439). Candidate sequences are never materialised: a candidate is
// (length so far, mismatches so far, var nodes, end position); mismatches are additive over appended nodes.
template <class W>
__device__ void labels_forward(W & S, const GR & g, const Loc & s, int koff, uint32_t RL, uint32_t & max_mm, int & wn)
{
  const DevRegion & R = g.R;
  int const w0 = wn;
  int nc = 1;
  uint32_t vb = 0, ve = 0;
  {
    typename W::Cand & c = S.cand_at(0);
    c.nvar = 0;
    c.pos = 0;
    if (s.type == 'V')
    {
      uint32_t const v = s.node;
      c.vars[0] = v;
      c.nvar = 1;
      uint32_t const n = g.var_len(v) - s.offset;
      c.len = n;
      c.mm = cmp_fwd(S, koff, RL, 0, g.var_dna(v) + s.offset, n);
      if (n >= RL)
      {
        uint32_t ep = g.var_reach(v) - (n - RL);
        uint32_t const rr = g.bubble_ref_reach(v);
        if (ep > rr)
          ep = g.special(ep, rr);
        c.pos = ep;
      }
      else
      {
        uint32_t const r = R.var_out_ref[v];
        vb = R.ref_var_off[r];
        ve = R.ref_var_off[r + 1];
        uint32_t const rn = g.ref_len(r);
        c.mm += cmp_fwd(S, koff, RL, c.len, g.ref_dna(r), rn);
        c.len += rn;
        c.pos = g.ref_reach(r) - (c.len - RL);
      }
    }
    else
    {
      uint32_t const r = s.node;
      vb = R.ref_var_off[r];
      ve = R.ref_var_off[r + 1];
      uint32_t const n = g.ref_len(r) - s.offset;
      c.len = n;
      c.mm = cmp_fwd(S, koff, RL, 0, g.ref_dna(r) + s.offset, n);
      c.pos = g.ref_reach(r) - (n - RL);
    }
  }

  if (ve > vb && S.cand_at(0).len < RL)
  {
    uint32_t r = R.var_out_ref[vb];
    bool all_long = false;
    while (!all_long && nc < 128 && ve > vb)
    {
      all_long = true;
      int orig = nc;
      uint32_t const rn = g.ref_len(r);
      const uint8_t * rdna = g.ref_dna(r);
      uint32_t const rreach = g.ref_reach(r);
      for (int j = 0; j < orig; ++j)
      {
        if (S.cand_at(j).len >= RL)
          continue;
        for (uint32_t v = vb; v + 1 < ve; ++v)
        {
          typename W::Cand const & base = S.cand_at(j);
          uint32_t const m = g.var_len(v);
          uint32_t mm = base.mm + cmp_fwd(S, koff, RL, base.len, g.var_dna(v), m);
          uint32_t len = base.len + m;
          bool const enough = len >= RL;
          if (!enough)
          {
            mm += cmp_fwd(S, koff, RL, len, rdna, rn);
            len += rn;
          }
          if (mm <= max_mm)
          {
            if (nc >= W::CAND_TOTAL)
            {
              S.overflow |= OV_CANDS;
              continue;
            }
            typename W::Cand & nw = S.cand_at(nc);
            nw = base;
            if (!cand_add_var(S, nw, v))
              continue;
            nw.len = len;
            nw.mm = mm;
            if (len < RL)
              all_long = false;
            if (enough)
            {
              uint32_t ep = g.var_reach(v) - (len - RL);
              uint32_t const rr = g.bubble_ref_reach(v);
              if (ep > rr)
                ep = g.special(ep, rr);
              nw.pos = ep;
            }
            else
              nw.pos = rreach - (len - RL);
            ++nc;
          }
        }
        // the last allele replaces candidate j in place (or erases it)
        {
          uint32_t const v = ve - 1;
          typename W::Cand & c = S.cand_at(j);
          uint32_t const m = g.var_len(v);
          uint32_t mm = c.mm + cmp_fwd(S, koff, RL, c.len, g.var_dna(v), m);
          uint32_t len = c.len + m;
          bool const enough = len >= RL;
          if (!enough)
          {
            mm += cmp_fwd(S, koff, RL, len, rdna, rn);
            len += rn;
          }
          if (mm <= max_mm && cand_add_var(S, c, v))
          {
            c.len = len;
            c.mm = mm;
            if (len < RL)
              all_long = false;
            if (enough)
            {
              uint32_t ep = g.var_reach(v) - (len - RL);
              uint32_t const rr = g.bubble_ref_reach(v);
              if (ep > rr)
                ep = g.special(ep, rr);
              c.pos = ep;
            }
            else
              c.pos = rreach - (len - RL);
          }
          else
          {
            for (int k = j; k + 1 < nc; ++k)
              S.cand_at(k) = S.cand_at(k + 1);
            --nc;
            --orig;
            --j;
          }
        }
      }
      if (!all_long)
      {
        vb = R.ref_var_off[r];
        ve = R.ref_var_off[r + 1];
        ++r;
      }
      else
        break;
    }
  }

  uint32_t start_pos = s.order + s.offset;
  if (s.type == 'V')
  {
    uint32_t const rr = g.bubble_ref_reach(s.node);
    if (start_pos > rr)
      start_pos = g.special(start_pos, rr);
  }
  for (int j = 0; j < nc; ++j)
  {
    typename W::Cand const & c = S.cand_at(j);
    if (c.len < RL)
      continue;
    if (c.mm > max_mm)
      continue;
    if (c.mm < max_mm)
    {
      max_mm = c.mm;
      wn = w0;
    }
    emit_labels(S, c, start_pos, 0, true, wn);
  }
}

// Graph::get_labels_backward (graph.cpp:1441-1701)
template <class W>
__device__ void labels_backward(W & S, const GR & g, const Loc & e, uint32_t RL, uint32_t & max_mm, int & wn)
{
  const DevRegion & R = g.R;
  int const w0 = wn;
  int nc = 1;
  uint32_t vb = 0, ve = 0;
  {
    typename W::Cand & c = S.cand_at(0);
    c.nvar = 0;
    c.pos = 0;
    if (e.type == 'V')
    {
      uint32_t const v = e.node;
      c.vars[0] = v;
      c.nvar = 1;
      uint32_t const n = e.offset + 1;
      c.len = n;
      c.mm = cmp_bwd(S, RL, 0, g.var_dna(v), n);
      if (n >= RL)
      {
        uint32_t sp = R.var_order[v] + (n - RL);
        uint32_t const rr = g.bubble_ref_reach(v);
        if (sp > rr)
          sp = g.special(sp, rr);
        c.pos = sp;
      }
      else
      {
        uint32_t const r = R.var_out_ref[v] - 1;
        uint32_t const rn = g.ref_len(r);
        c.mm += cmp_bwd(S, RL, c.len, g.ref_dna(r), rn);
        c.len += rn;
        c.pos = R.ref_order[r] + (c.len - RL);
        if (r != 0)
        {
          vb = R.ref_var_off[r - 1];
          ve = R.ref_var_off[r];
        }
      }
    }
    else
    {
      uint32_t const r = e.node;
      if (r != 0)
      {
        vb = R.ref_var_off[r - 1];
        ve = R.ref_var_off[r];
      }
      uint32_t const n = e.offset + 1;
      c.len = n;
      c.mm = cmp_bwd(S, RL, 0, g.ref_dna(r), n);
      c.pos = R.ref_order[r] + (n - RL);
    }
  }

  if (ve > vb && S.cand_at(0).len < RL)
  {
    uint32_t r = R.var_out_ref[vb] - 1;
    bool all_long = false;
    while (!all_long && nc < 128 && ve > vb)
    {
      all_long = true;
      int orig = nc;
      uint32_t const rn = g.ref_len(r);
      const uint8_t * rdna = g.ref_dna(r);
      uint32_t const rorder = R.ref_order[r];
      for (int j = 0; j < orig; ++j)
      {
        if (S.cand_at(j).len >= RL)
          continue;
        for (uint32_t v = vb; v + 1 < ve; ++v)
        {
          typename W::Cand const & base = S.cand_at(j);
          uint32_t const m = g.var_len(v);
          uint32_t mm = base.mm + cmp_bwd(S, RL, base.len, g.var_dna(v), m);
          uint32_t len = base.len + m;
          bool const enough = len >= RL;
          if (!enough)
          {
            mm += cmp_bwd(S, RL, len, rdna, rn);
            len += rn;
          }
          if (mm <= max_mm)
          {
            if (nc >= W::CAND_TOTAL)
            {
              S.overflow |= OV_CANDS;
              continue;
            }
            typename W::Cand & nw = S.cand_at(nc);
            nw = base;
            if (!cand_add_var(S, nw, v))
              continue;
            nw.len = len;
            nw.mm = mm;
            if (len < RL)
              all_long = false;
            if (enough)
            {
              uint32_t sp = R.var_order[v] + (len - RL);
              uint32_t const rr = g.bubble_ref_reach(v);
              if (sp > rr)
                sp = g.special(sp, rr);
              nw.pos = sp;
            }
            else
              nw.pos = rorder + (len - RL);
            ++nc;
          }
        }
        {
          uint32_t const v = ve - 1;
          typename W::Cand & c = S.cand_at(j);
          uint32_t const m = g.var_len(v);
          uint32_t mm = c.mm + cmp_bwd(S, RL, c.len, g.var_dna(v), m);
          uint32_t len = c.len + m;
          bool const enough = len >= RL;
          if (!enough)
          {
            mm += cmp_bwd(S, RL, len, rdna, rn);
            len += rn;
          }
          if (mm <= max_mm && cand_add_var(S, c, v))
          {
            c.len = len;
            c.mm = mm;
            if (len < RL)
              all_long = false;
            if (enough)
            {
              uint32_t sp = R.var_order[v] + (len - RL);
              uint32_t const rr = g.bubble_ref_reach(v);
              if (sp > rr)
                sp = g.special(sp, rr);
              c.pos = sp;
            }
            else
              c.pos = rorder + (len - RL);
          }
          else
          {
            for (int k = j; k + 1 < nc; ++k)
              S.cand_at(k) = S.cand_at(k + 1);
            --nc;
            --orig;
            --j;
          }
        }
      }
      if (!all_long)
      {
        if (r != 0)
        {
          --r;
          vb = R.ref_var_off[r];
          ve = R.ref_var_off[r + 1];
        }
        else
        {
          vb = ve = 0;
          break;
        }
      }
      else
        break;
    }
  }

  uint32_t end_pos = e.order + e.offset;
  if (e.type == 'V')
  {
    uint32_t const rr = g.bubble_ref_reach(e.node);
    if (end_pos > rr)
      end_pos = g.special(end_pos, rr);
  }
  for (int j = 0; j < nc; ++j)
  {
    typename W::Cand const & c = S.cand_at(j);
    if (c.len < RL)
      continue;
    if (c.mm < max_mm)
    {
      max_mm = c.mm;
      wn = w0;
      emit_labels(S, c, 0, end_pos, false, wn);
    }
    else if (c.mm == max_mm)
      emit_labels(S, c, 0, end_pos, false, wn);
  }
}

// walk_read_ends (forward = true, genotype_paths.cpp:483-553) / walk_read_starts (forward = false, :555-621)
template <class W>
__device__ void walk(W & S, const GR & g, uint32_t L, bool forward)
{
  if (S.npaths == 0 || psize(S.paths[0]) == L)
    return;
  // MAX_SEED_NUMBER_FOR_WALKING (256): no walk at all; MAX_SEED_NUMBER_ALLOWING_MISMATCHES (64): exact matches only
  // (genotype_paths.cpp:488-492,560-564; constants.hpp.in:42-43).  Only the huge tier can hold that many paths.
  if (W::MAXP > 256 && S.npaths > 256)
    return;
  bool const exact_only = W::MAXP > 64 && S.npaths > 64;
  {
    // read window this walk can touch: [min read_end_index, L) forwards, [0, max read_start_index] backwards
    int from = (int)L, to = 0;
    for (int pi = 0; pi < S.npaths; ++pi)
    {
      if (forward && S.paths[pi].re != L - 1)
        from = min(from, (int)S.paths[pi].re);
      if (!forward && S.paths[pi].rs != 0)
        to = max(to, (int)S.paths[pi].rs + 1);
    }
    if (forward)
      S.prepare(from, (int)L);
    else
      S.prepare(0, to);
  }
  uint32_t best_mm = 7;
  int nlists = 0;
  int committed = 0;
  S.wl_list_start[0] = 0;
  for (int pi = 0; pi < S.npaths; ++pi)
  {
    typename W::Path const & p = S.paths[pi];
    uint32_t klen;
    int koff = 0;
    int nloc;
    if (forward)
    {
      if (p.re == L - 1)
        continue;
      nloc = get_locations(S, g, p.end, p);
      klen = L - p.re;
      koff = p.re;
    }
    else
    {
      if (p.rs == 0)
        continue;
      klen = (uint32_t)p.rs + 1u;
      nloc = get_locations(S, g, p.start, p);
    }
    if (nloc == 0)
      continue;
    uint32_t mm = exact_only ? 0u : min(2u + klen / 11u, best_mm);
    // iterative_dfs (graph.cpp:1703-1754)
    int const pend0 = committed;
    int pend1 = committed;
    for (int li = 0; li < nloc; ++li)
    {
      uint32_t m2 = mm;
      int t1 = pend1;
      Loc const loc = S.locs[li];
      if (forward)
        labels_forward(S, g, loc, koff, klen, m2, t1);
      else
        labels_backward(S, g, loc, klen, m2, t1);
      if (t1 > pend1)
      {
        if (m2 < mm)
        {
          mm = m2;
          int const cnt = t1 - pend1;
          for (int k = 0; k < cnt; ++k)
            S.wl[pend0 + k] = S.wl[pend1 + k];
          pend1 = pend0 + cnt;
        }
        else if (m2 == mm)
          pend1 = t1;
      }
    }
    if (pend1 > pend0)
    {
      uint16_t const idx = forward ? p.re : p.rs;
      if (mm < best_mm)
      {
        int const cnt = pend1 - pend0;
        for (int k = 0; k < cnt; ++k)
          S.wl[k] = S.wl[pend0 + k];
        nlists = 1;
        S.wl_list_start[0] = 0;
        S.wl_list_start[1] = (uint16_t)cnt;
        S.wl_list_idx[0] = idx;
        best_mm = mm;
        committed = cnt;
      }
      else if (mm == best_mm)
      {
        S.wl_list_idx[nlists] = idx;
        S.wl_list_start[nlists + 1] = (uint16_t)pend1;
        ++nlists;
        committed = pend1;
      }
    }
  }
  for (int l = 0; l < nlists; ++l)
  {
    S.npp = 0;
    uint16_t const idx = S.wl_list_idx[l];
    for (int k = S.wl_list_start[l]; k < S.wl_list_start[l + 1]; ++k)
    {
      if (forward)
        pp_add_label(S, g, S.wl[k], idx, (uint16_t)(L - 1), (uint16_t)best_mm);
      else
        pp_add_label(S, g, S.wl[k], 0, idx, (uint16_t)best_mm);
    }
    if (forward)
      add_next(S, idx);
    else
      add_prev(S, idx);
  }
}

// remove_support_from_read_ends (genotype_paths.cpp:382-430), SV graphs only
template <class W>
__device__ void remove_support_from_read_ends(W & S, const GR & g)
{
  for (int pi = 0; pi < S.npaths; ++pi)
  {
    typename W::Path & p = S.paths[pi];
    if (p.nvar == 0)
      continue;
    if (!g.is_special(p.start) && !g.is_special(p.end))
      continue;
    int imin = 0, imax = 0;
    for (int k = 1; k < p.nvar; ++k)
    {
      if (p.order[k] < p.order[imin])
        imin = k;
      if (p.order[k] > p.order[imax]) // std::minmax_element returns the LAST max; orders are unique so no tie
        imax = k;
    }
    if (g.is_special(p.end) && (long long)g.actual_pos(p.end) <= (long long)p.order[imax] + 4)
      p.mask[imax] = 0;
    if (g.is_special(p.start))
    {
      bool amb;
      if (g.is_special(p.start + 4u))
        amb = g.ref_reach_pos(p.start) != g.ref_reach_pos(p.start + 4u);
      else
        amb = true;
      if (amb)
        p.mask[imin] = 0;
    }
  }
}

} // namespace

// ================================================================================================ task body (B..D)
namespace
{
// Phases B-D for one read orientation: chain the seed lists, extend both ends, filter (alignment.cpp:35-87).
// refs / list_start describe the 2*nslots seed lists as index bucket references in PHIndex::multi_get order.
template <class W>
__device__ void run_task(W & S, const GR & g, const uint2 * refs, const uint16_t * list_start, int nslots, int L)
{
  const DevRegion & R = g.R;
  // "all k-mers extremely common" bail-out (alignment.cpp:35-49)
  bool any_small = false;
  for (int i = 0; i < nslots && !any_small; ++i)
  {
    uint32_t c = 0;
    for (int k = list_start[2 * i]; k < list_start[2 * i + 1]; ++k)
      c += refs[k].y;
    if (c < 512u)
      any_small = true;
  }
  if (!any_small)
    return;
  for (int l = 0; l < 2 * nslots; ++l)
  {
    uint16_t const rs = (uint16_t)(31 * (l >> 1));
    if (list_start[l] == list_start[l + 1])
      continue; // add_next_kmer_labels with no labels changes nothing
    S.npp = 0;
    for (int k = list_start[l]; k < list_start[l + 1]; ++k)
    {
      uint2 const ref = refs[k];
      for (uint32_t q = 0; q < ref.y; ++q)
      {
        DevLabel const lab = R.labels[ref.x + q];
        pp_add_label(S, g, lab, rs, (uint16_t)(rs + 31), (uint16_t)(l & 1));
      }
    }
    add_next(S, rs);
  }
  remove_short_paths(S);
  walk(S, g, (uint32_t)L, false); // starts before ends (alignment.cpp:71-72)
  walk(S, g, (uint32_t)L, true);
  update_longest(S);
  remove_short_paths(S);
  // remove_paths_with_too_many_mismatches (genotype_paths.cpp:360-380)
  if (S.npaths > 0)
  {
    uint16_t mn = 10;
    for (int i = 0; i < S.npaths; ++i)
      mn = min(mn, S.paths[i].mm);
    int w = 0;
    for (int i = 0; i < S.npaths; ++i)
      if (S.paths[i].mm <= mn)
      {
        if (w != i)
          copy_path(S.paths[w], S.paths[i]);
        ++w;
      }
    S.npaths = w;
  }
  if (R.is_sv) // remove_fully_special_paths (genotype_paths.cpp:476-481)
  {
    int w = 0;
    for (int i = 0; i < S.npaths; ++i)
      if (g.ref_reach_pos(S.paths[i].start) != g.ref_reach_pos(S.paths[i].end))
      {
        if (w != i)
          copy_path(S.paths[w], S.paths[i]);
        ++w;
      }
    S.npaths = w;
  }
  // remove_non_ref_paths_when_read_matches_ref (genotype_paths.cpp:460-474)
  if (!all_paths_unique(S, g))
  {
    bool any_ref = false;
    for (int i = 0; i < S.npaths; ++i)
      if (path_is_reference(S.paths[i]))
      {
        any_ref = true;
        break;
      }
    if (any_ref)
    {
      int w = 0;
      for (int i = 0; i < S.npaths; ++i)
        if (path_is_reference(S.paths[i]))
        {
          if (w != i)
            copy_path(S.paths[w], S.paths[i]);
          ++w;
        }
      S.npaths = w;
    }
  }
  update_longest(S);
  remove_short_paths(S);
  if (R.is_sv)
    remove_support_from_read_ends(S, g);
}

// Writes the GenotypePaths record of a finished task; returns false when the path pool is exhausted.
template <class W>
__device__ bool write_result(const W & S, const GR & g, const LaunchParams & P, uint32_t task, uint32_t n_tasks_total)
{
  TaskSummary sum;
  sum.npaths = (uint16_t)S.npaths;
  sum.longest = (uint16_t)S.longest;
  sum.mm0 = 0;
  sum.altcalls = 0;
  sum.bits = TS_ALL_UNIQUE | TS_COMPUTED;
  sum.pad = 0;
  sum.path_off = task * INLINE_WORDS;
  bool ok = true;
  if (S.npaths > 0)
  {
    sum.mm0 = S.paths[0].mm;
    if (!all_paths_unique(S, g))
      sum.bits &= ~TS_ALL_UNIQUE;
    uint32_t words = 0, alt = 0;
    for (int i = 0; i < S.npaths; ++i)
    {
      words += PATH_HDR_WORDS + 2u * S.paths[i].nvar;
      for (int k = 0; k < S.paths[i].nvar; ++k)
        alt += (S.paths[i].mask[k] & 1u) == 0;
    }
    sum.altcalls = (uint16_t)min(alt, 0xFFFFu);
    unsigned long long off = (unsigned long long)task * INLINE_WORDS;
    if (words > INLINE_WORDS)
    {
      unsigned long long const base = (unsigned long long)n_tasks_total * INLINE_WORDS;
      off = base + atomicAdd(&P.counters->path_words, (unsigned long long)words);
      if (off + words > P.path_pool_cap)
        ok = false;
    }
    if (!ok)
    {
      sum.bits |= TS_OVERFLOW;
      sum.npaths = 0;
      atomicAdd(&P.counters->reasons[9], 1ull);
      atomicAdd(&P.counters->n_overflow, 1ull);
    }
    else
    {
      sum.path_off = (uint32_t)off;
      uint32_t * w = P.path_pool + off;
      for (int i = 0; i < S.npaths; ++i)
      {
        auto const & p = S.paths[i];
        *w++ = p.start;
        *w++ = p.end;
        *w++ = (uint32_t)p.rs | ((uint32_t)p.re << 16);
        *w++ = (uint32_t)p.mm | ((uint32_t)p.nvar << 16);
        for (int k = 0; k < p.nvar; ++k)
        {
          *w++ = p.order[k];
          *w++ = (uint32_t)p.mask[k];
        }
      }
    }
  }
  P.summaries[task] = sum;
  return ok;
}

// debug tap of the seed lists (lane 0 / one thread): labels of every list in multi_get order
__device__ bool write_seed_tap(const LaunchParams & P, const DevRegion & R, uint32_t task, const uint2 * refs,
                               const uint16_t * list_start, int nslots)
{
  uint32_t tot = 0;
  for (int l = 0; l < 2 * nslots; ++l)
  {
    uint32_t c = 0;
    for (int k = list_start[l]; k < list_start[l + 1]; ++k)
      c += refs[k].y;
    P.tap.list_count[(size_t)task * NLISTS + l] = c;
    tot += c;
  }
  for (int l = 2 * nslots; l < NLISTS; ++l)
    P.tap.list_count[(size_t)task * NLISTS + l] = 0;
  P.tap.nslots[task] = nslots;
  unsigned long long const base = atomicAdd(&P.counters->dbg_label_words, (unsigned long long)tot);
  if (base + tot > P.tap.pool_cap)
  {
    for (int l = 0; l < NLISTS; ++l)
      P.tap.list_count[(size_t)task * NLISTS + l] = 0;
    return false;
  }
  unsigned long long o = base;
  for (int l = 0; l < 2 * nslots; ++l)
  {
    P.tap.list_off[(size_t)task * NLISTS + l] = (uint32_t)o;
    for (int k = list_start[l]; k < list_start[l + 1]; ++k)
      for (uint32_t q = 0; q < refs[k].y; ++q)
        P.tap.pool[o++] = R.labels[refs[k].x + q];
  }
  return true;
}

// Seed record handed from probe_kernel to chain_kernel (global memory, 96 bytes)
struct SeedRec
{
  uint8_t cnt[NLISTS]; // bucket references per list
  uint8_t nslots;
  uint8_t slow;        // 1: needs the slow kernel (IUPAC/N seed, or more than SEED_INLINE references)
  uint16_t pad;
  uint2 refs[SEED_INLINE];
};
static_assert(sizeof(SeedRec) == SEED_REC_BYTES, "SeedRec layout");

__device__ __forceinline__ void push_slow(const LaunchParams & P, uint32_t task)
{
  P.pending[task] = 1;
  unsigned long long const i = atomicAdd(&P.counters->n_slow, 1ull);
  P.slow_tasks[i] = task;
}
__device__ __forceinline__ void push_slow2(const LaunchParams & P, uint32_t task) // from chain_general_kernel
{
  P.pending[task] = 1;
  P.slow2_tasks[atomicAdd(&P.counters->n_slow2, 1u)] = task;
}
} // namespace

// ================================================================================================ batch preparation
namespace
{
// orientations align_read aligns for a unit record: none below 2K-1 bases; forward only for unpaired reads and for properly
// oriented pairs within 1200 bp; else both (alignment.cpp:343-360)
__device__ __forceinline__ uint32_t orientations_of(const PrepParams & p, uint32_t k)
{
  uint32_t const L = p.lseq[k];
  if (L < 63u || L > (uint32_t)MAX_SEQ)
    return 0;
  uint32_t const flag = p.flag[k];
  int32_t const isz = p.isize[k];
  bool const fwd_only = (flag & 1u) == 0 || (p.same_tid[k] && isz > -1200 && isz < 1200 && (((flag & 16u) != 0) != ((flag & 32u) != 0)));
  return fwd_only ? 1u : 2u;
}
} // namespace

__global__ void __launch_bounds__(256) prep_flags_kernel(PrepParams p)
{
  uint32_t const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_records)
    return;
  uint32_t err = 0;
  int32_t const d = p.dup_of[k];
  bool is_unit = d < 0;
  if (d >= (int32_t)k)
  {
    err |= PREP_ERR_DUP;
    is_unit = true;
  }
  if (p.lseq[k] > (uint16_t)MAX_SEQ)
    err |= PREP_ERR_LEN;
  int32_t const s = p.sample[k];
  if (s < 0 || s >= (int32_t)p.regions[p.region[k]].n_samples)
  {
    err |= PREP_ERR_SAMPLE;
    p.sample[k] = 0;
  }
  if (p.mate[k] >= (int32_t)k)
  {
    err |= PREP_ERR_MATE;
    p.mate[k] = -1;
  }
  if (err)
    atomicOr(&p.counters->input_bits, err);
  p.scan[k] = is_unit ? ((1ull << 32) | orientations_of(p, k)) : 0ull;
}

__global__ void __launch_bounds__(256) prep_fill_kernel(PrepParams p)
{
  uint32_t const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_records)
    return;
  int32_t const d = p.dup_of[k];
  bool const is_unit = d < 0 || d >= (int32_t)k;
  unsigned long long const v = p.scan[k];
  uint32_t const u = (uint32_t)(v >> 32), a = (uint32_t)v;
  uint32_t n_or = 0;
  if (is_unit)
  {
    n_or = orientations_of(p, k);
    p.unit[k] = (int32_t)u;
    p.unit_record[u] = (int32_t)k;
    if (n_or >= 1)
      p.active[a] = u * 2;
    if (n_or == 2)
      p.active[a + 1] = u * 2 + 1;
  }
  else
  {
    int32_t r = d; // the record whose alignment is re-used may itself re-use an earlier one
    for (;;)
    {
      int32_t const dd = p.dup_of[r];
      if (dd < 0 || dd >= r)
        break;
      r = dd;
    }
    p.unit[k] = (int32_t)(p.scan[r] >> 32);
  }
  if (k == p.n_records - 1)
  {
    p.counters->n_units = u + (is_unit ? 1u : 0u);
    p.counters->n_active = a + n_or;
  }
}

void launch_prep_flags(const PrepParams & p, void * stream)
{
  if (p.n_records)
    prep_flags_kernel<<<(p.n_records + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
}

void launch_prep_fill(const PrepParams & p, void * stream)
{
  if (p.n_records)
    prep_fill_kernel<<<(p.n_records + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
}

// ================================================================================================ probe kernel
// Phase A only: seed keys -> index bucket references (PHIndex::get / multi_get, ph_index.cpp:24-107).
//
// One persistent block per SM (32 warps, one warp per task at a time).  A task asks for 4 x (1 + 96) keys, and 95+ of the
// 96 Hamming-1 neighbours of a seed are absent from the index; answering "absent" from global memory costs one L1
// wavefront per lane (every lane hits a different line), and ncu showed exactly that as the limiter (388 scattered 4-byte
// gathers per task, L1TEX tag stage ~70 % busy, issue slots idle).  So the presence filter of the region a block works on
// lives in SHARED memory (<= 128 KiB: the region's presence bitmap OR-folded to 2^20 bits, ~6 % of the bits set for a 50 kb
// region); a block walks its contiguous share of the task list region by region (tasks arrive grouped by region) and
// re-stages the filter when the region changes.  Only keys whose filter bit is set go to the table in HBM/L2: the exact key
// + ~6 neighbours per slot, compacted in key order onto lanes 0..n-1 so the table is walked once per slot.  Regions whose
// bitmap would have to be folded more than 4x (more than ~2.6e5 distinct k-mers) keep probing the global bitmap.
//
// The table hash is GF(2)-linear (gtb_device.cuh): hash(key) is one __reduce_xor_sync over per-lane words, the hash of a
// neighbour one XOR with a per-lane constant.  PHIndex::multi_get's ">75 labels => drop the slot" rule is a warp scan over
// the compacted hits.
#ifndef GTB_PROBE_BLOCK_WARPS
#define GTB_PROBE_BLOCK_WARPS 32
#endif
constexpr int PROBE_BLOCK_WARPS = GTB_PROBE_BLOCK_WARPS;
constexpr uint32_t FILTER_LOG2_BITS = 20; // 128 KiB
constexpr uint32_t FILTER_WORDS = 1u << (FILTER_LOG2_BITS - 5);
constexpr int FILTER_MAX_FOLD = 2;        // log2: the bitmap is folded at most 4x

namespace
{
__device__ __forceinline__ uint32_t or_fold2(uint32_t x) // bit j of the result = bit 2j | bit 2j+1 of x (16 result bits)
{
  x = (x | (x >> 1)) & 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}
} // namespace

__global__ void __launch_bounds__(PROBE_BLOCK_WARPS * 32, 1024 / (PROBE_BLOCK_WARPS * 32)) probe_kernel(LaunchParams P)
{
  extern __shared__ __align__(16) uint32_t s_filter[]; // FILTER_WORDS
  __shared__ uint16_t s_cand[PROBE_BLOCK_WARPS][4 * 97 + 4]; // candidates of a task, slot-major and in key order:
                                                             // slot << 8 | (0 = the exact key, 1 + k = neighbour k = bb * 3 + j)
  __shared__ uint2 s_key[PROBE_BLOCK_WARPS][MAX_SLOTS];      // seed keys (lo, hi) of the task a warp works on
  __shared__ uint32_t s_kh[PROBE_BLOCK_WARPS][MAX_SLOTS];    // and their hashes
  __shared__ uint2 s_refs[PROBE_BLOCK_WARPS][SEED_INLINE];
  __shared__ uint8_t s_reflist[PROBE_BLOCK_WARPS][SEED_INLINE + 2]; // seed list (slot * 2 + ham) of every reference
  __shared__ uint32_t s_hm[96];                        // hashes of the 96 neighbour masks
  __shared__ uint32_t s_seg_end;
  int const lane = threadIdx.x & 31;
  int const wib = threadIdx.x >> 5;
  uint32_t const lt = (1u << lane) - 1u;
  uint16_t * cand = s_cand[wib];
  uint2 * refs = s_refs[wib];
  uint8_t * reflist = s_reflist[wib];
  uint32_t const n_active = P.counters->n_active;
  // Per lane, task-independent:
  //   hm[q]   hash of the XOR mask of this lane's q-th Hamming-1 neighbour: key index k = q*32 + lane flips base k/3 by (k%3 + 1)
  //   hb[v]   hash of base value v at this lane's position of the key (base `lane` of a seed sits at bits 2(31-lane))
  uint32_t hm[3], hb[4];
#pragma unroll
  for (int q = 0; q < 3; ++q)
  {
    int const k = q * 32 + lane;
    int const bb = k / 3, j = k % 3 + 1;
    hm[q] = ((j & 1) ? c_hash_basis[2 * bb] : 0u) ^ ((j & 2) ? c_hash_basis[2 * bb + 1] : 0u);
    asm volatile("" : "+r"(hm[q])); // opaque: keep it in a register instead of re-deriving it inside the task loop
  }
  hb[0] = 0;
  hb[1] = c_hash_basis[2 * (31 - lane)];
  hb[2] = c_hash_basis[2 * (31 - lane) + 1];
  hb[3] = hb[1] ^ hb[2];
  asm volatile("" : "+r"(hb[1]), "+r"(hb[2]), "+r"(hb[3]));
  if (threadIdx.x < 96)
  {
    int const k = threadIdx.x, bb = k / 3, j = k % 3 + 1;
    s_hm[k] = ((j & 1) ? c_hash_basis[2 * bb] : 0u) ^ ((j & 2) ? c_hash_basis[2 * bb + 1] : 0u);
  }

  // this block's contiguous share of the task list
  uint32_t const t_begin = (uint32_t)(((unsigned long long)n_active * blockIdx.x) / gridDim.x);
  uint32_t const t_end = (uint32_t)(((unsigned long long)n_active * (blockIdx.x + 1)) / gridDim.x);
  auto region_of = [&](uint32_t t) { return (uint32_t)P.batch.region[P.batch.unit_record[P.active_tasks[t] >> 1]]; };

  for (uint32_t seg = t_begin; seg < t_end;)
  {
    // ---- segment = the run of tasks of one region starting at seg (tasks are grouped by region, in batch order)
    uint32_t const slot = region_of(seg);
    __syncthreads(); // every warp is done with the previous segment (filter, s_seg_end)
    if (threadIdx.x == 0)
      s_seg_end = t_end;
    __syncthreads();
    for (uint32_t t = seg + 1 + threadIdx.x; t < t_end; t += blockDim.x)
    {
      bool const differs = region_of(t) != slot;
      if (differs)
        atomicMin(&s_seg_end, t);
      if (__any_sync(__activemask(), differs))
        break; // later strides only find later positions
    }
    const DevRegion & R = P.regions[slot];
    int const log2cap = 64 - R.table_shift;
    int const fold = log2cap + 2 - (int)FILTER_LOG2_BITS; // log2 of (bitmap bits / filter bits); <= 0: bitmap fits as it is
    bool const use_filter = fold <= P.filter_max_fold;
    if (use_filter)
    {
      if (fold <= 0)
      {
        uint32_t const words = 1u << (log2cap + 2 - 5);
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x)
          s_filter[w] = __ldg(R.bitmap + w);
      }
      else if (fold == 1)
        for (uint32_t w = threadIdx.x; w < FILTER_WORDS; w += blockDim.x)
        {
          uint2 const v = __ldg(reinterpret_cast<const uint2 *>(R.bitmap) + w);
          s_filter[w] = or_fold2(v.x) | (or_fold2(v.y) << 16);
        }
      else
        for (uint32_t w = threadIdx.x; w < FILTER_WORDS; w += blockDim.x)
        {
          uint4 const v = __ldg(reinterpret_cast<const uint4 *>(R.bitmap) + w);
          uint32_t const a = or_fold2(v.x) | (or_fold2(v.y) << 16), b2 = or_fold2(v.z) | (or_fold2(v.w) << 16);
          s_filter[w] = or_fold2(a) | (or_fold2(b2) << 16);
        }
    }
    __syncthreads();
    uint32_t const seg_end = s_seg_end;
    int const bshift = R.table_shift - 34 + (use_filter && fold > 0 ? fold : 0); // hash -> presence bit index
    int const tshift = R.table_shift - 32;                                        // hash -> table slot
    const uint4 * const table = reinterpret_cast<const uint4 *>(R.table);
    const uint32_t * const fbase = use_filter ? s_filter : R.bitmap; // presence bits: shared filter, else the global bitmap

    // per-warp software pipeline over its tasks: the task id / record / length of the NEXT task are requested while the
    // current one is processed (a dependent chain of three global loads otherwise opens every task)
    uint32_t t = seg + wib;
    uint32_t nx_task = t < seg_end ? P.active_tasks[t] : 0u;
    int nx_rec = t < seg_end ? P.batch.unit_record[nx_task >> 1] : 0;
    int nx_L = t < seg_end ? P.batch.lseq[nx_rec] : 0;
    for (; t < seg_end; t += PROBE_BLOCK_WARPS)
    {
      uint32_t const task = nx_task;
      int const orient = task & 1;
      int const rec = nx_rec;
      uint32_t const t_next = t + PROBE_BLOCK_WARPS;
      if (t_next < seg_end)
        nx_task = P.active_tasks[t_next];
      int const L = nx_L;
      int const nslots = 1 + (L - 32) / 31; // get_num_kmers (kmer_help_functions.cpp:10-17)
      // this lane's base of every seed slot, straight from the 4-bit BAM bytes (slot i covers read[31 i .. 31 i + 31])
      uint32_t code[MAX_SLOTS];
      {
        const uint8_t * s4 = P.batch.seq4 + (size_t)rec * GTB_SEQ_STRIDE;
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; ++i)
        {
          code[i] = 1;
          if (i < nslots)
          {
            int const pos = 31 * i + lane;
            int const src = orient ? (L - 1 - pos) : pos;
            uint32_t const byte = __ldg(s4 + (src >> 1));
            uint32_t c = (src & 1) ? (byte & 15u) : (byte >> 4);
            if (orient)
              c = comp4((uint8_t)c);
            code[i] = c;
          }
        }
      }
      if (t_next < seg_end)
        nx_rec = P.batch.unit_record[nx_task >> 1];

      // ---- phase 1, all slots: key, key hash, presence bits of the 96 neighbours, candidate list (slot-major, key order)
      int n_ok = 0;       // slots with pure ACGT seeds, processed here
      bool slow = false;  // IUPAC / N in a seed: key expansion runs in the slow kernel
      int nc[MAX_SLOTS], cstart[MAX_SLOTS];
      int total_c = 0;
#pragma unroll
      for (int i = 0; i < MAX_SLOTS; ++i)
      {
        nc[i] = 0;
        cstart[i] = total_c;
        if (i >= nslots || slow)
          continue;
        uint32_t const c = code[i];
        if (!__all_sync(FULL, __popc(c) == 1))
        {
          slow = true;
          continue;
        }
        n_ok = i + 1;
        uint32_t const v = (uint32_t)__ffs((int)c) - 1u; // A0 C1 G2 T3
        uint32_t const kh = __reduce_xor_sync(FULL, v == 0 ? 0u : v == 1 ? hb[1] : v == 2 ? hb[2] : hb[3]);
        uint32_t nb[3], nw[3];
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
          nb[q] = (kh ^ hm[q]) >> bshift;
          nw[q] = fbase[nb[q] >> 5];
        }
        uint64_t const val = (uint64_t)v << (2 * (31 - lane));
        uint32_t const lo = __reduce_or_sync(FULL, (uint32_t)val);
        uint32_t const hi = __reduce_or_sync(FULL, (uint32_t)(val >> 32));
        bool const f0 = (nw[0] >> (nb[0] & 31u)) & 1u, f1 = (nw[1] >> (nb[1] & 31u)) & 1u, f2 = (nw[2] >> (nb[2] & 31u)) & 1u;
        uint32_t const m0 = __ballot_sync(FULL, f0), m1 = __ballot_sync(FULL, f1), m2 = __ballot_sync(FULL, f2);
        int const n0 = __popc(m0), n1 = __popc(m1), n2 = __popc(m2);
        nc[i] = 1 + n0 + n1 + n2; // candidate 0 = the exact key
        uint16_t * cl = cand + total_c;
        uint16_t const tag = (uint16_t)(i << 8);
        if (lane == 0)
        {
          cl[0] = tag;
          s_key[wib][i] = make_uint2(lo, hi);
          s_kh[wib][i] = kh;
        }
        if (f0)
          cl[1 + __popc(m0 & lt)] = (uint16_t)(tag | (1 + lane));
        if (f1)
          cl[1 + n0 + __popc(m1 & lt)] = (uint16_t)(tag | (33 + lane));
        if (f2)
          cl[1 + n0 + n1 + __popc(m2 & lt)] = (uint16_t)(tag | (65 + lane));
        total_c += nc[i];
      }
      __syncwarp();

      // ---- phase 2 + 3: walk the table for the candidates, then settle the slots in order
      // resolve(g): candidate g of the task's list
      auto resolve = [&](int g, bool & found, uint32_t & off, uint32_t & cnt) {
        uint32_t const e = cand[g];
        uint32_t const i = e >> 8, kc = e & 255u;
        uint2 const kk = s_key[wib][i];
        uint64_t ck = (uint64_t)kk.x | ((uint64_t)kk.y << 32);
        uint32_t ch = s_kh[wib][i];
        if (kc)
        {
          uint32_t const k = kc - 1u;
          uint32_t const bb = (k * 171u) >> 9; // k / 3 for k < 96
          ck ^= (uint64_t)(k - 3u * bb + 1u) << (2u * bb);
          ch ^= s_hm[k];
        }
        uint32_t h = ch >> tshift;
        uint4 sl = __ldg(table + h);
        while (sl.w != 0) // linear probing; count == 0 marks an empty slot
        {
          if (((uint64_t)sl.x | ((uint64_t)sl.y << 32)) == ck)
          {
            found = true;
            off = sl.z;
            cnt = sl.w;
            break;
          }
          h = (h + 1) & R.table_mask;
          sl = __ldg(table + h);
        }
      };
      int nrefs = 0;
      uint32_t cnts = 0, cnts_hi = 0; // 8 x 8-bit list counts
      // per-slot state while its neighbour hits are absorbed round by round
      uint32_t total = 0;
      bool dropped = false;
      auto absorb = [&](int i, bool hit, uint32_t off, uint32_t cnt) { // hit: this lane holds a found NEIGHBOUR of slot i
        unsigned const fm = __ballot_sync(FULL, hit);
        if (fm == 0 || dropped)
          return;
        // PHIndex::multi_get gives the slot up as soon as the running label count passes 75 (ph_index.cpp:84-89); every
        // bucket holds at least one label, so the running count peaks at its end: one warp sum decides
        uint32_t const sum = __reduce_add_sync(FULL, hit ? cnt : 0u);
        if (total + sum > 75u)
        {
          dropped = true;
          return;
        }
        int const pos = nrefs + __popc(fm & lt);
        if (hit && pos < SEED_INLINE)
        {
          refs[pos] = make_uint2(off, cnt);
          reflist[pos] = (uint8_t)(2 * i + 1);
        }
        nrefs += __popc(fm);
        total += sum;
      };
      auto exact_hit = [&](int i, uint32_t off, uint32_t cnt, int src_lane, uint32_t & c0) { // exact-match list: 1 key, never dropped
        if (nrefs < SEED_INLINE)
        {
          if (lane == src_lane)
          {
            refs[nrefs] = make_uint2(off, cnt);
            reflist[nrefs] = (uint8_t)(2 * i);
          }
        }
        else
          slow = true;
        ++nrefs;
        c0 = 1;
      };
      auto close_slot = [&](int i, uint32_t c0, int before) {
        if (dropped)
          nrefs = before;
        if (nrefs > SEED_INLINE)
          slow = true;
        uint32_t const c1 = (uint32_t)(nrefs - before);
        if (i < 2)
          cnts |= (c0 << (16 * i)) | (c1 << (16 * i + 8));
        else
          cnts_hi |= (c0 << (16 * (i - 2))) | (c1 << (16 * (i - 2) + 8));
      };

      if (total_c <= 32)
      {
        // the usual case: every candidate of every slot in ONE round of table walks (one memory latency per task, not per slot)
        bool found = false;
        uint32_t off = 0, cnt = 0;
        if (lane < total_c)
          resolve(lane, found, off, cnt);
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; ++i)
        {
          if (i >= n_ok || slow)
            continue;
          int const a = cstart[i];
          uint32_t c0 = 0;
          if (__shfl_sync(FULL, (int)found, a))
            exact_hit(i, off, cnt, a, c0);
          int const before = nrefs;
          total = 0;
          dropped = false;
          absorb(i, found && lane > a && lane < a + nc[i], off, cnt);
          close_slot(i, c0, before);
        }
      }
      else
      {
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; ++i)
        {
          if (i >= n_ok || slow)
            continue;
          uint32_t c0 = 0;
          int before = nrefs;
          total = 0;
          dropped = false;
          for (int base = 0; base < nc[i]; base += 32)
          {
            int const ci = base + lane;
            bool found = false;
            uint32_t off = 0, cnt = 0;
            if (ci < nc[i])
              resolve(cstart[i] + ci, found, off, cnt);
            if (base == 0)
            {
              if (__shfl_sync(FULL, (int)found, 0))
                exact_hit(i, off, cnt, 0, c0);
              before = nrefs;
              if (lane == 0)
                found = false;
            }
            absorb(i, found, off, cnt);
          }
          close_slot(i, c0, before);
        }
      }
      if (t_next < seg_end)
        nx_L = P.batch.lseq[nx_rec];
      __syncwarp();
      SeedRec * out = reinterpret_cast<SeedRec *>(P.seed_recs) + t;
      if (lane == 0)
      {
        uint32_t * w = reinterpret_cast<uint32_t *>(out);
        w[0] = cnts;
        w[1] = cnts_hi;
        w[2] = (uint32_t)nslots | ((slow ? 1u : 0u) << 8);
      }
      if (!slow && lane < nrefs)
        out->refs[lane] = refs[lane];
      // label record for chain_kernel: the labels themselves with the bubble order / allele number of their var node, one
      // reference per lane (a bucket holds one or two labels)
      {
        uint4 * lrec = reinterpret_cast<uint4 *>(static_cast<uint8_t *>(P.lab_recs) + (size_t)t * LAB_REC_BYTES);
        bool const mine = !slow && lane < nrefs;
        uint2 const rf = mine ? refs[lane] : make_uint2(0u, 0u);
        uint32_t const li = mine ? reflist[lane] : 0u;
        uint32_t inc = rf.y;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) // references sit in lanes 0 .. SEED_INLINE - 1
        {
          uint32_t const tt = __shfl_up_sync(FULL, inc, d);
          if (lane >= d)
            inc += tt;
        }
        uint32_t const T = __shfl_sync(FULL, inc, SEED_INLINE - 1);
        bool const fits = !slow && T <= (uint32_t)FT_LAB;
        uint32_t const lcnt = __reduce_add_sync(FULL, fits ? rf.y << (4 * li) : 0u);
        if (fits)
        {
          uint32_t o = inc - rf.y;
          for (uint32_t q = 0; q < rf.y; ++q, ++o)
          {
            DevLabel const lb = R.labels[rf.x + q];
            uint32_t order = 0, meta = li << 8;
            if (lb.var != INVALID)
            {
              order = R.var_order[lb.var];
              meta |= (uint32_t)R.var_num[lb.var] | (1u << 16);
            }
            lrec[1 + o] = make_uint4(lb.start, lb.end, order, meta);
          }
        }
        if (lane == 0)
          lrec[0] = make_uint4(lcnt, (fits ? T : 0xFFu) | ((uint32_t)nslots << 8) | ((slow ? 1u : 0u) << 16), 0u, 0u);
      }
      __syncwarp();
    }
    seg = seg_end;
  }
}

// ================================================================================================ chain kernel (fast tier)
// One THREAD per active task, working set in registers + 364 bytes of shared memory -- no local memory.
//
// What makes a small working set enough: in add_next_kmer_labels (genotype_paths.cpp:294-352) every existing path evolves
// on its own -- it is compared with the new label groups, replaced in place by its first successful merge, and only a
// SECOND success or an unmatched group appends a new path.  Paths therefore form a forest rooted at unmatched groups, the
// order of the survivors is the order of their roots, and remove_short_paths (alignment.cpp:68) keeps the paths of maximal
// size only.  A path that chains all seed slots (read_start_index 0 .. read_end_index 31 * nslots) can only be rooted at a
// group of slot 0 (exact list, then Hamming-1 list); as soon as one exists, everything rooted at a later slot is shorter
// and is removed before the read ends are walked, whatever it did in between.  So the tier follows the slot-0 roots ONE AT
// A TIME through the later slots (a path that would split in two hands the task to the general tier), keeps at most two
// full chains, walks their common tail with the general tier's own labels_forward on a small shared-memory candidate list,
// merges the walked labels back (a path only ever sees the groups of a list whose read_start_index equals its
// read_end_index) and applies the filters of alignment.cpp:73-87 to the one or two paths that are left.  Everything it
// cannot hold -- SV graphs, more than 8 seed labels, more than 4 bubbles on a path, a split, more than two full chains or
// none, a chain end inside a bubble or at a special position, a walk beyond 4 candidates -- goes to chain_general_kernel
// through a queue, which runs the unrestricted code on the compacted rest.
namespace
{
constexpr int FT_V = 4, FT_SURV = 2;
constexpr int FT_PATH_WORDS = 5 + 2 * FT_V;

struct FPath // registers: order[] / mask[] are only ever indexed with compile-time constants (for_v)
{
  uint32_t start, end, mm, nvar, re;
  uint32_t order[FT_V], mask[FT_V];
};

// f(integral_constant<int, 0>) ... f(integral_constant<int, FT_V - 1>): the index is a constant in the source, so the
// arrays are split into registers whatever the loop unroller decides
template <int I = 0, class F>
__device__ __forceinline__ void for_v(F && f)
{
  f(std::integral_constant<int, I>{});
  if constexpr (I + 1 < FT_V)
    for_v<I + 1>(f);
}
#define FT_I (decltype(ic)::value)

__device__ __forceinline__ void fp_store(uint32_t * w, const FPath & p)
{
  w[0] = p.start;
  w[1] = p.end;
  w[2] = p.mm;
  w[3] = p.nvar;
  w[4] = p.re;
  for_v([&](auto ic) {
    w[5 + FT_I] = p.order[FT_I];
    w[5 + FT_V + FT_I] = p.mask[FT_I];
  });
}

__device__ __forceinline__ void fp_load(FPath & p, const uint32_t * w)
{
  p.start = w[0];
  p.end = w[1];
  p.mm = w[2];
  p.nvar = w[3];
  p.re = w[4];
  for_v([&](auto ic) {
    p.order[FT_I] = w[5 + FT_I];
    p.mask[FT_I] = w[5 + FT_V + FT_I];
  });
}

// Path::merge_with_current for one label (path.cpp:105-129): OR into the bubble's allele set or append the bubble
__device__ __forceinline__ bool fp_add_var(FPath & p, uint32_t order, uint32_t bit)
{
  bool found = false;
  for_v([&](auto ic) {
    if ((uint32_t)FT_I < p.nvar && p.order[FT_I] == order)
    {
      p.mask[FT_I] |= bit;
      found = true;
    }
  });
  if (found)
    return true;
  if (p.nvar >= (uint32_t)FT_V)
    return false;
  for_v([&](auto ic) {
    if ((uint32_t)FT_I == p.nvar)
    {
      p.order[FT_I] = order;
      p.mask[FT_I] = bit;
    }
  });
  ++p.nvar;
  return true;
}

// Path(p1 = P, p2 = the group already in `out`) (path.cpp:38-82): 1 merged, 0 empty allele intersection, -1 too many bubbles
__device__ __forceinline__ int fp_merge_into(const FPath & P, FPath & out)
{
  int status = 1;
  if (out.nvar == 0) // the usual seed: no bubble inside the k-mer -- the merged path keeps P's bubbles
  {
    for_v([&](auto ic) {
      out.order[FT_I] = P.order[FT_I];
      out.mask[FT_I] = P.mask[FT_I];
    });
    out.nvar = P.nvar;
  }
  else if (P.nvar != 0)
  for_v([&](auto ic) {
    if ((uint32_t)FT_I < P.nvar && status == 1)
    {
      uint32_t const po = P.order[FT_I], pm = P.mask[FT_I];
      bool found = false, empty = false;
      for_v([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if ((uint32_t)J < out.nvar && out.order[J] == po)
        {
          out.mask[J] &= pm;
          empty = empty || out.mask[J] == 0;
          found = true;
        }
      });
      if (empty)
        status = 0;
      else if (!found)
      {
        if (out.nvar >= (uint32_t)FT_V)
          status = -1;
        else
        {
          for_v([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            if ((uint32_t)J == out.nvar)
            {
              out.order[J] = po;
              out.mask[J] = pm;
            }
          });
          ++out.nvar;
        }
      }
    }
  });
  if (status == 1)
  {
    out.start = P.start;
    out.mm += P.mm;
  }
  return status;
}

// candidate list / walked labels / decoded read tail of one task, in shared memory (interface of labels_forward)
struct TinyWalk
{
  static constexpr int CAND_TOTAL = 4, CANDV = 3, WLCAP = 6;
  struct Cand
  {
    uint32_t len, pos, mm, nvar;
    uint32_t vars[CANDV];
  };
  Cand cands[CAND_TOTAL];
  DevLabel wl[WLCAP];
  uint32_t tail[9]; // IUPAC characters of read[tail0 ..) (at most 31), padded for ld4_state
  uint32_t overflow;
  int tail0;
  __device__ __forceinline__ uint32_t rd4(int j) const { return ld4_state(reinterpret_cast<const uint8_t *>(tail), j - tail0); }
  __device__ __forceinline__ Cand & cand_at(int i) { return cands[i]; }
};
// per-thread shared memory: [ seed labels (4 words each, as probe_kernel wrote them) | later the TinyWalk ] [ full chains ]
constexpr int FT_LAB_WORDS = FT_LAB * 4;
constexpr int FT_UNION_WORDS = (int)(sizeof(TinyWalk) / 4) > FT_LAB_WORDS ? (int)(sizeof(TinyWalk) / 4) : FT_LAB_WORDS;
constexpr int FT_WORDS = (FT_UNION_WORDS + FT_SURV * FT_PATH_WORDS) | 1; // odd stride: threads of a warp hit distinct banks
static_assert(sizeof(TinyWalk) % 4 == 0, "TinyWalk is made of 32-bit words");
static_assert(FT_LAB <= 15, "labels per list are counted in 4 bits");
constexpr uint32_t LAB_HAS_VAR = 1u << 16;
// seed labels of one task in shared memory, 4 words each (see FT_LAB_WORDS)
__device__ __forceinline__ void group_init(FPath & G, const uint32_t * lab, int j, int l)
{
  G.start = lab[4 * j];
  G.end = lab[4 * j + 1];
  G.nvar = 0;
  G.mm = (uint32_t)(l & 1);
  G.re = (uint32_t)(31 * (l >> 1) + 31);
  for_v([&](auto ic) { G.order[FT_I] = G.mask[FT_I] = 0; });
}
// Path::merge_with_current with label j; false: more than FT_V bubbles
__device__ __forceinline__ bool group_add(FPath & G, const uint32_t * lab, int j)
{
  uint32_t const meta = lab[4 * j + 3];
  return (meta & LAB_HAS_VAR) == 0 || fp_add_var(G, lab[4 * j + 2], 1u << (meta & 0xFFu));
}
// first label of its (start, end) group within its list [j0, ..) (find_all_nonduplicated_paths, genotype_paths.cpp:32-66)
__device__ __forceinline__ bool is_head(const uint32_t * lab, int j, int j0)
{
  for (int i = j0; i < j; ++i)
    if (lab[4 * i] == lab[4 * j] && lab[4 * i + 1] == lab[4 * j + 1])
      return false;
  return true;
}
// the whole group headed by label h of list l = [.., jend)
__device__ __forceinline__ bool build_group(FPath & G, const uint32_t * lab, int h, int jend, int l)
{
  group_init(G, lab, h, l);
  for (int j = h; j < jend; ++j)
    if (lab[4 * j] == G.start && lab[4 * j + 1] == G.end && !group_add(G, lab, j))
      return false;
  return true;
}
} // namespace

__global__ void __launch_bounds__(CHAIN_THREADS, CHAIN_MIN_BLOCKS) chain_kernel(LaunchParams P)
{
  extern __shared__ __align__(16) uint32_t sm_all[]; // CHAIN_THREADS * FT_WORDS
  uint32_t const t = blockIdx.x * CHAIN_THREADS + threadIdx.x;
  if (t >= P.counters->n_active)
    return;
  uint32_t const task = P.active_tasks[t];
  const uint4 * const lrec = reinterpret_cast<const uint4 *>(static_cast<const uint8_t *>(P.lab_recs) + (size_t)t * LAB_REC_BYTES);
  uint4 const hdr = lrec[0]; // labels per list (4 bits each) | number of labels, nslots, slow flag
  if ((hdr.y >> 16) & 1u)
  {
    atomicAdd(&P.counters->fast_reasons[11], 1ull); // marked by probe_kernel (IUPAC/N seed or > SEED_INLINE references)
    push_slow(P, task);
    return;
  }
  auto to_general = [&](int why) {
    atomicAdd(&P.counters->t0_reasons[why], 1ull);
    P.gen_tasks[atomicAdd(&P.counters->n_gen, 1u)] = t;
    if (P.defer)
      P.pending[task] = 1; // the first score pass runs beside chain_general_kernel and leaves this read's records alone
  };
  uint32_t const unit = task >> 1;
  int const rec = P.batch.unit_record[unit];
  const DevRegion & R = P.regions[P.batch.region[rec]];
  int const L = P.batch.lseq[rec];
  int const nslots = (int)((hdr.y >> 8) & 0xFFu);
  if (P.tap.list_count)
  {
    const SeedRec * recp = reinterpret_cast<const SeedRec *>(P.seed_recs) + t;
    uint32_t const c_lo = reinterpret_cast<const uint32_t *>(recp)[0], c_hi = reinterpret_cast<const uint32_t *>(recp)[1];
    uint16_t list_start[NLISTS + 1];
    list_start[0] = 0;
    for (int l = 0; l < NLISTS; ++l)
      list_start[l + 1] = (uint16_t)(list_start[l] + (((l < 4 ? c_lo : c_hi) >> (8 * (l & 3))) & 0xFFu));
    write_seed_tap(P, R, task, recp->refs, list_start, nslots);
  }
  if (R.is_sv)
    return to_general(T0_SV);
  int const T = (int)(hdr.y & 0xFFu);
  if (T > FT_LAB)
    return to_general(T0_LABELS);

  uint32_t * const sm = sm_all + threadIdx.x * FT_WORDS;
  uint32_t * const lab = sm;                   // [FT_LAB][4]
  uint32_t * const surv = sm + FT_UNION_WORDS; // [FT_SURV][FT_PATH_WORDS]
  uint32_t const lcnt = hdr.x;
#pragma unroll
  for (int j = 0; j < FT_LAB; ++j)
    if (j < T)
    {
      uint4 const v = lrec[1 + j];
      lab[4 * j] = v.x;
      lab[4 * j + 1] = v.y;
      lab[4 * j + 2] = v.z;
      lab[4 * j + 3] = v.w;
    }
  // "all k-mers extremely common" (alignment.cpp:35-49) cannot hold here: every exact list has fewer than 512 labels

  // ---- chain the slot-0 roots through the later slots
  auto lc = [&](int l) { return (int)((lcnt >> (4 * l)) & 15u); };
  uint32_t const re_full = (uint32_t)(31 * nslots);
  int const n0 = lc(0), n01 = n0 + lc(1);
  int ns = 0;
  for (int h = 0; h < n01; ++h)
  {
    int const l0 = h < n0 ? 0 : 1, j00 = l0 ? n0 : 0, j01 = l0 ? n01 : n0;
    if (!is_head(lab, h, j00))
      continue;
    FPath Pth;
    if (!build_group(Pth, lab, h, j01, l0))
      return to_general(T0_VARS);
    int j0 = n01;
    for (int l = 2; l < 2 * nslots; ++l)
    {
      int const cnt = lc(l), jend = j0 + cnt;
      if (cnt != 0 && Pth.re == (uint32_t)(31 * (l >> 1)))
      {
        // the usual list holds one group that continues the path: collect it in one pass
        FPath M;
        bool found = false, other = false;
        for (int j = j0; j < jend; ++j)
        {
          if (lab[4 * j] != Pth.end)
            continue;
          if (!found)
          {
            group_init(M, lab, j, l);
            found = true;
          }
          else if (lab[4 * j + 1] != M.end)
          {
            other = true;
            continue;
          }
          if (!group_add(M, lab, j))
            return to_general(T0_VARS);
        }
        if (other)
        {
          // several candidate groups: every one is tried against the path as it was; a second success would split it
          int nm = 0, hsel = -1;
          for (int j = j0; j < jend; ++j)
          {
            if (lab[4 * j] != Pth.end || !is_head(lab, j, j0))
              continue;
            if (!build_group(M, lab, j, jend, l))
              return to_general(T0_VARS);
            int const r = fp_merge_into(Pth, M);
            if (r < 0)
              return to_general(T0_VARS);
            if (r == 1 && nm++ == 0)
              hsel = j;
          }
          if (nm >= 2)
            return to_general(T0_MULTI);
          found = nm == 1;
          if (found)
            build_group(M, lab, hsel, jend, l);
        }
        if (found)
        {
          int const r = fp_merge_into(Pth, M);
          if (r < 0)
            return to_general(T0_VARS);
          if (r == 1)
            Pth = M;
        }
      }
      j0 = jend;
    }
    if (Pth.re == re_full)
    {
      if (ns == FT_SURV)
        return to_general(T0_SURVIVORS);
      fp_store(surv + ns * FT_PATH_WORDS, Pth);
      ++ns;
    }
  }
  if (ns == 0)
    return to_general(T0_NO_CHAIN); // no full chain: shorter paths compete, the general tier sorts that out

  // ---- walk_read_ends over the full chains (walk_read_starts has nothing to do: read_start_index is 0 everywhere)
  GR g(R);
  if (re_full != (uint32_t)(L - 1))
  {
    TinyWalk & W = *reinterpret_cast<TinyWalk *>(sm); // the labels are no longer needed
    {
      const uint8_t * s4 = P.batch.seq4 + (size_t)rec * GTB_SEQ_STRIDE;
      int const orient = task & 1;
      W.tail0 = (int)re_full;
      uint32_t word = 0;
      int n = 0;
      for (int j = (int)re_full; j < L; ++j, ++n)
      {
        int const src = orient ? (L - 1 - j) : j;
        uint8_t const b = __ldg(s4 + (src >> 1));
        uint8_t c = (src & 1) ? (b & 15) : (b >> 4);
        if (orient)
          c = comp4(c);
        word |= (uint32_t)iupac_char(c) << (8 * (n & 3));
        if ((n & 3) == 3)
        {
          W.tail[n >> 2] = word;
          word = 0;
        }
      }
      W.tail[n >> 2] = word;
      W.tail[(n >> 2) + 1] = 0;
    }
    uint32_t const klen = (uint32_t)L - re_full;
    uint32_t best_mm = 7;
    int nlists = 0, committed = 0;
    int list_start[FT_SURV + 1];
    list_start[0] = 0;
    bool first_in_ref = false;
    for (int pi = 0; pi < ns; ++pi)
    {
      const uint32_t * const pw = surv + pi * FT_PATH_WORDS;
      uint32_t const endpos = pw[1];
      if (g.is_special(endpos))
        return to_general(T0_SPECIAL);
      if (endpos < R.ref_order[0])
        continue; // no location
      // A second chain that ends at the same position inside a ref node walks the same tail from the same location under a
      // budget that the first walk already lowered to its own result: it finds the same labels again (or nothing again), and
      // a second list of the same groups changes no path -- whoever could merge with them already has.
      if (pi > 0 && first_in_ref && endpos == surv[1])
        continue;
      // Graph::get_locations_of_a_position (graph.cpp:931-1029): the ref node holding the position, else every var node
      // holding it that the path allows, bubble by bubble backwards (PADDING 1000); each location is walked as it is found
      int rr = R.n_ref > 1 ? (int)last_ref_le(R, endpos) : 0;
      bool const in_ref = R.n_ref == 1 || endpos < R.ref_order[rr] + g.ref_len(rr);
      if (pi == 0)
        first_in_ref = in_ref;
      uint32_t v_it = in_ref ? 0u : R.ref_var_off[rr];
      uint32_t mm = min(2u + klen / 11u, best_mm);
      int const pend0 = committed;
      int pend1 = committed;
      for (;;) // iterative_dfs over the locations (graph.cpp:1703-1754)
      {
        Loc loc;
        if (in_ref)
          loc = Loc{'R', (uint32_t)rr, R.ref_order[rr], endpos - R.ref_order[rr]};
        else
        {
          bool have = false;
          while (rr >= 0 && (long long)g.ref_reach(rr) + 1000 > (long long)endpos && !have)
          {
            uint32_t const vb = R.ref_var_off[rr], ve = R.ref_var_off[rr + 1];
            for (; v_it < ve && !have; ++v_it)
            {
              uint32_t const vo = R.var_order[v_it];
              if (endpos < vo || endpos > g.var_reach(v_it))
                continue;
              uint32_t allowed = 0; // the path's allele set of this bubble (a bubble the path does not hold gives no location)
              for (uint32_t k = 0; k < pw[3]; ++k)
                if (pw[5 + k] == vo)
                  allowed = pw[5 + FT_V + k];
              if ((allowed >> (v_it - vb)) & 1u)
              {
                loc = Loc{'V', v_it, vo, endpos - vo};
                have = true;
              }
            }
            if (!have)
            {
              --rr;
              if (rr >= 0)
                v_it = R.ref_var_off[rr];
            }
          }
          if (!have)
            break;
        }
        uint32_t m2 = mm;
        int t1 = pend1;
        W.overflow = 0;
        labels_forward(W, g, loc, (int)re_full, klen, m2, t1);
        if (W.overflow)
          return to_general(T0_WALK_CAP);
        if (t1 > pend1)
        {
          if (m2 < mm)
          {
            mm = m2;
            int const cnt = t1 - pend1;
            for (int k = 0; k < cnt; ++k)
              W.wl[pend0 + k] = W.wl[pend1 + k];
            pend1 = pend0 + cnt;
          }
          else if (m2 == mm)
            pend1 = t1;
        }
        if (in_ref)
          break;
      }
      if (pend1 > pend0)
      {
        if (mm < best_mm)
        {
          int const cnt = pend1 - pend0;
          for (int k = 0; k < cnt; ++k)
            W.wl[k] = W.wl[pend0 + k];
          nlists = 1;
          list_start[1] = cnt;
          best_mm = mm;
          committed = cnt;
        }
        else if (mm == best_mm)
        {
          list_start[++nlists] = pend1;
          committed = pend1;
        }
      }
    }
    // the walked labels of each list, grouped by (start, end), against every chain that still ends at the seed boundary
    for (int l = 0; l < nlists; ++l)
      for (int pi = 0; pi < ns; ++pi)
      {
        uint32_t * const pw = surv + pi * FT_PATH_WORDS;
        if (pw[4] != re_full)
          continue;
        FPath Pth;
        fp_load(Pth, pw);
        int nm = 0;
        for (int h = list_start[l]; h < list_start[l + 1]; ++h)
        {
          uint32_t const gs = W.wl[h].start, ge = W.wl[h].end;
          bool head = true;
          for (int i = list_start[l]; i < h; ++i)
            head = head && !(W.wl[i].start == gs && W.wl[i].end == ge);
          if (!head || gs != Pth.end)
            continue;
          FPath M;
          M.start = gs;
          M.end = ge;
          M.nvar = 0;
          M.mm = best_mm;
          M.re = (uint32_t)(L - 1);
          for_v([&](auto ic) { M.order[FT_I] = M.mask[FT_I] = 0; });
          bool ok = true;
          for (int j = h; j < list_start[l + 1]; ++j)
            if (W.wl[j].start == gs && W.wl[j].end == ge && W.wl[j].var != INVALID)
            {
              uint32_t const v = W.wl[j].var;
              ok = ok && fp_add_var(M, R.var_order[v], 1u << R.var_num[v]);
            }
          if (!ok)
            return to_general(T0_VARS);
          int const r = fp_merge_into(Pth, M);
          if (r < 0)
            return to_general(T0_VARS);
          if (r == 1)
          {
            if (nm == 0)
              fp_store(pw, M);
            ++nm;
          }
        }
        if (nm >= 2)
          return to_general(T0_WALK_MULTI);
      }
  }

  // ---- filters (alignment.cpp:73-87) over the one or two chains; unmatched walk groups became paths of at most 31 bases
  //      and fall to the first remove_short_paths
  FPath A, B;
  fp_load(A, surv);
  bool keepA = true, keepB = ns > 1;
  fp_load(B, surv + (keepB ? FT_PATH_WORDS : 0));
  {
    uint32_t const longest = max(A.re, keepB ? B.re : 0u); // read_start_index is 0: size = re + 1
    keepA = A.re >= longest;
    keepB = keepB && B.re >= longest;
    uint32_t mn = 10; // remove_paths_with_too_many_mismatches (genotype_paths.cpp:360-380)
    if (keepA)
      mn = min(mn, A.mm);
    if (keepB)
      mn = min(mn, B.mm);
    keepA = keepA && A.mm <= mn;
    keepB = keepB && B.mm <= mn;
    if (keepA && keepB)
    {
      // remove_non_ref_paths_when_read_matches_ref (genotype_paths.cpp:460-474) when !all_paths_unique (:219-231)
      bool const uniq = !(g.ref_reach_pos(A.start) != g.ref_reach_pos(B.start) && g.ref_reach_pos(A.end) != g.ref_reach_pos(B.end));
      if (!uniq)
      {
        bool refA = true, refB = true;
        for_v([&](auto ic) {
          refA = refA && !((uint32_t)FT_I < A.nvar && (A.mask[FT_I] & 1u) == 0);
          refB = refB && !((uint32_t)FT_I < B.nvar && (B.mask[FT_I] & 1u) == 0);
        });
        if (refA || refB)
        {
          keepA = refA;
          keepB = refB;
        }
      }
    }
    // (update_longest_path_size + remove_short_paths again: both chains already have the same size when both are left)
  }
  bool const two = keepA && keepB;
  if (!keepA) // the first path left is path 0
  {
    A.start = B.start;
    A.end = B.end;
    A.mm = B.mm;
    A.nvar = B.nvar;
    A.re = B.re;
    for_v([&](auto ic) {
      A.order[FT_I] = B.order[FT_I];
      A.mask[FT_I] = B.mask[FT_I];
    });
  }
  bool const any = keepA || keepB;

  // ---- result record (same layout as write_result); two paths of at most 4 bubbles always fit the inline words
  static_assert(FT_SURV * (PATH_HDR_WORDS + 2 * FT_V) <= INLINE_WORDS, "fast-tier results are inline");
  TaskSummary sum;
  sum.npaths = (uint16_t)(two ? 2 : any ? 1 : 0);
  sum.longest = 0;
  sum.mm0 = 0;
  sum.altcalls = 0;
  sum.bits = TS_ALL_UNIQUE | TS_COMPUTED;
  sum.pad = 0;
  sum.path_off = task * INLINE_WORDS;
  if (any)
  {
    sum.longest = (uint16_t)(A.re + 1);
    sum.mm0 = (uint16_t)A.mm;
    if (two && g.ref_reach_pos(A.start) != g.ref_reach_pos(B.start) && g.ref_reach_pos(A.end) != g.ref_reach_pos(B.end))
      sum.bits &= ~TS_ALL_UNIQUE;
    uint32_t alt = 0;
    for_v([&](auto ic) {
      alt += ((uint32_t)FT_I < A.nvar && (A.mask[FT_I] & 1u) == 0) ? 1u : 0u;
      alt += (two && (uint32_t)FT_I < B.nvar && (B.mask[FT_I] & 1u) == 0) ? 1u : 0u;
    });
    sum.altcalls = (uint16_t)alt;
    uint32_t * w = P.path_pool + (size_t)task * INLINE_WORDS;
    w[0] = A.start;
    w[1] = A.end;
    w[2] = A.re << 16; // read_start_index 0
    w[3] = A.mm | (A.nvar << 16);
    for_v([&](auto ic) {
      if ((uint32_t)FT_I < A.nvar)
      {
        w[4 + 2 * FT_I] = A.order[FT_I];
        w[5 + 2 * FT_I] = A.mask[FT_I];
      }
    });
    if (two)
    {
      w += PATH_HDR_WORDS + 2 * A.nvar;
      w[0] = B.start;
      w[1] = B.end;
      w[2] = B.re << 16;
      w[3] = B.mm | (B.nvar << 16);
      for_v([&](auto ic) {
        if ((uint32_t)FT_I < B.nvar)
        {
          w[4 + 2 * FT_I] = B.order[FT_I];
          w[5 + 2 * FT_I] = B.mask[FT_I];
        }
      });
    }
  }
  P.summaries[task] = sum;
}

// ================================================================================================ chain kernel (general tier)
// The tasks chain_kernel queued, one THREAD per task: phases B-D on a per-thread working set in local memory (6 paths x 8
// bubbles).  Only `gen_lanes` lanes of a warp take a task: a warp takes as long as the union of its lanes' control paths,
// and the queue holds the irregular reads.  Tasks that exceed a capacity here are queued for slow_kernel.
__global__ void __launch_bounds__(CHAIN_THREADS, 8) chain_general_kernel(LaunchParams P)
{
  uint32_t const lanes = P.gen_lanes;
  uint32_t const lane = threadIdx.x & 31u;
  if (lane >= lanes)
    return;
  uint32_t const gt = ((blockIdx.x * CHAIN_THREADS + threadIdx.x) >> 5) * lanes + lane;
  if (gt >= P.counters->n_gen)
    return;
  uint32_t const t = P.gen_tasks[gt];
  uint32_t const task = P.active_tasks[t];
  const SeedRec * recp = reinterpret_cast<const SeedRec *>(P.seed_recs) + t;
  uint32_t const w2 = reinterpret_cast<const uint32_t *>(recp)[2];
  struct TaskTimer // profiling aid, see LaunchParams::task_times
  {
    unsigned long long * p;
    __device__ static unsigned long long now()
    {
      unsigned long long v;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
      return v;
    }
    __device__ TaskTimer(unsigned long long * q) : p(q)
    {
      if (p)
        p[0] = now();
    }
    __device__ ~TaskTimer()
    {
      if (p)
        p[1] = now();
    }
  } timer(P.task_times ? P.task_times + 2 * (size_t)gt : nullptr);
  uint32_t const unit = task >> 1;
  int const rec = P.batch.unit_record[unit];
  const DevRegion & R = P.regions[P.batch.region[rec]];
  GR g(R);
  FastState S;
  S.s4 = P.batch.seq4 + (size_t)rec * GTB_SEQ_STRIDE;
  S.L = P.batch.lseq[rec];
  S.orient = task & 1;
  S.npaths = 0;
  S.npp = 0;
  S.longest = 0;
  S.overflow = 0;
  int const nslots = (int)(w2 & 0xFFu);
  uint16_t list_start[NLISTS + 1];
  {
    uint32_t const c_lo = reinterpret_cast<const uint32_t *>(recp)[0], c_hi = reinterpret_cast<const uint32_t *>(recp)[1];
    list_start[0] = 0;
#pragma unroll
    for (int l = 0; l < NLISTS; ++l)
    {
      uint32_t const c = ((l < 4 ? c_lo : c_hi) >> (8 * (l & 3))) & 0xFFu;
      list_start[l + 1] = (uint16_t)(list_start[l] + c);
    }
  }
  run_task(S, g, recp->refs, list_start, nslots, S.L); // (the seed tap was written by chain_kernel)
  if (S.overflow)
  {
    for (int q = 0; q < 12; ++q)
      if ((S.overflow >> q) & 1u)
        atomicAdd(&P.counters->fast_reasons[q], 1ull);
    push_slow2(P, task);
    return;
  }
  write_result(S, g, P, task, P.batch.n_units * 2);
}

// ================================================================================================ slow kernel
// One warp per queued task, lane 0 runs the scalar logic.  Handles seeds with IUPAC/N bases (key expansion,
// type_conversions.cpp:207-266) and everything the previous tier's capacities cannot hold.  Two instances:
//   slow_kernel : SlowState in shared memory; a task that still overflows is queued for huge_kernel
//   huge_kernel : HugeState in a per-warp global slab, capacities = the reference's own limits; an overflow here is final
// `first` = first queue index of this warp (the queues of several chunks are dealt round-robin over the grid's warps),
// `warp_global` = the warp's own index, which owns the global scratch slabs.
template <class WST, bool LAST>
__device__ void warp_tier(const LaunchParams & P, WST & S, int lane, uint32_t first, uint32_t warp_global, uint32_t total_warps,
                          const uint32_t * queue, uint32_t n_queued)
{
  uint32_t const n_tasks = P.batch.n_units * 2;

  for (uint32_t si = first; si < n_queued; si += total_warps)
  {
    uint32_t const task = queue[si];
    uint32_t const unit = task >> 1;
    int const orient = task & 1;
    int const rec = P.batch.unit_record[unit];
    int const L = P.batch.lseq[rec];
    const DevRegion & R = P.regions[P.batch.region[rec]];
    GR g(R);
    __syncwarp();
    {
      const uint8_t * s4 = P.batch.seq4 + (size_t)rec * GTB_SEQ_STRIDE;
      for (int j = lane; j < L; j += 32)
      {
        int const src = orient ? (L - 1 - j) : j;
        uint8_t const byte = s4[src >> 1];
        uint8_t c = (src & 1) ? (byte & 15) : (byte >> 4);
        if (orient)
          c = comp4(c);
        S.seq[j] = c;
      }
      if (lane == 0)
      {
        S.overflow = 0;
        S.cand_spill = LAST ? nullptr : reinterpret_cast<typename WST::Cand *>(P.cand_spill) + (size_t)warp_global * CAND_SPILL;
        S.npaths = 0;
        S.npp = 0;
        S.longest = 0;
      }
    }
    __syncwarp();

    int const nslots = 1 + (L - 32) / 31;
    int nrefs = 0;
    if (lane == 0)
      S.list_start[0] = 0;
    for (int i = 0; i < nslots; ++i)
    {
      uint8_t const c = S.seq[31 * i + lane];
      bool const pure = __all_sync(FULL, __popc((unsigned)c) == 1);
      if (pure)
      {
        uint64_t const val = (uint64_t)(__ffs((int)c) - 1) << (2 * (31 - lane));
        uint32_t const lo = __reduce_or_sync(FULL, (uint32_t)val);
        uint32_t const hi = __reduce_or_sync(FULL, (uint32_t)(val >> 32));
        uint64_t const key = (uint64_t)lo | ((uint64_t)hi << 32);
        query_list(S, R, [key](int) { return key; }, 1, false, lane, nrefs);
        if (lane == 0)
          S.list_start[2 * i + 1] = (uint16_t)nrefs;
        query_list(S, R, [key](int k) { return key ^ ((uint64_t)(k % 3 + 1) << (2 * (k / 3))); }, 96, true, lane, nrefs);
        if (lane == 0)
          S.list_start[2 * i + 2] = (uint16_t)nrefs;
      }
      else
      {
        // IUPAC / N bases: expand on lane 0 into the (not yet used) path storage
        uint64_t * keys = reinterpret_cast<uint64_t *>(S.paths);
        int nk = 0;
        if (lane == 0)
        {
          nk = expand_keys(S.seq + 31 * i, keys, (int)(sizeof(typename WST::Path) * WST::MAXP * 2 / sizeof(uint64_t)));
          if (nk < 0)
          {
            S.overflow |= OV_KEYS;
            nk = 0;
          }
        }
        nk = __shfl_sync(FULL, nk, 0);
        __syncwarp();
        int const before = nrefs;
        if (nk > 0)
          query_list(S, R, [keys](int k) { return keys[k]; }, nk, nk > 1, lane, nrefs);
        if (lane == 0)
          S.list_start[2 * i + 1] = (uint16_t)nrefs;
        __syncwarp();
        // a slot without a unique exact key gets no Hamming-1 keys: the same keys are queried again
        int const cnt = nrefs - before;
        if (nrefs + cnt > WST::REFCAP)
        {
          if (lane == 0)
            S.overflow |= OV_REFS;
        }
        else
        {
          for (int k = lane; k < cnt; k += 32)
            S.refs[nrefs + k] = S.refs[before + k];
          nrefs += cnt;
        }
        if (lane == 0)
          S.list_start[2 * i + 2] = (uint16_t)nrefs;
        __syncwarp();
      }
    }
    __syncwarp();
    // codes -> IUPAC characters for the graph walk (all lanes)
    for (int j = lane; j < L; j += 32)
      S.seq[j] = iupac_char(S.seq[j]);
    __syncwarp();

    if (lane == 0)
    {
      if (P.tap.list_count && !write_seed_tap(P, R, task, S.refs, S.list_start, nslots))
        S.overflow |= OV_TAP;
      run_task(S, g, S.refs, S.list_start, nslots, L);
      if (S.overflow && !LAST && !(S.overflow & (OV_TAP | OV_LEN)))
      {
        // the next tier has room for it
        for (int q = 0; q < 12; ++q)
          if ((S.overflow >> q) & 1u)
            atomicAdd(&P.counters->reasons[q], 1ull);
        unsigned long long const hi = atomicAdd(&P.counters->n_huge, 1ull);
        P.huge_tasks[hi] = task;
      }
      else if (S.overflow)
      {
        TaskSummary sum;
        sum.npaths = 0;
        sum.longest = 0;
        sum.mm0 = 0;
        sum.altcalls = 0;
        sum.bits = TS_OVERFLOW | TS_COMPUTED;
        sum.pad = 0;
        sum.path_off = task * INLINE_WORDS;
        P.summaries[task] = sum;
        atomicAdd(&P.counters->n_overflow, 1ull);
        for (int q = 0; q < 12; ++q)
          if ((S.overflow >> q) & 1u)
            atomicAdd(&P.counters->final_reasons[q], 1ull);
      }
      else
        write_result(S, g, P, task, n_tasks);
    }
    __syncwarp();
  }
}

// The queues of all chunks of one submit are served by ONE launch: slow tasks are few (tens per 10^5 reads) and each is a
// long single-lane job, so a launch costs the latency of its slowest task no matter how many chunks feed it.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) slow_kernel(const __grid_constant__ MultiLaunch M, int which)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlowState * all = reinterpret_cast<SlowState *>(smem_raw);
  int const wib = threadIdx.x >> 5;
  uint32_t const warp = blockIdx.x * WARPS_PER_BLOCK + wib, total = gridDim.x * WARPS_PER_BLOCK;
  uint32_t base = 0;
  for (int c = 0; c < M.n; ++c)
  {
    uint32_t const n = which ? M.p[c].counters->n_slow2 : (uint32_t)M.p[c].counters->n_slow;
    if (n)
      warp_tier<SlowState, false>(M.p[c], all[wib], threadIdx.x & 31, (warp + total - base % total) % total, warp, total,
                                  which ? M.p[c].slow2_tasks : M.p[c].slow_tasks, n);
    base += n;
  }
}

// One warp per block (grid = SM count); exits at once when slow_kernel queued nothing, which is the normal case.
__global__ void __launch_bounds__(32) huge_kernel(const __grid_constant__ MultiLaunch M)
{
  uint32_t base = 0;
  for (int c = 0; c < M.n; ++c)
  {
    uint32_t const n = (uint32_t)M.p[c].counters->n_huge;
    if (n)
    {
      HugeState & S = reinterpret_cast<HugeState *>(M.p[c].huge_states)[blockIdx.x];
      warp_tier<HugeState, true>(M.p[c], S, threadIdx.x, (blockIdx.x + gridDim.x - base % gridDim.x) % gridDim.x, blockIdx.x,
                                 gridDim.x, M.p[c].huge_tasks, n);
    }
    base += n;
  }
}

// ================================================================================================ score kernel
namespace
{
constexpr uint16_t NO_COVERAGE = 0xFFFFu, MULTI_ALT_COVERAGE = 0xFFFEu, MULTI_REF_COVERAGE = 0xFFFDu;

struct Geno // one GenotypePaths as the pool loop sees it after update_paths (alignment.cpp:482-538)
{
  TaskSummary s;
  uint16_t flags;
  uint16_t read_length;
  uint8_t mapq;
  uint8_t score_diff;
  bool proper_pair; // GenotypePaths::is_proper_pair(): ml_insert_size != 0x7FFFFFFF
};

__device__ __forceinline__ uint32_t T_of(const Geno & g) { return g.s.npaths > 0 ? g.s.longest : 0u; }

// compare_pair_of_genotype_paths (genotype_paths.cpp:976-1169)
__device__ int compare_pairs(const Geno & a1, const Geno & a2, const Geno & b1, const Geno & b2)
{
  uint32_t const T11 = T_of(a1), T12 = T_of(a2), T21 = T_of(b1), T22 = T_of(b2);
  uint32_t const MAX1 = max(T11, T12), MAX2 = max(T21, T22);
  uint32_t const P1 = a1.read_length, P2 = a2.read_length, MIN = 94;
  bool const perf1 = T11 >= P1 && T12 >= P2, perf2 = T21 >= P1 && T22 >= P2;
  if (perf1 || perf2)
  {
    if (perf1 && perf2)
    {
      uint32_t const m1 = (uint32_t)a1.s.mm0 + a2.s.mm0, m2 = (uint32_t)b1.s.mm0 + b2.s.mm0;
      if (m1 < m2)
        return 1;
      if (m2 < m1)
        return 2;
      uint32_t const n1 = (uint32_t)a1.s.npaths + a2.s.npaths, n2 = (uint32_t)b1.s.npaths + b2.s.npaths;
      if (n1 < n2)
        return 1;
      if (n2 < n1)
        return 2;
      uint32_t const c1 = (uint32_t)a1.s.altcalls + a2.s.altcalls, c2 = (uint32_t)b1.s.altcalls + b2.s.altcalls;
      return c1 >= c2 ? 1 : 2;
    }
    return perf1 ? 1 : 2;
  }
  if (MAX2 >= MIN && MAX2 > MAX1)
    return 2;
  if (MAX1 >= MIN && MAX1 > MAX2)
    return 1;
  if (MAX1 >= MIN && MAX2 >= MIN)
  {
    uint32_t m1 = 10, m2 = 10;
    if (T11 == MAX1)
      m1 = min(m1, (uint32_t)a1.s.mm0);
    if (T12 == MAX1)
      m1 = min(m1, (uint32_t)a2.s.mm0);
    if (T21 == MAX2)
      m2 = min(m2, (uint32_t)b1.s.mm0);
    if (T22 == MAX2)
      m2 = min(m2, (uint32_t)b2.s.mm0);
    if (m1 < m2)
      return 1;
    if (m2 < m1)
      return 2;
    if (min(T11, T12) < min(T21, T22))
      return 1;
    if (min(T21, T22) < min(T11, T12))
      return 2;
    return 0;
  }
  if (MAX2 == 0u && T11 >= 63u && T12 >= 63u)
    return 1;
  if (MAX1 == 0u && T21 >= 63u && T22 >= 63u)
    return 2;
  return 1;
}

// are_genotype_paths_good (vcf_writer.cpp:28-60); all surviving paths have size == longest
__device__ bool geno_good(const DevRegion & R, const Geno & g)
{
  if (g.s.npaths == 0)
    return false;
  bool const fully = g.s.longest == g.read_length;
  bool const uniq = (g.s.bits & TS_ALL_UNIQUE) != 0;
  if (!fully && (!uniq || g.s.longest < 63))
    return false;
  double const ratio = (double)g.s.mm0 / (double)g.s.longest;
  if (ratio > 0.05)
    return false;
  if (!fully && ratio > 0.025)
    return false;
  if (R.is_sv)
    if (!fully || g.s.longest < 90 || ratio > 0.03)
      return false;
  return true;
}

__device__ __forceinline__ void add_coverage(uint16_t & coverage, uint16_t c) // haplotype.cpp:180-227
{
  if (coverage == NO_COVERAGE)
    coverage = c;
  else if (coverage == MULTI_ALT_COVERAGE)
  {
    if (c == 0)
      coverage = MULTI_REF_COVERAGE;
  }
  else if (coverage == MULTI_REF_COVERAGE)
  {
  }
  else if (coverage != c)
    coverage = (coverage == 0 || c == 0) ? MULTI_REF_COVERAGE : MULTI_ALT_COVERAGE;
}

// ---- phasing connections (vcf_writer.cpp:587-637 inside one read, :186-249 across the two mates)
__device__ __forceinline__ unsigned long long conn_key(int sample, uint32_t hap1, uint32_t a1, uint32_t hap2, uint32_t a2)
{
  return CONN_OCCUPIED | ((unsigned long long)(uint32_t)sample << 42) | ((unsigned long long)hap1 << 26) |
         ((unsigned long long)a1 << 21) | ((unsigned long long)hap2 << 5) | (unsigned long long)a2;
}

__device__ void conn_insert(const DevRegion & R, unsigned long long key, uint32_t inc)
{
  uint32_t const mask = R.conn_mask;
  uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
  for (uint32_t n = 0; n <= mask; ++n)
  {
    unsigned long long cur = R.conn_keys[h];
    if (cur == 0)
    {
      cur = atomicCAS(&R.conn_keys[h], 0ull, key);
      if (cur == 0)
      {
        atomicAdd(&R.conn_state[0], 1u);
        cur = key;
      }
    }
    if (cur == key)
    {
      atomicAdd(&R.conn_vals[h], inc);
      return;
    }
    h = (h + 1) & mask;
  }
  atomicExch(&R.conn_state[1], 1u); // table full: reported as GTB_ERR_CAPACITY by the host
}

// every allele pair of two touched bubbles, oriented from the earlier bubble to the later one
__device__ void conn_add_pairs(const DevRegion & R, int sample, uint32_t hap_a, allele_mask_t Ea, uint32_t hap_b, allele_mask_t Eb,
                               uint32_t inc)
{
  if (hap_a == hap_b || inc == 0)
    return;
  if (hap_a > hap_b)
  {
    uint32_t const th = hap_a;
    hap_a = hap_b;
    hap_b = th;
    allele_mask_t const te = Ea;
    Ea = Eb;
    Eb = te;
  }
  for (allele_mask_t m1 = Ea; m1; m1 &= m1 - 1)
    for (allele_mask_t m2 = Eb; m2; m2 &= m2 - 1)
      conn_insert(R, conn_key(sample, hap_a, (uint32_t)__ffs((int)m1) - 1, hap_b, (uint32_t)__ffs((int)m2) - 1), inc);
}

// (bubble, explained alleles) of one scored read: the keys of the reference's new_connections map
struct TouchList
{
  int n;
  uint32_t hap[MAX_TOUCH];
  allele_mask_t expl[MAX_TOUCH];
};

// the cross-mate part of update_haplotype_scores_geno's merge: +1 for every (key of mate 1, key of mate 2) on different bubbles
__device__ void conn_merge_mates(const DevRegion & R, int sample, const TouchList & a, const TouchList & b)
{
  for (int i = 0; i < a.n; ++i)
    for (int j = 0; j < b.n; ++j)
      conn_add_pairs(R, sample, a.hap[i], a.expl[i], b.hap[j], b.expl[j], 1u);
}

// push_to_haplotype_scores (vcf_writer.cpp:503-676) + explain_to_score / coverage_to_gts / *_to_stats
__device__ bool push_scores(const LaunchParams & P, const DevRegion & R, const Geno & geno, int sample, TouchList * touched = nullptr)
{
  GR g(R);
  int const clipped_bp = (int)geno.read_length - (int)geno.s.longest;
  bool const fully = clipped_bp == 0;
  bool const non_unique = (geno.s.bits & TS_ALL_UNIQUE) == 0;
  uint32_t const mismatches = geno.s.mm0;

  uint32_t t_hap[MAX_TOUCH];
  allele_mask_t t_expl[MAX_TOUCH];
  uint16_t t_cov[MAX_TOUCH];
  bool t_ovl[MAX_TOUCH];
  int nt = 0;

  const uint32_t * w = P.path_pool + geno.s.path_off;
  for (int pi = 0; pi < geno.s.npaths; ++pi)
  {
    uint32_t const start = w[0], end = w[1];
    uint32_t const nvar = w[3] >> 16;
    w += PATH_HDR_WORDS;
    long long const rs = (long long)g.ref_reach_pos(start), re = (long long)g.ref_reach_pos(end);
    for (uint32_t k = 0; k < nvar; ++k)
    {
      uint32_t const order = w[2 * k];
      allele_mask_t const mask = (allele_mask_t)w[2 * k + 1];
      if (mask == 0)
        continue;
      // id2hap (vcf_writer.cpp:84): bubble index of this var order -- direct table, binary search as the fallback
      uint32_t hap = 0xFFFFu;
      if (R.hap_of_order && order - R.hap_base < R.hap_span)
        hap = R.hap_of_order[order - R.hap_base];
      if (hap == 0xFFFFu)
      {
        int lo = 0, hi = (int)R.n_bubbles;
        while (lo < hi)
        {
          int const mid = (lo + hi) >> 1;
          if (R.bubble_order[mid] <= order)
            lo = mid + 1;
          else
            hi = mid;
        }
        hap = (uint32_t)(lo - 1);
      }
      int t = -1;
      for (int q = 0; q < nt; ++q)
        if (t_hap[q] == hap)
        {
          t = q;
          break;
        }
      if (t < 0)
      {
        if (nt >= MAX_TOUCH)
          return false;
        t = nt++;
        t_hap[t] = hap;
        t_expl[t] = 0;
        t_cov[t] = NO_COVERAGE;
        t_ovl[t] = false;
      }
      bool const overlapping = (rs + 3) <= (long long)order && (re - 3) > (long long)order;
      t_ovl[t] = t_ovl[t] || overlapping;
      t_expl[t] |= mask;
      if (__popc(mask) == 1)
        add_coverage(t_cov[t], (uint16_t)(__ffs((int)mask) - 1));
      else
      {
        add_coverage(t_cov[t], 1);
        add_coverage(t_cov[t], (mask & 1u) ? 0 : 2);
      }
    }
    w += 2 * nvar;
  }

  if (touched) // "check connections": weight n1*n2, a pair of uniquely explained bubbles counts once, 3 <= weight <= 6 -> 6/weight
  {
    touched->n = nt;
    for (int i = 0; i < nt; ++i)
    {
      touched->hap[i] = t_hap[i];
      touched->expl[i] = t_expl[i];
      for (int j = i + 1; j < nt; ++j)
      {
        uint32_t const weight = (uint32_t)__popc(t_expl[i]) * (uint32_t)__popc(t_expl[j]);
        conn_add_pairs(R, sample, t_hap[i], t_expl[i], t_hap[j], t_expl[j], weight >= 3u ? 6u / weight : 1u);
      }
    }
  }

  uint32_t const NS = R.n_samples;
  for (int t = 0; t < nt; ++t)
  {
    uint32_t const b = t_hap[t];
    uint16_t const cov = t_cov[t];
    uint32_t const c0 = R.cov_off[b];
    uint32_t const cnum = R.cov_off[b + 1] - c0;
    // *_to_stats (haplotype.cpp:229-313)
    if (clipped_bp != 0)
    {
      long const scaled = (clipped_bp * 1000l) / geno.read_length;
      if (cov != NO_COVERAGE)
        atomicAdd(&R.vs_clipped_reads[b], 1ull);
      if (cov < MULTI_REF_COVERAGE)
        atomicAdd(&R.pa_clipped_bp[c0 + cov], (unsigned long long)scaled);
    }
    if (geno.mapq != 255)
    {
      unsigned long long const sq = (unsigned long long)geno.mapq * geno.mapq;
      if (cov != NO_COVERAGE)
        atomicAdd(&R.vs_mapq_squared[b], sq);
      if (cov < MULTI_REF_COVERAGE)
        atomicAdd(&R.pa_mapq_squared[c0 + cov], sq);
    }
    if (cov < MULTI_REF_COVERAGE)
    {
      bool const fwd = (geno.flags & F_REV) == 0, first = (geno.flags & F_FIRST) != 0;
      atomicAdd(&R.read_strand[(size_t)(c0 + cov) * 4 + (first ? 0 : 2) + (fwd ? 0 : 1)], 1u);
      uint32_t const mm8 = mismatches & 0xFFu;
      if (mm8 != 0)
        atomicAdd(&R.pa_mismatches[c0 + cov], (unsigned long long)((mm8 * 1000l) / geno.read_length));
      if (geno.score_diff != 0)
        atomicAdd(&R.pa_score_diff[c0 + cov], (unsigned long long)geno.score_diff);
    }
    // explain_to_score (haplotype.cpp:462-585)
    long eps = 12;
    eps -= (long)mismatches;
    if (non_unique)
      eps -= 3;
    if (geno.flags & F_MAPQ_BAD)
      eps -= 2;
    if (!fully)
      eps -= 3;
    if (!t_ovl[t])
      eps -= 1;
    uint32_t const e = (uint32_t)(max(eps, 8l) - 4);
    size_t const bs = (size_t)b * NS + sample;
    atomicAdd(&R.max_log_score[bs], e);
    {
      uint32_t const tri = R.score_off[b + 1] - R.score_off[b];
      uint32_t * ls = R.log_score + (size_t)R.score_off[b] * NS + (size_t)sample * tri;
      allele_mask_t const E = t_expl[t];
      uint32_t i = 0;
      for (uint32_t y = 0; y < cnum; ++y)
      {
        bool const ey = (E >> y) & 1u;
        for (uint32_t x = 0; x <= y; ++x, ++i)
        {
          bool const ex = (E >> x) & 1u;
          if (ex && ey)
            atomicAdd(&ls[i], e);
          else if (ex || ey)
            atomicAdd(&ls[i], e - 1);
        }
      }
    }
    // coverage_to_gts (haplotype.cpp:315-361); saturation applied on download
    if (cov == MULTI_REF_COVERAGE)
      atomicAdd(&R.amb[bs], 1u);
    else if (cov == MULTI_ALT_COVERAGE)
    {
      atomicAdd(&R.amb[bs], 1u);
      atomicAdd(&R.amb_alt[bs], 1u);
      if (geno.proper_pair)
        atomicAdd(&R.alt_pp[bs], 1u);
    }
    else if (cov != NO_COVERAGE)
    {
      atomicAdd(&R.gt_cov[(size_t)c0 * NS + (size_t)sample * cnum + cov], 1u);
      if (cov > 0 && geno.proper_pair)
        atomicAdd(&R.alt_pp[bs], 1u);
    }
  }
  return true;
}

// ReferenceDepth::add_genotype_paths (src/graph/reference_depth.cpp:114-203) on a per-sample difference array:
// +1 at the first covered index, -1 one past the last; one path = its span, several paths = the union of their
// (4-bp trimmed) spans.  Saturation (u16) is applied after the prefix sum on download.
__device__ void add_ref_depth(const LaunchParams & P, const DevRegion & R, const Geno & geno, int sample)
{
  if (geno.s.npaths == 0 || geno.s.longest < 63)
    return;
  GR g(R);
  long long const size = R.depth_size, offset = R.reference_offset;
  int * delta = R.ref_depth_delta + (size_t)sample * (size_t)(size + 1);
  const uint32_t * w = P.path_pool + geno.s.path_off;
  long long const RL = geno.read_length;
  if (geno.s.npaths == 1)
  {
    long long const sp = (long long)g.ref_reach_pos(w[0]) - (long long)(w[2] & 0xFFFFu);
    long long const ep = (long long)g.ref_reach_pos(w[1]) + (RL - 1 - (long long)(w[2] >> 16));
    long long const si = sp < offset ? 0 : sp - offset;
    long long ei = ep > offset + size ? size : ep + 1 - offset;
    if (si < size)
    {
      if (ei < si || ei > size)
        ei = size; // the reference's iterator loop runs to the end of the track in that case
      if (ei > si)
      {
        atomicAdd(delta + si, 1);
        atomicAdd(delta + ei, -1);
      }
    }
    return;
  }
  if (geno.s.npaths > MAXP)
  {
    // (huge tier) more paths than the sorted-interval buffer holds: union by repeated sweeps over the path records, no
    // storage.  Any exact decomposition of the union gives the same difference array.
    auto for_each_interval = [&](auto && fn) {
      const uint32_t * v = w;
      for (int pi = 0; pi < geno.s.npaths; ++pi)
      {
        uint32_t const nvar = v[3] >> 16;
        long long sp = (long long)g.ref_reach_pos(v[0]) - (long long)(v[2] & 0xFFFFu);
        long long ep = (long long)g.ref_reach_pos(v[1]) + (RL - 1 - (long long)(v[2] >> 16));
        v += PATH_HDR_WORDS + 2 * nvar;
        if (ep - sp >= 50)
        {
          sp += 4;
          ep -= 4;
        }
        if (ep < offset)
          continue;
        long long const si = sp < offset ? 0 : sp - offset;
        long long ei = ep > offset + size ? size : ep + 1 - offset;
        if (si >= size)
          continue;
        if (ei < si || ei > size)
          ei = size;
        if (ei > si)
          fn(si, ei);
      }
    };
    long long done = -1; // every index < done is handled
    for (;;)
    {
      long long cs = -1, ce = -1;
      for_each_interval([&](long long si, long long ei) {
        if (ei <= done)
          return;
        si = max(si, done);
        if (cs < 0 || si < cs)
        {
          cs = si;
          ce = ei;
        }
      });
      if (cs < 0)
        return;
      bool grew = true;
      while (grew)
      {
        grew = false;
        for_each_interval([&](long long si, long long ei) {
          if (si <= ce && ei > ce)
          {
            ce = ei;
            grew = true;
          }
        });
      }
      atomicAdd(delta + cs, 1);
      atomicAdd(delta + ce, -1);
      done = ce;
    }
  }
  long long iv[MAXP][2];
  int n = 0;
  for (int pi = 0; pi < geno.s.npaths && n < MAXP; ++pi)
  {
    uint32_t const nvar = w[3] >> 16;
    long long sp = (long long)g.ref_reach_pos(w[0]) - (long long)(w[2] & 0xFFFFu);
    long long ep = (long long)g.ref_reach_pos(w[1]) + (RL - 1 - (long long)(w[2] >> 16));
    w += PATH_HDR_WORDS + 2 * nvar;
    if (ep - sp >= 50)
    {
      sp += 4;
      ep -= 4;
    }
    if (ep < offset)
      continue;
    long long const si = sp < offset ? 0 : sp - offset;
    long long ei = ep > offset + size ? size : ep + 1 - offset;
    if (si >= size)
      continue;
    if (ei < si || ei > size)
      ei = size;
    if (ei <= si)
      continue;
    // insertion sort by start
    int k = n++;
    while (k > 0 && iv[k - 1][0] > si)
    {
      iv[k][0] = iv[k - 1][0];
      iv[k][1] = iv[k - 1][1];
      --k;
    }
    iv[k][0] = si;
    iv[k][1] = ei;
  }
  if (n == 0)
    return;
  long long cs = iv[0][0], ce = iv[0][1];
  for (int k = 1; k < n; ++k)
  {
    if (iv[k][0] <= ce)
      ce = max(ce, iv[k][1]);
    else
    {
      atomicAdd(delta + cs, 1);
      atomicAdd(delta + ce, -1);
      cs = iv[k][0];
      ce = iv[k][1];
    }
  }
  atomicAdd(delta + cs, 1);
  atomicAdd(delta + ce, -1);
}

// GenotypePaths pair of one record as update_paths leaves it
__device__ void make_genos(const LaunchParams & P, int rec, Geno & first, Geno & second)
{
  int const unit = P.batch.unit[rec];
  uint16_t const flag = P.batch.flag[rec];
  first.s = P.summaries[2 * unit];
  second.s = P.summaries[2 * unit + 1];
  first.read_length = second.read_length = P.batch.lseq[rec];
  first.mapq = second.mapq = P.batch.mapq[rec];
  first.score_diff = second.score_diff = P.batch.score_diff[rec];
  int32_t const isz = P.batch.isize[rec];
  bool const pp = (isz < 0 ? -(long long)isz : (long long)isz) != 0x7FFFFFFFll;
  first.proper_pair = second.proper_pair = pp;
  first.flags = flag & ~F_PROPER;
  if (P.batch.mapq[rec] < 25)
    first.flags |= F_MAPQ_BAD;
  second.flags = (flag ^ F_REV) & ~F_PROPER; // note: no IS_MAPQ_BAD on the flipped orientation (alignment.cpp:520)
  if (P.batch.clipped[rec])
  {
    first.flags |= 8192;
    second.flags |= 8192;
  }
}
} // namespace

// CONN: also build the phasing connections (regions whose conn_mask is 0 skip them at run time)
template <bool CONN>
__device__ void score_record(const LaunchParams & P, uint32_t i);

// first pass: one thread per record of a chunk
template <bool CONN>
__global__ void __launch_bounds__(128) score_kernel(LaunchParams P)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.batch.n_records)
    return;
  if (P.defer)
  {
    // anything this record needs still queued for the slower tiers?  (own alignment unit, and the mate's)
    int const u = P.batch.unit[i];
    int32_t const m = P.batch.mate[i];
    uint32_t pend = (uint32_t)P.pending[2 * u] | (uint32_t)P.pending[2 * u + 1];
    if (m >= 0)
    {
      int const um = P.batch.unit[m];
      pend |= (uint32_t)P.pending[2 * um] | (uint32_t)P.pending[2 * um + 1];
    }
    if (pend)
    {
      P.deferred[atomicAdd(&P.counters->n_deferred, 1u)] = i;
      return;
    }
  }
  score_record<CONN>(P, i);
}

// second pass: the deferred records of all chunks of a submit, after slow_kernel and huge_kernel
template <bool CONN>
__global__ void __launch_bounds__(128) score_deferred_kernel(const __grid_constant__ MultiLaunch M, uint32_t conn_chunks)
{
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  for (int c = 0; c < M.n; ++c)
  {
    uint32_t const n = M.p[c].counters->n_deferred;
    if (j < n)
    {
      if (CONN && ((conn_chunks >> c) & 1u))
        score_record<true>(M.p[c], M.p[c].deferred[j]);
      else
        score_record<false>(M.p[c], M.p[c].deferred[j]);
      return;
    }
    j -= n;
  }
}

template <bool CONN>
__device__ void score_record(const LaunchParams & P, uint32_t const i)
{
  uint16_t const flag = P.batch.flag[i];
  int32_t const m = P.batch.mate[i];
  const DevRegion & R = P.regions[P.batch.region[i]];
  int const sample = P.batch.sample[i];
  bool const conn = CONN && R.conn_mask != 0;

  if (m >= 0)
  {
    Geno g[4]; // prev.first, prev.second, cur.first, cur.second
    make_genos(P, m, g[0], g[1]);
    make_genos(P, (int)i, g[2], g[3]);
    if (((g[0].s.bits | g[1].s.bits | g[2].s.bits | g[3].s.bits) & TS_OVERFLOW) != 0)
      return; // counted by the align kernel; the host reports GTB_ERR_CAPACITY
    if ((g[2].flags & F_FIRST) == (g[0].flags & F_FIRST))
    {
      atomicAdd(&P.counters->n_input_error, 1ull);
      return;
    }
    // get_better_paths (alignment.cpp:557-622)
    int arr[4] = {-1, -1, -1, -1};
    for (int k = 0; k < 4; ++k)
      arr[((g[k].flags & F_FIRST) != 0) + 2 * ((g[k].flags & F_REV) == 0)] = k;
    if (arr[0] < 0 || arr[1] < 0 || arr[2] < 0 || arr[3] < 0)
      return;
    int const c = compare_pairs(g[arr[3]], g[arr[0]], g[arr[1]], g[arr[2]]);
    int s1 = -1, s2 = -1;
    if (c == 1)
    {
      s1 = arr[3];
      s2 = arr[0];
    }
    else if (c == 2)
    {
      s1 = arr[1];
      s2 = arr[2];
    }
    if (s1 < 0)
      return;
    if (R.is_sv) // hts_parallel_reader.cpp:321-326
    {
      add_ref_depth(P, R, g[s1], sample);
      add_ref_depth(P, R, g[s2], sample);
    }
    bool ok = true;
    if (CONN && conn)
    {
      TouchList k1, k2;
      k1.n = k2.n = 0;
      if (geno_good(R, g[s1]))
        ok = push_scores(P, R, g[s1], sample, &k1) && ok;
      if (geno_good(R, g[s2]))
        ok = push_scores(P, R, g[s2], sample, &k2) && ok;
      if (ok)
        conn_merge_mates(R, sample, k1, k2);
    }
    else
    {
      if (geno_good(R, g[s1]))
        ok = push_scores(P, R, g[s1], sample) && ok;
      if (geno_good(R, g[s2]))
        ok = push_scores(P, R, g[s2], sample) && ok;
    }
    if (!ok)
      atomicAdd(&P.counters->n_overflow, 1ull);
    atomicAdd(&P.counters->n_pairs_scored, 1ull);
  }
  else if (R.is_sv && P.batch.leftover[i])
  {
    // leftover mate at pool end (SV calling only, hts_parallel_reader.cpp:719-772): paired against a copy of itself
    // with IS_FIRST_IN_PAIR and IS_SEQ_REVERSED toggled; the first-in-pair side of the better pairing is scored alone
    Geno g[4];
    make_genos(P, (int)i, g[0], g[1]);
    if (((g[0].s.bits | g[1].s.bits) & TS_OVERFLOW) != 0)
      return;
    g[2] = g[0];
    g[3] = g[1];
    g[2].flags ^= (F_FIRST | F_REV);
    g[3].flags ^= (F_FIRST | F_REV);
    int arr[4] = {-1, -1, -1, -1};
    for (int k = 0; k < 4; ++k)
      arr[((g[k].flags & F_FIRST) != 0) + 2 * ((g[k].flags & F_REV) == 0)] = k;
    if (arr[0] < 0 || arr[1] < 0 || arr[2] < 0 || arr[3] < 0)
      return;
    int const c = compare_pairs(g[arr[3]], g[arr[0]], g[arr[1]], g[arr[2]]);
    int const s1 = c == 1 ? arr[3] : c == 2 ? arr[1] : -1;
    if (s1 < 0)
      return;
    add_ref_depth(P, R, g[s1], sample);
    if (geno_good(R, g[s1]))
    {
      TouchList k1;
      if (!push_scores(P, R, g[s1], sample, (CONN && conn) ? &k1 : nullptr))
        atomicAdd(&P.counters->n_overflow, 1ull);
      atomicAdd(&P.counters->n_singles_scored, 1ull);
    }
  }
  else if ((flag & F_PAIRED) == 0)
  {
    // update_unpaired_read_paths (alignment.cpp:365-449) + compare (genotype_paths.cpp:943-974)
    Geno a, b;
    make_genos(P, (int)i, a, b);
    if (((a.s.bits | b.s.bits) & TS_OVERFLOW) != 0)
      return;
    a.proper_pair = b.proper_pair = false; // ml_insert_size stays INSERT_SIZE_WHEN_NOT_PROPER_PAIR
    uint32_t const T1 = a.s.longest, T2 = b.s.longest; // longest_path_size() (0 when there are no paths)
    int c = 0;
    if (T1 > T2 && T1 > 94)
      c = 1;
    else if (T2 > T1 && T2 > 94)
      c = 2;
    else if (T2 == T1 && T1 > 94)
      c = (b.s.mm0 < a.s.mm0) ? 2 : 1;
    if (c == 0)
      return;
    Geno & sel = c == 1 ? a : b;
    if (c == 2 && P.batch.mapq[i] < 25)
      sel.flags |= F_MAPQ_BAD; // the unpaired path does set IS_MAPQ_BAD on the flipped orientation (alignment.cpp:419)
    if (geno_good(R, sel))
    {
      TouchList k1;
      if (!push_scores(P, R, sel, sample, (CONN && conn) ? &k1 : nullptr))
        atomicAdd(&P.counters->n_overflow, 1ull);
      atomicAdd(&P.counters->n_singles_scored, 1ull);
    }
  }
}

// ================================================================================================ segment gather / zero
__global__ void __launch_bounds__(256) gather_segments_kernel(const __grid_constant__ SegmentTable tab, uint8_t * dst)
{
  Segment const & sg = tab.s[blockIdx.y];
  const uint4 * src = static_cast<const uint4 *>(sg.ptr);
  uint4 * out = reinterpret_cast<uint4 *>(dst + sg.dst_off);
  size_t const n = sg.bytes / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = src[i];
}

__global__ void __launch_bounds__(256) zero_segments_kernel(const __grid_constant__ SegmentTable tab)
{
  Segment const & sg = tab.s[blockIdx.y];
  uint4 * out = static_cast<uint4 *>(sg.ptr);
  size_t const n = sg.bytes / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_uint4(0, 0, 0, 0);
}

static dim3 segment_grid(int n, unsigned long long max_bytes)
{
  unsigned long long const per_block = 256ull * 16 * 4; // four 16-byte words per thread
  unsigned const gx = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>((max_bytes + per_block - 1) / per_block, 1), 1024);
  return dim3(gx, (unsigned)n, 1);
}

void launch_gather_segments(const Segment * seg, int n, unsigned long long max_bytes, void * dst, void * stream)
{
  for (int at = 0; at < n; at += SEGMENTS_PER_LAUNCH)
  {
    int const m = std::min(SEGMENTS_PER_LAUNCH, n - at);
    SegmentTable tab;
    memcpy(tab.s, seg + at, (size_t)m * sizeof(Segment));
    gather_segments_kernel<<<segment_grid(m, max_bytes), 256, 0, (cudaStream_t)stream>>>(tab, static_cast<uint8_t *>(dst));
  }
}

void launch_zero_segments(const Segment * seg, int n, unsigned long long max_bytes, void * stream)
{
  for (int at = 0; at < n; at += SEGMENTS_PER_LAUNCH)
  {
    int const m = std::min(SEGMENTS_PER_LAUNCH, n - at);
    SegmentTable tab;
    memcpy(tab.s, seg + at, (size_t)m * sizeof(Segment));
    zero_segments_kernel<<<segment_grid(m, max_bytes), 256, 0, (cudaStream_t)stream>>>(tab);
  }
}

// ================================================================================================ column gather (zero copy)
// One column of one tile: m elements from mapped host memory to the chunk's device block, f applied to every element.
// 16-byte loads wherever the source allows it (a PCIe read request per 16 bytes of a lane, 512 bytes per warp), several of
// them in flight per thread; the destination is written with 16-byte stores when it happens to be aligned as well.
template <typename T, typename F>
__device__ __forceinline__ void gather_col(const T * __restrict__ src, T * __restrict__ dst, uint32_t m, F f)
{
  constexpr uint32_t PER = 16 / sizeof(T);
  uint32_t head = (uint32_t)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) / sizeof(T));
  head = head < m ? head : m;
  for (uint32_t i = threadIdx.x; i < head; i += blockDim.x)
    dst[i] = f(src[i]);
  uint32_t const nv = (m - head) / PER;
  const uint4 * sv = reinterpret_cast<const uint4 *>(src + head);
  T * d = dst + head;
  bool const aligned_dst = (reinterpret_cast<uintptr_t>(d) & 15u) == 0;
#pragma unroll 4
  for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x)
  {
    union
    {
      uint4 v;
      T e[PER];
    } u;
    u.v = sv[i];
#pragma unroll
    for (uint32_t j = 0; j < PER; ++j)
      u.e[j] = f(u.e[j]);
    if (aligned_dst)
      reinterpret_cast<uint4 *>(d)[i] = u.v;
    else
    {
#pragma unroll
      for (uint32_t j = 0; j < PER; ++j)
        d[i * PER + j] = u.e[j];
    }
  }
  for (uint32_t i = head + nv * PER + threadIdx.x; i < m; i += blockDim.x)
    dst[i] = f(src[i]);
}

template <typename T>
__device__ __forceinline__ void fill_col(T * dst, uint32_t m, T value)
{
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
    dst[i] = value;
}

__global__ void __launch_bounds__(512) gather_columns_kernel(const __grid_constant__ ColumnGather g)
{
  for (uint32_t w = blockIdx.x; w < g.n_tiles; w += gridDim.x)
  {
    uint32_t ji = 0;
    while (ji + 1 < g.n_jobs && g.job[ji + 1].tile_begin <= w)
      ++ji;
    ColumnJob const & J = g.job[ji];
    uint32_t const k0 = (w - J.tile_begin) * GATHER_TILE;
    uint32_t const m = J.n - k0 < GATHER_TILE ? J.n - k0 : GATHER_TILE;
    size_t const at = (size_t)J.rec_base + k0;
    auto same = [](auto v) { return v; };
    int32_t const rb = (int32_t)J.rec_base;
    auto rebase = [rb](int32_t v) { return v < 0 ? -1 : rb + v; };
    gather_col(J.lseq + k0, reinterpret_cast<uint16_t *>(g.dst + g.o_lseq) + at, m, same);
    gather_col(J.flag + k0, reinterpret_cast<uint16_t *>(g.dst + g.o_flag) + at, m, same);
    gather_col(J.mapq + k0, g.dst + g.o_mapq + at, m, same);
    gather_col(J.same_tid + k0, g.dst + g.o_same + at, m, same);
    gather_col(J.score_diff + k0, g.dst + g.o_sd + at, m, same);
    if (J.clipped)
      gather_col(J.clipped + k0, g.dst + g.o_clip + at, m, same);
    else
      fill_col<uint8_t>(g.dst + g.o_clip + at, m, 0);
    if (J.leftover)
      gather_col(J.leftover + k0, g.dst + g.o_left + at, m, same);
    else
      fill_col<uint8_t>(g.dst + g.o_left + at, m, 0);
    gather_col(J.isize + k0, reinterpret_cast<int32_t *>(g.dst + g.o_isize) + at, m, same);
    gather_col(J.sample + k0, reinterpret_cast<int32_t *>(g.dst + g.o_sample) + at, m, same);
    // links are batch-local; a link at or beyond its own record stays >= its chunk-global index and is reported by
    // prep_flags_kernel
    if (J.mate)
      gather_col(J.mate + k0, reinterpret_cast<int32_t *>(g.dst + g.o_mate) + at, m, rebase);
    else
      fill_col<int32_t>(reinterpret_cast<int32_t *>(g.dst + g.o_mate) + at, m, -1);
    if (J.dup_of)
      gather_col(J.dup_of + k0, reinterpret_cast<int32_t *>(g.dst + g.o_dup) + at, m, rebase);
    else
      fill_col<int32_t>(reinterpret_cast<int32_t *>(g.dst + g.o_dup) + at, m, -1);
    fill_col<uint16_t>(reinterpret_cast<uint16_t *>(g.dst + g.o_region) + at, m, (uint16_t)J.slot);
  }
}

// A few fat blocks: the kernel waits on PCIe reads, not on issue slots, and every SM it sits on is an SM the persistent
// probe_kernel of another pool cannot start on.  24 blocks x 512 threads x 4 loads of 16 bytes = 0.8 MB in flight.
void launch_gather_columns(const ColumnGather & g, void * stream)
{
  if (g.n_tiles == 0)
    return;
  unsigned const blocks = std::min<unsigned>(g.n_tiles, 24u);
  gather_columns_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(g);
}

// ================================================================================================ connection table upkeep
__global__ void __launch_bounds__(256) conn_rehash_kernel(const unsigned long long * keys, const uint32_t * vals, uint32_t n_slots,
                                                          DevRegion R)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots && keys[i] != 0)
    conn_insert(R, keys[i], vals[i]);
}

__global__ void __launch_bounds__(256) conn_compact_kernel(DevRegion R, unsigned long long * out_keys, uint32_t * out_vals,
                                                           uint32_t * out_n)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > R.conn_mask || R.conn_keys[i] == 0)
    return;
  uint32_t const o = atomicAdd(out_n, 1u);
  out_keys[o] = R.conn_keys[i];
  out_vals[o] = R.conn_vals[i];
}

// ================================================================================================ table build
// Inserts the region's distinct k-mers {key, label offset, count} into the zero-initialised open-addressing table.
// A slot is claimed by CAS on its (offset, count) word -- count == 0 means empty -- then the key is written;
// lookups only happen in later kernels.
__global__ void __launch_bounds__(256) build_table_kernel(const IndexSlot * uniq, uint32_t n, IndexSlot * table,
                                                          uint32_t mask, int shift, uint32_t * bitmap)
{
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  IndexSlot const u = uniq[i];
  unsigned long long const val = (unsigned long long)u.off | ((unsigned long long)u.cnt << 32);
  uint32_t const hs = hash32_tab(g_hash_tab, u.key);
  uint32_t const bi = hs >> (shift - 34); // presence bitmap: 4 bits per table slot
  atomicOr(&bitmap[bi >> 5], 1u << (bi & 31u));
  uint32_t h = hs >> (shift - 32);
  while (true)
  {
    unsigned long long * w = reinterpret_cast<unsigned long long *>(&table[h]) + 1;
    if (atomicCAS(w, 0ull, val) == 0ull)
    {
      table[h].key = u.key;
      return;
    }
    h = (h + 1) & mask;
  }
}

void launch_build_table(const IndexSlot * uniq, uint32_t n, IndexSlot * table, uint32_t mask, int shift, uint32_t * bitmap,
                        void * stream)
{
  if (n == 0)
    return;
  build_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(uniq, n, table, mask, shift, bitmap);
}

// ================================================================================================ launchers
static int g_align_blocks_per_sm = 0;

int align_kernel_blocks_per_sm()
{
  if (g_align_blocks_per_sm == 0)
  {
    size_t const smem = sizeof(SlowState) * WARPS_PER_BLOCK;
    cudaFuncSetAttribute(slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, slow_kernel, WARPS_PER_BLOCK * 32, smem) != cudaSuccess || nb < 1)
      nb = 1;
    g_align_blocks_per_sm = nb;
  }
  return g_align_blocks_per_sm;
}

static int sm_count()
{
  static int sms = 0;
  if (sms == 0)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

size_t align_spill_bytes()
{
  return (size_t)sm_count() * align_kernel_blocks_per_sm() * WARPS_PER_BLOCK * CAND_SPILL * sizeof(SlowState::Cand);
}

void launch_probe(const LaunchParams & p, void * stream)
{
  if (p.n_active == 0)
    return;
  // one persistent block per SM (p.n_active is an upper bound; the exact count is read on the device); small batches
  // get fewer blocks so that a block still has a few hundred tasks to amortise staging its region's filter
  {
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev])
    {
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FILTER_WORDS * 4));
      if (dev >= 0 && dev < 64)
        attr_set[dev] = true;
    }
  }
  uint32_t const want = (p.n_active + 255) / 256;
  uint32_t const grid = std::max(1u, std::min(want, (uint32_t)sm_count()));
  LaunchParams q = p;
  const char * e = getenv("GTB_PROBE_FILTER_FOLD"); // tests: -1 forces the global-bitmap path on every region
  q.filter_max_fold = e ? std::min(atoi(e), FILTER_MAX_FOLD) : FILTER_MAX_FOLD;
  probe_kernel<<<grid, PROBE_BLOCK_WARPS * 32, FILTER_WORDS * 4, (cudaStream_t)stream>>>(q);
}

void launch_chain(const LaunchParams & p, void * stream)
{
  if (p.n_active == 0)
    return;
  uint32_t const grid = (p.n_active + CHAIN_THREADS - 1) / CHAIN_THREADS;
  size_t const smem = (size_t)CHAIN_THREADS * FT_WORDS * 4;
  {
    static bool attr_set[64] = {false}; // > 48 KB of dynamic shared memory is a per-device opt-in
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev])
    {
      cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (dev >= 0 && dev < 64)
        attr_set[dev] = true;
    }
  }
  chain_kernel<<<grid, CHAIN_THREADS, smem, (cudaStream_t)stream>>>(p);
}

// Grid for the upper bound of queued tasks (the exact count is read on the device; blocks beyond it return at once).
void launch_chain_general(const LaunchParams & p, void * stream)
{
  if (p.n_active == 0)
    return;
  uint32_t const lanes = std::max(1u, std::min(32u, p.gen_lanes));
  uint32_t const warps = (p.n_active + lanes - 1) / lanes;
  uint32_t const grid = (warps + CHAIN_THREADS / 32 - 1) / (CHAIN_THREADS / 32);
  LaunchParams q = p;
  q.gen_lanes = lanes;
  chain_general_kernel<<<grid, CHAIN_THREADS, 0, (cudaStream_t)stream>>>(q);
}

// persistent grid (a multiple of the SM count); the number of queued tasks is read on the device
void launch_slow(const MultiLaunch & m, int which, void * stream)
{
  bool any = false;
  for (int c = 0; c < m.n; ++c)
    any = any || m.p[c].n_active != 0;
  if (!any)
    return;
  size_t const smem = sizeof(SlowState) * WARPS_PER_BLOCK;
  uint32_t const grid = (uint32_t)(sm_count() * align_kernel_blocks_per_sm());
  {
    // > 48 KB of dynamic shared memory is a per-device opt-in (align_kernel_blocks_per_sm() sets it on the first device only)
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev])
    {
      cudaFuncSetAttribute(slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (dev >= 0 && dev < 64)
        attr_set[dev] = true;
    }
  }
  slow_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(m, which);
  if (which)
    huge_kernel<<<(uint32_t)sm_count(), 32, 0, (cudaStream_t)stream>>>(m);
}

size_t huge_state_bytes() { return (size_t)sm_count() * sizeof(HugeState); }

void launch_score(const LaunchParams & p, bool with_connections, void * stream)
{
  if (p.batch.n_records == 0)
    return;
  uint32_t const grid = (p.batch.n_records + 127) / 128;
  if (with_connections)
    score_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  else
    score_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_score_deferred(const MultiLaunch & m, const bool * with_connections, void * stream)
{
  // grid bound: slow tasks are rare; the kernel reads the exact counts on the device.  One block per 128 records of the
  // largest plausible backlog (1/64 of the records, at least 4 blocks); a backlog beyond that is handled by more passes.
  uint32_t total = 0, conn_chunks = 0;
  for (int c = 0; c < m.n; ++c)
  {
    total += m.p[c].batch.n_records;
    if (with_connections[c])
      conn_chunks |= 1u << c;
  }
  if (total == 0)
    return;
  uint32_t const grid = (total + 127) / 128;
  if (conn_chunks)
    score_deferred_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(m, conn_chunks);
  else
    score_deferred_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(m, conn_chunks);
}

void launch_conn_rehash(const unsigned long long * keys, const uint32_t * vals, uint32_t n_slots, const DevRegion & R, void * stream)
{
  if (n_slots)
    conn_rehash_kernel<<<(n_slots + 255) / 256, 256, 0, (cudaStream_t)stream>>>(keys, vals, n_slots, R);
}

void launch_conn_compact(const DevRegion & R, unsigned long long * out_keys, uint32_t * out_vals, uint32_t * out_n, void * stream)
{
  if (R.conn_mask)
    conn_compact_kernel<<<(R.conn_mask + 256) / 256, 256, 0, (cudaStream_t)stream>>>(R, out_keys, out_vals, out_n);
}

} // namespace gtb
