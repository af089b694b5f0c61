// oracle/gtb_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's algorithm).
//
// Scalar, single-threaded restatement of graphtyper's genotyping hot path on the flat views of
// include/gtb200.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this;
// the product (graphtyper_b200/csrc) never links or calls it.
//
// PARITY PINNING: the reference's own unit tests pin only the k-mer codec and index contents
// (test/index/test_index.cpp, test/utilities/test_kmer_help_functions.cpp); alignment, path scoring and
// likelihoods are unpinned there.  This oracle is therefore pinned against outputs of the
// reference ITSELF compiled here (oracle/_ref/bin/gt_probe, oracle/ref_build/): graph index contents,
// per-seed label lists, per-read GenotypePaths, per-bubble accumulators, SampleCall PL/GT/GQ -- see
// tests/test_oracle_vs_reference.py and the committed vectors under tests/golden/.
//
// Each function cites the reference file:line it follows (paths relative to /root/reference).

#include <algorithm>
#include <array>
#include <numeric>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/gtb200.h"

namespace
{
constexpr uint32_t K = 32;
constexpr uint32_t INVALID_ID = 0xFFFFFFFFu;
constexpr uint32_t SPECIAL_START = 0xD0000000u;
constexpr uint32_t MAX_UNIQUE_KMER_POSITIONS = 512;     // constants.hpp.in:39
constexpr uint32_t MAX_SEED_NUMBER_ALLOWING_MISMATCHES = 64;
constexpr uint32_t MAX_SEED_NUMBER_FOR_WALKING = 256;
constexpr uint32_t MAX_NUM_LOCATIONS_PER_PATH = 256;
constexpr long MAX_INDEX_LABELS = 75;                   // options.hpp:82
constexpr uint16_t IS_PAIRED = 1, IS_PROPER_PAIR = 2, IS_UNMAPPED = 4, IS_SEQ_REVERSED = 16,
                   IS_MATE_SEQ_REVERSED = 32, IS_FIRST_IN_PAIR = 64, IS_MAPQ_BAD = 4096, IS_CLIPPED = 8192;
constexpr int32_t INSERT_SIZE_WHEN_NOT_PROPER_PAIR = 0x7FFFFFFF;

std::string g_err;

// ------------------------------------------------------------------------------------------------ graph
struct G
{
  gtb_graph_view v;
  uint32_t ref_len(uint32_t r) const { return (uint32_t)(v.ref_seq_off[r + 1] - v.ref_seq_off[r]); }
  uint32_t var_len(uint32_t x) const { return (uint32_t)(v.var_seq_off[x + 1] - v.var_seq_off[x]); }
  const uint8_t * ref_dna(uint32_t r) const { return v.seq + v.ref_seq_off[r]; }
  const uint8_t * var_dna(uint32_t x) const { return v.seq + v.var_seq_off[x]; }
  uint32_t ref_reach(uint32_t r) const { return v.ref_order[r] + ref_len(r) - 1; } // label.cpp:37-40
  uint32_t var_reach(uint32_t x) const { return v.var_order[x] + var_len(x) - 1; }
  uint32_t out_degree(uint32_t r) const { return v.ref_var_off[r + 1] - v.ref_var_off[r]; }
  // graph.cpp:341-345
  uint16_t variant_num(uint32_t x) const { return (uint16_t)(x - v.ref_var_off[v.var_out_ref[x] - 1]); }
  // reach of the reference allele of x's bubble
  uint32_t bubble_ref_reach(uint32_t x) const { return var_reach(v.ref_var_off[v.var_out_ref[x] - 1]); }
  bool is_special_pos(uint32_t p) const { return p >= SPECIAL_START && (p - SPECIAL_START) < v.n_special; }
  uint32_t get_ref_reach_pos(uint32_t p) const { return is_special_pos(p) ? v.ref_reach_poses[p - SPECIAL_START] : p; }
  uint32_t get_actual_pos(uint32_t p) const { return is_special_pos(p) ? v.actual_poses[p - SPECIAL_START] : p; }
  // graph.cpp:1775-1782
  uint32_t get_special_pos(uint32_t pos, uint32_t ref_reach) const
  {
    const uint32_t * b = v.sp_keys;
    const uint32_t * e = v.sp_keys + v.n_sp_keys;
    const uint32_t * it = std::lower_bound(b, e, ref_reach);
    if (it == e || *it != ref_reach)
      return INVALID_ID; // the reference would throw (unordered_map::at)
    uint32_t k = (uint32_t)(it - b);
    uint32_t idx = pos - ref_reach - 1;
    if (v.sp_off[k] + idx >= v.sp_off[k + 1])
      return INVALID_ID;
    return v.sp_list[v.sp_off[k] + idx];
  }
};

// ------------------------------------------------------------------------------------------------ index (A1)
struct Label
{
  uint32_t start, end, var;
};

using Index = std::unordered_map<uint64_t, std::vector<Label>>;

// include/graphtyper/index/index_entry.hpp:17-33 (valid is always 0: non-ACGT bases never reach add_to_dna)
struct Entry
{
  uint64_t dna = 0;
  uint32_t start = 0;
  std::vector<uint32_t> var_ids; // std::set<uint32_t> semantics: kept sorted unique
  uint32_t total_var_num = 1;
  uint32_t total_var_count = 0;
  std::vector<int64_t> events, anti_events; // unordered_set semantics: kept sorted unique
};
using Sub = std::vector<Entry>;
using List = std::deque<Sub>;

inline void sorted_insert(std::vector<uint32_t> & v, uint32_t x)
{
  auto it = std::lower_bound(v.begin(), v.end(), x);
  if (it == v.end() || *it != x)
    v.insert(it, x);
}
inline void sorted_insert64(std::vector<int64_t> & v, int64_t x)
{
  auto it = std::lower_bound(v.begin(), v.end(), x);
  if (it == v.end() || *it != x)
    v.insert(it, x);
}
inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
inline void add_to_dna(Entry & e, uint8_t base) // index_entry.cpp:20-54
{
  e.dna <<= 2;
  e.dna += base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : 3;
}

// indexer.cpp:26-81
void index_reference_label(const G & g, Index & idx, List & mers, uint32_t r)
{
  uint32_t const order = g.v.ref_order[r];
  uint32_t const n = g.ref_len(r);
  const uint8_t * dna = g.ref_dna(r);
  for (uint32_t d = 0; d < n; ++d)
  {
    uint8_t const b = dna[d];
    if (!is_acgt(b))
    {
      mers.clear();
      continue;
    }
    for (auto & sub : mers)
      for (auto & e : sub)
        add_to_dna(e, b);
    {
      Entry e;
      e.start = order + d;
      add_to_dna(e, b);
      mers.push_front(Sub(1, e));
    }
    if (mers.size() >= K)
    {
      for (auto const & q : mers.back())
      {
        auto & bucket = idx[q.dna];
        if (q.var_ids.empty())
          bucket.push_back({q.start, order + d, INVALID_ID});
        else
          for (uint32_t vid : q.var_ids)
            bucket.push_back({q.start, order + d, vid});
      }
      mers.pop_back();
    }
  }
}

// indexer.cpp:83-178
void insert_variant_label(const G & g, Index & idx, List & mers, uint32_t v, bool is_reference, unsigned var_count,
                          uint32_t ref_reach)
{
  uint32_t const order = g.v.var_order[v];
  uint32_t const n = g.var_len(v);
  const uint8_t * dna = g.var_dna(v);
  const int64_t * ev_b = g.v.var_ev ? g.v.var_ev + g.v.var_ev_off[v] : nullptr;
  const int64_t * ev_e = g.v.var_ev ? g.v.var_ev + g.v.var_ev_off[v + 1] : nullptr;
  const int64_t * aev_b = g.v.var_aev ? g.v.var_aev + g.v.var_aev_off[v] : nullptr;
  const int64_t * aev_e = g.v.var_aev ? g.v.var_aev + g.v.var_aev_off[v + 1] : nullptr;

  for (uint32_t d = 0; d < n; ++d)
  {
    uint8_t const b = dna[d];
    if (!is_acgt(b))
    {
      mers.clear();
      continue;
    }
    for (auto & sub : mers)
    {
      for (size_t k = 0; k < sub.size();)
      {
        Entry & e = sub[k];
        bool ok = true;
        for (int64_t ae : e.anti_events)
          if (ev_b && std::binary_search(ev_b, ev_e, ae))
          {
            ok = false;
            break;
          }
        if (ok)
        {
          add_to_dna(e, b);
          for (const int64_t * p = ev_b; p != ev_e; ++p)
            sorted_insert64(e.events, *p);
          for (const int64_t * p = aev_b; p != aev_e; ++p)
            sorted_insert64(e.anti_events, *p);
          sorted_insert(e.var_ids, v);
          ++k;
        }
        else
        {
          sub.erase(sub.begin() + k);
        }
      }
    }
    uint32_t pos = order + d;
    if (pos > ref_reach)
      pos = g.get_special_pos(pos, ref_reach);
    Entry ne;
    ne.start = pos;
    ne.total_var_num = var_count;
    ne.total_var_count = is_reference ? 0u : 1u;
    ne.var_ids.push_back(v);
    add_to_dna(ne, b);
    if (ev_b)
      ne.events.assign(ev_b, ev_e);
    if (aev_b)
      ne.anti_events.assign(aev_b, aev_e);
    mers.push_front(Sub(1, ne));
    if (mers.size() >= K)
    {
      for (auto const & q : mers.back())
      {
        auto & bucket = idx[q.dna];
        for (uint32_t vid : q.var_ids)
          bucket.push_back({q.start, pos, vid});
      }
      mers.pop_back();
    }
  }
}

// indexer.cpp:180-196
void append_list(List & mers, List && list)
{
  if (mers.size() < list.size())
    mers.resize(list.size());
  auto m = mers.begin();
  for (auto & sub : list)
  {
    for (auto & e : sub)
      m->push_back(std::move(e));
    ++m;
  }
}

// indexer.cpp:13-20,198-211
void remove_large_variants_from_list(List & list, unsigned var_count)
{
  for (auto & sub : list)
  {
    for (auto & e : sub)
    {
      e.total_var_num *= var_count;
      ++e.total_var_count;
    }
    sub.erase(std::remove_if(sub.begin(), sub.end(),
                             [](Entry const & e)
                             { return e.total_var_count > 1 && (e.total_var_num > 181u || e.total_var_count > 4u); }),
              sub.end());
  }
}

// indexer.cpp:213-244
void index_variant(const G & g, Index & idx, List & mers, unsigned var_count, uint32_t v)
{
  List clean(mers);
  uint32_t const ref_reach = g.var_reach(v);
  insert_variant_label(g, idx, mers, v, true, 1, ref_reach);
  remove_large_variants_from_list(clean, var_count);
  unsigned const var_num = var_count;
  while (var_count > 2)
  {
    --var_count;
    ++v;
    List nl(clean);
    insert_variant_label(g, idx, nl, v, false, var_num, ref_reach);
    append_list(mers, std::move(nl));
  }
  ++v;
  if (v < g.v.n_var)
  {
    insert_variant_label(g, idx, clean, v, false, var_num, ref_reach);
    append_list(mers, std::move(clean));
  }
}

// indexer.cpp:246-291
void index_graph(const G & g, Index & idx)
{
  List mers;
  for (uint32_t r = 0; r + 1 < g.v.n_ref; ++r)
  {
    index_reference_label(g, idx, mers, r);
    if (g.out_degree(r) > 0)
      index_variant(g, idx, mers, g.out_degree(r), g.v.ref_var_off[r]);
  }
  if (g.v.n_ref > 0)
    index_reference_label(g, idx, mers, g.v.n_ref - 1);
}

struct IndexHandle
{
  Index idx;
  std::vector<uint64_t> sorted_keys;
};

// ------------------------------------------------------------------------------------------------ seeds (A2, A3)
// src/utilities/type_conversions.cpp:207-266 ; s = 4-bit BAM codes (A1 C2 G4 T8 N15, '='0)
std::vector<uint64_t> to_uint64_vec(const uint8_t * s, size_t i)
{
  std::vector<uint64_t> u(1, 0);
  for (size_t const j = i + 32; i < j; ++i)
  {
    size_t const origin = u.size();
    if (origin > 97)
      return std::vector<uint64_t>();
    for (size_t k = 0; k < origin; ++k)
    {
      uint8_t const c = s[i];
      if (c == 15 || c == 0)
      {
        u.push_back(u[k] * 4 + 0);
        u.push_back(u[k] * 4 + 1);
        u.push_back(u[k] * 4 + 2);
        u[k] <<= 2;
        u[k] += 3;
      }
      else
      {
        int set_count = __builtin_popcount(c);
        for (int b = 0; b < 4; ++b)
          if (c & (1 << b))
          {
            if (set_count == 1)
            {
              u[k] *= 4;
              u[k] += b;
            }
            else
              u.push_back(u[k] * 4 + b);
            --set_count;
          }
      }
    }
  }
  return u;
}

// ph_index.cpp:66-107
std::vector<std::vector<Label>> multi_get(const Index & idx, const std::vector<std::vector<uint64_t>> & keys)
{
  std::vector<std::vector<Label>> labels(keys.size());
  for (size_t i = 0; i < keys.size(); ++i)
  {
    long num_results = 0;
    long const NK = (long)keys[i].size();
    std::vector<const std::vector<Label> *> res;
    for (long j = 0; j < NK; ++j)
    {
      auto it = idx.find(keys[i][j]);
      if (it != idx.end())
      {
        num_results += (long)it->second.size();
        if (NK > 1 && num_results > MAX_INDEX_LABELS)
        {
          res.clear();
          break;
        }
        res.push_back(&it->second);
      }
    }
    for (auto p : res)
      labels[i].insert(labels[i].end(), p->begin(), p->end());
  }
  return labels;
}

inline size_t get_num_kmers(size_t len) { return len < K ? 0 : 1 + (len - K) / (K - 1); } // kmer_help_functions.cpp:10-17

// kmer_help_functions.cpp:51-61
std::vector<std::vector<Label>> query_index(const Index & idx, const uint8_t * s, size_t len)
{
  std::vector<std::vector<uint64_t>> mk;
  size_t const n = get_num_kmers(len);
  for (size_t i = 0; i < n; ++i)
    mk.push_back(to_uint64_vec(s, (K - 1) * i));
  return multi_get(idx, mk);
}

// kmer_help_functions.cpp:93-119 ; type_conversions.cpp:272-288
std::vector<std::vector<Label>> query_index_ham1(const Index & idx, const uint8_t * s, size_t len)
{
  std::vector<std::vector<uint64_t>> mk;
  size_t const n = get_num_kmers(len);
  for (size_t i = 0; i < n; ++i)
    mk.push_back(to_uint64_vec(s, (K - 1) * i));
  for (size_t i = 0; i < n; ++i)
  {
    if (mk[i].size() != 1)
      continue;
    uint64_t const key = mk[i][0];
    mk[i].clear();
    for (unsigned bb = 0; bb < 32; ++bb)
      for (uint64_t j = 1; j <= 3; ++j)
        mk[i].push_back((j << (bb * 2)) ^ key);
  }
  return multi_get(idx, mk);
}

// ------------------------------------------------------------------------------------------------ paths (A6)
struct Path // include/graphtyper/typer/path.hpp:18-79
{
  uint32_t start = 0, end = 0;
  uint16_t rs = 0, re = 0;
  std::vector<uint32_t> var_order;
  std::vector<std::set<uint16_t>> nums;
  uint16_t mm = 0;
  uint32_t size() const { return (uint32_t)re - rs + 1u; } // path.cpp:165-169
  bool is_reference() const
  {
    for (auto const & n : nums)
      if (n.count(0) == 0)
        return false;
    return true;
  }
  bool is_empty() const { return start == end; }
};

Path make_path(const G & g, const Label & l, uint16_t rs, uint16_t re, uint16_t mm) // path.cpp:13-36
{
  Path p;
  p.start = l.start;
  p.end = l.end;
  p.rs = rs;
  p.re = re;
  p.mm = mm;
  if (l.var != INVALID_ID)
  {
    p.var_order.push_back(g.v.var_order[l.var]);
    p.nums.push_back({g.variant_num(l.var)});
  }
  return p;
}

Path merge_paths(const Path & p1, const Path & p2) // path.cpp:38-82
{
  Path r(p2);
  for (size_t i = 0; i < p1.var_order.size(); ++i)
  {
    bool found = false;
    for (size_t j = 0; j < r.var_order.size(); ++j)
    {
      if (p1.var_order[i] == r.var_order[j])
      {
        for (auto it = r.nums[j].begin(); it != r.nums[j].end();)
        {
          if (p1.nums[i].count(*it) == 0)
            it = r.nums[j].erase(it);
          else
            ++it;
        }
        if (r.nums[j].empty())
          return r;
        found = true;
        break;
      }
    }
    if (!found)
    {
      r.var_order.push_back(p1.var_order[i]);
      r.nums.push_back(p1.nums[i]);
    }
  }
  r.rs = p1.rs;
  r.start = p1.start;
  r.mm += p1.mm;
  return r;
}

void merge_with_current(const G & g, Path & p, const Label & l) // path.cpp:105-129
{
  if (l.var == INVALID_ID)
    return;
  uint32_t const vo = g.v.var_order[l.var];
  uint16_t const vn = g.variant_num(l.var);
  for (size_t i = 0; i < p.var_order.size(); ++i)
    if (p.var_order[i] == vo)
    {
      p.nums[i].insert(vn);
      return;
    }
  p.var_order.push_back(vo);
  p.nums.push_back({vn});
}

// genotype_paths.cpp:32-66
std::vector<Path> find_all_nonduplicated_paths(const G & g, const std::vector<Label> & ll, uint16_t rs, uint16_t re,
                                               uint16_t mm)
{
  std::vector<Path> paths;
  if (ll.empty())
    return paths;
  paths.push_back(make_path(g, ll[0], rs, re, mm));
  for (size_t i = 1; i < ll.size(); ++i)
  {
    bool nothing = true;
    for (auto & p : paths)
      if (ll[i].start == p.start && ll[i].end == p.end)
      {
        merge_with_current(g, p, ll[i]);
        nothing = false;
        break;
      }
    if (nothing)
      paths.push_back(make_path(g, ll[i], rs, re, mm));
  }
  return paths;
}

struct GenoPaths // include/graphtyper/typer/genotype_paths.hpp:27-114
{
  std::vector<Path> paths;
  uint16_t read_length = 0;
  uint16_t flags = 0;
  uint32_t longest = 0;
  uint8_t score_diff = 0;
  uint8_t mapq = 255;
  int32_t ml_insert_size = INSERT_SIZE_WHEN_NOT_PROPER_PAIR;

  // genotype_paths.cpp:294-352
  void add_next(const G & g, const std::vector<Label> & ll, uint32_t rs, uint32_t re, int mm)
  {
    std::vector<Path> const pp = find_all_nonduplicated_paths(g, ll, (uint16_t)rs, (uint16_t)re, (uint16_t)mm);
    size_t const orig = paths.size();
    std::vector<uint8_t> matched(pp.size(), 0);
    for (size_t i = 0; i < orig; ++i)
    {
      if (paths[i].re != rs)
        continue;
      bool once = false;
      Path const op = paths[i];
      for (size_t j = 0; j < pp.size(); ++j)
      {
        if (op.end == pp[j].start && op.re == pp[j].rs)
        {
          Path np = merge_paths(op, pp[j]);
          if (np.start != op.start || np.rs != op.rs)
            continue;
          matched[j] = 1;
          if (once)
            paths.push_back(std::move(np));
          else
          {
            longest = std::max(np.size(), longest);
            paths[i] = std::move(np);
            once = true;
          }
        }
      }
    }
    for (size_t j = 0; j < pp.size(); ++j)
      if (!matched[j])
      {
        longest = std::max(pp[j].size(), longest);
        paths.push_back(pp[j]);
      }
  }

  // genotype_paths.cpp:233-292
  void add_prev(const G & g, const std::vector<Label> & ll, uint32_t rs, uint32_t re, int mm)
  {
    std::vector<Path> const pp = find_all_nonduplicated_paths(g, ll, (uint16_t)rs, (uint16_t)re, (uint16_t)mm);
    size_t const orig = paths.size();
    std::vector<uint8_t> matched(pp.size(), 0);
    for (size_t i = 0; i < orig; ++i)
    {
      if (paths[i].rs != re)
        continue;
      bool once = false;
      Path const op = paths[i];
      for (size_t j = 0; j < pp.size(); ++j)
      {
        if (pp[j].end == op.start && pp[j].re == op.rs)
        {
          Path np = merge_paths(pp[j], op);
          if (np.rs != pp[j].rs)
            continue;
          matched[j] = 1;
          if (once)
            paths.push_back(std::move(np));
          else
          {
            longest = std::max(np.size(), longest);
            paths[i] = std::move(np);
            once = true;
          }
        }
      }
    }
    for (size_t j = 0; j < pp.size(); ++j)
      if (!matched[j])
      {
        longest = std::max(pp[j].size(), longest);
        paths.push_back(pp[j]);
      }
  }

  void remove_short_paths() // genotype_paths.cpp:824-834
  {
    if (longest <= 1)
      return;
    uint32_t const L = longest;
    paths.erase(std::remove_if(paths.begin(), paths.end(), [L](Path const & p) { return p.size() < L; }), paths.end());
  }
  void update_longest() // genotype_paths.cpp:858-864
  {
    longest = 0;
    for (auto const & p : paths)
      longest = std::max(p.size(), longest);
  }
  void remove_paths_with_too_many_mismatches() // genotype_paths.cpp:360-380
  {
    if (paths.empty())
      return;
    uint16_t mn = 10;
    for (auto const & p : paths)
      mn = std::min(p.mm, mn);
    paths.erase(std::remove_if(paths.begin(), paths.end(), [mn](Path const & p) { return p.mm > mn; }), paths.end());
  }
  bool all_paths_unique(const G & g) const // genotype_paths.cpp:219-231
  {
    for (size_t i = 1; i < paths.size(); ++i)
      if (g.get_ref_reach_pos(paths[0].start) != g.get_ref_reach_pos(paths[i].start) &&
          g.get_ref_reach_pos(paths[0].end) != g.get_ref_reach_pos(paths[i].end))
        return false;
    return true;
  }
  void remove_non_ref_paths_when_read_matches_ref(const G & g) // genotype_paths.cpp:460-474
  {
    if (all_paths_unique(g))
      return;
    bool any = false;
    for (auto const & p : paths)
      if (p.is_reference())
      {
        any = true;
        break;
      }
    if (any)
      paths.erase(std::remove_if(paths.begin(), paths.end(), [](Path const & p) { return !p.is_reference(); }),
                  paths.end());
  }
  void remove_fully_special_paths(const G & g) // genotype_paths.cpp:476-481
  {
    paths.erase(std::remove_if(paths.begin(), paths.end(),
                               [&g](Path const & p)
                               { return g.get_ref_reach_pos(p.start) == g.get_ref_reach_pos(p.end); }),
                paths.end());
  }
  void remove_support_from_read_ends(const G & g) // genotype_paths.cpp:382-430
  {
    long constexpr MIN_OFFSET = 4;
    for (Path & path : paths)
    {
      if (path.var_order.empty())
        continue;
      if (!g.is_special_pos(path.start) && !g.is_special_pos(path.end))
        continue;
      auto mm = std::minmax_element(path.var_order.begin(), path.var_order.end());
      if (g.is_special_pos(path.end) && g.get_actual_pos(path.end) <= (*mm.second) + MIN_OFFSET)
        path.nums[mm.second - path.var_order.begin()].clear();
      if (g.is_special_pos(path.start))
      {
        bool amb;
        if (g.is_special_pos(path.start + (uint32_t)MIN_OFFSET))
          amb = (long)g.get_ref_reach_pos(path.start) != (long)g.get_ref_reach_pos(path.start + (uint32_t)MIN_OFFSET);
        else
          amb = true;
        if (amb)
          path.nums[mm.first - path.var_order.begin()].clear();
      }
    }
  }
  bool all_paths_fully_aligned() const // genotype_paths.cpp:836-845
  {
    for (auto const & p : paths)
      if (p.size() != read_length)
        return false;
    return true;
  }
  bool is_proper_pair() const { return ml_insert_size != INSERT_SIZE_WHEN_NOT_PROPER_PAIR; }
};

// ------------------------------------------------------------------------------------------------ graph walk (A7)
struct Location // include/graphtyper/graph/location.hpp
{
  char type = 'U';
  uint32_t node_index = 0, node_order = 0, offset = 0;
};

// graph.cpp:931-1029
std::vector<Location> get_locations_of_an_actual_position(const G & g, uint32_t pos, const Path & path, bool is_special)
{
  std::vector<Location> locs;
  uint32_t const NR = g.v.n_ref;
  if (pos < g.v.ref_order[0])
    return locs;
  if (NR == 1)
  {
    locs.push_back({'R', 0, g.v.ref_order[0], pos - g.v.ref_order[0]});
    return locs;
  }
  for (uint32_t r = 1; r <= NR; ++r)
  {
    if (r < NR && g.v.ref_order[r] <= pos)
      continue;
    int rr = (int)r - 1;
    if (pos < g.v.ref_order[rr] + g.ref_len(rr))
    {
      if (!is_special)
      {
        locs.push_back({'R', (uint32_t)rr, g.v.ref_order[rr], pos - g.v.ref_order[rr]});
        break;
      }
      --rr;
    }
    long const PADDING = g.v.is_sv_graph ? 1000000 : 1000;
    while (rr >= 0 && (long)g.ref_reach(rr) + PADDING > (long)pos)
    {
      for (uint32_t i = 0; i < g.out_degree(rr); ++i)
      {
        uint32_t const v = g.v.ref_var_off[rr] + i;
        if (pos >= g.v.var_order[v] && pos <= g.var_reach(v))
        {
          auto it = std::find(path.var_order.begin(), path.var_order.end(), g.v.var_order[v]);
          if (it == path.var_order.end())
            continue;
          long const j = it - path.var_order.begin();
          if (path.is_empty() || (j < (long)path.nums.size() && path.nums[j].count((uint16_t)i)))
            locs.push_back({'V', v, g.v.var_order[v], pos - g.v.var_order[v]});
        }
      }
      --rr;
    }
    break;
  }
  return locs;
}

// graph.cpp:1154-1185
std::vector<Location> get_locations_of_a_position(const G & g, uint32_t pos, const Path & path)
{
  bool const sp = g.is_special_pos(pos);
  if (sp)
    pos = g.v.actual_poses[pos - SPECIAL_START];
  return get_locations_of_an_actual_position(g, pos, path, sp);
}

// include/graphtyper/graph/graph_utils.hpp:7-37
uint32_t count_mismatches(const std::vector<char> & read, const std::vector<char> & dna, uint32_t max_mm)
{
  uint32_t mm = 0;
  size_t const n = std::min(read.size(), dna.size());
  for (size_t i = 0; i < n; ++i)
  {
    char const gc = dna[i], rc = read[i];
    if (gc == '>' || gc == '<')
      return max_mm + 1;
    if (gc != rc && rc != 'N' && gc != 'N')
    {
      ++mm;
      if (mm > max_mm)
        return mm;
    }
  }
  return mm;
}
// graph_utils.hpp:39-69
uint32_t count_mismatches_backward(const std::vector<char> & read, const std::vector<char> & dna, uint32_t max_mm)
{
  uint32_t mm = 0;
  size_t const n = std::min(read.size(), dna.size());
  for (size_t i = 0; i < n; ++i)
  {
    char const gc = dna[dna.size() - 1 - i], rc = read[read.size() - 1 - i];
    if (gc == '>' || gc == '<')
      return max_mm + 1;
    if (gc != rc && rc != 'N' && gc != 'N')
    {
      ++mm;
      if (mm > max_mm)
        return mm;
    }
  }
  return mm;
}

inline void append(std::vector<char> & s, const uint8_t * b, size_t n) { s.insert(s.end(), b, b + n); }
inline void prepend(std::vector<char> & s, const uint8_t * b, size_t n) { s.insert(s.begin(), b, b + n); }

// graph.cpp:1187-1439
std::vector<Label> get_labels_forward(const G & g, const Location & s, const std::vector<char> & read, uint32_t & max_mm)
{
  std::vector<Label> labels;
  std::vector<std::vector<char>> seqs(1);
  std::vector<std::vector<uint32_t>> var_ids(1);
  std::vector<uint32_t> end_pos(1, 0u);
  uint32_t vb = 0, ve = 0; // current bubble's var range [vb, ve)
  size_t const RL = read.size();

  if (s.type == 'V')
  {
    uint32_t const v = s.node_index;
    var_ids[0].push_back(v);
    seqs[0].assign(g.var_dna(v) + s.offset, g.var_dna(v) + g.var_len(v));
    if (seqs[0].size() >= RL)
    {
      end_pos[0] = (uint32_t)(g.var_reach(v) - (seqs[0].size() - RL));
      uint32_t const rr = g.bubble_ref_reach(v);
      if (end_pos[0] > rr)
        end_pos[0] = g.get_special_pos(end_pos[0], rr);
    }
    else
    {
      uint32_t const r = g.v.var_out_ref[v];
      vb = g.v.ref_var_off[r];
      ve = g.v.ref_var_off[r + 1];
      append(seqs[0], g.ref_dna(r), g.ref_len(r));
      end_pos[0] = (uint32_t)(g.ref_reach(r) - (seqs[0].size() - RL));
    }
  }
  else
  {
    uint32_t const r = s.node_index;
    vb = g.v.ref_var_off[r];
    ve = g.v.ref_var_off[r + 1];
    seqs[0].assign(g.ref_dna(r) + s.offset, g.ref_dna(r) + g.ref_len(r));
    end_pos[0] = (uint32_t)(g.ref_reach(r) - (seqs[0].size() - RL));
  }

  if (ve > vb && seqs[0].size() < RL)
  {
    uint32_t r = g.v.var_out_ref[vb];
    bool all_long = false;
    size_t const MAXC = 128;
    while (!all_long && seqs.size() < MAXC && ve > vb)
    {
      all_long = true;
      size_t orig = seqs.size();
      for (size_t j = 0; j < orig; ++j)
      {
        if (seqs[j].size() >= RL)
          continue;
        for (uint32_t v = vb; v + 1 < ve; ++v)
        {
          std::vector<char> ns(seqs[j]);
          append(ns, g.var_dna(v), g.var_len(v));
          bool const enough = ns.size() >= RL;
          if (!enough)
            append(ns, g.ref_dna(r), g.ref_len(r));
          if (count_mismatches(read, ns, max_mm) <= max_mm)
          {
            std::vector<uint32_t> nv(var_ids[j]);
            nv.push_back(v);
            var_ids.push_back(std::move(nv));
            if (ns.size() < RL)
              all_long = false;
            if (enough)
            {
              uint32_t ep = (uint32_t)(g.var_reach(v) - (ns.size() - RL));
              uint32_t const rr = g.bubble_ref_reach(v);
              if (ep > rr)
                ep = g.get_special_pos(ep, rr);
              end_pos.push_back(ep);
            }
            else
              end_pos.push_back((uint32_t)(g.ref_reach(r) - (ns.size() - RL)));
            seqs.push_back(std::move(ns));
          }
        }
        uint32_t const v = ve - 1;
        append(seqs[j], g.var_dna(v), g.var_len(v));
        bool const enough = seqs[j].size() >= RL;
        if (!enough)
          append(seqs[j], g.ref_dna(r), g.ref_len(r));
        if (count_mismatches(read, seqs[j], max_mm) <= max_mm)
        {
          var_ids[j].push_back(v);
          if (all_long && seqs[j].size() < RL)
            all_long = false;
          if (enough)
          {
            end_pos[j] = (uint32_t)(g.var_reach(v) - (seqs[j].size() - RL));
            uint32_t const rr = g.bubble_ref_reach(v);
            if (end_pos[j] > rr)
              end_pos[j] = g.get_special_pos(end_pos[j], rr);
          }
          else
            end_pos[j] = (uint32_t)(g.ref_reach(r) - (seqs[j].size() - RL));
        }
        else
        {
          seqs.erase(seqs.begin() + j);
          var_ids.erase(var_ids.begin() + j);
          end_pos.erase(end_pos.begin() + j);
          --orig;
          --j;
        }
      }
      if (!all_long)
      {
        vb = g.v.ref_var_off[r];
        ve = g.v.ref_var_off[r + 1];
        ++r;
      }
      else
        break;
    }
  }

  std::vector<std::vector<uint32_t>> best_ids;
  std::vector<uint32_t> best_end;
  for (size_t j = 0; j < seqs.size(); ++j)
  {
    if (seqs[j].size() < RL)
      continue;
    uint32_t const mm = count_mismatches(read, seqs[j], max_mm);
    if (mm > max_mm)
      continue;
    else if (mm < max_mm)
    {
      max_mm = mm;
      best_ids.clear();
      best_end.clear();
    }
    best_ids.push_back(var_ids[j]);
    best_end.push_back(end_pos[j]);
  }
  for (size_t j = 0; j < best_ids.size(); ++j)
  {
    uint32_t start_pos = s.node_order + s.offset;
    if (s.type == 'V')
    {
      uint32_t const rr = g.bubble_ref_reach(s.node_index);
      if (start_pos > rr)
        start_pos = g.get_special_pos(start_pos, rr);
    }
    if (best_ids[j].empty())
      labels.push_back({start_pos, best_end[j], INVALID_ID});
    else
      for (uint32_t v : best_ids[j])
        labels.push_back({start_pos, best_end[j], v});
  }
  return labels;
}

// graph.cpp:1441-1701
std::vector<Label> get_labels_backward(const G & g, const Location & e, const std::vector<char> & read, uint32_t & max_mm)
{
  std::vector<Label> labels;
  std::vector<std::vector<char>> seqs(1);
  std::vector<std::vector<uint32_t>> var_ids(1);
  std::vector<uint32_t> start_pos(1, 0u);
  uint32_t vb = 0, ve = 0;
  size_t const RL = read.size();

  if (e.type == 'V')
  {
    uint32_t const v = e.node_index;
    var_ids[0].push_back(v);
    seqs[0].assign(g.var_dna(v), g.var_dna(v) + e.offset + 1);
    if (seqs[0].size() >= RL)
    {
      start_pos[0] = (uint32_t)(g.v.var_order[v] + (seqs[0].size() - RL));
      uint32_t const rr = g.bubble_ref_reach(v);
      if (start_pos[0] > rr)
        start_pos[0] = g.get_special_pos(start_pos[0], rr);
    }
    else
    {
      uint32_t const r = g.v.var_out_ref[v] - 1;
      prepend(seqs[0], g.ref_dna(r), g.ref_len(r));
      start_pos[0] = (uint32_t)(g.v.ref_order[r] + (seqs[0].size() - RL));
      if (r != 0)
      {
        vb = g.v.ref_var_off[r - 1];
        ve = g.v.ref_var_off[r];
      }
    }
  }
  else
  {
    uint32_t const r = e.node_index;
    if (r != 0)
    {
      vb = g.v.ref_var_off[r - 1];
      ve = g.v.ref_var_off[r];
    }
    seqs[0].assign(g.ref_dna(r), g.ref_dna(r) + e.offset + 1);
    start_pos[0] = (uint32_t)(g.v.ref_order[r] + (seqs[0].size() - RL));
  }

  if (ve > vb && seqs[0].size() < RL)
  {
    uint32_t r = g.v.var_out_ref[vb] - 1;
    bool all_long = false;
    size_t const MAXC = 128;
    while (!all_long && seqs.size() < MAXC && ve > vb)
    {
      all_long = true;
      size_t orig = seqs.size();
      for (size_t j = 0; j < orig; ++j)
      {
        if (seqs[j].size() >= RL)
          continue;
        for (uint32_t v = vb; v + 1 < ve; ++v)
        {
          std::vector<char> ns(g.var_dna(v), g.var_dna(v) + g.var_len(v));
          ns.insert(ns.end(), seqs[j].begin(), seqs[j].end());
          bool const enough = ns.size() >= RL;
          if (!enough)
            prepend(ns, g.ref_dna(r), g.ref_len(r));
          if (count_mismatches_backward(read, ns, max_mm) <= max_mm)
          {
            std::vector<uint32_t> nv(var_ids[j]);
            nv.push_back(v);
            var_ids.push_back(std::move(nv));
            if (ns.size() < RL)
              all_long = false;
            if (enough)
            {
              uint32_t sp = (uint32_t)(g.v.var_order[v] + (ns.size() - RL));
              uint32_t const rr = g.bubble_ref_reach(v);
              if (sp > rr)
                sp = g.get_special_pos(sp, rr);
              start_pos.push_back(sp);
            }
            else
              start_pos.push_back((uint32_t)(g.v.ref_order[r] + (ns.size() - RL)));
            seqs.push_back(std::move(ns));
          }
        }
        uint32_t const v = ve - 1;
        prepend(seqs[j], g.var_dna(v), g.var_len(v));
        bool const enough = seqs[j].size() >= RL;
        if (!enough)
          prepend(seqs[j], g.ref_dna(r), g.ref_len(r));
        if (count_mismatches_backward(read, seqs[j], max_mm) <= max_mm)
        {
          var_ids[j].push_back(v);
          if (seqs[j].size() < RL)
            all_long = false;
          if (enough)
          {
            start_pos[j] = (uint32_t)(g.v.var_order[v] + (seqs[j].size() - RL));
            uint32_t const rr = g.bubble_ref_reach(v);
            if (start_pos[j] > rr)
              start_pos[j] = g.get_special_pos(start_pos[j], rr);
          }
          else
            start_pos[j] = (uint32_t)(g.v.ref_order[r] + (seqs[j].size() - RL));
        }
        else
        {
          seqs.erase(seqs.begin() + j);
          var_ids.erase(var_ids.begin() + j);
          start_pos.erase(start_pos.begin() + j);
          --orig;
          --j;
        }
      }
      if (!all_long)
      {
        if (r != 0)
        {
          --r;
          vb = g.v.ref_var_off[r];
          ve = g.v.ref_var_off[r + 1];
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

  std::vector<std::vector<uint32_t>> best_ids;
  std::vector<uint32_t> best_start;
  for (size_t j = 0; j < seqs.size(); ++j)
  {
    if (seqs[j].size() < RL)
      continue;
    uint32_t const mm = count_mismatches_backward(read, seqs[j], max_mm);
    if (mm < max_mm)
    {
      max_mm = mm;
      best_ids.clear();
      best_start.clear();
      best_ids.push_back(var_ids[j]);
      best_start.push_back(start_pos[j]);
    }
    else if (mm == max_mm)
    {
      best_ids.push_back(var_ids[j]);
      best_start.push_back(start_pos[j]);
    }
  }
  for (size_t j = 0; j < best_ids.size(); ++j)
  {
    uint32_t end_pos = e.node_order + e.offset;
    if (e.type == 'V')
    {
      uint32_t const rr = g.bubble_ref_reach(e.node_index);
      if (end_pos > rr)
        end_pos = g.get_special_pos(end_pos, rr);
    }
    if (best_ids[j].empty())
      labels.push_back({best_start[j], end_pos, INVALID_ID});
    else
      for (uint32_t v : best_ids[j])
        labels.push_back({best_start[j], end_pos, v});
  }
  return labels;
}

// graph.cpp:1703-1754
std::vector<Label> iterative_dfs(const G & g, const std::vector<Location> & s_locs, const std::vector<Location> & e_locs,
                                 const std::vector<char> & sub, uint32_t & max_mm)
{
  std::vector<Label> labels;
  if (s_locs.size() > 1024 || e_locs.size() > 1024)
    return labels;
  auto add_if_better = [&](std::vector<Label> && nl, uint32_t mm)
  {
    if (!nl.empty())
    {
      if (mm < max_mm)
      {
        max_mm = mm;
        labels = std::move(nl);
      }
      else if (mm == max_mm)
        labels.insert(labels.end(), nl.begin(), nl.end());
    }
  };
  if (s_locs.size() == 1 && s_locs[0].type == 'U')
  {
    for (auto const & e : e_locs)
    {
      uint32_t mm = max_mm;
      auto nl = get_labels_backward(g, e, sub, mm);
      add_if_better(std::move(nl), mm);
    }
  }
  else
  {
    for (auto const & s : s_locs)
    {
      uint32_t mm = max_mm;
      auto nl = get_labels_forward(g, s, sub, mm);
      add_if_better(std::move(nl), mm);
    }
  }
  return labels;
}

const char IUPAC_CHAR[17] = "UACMGRSVTWYHKDBN"; // seqan alphabet_residue_tabs.h:222-240

// genotype_paths.cpp:483-553
void walk_read_ends(const G & g, GenoPaths & gp, const std::vector<char> & seq)
{
  auto & paths = gp.paths;
  if (paths.empty() || paths[0].size() == seq.size())
    return;
  if (paths.size() > MAX_SEED_NUMBER_FOR_WALKING)
    return;
  int maximum_mismatches = -1;
  if (paths.size() > MAX_SEED_NUMBER_ALLOWING_MISMATCHES)
    maximum_mismatches = 0;
  size_t best_mm = 7;
  std::vector<uint32_t> best_idx;
  std::vector<std::vector<Label>> best_labels;
  for (auto & path : paths)
  {
    if (path.re == seq.size() - 1)
      continue;
    std::vector<Location> s_locs = get_locations_of_a_position(g, path.end, path);
    if (s_locs.empty() || s_locs.size() > MAX_NUM_LOCATIONS_PER_PATH)
      continue;
    std::vector<char> kmer(seq.begin() + path.re, seq.end());
    std::vector<Location> e_locs(1);
    uint32_t mm = maximum_mismatches < 0 ? (uint32_t)std::min((size_t)(2 + kmer.size() / 11), best_mm)
                                         : (uint32_t)maximum_mismatches;
    std::vector<Label> nl = iterative_dfs(g, s_locs, e_locs, kmer, mm);
    if (!nl.empty())
    {
      if (mm < best_mm)
      {
        best_labels.clear();
        best_idx.clear();
        best_labels.push_back(std::move(nl));
        best_idx.push_back(path.re);
        best_mm = mm;
      }
      else if (mm == best_mm)
      {
        best_labels.push_back(std::move(nl));
        best_idx.push_back(path.re);
      }
    }
  }
  for (size_t i = 0; i < best_labels.size(); ++i)
    gp.add_next(g, best_labels[i], best_idx[i], (uint32_t)seq.size() - 1, (int)best_mm);
}

// genotype_paths.cpp:555-621
void walk_read_starts(const G & g, GenoPaths & gp, const std::vector<char> & seq)
{
  auto & paths = gp.paths;
  if (paths.empty() || paths[0].size() == seq.size())
    return;
  if (paths.size() > MAX_SEED_NUMBER_FOR_WALKING)
    return;
  int maximum_mismatches = -1;
  if (paths.size() > MAX_SEED_NUMBER_ALLOWING_MISMATCHES)
    maximum_mismatches = 0;
  size_t best_mm = 7;
  std::vector<uint32_t> best_idx;
  std::vector<std::vector<Label>> best_labels;
  for (auto & path : paths)
  {
    if (path.rs == 0)
      continue;
    std::vector<char> kmer(seq.begin(), seq.begin() + path.rs + 1);
    std::vector<Location> e_locs = get_locations_of_a_position(g, path.start, path);
    if (e_locs.empty() || e_locs.size() > MAX_NUM_LOCATIONS_PER_PATH)
      continue;
    std::vector<Location> s_locs(1);
    uint32_t mm = maximum_mismatches < 0 ? (uint32_t)std::min((size_t)(2 + kmer.size() / 11), best_mm)
                                         : (uint32_t)maximum_mismatches;
    std::vector<Label> nl = iterative_dfs(g, s_locs, e_locs, kmer, mm);
    if (!nl.empty())
    {
      if (mm < best_mm)
      {
        best_labels.clear();
        best_idx.clear();
        best_labels.push_back(std::move(nl));
        best_idx.push_back(path.rs);
        best_mm = mm;
      }
      else if (mm == best_mm)
      {
        best_labels.push_back(std::move(nl));
        best_idx.push_back(path.rs);
      }
    }
  }
  for (size_t i = 0; i < best_labels.size(); ++i)
    gp.add_prev(g, best_labels[i], 0, best_idx[i], (int)best_mm);
}

struct SeedTap
{
  std::vector<std::vector<Label>> h0, h1;
  bool computed = false;
};

// alignment.cpp:23-103
void find_genotype_paths_of_one_of_the_sequences(const G & g, const Index & idx, const uint8_t * codes, size_t len,
                                                 GenoPaths & geno, SeedTap * tap)
{
  auto h0 = query_index(idx, codes, len);
  auto h1 = query_index_ham1(idx, codes, len);
  if (tap)
  {
    tap->h0 = h0;
    tap->h1 = h1;
    tap->computed = true;
  }
  {
    size_t i = 0;
    for (;;)
    {
      if (h0[i].size() < MAX_UNIQUE_KMER_POSITIONS)
        break;
      ++i;
      if (i == h0.size())
        return;
    }
  }
  uint32_t rs = 0;
  for (size_t i = 0; i < h0.size(); ++i)
  {
    geno.add_next(g, h0[i], rs, rs + (K - 1), 0);
    geno.add_next(g, h1[i], rs, rs + (K - 1), 1);
    rs += (K - 1);
  }
  geno.remove_short_paths();
  std::vector<char> seq(len);
  for (size_t i = 0; i < len; ++i)
    seq[i] = IUPAC_CHAR[codes[i]];
  walk_read_starts(g, geno, seq);
  walk_read_ends(g, geno, seq);
  geno.update_longest();
  geno.remove_short_paths();
  geno.remove_paths_with_too_many_mismatches();
  if (g.v.is_sv_graph)
    geno.remove_fully_special_paths(g);
  geno.remove_non_ref_paths_when_read_matches_ref(g);
  geno.update_longest();
  geno.remove_short_paths();
  if (g.v.is_sv_graph)
    geno.remove_support_from_read_ends(g);
}

inline uint8_t comp4(uint8_t c) { return (uint8_t)(((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3)); }

struct Alignment
{
  GenoPaths first, second;
  SeedTap tap[2];
};

// alignment.cpp:331-363 (force_align_both_orientations = false, options.hpp:89)
void align_read(const G & g, const Index & idx, const gtb_read_batch & b, uint32_t i, Alignment & out, bool tap)
{
  uint16_t const flag = b.flag[i];
  uint16_t const len = b.lseq[i];
  out.first = GenoPaths();
  out.second = GenoPaths();
  out.first.flags = out.second.flags = flag;
  out.first.read_length = out.second.read_length = len;
  if (len < 2 * K - 1)
    return;
  std::vector<uint8_t> codes(len), rcodes(len);
  const uint8_t * s4 = b.seq4 + (size_t)i * b.seq_stride;
  for (uint32_t j = 0; j < len; ++j)
    codes[j] = (j & 1) ? (s4[j >> 1] & 15) : (s4[j >> 1] >> 4);
  for (uint32_t j = 0; j < len; ++j)
    rcodes[j] = comp4(codes[len - 1 - j]);
  int32_t const isize = b.isize[i];
  bool const fwd_only = (flag & IS_PAIRED) == 0u ||
                        (b.same_tid[i] && isize > -1200 && isize < 1200 &&
                         (((flag & IS_SEQ_REVERSED) != 0u) != ((flag & IS_MATE_SEQ_REVERSED) != 0u)));
  find_genotype_paths_of_one_of_the_sequences(g, idx, codes.data(), len, out.first, tap ? &out.tap[0] : nullptr);
  if (!fwd_only)
    find_genotype_paths_of_one_of_the_sequences(g, idx, rcodes.data(), len, out.second, tap ? &out.tap[1] : nullptr);
}

// alignment.cpp:482-538
void update_paths(GenoPaths & g1, GenoPaths & g2, const gtb_read_batch & b, uint32_t i)
{
  uint16_t const flag = b.flag[i];
  g1.flags = flag & ~IS_PROPER_PAIR;
  g1.mapq = b.mapq[i];
  g1.ml_insert_size = std::abs(b.isize[i]);
  if (b.mapq[i] < 25)
    g1.flags |= IS_MAPQ_BAD;
  if (b.clipped && b.clipped[i])
  {
    g1.flags |= IS_CLIPPED;
    g2.flags |= IS_CLIPPED;
  }
  g1.score_diff = g2.score_diff = b.score_diff[i];
  g2.flags = (flag ^ IS_SEQ_REVERSED) & ~IS_PROPER_PAIR;
  g2.mapq = g1.mapq;
  g2.ml_insert_size = g1.ml_insert_size;
}

// genotype_paths.cpp:943-974
int compare_single(const GenoPaths & a, const GenoPaths & b)
{
  size_t const T1 = a.longest, T2 = b.longest, MIN = 94;
  if (T1 > T2 && T1 > MIN)
    return 1;
  else if (T2 > T1 && T2 > MIN)
    return 2;
  else if (T2 == T1 && T1 > MIN)
  {
    size_t const m1 = a.paths[0].mm, m2 = b.paths[0].mm;
    if (m1 < m2)
      return 1;
    else if (m2 < m1)
      return 2;
    return 1;
  }
  return 0;
}

size_t alternative_call_count(const std::vector<Path> & paths)
{
  size_t c = 0;
  for (auto const & p : paths)
    for (auto const & n : p.nums)
      c += (n.count(0) == 0);
  return c;
}

// genotype_paths.cpp:976-1169
int compare_pairs(const GenoPaths & a1, const GenoPaths & a2, const GenoPaths & b1, const GenoPaths & b2)
{
  size_t const T11 = a1.paths.size() > 0 ? a1.longest : 0;
  size_t const T12 = a2.paths.size() > 0 ? a2.longest : 0;
  size_t const T21 = b1.paths.size() > 0 ? b1.longest : 0;
  size_t const T22 = b2.paths.size() > 0 ? b2.longest : 0;
  size_t const MAX1 = std::max(T11, T12), MAX2 = std::max(T21, T22);
  size_t const P1 = a1.read_length, P2 = a2.read_length, MIN = 94;
  bool const perf1 = T11 >= P1 && T12 >= P2, perf2 = T21 >= P1 && T22 >= P2;
  if (perf1 || perf2)
  {
    if (perf1 && perf2)
    {
      size_t const m1 = a1.paths[0].mm + a2.paths[0].mm, m2 = b1.paths[0].mm + b2.paths[0].mm;
      if (m1 < m2)
        return 1;
      else if (m2 < m1)
        return 2;
      size_t const n1 = a1.paths.size() + a2.paths.size(), n2 = b1.paths.size() + b2.paths.size();
      if (n1 < n2)
        return 1;
      else if (n2 < n1)
        return 2;
      size_t const c1 = alternative_call_count(a1.paths) + alternative_call_count(a2.paths);
      size_t const c2 = alternative_call_count(b1.paths) + alternative_call_count(b2.paths);
      return c1 >= c2 ? 1 : 2;
    }
    else if (perf1)
      return 1;
    else
      return 2;
  }
  else if (MAX2 >= MIN && MAX2 > MAX1)
    return 2;
  else if (MAX1 >= MIN && MAX1 > MAX2)
    return 1;
  else if (MAX1 >= MIN && MAX2 >= MIN)
  {
    uint16_t m1 = 10, m2 = 10;
    if (T11 == MAX1)
      m1 = std::min(m1, a1.paths[0].mm);
    if (T12 == MAX1)
      m1 = std::min(m1, a2.paths[0].mm);
    if (T21 == MAX2)
      m2 = std::min(m2, b1.paths[0].mm);
    if (T22 == MAX2)
      m2 = std::min(m2, b2.paths[0].mm);
    if (m1 < m2)
      return 1;
    else if (m2 < m1)
      return 2;
    if (std::min(T11, T12) < std::min(T21, T22))
      return 1;
    else if (std::min(T21, T22) < std::min(T11, T12))
      return 2;
    return 0;
  }
  else if (MAX2 == 0u && T11 >= 63u && T12 >= 63u)
    return 1;
  else if (MAX1 == 0u && T21 >= 63u && T22 >= 63u)
    return 2;
  return 1;
}

// ------------------------------------------------------------------------------------------------ accumulate (A11, A12)
constexpr uint16_t NO_COVERAGE = 0xFFFFu, MULTI_ALT_COVERAGE = 0xFFFEu, MULTI_REF_COVERAGE = 0xFFFDu;

struct HapSample // include/graphtyper/graph/haplotype.hpp:25-75
{
  std::vector<uint16_t> log_score, gt_coverage;
  uint16_t max_log_score = 0;
  uint8_t ambiguous_depth = 0, ambiguous_depth_alt = 0, alt_proper_pair_depth = 0;
  uint32_t saturated = 0;
};

struct Hap
{
  uint32_t id = 0, num = 0;
  std::vector<HapSample> samples;
  uint16_t coverage = NO_COVERAGE;
  std::set<uint64_t> explains;
  uint64_t clipped_reads = 0, mapq_squared = 0;
  std::vector<uint64_t> pa_clipped_bp, pa_mapq_squared, pa_score_diff, pa_mismatches;
  std::vector<uint32_t> read_strand; // 4 per allele

  void add_coverage(uint16_t c) // haplotype.cpp:180-227
  {
    switch (coverage)
    {
    case NO_COVERAGE:
      coverage = c;
      break;
    case MULTI_ALT_COVERAGE:
      if (c == 0)
        coverage = MULTI_REF_COVERAGE;
      break;
    case MULTI_REF_COVERAGE:
      break;
    default:
      if (coverage != c)
        coverage = (coverage == 0 || c == 0) ? MULTI_REF_COVERAGE : MULTI_ALT_COVERAGE;
      break;
    }
  }
};

// push_to_haplotype_scores' return value (vcf_writer.cpp:502-519): (hap, allele) -> list of (later hap, allele)
using HapAllele = std::pair<uint16_t, uint16_t>;
using ConnMap = std::map<HapAllele, std::vector<HapAllele>>;

struct Writer
{
  std::vector<Hap> haps;
  std::vector<uint32_t> hap_orders; // ascending bubble orders for id2hap (vcf_writer.cpp:84)
  // HapSample::connections (haplotype.hpp:42) of all samples as one sparse map:
  // (sample, hap1, allele1, hap2, allele2) -> uint16 support count
  bool conn_on = false;
  std::map<std::array<uint32_t, 5>, uint16_t> conn;

  // "add new connections to hap_samples" (vcf_writer.cpp:119-140, 229-249)
  void add_connections(ConnMap const & merged, long pn)
  {
    for (auto const & kv : merged)
      for (auto const & to : kv.second)
        ++conn[{(uint32_t)pn, kv.first.first, kv.first.second, to.first, to.second}];
  }

  void init(const G & g, int n_samples) // graph.cpp:680-702, haplotype.cpp:122-147
  {
    for (uint32_t r = 0; r + 1 < g.v.n_ref; ++r)
    {
      Hap h;
      h.id = g.v.var_order[g.v.ref_var_off[r]];
      h.num = g.out_degree(r);
      h.samples.resize(n_samples);
      for (auto & s : h.samples)
      {
        s.log_score.assign((size_t)h.num * (h.num + 1) / 2, 0);
        s.gt_coverage.assign(h.num, 0);
      }
      h.pa_clipped_bp.assign(h.num, 0);
      h.pa_mapq_squared.assign(h.num, 0);
      h.pa_score_diff.assign(h.num, 0);
      h.pa_mismatches.assign(h.num, 0);
      h.read_strand.assign((size_t)h.num * 4, 0);
      haps.push_back(std::move(h));
    }
    for (auto const & h : haps)
      hap_orders.push_back(h.id);
  }

  uint32_t id2hap(uint32_t order) const
  {
    // unordered_map<id, index>: on duplicate ids the last insertion wins (vcf_writer.cpp:84)
    auto it = std::upper_bound(hap_orders.begin(), hap_orders.end(), order);
    return (uint32_t)(it - hap_orders.begin()) - 1;
  }
};

// vcf_writer.cpp:28-60 (hq_reads = false)
bool are_genotype_paths_good(const G & g, const GenoPaths & geno)
{
  if (geno.paths.empty())
    return false;
  bool const fully = geno.all_paths_fully_aligned();
  if (!fully && (!geno.all_paths_unique(g) || geno.paths[0].size() < 63))
    return false;
  double const ratio = (double)geno.paths[0].mm / (double)geno.paths[0].size();
  if (ratio > 0.05)
    return false;
  if (!fully && ratio > 0.025)
    return false;
  if (g.v.is_sv_graph)
    if (!fully || geno.paths[0].size() < 90 || ratio > 0.03)
      return false;
  return true;
}

// vcf_writer.cpp:503-676; the returned connections are only built when Writer::conn_on (they are consumed when
// is_writing_hap, hts_parallel_reader.cpp:782)
ConnMap push_to_haplotype_scores(const G & g, Writer & w, const GenoPaths & geno, long pn)
{
  ConnMap new_connections;
  int const clipped_bp = (int)geno.read_length - (int)geno.longest;
  bool const fully_aligned = clipped_bp == 0;
  bool const non_unique = !geno.all_paths_unique(g);
  size_t const mismatches = geno.paths[0].mm;
  std::map<uint32_t, bool> recent;

  for (auto const & p : geno.paths)
  {
    for (size_t i = 0; i < p.var_order.size(); ++i)
    {
      uint32_t const hap_id = w.id2hap(p.var_order[i]);
      if (p.nums[i].empty())
        continue;
      Hap & hap = w.haps[hap_id];
      auto const & num = p.nums[i];
      long const MIN_OFFSET = 3;
      bool const overlapping = ((long)g.get_ref_reach_pos(p.start) + MIN_OFFSET) <= (long)p.var_order[i] &&
                               ((long)g.get_ref_reach_pos(p.end) - MIN_OFFSET) > (long)p.var_order[i];
      recent[hap_id] |= overlapping;
      hap.explains.insert(num.begin(), num.end());
      if (num.size() == 1)
        hap.add_coverage(*num.begin());
      else
      {
        hap.add_coverage(1);
        if (num.count(0) == 1)
          hap.add_coverage(0);
        else
          hap.add_coverage(2);
      }
    }
  }

  // "check connections" (vcf_writer.cpp:587-637)
  if (w.conn_on)
    for (auto it = recent.begin(); it != recent.end(); ++it)
    {
      Hap const & h1 = w.haps[it->first];
      long const n1 = (long)h1.explains.size();
      if (n1 == 0 || n1 > 64)
        continue;
      for (uint64_t b1 : h1.explains) // ascending alleles, as the b1 counter loop visits them
      {
        auto & conn = new_connections[{(uint16_t)it->first, (uint16_t)b1}];
        for (auto it2 = std::next(it); it2 != recent.end(); ++it2)
        {
          Hap const & h2 = w.haps[it2->first];
          long const n2 = (long)h2.explains.size();
          if (n2 == 0 || n2 > 64)
            continue;
          long const weight = n1 * n2;
          long const repeat = weight >= 3 ? 6 / weight : 1;
          for (uint64_t b2 : h2.explains)
            for (long r = 0; r < repeat; ++r)
              conn.push_back({(uint16_t)it2->first, (uint16_t)b2});
        }
      }
    }

  for (auto const & kv : recent)
  {
    Hap & h = w.haps[kv.first];
    HapSample & hs = h.samples[pn];
    uint16_t const cov = h.coverage;
    // haplotype.cpp:229-313
    if (clipped_bp != 0)
    {
      long const scaled = (clipped_bp * 1000l) / geno.read_length;
      if (cov != NO_COVERAGE)
        ++h.clipped_reads;
      if (cov < MULTI_REF_COVERAGE)
        h.pa_clipped_bp[cov] += scaled;
    }
    if (geno.mapq != 255)
    {
      uint64_t const sq = (uint64_t)geno.mapq * geno.mapq;
      if (cov != NO_COVERAGE)
        h.mapq_squared += sq;
      if (cov < MULTI_REF_COVERAGE)
        h.pa_mapq_squared[cov] += sq;
    }
    if (cov < MULTI_REF_COVERAGE)
    {
      bool const fwd = (geno.flags & IS_SEQ_REVERSED) == 0, first = (geno.flags & IS_FIRST_IN_PAIR) != 0;
      ++h.read_strand[(size_t)cov * 4 + (first ? 0 : 2) + (fwd ? 0 : 1)];
    }
    {
      uint8_t const mm8 = (uint8_t)mismatches;
      if (mm8 != 0 && cov < MULTI_REF_COVERAGE)
        h.pa_mismatches[cov] += (mm8 * 1000l) / geno.read_length;
    }
    if (geno.score_diff != 0 && cov < MULTI_REF_COVERAGE)
      h.pa_score_diff[cov] += geno.score_diff;

    // haplotype.cpp:462-585
    long eps = 12;
    eps -= (long)mismatches;
    if (non_unique)
      eps -= 3;
    if (geno.flags & IS_MAPQ_BAD)
      eps -= 2;
    if (!fully_aligned)
      eps -= 3;
    if (!kv.second)
      eps -= 1;
    uint16_t const e = (uint16_t)(std::max(eps, 8l) - 4);
    if (hs.max_log_score < (0xFFFFul - e))
    {
      hs.max_log_score += e;
      size_t i = 0;
      for (size_t y = 0; y < h.num; ++y)
      {
        bool const ey = h.explains.count(y) == 1;
        for (size_t x = 0; x <= y; ++x, ++i)
        {
          bool const ex = h.explains.count(x) == 1;
          if (ex && ey)
            hs.log_score[i] += e;
          else if (ex || ey)
            hs.log_score[i] += e - 1;
        }
      }
    }
    else
      hs.saturated = 1;

    // haplotype.cpp:315-361
    bool const pp = geno.is_proper_pair();
    auto inc8 = [](uint8_t & x)
    {
      if (x < 0xFF)
        ++x;
    };
    switch (cov)
    {
    case NO_COVERAGE:
      break;
    case MULTI_REF_COVERAGE:
      inc8(hs.ambiguous_depth);
      break;
    case MULTI_ALT_COVERAGE:
      inc8(hs.ambiguous_depth);
      inc8(hs.ambiguous_depth_alt);
      if (pp)
        inc8(hs.alt_proper_pair_depth);
      break;
    default:
      if (hs.gt_coverage[cov] < 0xFFFF)
        ++hs.gt_coverage[cov];
      if (cov > 0 && pp)
        inc8(hs.alt_proper_pair_depth);
      break;
    }
    h.coverage = NO_COVERAGE;
    h.explains.clear();
  }
  return new_connections;
}

// VcfWriter::update_haplotype_scores_geno, single read (vcf_writer.cpp:88-141); the caller has checked are_genotype_paths_good
void update_scores_single(const G & g, Writer & w, const GenoPaths & geno, long pn)
{
  ConnMap const con1 = push_to_haplotype_scores(g, w, geno, pn);
  if (w.conn_on)
    w.add_connections(con1, pn);
}

// VcfWriter::update_haplotype_scores_geno, read pair (vcf_writer.cpp:143-250)
void update_scores_pair(const G & g, Writer & w, const GenoPaths & geno1, const GenoPaths & geno2, long pn)
{
  ConnMap con1, con2;
  if (are_genotype_paths_good(g, geno1))
    con1 = push_to_haplotype_scores(g, w, geno1, pn);
  if (are_genotype_paths_good(g, geno2))
    con2 = push_to_haplotype_scores(g, w, geno2, pn);
  if (!w.conn_on || (con1.empty() && con2.empty()))
    return;
  ConnMap merged;
  for (auto const & kv1 : con1) // :191-208
  {
    auto & dst = merged[kv1.first];
    dst = kv1.second;
    for (auto const & kv2 : con2)
      if (kv2.first.first > kv1.first.first)
        dst.push_back(kv2.first);
  }
  for (auto const & kv2 : con2) // :211-226
  {
    auto ins = merged.insert(kv2);
    if (!ins.second)
      ins.first->second.insert(ins.first->second.end(), kv2.second.begin(), kv2.second.end());
    for (auto const & kv1 : con1)
      if (kv1.first.first > kv2.first.first)
        ins.first->second.push_back(kv1.first);
  }
  w.add_connections(merged, pn);
}

// ------------------------------------------------------------------------------------------------ reference depth (A16)
// src/graph/reference_depth.cpp:17-36,114-232
struct RefDepth
{
  long offset = 0;
  std::vector<std::vector<uint16_t>> depths;
  long start_idx(long p) const { return p < offset ? 0 : p - offset; }
  long end_idx(long e, long size) const { return e > offset + size ? size : e + 1 - offset; }

  void add_genotype_paths(const G & g, const GenoPaths & geno, long sample)
  {
    if (sample >= (long)depths.size())
      return;
    if (geno.paths.empty() || geno.paths[0].size() < 63)
      return;
    auto & depth = depths[sample];
    long const size = (long)depth.size();
    if (geno.paths.size() == 1)
    {
      auto const & p = geno.paths[0];
      long const sp = (long)g.get_ref_reach_pos(p.start) - p.rs;
      long const ep = (long)g.get_ref_reach_pos(p.end) + ((long)geno.read_length - 1 - p.re);
      long const si = start_idx(sp), ei = end_idx(ep, size);
      if (si < size)
        for (long k = si; k != ei && k != size; ++k)
          ++depth[k]; // not saturating in the reference
    }
    else
    {
      std::set<long> local;
      for (auto const & p : geno.paths)
      {
        long sp = (long)g.get_ref_reach_pos(p.start) - p.rs;
        long ep = (long)g.get_ref_reach_pos(p.end) + ((long)geno.read_length - 1 - p.re);
        if (ep - sp >= 50)
        {
          sp += 4;
          ep -= 4;
        }
        if (ep < offset)
          continue;
        long const si = start_idx(sp), ei = end_idx(ep, size);
        if (si < size)
          for (long k = si; k != ei && k != size; ++k)
            local.insert(k);
      }
      for (long k : local)
        if (depth[k] < 0xFFFF)
          ++depth[k];
    }
  }
};

// ------------------------------------------------------------------------------------------------ pool driver (A10)
struct PoolResult
{
  Writer w;
  RefDepth rd;
  int n_samples = 0;
  // debug taps: units = non-duplicate records in batch order
  std::vector<uint32_t> unit_record;
  std::vector<Alignment> units;
  gtb_submit_stats stats{};
};

// hts_parallel_reader.cpp:245-338 (+ alignment.cpp:365-449,557-622)
bool g_connections = false; // gto_set_connections

int run_pool(const G & g, const Index & idx, int n_samples, const gtb_read_batch & b, PoolResult & R, bool tap)
{
  R.w.init(g, n_samples);
  R.w.conn_on = g_connections;
  R.n_samples = n_samples;
  bool const SV = g.v.is_sv_graph != 0;
  if (SV && g.v.n_ref > 0)
  {
    R.rd.offset = g.v.ref_order[0];
    long const size = (long)g.ref_reach(g.v.n_ref - 1) - (long)g.v.ref_order[0] + 1;
    R.rd.depths.assign(n_samples, std::vector<uint16_t>(size, 0));
  }
  uint32_t const n = b.n_reads;
  std::vector<int32_t> unit_of(n, -1);
  // per record: the pair<GenotypePaths, GenotypePaths> after update_paths, kept while waiting for the mate
  std::vector<std::pair<GenoPaths, GenoPaths>> waiting(n);
  std::vector<uint8_t> is_waiting(n, 0);

  for (uint32_t i = 0; i < n; ++i)
  {
    int32_t const d = b.dup_of ? b.dup_of[i] : -1;
    if (d < 0)
    {
      unit_of[i] = (int32_t)R.units.size();
      R.unit_record.push_back(i);
      R.units.emplace_back();
      align_read(g, idx, b, i, R.units.back(), tap);
      ++R.stats.n_alignments;
    }
    else
    {
      if (d >= (int32_t)i || unit_of[d] < 0)
      {
        g_err = "dup_of must reference an earlier record";
        return GTB_ERR_ARG;
      }
      unit_of[i] = unit_of[d];
    }
    std::pair<GenoPaths, GenoPaths> gp(R.units[unit_of[i]].first, R.units[unit_of[i]].second);
    uint16_t const flag = b.flag[i];
    int32_t const m = b.mate ? b.mate[i] : -1;
    int32_t const sample = b.sample[i];

    if (m < 0)
    {
      if (flag & IS_PAIRED)
      {
        update_paths(gp.first, gp.second, b, i);
        waiting[i] = std::move(gp);
        is_waiting[i] = 1;
      }
      else
      {
        // update_unpaired_read_paths, alignment.cpp:365-449
        GenoPaths * sel = nullptr;
        int const c = compare_single(gp.first, gp.second);
        if (c == 1)
        {
          sel = &gp.first;
          sel->flags = flag & ~IS_PROPER_PAIR;
        }
        else if (c == 2)
        {
          sel = &gp.second;
          sel->flags = (flag ^ IS_SEQ_REVERSED) & ~IS_PROPER_PAIR;
        }
        if (sel)
        {
          sel->mapq = b.mapq[i];
          if (b.mapq[i] < 25)
            sel->flags |= IS_MAPQ_BAD;
          if (b.clipped && b.clipped[i])
            sel->flags |= IS_CLIPPED;
          sel->score_diff = b.score_diff[i];
          if (are_genotype_paths_good(g, *sel)) // vcf_writer.cpp:88-141
          {
            update_scores_single(g, R.w, *sel, sample);
            ++R.stats.n_singles_scored;
          }
        }
      }
    }
    else
    {
      if (m >= (int32_t)i || !is_waiting[m])
      {
        g_err = "mate must reference an earlier, still waiting record";
        return GTB_ERR_ARG;
      }
      update_paths(gp.first, gp.second, b, i);
      auto & prev = waiting[m];
      if ((gp.first.flags & IS_FIRST_IN_PAIR) == (prev.first.flags & IS_FIRST_IN_PAIR))
      {
        g_err = "two mates with the same IS_FIRST_IN_PAIR (reference exits, hts_parallel_reader.cpp:306-315)";
        return GTB_ERR_INPUT;
      }
      // get_better_paths, alignment.cpp:557-622
      GenoPaths * arr[4] = {nullptr, nullptr, nullptr, nullptr};
      auto get_index = [](uint16_t f) { return ((f & IS_FIRST_IN_PAIR) != 0) + 2 * ((f & IS_SEQ_REVERSED) == 0); };
      arr[get_index(prev.first.flags)] = &prev.first;
      arr[get_index(prev.second.flags)] = &prev.second;
      arr[get_index(gp.first.flags)] = &gp.first;
      arr[get_index(gp.second.flags)] = &gp.second;
      if (arr[0] && arr[1] && arr[2] && arr[3])
      {
        int const c = compare_pairs(*arr[3], *arr[0], *arr[1], *arr[2]);
        GenoPaths * s1 = nullptr, * s2 = nullptr;
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
        if (s1)
        {
          s1->flags |= IS_PROPER_PAIR;
          s2->flags |= IS_PROPER_PAIR;
          if (SV) // hts_parallel_reader.cpp:321-326
          {
            R.rd.add_genotype_paths(g, *s1, sample);
            R.rd.add_genotype_paths(g, *s2, sample);
          }
          // update_haplotype_scores_geno (pair), vcf_writer.cpp:143-250
          update_scores_pair(g, R.w, *s1, *s2, sample);
          ++R.stats.n_pairs_scored;
        }
      }
      is_waiting[m] = 0;
      waiting[m] = std::pair<GenoPaths, GenoPaths>();
    }
  }
  // leftover mates (SV calling only), hts_parallel_reader.cpp:719-772
  if (SV)
  {
    for (uint32_t i = 0; i < n; ++i)
    {
      if (!is_waiting[i] || !(b.leftover && b.leftover[i]))
        continue;
      auto & orig = waiting[i];
      std::pair<GenoPaths, GenoPaths> copy(orig);
      copy.first.flags ^= (IS_FIRST_IN_PAIR | IS_SEQ_REVERSED);
      copy.second.flags ^= (IS_FIRST_IN_PAIR | IS_SEQ_REVERSED);
      GenoPaths * arr[4] = {nullptr, nullptr, nullptr, nullptr};
      auto get_index = [](uint16_t f) { return ((f & IS_FIRST_IN_PAIR) != 0) + 2 * ((f & IS_SEQ_REVERSED) == 0); };
      arr[get_index(orig.first.flags)] = &orig.first;
      arr[get_index(orig.second.flags)] = &orig.second;
      arr[get_index(copy.first.flags)] = &copy.first;
      arr[get_index(copy.second.flags)] = &copy.second;
      if (!(arr[0] && arr[1] && arr[2] && arr[3]))
        continue;
      int const c = compare_pairs(*arr[3], *arr[0], *arr[1], *arr[2]);
      GenoPaths * s1 = c == 1 ? arr[3] : c == 2 ? arr[1] : nullptr;
      if (!s1)
        continue;
      s1->flags |= IS_PROPER_PAIR;
      R.rd.add_genotype_paths(g, *s1, b.sample[i]);
      if (are_genotype_paths_good(g, *s1))
      {
        update_scores_single(g, R.w, *s1, b.sample[i]);
        ++R.stats.n_singles_scored;
      }
    }
  }
  R.stats.n_records = n;
  return 0;
}

void accumulator_layout(const Writer & w, std::vector<uint64_t> & score_off, std::vector<uint64_t> & cov_off)
{
  score_off.assign(1, 0);
  cov_off.assign(1, 0);
  for (auto const & h : w.haps)
  {
    score_off.push_back(score_off.back() + (uint64_t)h.num * (h.num + 1) / 2);
    cov_off.push_back(cov_off.back() + h.num);
  }
}

} // namespace

// ================================================================================================ C API (ctypes)
extern "C"
{
const char * gto_last_error() { return g_err.c_str(); }

int gto_index_build(const gtb_graph_view * gv, void ** handle)
{
  G g{*gv};
  auto * h = new IndexHandle();
  index_graph(g, h->idx);
  h->sorted_keys.reserve(h->idx.size());
  for (auto const & kv : h->idx)
    h->sorted_keys.push_back(kv.first);
  std::sort(h->sorted_keys.begin(), h->sorted_keys.end());
  *handle = h;
  return 0;
}

int gto_index_size(void * handle, uint64_t * n_keys, uint64_t * n_labels)
{
  auto * h = (IndexHandle *)handle;
  *n_keys = h->sorted_keys.size();
  uint64_t nl = 0;
  for (auto const & kv : h->idx)
    nl += kv.second.size();
  *n_labels = nl;
  return 0;
}

int gto_index_export(void * handle, uint64_t * keys, uint32_t * label_off, gtb_label * labels)
{
  auto * h = (IndexHandle *)handle;
  uint32_t off = 0;
  size_t i = 0;
  for (uint64_t k : h->sorted_keys)
  {
    keys[i] = k;
    label_off[i] = off;
    for (auto const & l : h->idx.at(k))
    {
      labels[off].start = l.start;
      labels[off].end = l.end;
      labels[off].var_id = l.var;
      ++off;
    }
    ++i;
  }
  label_off[i] = off;
  return 0;
}

void gto_index_free(void * handle) { delete (IndexHandle *)handle; }

int gto_pool_run(const gtb_graph_view * gv, void * index, int n_samples, const gtb_read_batch * batch, int tap,
                 void ** result)
{
  G g{*gv};
  auto * R = new PoolResult();
  int rc = run_pool(g, ((IndexHandle *)index)->idx, n_samples, *batch, *R, tap != 0);
  if (rc != 0)
  {
    delete R;
    return rc;
  }
  *result = R;
  return 0;
}

void gto_result_free(void * r) { delete (PoolResult *)r; }

int gto_result_stats(void * r, gtb_submit_stats * s)
{
  *s = ((PoolResult *)r)->stats;
  return 0;
}

int gto_result_accum_sizes(void * r, uint32_t * n_bubbles, uint64_t * n_scores, uint64_t * n_cov)
{
  auto * R = (PoolResult *)r;
  std::vector<uint64_t> so, co;
  accumulator_layout(R->w, so, co);
  *n_bubbles = (uint32_t)R->w.haps.size();
  *n_scores = so.back();
  *n_cov = co.back();
  return 0;
}

int gto_result_accum(void * r, gtb_accumulators * out)
{
  auto * R = (PoolResult *)r;
  std::vector<uint64_t> so, co;
  accumulator_layout(R->w, so, co);
  uint32_t const NB = (uint32_t)R->w.haps.size();
  uint32_t const NS = (uint32_t)R->n_samples;
  out->n_bubbles = NB;
  out->n_samples = NS;
  if (out->ref_depth && !R->rd.depths.empty())
  {
    size_t const sz = R->rd.depths[0].size();
    out->depth_size = (uint32_t)sz;
    out->reference_offset = (uint32_t)R->rd.offset;
    for (uint32_t s = 0; s < NS; ++s)
      memcpy(out->ref_depth + (size_t)s * sz, R->rd.depths[s].data(), sz * 2);
  }
  for (uint32_t b = 0; b <= NB; ++b)
  {
    out->score_off[b] = so[b];
    out->cov_off[b] = co[b];
  }
  for (uint32_t b = 0; b < NB; ++b)
  {
    Hap const & h = R->w.haps[b];
    out->bubble_id[b] = h.id;
    out->n_alleles[b] = h.num;
    uint64_t const tri = so[b + 1] - so[b];
    for (uint32_t s = 0; s < NS; ++s)
    {
      HapSample const & hs = h.samples[s];
      memcpy(out->log_score + so[b] * NS + s * tri, hs.log_score.data(), tri * 2);
      memcpy(out->gt_coverage + co[b] * NS + (uint64_t)s * h.num, hs.gt_coverage.data(), (size_t)h.num * 2);
      out->max_log_score[(uint64_t)b * NS + s] = hs.max_log_score;
      out->ambiguous_depth[(uint64_t)b * NS + s] = hs.ambiguous_depth;
      out->ambiguous_depth_alt[(uint64_t)b * NS + s] = hs.ambiguous_depth_alt;
      out->alt_proper_pair_depth[(uint64_t)b * NS + s] = hs.alt_proper_pair_depth;
      out->saturated[(uint64_t)b * NS + s] = hs.saturated;
    }
    out->vs_clipped_reads[b] = h.clipped_reads;
    out->vs_mapq_squared[b] = h.mapq_squared;
    for (uint32_t a = 0; a < h.num; ++a)
    {
      out->pa_clipped_bp[co[b] + a] = h.pa_clipped_bp[a];
      out->pa_mapq_squared[co[b] + a] = h.pa_mapq_squared[a];
      out->pa_score_diff[co[b] + a] = h.pa_score_diff[a];
      out->pa_mismatches[co[b] + a] = h.pa_mismatches[a];
      for (int k = 0; k < 4; ++k)
        out->read_strand[(co[b] + a) * 4 + k] = h.read_strand[(size_t)a * 4 + k];
    }
  }
  return 0;
}

int gto_result_ref_depth_size(void * r, uint32_t * depth_size, uint32_t * reference_offset)
{
  auto * R = (PoolResult *)r;
  *depth_size = R->rd.depths.empty() ? 0u : (uint32_t)R->rd.depths[0].size();
  *reference_offset = (uint32_t)R->rd.offset;
  return 0;
}

int gto_result_seed_sizes(void * r, uint64_t * n_units, uint64_t * n_slots, uint64_t * n_labels)
{
  auto * R = (PoolResult *)r;
  uint64_t ns = 0, nl = 0;
  for (auto const & u : R->units)
    for (int o = 0; o < 2; ++o)
      if (u.tap[o].computed)
        for (int h = 0; h < 2; ++h)
          for (auto const & s : (h ? u.tap[o].h1 : u.tap[o].h0))
          {
            ++ns;
            nl += s.size();
          }
  *n_units = R->units.size();
  *n_slots = ns;
  *n_labels = nl;
  return 0;
}

// nslots: [n_units][orientation][ham]
int gto_result_seeds(void * r, uint32_t * unit_record, uint32_t * nslots, uint32_t * nlabels, gtb_label * labels)
{
  auto * R = (PoolResult *)r;
  size_t si = 0, li = 0;
  for (size_t u = 0; u < R->units.size(); ++u)
  {
    unit_record[u] = R->unit_record[u];
    for (int o = 0; o < 2; ++o)
      for (int h = 0; h < 2; ++h)
      {
        auto const & T = R->units[u].tap[o];
        auto const & ll = h ? T.h1 : T.h0;
        nslots[(u * 2 + o) * 2 + h] = T.computed ? (uint32_t)ll.size() : 0u;
        if (!T.computed)
          continue;
        for (auto const & s : ll)
        {
          nlabels[si++] = (uint32_t)s.size();
          for (auto const & l : s)
          {
            labels[li].start = l.start;
            labels[li].end = l.end;
            labels[li].var_id = l.var;
            ++li;
          }
        }
      }
  }
  return 0;
}

int gto_result_path_sizes(void * r, uint64_t * n_units, uint64_t * n_paths, uint64_t * n_vars, uint64_t * n_nums)
{
  auto * R = (PoolResult *)r;
  uint64_t np = 0, nv = 0, nn = 0;
  for (auto const & u : R->units)
    for (int o = 0; o < 2; ++o)
      for (auto const & p : (o ? u.second : u.first).paths)
      {
        ++np;
        nv += p.var_order.size();
        for (auto const & s : p.nums)
          nn += s.size();
      }
  *n_units = R->units.size();
  *n_paths = np;
  *n_vars = nv;
  *n_nums = nn;
  return 0;
}

int gto_result_paths(void * r, uint32_t * gp_npaths, uint32_t * gp_longest, uint32_t * p_fields, uint32_t * v_order,
                     uint32_t * v_nnum, uint16_t * v_nums)
{
  auto * R = (PoolResult *)r;
  size_t pi = 0, vi = 0, ni = 0;
  for (size_t u = 0; u < R->units.size(); ++u)
    for (int o = 0; o < 2; ++o)
    {
      GenoPaths const & gp = o ? R->units[u].second : R->units[u].first;
      gp_npaths[u * 2 + o] = (uint32_t)gp.paths.size();
      gp_longest[u * 2 + o] = gp.longest;
      for (auto const & p : gp.paths)
      {
        uint32_t * f = p_fields + pi * 6;
        f[0] = p.start;
        f[1] = p.end;
        f[2] = p.rs;
        f[3] = p.re;
        f[4] = p.mm;
        f[5] = (uint32_t)p.var_order.size();
        ++pi;
        for (size_t k = 0; k < p.var_order.size(); ++k)
        {
          v_order[vi] = p.var_order[k];
          v_nnum[vi] = (uint32_t)p.nums[k].size();
          ++vi;
          for (uint16_t x : p.nums[k])
            v_nums[ni++] = x;
        }
      }
    }
  return 0;
}

// ---- record parsing (N3): what genotype_only's caller derives per record, restated on the CPU
// get_score_diff (src/typer/alignment.cpp:140-325): walk of the aux block; AS / XS of integer type, anything but
// A Z c C s S i I f ends the walk
static uint8_t oracle_score_diff(const uint8_t * it, long l_aux)
{
  long i = 0;
  int64_t as = -1, xs = -1;
  while (i < l_aux)
  {
    i += 3;
    char const type = (char)it[i - 1];
    bool const s_tag = it[i - 2] == 'S';
    bool const is_as = s_tag && it[i - 3] == 'A', is_xs = s_tag && it[i - 3] == 'X';
    int64_t num = 0;
    bool have = false;
    switch (type)
    {
    case 'A':
      ++i;
      break;
    case 'Z':
      while (it[i] != '\0' && it[i] != '\n')
        ++i;
      ++i;
      break;
    case 'c': { int8_t v; memcpy(&v, it + i, 1); num = v; have = true; i += 1; break; }
    case 'C': { uint8_t v; memcpy(&v, it + i, 1); num = v; have = true; i += 1; break; }
    case 's': { int16_t v; memcpy(&v, it + i, 2); num = v; have = true; i += 2; break; }
    case 'S': { uint16_t v; memcpy(&v, it + i, 2); num = v; have = true; i += 2; break; }
    case 'i': { int32_t v; memcpy(&v, it + i, 4); num = v; have = true; i += 4; break; }
    case 'I': { uint32_t v; memcpy(&v, it + i, 4); num = v; have = true; i += 4; break; }
    case 'f':
      i += 4;
      break;
    default:
      i = l_aux;
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
  long const diff = (long)(as - xs);
  return diff < 255 ? (uint8_t)diff : 255;
}

// Fills the gtb_read_batch columns (caller-allocated, n_reads entries; seq4 zero-initialised with stride GTB_SEQ_STRIDE)
int gto_parse_bam(const gtb_bam_batch * b, int is_sv, uint8_t * seq4, uint16_t * lseq, uint16_t * flag, uint8_t * mapq,
                  int32_t * isize, uint8_t * same_tid, uint8_t * score_diff, int32_t * mate, int32_t * dup_of, uint8_t * leftover)
{
  uint32_t const n = b->n_reads;
  std::vector<std::unordered_map<std::string, int32_t>> maps; // one read-name map per read group (hts_parallel_reader.cpp:270-337)
  int32_t prev = -1;                                          // last record that was not a duplicate (:666-684)
  for (uint32_t k = 0; k < n; ++k)
  {
    gtb_bam_core const & c = b->core[k];
    const uint8_t * d = b->data + b->data_off[k];
    long const l_data = (long)(b->data_off[k + 1] - b->data_off[k]);
    long const o_seq = (long)c.l_qname + 4l * c.n_cigar, n_seq = (c.l_qseq + 1l) / 2l, o_aux = o_seq + n_seq + c.l_qseq;
    if (c.l_qseq < 0 || c.l_qseq > (int32_t)(2 * GTB_SEQ_STRIDE) || o_aux > l_data)
    {
      g_err = "malformed record";
      return GTB_ERR_ARG;
    }
    memcpy(seq4 + (size_t)k * GTB_SEQ_STRIDE, d + o_seq, (size_t)n_seq);
    lseq[k] = (uint16_t)c.l_qseq;
    flag[k] = c.flag;
    mapq[k] = c.mapq;
    isize[k] = (int32_t)std::max<int64_t>(std::min<int64_t>(c.isize, INT32_MAX), INT32_MIN);
    same_tid[k] = c.tid == c.mtid;
    score_diff[k] = oracle_score_diff(d + o_aux, l_data - o_aux);
    leftover[k] = 0;
    // duplicate shortcut: equal_pos_seq against the last non-duplicate record (hts_utils.hpp:110-128)
    dup_of[k] = -1;
    if (prev >= 0)
    {
      gtb_bam_core const & p = b->core[prev];
      if (p.tid == c.tid && p.pos == c.pos && p.l_qseq == c.l_qseq &&
          memcmp(seq4 + (size_t)prev * GTB_SEQ_STRIDE, seq4 + (size_t)k * GTB_SEQ_STRIDE, (size_t)n_seq) == 0)
        dup_of[k] = prev;
    }
    if (dup_of[k] < 0)
      prev = (int32_t)k;
    // read-name map
    if ((size_t)b->rg[k] >= maps.size())
      maps.resize((size_t)b->rg[k] + 1);
    auto & m = maps[b->rg[k]];
    std::string const name(reinterpret_cast<const char *>(d));
    auto it = m.find(name);
    mate[k] = -1;
    if (it != m.end())
    {
      mate[k] = it->second;
      m.erase(it);
    }
    else if (c.flag & IS_PAIRED)
      m[name] = (int32_t)k;
  }
  if (is_sv) // whatever still waits at the end of the pool is processed alone (hts_parallel_reader.cpp:719-772)
    for (auto const & m : maps)
      for (auto const & kv : m)
        leftover[kv.second] = 1;
  return 0;
}

void gto_set_connections(int on) { g_connections = on != 0; }

int gto_result_connections_size(void * r, uint64_t * n)
{
  uint64_t k = 0;
  for (auto const & kv : ((PoolResult *)r)->w.conn)
    if (kv.second != 0)
      ++k;
  *n = k;
  return 0;
}

int gto_result_connections(void * r, gtb_connection * out)
{
  for (auto const & kv : ((PoolResult *)r)->w.conn)
    if (kv.second != 0)
    {
      out->sample = kv.first[0];
      out->hap1 = (uint16_t)kv.first[1];
      out->allele1 = (uint16_t)kv.first[2];
      out->hap2 = (uint16_t)kv.first[3];
      out->allele2 = (uint16_t)kv.first[4];
      out->count = kv.second;
      ++out;
    }
  return 0;
}

// the `ph` block of parallel_reader_genotype_only (hts_parallel_reader.cpp:782-893)
int gto_phase_support(const gtb_accumulators * acc, uint64_t n_conn, const gtb_connection * conn, uint64_t * n_out,
                      gtb_phase_support_entry * out)
{
  long const NB = acc->n_bubbles, NS = acc->n_samples;
  // hap_samples[s].connections[cov1] : hap2 -> vector<uint16_t>(hap2.gt.num)
  std::map<std::array<uint32_t, 4>, std::vector<uint16_t>> C;
  for (uint64_t i = 0; i < n_conn; ++i)
  {
    auto & v = C[{conn[i].sample, conn[i].hap1, conn[i].allele1, conn[i].hap2}];
    if (v.empty())
      v.assign(acc->n_alleles[conn[i].hap2], 0);
    v[conn[i].allele2] = (uint16_t)conn[i].count;
  }
  std::map<std::pair<HapAllele, HapAllele>, int8_t> ph;
  for (long ps1 = 0; ps1 < NB - 1; ++ps1)
  {
    long const order1 = acc->bubble_id[ps1];
    long const num1 = acc->n_alleles[ps1];
    for (long ps2 = ps1 + 1; ps2 < NB; ++ps2)
    {
      long const order2 = acc->bubble_id[ps2];
      long const num2 = acc->n_alleles[ps2];
      if (order2 >= order1 + 100)
        break;
      for (long s = 0; s < NS; ++s)
      {
        const uint16_t * cov1v = acc->gt_coverage + acc->cov_off[ps1] * NS + s * num1;
        const uint16_t * cov2v = acc->gt_coverage + acc->cov_off[ps2] * NS + s * num2;
        double const sum1 = std::accumulate(cov1v, cov1v + num1, 0.0), sum2 = std::accumulate(cov2v, cov2v + num2, 0.0);
        for (long cov1 = 1; cov1 < num1; ++cov1)
        {
          auto f = C.find({(uint32_t)s, (uint32_t)ps1, (uint32_t)cov1, (uint32_t)(uint16_t)ps2});
          if (f == C.end())
            continue;
          bool const clearly1 = cov1v[cov1] >= 4 || ((double)cov1v[cov1] / sum1) >= 0.28;
          bool const not1 = cov1v[cov1] <= 2 || ((double)cov1v[cov1] / sum1) < 0.22;
          std::vector<uint16_t> const & support_vec = f->second;
          long const total_support = std::accumulate(support_vec.begin(), support_vec.end(), 0l);
          for (long cov2 = 1; cov2 < (long)support_vec.size(); ++cov2)
          {
            double const support = (double)support_vec[cov2];
            int8_t is_good = 0;
            bool const clearly2 = cov2v[cov2] >= 4 || ((double)cov2v[cov2] / sum2) >= 0.28;
            bool const not2 = cov2v[cov2] <= 2 || ((double)cov2v[cov2] / sum2) < 0.22;
            if (not1 && not2)
              continue;
            if ((not1 && clearly2) || (not2 && clearly1))
              is_good = 2;
            else
            {
              if (total_support <= 2)
                continue;
              if (clearly1 && clearly2 && support / (double)total_support > 0.78)
                is_good = 1;
              else if (support / (double)total_support < 0.22)
                is_good = 2;
              else
                continue;
            }
            ph[{{(uint16_t)ps1, (uint16_t)cov1}, {(uint16_t)ps2, (uint16_t)cov2}}] |= is_good;
          }
        }
      }
    }
  }
  if (out)
  {
    if (*n_out < ph.size())
    {
      g_err = "phase support output too small";
      return GTB_ERR_ARG;
    }
    for (auto const & kv : ph)
    {
      memset(out, 0, sizeof(*out));
      out->hap1 = kv.first.first.first;
      out->allele1 = kv.first.first.second;
      out->hap2 = kv.first.second.first;
      out->allele2 = kv.first.second.second;
      out->flags = kv.second;
      ++out;
    }
  }
  *n_out = ph.size();
  return 0;
}

// get_haplotype_phred (vcf.cpp:47-81) + SampleCall::get_gt_call/get_gq (sample_call.cpp:78-131)
int gto_calls_from_accumulators(const gtb_accumulators * acc, uint8_t * phred, uint16_t * gt, uint8_t * gq)
{
  uint32_t const NB = acc->n_bubbles, NS = acc->n_samples;
  for (uint32_t b = 0; b < NB; ++b)
  {
    uint64_t const tri = acc->score_off[b + 1] - acc->score_off[b];
    uint32_t const cnum = acc->n_alleles[b];
    for (uint32_t s = 0; s < NS; ++s)
    {
      const uint16_t * ls = acc->log_score + acc->score_off[b] * NS + s * tri;
      uint8_t * ph = phred + acc->score_off[b] * NS + s * tri;
      uint16_t const mx = *std::max_element(ls, ls + tri);
      bool all_same = true;
      for (uint64_t i = 0; i < tri; ++i)
        if (ls[i] != mx)
          all_same = false;
      for (uint64_t i = 0; i < tri; ++i)
      {
        if (all_same)
          ph[i] = 0;
        else
        {
          long const sc = std::llround((mx - ls[i]) * 3.01029995663981195213738894724493026768189881462108541);
          ph[i] = sc < 255 ? (uint8_t)sc : 255;
        }
      }
      uint16_t gx = 0, gy = 0;
      {
        uint64_t i = 0;
        bool done = false;
        for (uint32_t y = 0; y < cnum && !done; ++y)
          for (uint32_t x = 0; x <= y; ++x, ++i)
            if (ph[i] == 0)
            {
              gx = (uint16_t)x;
              gy = (uint16_t)y;
              done = true;
              break;
            }
      }
      gt[((uint64_t)b * NS + s) * 2 + 0] = gx;
      gt[((uint64_t)b * NS + s) * 2 + 1] = gy;
      bool seen_zero = false;
      uint8_t nl = 255;
      bool two_zero = false;
      for (uint64_t i = 0; i < tri; ++i)
      {
        if (ph[i] == 0)
        {
          if (!seen_zero)
            seen_zero = true;
          else
          {
            two_zero = true;
            break;
          }
        }
        else if (ph[i] < nl)
          nl = ph[i];
      }
      gq[(uint64_t)b * NS + s] = two_zero ? 0 : nl;
    }
  }
  return 0;
}
}
