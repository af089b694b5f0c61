// gtb_index_host.hpp -- host-side k-mer index construction for one region graph (product code).
//
// Replaces PHIndex index_graph(Graph const&) (reference: src/index/indexer.cpp:246-291).  The reference simulates
// a deque-of-deques of open IndexEntry objects while sweeping the graph forwards.  Here the index is stated
// as a SET PROPERTY instead and enumerated backwards, one independent job per k-mer END position:
//
//   The index holds one entry per 32-base walk through the bubble graph (starting anywhere, also inside an
//   allele) that (a) contains only A/C/G/T, (b) satisfies the allele-combination limit of
//   indexer.cpp:13-20,198-211 -- with c = number of non-reference alleles on the walk and n = product of the
//   allele counts of those bubbles: not (c > 1 and (n > 181 or c > 4)) -- and (c) never enters a var node whose
//   events intersect the anti_events of an earlier var node of the walk (indexer.cpp:114-121).
//   Its labels are (start, end, var_id) for every var node touched, var ids ascending, or one variant-less label;
//   start/end inside an alternative allele beyond the reference allele's reach are special-position encoded
//   (indexer.cpp:146-147).
//
// Bucket order (which the aligner's path order depends on, SURVEY.md section 7 "hard parts") is the reference's
// emission order: END positions in sweep order (ref node r, then alleles 0..n-1 of bubble r, ...), and within
// one END position walks ordered by their allele choices compared from the most recent bubble backwards --
// exactly the order a depth-first backward enumeration with ascending allele loops produces.
//
// Because every END position is an independent job the same formulation maps 1:1 onto a device kernel
// (enumerate -> radix sort by (key, sequence number) -> table); see DESIGN.md "next".
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/gtb200.h"

namespace gtb
{
struct IndexSlot // 16-byte open-addressing slot; cnt == 0 means empty
{
  uint64_t key;
  uint32_t off;
  uint32_t cnt;
};

struct HostIndex
{
  std::vector<gtb_label> labels;   // grouped by key (first-seen key order); bucket order = reference insertion order
  std::vector<IndexSlot> uniq;     // one {key, label offset, label count} per distinct k-mer, first-seen order
  uint64_t n_keys = 0;
  // geometry of the DEVICE table (built on the device from `uniq`): power-of-two capacity, load <= 0.25
  uint32_t table_cap = 16;
  uint32_t table_mask = 15;
  int table_shift = 60;

  // PHIndex-style view for inspection / parity tests: keys ascending, labels concatenated in that order
  void export_sorted(std::vector<uint64_t> & keys, std::vector<uint32_t> & label_off, std::vector<gtb_label> & out) const
  {
    std::vector<uint32_t> order(uniq.size());
    for (uint32_t i = 0; i < order.size(); ++i)
      order[i] = i;
    std::sort(order.begin(), order.end(), [this](uint32_t a, uint32_t b) { return uniq[a].key < uniq[b].key; });
    keys.clear();
    label_off.assign(1, 0);
    out.clear();
    out.reserve(labels.size());
    for (uint32_t i : order)
    {
      keys.push_back(uniq[i].key);
      out.insert(out.end(), labels.begin() + uniq[i].off, labels.begin() + uniq[i].off + uniq[i].cnt);
      label_off.push_back((uint32_t)out.size());
    }
  }
};

inline uint64_t hash_key(uint64_t k) { return k * 0x9E3779B97F4A7C15ull; }

class IndexBuilder
{
public:
  explicit IndexBuilder(const gtb_graph_view & g) : g_(g) {}

  // returns false and sets err on a malformed view
  bool build(HostIndex & out, const char ** err)
  {
    if (g_.n_ref == 0)
    {
      finish(out);
      return true;
    }
    tuples_.clear();
    uint32_t const NR = g_.n_ref;
    for (uint32_t r = 0; r < NR; ++r)
    {
      uint32_t const n = ref_len(r);
      // A window that lies entirely inside this ref node is the only walk ending there: roll its key along the
      // node instead of walking backwards (the bulk of all end positions).
      const uint8_t * dna = ref_dna(r);
      uint64_t key = 0;
      uint32_t run = 0; // consecutive ACGT bases ending at d
      for (uint32_t d = 0; d < n; ++d)
      {
        int const c = code_of(dna[d]);
        if (c < 0)
        {
          run = 0;
          key = 0;
          continue; // a non-ACGT base kills every walk through it
        }
        key = (key << 2) | (uint64_t)c;
        ++run;
        if (d >= 31)
        {
          if (run >= 32)
            tuples_.push_back({key, {g_.ref_order[r] + d - 31, g_.ref_order[r] + d, GTB_INVALID_ID}});
        }
        else
          enumerate_end(false, r, d);
      }
      if (r + 1 < NR)
      {
        for (uint32_t v = g_.ref_var_off[r]; v < g_.ref_var_off[r + 1]; ++v)
        {
          uint32_t const m = var_len(v);
          for (uint32_t d = 0; d < m; ++d)
            enumerate_end(true, v, d);
        }
      }
    }
    if (bad_)
    {
      *err = "special position lookup failed while indexing (inconsistent graph view)";
      return false;
    }
    finish(out);
    return true;
  }

private:
  struct Tuple // in emission order (= the reference's insertion order)
  {
    uint64_t key;
    gtb_label label;
  };

  struct Walk // state of the backward DFS
  {
    uint64_t key = 0;
    int depth = 0;        // bases collected so far (from the end)
    uint32_t vars[40];    // var nodes touched, in backward order
    uint8_t var_bases[40]; // bases of that var node on the walk (>= 2 enables the self anti-event corner)
    int nvars = 0;
    uint32_t alt_count = 0; // non-reference alleles on the walk so far
    uint64_t alt_prod = 1;  // product of their bubbles' allele counts
  };

  const gtb_graph_view & g_;
  std::vector<Tuple> tuples_;
  bool bad_ = false;
  uint32_t end_pos_ = 0; // encoded end position of the current job

  uint32_t ref_len(uint32_t r) const { return (uint32_t)(g_.ref_seq_off[r + 1] - g_.ref_seq_off[r]); }
  uint32_t var_len(uint32_t v) const { return (uint32_t)(g_.var_seq_off[v + 1] - g_.var_seq_off[v]); }
  const uint8_t * ref_dna(uint32_t r) const { return g_.seq + g_.ref_seq_off[r]; }
  const uint8_t * var_dna(uint32_t v) const { return g_.seq + g_.var_seq_off[v]; }
  uint32_t bubble_of(uint32_t v) const { return g_.var_out_ref[v] - 1; }
  uint32_t bubble_ref_reach(uint32_t v) const
  {
    uint32_t const v0 = g_.ref_var_off[bubble_of(v)];
    return g_.var_order[v0] + var_len(v0) - 1;
  }

  uint32_t special(uint32_t pos, uint32_t ref_reach)
  {
    const uint32_t * b = g_.sp_keys;
    const uint32_t * e = g_.sp_keys + g_.n_sp_keys;
    const uint32_t * it = std::lower_bound(b, e, ref_reach);
    if (it == e || *it != ref_reach)
    {
      bad_ = true;
      return GTB_INVALID_ID;
    }
    uint32_t const k = (uint32_t)(it - b);
    uint32_t const idx = pos - ref_reach - 1;
    if (g_.sp_off[k] + idx >= g_.sp_off[k + 1])
    {
      bad_ = true;
      return GTB_INVALID_ID;
    }
    return g_.sp_list[g_.sp_off[k] + idx];
  }

  uint32_t encode_var_pos(uint32_t v, uint32_t d)
  {
    uint32_t pos = g_.var_order[v] + d;
    uint32_t const rr = bubble_ref_reach(v);
    if (pos > rr)
      pos = special(pos, rr);
    return pos;
  }

  static inline int code_of(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

  // consume bases dna[hi], dna[hi-1], ... dna[0] (as many as still needed); returns false on a non-ACGT base
  static inline bool take_backward(Walk & w, const uint8_t * dna, int hi, int & taken)
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

  bool events_ok(const Walk & w) const
  {
    if (!g_.var_ev || !g_.var_aev)
      return true;
    // forward order = reverse of w.vars; accumulate anti_events of earlier nodes
    for (int i = w.nvars - 1; i >= 0; --i)
    {
      uint32_t const v = w.vars[i];
      const int64_t * eb = g_.var_ev + g_.var_ev_off[v];
      const int64_t * ee = g_.var_ev + g_.var_ev_off[v + 1];
      if (eb == ee)
        continue;
      for (int m = w.nvars - 1; m >= i; --m)
      {
        if (m == i && w.var_bases[i] < 2)
          continue; // a node's own anti_events only apply from its second base on
        uint32_t const u = w.vars[m];
        for (const int64_t * a = g_.var_aev + g_.var_aev_off[u]; a != g_.var_aev + g_.var_aev_off[u + 1]; ++a)
          if (std::binary_search(eb, ee, *a))
            return false;
      }
    }
    return true;
  }

  void emit(const Walk & w, uint32_t start_pos)
  {
    if (!events_ok(w))
      return;
    if (w.nvars == 0)
    {
      tuples_.push_back({w.key, {start_pos, end_pos_, GTB_INVALID_ID}});
      return;
    }
    for (int i = w.nvars - 1; i >= 0; --i) // ascending var id = forward order
      tuples_.push_back({w.key, {start_pos, end_pos_, w.vars[i]}});
  }

  // continue the walk backwards from the END of ref node r (all of r's bases are candidates)
  void back_into_ref(Walk w, uint32_t r)
  {
    int const n = (int)ref_len(r);
    int taken;
    if (!take_backward(w, ref_dna(r), n - 1, taken))
      return;
    if (w.depth == 32)
    {
      emit(w, g_.ref_order[r] + (uint32_t)(n - taken));
      return;
    }
    if (r == 0)
      return; // ran out of graph
    back_into_bubble(w, r - 1);
  }

  // continue backwards through bubble b (the alleles between ref node b and b+1), ascending allele order
  void back_into_bubble(const Walk & w0, uint32_t b)
  {
    uint32_t const vb = g_.ref_var_off[b], ve = g_.ref_var_off[b + 1];
    uint32_t const nalleles = ve - vb;
    for (uint32_t v = vb; v < ve; ++v)
    {
      Walk w = w0;
      if (v != vb)
      {
        ++w.alt_count;
        w.alt_prod *= nalleles;
        if (w.alt_prod > 0xFFFFFFFFull)
          w.alt_prod = 0xFFFFFFFFull;
        if (w.alt_count > 1 && (w.alt_prod > 181 || w.alt_count > 4))
          continue;
      }
      int const m = (int)var_len(v);
      int taken;
      if (!take_backward(w, var_dna(v), m - 1, taken))
        continue;
      if (taken > 0)
      {
        if (w.nvars >= 40)
          continue;
        w.vars[w.nvars] = v;
        w.var_bases[w.nvars] = (uint8_t)std::min(taken, 255);
        ++w.nvars;
      }
      if (w.depth == 32)
      {
        emit(w, encode_var_pos(v, (uint32_t)(m - taken)));
        continue;
      }
      back_into_ref(w, b);
    }
  }

  void enumerate_end(bool in_var, uint32_t node, uint32_t d)
  {
    Walk w;
    int taken;
    if (!in_var)
    {
      end_pos_ = g_.ref_order[node] + d;
      if (!take_backward(w, ref_dna(node), (int)d, taken))
        return;
      if (w.depth == 32)
      {
        emit(w, g_.ref_order[node] + d + 1 - (uint32_t)taken);
        return;
      }
      if (node == 0)
        return;
      back_into_bubble(w, node - 1);
    }
    else
    {
      uint32_t const v = node;
      uint32_t const b = bubble_of(v);
      uint32_t const vb = g_.ref_var_off[b];
      end_pos_ = encode_var_pos(v, d);
      if (v != vb)
      {
        w.alt_count = 1;
        w.alt_prod = g_.ref_var_off[b + 1] - vb;
      }
      if (!take_backward(w, var_dna(v), (int)d, taken))
        return;
      w.vars[0] = v;
      w.var_bases[0] = (uint8_t)std::min(taken, 255);
      w.nvars = 1;
      if (w.depth == 32)
      {
        emit(w, encode_var_pos(v, d + 1 - (uint32_t)taken));
        return;
      }
      back_into_ref(w, b);
    }
  }

  void finish(HostIndex & out)
  {
    size_t const n = tuples_.size();
    // host-side grouping table (load <= 0.5, values = index into out.uniq); no sort: the labels of one key keep
    // their emission order by construction
    size_t gcap = 16;
    while (gcap < n * 2 + 2)
      gcap <<= 1;
    int gshift = 64;
    for (size_t c = gcap; c > 1; c >>= 1)
      --gshift;
    std::vector<uint32_t> gtab(gcap, 0xFFFFFFFFu);
    std::vector<uint32_t> bucket_of(n);
    out.uniq.clear();
    out.uniq.reserve(n);
    for (size_t i = 0; i < n; ++i)
    {
      uint64_t const k = tuples_[i].key;
      size_t h = (size_t)(hash_key(k) >> gshift);
      while (gtab[h] != 0xFFFFFFFFu && out.uniq[gtab[h]].key != k)
        h = (h + 1) & (gcap - 1);
      if (gtab[h] == 0xFFFFFFFFu)
      {
        gtab[h] = (uint32_t)out.uniq.size();
        out.uniq.push_back(IndexSlot{k, 0, 0});
      }
      bucket_of[i] = gtab[h];
      ++out.uniq[gtab[h]].cnt;
    }
    out.n_keys = out.uniq.size();
    uint32_t running = 0;
    for (auto & u : out.uniq) // first-seen key order
    {
      u.off = running;
      running += u.cnt;
      u.cnt = 0; // reused as write cursor, restored below
    }
    out.labels.resize(n);
    for (size_t i = 0; i < n; ++i)
    {
      IndexSlot & u = out.uniq[bucket_of[i]];
      out.labels[u.off + u.cnt++] = tuples_[i].label;
    }
    size_t cap = 16;
    while (cap < n * 4 + 2) // load factor <= 0.25: an unsuccessful lookup (the 96 neighbours) averages ~1.2 probes
      cap <<= 1;
    out.table_cap = (uint32_t)cap;
    out.table_mask = (uint32_t)(cap - 1);
    int shift = 64;
    for (size_t c = cap; c > 1; c >>= 1)
      --shift;
    out.table_shift = shift;
    tuples_.clear();
    tuples_.shrink_to_fit();
  }
};

} // namespace gtb
