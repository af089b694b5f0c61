// gtb_shim.hpp -- the reference-side binding of libgtb200: what a graphtyper maintainer adds to the tree.
//
// Header-only C++17 on top of the reference's own headers and include/gtb200.h.  It does three things and nothing else:
//   flatten()   gyper::Graph            -> gtb_graph_view        (input of gtb_region_begin; replaces index_graph's argument)
//   Records     HtsRecord stream        -> gtb_bam_batch         (input of gtb_submit_bam_records; replaces genotype_only's
//                                                                 per-record work, src/utilities/hts_parallel_reader.cpp:245-338)
//   die()       gtb_status              -> print_log(error) + std::exit(1), the reference's error convention
//
// Compiled and exercised against the unmodified reference objects by oracle/ref_build/shim_probe.cpp
// (tests/test_shim_against_reference.py).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <vector>

#include <graphtyper/graph/graph.hpp>
#include <graphtyper/utilities/hts_parallel_reader.hpp>
#include <graphtyper/utilities/logging.hpp>

#include <gtb200.h>

namespace gtb_shim
{
inline void die(int rc)
{
  if (rc != 0)
  {
    gyper::print_log(gyper::log_severity::error, "[gtb200] ", gtb_last_error());
    std::exit(1);
  }
}

// Owns the arrays behind a gtb_graph_view.
struct FlatGraph
{
  std::vector<uint32_t> ref_order, ref_var_off, var_order, var_out_ref, var_ev_off, var_aev_off, sp_keys, sp_off, sp_list,
    actual_poses, ref_reach_poses;
  std::vector<uint64_t> ref_seq_off, var_seq_off;
  std::vector<uint8_t> seq;
  std::vector<int64_t> var_ev, var_aev;
  gtb_graph_view view{};
};

// include/graphtyper/graph/graph.hpp:39-171: ref node r is followed by the consecutive var nodes of its bubble, which all lead
// to ref node r + 1 (graph.cpp:549-620).
inline FlatGraph flatten(gyper::Graph const & g)
{
  FlatGraph f;
  f.ref_var_off.push_back(0);
  f.var_ev_off.push_back(0);
  f.var_aev_off.push_back(0);
  for (auto const & rn : g.ref_nodes)
  {
    f.ref_order.push_back(rn.get_label().order);
    f.ref_seq_off.push_back(f.seq.size());
    f.seq.insert(f.seq.end(), rn.get_label().dna.begin(), rn.get_label().dna.end());
    f.ref_var_off.push_back(f.ref_var_off.back() + rn.out_degree());
  }
  f.ref_seq_off.push_back(f.seq.size());
  for (auto const & vn : g.var_nodes)
  {
    f.var_order.push_back(vn.get_label().order);
    f.var_out_ref.push_back(vn.get_out_ref_index());
    f.var_seq_off.push_back(f.seq.size());
    f.seq.insert(f.seq.end(), vn.get_label().dna.begin(), vn.get_label().dna.end());
    std::vector<int64_t> ev(vn.events.begin(), vn.events.end()), aev(vn.anti_events.begin(), vn.anti_events.end());
    std::sort(ev.begin(), ev.end());
    std::sort(aev.begin(), aev.end());
    f.var_ev.insert(f.var_ev.end(), ev.begin(), ev.end());
    f.var_ev_off.push_back(f.var_ev.size());
    f.var_aev.insert(f.var_aev.end(), aev.begin(), aev.end());
    f.var_aev_off.push_back(f.var_aev.size());
  }
  f.var_seq_off.push_back(f.seq.size());
  f.actual_poses.assign(g.actual_poses.begin(), g.actual_poses.end());
  f.ref_reach_poses.assign(g.ref_reach_poses.begin(), g.ref_reach_poses.end());
  f.sp_off.push_back(0);
  {
    std::map<uint32_t, std::vector<uint32_t>> sorted(g.ref_reach_to_special_pos.begin(), g.ref_reach_to_special_pos.end());
    for (auto const & kv : sorted)
    {
      f.sp_keys.push_back(kv.first);
      f.sp_list.insert(f.sp_list.end(), kv.second.begin(), kv.second.end());
      f.sp_off.push_back(f.sp_list.size());
    }
  }
  gtb_graph_view & v = f.view;
  v.n_ref = (uint32_t)f.ref_order.size();
  v.n_var = (uint32_t)f.var_order.size();
  v.is_sv_graph = g.is_sv_graph ? 1 : 0;
  v.ref_order = f.ref_order.data();
  v.ref_seq_off = f.ref_seq_off.data();
  v.ref_var_off = f.ref_var_off.data();
  v.var_order = f.var_order.data();
  v.var_seq_off = f.var_seq_off.data();
  v.var_out_ref = f.var_out_ref.data();
  v.seq = f.seq.data();
  v.seq_len = f.seq.size();
  v.var_ev_off = f.var_ev_off.data();
  v.var_ev = f.var_ev.data();
  v.var_aev_off = f.var_aev_off.data();
  v.var_aev = f.var_aev.data();
  v.n_special = (uint32_t)f.actual_poses.size();
  v.actual_poses = f.actual_poses.data();
  v.ref_reach_poses = f.ref_reach_poses.data();
  v.n_sp_keys = (uint32_t)f.sp_keys.size();
  v.sp_keys = f.sp_keys.data();
  v.sp_off = f.sp_off.data();
  v.sp_list = f.sp_list.data();
  return f;
}

// The records of one pool in merge order, as htslib holds them (gtb_bam_batch).  add() is the whole per-record work left on
// the host: copy the core fields and the data block.
struct Records
{
  std::vector<gtb_bam_core> core;
  std::vector<uint8_t> data;
  std::vector<uint64_t> off{0};
  std::vector<int32_t> sample, rg;

  void add(gyper::HtsParallelReader const & reader, gyper::HtsRecord const & rec)
  {
    bam1_t const * r = rec.record;
    long sample_i = 0, rg_i = 0;
    reader.get_sample_and_rg_index(sample_i, rg_i, rec);
    gtb_bam_core c{};
    c.pos = r->core.pos;
    c.mpos = r->core.mpos;
    c.isize = r->core.isize;
    c.tid = r->core.tid;
    c.mtid = r->core.mtid;
    c.l_qseq = r->core.l_qseq;
    c.n_cigar = r->core.n_cigar;
    c.flag = r->core.flag;
    c.l_qname = r->core.l_qname;
    c.mapq = r->core.qual;
    core.push_back(c);
    data.insert(data.end(), r->data, r->data + r->l_data);
    off.push_back(data.size());
    sample.push_back((int32_t)sample_i);
    rg.push_back((int32_t)rg_i);
  }

  gtb_bam_batch view() const
  {
    gtb_bam_batch b{};
    b.n_reads = (uint32_t)core.size();
    b.core = core.data();
    b.data = data.data();
    b.data_off = off.data();
    b.sample = sample.data();
    b.rg = rg.data();
    return b;
  }
};
} // namespace gtb_shim
