// oracle/ref_build/gt_probe.cpp -- TEST INFRASTRUCTURE (golden-vector generator), not product code.
//
// Links against the UNMODIFIED reference objects (oracle/ref_build/Makefile) and drives the reference's
// own functions for the genotyping hot path, dumping every intermediate the parity tests pin:
//
//   <out>.graph.gtba  flat view of gyper::graph after construct_graph()              (src/graph/constructor.cpp:1597)
//   <out>.index.gtba  PHIndex contents of index_graph(), labels in bucket order      (src/index/indexer.cpp:246)
//   <out>.reads.gtba  every record the pool loop passes to genotype_only(), its seed-query results
//                     (query_index / query_index_hamming_distance1_without_index,     src/utilities/kmer_help_functions.cpp:51,93)
//                     and its align_read() GenotypePaths                              (src/typer/alignment.cpp:331)
//   <out>.accum.gtba  VcfWriter::haplotypes[*].hap_samples[*] + var_stats after the pool loop
//                     (src/typer/vcf_writer.cpp:143, src/graph/haplotype.cpp:462) and the SampleCall/VarStats
//                     produced by Vcf::add_haplotype + Variant::scan_calls            (src/typer/vcf.cpp:1507, variant.cpp:230)
//
// The pool loop below is our own restatement of parallel_reader_genotype_only()'s driver
// (src/utilities/hts_parallel_reader.cpp:458-716) expressed as calls into the reference's functions
// (HtsParallelReader, genotype_only, align_read, VcfWriter); nothing is copied from the reference.
//
// File format "GTBA": 8-byte magic, u64 n_arrays, then per array: char name[48], u32 elem_size, u32 kind
// (0 unsigned, 1 signed, 2 float), u64 count, raw little-endian data padded to 8 bytes.

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <graphtyper/constants.hpp>
#include <graphtyper/graph/absolute_position.hpp>
#include <graphtyper/graph/constructor.hpp>
#include <graphtyper/graph/genomic_region.hpp>
#include <graphtyper/graph/graph.hpp>
#include <graphtyper/graph/reference_depth.hpp>
#include <graphtyper/index/indexer.hpp>
#include <graphtyper/index/ph_index.hpp>
#include <graphtyper/typer/alignment.hpp>
#include <graphtyper/typer/genotype_paths.hpp>
#include <graphtyper/typer/primers.hpp>
#include <graphtyper/typer/variant.hpp>
#include <graphtyper/typer/vcf.hpp>
#include <graphtyper/typer/vcf_writer.hpp>
#include <graphtyper/utilities/hts_parallel_reader.hpp>
#include <graphtyper/utilities/kmer_help_functions.hpp>
#include <graphtyper/utilities/logging.hpp>
#include <graphtyper/utilities/options.hpp>

#include <seqan/sequence.h>

namespace gyper
{
// defined with external linkage in src/utilities/hts_parallel_reader.cpp:226,245 (not declared in a header)
void get_sequence(seqan::IupacString & seq, seqan::IupacString & rseq, bam1_t const * rec);

void genotype_only(HtsParallelReader const & hts_preader,
                   VcfWriter & writer,
                   ReferenceDepth & reference_depth,
                   std::vector<std::unordered_map<std::string, std::pair<GenotypePaths, GenotypePaths>>> & maps,
                   std::pair<GenotypePaths, GenotypePaths> & prev_paths,
                   PHIndex const & ph_index,
                   Primers const * primers,
                   HtsRecord const & hts_rec,
                   seqan::IupacString & seq,
                   seqan::IupacString & rseq,
                   bool update_prev_paths,
                   bool const IS_SV_CALLING);
} // namespace gyper

namespace
{
struct ArrayFile
{
  struct Arr
  {
    std::string name;
    uint32_t esize;
    uint32_t kind;
    std::vector<uint8_t> bytes;
  };
  std::vector<Arr> arrs;

  template <typename T>
  void add(std::string const & name, std::vector<T> const & v, uint32_t kind = 0)
  {
    Arr a;
    a.name = name;
    a.esize = sizeof(T);
    a.kind = kind;
    a.bytes.resize(v.size() * sizeof(T));
    if (!v.empty())
      memcpy(a.bytes.data(), v.data(), a.bytes.size());
    arrs.push_back(std::move(a));
  }

  void write(std::string const & path) const
  {
    FILE * f = fopen(path.c_str(), "wb");
    if (!f)
    {
      fprintf(stderr, "cannot write %s\n", path.c_str());
      exit(2);
    }
    fwrite("GTBA0001", 1, 8, f);
    uint64_t n = arrs.size();
    fwrite(&n, 8, 1, f);
    for (auto const & a : arrs)
    {
      char name[48];
      memset(name, 0, sizeof(name));
      strncpy(name, a.name.c_str(), 47);
      fwrite(name, 1, 48, f);
      fwrite(&a.esize, 4, 1, f);
      fwrite(&a.kind, 4, 1, f);
      uint64_t count = a.bytes.size() / a.esize;
      fwrite(&count, 8, 1, f);
      fwrite(a.bytes.data(), 1, a.bytes.size(), f);
      static const char zeros[8] = {0};
      size_t pad = (8 - a.bytes.size() % 8) % 8;
      fwrite(zeros, 1, pad, f);
    }
    fclose(f);
  }
};

std::vector<std::string> split(std::string const & s, char d)
{
  std::vector<std::string> out;
  size_t b = 0;
  while (b <= s.size())
  {
    size_t e = s.find(d, b);
    if (e == std::string::npos)
      e = s.size();
    if (e > b)
      out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
  return out;
}

void dump_graph(std::string const & path)
{
  using namespace gyper;
  Graph const & g = graph;
  ArrayFile af;
  std::vector<uint64_t> meta = {g.is_sv_graph ? 1ull : 0ull,
                                g.ref_nodes.size(),
                                g.var_nodes.size(),
                                (uint64_t)g.genomic_region.begin,
                                (uint64_t)g.genomic_region.end,
                                (uint64_t)g.genomic_region.get_absolute_begin_position(),
                                (uint64_t)g.genomic_region.get_absolute_end_position()};
  af.add("meta", meta);

  std::vector<uint32_t> ref_order, ref_var_off(1, 0u), var_order, var_out_ref, var_ev_off(1, 0u), var_aev_off(1, 0u);
  std::vector<uint64_t> ref_seq_off, var_seq_off;
  std::vector<uint8_t> seq;
  std::vector<int64_t> var_ev, var_aev;

  for (auto const & rn : g.ref_nodes)
  {
    ref_order.push_back(rn.get_label().order);
    ref_seq_off.push_back(seq.size());
    seq.insert(seq.end(), rn.get_label().dna.begin(), rn.get_label().dna.end());
    // out vars must be consecutive (graph.cpp:549-620); record begin/end
    if (rn.out_degree() > 0)
    {
      if (rn.get_var_index(0) != ref_var_off.back())
      {
        fprintf(stderr, "non-consecutive var ids\n");
        exit(3);
      }
      for (unsigned i = 1; i < rn.out_degree(); ++i)
        if (rn.get_var_index(i) != rn.get_var_index(0) + i)
        {
          fprintf(stderr, "non-consecutive var ids within a bubble\n");
          exit(3);
        }
    }
    ref_var_off.push_back(ref_var_off.back() + rn.out_degree());
  }
  ref_seq_off.push_back(seq.size());

  for (auto const & vn : g.var_nodes)
  {
    var_order.push_back(vn.get_label().order);
    var_out_ref.push_back(vn.get_out_ref_index());
    var_seq_off.push_back(seq.size());
    seq.insert(seq.end(), vn.get_label().dna.begin(), vn.get_label().dna.end());
    std::vector<int64_t> ev(vn.events.begin(), vn.events.end());
    std::sort(ev.begin(), ev.end());
    var_ev.insert(var_ev.end(), ev.begin(), ev.end());
    var_ev_off.push_back(var_ev.size());
    std::vector<int64_t> aev(vn.anti_events.begin(), vn.anti_events.end());
    std::sort(aev.begin(), aev.end());
    var_aev.insert(var_aev.end(), aev.begin(), aev.end());
    var_aev_off.push_back(var_aev.size());
  }
  var_seq_off.push_back(seq.size());

  af.add("ref_order", ref_order);
  af.add("ref_seq_off", ref_seq_off);
  af.add("ref_var_off", ref_var_off);
  af.add("var_order", var_order);
  af.add("var_seq_off", var_seq_off);
  af.add("var_out_ref", var_out_ref);
  af.add("seq", seq);
  af.add("var_ev_off", var_ev_off);
  af.add("var_ev", var_ev, 1);
  af.add("var_aev_off", var_aev_off);
  af.add("var_aev", var_aev, 1);
  af.add("actual_poses", g.actual_poses);
  af.add("ref_reach_poses", g.ref_reach_poses);

  // ref_reach -> special position codes (graph.cpp:1759-1782), keys ascending
  std::vector<uint32_t> sp_keys, sp_off(1, 0u), sp_list;
  {
    std::map<uint32_t, std::vector<uint32_t>> sorted(g.ref_reach_to_special_pos.begin(),
                                                     g.ref_reach_to_special_pos.end());
    for (auto const & kv : sorted)
    {
      sp_keys.push_back(kv.first);
      sp_list.insert(sp_list.end(), kv.second.begin(), kv.second.end());
      sp_off.push_back(sp_list.size());
    }
  }
  af.add("sp_keys", sp_keys);
  af.add("sp_off", sp_off);
  af.add("sp_list", sp_list);

  std::vector<uint64_t> contig_len, contig_off;
  std::vector<uint8_t> contig_names;
  for (size_t i = 0; i < g.contigs.size(); ++i)
  {
    contig_len.push_back(g.contigs[i].length);
    contig_off.push_back(i < absolute_pos.offsets.size() ? absolute_pos.offsets[i] : 0);
    contig_names.insert(contig_names.end(), g.contigs[i].name.begin(), g.contigs[i].name.end());
    contig_names.push_back('\n');
  }
  af.add("contig_len", contig_len);
  af.add("contig_off", contig_off);
  af.add("contig_names", contig_names);
  af.write(path);
}

void dump_index(gyper::PHIndex const & idx, std::string const & path)
{
  using namespace gyper;
  std::vector<uint64_t> keys;
  keys.reserve(idx.hamming0.size());
  for (auto const & kv : idx.hamming0)
    keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  std::vector<uint32_t> off(1, 0u), labels;
  for (uint64_t k : keys)
  {
    auto const & v = idx.hamming0.at(k);
    for (auto const & l : v)
    {
      labels.push_back(l.start_index);
      labels.push_back(l.end_index);
      labels.push_back(l.variant_id);
    }
    off.push_back(labels.size() / 3);
  }
  ArrayFile af;
  af.add("keys", keys);
  af.add("label_off", off);
  af.add("labels", labels);
  af.write(path);
}

struct PathDump
{
  // per (record, orientation)
  std::vector<uint32_t> gp_npaths, gp_longest;
  // per path
  std::vector<uint32_t> p_start, p_end, p_rs, p_re, p_mm, p_nvar;
  // per path-variant
  std::vector<uint32_t> v_order, v_nnum;
  std::vector<uint16_t> v_nums;

  void add(gyper::GenotypePaths const & g)
  {
    gp_npaths.push_back(g.paths.size());
    gp_longest.push_back(g.longest_path_length);
    for (auto const & p : g.paths)
    {
      p_start.push_back(p.start);
      p_end.push_back(p.end);
      p_rs.push_back(p.read_start_index);
      p_re.push_back(p.read_end_index);
      p_mm.push_back(p.mismatches);
      p_nvar.push_back(p.var_order.size());
      for (size_t i = 0; i < p.var_order.size(); ++i)
      {
        v_order.push_back(p.var_order[i]);
        std::vector<uint16_t> nums(p.nums[i].begin(), p.nums[i].end());
        std::sort(nums.begin(), nums.end());
        v_nnum.push_back(nums.size());
        v_nums.insert(v_nums.end(), nums.begin(), nums.end());
      }
    }
  }

  void write(ArrayFile & af, std::string const & pre) const
  {
    af.add(pre + "gp_npaths", gp_npaths);
    af.add(pre + "gp_longest", gp_longest);
    af.add(pre + "p_start", p_start);
    af.add(pre + "p_end", p_end);
    af.add(pre + "p_rs", p_rs);
    af.add(pre + "p_re", p_re);
    af.add(pre + "p_mm", p_mm);
    af.add(pre + "p_nvar", p_nvar);
    af.add(pre + "v_order", v_order);
    af.add(pre + "v_nnum", v_nnum);
    af.add(pre + "v_nums", v_nums);
  }
};

struct SeedDump
{
  // per (record, orientation, ham) : nslots ; per slot: nlabels ; labels (start,end,var)
  std::vector<uint32_t> nslots, nlabels, labels;

  void add(gyper::TKmerLabels const & ll)
  {
    nslots.push_back(ll.size());
    for (auto const & s : ll)
    {
      nlabels.push_back(s.size());
      for (auto const & l : s)
      {
        labels.push_back(l.start_index);
        labels.push_back(l.end_index);
        labels.push_back(l.variant_id);
      }
    }
  }
};

// SV-only read pre-filter, restated from hts_parallel_reader.cpp:528-568
bool is_good_read_sv(bam1_t * record)
{
  auto const & core = record->core;
  if ((core.flag & gyper::IS_UNMAPPED) != 0u)
    return false;
  uint8_t * cigar_it = record->data + core.l_qname;
  long const n_cigar = core.n_cigar;
  bool const mate_far = core.tid != core.mtid || std::abs(core.pos - core.mpos) > 200000;
  if (core.qual <= 15 && mate_far)
    return false;
  if (n_cigar >= 2)
  {
    uint32_t a, b;
    memcpy(&a, cigar_it, 4);
    memcpy(&b, cigar_it + 4 * (n_cigar - 1), 4);
    bool const front_s = (a & 15) == 4, back_s = (b & 15) == 4;
    bool const one = (front_s && (a >> 4) >= 12) || (back_s && (b >> 4) >= 12);
    if ((front_s && back_s) || (core.qual <= 15 && one))
      return false;
  }
  return true;
}

} // namespace

int main(int argc, char ** argv)
{
  using namespace gyper;
  std::string ref_fn, vcf_fn, region_str, out, sams_arg;
  bool is_sv = false, dump_seeds = true, light = false;
  long pad = 1000;

  for (int i = 1; i < argc; ++i)
  {
    std::string a = argv[i];
    auto next = [&]() -> std::string
    {
      if (i + 1 >= argc)
      {
        fprintf(stderr, "missing value for %s\n", a.c_str());
        exit(2);
      }
      return argv[++i];
    };
    if (a == "--ref")
      ref_fn = next();
    else if (a == "--vcf")
      vcf_fn = next();
    else if (a == "--region")
      region_str = next();
    else if (a == "--out")
      out = next();
    else if (a == "--sams")
      sams_arg = next();
    else if (a == "--sv")
      is_sv = true;
    else if (a == "--pad")
      pad = std::stol(next());
    else if (a == "--no-seeds")
      dump_seeds = false;
    else if (a == "--light") // large pools: no index / per-read dumps, only the record stream's identity + accumulators
      light = true;
    else
    {
      fprintf(stderr, "unknown arg %s\n", a.c_str());
      return 2;
    }
  }

  if (ref_fn.empty() || vcf_fn.empty() || region_str.empty() || out.empty())
  {
    fprintf(stderr,
            "usage: gt_probe --ref R.fa --vcf V.vcf.gz --region chr:b-e --out PREFIX [--sams a.sam,b.sam] [--sv] "
            "[--pad N]\n");
    return 2;
  }

  Options & opts = *Options::instance();
  // same as setup_logger() in src/main.cpp:234-251 (warning level to std::clog)
  gyper::log_singleton =
    std::unique_ptr<gyper::log_singleton_t>{new gyper::log_singleton_t{gyper::log_severity::warning, std::clog}};
  opts.vcf = vcf_fn;
  opts.threads = 1;
  opts.no_bamshrink = true;

  // same steps as genotype() -> genotype_only_with_a_vcf (src/utilities/genotype.cpp:401-402,262-307)
  GenomicRegion region(region_str);
  GenomicRegion padded(region);
  padded.pad(pad);
  construct_graph(ref_fn, vcf_fn, padded.to_string(), is_sv, true);
  absolute_pos.calculate_offsets(graph.contigs);
  dump_graph(out + ".graph.gtba");

  PHIndex ph_index = index_graph(graph);
  if (!light)
    dump_index(ph_index, out + ".index.gtba");

  if (sams_arg.empty())
    return 0;

  std::vector<std::string> sams = split(sams_arg, ',');

  // ---- pool loop (restated driver of hts_parallel_reader.cpp:458-716, non-SV coverage filter omitted:
  //      avg_cov_by_readlen is -1 in genotype_only_with_a_vcf so update_bin_count() is a no-op there)
  HtsParallelReader hts_preader;
  hts_preader.open(sams, "", ".");
  VcfWriter writer(opts.split_var_threshold - 1);
  writer.set_samples(hts_preader.get_samples());
  ReferenceDepth reference_depth;
  if (graph.is_sv_graph)
    reference_depth.set_depth_sizes(writer.pns.size());
  std::vector<std::unordered_map<std::string, std::pair<GenotypePaths, GenotypePaths>>> maps;
  maps.resize(hts_preader.get_num_rg());

  // record columns
  std::vector<uint16_t> r_flag;
  std::vector<int64_t> r_pos, r_mpos, r_isize;
  std::vector<int32_t> r_tid, r_mtid, r_lseq, r_sample, r_rg, r_file;
  std::vector<uint8_t> r_mapq, r_isdup, r_seq4; // r_seq4: (l+1)/2 bytes per record, BAM nibbles
  std::vector<uint64_t> r_seq_off(1, 0), r_name_off(1, 0), r_cigar_off(1, 0);
  std::vector<uint8_t> r_names;
  std::vector<uint32_t> r_cigar;
  std::vector<uint8_t> r_score_diff_first; // score_diff as set by update_paths on geno.first
  // the records themselves as htslib holds them: bam1_t::data (qname | cigar | seq | qual | aux) + the core fields the
  // columns above do not carry -- input of the record-parsing path (gtb_submit_bam_records)
  std::vector<uint8_t> r_bam_data;
  std::vector<uint64_t> r_bam_off(1, 0);
  std::vector<uint32_t> r_lqname;
  PathDump pd;
  SeedDump sd0, sd1;
  // --light: which record of which file every processed record is (ordinal counts every record the reader delivered)
  std::vector<int32_t> r_ord;
  std::vector<int64_t> file_count(sams.size(), 0);

  std::pair<GenotypePaths, GenotypePaths> prev_paths;
  HtsRecord prev, curr;
  seqan::IupacString seq, rseq;
  bool const IS_SV = graph.is_sv_graph;
  bool have_prev = false;

  auto record_columns = [&](HtsRecord const & h, bool is_dup)
  {
    bam1_t * b = h.record;
    auto const & c = b->core;
    long sample_i = 0, rg_i = 0;
    hts_preader.get_sample_and_rg_index(sample_i, rg_i, h);
    if (light)
    {
      r_flag.push_back(c.flag);
      r_sample.push_back(sample_i);
      r_file.push_back(h.file_index);
      r_ord.push_back((int32_t)(file_count[h.file_index] - 1));
      {
        // record identity for the regenerating harness: the trailing decimal number of the read name (the generator's pair
        // id); the reader re-sorts same-position records of a file by sequence, so a delivery ordinal is not a file line
        char const * qn = bam_get_qname(b);
        size_t const len = strlen(qn);
        size_t k = len;
        while (k > 0 && qn[k - 1] >= '0' && qn[k - 1] <= '9')
          --k;
        r_pos.push_back(k < len ? std::strtoll(qn + k, nullptr, 10) : -1);
      }
      r_isdup.push_back(is_dup ? 1 : 0);
      std::pair<GenotypePaths, GenotypePaths> scratch =
        std::make_pair(GenotypePaths(c.flag, c.l_qseq), GenotypePaths(c.flag, c.l_qseq));
      update_paths(scratch, b);
      r_score_diff_first.push_back(scratch.first.score_diff);
      return;
    }
    r_flag.push_back(c.flag);
    r_pos.push_back(c.pos);
    r_mpos.push_back(c.mpos);
    r_isize.push_back(c.isize);
    r_tid.push_back(c.tid);
    r_mtid.push_back(c.mtid);
    r_lseq.push_back(c.l_qseq);
    r_sample.push_back(sample_i);
    r_rg.push_back(rg_i);
    r_file.push_back(h.file_index);
    r_mapq.push_back(c.qual);
    r_isdup.push_back(is_dup ? 1 : 0);
    uint8_t * s = bam_get_seq(b);
    r_seq4.insert(r_seq4.end(), s, s + (c.l_qseq + 1) / 2);
    r_seq_off.push_back(r_seq4.size());
    char const * qn = bam_get_qname(b);
    r_names.insert(r_names.end(), qn, qn + strlen(qn));
    r_name_off.push_back(r_names.size());
    uint32_t * cg = bam_get_cigar(b);
    r_cigar.insert(r_cigar.end(), cg, cg + c.n_cigar);
    r_cigar_off.push_back(r_cigar.size());
    r_bam_data.insert(r_bam_data.end(), b->data, b->data + b->l_data);
    r_bam_off.push_back(r_bam_data.size());
    r_lqname.push_back(c.l_qname);
    // AS-XS exactly as the reference computes it: run update_paths on a scratch pair
    std::pair<GenotypePaths, GenotypePaths> scratch =
      std::make_pair(GenotypePaths(c.flag, c.l_qseq), GenotypePaths(c.flag, c.l_qseq));
    update_paths(scratch, b);
    r_score_diff_first.push_back(scratch.first.score_diff);
  };

  auto process = [&](HtsRecord const & h, bool update_prev)
  {
    record_columns(h, !update_prev);
    if (update_prev && !light)
    {
      // golden per-read alignment (same call genotype_only makes internally)
      seqan::IupacString s1, s2;
      get_sequence(s1, s2, h.record);
      std::pair<GenotypePaths, GenotypePaths> gp = align_read(h.record, s1, s2, ph_index);
      pd.add(gp.first);
      pd.add(gp.second);
      if (dump_seeds && seqan::length(s1) >= 32)
      {
        sd0.add(query_index(s1, ph_index));
        sd1.add(query_index_hamming_distance1_without_index(s1, ph_index));
        sd0.add(query_index(s2, ph_index));
        sd1.add(query_index_hamming_distance1_without_index(s2, ph_index));
      }
      else if (dump_seeds)
      {
        for (int k = 0; k < 2; ++k)
        {
          sd0.nslots.push_back(0);
          sd1.nslots.push_back(0);
        }
      }
    }
    genotype_only(hts_preader, writer, reference_depth, maps, prev_paths, ph_index, nullptr, h, seq, rseq,
                  update_prev, IS_SV);
  };

  auto read_next = [&](HtsRecord & h)
  {
    bool const ok = hts_preader.read_record(h);
    if (ok)
      ++file_count[h.file_index];
    return ok;
  };
  bool is_done = !read_next(prev);
  while (!is_done && ((prev.record->core.flag & opts.sam_flag_filter) != 0u || (IS_SV && !is_good_read_sv(prev.record))))
    is_done = !read_next(prev);

  if (!is_done)
  {
    process(prev, true);
    have_prev = true;
    while (read_next(curr))
    {
      if ((curr.record->core.flag & opts.sam_flag_filter) != 0u || (IS_SV && !is_good_read_sv(curr.record)))
        continue;
      if (equal_pos_seq(prev.record, curr.record))
      {
        process(curr, false);
      }
      else
      {
        process(curr, true);
        hts_preader.move_record(prev, curr);
      }
    }
  }
  (void)have_prev;

  // leftover mates: SV calling only -- restated driver of hts_parallel_reader.cpp:719-772 on the reference's functions
  if (IS_SV)
  {
    auto process_leftovers = [&](long sample_i, long rg_i)
    {
      auto & map_gpaths = maps[rg_i];
      for (auto read_it = map_gpaths.begin(); read_it != map_gpaths.end(); ++read_it)
      {
        std::pair<GenotypePaths, GenotypePaths> copy(read_it->second);
        toggle_bit(copy.first.flags, IS_FIRST_IN_PAIR | IS_SEQ_REVERSED);
        toggle_bit(copy.second.flags, IS_FIRST_IN_PAIR | IS_SEQ_REVERSED);
        std::pair<GenotypePaths *, GenotypePaths *> better = get_better_paths(read_it->second, copy);
        if (better.first)
        {
          reference_depth.add_genotype_paths(*better.first, sample_i);
          writer.update_haplotype_scores_geno(*better.first, sample_i, nullptr);
        }
      }
    };
    for (long file_i = 0; file_i < static_cast<long>(hts_preader.hts_files.size()); ++file_i)
    {
      auto const & hts_f = hts_preader.hts_files[file_i];
      if (hts_f.rg2sample_i.size() <= 1)
        process_leftovers(hts_f.sample_index_offset, hts_f.rg_index_offset);
      else
        for (long rg_i = 0; rg_i < static_cast<long>(hts_f.rg2sample_i.size()); ++rg_i)
          process_leftovers(hts_f.sample_index_offset + hts_f.rg2sample_i[rg_i], hts_f.rg_index_offset + rg_i);
    }
    maps.clear();
  }

  if (light)
  {
    ArrayFile af;
    af.add("flag", r_flag);
    af.add("sample", r_sample, 1);
    af.add("file", r_file, 1);
    af.add("ord", r_ord, 1);
    af.add("name_id", r_pos, 1);
    af.add("isdup", r_isdup);
    af.add("score_diff", r_score_diff_first);
    std::vector<uint8_t> sample_names;
    for (auto const & s : writer.pns)
    {
      sample_names.insert(sample_names.end(), s.begin(), s.end());
      sample_names.push_back('\n');
    }
    af.add("sample_names", sample_names);
    af.write(out + ".stream.gtba");
  }
  else
  {
    ArrayFile af;
    af.add("flag", r_flag);
    af.add("pos", r_pos, 1);
    af.add("mpos", r_mpos, 1);
    af.add("isize", r_isize, 1);
    af.add("tid", r_tid, 1);
    af.add("mtid", r_mtid, 1);
    af.add("lseq", r_lseq, 1);
    af.add("sample", r_sample, 1);
    af.add("rg", r_rg, 1);
    af.add("file", r_file, 1);
    af.add("mapq", r_mapq);
    af.add("isdup", r_isdup);
    af.add("seq4", r_seq4);
    af.add("seq_off", r_seq_off);
    af.add("names", r_names);
    af.add("name_off", r_name_off);
    af.add("cigar", r_cigar);
    af.add("cigar_off", r_cigar_off);
    af.add("score_diff", r_score_diff_first);
    af.add("bam_data", r_bam_data);
    af.add("bam_off", r_bam_off);
    af.add("bam_lqname", r_lqname);
    pd.write(af, "");
    af.add("s0_nslots", sd0.nslots);
    af.add("s0_nlabels", sd0.nlabels);
    af.add("s0_labels", sd0.labels);
    af.add("s1_nslots", sd1.nslots);
    af.add("s1_nlabels", sd1.nlabels);
    af.add("s1_labels", sd1.labels);
    std::vector<uint8_t> sample_names;
    for (auto const & s : writer.pns)
    {
      sample_names.insert(sample_names.end(), s.begin(), s.end());
      sample_names.push_back('\n');
    }
    af.add("sample_names", sample_names);
    af.write(out + ".reads.gtba");
  }

  // ---- accumulators
  {
    ArrayFile af;
    long const NS = writer.pns.size();
    std::vector<uint32_t> hap_id, hap_num, hap_first_var, score_off(1, 0u), cov_off(1, 0u);
    std::vector<uint16_t> log_score, gt_cov, max_log_score;
    std::vector<uint8_t> amb, amb_alt, alt_pp;
    std::vector<uint64_t> vs_clipped_reads, vs_mapq_sq, pa_clipped_bp, pa_mapq_sq, pa_score_diff, pa_mismatches;
    std::vector<uint32_t> rs_counts; // per allele: r1f r1r r2f r2r

    for (auto const & hap : writer.haplotypes)
    {
      hap_id.push_back(hap.gt.id);
      hap_num.push_back(hap.gt.num);
      hap_first_var.push_back(hap.gt.first_variant_node);
      for (long s = 0; s < NS; ++s)
      {
        auto const & hs = hap.hap_samples[s];
        log_score.insert(log_score.end(), hs.log_score.begin(), hs.log_score.end());
        score_off.push_back(log_score.size());
        gt_cov.insert(gt_cov.end(), hs.gt_coverage.begin(), hs.gt_coverage.end());
        cov_off.push_back(gt_cov.size());
        max_log_score.push_back(hs.max_log_score);
        amb.push_back(hs.get_ambiguous_depth());
        amb_alt.push_back(hs.get_ambiguous_depth_alt());
        alt_pp.push_back(hs.get_alt_proper_pair_depth());
      }
      vs_clipped_reads.push_back(hap.var_stats.clipped_reads);
      vs_mapq_sq.push_back(hap.var_stats.mapq_squared);
      for (size_t a = 0; a < hap.var_stats.per_allele.size(); ++a)
      {
        auto const & pa = hap.var_stats.per_allele[a];
        pa_clipped_bp.push_back(pa.clipped_bp);
        pa_mapq_sq.push_back(pa.mapq_squared);
        pa_score_diff.push_back(pa.score_diff);
        pa_mismatches.push_back(pa.mismatches);
        auto const & rs = hap.var_stats.read_strand[a];
        rs_counts.push_back(rs.r1_forward);
        rs_counts.push_back(rs.r1_reverse);
        rs_counts.push_back(rs.r2_forward);
        rs_counts.push_back(rs.r2_reverse);
      }
    }
    std::vector<uint64_t> meta = {(uint64_t)NS, (uint64_t)writer.haplotypes.size()};
    af.add("meta", meta);
    af.add("hap_id", hap_id);
    af.add("hap_num", hap_num);
    af.add("hap_first_var", hap_first_var);
    af.add("score_off", score_off);
    af.add("log_score", log_score);
    af.add("cov_off", cov_off);
    af.add("gt_cov", gt_cov);
    af.add("max_log_score", max_log_score);
    af.add("amb", amb);
    af.add("amb_alt", amb_alt);
    af.add("alt_pp", alt_pp);
    af.add("vs_clipped_reads", vs_clipped_reads);
    af.add("vs_mapq_sq", vs_mapq_sq);
    af.add("pa_clipped_bp", pa_clipped_bp);
    af.add("pa_mapq_sq", pa_mapq_sq);
    af.add("pa_score_diff", pa_score_diff);
    af.add("pa_mismatches", pa_mismatches);
    af.add("rs_counts", rs_counts);

    if (IS_SV)
    {
      std::vector<uint16_t> depth;
      for (auto const & d : reference_depth.depths)
        depth.insert(depth.end(), d.begin(), d.end());
      af.add("ref_depth", depth);
      std::vector<uint64_t> rdm = {(uint64_t)reference_depth.reference_offset,
                                   (uint64_t)(reference_depth.depths.empty() ? 0 : reference_depth.depths[0].size())};
      af.add("ref_depth_meta", rdm);
    }

    // ---- phasing connections: hap_samples[s].connections[b1][hap2][b2] (vcf_writer.cpp:119-140,229-249), non-zero
    //      counts as sorted (sample, hap1, b1, hap2, b2) tuples
    {
      std::vector<std::array<uint32_t, 5>> keys;
      std::map<std::array<uint32_t, 5>, uint16_t> sorted;
      for (uint32_t h1 = 0; h1 < writer.haplotypes.size(); ++h1)
        for (long s = 0; s < NS; ++s)
        {
          auto const & conn_vec = writer.haplotypes[h1].hap_samples[s].connections;
          for (uint32_t b1 = 0; b1 < conn_vec.size(); ++b1)
            for (auto const & kv : conn_vec[b1])
              for (uint32_t b2 = 0; b2 < kv.second.size(); ++b2)
                if (kv.second[b2] != 0)
                  sorted[{(uint32_t)s, h1, b1, (uint32_t)kv.first, b2}] = kv.second[b2];
        }
      std::vector<uint32_t> conn_tuples;
      std::vector<uint16_t> conn_counts;
      for (auto const & kv : sorted)
      {
        conn_tuples.insert(conn_tuples.end(), kv.first.begin(), kv.first.end());
        conn_counts.push_back(kv.second);
      }
      af.add("conn_tuples", conn_tuples);
      af.add("conn_counts", conn_counts);
    }

    // ---- the phase-support map `ph` exactly as the reference derives it: a second, independent pass that calls the
    //      reference's own pool function with is_writing_hap = true (hts_parallel_reader.cpp:458,782-893)
    if (!IS_SV)
    {
      std::vector<std::map<std::pair<uint16_t, uint16_t>, std::map<std::pair<uint16_t, uint16_t>, int8_t>>> ph_vec(1);
      std::string out_path, tmp_dir = out + ".phtmp", reference_fn = "", reg = ".";
      std::vector<double> avg_cov(sams.size(), -1.0);
      std::string const cmd = "mkdir -p '" + tmp_dir + "'";
      if (system(cmd.c_str()) != 0)
        return 3;
      parallel_reader_genotype_only(0, &out_path, &sams, &avg_cov, &tmp_dir, &reference_fn, &reg, &ph_index, nullptr, &ph_vec,
                                    true, true, nullptr);
      std::string const rm = "rm -rf '" + tmp_dir + "'";
      if (system(rm.c_str()) != 0)
        return 3;
      std::vector<uint32_t> ph_tuples; // ps1 cov1 ps2 cov2 flags
      for (auto const & a : ph_vec[0])
        for (auto const & b : a.second)
          ph_tuples.insert(ph_tuples.end(), {(uint32_t)a.first.first, (uint32_t)a.first.second, (uint32_t)b.first.first,
                                             (uint32_t)b.first.second, (uint32_t)(uint8_t)b.second});
      af.add("ph_tuples", ph_tuples);
    }

    // ---- pool finalisation: SampleCall + VarStats (vcf.cpp:1507, variant.cpp:230)
    Vcf vcf;
    vcf.sample_names = writer.pns;
    for (long ps = 0; ps < static_cast<long>(writer.haplotypes.size()); ++ps)
      vcf.add_haplotype(writer.haplotypes[ps], static_cast<int32_t>(ps));
    if (!IS_SV)
      for (Variant & var : vcf.variants)
        var.scan_calls();

    std::vector<uint8_t> phred, c_amb, c_altpp;
    std::vector<uint16_t> c_cov, c_ref_total, c_alt_total, gt_call;
    std::vector<uint8_t> gq;
    std::vector<uint64_t> st_u64; // per variant: n_genotyped n_calls n_passed_calls n_max_alt_pp seqdepth het_ad0 het_ad1 hom_ad0 hom_ad1
    std::vector<uint64_t> sta_u64; // per allele: qd_qual qd_depth total_depth ac pass_ac n_ref_ref n_ref_alt n_alt_alt max_alt_support het_multi0 het_multi1 hom_multi0 hom_multi1
    std::vector<double> sta_ratio;
    for (auto const & var : vcf.variants)
    {
      for (auto const & call : var.calls)
      {
        phred.insert(phred.end(), call.phred.begin(), call.phred.end());
        c_cov.insert(c_cov.end(), call.coverage.begin(), call.coverage.end());
        c_ref_total.push_back(call.ref_total_depth);
        c_alt_total.push_back(call.alt_total_depth);
        c_amb.push_back(call.ambiguous_depth);
        c_altpp.push_back(call.alt_proper_pair_depth);
        auto gt = call.get_gt_call();
        gt_call.push_back(gt.first);
        gt_call.push_back(gt.second);
        gq.push_back(call.get_gq());
      }
      auto const & st = var.stats;
      st_u64.insert(st_u64.end(),
                    {st.n_genotyped, st.n_calls, st.n_passed_calls, st.n_max_alt_proper_pairs, st.seqdepth,
                     st.het_allele_depth.first, st.het_allele_depth.second, st.hom_allele_depth.first,
                     st.hom_allele_depth.second});
      for (auto const & pa : st.per_allele)
      {
        sta_u64.insert(sta_u64.end(),
                       {pa.qd_qual, pa.qd_depth, pa.total_depth, pa.ac, pa.pass_ac, pa.n_ref_ref, pa.n_ref_alt,
                        pa.n_alt_alt, pa.maximum_alt_support, pa.het_multi_allele_depth.first,
                        pa.het_multi_allele_depth.second, pa.hom_multi_allele_depth.first,
                        pa.hom_multi_allele_depth.second});
        sta_ratio.push_back(pa.maximum_alt_support_ratio);
      }
    }
    af.add("call_phred", phred);
    af.add("call_cov", c_cov);
    af.add("call_ref_total", c_ref_total);
    af.add("call_alt_total", c_alt_total);
    af.add("call_amb", c_amb);
    af.add("call_altpp", c_altpp);
    af.add("call_gt", gt_call);
    af.add("call_gq", gq);
    af.add("stats_var", st_u64);
    af.add("stats_allele", sta_u64);
    af.add("stats_allele_ratio", sta_ratio, 2);
    af.write(out + ".accum.gtba");
  }

  return 0;
}
