// gtb_shim.hpp -- the reference-side binding of libgtb200: what a graphtyper maintainer adds to the tree.
//
// Header-only C++17 on top of the reference's own headers and include/gtb200.h:
//   flatten()       gyper::Graph            -> gtb_graph_view    (input of gtb_region_begin; replaces index_graph's argument)
//   Records         HtsRecord stream        -> gtb_bam_batch     (input of gtb_submit_bam_records; replaces genotype_only's
//                                                                 per-record work, src/utilities/hts_parallel_reader.cpp:245-338)
//   is_good_read()  the SV pre-filter of the pool loop           (hts_parallel_reader.cpp:528-568)
//   CoverageCap     the per-sample 50-bp bin cap of the pool loop (hts_parallel_reader.cpp:599-633)
//   Accumulators    caller-owned buffers of gtb_pool_finish
//   fill_writer()   gtb_accumulators (+ connections) -> VcfWriter::haplotypes[*].hap_samples[*] / var_stats, ReferenceDepth:
//                   the state the reference's own pool finalisation (Vcf::add_haplotype, src/typer/vcf.cpp:1507) reads
//   die()           gtb_status              -> print_log(error) + std::exit(1), the reference's error convention
// integration/gtb_pool_reader.cpp puts them together into replacements of index_graph and parallel_reader_genotype_only.
//
// Compiled and exercised against the unmodified reference objects by oracle/ref_build/shim_probe.cpp
// (tests/test_shim_against_reference.py).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <graphtyper/graph/graph.hpp>
#include <graphtyper/graph/reference_depth.hpp>
#include <graphtyper/typer/vcf_writer.hpp>
#include <graphtyper/utilities/hts_parallel_reader.hpp>
#include <graphtyper/utilities/logging.hpp>

#include <htslib/bgzf.h>
#include <htslib/hts.h>
#include <htslib/sam.h>

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

// The pool loop's read filter when calling SVs (src/utilities/hts_parallel_reader.cpp:528-568): mapped; not (MAPQ <= 15 with
// the mate on another contig or more than 200 kb away); not soft-clipped at both ends; not (MAPQ <= 15 and >= 12 bases
// soft-clipped at one end).
inline bool is_good_read(bam1_t const * r)
{
  auto const & c = r->core;
  if ((c.flag & BAM_FUNMAP) != 0)
    return false;
  bool const mate_far = c.tid != c.mtid || std::abs((long)(c.pos - c.mpos)) > 200000l;
  if (c.qual <= 15 && mate_far)
    return false;
  if (c.n_cigar >= 2)
  {
    uint32_t const * cig = bam_get_cigar(r);
    uint32_t const front = cig[0], back = cig[c.n_cigar - 1];
    bool const front_clip = bam_cigar_op(front) == BAM_CSOFT_CLIP, back_clip = bam_cigar_op(back) == BAM_CSOFT_CLIP;
    bool const one_long = (front_clip && bam_cigar_oplen(front) >= 12) || (back_clip && bam_cigar_oplen(back) >= 12);
    if ((front_clip && back_clip) || (c.qual <= 15 && one_long))
      return false;
  }
  return true;
}

// The (extreme) coverage filter of the pool loop when calling SVs (hts_parallel_reader.cpp:587-633): per sample, reads are
// counted in 50-bp bins from the first record's position; a bin that already holds more than avg_cov_by_readlen * 150 reads
// takes no more.  A record that opens a new bin is always kept; duplicates of the previous record are counted but never
// skipped (the caller decides that, exactly like the reference's loop: :666-707).
struct CoverageCap
{
  bool active = false;
  long first_pos = 0;
  std::vector<double> const * avg_cov = nullptr;
  std::vector<std::vector<uint16_t>> bins;

  void start(bool is_active, long n_samples, long first_record_pos, std::vector<double> const * avg_cov_by_readlen)
  {
    active = is_active;
    first_pos = first_record_pos;
    avg_cov = avg_cov_by_readlen;
    if (active)
      bins.assign(n_samples, {});
  }

  // counts the record in; false = its bin is full
  bool admit(long sample_i, long pos)
  {
    if (!active || !avg_cov || sample_i >= (long)avg_cov->size() || (*avg_cov)[sample_i] <= 0.0)
      return true;
    uint16_t const cap = (uint16_t)std::min(65535l, (long)((*avg_cov)[sample_i] * 50.0 * 3.0 + 0.5));
    auto & b = bins[sample_i];
    long const bin = (pos - first_pos) / 50l;
    if (bin >= (long)b.size())
    {
      b.resize(bin + 1, 0u);
      ++b[bin];
      return true;
    }
    if (b[bin] > cap)
      return false;
    ++b[bin];
    return true;
  }
};

// The pool's files as COMPRESSED bytes for gtb_submit_bgzf: for every BAM file the chunks of its region iterator
// (hts_itr_t::off, which HtsReader::open has already computed from the index: src/utilities/hts_reader.cpp:99-118), adjacent
// chunks merged -- or, with region "." (how `genotype` reads its pools' bamshrink output), everything behind the header --
// each read from the file as it is: no inflate, no record parsing, no heap on the host.  collect() says why
// when the pool has to stay with the reference's reader: CRAM / SAM, several read groups in one file
// (sample and read group are per file on the device), files that number the contig differently.
struct BgzfPool
{
  std::vector<std::vector<uint8_t>> bytes; // one buffer per segment
  std::vector<std::vector<gtb_bgzf_segment>> segs;
  std::vector<gtb_bgzf_file> files;
  gtb_bgzf_query query{};
  uint64_t n_bytes = 0;
  std::vector<uint64_t> whole_u; // whole-file reading: virtual offset of every file's first record
  // the device needs ~4 bytes of working memory per inflated byte (record lists, sort buffers): pools beyond this stay with htslib
  static constexpr uint64_t max_pool_bytes = 256ull << 20;

  bool collect(gyper::HtsParallelReader const & reader, uint32_t flag_filter, bool is_sv, std::string & why_not)
  {
    size_t const nf = reader.hts_files.size();
    segs.assign(nf, {});
    files.assign(nf, gtb_bgzf_file{});
    for (size_t i = 0; i < nf; ++i)
    {
      gyper::HtsReader const & f = reader.hts_files[i];
      if (!f.fp || f.fp->format.format != bam || f.fp->format.compression != bgzf)
        return why_not = "not a BAM file", false;
      hts_itr_t const * it = f.hts_iter;
      bool const whole = it == nullptr; // region ".": HtsReader reads the file with sam_read1 (hts_reader.cpp:94-97)
      if (i == 0)
        query.whole_file = whole ? 1u : 0u;
      else if ((query.whole_file != 0) != whole)
        return why_not = "files disagree on region / whole-file reading", false;
      if (!whole && (it->read_rest || it->nocoor || it->multi || it->tid < 0))
        return why_not = "no single-region iterator", false;
      if (f.rg2sample_i.size() > 1)
        return why_not = "several read groups in one file", false;
      if (whole)
      {
        // where the records start: behind the header, found by reading the header once more on a handle of our own
        htsFile * hf = hts_open(f.fp->fn, "r");
        if (!hf)
          return why_not = "cannot reopen the file", false;
        sam_hdr_t * hh = sam_hdr_read(hf);
        uint64_t const u = hh ? (uint64_t)bgzf_tell(hf->fp.bgzf) : 0;
        if (hh)
          sam_hdr_destroy(hh);
        hts_close(hf);
        if (!hh)
          return why_not = "cannot read the header", false;
        struct stat st0;
        if (::stat(f.fp->fn, &st0) != 0 || (uint64_t)st0.st_size - (u >> 16) > max_pool_bytes)
          return why_not = "more than 256 MiB to read without a region", false;
        whole_u.push_back(u);
      }
      else if (i == 0)
      {
        query.tid = it->tid;
        query.beg = it->beg;
        query.end = it->end;
      }
      else if (query.tid != it->tid || query.beg != it->beg || query.end != it->end)
        return why_not = "files disagree on the region's contig index", false;
      int const fd = ::open(f.fp->fn, O_RDONLY);
      if (fd < 0)
        return why_not = "cannot open the file for raw reads", false;
      struct stat st;
      if (::fstat(fd, &st) != 0)
      {
        ::close(fd);
        return why_not = "cannot stat the file", false;
      }
      uint64_t const file_size = (uint64_t)st.st_size;
      int const n_chunks = whole ? 1 : it->n_off;
      for (int k = 0; k < n_chunks; ++k)
      {
        uint64_t const u = whole ? whole_u.back() : it->off[k].u;
        uint64_t v = whole ? (file_size << 16) : it->off[k].v;
        while (!whole && k + 1 < it->n_off && it->off[k + 1].u == v) // adjacent chunks: hts_itr_next reads on without a seek
          v = it->off[++k].v;
        uint64_t const begin = u >> 16;
        // the last record that is read starts before v and may reach into the blocks behind v's: three more block sizes
        uint64_t const end = whole ? file_size : std::min<uint64_t>(file_size, (v >> 16) + 4 * 65536ull);
        if (begin >= end)
          continue;
        std::vector<uint8_t> buf(end - begin);
        uint64_t got = 0;
        while (got < buf.size())
        {
          ssize_t const r = ::pread(fd, buf.data() + got, buf.size() - got, (off_t)(begin + got));
          if (r <= 0)
            break;
          got += (uint64_t)r;
        }
        if (got != buf.size())
        {
          ::close(fd);
          return why_not = "short read", false;
        }
        gtb_bgzf_segment g{};
        g.comp_bytes = buf.size();
        g.file_offset = begin;
        g.v_end = v;
        g.first_offset = (uint32_t)(u & 0xFFFFu);
        g.to_eof = end == file_size ? 1u : 0u;
        n_bytes += buf.size();
        if (n_bytes > max_pool_bytes)
        {
          ::close(fd);
          return why_not = "more than 256 MiB of compressed bytes in one pool", false;
        }
        bytes.push_back(std::move(buf));
        segs[i].push_back(g);
      }
      ::close(fd);
      files[i].sample = f.sample_index_offset;
      files[i].rg = f.rg_index_offset;
    }
    // pointers last: the vectors above no longer move
    size_t b = 0;
    for (size_t i = 0; i < nf; ++i)
    {
      for (auto & g : segs[i])
        g.comp = bytes[b++].data();
      files[i].n_segments = (uint32_t)segs[i].size();
      files[i].segments = segs[i].data();
    }
    query.flag_filter = flag_filter;
    query.sv_read_filter = is_sv ? 1u : 0u;
    query.check_crc = 1u;
    return true;
  }
};

// Caller-owned buffers behind a gtb_accumulators (sized with gtb_accumulator_sizes / gtb_ref_depth_size).
struct Accumulators
{
  std::vector<uint32_t> bubble_id, n_alleles, saturated, read_strand;
  std::vector<uint64_t> score_off, cov_off, vs_clipped_reads, vs_mapq_squared, pa_clipped_bp, pa_mapq_squared, pa_score_diff,
    pa_mismatches;
  std::vector<uint16_t> log_score, gt_coverage, max_log_score, ref_depth;
  std::vector<uint8_t> amb, amb_alt, alt_pp;
  gtb_accumulators view{};

  void resize(uint32_t nb, uint64_t n_scores, uint64_t n_cov, uint32_t ns, uint32_t depth_size)
  {
    size_t const cells = (size_t)nb * ns;
    bubble_id.assign(nb, 0);
    n_alleles.assign(nb, 0);
    score_off.assign(nb + 1, 0);
    cov_off.assign(nb + 1, 0);
    log_score.assign(n_scores * ns, 0);
    gt_coverage.assign(n_cov * ns, 0);
    max_log_score.assign(cells, 0);
    amb.assign(cells, 0);
    amb_alt.assign(cells, 0);
    alt_pp.assign(cells, 0);
    saturated.assign(cells, 0);
    vs_clipped_reads.assign(nb, 0);
    vs_mapq_squared.assign(nb, 0);
    pa_clipped_bp.assign(n_cov, 0);
    pa_mapq_squared.assign(n_cov, 0);
    pa_score_diff.assign(n_cov, 0);
    pa_mismatches.assign(n_cov, 0);
    read_strand.assign(n_cov * 4, 0);
    ref_depth.assign((size_t)depth_size * ns, 0);
    view = gtb_accumulators{};
    view.n_bubbles = nb;
    view.n_samples = ns;
    view.bubble_id = bubble_id.data();
    view.n_alleles = n_alleles.data();
    view.score_off = score_off.data();
    view.cov_off = cov_off.data();
    view.log_score = log_score.data();
    view.gt_coverage = gt_coverage.data();
    view.max_log_score = max_log_score.data();
    view.ambiguous_depth = amb.data();
    view.ambiguous_depth_alt = amb_alt.data();
    view.alt_proper_pair_depth = alt_pp.data();
    view.saturated = saturated.data();
    view.vs_clipped_reads = vs_clipped_reads.data();
    view.vs_mapq_squared = vs_mapq_squared.data();
    view.pa_clipped_bp = pa_clipped_bp.data();
    view.pa_mapq_squared = pa_mapq_squared.data();
    view.pa_score_diff = pa_score_diff.data();
    view.pa_mismatches = pa_mismatches.data();
    view.read_strand = read_strand.data();
    view.depth_size = depth_size;
    view.ref_depth = depth_size ? ref_depth.data() : nullptr;
  }
};

// The pool's results into the objects the reference's pool finalisation reads (hts_parallel_reader.cpp:782-1029):
// VcfWriter::haplotypes[b].hap_samples[s] (include/graphtyper/graph/haplotype.hpp:25-75), Haplotype::var_stats
// (include/graphtyper/typer/var_stats.hpp:15-84), and for SV graphs the per-sample depth tracks of ReferenceDepth.
// `writer` comes straight from set_samples (all zero); bubble b of the accumulators is haplotype b (both follow the graph's
// bubble order, checked through Genotype::id).
inline void fill_writer(gyper::VcfWriter & writer, gyper::ReferenceDepth & reference_depth, gtb_accumulators const & A,
                        gtb_connection const * conn, uint64_t n_conn)
{
  using namespace gyper;
  size_t const NS = A.n_samples;
  if (writer.haplotypes.size() != A.n_bubbles)
  {
    print_log(log_severity::error, "[gtb200] ", writer.haplotypes.size(), " haplotypes in the writer but ", A.n_bubbles,
              " bubbles in the accumulators");
    std::exit(1);
  }
  for (size_t b = 0; b < A.n_bubbles; ++b)
  {
    Haplotype & hap = writer.haplotypes[b];
    size_t const cnum = A.n_alleles[b], tri = cnum * (cnum + 1) / 2;
    if (hap.gt.id != A.bubble_id[b] || hap.gt.num != cnum || hap.hap_samples.size() != NS)
    {
      print_log(log_severity::error, "[gtb200] haplotype ", b, " does not match bubble ", A.bubble_id[b]);
      std::exit(1);
    }
    for (size_t s = 0; s < NS; ++s)
    {
      HapSample & hs = hap.hap_samples[s];
      uint16_t const * ls = A.log_score + A.score_off[b] * NS + s * tri;
      uint16_t const * gc = A.gt_coverage + A.cov_off[b] * NS + s * cnum;
      hs.log_score.assign(ls, ls + tri);
      hs.gt_coverage.assign(gc, gc + cnum);
      hs.max_log_score = A.max_log_score[b * NS + s];
      // the three depths are private saturating uint8 counters: counted up to their (already clamped) values
      for (unsigned k = 0; k < A.ambiguous_depth[b * NS + s]; ++k)
        hs.increment_ambiguous_depth();
      for (unsigned k = 0; k < A.ambiguous_depth_alt[b * NS + s]; ++k)
        hs.increment_ambiguous_depth_alt();
      for (unsigned k = 0; k < A.alt_proper_pair_depth[b * NS + s]; ++k)
        hs.increment_alt_proper_pair_depth();
    }
    VarStats & vs = hap.var_stats;
    vs.clipped_reads = (uint32_t)A.vs_clipped_reads[b];
    vs.mapq_squared = A.vs_mapq_squared[b];
    for (size_t a = 0; a < cnum && a < vs.per_allele.size(); ++a)
    {
      size_t const i = A.cov_off[b] + a;
      vs.per_allele[a].clipped_bp = A.pa_clipped_bp[i];
      vs.per_allele[a].mapq_squared = A.pa_mapq_squared[i];
      vs.per_allele[a].score_diff = (uint32_t)A.pa_score_diff[i];
      vs.per_allele[a].mismatches = (uint32_t)A.pa_mismatches[i];
      vs.read_strand[a].r1_forward = A.read_strand[i * 4];
      vs.read_strand[a].r1_reverse = A.read_strand[i * 4 + 1];
      vs.read_strand[a].r2_forward = A.read_strand[i * 4 + 2];
      vs.read_strand[a].r2_reverse = A.read_strand[i * 4 + 3];
    }
  }
  // HapSample::connections[allele1][hap2] = support per allele of hap2 (haplotype.hpp:42)
  for (uint64_t k = 0; k < n_conn; ++k)
  {
    gtb_connection const & c = conn[k];
    auto & m = writer.haplotypes[c.hap1].hap_samples[c.sample].connections[c.allele1];
    auto & v = m[c.hap2];
    if (v.empty())
      v.assign(writer.haplotypes[c.hap2].gt.num, 0u);
    v[c.allele2] = (uint16_t)c.count;
  }
  if (A.ref_depth && A.depth_size)
  {
    reference_depth.reference_offset = A.reference_offset;
    for (size_t s = 0; s < NS && s < reference_depth.depths.size(); ++s)
    {
      auto & d = reference_depth.depths[s];
      size_t const n = std::min<size_t>(d.size(), A.depth_size);
      std::copy(A.ref_depth + s * A.depth_size, A.ref_depth + s * A.depth_size + n, d.begin());
    }
  }
}
} // namespace gtb_shim
