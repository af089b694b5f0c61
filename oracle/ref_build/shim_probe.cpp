// oracle/ref_build/shim_probe.cpp -- TEST INFRASTRUCTURE: the reference-side shim (integration/gtb_shim.hpp) compiled against
// the UNMODIFIED reference objects and libgtb200.so, exercised where no GPU is needed:
//   1. construct_graph (reference) -> gtb_shim::flatten -> gtb_region_begin on a host-only context (host index builder)
//      -> gtb_index_export, compared key by key and label by label (bucket order included) with index_graph (reference);
//   2. the pool's records gathered by gtb_shim::Records through the reference's own HtsParallelReader + flag filter;
//   3. the compute entry points refuse to run without a device (no CPU fallback);
//   4. with --gpu (on the GPU box): region + pool + records through the C ABI on the device, every HapSample and per-bubble
//      statistic compared with the reference's own pool loop (genotype_only + VcfWriter) run in the same process.
//   5. with --bgzf REGION a.bam[,b.bam] [--sv]: gtb_shim::BgzfPool::collect on the pool's files as the reference's own reader
//      opened them (the chunks of the real index, or everything behind the header for region "."), decoded by the CPU run of
//      the device's source functions (gtb_debug_bgzf_host), compared record by record with what HtsParallelReader::read_record
//      + the pool loop's filters hand out.  No GPU needed.
// Prints "SHIM PASS ..." (and "SHIM GPU PASS ...", "SHIM BGZF PASS ...") or the first difference.
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <graphtyper/constants.hpp>
#include <graphtyper/graph/absolute_position.hpp>
#include <graphtyper/graph/constructor.hpp>
#include <graphtyper/graph/genomic_region.hpp>
#include <graphtyper/graph/graph.hpp>
#include <graphtyper/index/indexer.hpp>
#include <graphtyper/index/ph_index.hpp>
#include <graphtyper/utilities/hts_parallel_reader.hpp>
#include <graphtyper/utilities/logging.hpp>
#include <graphtyper/utilities/options.hpp>

#include <graphtyper/graph/reference_depth.hpp>
#include <graphtyper/typer/genotype_paths.hpp>
#include <graphtyper/typer/primers.hpp>
#include <graphtyper/typer/vcf_writer.hpp>
#include <graphtyper/utilities/hts_utils.hpp>

#include <seqan/sequence.h>

#include "../../integration/gtb_shim.hpp"

namespace gyper
{
// defined with external linkage in src/utilities/hts_parallel_reader.cpp:245 (not declared in a header)
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
// 4. (--gpu) the whole path through the shim on the device against the reference's own pool loop, in this process:
// every HapSample (log_score, gt_coverage, max_log_score, ambiguous / alt proper-pair depths) and the per-bubble statistics
int compare_with_reference_loop(std::vector<std::string> const & sams, gyper::PHIndex const & ph_index, gtb_shim::FlatGraph const & fg,
                                gtb_bam_batch const & batch)
{
  using namespace gyper;
  Options const & opts = *Options::const_instance();
  // reference: driver of hts_parallel_reader.cpp:655-708 on the reference's own functions (same as gt_probe.cpp)
  HtsParallelReader reader;
  reader.open(sams, "", ".");
  VcfWriter writer(opts.split_var_threshold - 1);
  writer.set_samples(reader.get_samples());
  ReferenceDepth reference_depth;
  std::vector<std::unordered_map<std::string, std::pair<GenotypePaths, GenotypePaths>>> maps(reader.get_num_rg());
  std::pair<GenotypePaths, GenotypePaths> prev_paths;
  HtsRecord prev, curr;
  seqan::IupacString seq, rseq;
  bool done = !reader.read_record(prev);
  while (!done && (prev.record->core.flag & opts.sam_flag_filter) != 0u)
    done = !reader.read_record(prev);
  if (!done)
  {
    genotype_only(reader, writer, reference_depth, maps, prev_paths, ph_index, nullptr, prev, seq, rseq, true, false);
    while (reader.read_record(curr))
    {
      if ((curr.record->core.flag & opts.sam_flag_filter) != 0u)
        continue;
      if (equal_pos_seq(prev.record, curr.record))
        genotype_only(reader, writer, reference_depth, maps, prev_paths, ph_index, nullptr, curr, seq, rseq, false, false);
      else
      {
        genotype_only(reader, writer, reference_depth, maps, prev_paths, ph_index, nullptr, curr, seq, rseq, true, false);
        reader.move_record(prev, curr);
      }
    }
  }
  // device: region + pool + records through the C ABI
  int const NS = (int)writer.pns.size();
  gtb_ctx * ctx = nullptr;
  gtb_shim::die(gtb_create(0, &ctx));
  gtb_shim::die(gtb_region_begin(ctx, 0, &fg.view));
  gtb_shim::die(gtb_pool_begin(ctx, 0, NS));
  gtb_submit_stats st{};
  gtb_shim::die(gtb_submit_bam_records(ctx, 0, &batch, &st));
  uint32_t nb = 0;
  uint64_t n_scores = 0, n_cov = 0;
  gtb_shim::die(gtb_accumulator_sizes(ctx, 0, &nb, &n_scores, &n_cov));
  std::vector<uint32_t> bubble_id(nb), n_alleles(nb), saturated((size_t)nb * NS), read_strand(n_cov * 4);
  std::vector<uint64_t> score_off(nb + 1), cov_off(nb + 1), vs_cr(nb), vs_mq(nb), pa_cb(n_cov), pa_mq(n_cov), pa_sd(n_cov), pa_mm(n_cov);
  std::vector<uint16_t> log_score(n_scores * NS), gt_cov(n_cov * NS), max_ls((size_t)nb * NS);
  std::vector<uint8_t> amb((size_t)nb * NS), amb_alt((size_t)nb * NS), alt_pp((size_t)nb * NS);
  gtb_accumulators acc{};
  acc.bubble_id = bubble_id.data();
  acc.n_alleles = n_alleles.data();
  acc.score_off = score_off.data();
  acc.cov_off = cov_off.data();
  acc.log_score = log_score.data();
  acc.gt_coverage = gt_cov.data();
  acc.max_log_score = max_ls.data();
  acc.ambiguous_depth = amb.data();
  acc.ambiguous_depth_alt = amb_alt.data();
  acc.alt_proper_pair_depth = alt_pp.data();
  acc.saturated = saturated.data();
  acc.vs_clipped_reads = vs_cr.data();
  acc.vs_mapq_squared = vs_mq.data();
  acc.pa_clipped_bp = pa_cb.data();
  acc.pa_mapq_squared = pa_mq.data();
  acc.pa_score_diff = pa_sd.data();
  acc.pa_mismatches = pa_mm.data();
  acc.read_strand = read_strand.data();
  gtb_shim::die(gtb_pool_finish(ctx, 0, &acc));
  gtb_destroy(ctx);
  if (writer.haplotypes.size() != nb)
  {
    printf("SHIM GPU FAIL: %zu haplotypes in the reference, %u bubbles on the device\n", writer.haplotypes.size(), nb);
    return 1;
  }
  for (uint32_t b = 0; b < nb; ++b)
  {
    Haplotype const & hap = writer.haplotypes[b];
    uint64_t const tri = score_off[b + 1] - score_off[b], cnum = cov_off[b + 1] - cov_off[b];
    bool ok = hap.gt.id == bubble_id[b] && hap.gt.num == n_alleles[b] && hap.var_stats.clipped_reads == vs_cr[b] &&
              hap.var_stats.mapq_squared == vs_mq[b];
    for (int s = 0; ok && s < NS; ++s)
    {
      auto const & hs = hap.hap_samples[s];
      ok = hs.log_score.size() == tri && hs.gt_coverage.size() == cnum &&
           memcmp(hs.log_score.data(), log_score.data() + score_off[b] * NS + s * tri, tri * 2) == 0 &&
           memcmp(hs.gt_coverage.data(), gt_cov.data() + cov_off[b] * NS + s * cnum, cnum * 2) == 0 &&
           hs.max_log_score == max_ls[(size_t)b * NS + s] && hs.get_ambiguous_depth() == amb[(size_t)b * NS + s] &&
           hs.get_ambiguous_depth_alt() == amb_alt[(size_t)b * NS + s] &&
           hs.get_alt_proper_pair_depth() == alt_pp[(size_t)b * NS + s];
    }
    for (uint64_t a = 0; ok && a < cnum; ++a)
    {
      auto const & pa = hap.var_stats.per_allele[a];
      auto const & rs = hap.var_stats.read_strand[a];
      uint64_t const i = cov_off[b] + a;
      ok = pa.clipped_bp == pa_cb[i] && pa.mapq_squared == pa_mq[i] && pa.score_diff == pa_sd[i] && pa.mismatches == pa_mm[i] &&
           rs.r1_forward == read_strand[i * 4] && rs.r1_reverse == read_strand[i * 4 + 1] && rs.r2_forward == read_strand[i * 4 + 2] &&
           rs.r2_reverse == read_strand[i * 4 + 3];
    }
    if (!ok)
    {
      printf("SHIM GPU FAIL: bubble %u differs from the reference's haplotype\n", b);
      return 1;
    }
  }
  printf("SHIM GPU PASS bubbles=%u samples=%d records=%llu pairs_scored=%llu\n", nb, NS, (unsigned long long)st.n_records,
         (unsigned long long)st.n_pairs_scored);
  return 0;
}
} // namespace

namespace
{
std::vector<std::string> split_list(std::string const & s)
{
  std::vector<std::string> out;
  size_t b = 0;
  while (b <= s.size())
  {
    size_t e = s.find(',', b);
    if (e == std::string::npos)
      e = s.size();
    if (e > b)
      out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
  return out;
}

// 5. (--bgzf) the compressed-bytes hand-off against the reference's reader on the same files and region
int bgzf_check(std::string const & region, std::vector<std::string> const & paths, bool is_sv)
{
  using namespace gyper;
  Options const & opts = *Options::const_instance();
  HtsParallelReader a;
  a.open(paths, "", region);
  gtb_shim::BgzfPool pool;
  std::string why;
  if (!pool.collect(a, (uint32_t)opts.sam_flag_filter, is_sv, why))
  {
    printf("SHIM BGZF DECLINED: %s\n", why.c_str());
    return 0;
  }
  uint32_t n = 0;
  uint64_t nd = 0;
  if (gtb_debug_bgzf_host((int)pool.files.size(), pool.files.data(), &pool.query, &n, &nd, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr) != 0)
  {
    printf("SHIM BGZF FAIL: %s\n", gtb_last_error());
    return 1;
  }
  std::vector<gtb_bam_core> core(n);
  std::vector<uint8_t> data(nd + 1);
  std::vector<uint64_t> off(n + 1);
  std::vector<int32_t> sample(n), rg(n);
  if (n && gtb_debug_bgzf_host((int)pool.files.size(), pool.files.data(), &pool.query, &n, &nd, core.data(), data.data(), off.data(),
                               sample.data(), rg.data(), nullptr, nullptr) != 0)
  {
    printf("SHIM BGZF FAIL: %s\n", gtb_last_error());
    return 1;
  }
  // the reference's reader on the same files: records in its order, the pool loop's filters applied
  HtsParallelReader b;
  b.open(paths, "", region);
  HtsRecord rec;
  uint64_t k = 0;
  while (b.read_record(rec))
  {
    bam1_t const * r = rec.record;
    if ((r->core.flag & opts.sam_flag_filter) != 0u || (is_sv && !gtb_shim::is_good_read(r)))
      continue;
    if (k >= n)
    {
      printf("SHIM BGZF FAIL: the reference reads more than %u records\n", n);
      return 1;
    }
    long s_i = 0, rg_i = 0;
    b.get_sample_and_rg_index(s_i, rg_i, rec);
    gtb_bam_core const & c = core[k];
    const uint8_t * d = data.data() + off[k];
    uint64_t const len = off[k + 1] - off[k];
    size_t const name_len = strlen(bam_get_qname(r)) + 1; // the file's l_read_name; htslib pads the name in memory
    bool ok = c.pos == r->core.pos && c.mpos == r->core.mpos && c.isize == r->core.isize && c.tid == r->core.tid && c.mtid == r->core.mtid &&
              c.l_qseq == r->core.l_qseq && c.n_cigar == r->core.n_cigar && c.flag == r->core.flag && c.mapq == r->core.qual &&
              c.l_qname == name_len && sample[k] == s_i && rg[k] == rg_i;
    uint64_t const rest = (uint64_t)r->l_data - r->core.l_qname;
    ok = ok && len == name_len + rest && memcmp(d, r->data, name_len) == 0 && memcmp(d + name_len, r->data + r->core.l_qname, rest) == 0;
    if (!ok)
    {
      printf("SHIM BGZF FAIL: record %llu (%s at %lld) differs\n", (unsigned long long)k, bam_get_qname(r), (long long)r->core.pos);
      return 1;
    }
    ++k;
  }
  if (k != n)
  {
    printf("SHIM BGZF FAIL: %u records decoded, the reference reads %llu\n", n, (unsigned long long)k);
    return 1;
  }
  size_t n_seg = 0;
  for (auto const & f : pool.files)
    n_seg += f.n_segments;
  printf("SHIM BGZF PASS records=%u files=%zu segments=%zu compressed_bytes=%llu whole_file=%u\n", n, pool.files.size(), n_seg,
         (unsigned long long)pool.n_bytes, pool.query.whole_file);
  return 0;
}
} // namespace

int main(int argc, char ** argv)
{
  using namespace gyper;
  if (argc >= 4 && std::string(argv[1]) == "--bgzf")
  {
    Options::instance();
    gyper::log_singleton = std::unique_ptr<gyper::log_singleton_t>{new gyper::log_singleton_t{gyper::log_severity::warning, std::clog}};
    return bgzf_check(argv[2], split_list(argv[3]), argc > 4 && std::string(argv[4]) == "--sv");
  }
  if (argc < 5)
  {
    fprintf(stderr, "usage: shim_probe REF.fa VCF.gz chr:b-e a.sam[,b.sam]\n");
    return 2;
  }
  std::string const ref_fn = argv[1], vcf_fn = argv[2], region_str = argv[3];
  std::vector<std::string> sams;
  {
    std::string s = argv[4];
    size_t b = 0;
    while (b <= s.size())
    {
      size_t e = s.find(',', b);
      if (e == std::string::npos)
        e = s.size();
      if (e > b)
        sams.push_back(s.substr(b, e - b));
      b = e + 1;
    }
  }
  Options & opts = *Options::instance();
  gyper::log_singleton = std::unique_ptr<gyper::log_singleton_t>{new gyper::log_singleton_t{gyper::log_severity::warning, std::clog}};
  opts.vcf = vcf_fn;
  opts.threads = 1;
  opts.no_bamshrink = true;
  GenomicRegion padded{GenomicRegion(region_str)};
  padded.pad(1000);
  construct_graph(ref_fn, vcf_fn, padded.to_string(), false, true);
  absolute_pos.calculate_offsets(graph.contigs);

  // 1. index through the shim and the C ABI vs the reference's index_graph
  gtb_shim::FlatGraph fg = gtb_shim::flatten(graph);
  gtb_ctx * ctx = nullptr;
  gtb_shim::die(gtb_create(-1, &ctx));
  gtb_shim::die(gtb_region_begin(ctx, 0, &fg.view));
  uint64_t n_keys = 0, n_labels = 0;
  gtb_shim::die(gtb_index_size(ctx, 0, &n_keys, &n_labels));
  std::vector<uint64_t> keys(n_keys);
  std::vector<uint32_t> off(n_keys + 1);
  std::vector<gtb_label> labels(n_labels);
  gtb_shim::die(gtb_index_export(ctx, 0, keys.data(), off.data(), labels.data()));
  PHIndex const ref_index = index_graph(graph);
  if (ref_index.hamming0.size() != n_keys)
  {
    printf("SHIM FAIL: %zu keys in the reference index, %llu through the shim\n", ref_index.hamming0.size(), (unsigned long long)n_keys);
    return 1;
  }
  for (uint64_t i = 0; i < n_keys; ++i)
  {
    auto it = ref_index.hamming0.find(keys[i]);
    if (it == ref_index.hamming0.end() || it->second.size() != off[i + 1] - off[i])
    {
      printf("SHIM FAIL: key %llu\n", (unsigned long long)keys[i]);
      return 1;
    }
    for (size_t q = 0; q < it->second.size(); ++q)
    {
      gtb_label const & l = labels[off[i] + q];
      if (l.start != it->second[q].start_index || l.end != it->second[q].end_index || l.var_id != it->second[q].variant_id)
      {
        printf("SHIM FAIL: label %zu of key %llu\n", q, (unsigned long long)keys[i]);
        return 1;
      }
    }
  }

  // 2. the records of the pool, gathered the way INTEGRATION.md section 2b shows
  HtsParallelReader reader;
  reader.open(sams, "", ".");
  gtb_shim::Records recs;
  HtsRecord cur;
  while (reader.read_record(cur))
    if ((cur.record->core.flag & opts.sam_flag_filter) == 0u)
      recs.add(reader, cur);
  gtb_bam_batch const batch = recs.view();

  // 3. no CPU fallback: a host-only context must refuse the compute entry points
  int const rc_pool = gtb_pool_begin(ctx, 0, (int)reader.get_num_samples());
  int const rc_sub = gtb_submit_bam_records(ctx, 0, &batch, nullptr);
  if (rc_pool != GTB_ERR_CUDA || rc_sub != GTB_ERR_CUDA)
  {
    printf("SHIM FAIL: host-only context accepted a compute call (%d, %d)\n", rc_pool, rc_sub);
    return 1;
  }
  gtb_destroy(ctx);
  if (argc > 5 && std::string(argv[5]) == "--gpu")
    if (int rc = compare_with_reference_loop(sams, ref_index, fg, batch))
      return rc;
  printf("SHIM PASS keys=%llu labels=%llu records=%u data_bytes=%llu samples=%ld\n", (unsigned long long)n_keys,
         (unsigned long long)n_labels, batch.n_reads, (unsigned long long)recs.data.size(), reader.get_num_samples());
  return 0;
}
