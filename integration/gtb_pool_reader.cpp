// gtb_pool_reader.cpp -- the drop-in translation unit: libgtb200 behind the reference's own seams.
//
// Linked INTO the unmodified reference (oracle/ref_build/Makefile, target graphtyper_gtb) it replaces two functions and
// nothing else, so `graphtyper_gtb genotype ... --vcf=V` and `graphtyper_gtb genotype_sv ...` keep the CLI, graph
// construction, htslib I/O, per-region iteration, VCF merge and writing of the reference:
//
//   gyper::index_graph(Graph const &)                      include/graphtyper/index/indexer.hpp:16, src/index/indexer.cpp:246-291
//       -> flattens gyper::graph, uploads it and builds the k-mer index on the GPU (gtb_region_begin); returns an empty
//          PHIndex (nothing on the GPU path reads it)
//   gyper::parallel_reader_genotype_only(thread_id, ...)   include/graphtyper/utilities/hts_parallel_reader.hpp:69-82,
//                                                          src/utilities/hts_parallel_reader.cpp:458-1029
//       -> the pool thread entry: hands the pool's BAM files to the GPU as COMPRESSED bytes (gtb_submit_bgzf: the chunks of
//          the region iterators HtsReader::open computed from the index; inflate, filters, merge order, parsing, pairing on
//          the device) -- or, for what that entry declines (CRAM, several read groups per file, the SV coverage cap), reads
//          the records with the reference's HtsParallelReader (flag filter, SV read filter, coverage-bin cap exactly as the
//          loop applies them) and hands them over in ONE gtb_submit_bam_records call; fills VcfWriter::haplotypes /
//          ReferenceDepth from the accumulators and then finishes the pool with the reference's own code:
//          Vcf::add_haplotype, Variant::scan_calls, reformat_sv_vcf_records, save_vcf.
//
// The originals stay reachable as index_graph_cpu / parallel_reader_genotype_only_cpu (the Makefile renames the symbols in
// copies of indexer.o / hts_parallel_reader.o with objcopy): flows this path does not serve -- discovery iterations, primers,
// HLA / camou calling, GTB200_DISABLE=1, or a graph beyond the device capacities (a bubble with more than 32 alleles) -- run
// the reference's CPU code, announced in the log, never silently.
//
// One pool thread = one gtb_ctx sharing the region (gtb_region_attach), as the reference shares one PHIndex between its
// paw::Station workers (src/typer/caller.cpp:272-391).
#include <cstdlib>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include <graphtyper/constants.hpp>
#include <graphtyper/graph/graph.hpp>
#include <graphtyper/graph/reference_depth.hpp>
#include <graphtyper/graph/sv.hpp>
#include <graphtyper/index/indexer.hpp>
#include <graphtyper/index/ph_index.hpp>
#include <graphtyper/typer/primers.hpp>
#include <graphtyper/typer/vcf.hpp>
#include <graphtyper/typer/vcf_writer.hpp>
#include <graphtyper/utilities/hts_parallel_reader.hpp>
#include <graphtyper/utilities/hts_utils.hpp>
#include <graphtyper/utilities/logging.hpp>
#include <graphtyper/utilities/options.hpp>

#include "gtb_shim.hpp"

namespace gyper
{
// the reference's own implementations under their build-time aliases (see the header of this file)
PHIndex index_graph_cpu(Graph const & graph);
void parallel_reader_genotype_only_cpu(
  long thread_id,
  std::string * out_path,
  std::vector<std::string> const * hts_paths_ptr,
  std::vector<double> const * avg_cov_ptr,
  std::string const * output_dir_ptr,
  std::string const * reference_fn_ptr,
  std::string const * region_ptr,
  PHIndex const * ph_index_ptr,
  Primers const * primers,
  std::vector<std::map<std::pair<uint16_t, uint16_t>, std::map<std::pair<uint16_t, uint16_t>, int8_t>>> * ph_ptr,
  bool is_writing_calls_vcf,
  bool is_writing_hap,
  std::vector<std::unordered_map<uint32_t, uint32_t>> * allele_hap_gts_ptr);
} // namespace gyper

namespace
{
// Process-wide state of the GPU path, like gyper::graph itself: one region at a time, built by index_graph, used by the pool
// threads of the call() that follows.
struct Session
{
  std::mutex m;
  gtb_ctx * owner = nullptr;
  bool region_open = false;   // a region is resident in `owner`
  bool region_on_gpu = false; // the pools of the current graph run on the GPU
  int region_id = 0;
  std::vector<gtb_ctx *> idle; // pool-thread contexts between calls

  static Session & get()
  {
    static Session s;
    return s;
  }

  static int device()
  {
    const char * e = std::getenv("GTB200_DEVICE");
    return e ? std::atoi(e) : 0;
  }

  gtb_ctx * acquire()
  {
    {
      std::lock_guard<std::mutex> lk(m);
      if (!idle.empty())
      {
        gtb_ctx * c = idle.back();
        idle.pop_back();
        return c;
      }
    }
    gtb_ctx * c = nullptr;
    gtb_shim::die(gtb_create(device(), &c));
    return c;
  }

  void release(gtb_ctx * c)
  {
    std::lock_guard<std::mutex> lk(m);
    idle.push_back(c);
  }
};

bool gpu_path_wanted(gyper::Graph const & g)
{
  if (std::getenv("GTB200_DISABLE"))
    return false;
  // `genotype --vcf` (genotype_only_with_a_vcf, src/utilities/genotype.cpp:262-307) and `genotype_sv`
  // (src/utilities/genotype_sv.cpp:26-164): the two flows whose every index consumer is parallel_reader_genotype_only
  return g.is_sv_graph || !gyper::Options::const_instance()->vcf.empty();
}
} // namespace

namespace gyper
{
PHIndex index_graph(Graph const & g)
{
  Session & S = Session::get();
  std::lock_guard<std::mutex> lk(S.m);
  if (S.region_open)
  {
    gtb_shim::die(gtb_region_end(S.owner, S.region_id));
    S.region_open = false;
  }
  S.region_on_gpu = false;
  if (!gpu_path_wanted(g))
    return index_graph_cpu(g);
  if (!S.owner)
    gtb_shim::die(gtb_create(Session::device(), &S.owner));
  gtb_shim::FlatGraph const flat = gtb_shim::flatten(g);
  int const rc = gtb_region_begin(S.owner, ++S.region_id, &flat.view);
  if (rc == GTB_ERR_CAPACITY)
  {
    print_log(log_severity::warning, "[gtb200] ", gtb_last_error(), " -- this region runs on the reference's CPU path");
    return index_graph_cpu(g);
  }
  gtb_shim::die(rc);
  S.region_open = true;
  S.region_on_gpu = true;
  print_log(log_severity::debug, "[gtb200] graph uploaded and indexed on the device: ", flat.view.n_ref, " reference nodes, ",
            flat.view.n_var, " variant nodes");
  return PHIndex();
}

void parallel_reader_genotype_only(
  long const thread_id,
  std::string * out_path,
  std::vector<std::string> const * hts_paths_ptr,
  std::vector<double> const * avg_cov_ptr,
  std::string const * output_dir_ptr,
  std::string const * reference_fn_ptr,
  std::string const * region_ptr,
  PHIndex const * ph_index_ptr,
  Primers const * primers,
  std::vector<std::map<std::pair<uint16_t, uint16_t>, std::map<std::pair<uint16_t, uint16_t>, int8_t>>> * ph_ptr,
  bool const is_writing_calls_vcf,
  bool const is_writing_hap,
  std::vector<std::unordered_map<uint32_t, uint32_t>> * allele_hap_gts_ptr)
{
  Session & S = Session::get();
  bool on_gpu;
  int owner_region;
  {
    std::lock_guard<std::mutex> lk(S.m);
    on_gpu = S.region_on_gpu;
    owner_region = S.region_id;
  }
  if (!on_gpu || primers || allele_hap_gts_ptr)
  {
    if (on_gpu)
    {
      // primer trimming / HLA allele calls are not on the device path: this pool runs the reference's CPU code, which
      // needs the CPU index index_graph skipped
      print_log(log_severity::warning, "[gtb200] pool ", thread_id, " uses primers or HLA allele calls: reference CPU path");
      PHIndex const cpu_index = index_graph_cpu(graph);
      parallel_reader_genotype_only_cpu(thread_id, out_path, hts_paths_ptr, avg_cov_ptr, output_dir_ptr, reference_fn_ptr,
                                        region_ptr, &cpu_index, primers, ph_ptr, is_writing_calls_vcf, is_writing_hap,
                                        allele_hap_gts_ptr);
      return;
    }
    parallel_reader_genotype_only_cpu(thread_id, out_path, hts_paths_ptr, avg_cov_ptr, output_dir_ptr, reference_fn_ptr,
                                      region_ptr, ph_index_ptr, primers, ph_ptr, is_writing_calls_vcf, is_writing_hap,
                                      allele_hap_gts_ptr);
    return;
  }

  auto const & output_dir = *output_dir_ptr;
  Options const & opts = *Options::const_instance();

  // ---- the pool's reader and writer, set up as the reference does (hts_parallel_reader.cpp:488-510)
  HtsParallelReader hts_preader;
  hts_preader.open(*hts_paths_ptr, *reference_fn_ptr, *region_ptr);
  VcfWriter writer(opts.split_var_threshold - 1);
  writer.set_samples(hts_preader.get_samples());
  if (writer.pns.empty())
  {
    print_log(log_severity::error, __HERE__, " No samples were extracted.");
    std::exit(1);
  }
  std::string const & first_sample = writer.pns[0];
  long const n_samples = writer.pns.size();
  bool const is_sv = graph.is_sv_graph;
  ReferenceDepth reference_depth;
  if (is_sv)
    reference_depth.set_depth_sizes(n_samples);

  // ---- the device part: one context per pool thread on the shared region
  gtb_ctx * ctx = S.acquire();
  int const rid = 1;
  gtb_shim::die(gtb_region_attach(ctx, rid, S.owner, owner_region));
  gtb_shim::die(gtb_set_connections(ctx, is_writing_hap ? 1 : 0)); // HapSample::connections are only read when is_writing_hap
  gtb_shim::die(gtb_pool_begin(ctx, rid, (int)n_samples));

  // ---- first choice: the files' COMPRESSED bytes go to the device (gtb_submit_bgzf): inflate, record boundaries, the
  //      iterator's and the pool loop's filters, the merge order and everything after it happen there.  Not with the
  //      coverage-bin cap (order dependent, per sample, on the host: hts_parallel_reader.cpp:587-633), and not for what
  //      BgzfPool::collect declines; a stream the device rejects leaves the pool untouched.  GTB200_BGZF=0 switches it off.
  bool submitted = false;
  {
    static bool const bgzf_on = []() { const char * e = getenv("GTB200_BGZF"); return !e || atoi(e) != 0; }();
    bool cap_possible = false;
    if (is_sv && !opts.no_filter_on_coverage && avg_cov_ptr)
      for (double c : *avg_cov_ptr)
        cap_possible = cap_possible || c > 0.0;
    if (bgzf_on && !cap_possible)
    {
      gtb_shim::BgzfPool pool;
      std::string why;
      if (pool.collect(hts_preader, (uint32_t)opts.sam_flag_filter, is_sv, why))
      {
        gtb_submit_stats st{};
        int const rc = gtb_submit_bgzf(ctx, rid, (int)pool.files.size(), pool.files.data(), &pool.query, &st);
        if (rc == 0)
        {
          submitted = true;
          print_log(log_severity::info, "[gtb200] pool ", thread_id, ": ", st.n_records, " records decoded on the device from ",
                    pool.n_bytes, " compressed bytes");
        }
        else if (rc == GTB_ERR_INPUT || rc == GTB_ERR_CAPACITY)
          print_log(log_severity::warning, "[gtb200] pool ", thread_id, ": device decode declined (", gtb_last_error(),
                    "): reading with htslib");
        else
          gtb_shim::die(rc);
      }
      else
        print_log(log_severity::debug, "[gtb200] pool ", thread_id, ": reading with htslib (", why, ")");
    }
  }

  // ---- second choice: the record loop (hts_parallel_reader.cpp:570-716) without the per-record work: which records the pool
  //      sees, in merge order, read by the reference's HtsParallelReader.  Alignment, the duplicate shortcut, mate pairing
  //      and the accumulation happen on the device.
  gtb_shim::Records recs;
  long n_dup = 0;
  bool too_long = false; // a read beyond the device's read length (GTB_SEQ_STRIDE * 2 bases): found before anything is submitted
  if (!submitted)
  {
    auto rejected = [&](HtsRecord const & h)
    { return (h.record->core.flag & opts.sam_flag_filter) != 0u || (is_sv && !gtb_shim::is_good_read(h.record)); };
    auto sample_of = [&](HtsRecord const & h)
    {
      long s = 0, rg = 0;
      hts_preader.get_sample_and_rg_index(s, rg, h);
      return s;
    };
    HtsRecord prev, curr;
    bool more = hts_preader.read_record(prev);
    while (more && rejected(prev))
      more = hts_preader.read_record(prev);
    if (more)
    {
      gtb_shim::CoverageCap cap;
      cap.start(is_sv && !opts.no_filter_on_coverage, n_samples, prev.record->core.pos, avg_cov_ptr);
      cap.admit(sample_of(prev), prev.record->core.pos);
      recs.add(hts_preader, prev);
      too_long = prev.record->core.l_qseq > (int)(GTB_SEQ_STRIDE * 2);
      while (!too_long && hts_preader.read_record(curr))
      {
        if (rejected(curr))
          continue;
        too_long = curr.record->core.l_qseq > (int)(GTB_SEQ_STRIDE * 2);
        if (equal_pos_seq(prev.record, curr.record))
        {
          // same position and bases as the last aligned record: counted in its bin, never skipped
          cap.admit(sample_of(curr), curr.record->core.pos);
          ++n_dup;
          recs.add(hts_preader, curr);
        }
        else
        {
          if (!cap.admit(sample_of(curr), curr.record->core.pos))
            continue; // bin full
          recs.add(hts_preader, curr);
          hts_preader.move_record(prev, curr);
        }
      }
    }
    print_log(log_severity::debug, "[gtb200] pool ", thread_id, ": ", recs.core.size(), " records (", n_dup,
              " equal to their predecessor)");
  }

  if (too_long)
  {
    // The reference has no read-length limit in call() (MAX_READ_LENGTH is only a default variant distance); the device
    // path holds 152 bases per read.  Nothing has been added to the pool on the device: it runs the reference's CPU code.
    print_log(log_severity::warning, "[gtb200] pool ", thread_id, " holds a read longer than ", GTB_SEQ_STRIDE * 2,
              " bases: reference CPU path");
    gtb_shim::die(gtb_region_end(ctx, rid));
    S.release(ctx);
    recs = gtb_shim::Records();
    PHIndex const cpu_index = index_graph_cpu(graph);
    parallel_reader_genotype_only_cpu(thread_id, out_path, hts_paths_ptr, avg_cov_ptr, output_dir_ptr, reference_fn_ptr,
                                      region_ptr, &cpu_index, primers, ph_ptr, is_writing_calls_vcf, is_writing_hap,
                                      allele_hap_gts_ptr);
    return;
  }

  if (!submitted && !recs.core.empty())
  {
    gtb_bam_batch const batch = recs.view();
    gtb_shim::die(gtb_submit_bam_records(ctx, rid, &batch, nullptr));
  }
  gtb_shim::Accumulators acc;
  {
    uint32_t nb = 0, depth_size = 0, ref_off = 0;
    uint64_t n_scores = 0, n_cov = 0;
    gtb_shim::die(gtb_accumulator_sizes(ctx, rid, &nb, &n_scores, &n_cov));
    gtb_shim::die(gtb_ref_depth_size(ctx, rid, &depth_size, &ref_off));
    acc.resize(nb, n_scores, n_cov, (uint32_t)n_samples, depth_size);
    gtb_shim::die(gtb_pool_finish(ctx, rid, &acc.view));
  }
  std::vector<gtb_connection> conn;
  if (is_writing_hap)
  {
    uint64_t n = 0;
    gtb_shim::die(gtb_connections_size(ctx, rid, &n));
    conn.resize(n);
    if (n)
      gtb_shim::die(gtb_connections(ctx, rid, conn.data()));
  }
  gtb_shim::die(gtb_region_end(ctx, rid));
  S.release(ctx);
  recs = gtb_shim::Records();
  for (uint32_t v : acc.saturated)
    if (v)
    {
      print_log(log_severity::warning, "[gtb200] a likelihood reached the uint16 ceiling (haplotype.cpp:561, > ~6000x): the "
                                       "reference's result depends on the read order there");
      break;
    }
  gtb_shim::fill_writer(writer, reference_depth, acc.view, conn.data(), conn.size());

  // ---- pool finalisation: the reference's own objects and functions from here on (hts_parallel_reader.cpp:782-1029)
  if (is_writing_hap)
  {
    // the phase-support map of this pool (:782-893) is a pure function of coverage + connections: gtb_phase_support
    auto & ph = (*ph_ptr)[thread_id];
    uint64_t n = 0;
    gtb_shim::die(gtb_phase_support(&acc.view, conn.size(), conn.data(), &n, nullptr));
    std::vector<gtb_phase_support_entry> e(n);
    if (n)
      gtb_shim::die(gtb_phase_support(&acc.view, conn.size(), conn.data(), &n, e.data()));
    for (auto const & x : e)
      ph[{x.hap1, x.allele1}][{x.hap2, x.allele2}] |= x.flags;
  }
  auto calls_of_pool = [&](Vcf & vcf)
  {
    vcf.sample_names = writer.pns;
    for (long ps = 0; ps < (long)writer.haplotypes.size(); ++ps)
    {
      vcf.add_haplotype(writer.haplotypes[ps], (int32_t)ps);
      writer.haplotypes[ps].clear();
    }
  };
  if (is_writing_hap && !is_writing_calls_vcf)
  {
    Vcf vcf;
    calls_of_pool(vcf);
    for (Variant & var : vcf.variants)
    {
      var.scan_calls();
      var.calls.clear();
    }
    save_vcf(vcf, output_dir + "/" + first_sample);
  }
  if (is_writing_calls_vcf || opts.force_ignore_segment)
  {
    print_log(log_severity::debug, "[gtb200] Writing calls to '", output_dir, "/", first_sample, "_*'");
    Vcf vcf;
    calls_of_pool(vcf);
    if (is_sv)
    {
      reformat_sv_vcf_records(vcf.variants, reference_depth);
      // non-SV records end up last in reformat_sv_vcf_records: back into position order
      std::sort(vcf.variants.begin(), vcf.variants.end(), [](Variant const & a, Variant const & b)
                { return a.abs_pos < b.abs_pos || (a.abs_pos == b.abs_pos && a.seqs < b.seqs); });
      for (Variant & var : vcf.variants)
        var.stats.clear();
    }
    else if (!opts.is_segment_calling)
    {
      for (Variant & var : vcf.variants)
        var.scan_calls();
    }
    save_vcf(vcf, output_dir + "/" + first_sample);
  }
  *out_path = output_dir + "/" + first_sample;
}
} // namespace gyper
