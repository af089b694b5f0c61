// oracle/ref_build/shim_probe.cpp -- TEST INFRASTRUCTURE: the reference-side shim (integration/gtb_shim.hpp) compiled against
// the UNMODIFIED reference objects and libgtb200.so, exercised where no GPU is needed:
//   1. construct_graph (reference) -> gtb_shim::flatten -> gtb_region_begin on a host-only context (host index builder)
//      -> gtb_index_export, compared key by key and label by label (bucket order included) with index_graph (reference);
//   2. the pool's records gathered by gtb_shim::Records through the reference's own HtsParallelReader + flag filter;
//   3. the compute entry points refuse to run without a device (no CPU fallback).
// Prints "SHIM PASS ..." or the first difference.
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

#include "../../integration/gtb_shim.hpp"

int main(int argc, char ** argv)
{
  using namespace gyper;
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
  printf("SHIM PASS keys=%llu labels=%llu records=%u data_bytes=%llu samples=%ld\n", (unsigned long long)n_keys,
         (unsigned long long)n_labels, batch.n_reads, (unsigned long long)recs.data.size(), reader.get_num_samples());
  return 0;
}
