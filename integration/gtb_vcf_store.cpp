// gtb_vcf_store.cpp -- pool results reach the merge without the filesystem.
//
// The reference hands every pool's calls to vcf_merge_and_break / vcf_merge_and_filter (src/typer/vcf_operations.cpp:
// 28-104,294-353,500-578) through files: save_vcf cuts the pool's Vcf into batches, serialises each with cereal through a
// gzip stream into <dir>/<n>, and load_vcf / append_vcf inflate and parse them again (src/typer/vcf.cpp:1662-1811) --
// about 13 % of a region's wall time.  Linked into the drop-in binary (oracle/ref_build/Makefile, graphtyper_gtb) this unit
// replaces those three functions: the batches keep the reference's own cereal serialisation (Vcf::serialize, so what
// comes back is bit for bit what the file round trip gives -- the final VCF stays byte identical) but live in a
// process-wide map keyed by the batch's would-be file name; no gzip, no file.  Paths that are real files (anything ending in
// .vcf.gz, or a batch this process never saved) still go to the originals, kept as save_vcf_cpu / load_vcf_cpu /
// append_vcf_cpu (objcopy-renamed copy of vcf.o).  GTB200_VCF_FILES=1 restores the files.
#include <cstdlib>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>

#include <cereal/archives/binary.hpp>

#include <graphtyper/typer/vcf.hpp>
#include <graphtyper/utilities/logging.hpp>
#include <graphtyper/utilities/options.hpp>

namespace gyper
{
void save_vcf_cpu(Vcf const & vcf, std::string const & filename);
void load_vcf_cpu(Vcf & vcf, std::string const & filename, long n_batch);
bool append_vcf_cpu(Vcf & vcf, std::string const & filename, long n_batch);
} // namespace gyper

namespace
{
struct Store
{
  std::mutex m;
  std::unordered_map<std::string, std::string> batches; // "<dir>/<n>" -> cereal bytes
  static Store & get()
  {
    static Store s;
    return s;
  }
};

bool files_wanted()
{
  static bool const v = std::getenv("GTB200_VCF_FILES") != nullptr;
  return v;
}

void put(std::string const & key, gyper::Vcf const & batch)
{
  std::ostringstream os(std::ios::binary);
  {
    cereal::BinaryOutputArchive oa(os);
    oa << batch;
  }
  Store & S = Store::get();
  std::lock_guard<std::mutex> lk(S.m);
  S.batches[key] = os.str();
}

// moves the batch out of the store: every batch is read once (the merge walks the pools batch by batch)
bool take(std::string const & key, gyper::Vcf & into)
{
  std::string bytes;
  {
    Store & S = Store::get();
    std::lock_guard<std::mutex> lk(S.m);
    auto it = S.batches.find(key);
    if (it == S.batches.end())
      return false;
    bytes = std::move(it->second);
    S.batches.erase(it);
  }
  std::istringstream is(bytes, std::ios::binary);
  cereal::BinaryInputArchive ia(is);
  ia >> into;
  return true;
}
} // namespace

namespace gyper
{
// save_vcf (vcf.cpp:1662-1751): batches of at most num_alleles_in_batch squared alternative alleles, the first one carries
// the sample names
void save_vcf(Vcf const & vcf, std::string const & filename)
{
  if (files_wanted())
    return save_vcf_cpu(vcf, filename);
  long const limit = Options::const_instance()->num_alleles_in_batch;
  long n_batch = 0, begin = 0, weight = 0;
  auto flush = [&](long end)
  {
    Vcf batch;
    if (n_batch == 0)
      batch.sample_names = vcf.sample_names;
    batch.variants.assign(vcf.variants.begin() + begin, vcf.variants.begin() + end);
    put(filename + "/" + std::to_string(n_batch), batch);
    begin = end;
    ++n_batch;
    weight = 0;
  };
  for (long v = 0; v < (long)vcf.variants.size(); ++v)
  {
    long const n_alts = (long)vcf.variants[v].seqs.size() - 1;
    weight += n_alts * n_alts;
    if (weight >= limit)
      flush(v + 1);
  }
  flush((long)vcf.variants.size()); // the rest (possibly empty), as the reference always writes a last batch
}

void load_vcf(Vcf & vcf, std::string const & filename, long n_batch)
{
  if (!take(filename + "/" + std::to_string(n_batch), vcf))
    load_vcf_cpu(vcf, filename, n_batch);
}

bool append_vcf(Vcf & vcf, std::string const & filename, long n_batch)
{
  Vcf batch;
  if (!take(filename + "/" + std::to_string(n_batch), batch))
    return append_vcf_cpu(vcf, filename, n_batch);
  std::move(batch.variants.begin(), batch.variants.end(), std::back_inserter(vcf.variants));
  return true;
}
} // namespace gyper
