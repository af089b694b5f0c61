// oracle/ref_build/paw_probe.cpp -- TEST INFRASTRUCTURE (golden-vector generator for the discovery re-alignment, N1).
//
// Runs the UNMODIFIED paw::pairwise_alignment (paw/include/paw/align/pairwise_alignment.hpp:146-388, AVX512 build with
// runtime dispatch) exactly as graphtyper's realign_to_indels configures it (src/typer/caller.cpp:1864-1870:
// match 1, mismatch 4, gap open 7, gap extend 1, clip 5, both end columns free, clipping on) on every
// (query, database) pair of a text file ("query<TAB>database" per line) and prints
//   score  database_begin  database_end  clip_begin  clip_end
// With "--time T" it instead loads all pairs, aligns them on T threads (one AlignmentOptions per thread, like one
// per graphtyper worker) and prints "pairs N seconds S threads T" -- the CPU baseline of tools/sw_bench.py.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include <paw/align.hpp>

int main(int argc, char ** argv)
{
  if (argc < 2)
  {
    fprintf(stderr, "usage: paw_probe pairs.tsv\n");
    return 2;
  }
  std::ifstream in(argv[1]);
  std::string line;
  using Tuint = uint16_t;
  if (argc >= 4 && std::string(argv[2]) == "--time")
  {
    int const T = std::max(1, atoi(argv[3]));
    std::vector<std::pair<std::string, std::string>> pairs;
    while (std::getline(in, line))
    {
      size_t const tab = line.find('\t');
      if (tab != std::string::npos)
        pairs.emplace_back(line.substr(0, tab), line.substr(tab + 1));
    }
    std::vector<long> sums(T, 0);
    auto const t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        paw::AlignmentOptions<Tuint> o;
        o.set_match(1).set_mismatch(4);
        o.set_gap_open(7).set_gap_extend(1);
        o.left_column_free = true;
        o.right_column_free = true;
        o.is_clip = true;
        for (size_t k = t; k < pairs.size(); k += T)
        {
          paw::pairwise_alignment(pairs[k].first, pairs[k].second, o);
          sums[t] += o.get_alignment_results()->score;
        }
      });
    for (auto & x : th)
      x.join();
    double const sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long tot = 0;
    for (long v : sums)
      tot += v;
    printf("pairs %zu seconds %.6f threads %d checksum %ld\n", pairs.size(), sec, T, tot);
    return 0;
  }
  paw::AlignmentOptions<Tuint> opts;
  opts.set_match(1).set_mismatch(4);
  opts.set_gap_open(7).set_gap_extend(1);
  opts.left_column_free = true;
  opts.right_column_free = true;
  opts.is_clip = true;
  while (std::getline(in, line))
  {
    size_t const tab = line.find('\t');
    if (tab == std::string::npos)
      continue;
    std::string const q = line.substr(0, tab), d = line.substr(tab + 1);
    paw::pairwise_alignment(q, d, opts);
    paw::AlignmentResults const & ar = *opts.get_alignment_results();
    printf("%d\t%ld\t%ld\t%ld\t%ld\n", (int)(int32_t)ar.score, (long)ar.database_begin, (long)ar.database_end, (long)ar.clip_begin, (long)ar.clip_end);
  }
  return 0;
}
