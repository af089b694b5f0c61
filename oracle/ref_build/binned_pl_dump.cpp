// binned_pl_dump.cpp -- test infrastructure: prints the reference's own PL / GQ output binning table
// (include/graphtyper/typer/binned_pl.hpp, applied when a VCF record is printed, src/typer/vcf.cpp:1107-1113) as 256
// numbers, so that a checker can compare raw PHRED calls with the sample columns of a VCF the reference CLI wrote.
// The table itself stays out of the repository: the output goes to oracle/_ref/gen/binned_pl.txt.
#include <cstdio>

#include <graphtyper/typer/binned_pl.hpp>

int main()
{
  for (unsigned v : binned_pl)
    std::printf("%u\n", v);
  return 0;
}
