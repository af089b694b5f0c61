/* sam2bam.c -- test infrastructure: SAM -> BAM + .bai with the vendored htslib (no samtools in this image).
 * `graphtyper genotype_sv` reads the original BAMs through their index (src/utilities/hts_reader.cpp:99-118), so the
 * CLI-level tests need indexed BAMs.  Usage: sam2bam in.sam out.bam */
#include <stdio.h>

#include <htslib/sam.h>

int main(int argc, char ** argv)
{
  if (argc != 3)
  {
    fprintf(stderr, "usage: sam2bam in.sam out.bam\n");
    return 2;
  }
  samFile * in = sam_open(argv[1], "r");
  samFile * out = sam_open(argv[2], "wb");
  if (!in || !out)
  {
    fprintf(stderr, "sam2bam: cannot open files\n");
    return 1;
  }
  sam_hdr_t * h = sam_hdr_read(in);
  if (!h || sam_hdr_write(out, h) < 0)
  {
    fprintf(stderr, "sam2bam: header\n");
    return 1;
  }
  bam1_t * b = bam_init1();
  int r;
  while ((r = sam_read1(in, h, b)) >= 0)
    if (sam_write1(out, h, b) < 0)
    {
      fprintf(stderr, "sam2bam: write\n");
      return 1;
    }
  bam_destroy1(b);
  sam_hdr_destroy(h);
  sam_close(in);
  if (sam_close(out) < 0 || r < -1)
    return 1;
  if (sam_index_build(argv[2], 0) < 0)
  {
    fprintf(stderr, "sam2bam: index\n");
    return 1;
  }
  return 0;
}
