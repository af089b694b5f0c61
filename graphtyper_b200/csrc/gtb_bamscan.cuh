// gtb_bamscan.cuh -- BAM records inside an inflated BGZF stream: the per-record rules of the reference's readers as host/device
// functions (product code).  The kernels of gtb_bgzf.cu and the CPU emulation behind
// gtb_debug_bgzf_host call exactly these functions, so the rules are pinned by the CPU test-suite and the GPU run adds only
// the parallel plumbing.
//
//   what                         reference
//   record layout                htslib sam.c: bam_read1 (32 fixed bytes after block_size; qname | cigar | seq | qual | aux)
//   end position                 htslib sam.c: bam_endpos / bam_cigar2rlen (unmapped or empty: pos + 1)
//   region iterator              htslib hts.c:4046-4098 hts_itr_next: a record is read while the offset after the previous one is
//                                below the chunk's end; the first record on another contig or at / beyond the region's end ends
//                                the iteration; records that do not overlap [beg, end) are skipped
//   flag filter                  src/utilities/hts_parallel_reader.cpp:655-663 (Options::sam_flag_filter)
//   SV read filter               src/utilities/hts_parallel_reader.cpp:528-568 (is_good_read)
//   order of the pool's records  src/utilities/hts_reader.cpp:166-303 (records of one position sorted by gt_pos_seq_same_pos) and
//                                hts_parallel_reader.cpp:66-136 (heap over the files by gt_pos_seq), include/graphtyper/utilities/
//                                hts_utils.hpp:48-108: ascending (contig, position, sequence length, packed sequence bytes).
//                                Records that tie in all four are exact duplicates for the pool loop (equal_pos_seq): whichever
//                                comes first is aligned, the others re-use its alignment, and every accumulator is a sum over
//                                records -- std::sort / the heap leave their order unspecified, here it is (file, file order).
#pragma once
#include <cstdint>

#include "gtb_inflate.cuh"

namespace gtb
{
struct BamQuery
{
  int32_t tid;
  long long beg, end;   // 0-based half-open, as hts_itr_t holds them
  uint32_t flag_filter; // records with any of these flag bits are dropped
  uint32_t sv_filter;   // 1: is_good_read
  uint32_t max_lseq;    // longer reads are an error (device read-length capacity)
  uint32_t whole_file;  // no region: every record of the file is returned (sam_read1 instead of the iterator; `genotype` reads
                        // its pools' files this way, src/utilities/hts_reader.cpp:94-97)
};

GTB_HD uint32_t le32(const uint8_t * p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
GTB_HD uint32_t le16(const uint8_t * p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

struct BamFixed
{
  int32_t block_size, tid, pos;
  uint32_t l_read_name, mapq, n_cigar, flag;
  int32_t l_seq, mtid, mpos, tlen;
};
// rec points at the record's block_size field; 36 bytes are readable
GTB_HD BamFixed bam_fixed(const uint8_t * rec)
{
  BamFixed f;
  f.block_size = (int32_t)le32(rec);
  f.tid = (int32_t)le32(rec + 4);
  f.pos = (int32_t)le32(rec + 8);
  f.l_read_name = rec[12];
  f.mapq = rec[13];
  f.n_cigar = le16(rec + 16);
  f.flag = le16(rec + 18);
  f.l_seq = (int32_t)le32(rec + 20);
  f.mtid = (int32_t)le32(rec + 24);
  f.mpos = (int32_t)le32(rec + 28);
  f.tlen = (int32_t)le32(rec + 32);
  return f;
}
// the variable part fits the record (bam_read1 rejects such records as truncated)
GTB_HD bool bam_layout_ok(const BamFixed & f)
{
  if (f.block_size < 32 || f.l_seq < 0 || f.l_read_name == 0)
    return false;
  unsigned long long const need = 32ull + f.l_read_name + 4ull * f.n_cigar + (unsigned long long)((f.l_seq + 1ll) / 2) + (unsigned long long)f.l_seq;
  return need <= (unsigned long long)f.block_size;
}
GTB_HD const uint8_t * bam_cigar_ptr(const uint8_t * rec, const BamFixed & f) { return rec + 36 + f.l_read_name; }
GTB_HD const uint8_t * bam_seq_ptr(const uint8_t * rec, const BamFixed & f) { return rec + 36 + f.l_read_name + 4ull * f.n_cigar; }

GTB_HD long long bam_end_position(const uint8_t * rec, const BamFixed & f)
{
  long long rlen = 0;
  if ((f.flag & 4u) == 0)
  {
    const uint8_t * cg = bam_cigar_ptr(rec, f);
    for (uint32_t k = 0; k < f.n_cigar; ++k)
    {
      uint32_t const c = le32(cg + 4ull * k), op = c & 15u;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) // M D N = X consume the reference
        rlen += c >> 4;
    }
  }
  return (long long)f.pos + (rlen == 0 ? 1 : rlen);
}

GTB_HD bool bam_good_read_sv(const uint8_t * rec, const BamFixed & f)
{
  if ((f.flag & 4u) != 0)
    return false;
  long long const d = (long long)f.pos - (long long)f.mpos;
  bool const mate_far = f.tid != f.mtid || (d < 0 ? -d : d) > 200000ll;
  if (f.mapq <= 15 && mate_far)
    return false;
  if (f.n_cigar >= 2)
  {
    const uint8_t * cg = bam_cigar_ptr(rec, f);
    uint32_t const front = le32(cg), back = le32(cg + 4ull * (f.n_cigar - 1));
    bool const front_clip = (front & 15u) == 4u, back_clip = (back & 15u) == 4u;
    bool const one_long = (front_clip && (front >> 4) >= 12) || (back_clip && (back >> 4) >= 12);
    if ((front_clip && back_clip) || (f.mapq <= 15 && one_long))
      return false;
  }
  return true;
}

constexpr int BAM_SKIP = 0, BAM_KEEP = 1, BAM_STOP = 2, BAM_TOO_LONG = 3, BAM_FILTERED = 4;
// What the iterator + the pool loop's filters do with one record that was read.  BAM_FILTERED records are returned by the
// iterator and dropped by the pool loop only after the merge: they take part in the ORDER of the pool's records (the per-file
// sort and the heap see them) and leave afterwards.
GTB_HD int bam_classify(const uint8_t * rec, const BamFixed & f, const BamQuery & q)
{
  if (!q.whole_file)
  {
    if (f.tid != q.tid || (long long)f.pos >= q.end)
      return BAM_STOP;
    if (!(bam_end_position(rec, f) > q.beg))
      return BAM_SKIP;
  }
  if ((f.flag & q.flag_filter) != 0)
    return BAM_FILTERED;
  if (q.sv_filter && !bam_good_read_sv(rec, f))
    return BAM_FILTERED;
  if ((uint32_t)f.l_seq > q.max_lseq)
    return BAM_TOO_LONG;
  return BAM_KEEP;
}

// first sort key: contig, position, sequence length as gt_pos_seq compares them (signed contig index: -1 sorts first).
// 16 | 32 | 16 bits; bam_key_fits says whether a record fits them.
GTB_HD unsigned long long bam_order_key(const BamFixed & f)
{
  return ((unsigned long long)(uint32_t)(f.tid + 1) << 48) | ((unsigned long long)(uint32_t)(f.pos + 1) << 16) | (uint32_t)f.l_seq;
}
GTB_HD bool bam_key_fits(const BamFixed & f) { return f.tid >= -1 && f.tid < 65535 && f.pos >= -1 && f.l_seq >= 0 && f.l_seq < 65536; }
// (contig, position) part and position part of a key
GTB_HD unsigned long long key_place(unsigned long long key) { return key >> 16; }
GTB_HD uint32_t key_pos(unsigned long long key) { return (uint32_t)(key >> 16); }
// ties of the first key: packed sequence bytes (cmp < 0, 0, > 0)
GTB_HD int bam_seq_compare(const uint8_t * a, const BamFixed & fa, const uint8_t * b, const BamFixed & fb)
{
  const uint8_t * sa = bam_seq_ptr(a, fa);
  const uint8_t * sb = bam_seq_ptr(b, fb);
  uint32_t const nb = (uint32_t)((fa.l_seq + 1) / 2);
  for (uint32_t j = 0; j < nb; ++j)
    if (sa[j] != sb[j])
      return sa[j] < sb[j] ? -1 : 1;
  return 0;
}
// Records that tie in the sequence bytes as well are exact duplicates; the reference leaves their order to std::sort (records
// of one position within a file) and to std::push_heap / pop_heap (between files).  Within a file libstdc++ sorts groups of up
// to 16 records by a stable insertion sort in DESCENDING order and the reader pops from the back (hts_reader.cpp:231-233), so
// exact ties come out in REVERSE file order: that is the device order.  Between files (and for larger groups) the order
// depends on the history of the heap (on introsort's pivots): reference_merge_order() below replays both with the standard
// library's own algorithms on the dense ranks the device sort produces.  Device order between files: ascending file index.
GTB_HD bool bam_tie_less(const uint8_t * a, const BamFixed & fa, uint32_t file_a, uint32_t ia, const uint8_t * b, const BamFixed & fb,
                         uint32_t file_b, uint32_t ib)
{
  int const c = bam_seq_compare(a, fa, b, fb);
  if (c != 0)
    return c < 0;
  if (file_a != file_b)
    return file_a < file_b;
  return ia > ib;
}

// ---- layout of one gtb_submit_bgzf call
struct BgzfBlock
{
  unsigned long long comp_off; // in the concatenated compressed bytes
  unsigned long long file_off; // of the block in its file (virtual offsets)
  unsigned long long out_off;  // in the inflated bytes
  uint32_t comp_bytes, isize;
  uint32_t segment;            // index of the block's segment
  uint32_t slot_base;          // first of the block's isize / 36 + 1 slots in the per-block record list
};
struct BgzfSegment
{
  uint32_t block_begin, block_end;
  unsigned long long out_begin, out_end;
  unsigned long long end_file_off; // file offset behind the segment's last block
  unsigned long long v_end;        // virtual offset at which the chunk ends
  uint32_t first_offset;           // of the first record inside the first block
  uint32_t to_eof;                 // the bytes run to the end of the file: running out of them is the end of the iteration
};
struct BgzfFile
{
  uint32_t seg_begin, seg_end;
  uint32_t rec_base, rec_cap; // slots of this file in the scanned-record arrays
  int32_t sample, rg;
};

// virtual offset of inflated offset g of a segment (g in [out_begin, out_end]), in the form bgzf_tell() returns it: the offset
// behind the last byte of a block is offset 0 of the next block (bgzf.c:1266-1269)
GTB_HD unsigned long long bgzf_voffset(const BgzfBlock * blocks, const BgzfSegment & s, unsigned long long g)
{
  if (g >= s.out_end)
    return s.end_file_off << 16;
  uint32_t lo = s.block_begin, hi = s.block_end; // last block with out_off <= g
  while (hi - lo > 1)
  {
    uint32_t const mid = (lo + hi) >> 1;
    if (blocks[mid].out_off <= g)
      lo = mid;
    else
      hi = mid;
  }
  return (blocks[lo].file_off << 16) | (g - blocks[lo].out_off);
}

constexpr int SCAN_OK = 0, SCAN_ERR_TRUNCATED = -20 /* a record that must be read runs past the bytes handed over */,
              SCAN_ERR_RECORD = -21 /* malformed record */, SCAN_ERR_CAPACITY = -22 /* more records than slots */,
              SCAN_ERR_KEY = -23 /* contig index / length beyond the sort key's fields */,
              SCAN_ERR_UNSORTED = -24 /* a file is not in coordinate order: the merge would not be the sorted order */;

// The records one file's iterator reads, in file order: start offsets into the inflated bytes.  Serial by nature (every
// record's length sits in its own first four bytes).  Returns the number of records read and the status.
GTB_HD uint32_t bam_walk_file(const uint8_t * out, const BgzfBlock * blocks, const BgzfSegment * segs, const BgzfFile & file,
                              const BamQuery & q, unsigned long long * rec_start, int * status)
{
  uint32_t n = 0;
  *status = SCAN_OK;
  for (uint32_t si = file.seg_begin; si < file.seg_end; ++si)
  {
    BgzfSegment const & s = segs[si];
    unsigned long long g = s.out_begin + s.first_offset;
    bool first = true;
    for (;;)
    {
      // hts_itr_next: before every read, "curr_off >= off[i].v" ends the chunk (curr_off = the offset behind the last record)
      if (!first && bgzf_voffset(blocks, s, g) >= s.v_end)
        break;
      first = false;
      if (g >= s.out_end)
      {
        if (s.to_eof)
          return n; // end of the file: readrec fails and the iterator is finished
        *status = SCAN_ERR_TRUNCATED; // the chunk is not over, but the bytes are
        return n;
      }
      if (g + 36 > s.out_end)
      {
        *status = SCAN_ERR_TRUNCATED;
        return n;
      }
      BamFixed const f = bam_fixed(out + g);
      if (!bam_layout_ok(f))
      {
        *status = SCAN_ERR_RECORD;
        return n;
      }
      if (g + 4ull + (unsigned long long)f.block_size > s.out_end)
      {
        *status = SCAN_ERR_TRUNCATED;
        return n;
      }
      if (n >= file.rec_cap)
      {
        *status = SCAN_ERR_CAPACITY;
        return n;
      }
      rec_start[file.rec_base + n++] = g;
      if (!q.whole_file && (f.tid != q.tid || (long long)f.pos >= q.end))
        return n; // the iterator is finished: later chunks are not read either
      g += 4ull + (unsigned long long)f.block_size;
    }
  }
  return n;
}

// ---- the same walk, one thread per BGZF block.  htslib never lets a record straddle two blocks (bam_write1 flushes the block
// when the next record does not fit, sam.c / bgzf_flush_try), so in practice every block starts at a record boundary.  That is
// only a guess here: bam_walk_block walks ONE block from its guessed entry and notes where it came out; bam_stitch_file then
// follows the file's blocks in order and accepts a block only if the walk before it came out exactly at its guessed entry
// (by induction from the chunk's known first record, accepted blocks hold true record boundaries); the first mismatch sends
// the whole file to the serial walk above.  Events that end a serial walk (chunk end, iterator finished, malformed or
// truncated record) end the block's walk at the same record, so the stitched result is the serial result.
constexpr uint32_t WALK_NONE = 0, WALK_VCUT = 1 /* the record at `halt` is not read: chunk end */, WALK_STOP = 2 /* read, and the
                   iterator is finished */, WALK_BAD = 3, WALK_TRUNC = 4;
struct BlockWalk
{
  unsigned long long exit; // inflated offset behind the last record walked (reason NONE: >= the block's end)
  uint32_t count;          // records recorded (for STOP including the one that stops)
  uint32_t reason;
};

GTB_HD BlockWalk bam_walk_block(const uint8_t * out, const BgzfBlock & blk, const BgzfSegment & s, bool first_of_segment,
                                const BamQuery & q, unsigned long long * slot)
{
  BlockWalk w{0, 0, WALK_NONE};
  unsigned long long g = blk.out_off + (first_of_segment ? s.first_offset : 0u);
  unsigned long long const block_end = blk.out_off + blk.isize;
  bool first = first_of_segment;
  while (g < block_end)
  {
    if (!first && ((blk.file_off << 16) | (g - blk.out_off)) >= s.v_end)
    {
      w.reason = WALK_VCUT;
      break;
    }
    first = false;
    if (g + 36 > s.out_end)
    {
      w.reason = WALK_TRUNC;
      break;
    }
    BamFixed const f = bam_fixed(out + g);
    if (!bam_layout_ok(f))
    {
      w.reason = WALK_BAD;
      break;
    }
    if (g + 4ull + (unsigned long long)f.block_size > s.out_end)
    {
      w.reason = WALK_TRUNC;
      break;
    }
    slot[w.count++] = g;
    g += 4ull + (unsigned long long)f.block_size;
    if (!q.whole_file && (f.tid != q.tid || (long long)f.pos >= q.end))
    {
      w.reason = WALK_STOP;
      break;
    }
  }
  w.exit = g;
  return w;
}

// Follows the blocks of a file.  take[b] = records of block b that the iterator reads, dst[b] = where they go in rec_start
// (file-major).  Returns false when a guess failed (the caller runs bam_walk_file instead); otherwise *n_rec / *status are
// what bam_walk_file would have returned.
GTB_HD bool bam_stitch_file(const BgzfBlock * blocks, const BgzfSegment * segs, const BgzfFile & file, const BlockWalk * walks,
                            uint32_t * take, uint32_t * dst, uint32_t * n_rec, int * status)
{
  uint32_t n = 0;
  *status = SCAN_OK;
  bool done = false;
  for (uint32_t si = file.seg_begin; si < file.seg_end; ++si)
  {
    BgzfSegment const & s = segs[si];
    for (uint32_t b = s.block_begin; b < s.block_end; ++b)
      take[b] = 0;
    if (done)
      continue;
    unsigned long long g = s.out_begin + s.first_offset;
    bool any = false, segment_over = false;
    for (uint32_t b = s.block_begin; b < s.block_end && !segment_over && !done; ++b)
    {
      unsigned long long const entry = blocks[b].out_off + (b == s.block_begin ? s.first_offset : 0u);
      if (g != entry)
      {
        if (g > entry && g >= blocks[b].out_off + blocks[b].isize)
          continue; // the last record ran over this whole block (a record larger than a block): nothing starts here
        return false;
      }
      BlockWalk const & w = walks[b];
      if (n + w.count > file.rec_cap)
      {
        *status = SCAN_ERR_CAPACITY;
        *n_rec = n;
        return true;
      }
      take[b] = w.count;
      dst[b] = file.rec_base + n;
      n += w.count;
      any = any || w.count > 0;
      g = w.exit;
      switch (w.reason)
      {
      case WALK_VCUT: segment_over = true; break;
      case WALK_STOP: done = true; break;
      case WALK_BAD: *status = SCAN_ERR_RECORD; *n_rec = n; return true;
      case WALK_TRUNC: *status = SCAN_ERR_TRUNCATED; *n_rec = n; return true;
      default: break;
      }
    }
    if (done || segment_over)
      continue;
    // the bytes of the segment are used up (g >= out_end), as in bam_walk_file
    if (g < s.out_end)
      return false; // a walk that stopped inside the segment without a reason cannot happen; be safe
    if (any && (s.end_file_off << 16) >= s.v_end)
      continue;
    if (s.to_eof)
      done = true;
    else
    {
      *status = SCAN_ERR_TRUNCATED;
      *n_rec = n;
      return true;
    }
  }
  *n_rec = n;
  return true;
}
} // namespace gtb
