"""oracle/bam_oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-Python restatement of what the reference's readers do with the records of a pool's BAM files, the independent
expectation of tests/test_bgzf_host.py for the BGZF entry (gtb_submit_bgzf):

  ParsedRecord.endpos        htslib sam.c: bam_endpos / bam_cigar2rlen (unmapped or empty CIGAR: pos + 1)
  ParsedRecord.good_read_sv  src/utilities/hts_parallel_reader.cpp:528-568 (is_good_read)
  iterate_region             htslib hts.c:4046-4098 (hts_itr_next over the index chunks: a chunk is read while the offset
                             behind the last record is below its end; the first record on another contig or at / beyond the
                             region's end finishes the iteration; records that do not overlap the region are skipped)
  expected_pool_records      src/utilities/hts_parallel_reader.cpp:655-663 (flag filter, after the merge) and the merge order of
                             hts_reader.cpp:166-303 + hts_parallel_reader.cpp:66-136 (include/graphtyper/utilities/
                             hts_utils.hpp:48-108): ascending (contig, position, length, packed sequence bytes).  The order of
                             EXACT duplicates is the reference's standard library's (std::sort, the heap's history); it is
                             pinned against the compiled reference itself by tests/test_bgzf_vs_reference.py, here they are
                             ordered by a fixed rule and compared modulo that order.
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

class ParsedRecord:
    __slots__ = ("raw", "tid", "pos", "l_name", "mapq", "n_cigar", "flag", "l_seq", "mtid", "mpos", "tlen", "cigar", "seq")

    def __init__(self, raw: bytes):
        self.raw = raw  # without block_size
        (self.tid, self.pos, self.l_name, self.mapq, _bin, self.n_cigar, self.flag, self.l_seq, self.mtid, self.mpos,
         self.tlen) = struct.unpack_from("<iiBBHHHiiii", raw, 0)
        o = 32 + self.l_name
        self.cigar = struct.unpack_from("<%dI" % self.n_cigar, raw, o)
        o += 4 * self.n_cigar
        self.seq = raw[o:o + (self.l_seq + 1) // 2]

    def endpos(self) -> int:
        rlen = 0
        if not (self.flag & 4):
            rlen = sum(c >> 4 for c in self.cigar if (c & 15) in (0, 2, 3, 7, 8))
        return self.pos + (rlen if rlen else 1)

    def good_read_sv(self) -> bool:
        if self.flag & 4:
            return False
        far = self.tid != self.mtid or abs(self.pos - self.mpos) > 200000
        if self.mapq <= 15 and far:
            return False
        if self.n_cigar >= 2:
            f, b = self.cigar[0], self.cigar[-1]
            fc, bc = (f & 15) == 4, (b & 15) == 4
            one_long = (fc and (f >> 4) >= 12) or (bc and (b >> 4) >= 12)
            if (fc and bc) or (self.mapq <= 15 and one_long):
                return False
        return True


def iterate_region(stream: bytes, blocks: List[Tuple[int, int, int]], chunks: Sequence[Tuple[int, int]], tid: int, beg: int,
                   end: int) -> List[ParsedRecord]:
    """hts_itr_next over `chunks` [(u, v)] of virtual offsets on the uncompressed `stream` cut into `blocks`."""
    file_end = blocks[-1][0] + 1 if blocks else 0

    def to_unc(v):
        co, uo = v >> 16, v & 0xFFFF
        for fo, at, n in blocks:
            if fo == co:
                return at + uo
        raise ValueError("virtual offset outside the blocks")

    def tell(unc):
        for i, (fo, at, n) in enumerate(blocks):
            if at <= unc < at + n:
                return (fo << 16) | (unc - at)
        return None  # at the end: the address behind the last data block (set by the caller)

    out = []
    i = -1
    curr = 0
    unc = 0
    n_off = len(chunks)
    while True:
        if curr == 0 or curr >= chunks[i][1]:
            if i == n_off - 1:
                break
            if i < 0 or chunks[i][1] != chunks[i + 1][0]:
                unc = to_unc(chunks[i + 1][0])
                curr = chunks[i + 1][0]
            i += 1
        if unc + 4 > len(stream):
            break
        (bs,) = struct.unpack_from("<i", stream, unc)
        rec = ParsedRecord(stream[unc + 4:unc + 4 + bs])
        unc += 4 + bs
        t = tell(unc)
        curr = t if t is not None else ((blocks[-1][0] + 0x10000) << 16)  # beyond every chunk end
        if rec.tid != tid or rec.pos >= end:
            break
        if rec.endpos() > beg and end > rec.pos:
            out.append(rec)
    return out


def expected_pool_records(per_file: Sequence[List[ParsedRecord]], flag_filter: int, sv_filter: bool):
    """Pool loop filters + merge order: [(file index, record)] ascending (pos, l_seq, packed seq bytes); exact ties by file,
    within a file in reverse file order (libstdc++'s stable descending insertion sort, popped from the back)."""
    rows = []
    for fi, recs in enumerate(per_file):
        for k, r in enumerate(recs):
            if r.flag & flag_filter:
                continue
            if sv_filter and not r.good_read_sv():
                continue
            rows.append(((r.tid, r.pos, r.l_seq, r.seq, -fi, -k), fi, r))
    rows.sort(key=lambda x: x[0])
    return [(fi, r) for _, fi, r in rows]


