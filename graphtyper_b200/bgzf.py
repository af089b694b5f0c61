"""BGZF / BAM helpers of the harness (tests, bench): writes BAM files from record batches and cuts them into the compressed
segments gtb_submit_bgzf takes (ctypes views of include/gtb200.h).  The plain-Python restatement of htslib's region iterator
and the pool's merge order that the tests check against lives in oracle/bam_oracle.py.  Nothing here is on the product path."""
from __future__ import annotations

import ctypes as C
import struct
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_compress(data: bytes, block_size: int = 0xFF00, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
                  eof: bool = True) -> Tuple[bytes, List[Tuple[int, int, int]]]:
    """data -> BGZF file bytes; also [(file offset, uncompressed offset, uncompressed size)] per data block."""
    out = bytearray()
    blocks = []
    for at in range(0, len(data), block_size):
        chunk = data[at:at + block_size]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8
        assert bsize <= 65536
        blocks.append((len(out), at, len(chunk)))
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize - 1)
        out += body
        out += struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    if eof:
        out += EOF_BLOCK
    return bytes(out), blocks


def bam_header(refs: Sequence[Tuple[str, int]], text: str = "@HD\tVN:1.6\tSO:coordinate\n") -> bytes:
    h = bytearray(b"BAM\x01")
    t = text.encode()
    h += struct.pack("<i", len(t)) + t + struct.pack("<i", len(refs))
    for name, length in refs:
        nm = name.encode() + b"\0"
        h += struct.pack("<i", len(nm)) + nm + struct.pack("<i", length)
    return bytes(h)


def bam_record(tid: int, pos: int, mapq: int, flag: int, name: bytes, cigar: Sequence[int], seq4: bytes, l_seq: int,
               qual: bytes, aux: bytes, mtid: int = -1, mpos: int = -1, tlen: int = 0) -> bytes:
    """One record in FILE layout (block_size first).  name without NUL; cigar as packed uint32 (len << 4 | op)."""
    nm = name + b"\0"
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, 0, len(cigar), flag, l_seq, mtid, mpos, tlen)
    body += nm + b"".join(struct.pack("<I", c) for c in cigar) + seq4 + qual + aux
    return struct.pack("<i", len(body)) + body


def records_from_batch(bam: abi.HostBamBatch, sel: Optional[Sequence[int]] = None) -> List[bytes]:
    """File-layout records of a HostBamBatch (htslib memory layout: l_qname counts the padding NULs htslib adds)."""
    out = []
    idx = range(len(bam)) if sel is None else sel
    data = bam.data.tobytes()
    for k in idx:
        c = bam.core[k]
        d = data[int(bam.data_off[k]):int(bam.data_off[k + 1])]
        lq = int(c["l_qname"])
        name = d[:lq].split(b"\0")[0]
        rest = d[lq:]
        nm = name + b"\0"
        body = struct.pack("<iiBBHHHiiii", int(c["tid"]), int(c["pos"]), len(nm), int(c["mapq"]), 0, int(c["n_cigar"]),
                           int(c["flag"]), int(c["l_qseq"]), int(c["mtid"]), int(c["mpos"]), int(c["isize"]))
        body += nm + rest
        out.append(struct.pack("<i", len(body)) + body)
    return out


# ------------------------------------------------------------------------------------------------ ctypes views
class BgzfSegment(C.Structure):
    _fields_ = [("comp", C.c_void_p), ("comp_bytes", C.c_uint64), ("file_offset", C.c_uint64), ("v_end", C.c_uint64),
                ("first_offset", C.c_uint32), ("to_eof", C.c_uint32)]


class BgzfFile(C.Structure):
    _fields_ = [("n_segments", C.c_uint32), ("reserved", C.c_uint32), ("segments", C.POINTER(BgzfSegment)),
                ("sample", C.c_int32), ("rg", C.c_int32)]


class BgzfQuery(C.Structure):
    _fields_ = [("tid", C.c_int32), ("flag_filter", C.c_uint32), ("beg", C.c_int64), ("end", C.c_int64),
                ("sv_read_filter", C.c_uint32), ("check_crc", C.c_uint32), ("whole_file", C.c_uint32), ("reserved", C.c_uint32)]


class HostBgzfFiles:
    """Owns the compressed bytes and the gtb_bgzf_file array of one pool.
    files: [(file bytes (np.uint8 array or bytes), [(u, v, to_eof)] chunks of virtual offsets, sample, rg)].  A chunk's bytes
    run from the block of u to `tail_blocks` blocks behind the block of v (or to the end of the file)."""

    def __init__(self, files, block_offsets: Sequence[Sequence[int]], tail_blocks: int = 2, pin=None):
        self.keep = []
        self.n_files = len(files)
        self.files = (BgzfFile * max(1, self.n_files))()
        self.comp_bytes = 0
        for fi, (raw, chunks, sample, rg) in enumerate(files):
            arr = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
            offs = sorted(block_offsets[fi])  # file offsets of all block starts, incl. the EOF block; + file length
            segs = (BgzfSegment * max(1, len(chunks)))()
            for si, (u, v, to_eof) in enumerate(chunks):
                start = u >> 16
                if to_eof:
                    stop = len(arr)
                else:
                    vb = v >> 16
                    later = [o for o in offs if o > vb]
                    stop = later[min(tail_blocks, len(later)) - 1] if later else len(arr)
                piece = np.ascontiguousarray(arr[start:stop])
                if pin is not None:
                    piece = pin.take(piece)
                self.keep.append(piece)
                segs[si].comp = piece.ctypes.data
                segs[si].comp_bytes = len(piece)
                segs[si].file_offset = start
                segs[si].v_end = v
                segs[si].first_offset = u & 0xFFFF
                segs[si].to_eof = 1 if (to_eof or stop == len(arr)) else 0
                self.comp_bytes += len(piece)
            self.keep.append(segs)
            self.files[fi].n_segments = len(chunks)
            self.files[fi].segments = segs
            self.files[fi].sample = sample
            self.files[fi].rg = rg


def query(tid: int, beg: int, end: int, flag_filter: int = 3840, sv: bool = False, check_crc: bool = True,
          whole_file: bool = False) -> BgzfQuery:
    q = BgzfQuery()
    q.whole_file = 1 if whole_file else 0
    q.tid, q.beg, q.end, q.flag_filter = tid, beg, end, flag_filter
    q.sv_read_filter = 1 if sv else 0
    q.check_crc = 1 if check_crc else 0
    return q


def build_pool_files(bam: abi.HostBamBatch, refs: Sequence[Tuple[str, int]], block_size: int = 0xFF00, level: int = 6,
                     strategy: int = zlib.Z_DEFAULT_STRATEGY, decoys=None):
    """One BAM file per (sample, read group) of the batch, records in batch order (= coordinate order).
    decoys(file index, records) may insert extra file-layout records.  Returns [(file bytes, stream, blocks, header length,
    sample, rg)]."""
    keys = sorted({(int(s), int(g)) for s, g in zip(bam.sample, bam.rg)})
    out = []
    for fi, (s, g) in enumerate(keys):
        sel = [k for k in range(len(bam)) if int(bam.sample[k]) == s and int(bam.rg[k]) == g]
        recs = records_from_batch(bam, sel)
        if decoys is not None:
            recs = decoys(fi, recs)
        hdr = bam_header(refs)
        stream = hdr + b"".join(recs)
        raw, blocks = bgzf_compress(stream, block_size, level, strategy)
        out.append((raw, stream, blocks, len(hdr), s, g))
    return out


def voffset_of(blocks, unc: int) -> int:
    for fo, at, n in blocks:
        if at <= unc < at + n:
            return (fo << 16) | (unc - at)
    raise ValueError("offset outside the data blocks")


def scan_blocks(raw: bytes) -> List[Tuple[int, int, int]]:
    """[(file offset, uncompressed offset, uncompressed size)] of the data blocks of a BGZF file (headers only)."""
    out = []
    at = unc = 0
    while at + 18 <= len(raw):
        (bsize,) = struct.unpack_from("<H", raw, at + 16)
        (isize,) = struct.unpack_from("<I", raw, at + bsize + 1 - 4)
        if isize:
            out.append((at, unc, isize))
        unc += isize
        at += bsize + 1
    return out


def inflate_file(raw: bytes) -> bytes:
    """The whole uncompressed stream of a BGZF file, by zlib (harness only)."""
    out = bytearray()
    at = 0
    while at + 18 <= len(raw):
        (bsize,) = struct.unpack_from("<H", raw, at + 16)
        (xlen,) = struct.unpack_from("<H", raw, at + 10)
        out += zlib.decompress(raw[at + 12 + xlen:at + bsize + 1 - 8], -15)
        at += bsize + 1
    return bytes(out)


def bam_header_length(stream: bytes) -> int:
    assert stream[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", stream, 4)
    at = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", stream, at)
    at += 4
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", stream, at)
        at += 4 + l_name + 4
    return at


def bgzf_compress_records(header: bytes, records: Sequence[bytes], block_size: int = 0xFF00, level: int = 6):
    """BGZF bytes as htslib writes BAM: the header in blocks of its own, then records, a block being flushed when the next
    record does not fit (bam_write1 / bgzf_flush_try: no record straddles two blocks).  Returns (bytes, blocks)."""
    chunks = [header[at:at + block_size] for at in range(0, len(header), block_size)]
    cur = bytearray()
    for r in records:
        if len(cur) + len(r) > block_size and cur:
            chunks.append(bytes(cur))
            cur = bytearray()
        cur += r
    if cur:
        chunks.append(bytes(cur))
    out = bytearray()
    blocks = []
    unc = 0
    for chunk in chunks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8
        blocks.append((len(out), unc, len(chunk)))
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
        unc += len(chunk)
    out += EOF_BLOCK
    return bytes(out), blocks
