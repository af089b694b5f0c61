"""Shared builders of the BGZF tests: BAM files made from the golden fixtures' records (one file per sample / read group, with
decoy records the readers must drop) and the segment lists gtb_submit_bgzf takes."""
import struct
import zlib

import numpy as np

from oracle import bam_oracle
from graphtyper_b200 import abi, bgzf, gtba

REFS = [("chr1", 250000000), ("chr2", 240000000)]


def fixture_batch(pre):
    rd = gtba.load(pre + ".reads.gtba")
    return rd, abi.HostBamBatch.from_probe(rd)


def region_of(bam):
    """A query region that every record of the batch overlaps the way the reference's run did: [first pos, last pos + 1)."""
    pos = bam.core["pos"].astype(np.int64)
    return int(bam.core["tid"][0]) if len(bam) else 0, int(pos.min()) if len(bam) else 0, int(pos.max()) + 1 if len(bam) else 1


def decoy_maker(tid, beg, end, seed=0):
    """Records every reader must drop: filtered flags inside the region, records that end at or before `beg`, and a tail at /
    behind `end` or on the next contig (the first of them ends the file's iteration)."""
    rng = np.random.default_rng(seed)

    def rec(name, t, pos, flag, l=60, cig=None):
        seq = bytes(rng.integers(0x11, 0x89, (l + 1) // 2, dtype=np.uint8))
        return bgzf.bam_record(t, pos, 30, flag, name, cig if cig is not None else [(l << 4) | 0], seq, l, bytes([30]) * l,
                               b"ASC\x10", t, pos + 5, 100)

    def make(fi, recs):
        out = []
        if beg >= 80:  # ends before the region starts: skipped by the iterator
            out.append(rec(b"before%d" % fi, tid, beg - 70, 99))
            out.append(rec(b"unmapped_before%d" % fi, tid, beg - 1, 4 | 1))  # unmapped: end = pos + 1 = beg -> no overlap
        k = 0
        for r in recs:
            out.append(r)
            k += 1
            if k % 97 == 0:  # filtered flags at a real record's position
                (bs, t, p) = struct.unpack_from("<iii", r, 0)
                for flag in (0x400 | 99, 0x100 | 83, 0x800 | 163, 0x200 | 99):
                    out.append(rec(b"flt%d_%d_%x" % (fi, k, flag), t, p, flag))
        out.append(rec(b"beyond%d" % fi, tid, end, 99))          # at the region's end: the iterator finishes here
        out.append(rec(b"beyond2_%d" % fi, tid, end + 10, 99))
        out.append(rec(b"other%d" % fi, tid + 1, 5, 99))
        return out
    return make


def pool_files(bam, tid, beg, end, block_size=0xFF00, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, with_decoys=True, seed=0):
    made = bgzf.build_pool_files(bam, REFS, block_size, level, strategy, decoy_maker(tid, beg, end, seed) if with_decoys else None)
    return made


def whole_file_segments(made, tail_blocks=2, pin=None):
    """One chunk per file: from the first record behind the header to the end of the file."""
    files, offs = [], []
    for raw, stream, blocks, hlen, s, g in made:
        u = bgzf.voffset_of(blocks, hlen) if len(stream) > hlen else (blocks[-1][0] << 16)
        files.append((raw, [(u, len(raw) << 16, True)], s, g))
        offs.append([b[0] for b in blocks] + [len(raw) - len(bgzf.EOF_BLOCK), len(raw)])
    return bgzf.HostBgzfFiles(files, offs, tail_blocks, pin)


def record_starts(stream, hlen):
    out = []
    at = hlen
    while at + 4 <= len(stream):
        (bs,) = struct.unpack_from("<i", stream, at)
        out.append(at)
        at += 4 + bs
    return out


def chunked_segments(made, n_chunks=3, gap=5, tail_blocks=2):
    """Several non-adjacent chunks per file: chunk boundaries at record starts, `gap` records between two chunks are not
    covered (the iterator still reads the first of them: it reads until the offset BEHIND a record reaches the chunk's end)."""
    files, offs, chunk_lists = [], [], []
    for raw, stream, blocks, hlen, s, g in made:
        starts = record_starts(stream, hlen)
        n = len(starts)
        chunks = []
        per = max(1, n // n_chunks)
        for c in range(n_chunks):
            a = c * per
            b = n if c == n_chunks - 1 else max(a + 1, (c + 1) * per - gap)
            if a >= n:
                break
            u = bgzf.voffset_of(blocks, starts[a])
            v = bgzf.voffset_of(blocks, starts[b]) if b < n else (len(raw) - len(bgzf.EOF_BLOCK)) << 16
            chunks.append((u, v, b >= n))
        files.append((raw, chunks, s, g))
        offs.append([b[0] for b in blocks] + [len(raw) - len(bgzf.EOF_BLOCK), len(raw)])
        chunk_lists.append([(u, v) for u, v, _ in chunks])
    return bgzf.HostBgzfFiles(files, offs, tail_blocks), chunk_lists


def expected_batch(made, chunk_lists, tid, beg, end, flag_filter=3840, sv=False):
    """Plain-Python expectation as a HostBamBatch in file layout (l_qname = name length + 1, no padding)."""
    per_file = []
    for (raw, stream, blocks, hlen, s, g), chunks in zip(made, chunk_lists):
        per_file.append(bam_oracle.iterate_region(stream, blocks, chunks, tid, beg, end))
    rows = bam_oracle.expected_pool_records(per_file, flag_filter, sv)
    n = len(rows)
    core = np.zeros(n, abi.BAM_CORE_DTYPE)
    data = bytearray()
    off = np.zeros(n + 1, np.uint64)
    smp, rg = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for j, (fi, r) in enumerate(rows):
        core[j]["pos"], core[j]["mpos"], core[j]["isize"] = r.pos, r.mpos, r.tlen
        core[j]["tid"], core[j]["mtid"], core[j]["l_qseq"] = r.tid, r.mtid, r.l_seq
        core[j]["n_cigar"], core[j]["flag"], core[j]["l_qname"], core[j]["mapq"] = r.n_cigar, r.flag, r.l_name, r.mapq
        data += r.raw[32:]
        off[j + 1] = len(data)
        smp[j], rg[j] = made[fi][4], made[fi][5]
    return abi.HostBamBatch(core, np.frombuffer(bytes(data), np.uint8), off, smp, rg)


def whole_file_chunks(made):
    out = []
    for raw, stream, blocks, hlen, s, g in made:
        u = bgzf.voffset_of(blocks, hlen) if len(stream) > hlen else (blocks[-1][0] << 16)
        out.append([(u, len(raw) << 16)])
    return out


def _records(b):
    raw = b.data.tobytes()
    out = []
    for k in range(len(b)):
        c = b.core[k]
        d = raw[int(b.data_off[k]):int(b.data_off[k + 1])]
        o = int(c["l_qname"]) + 4 * int(c["n_cigar"])
        seq = d[o:o + (int(c["l_qseq"]) + 1) // 2]
        out.append(((int(c["pos"]), int(c["l_qseq"]), seq), (int(b.sample[k]), int(b.rg[k]), tuple(int(c[f]) for f in c.dtype.names if f != "reserved"), d)))
    return out


def assert_batches_equal(a, b, what="", modulo_ties=False):
    """Exact equality of two record batches; modulo_ties: exact duplicates (same position, length, sequence) may come in any
    order -- between files the reference's order is its heap's history (pinned by test_bgzf_vs_reference.py), which the
    plain-Python expectation does not replay."""
    assert len(a) == len(b), (what, len(a), len(b))
    if modulo_ties:
        ra, rb = _records(a), _records(b)
        assert [k for k, _ in ra] == [k for k, _ in rb], what  # the order of everything that is not an exact tie
        assert sorted(ra) == sorted(rb), what
        return
    for f in a.core.dtype.names:
        if f == "reserved":
            continue
        assert np.array_equal(a.core[f], b.core[f]), (what, f)
    assert np.array_equal(a.data_off, b.data_off), what
    assert np.array_equal(a.data, b.data), what
    assert np.array_equal(a.sample, b.sample) and np.array_equal(a.rg, b.rg), what
