"""BGZF decode + record selection + merge order: the host/device source functions of csrc/gtb_inflate.cuh and
csrc/gtb_bamscan.cuh, run serially on the CPU through gtb_debug_bgzf_host, against zlib and a plain-Python restatement of
htslib's region iterator and the reference's merge order (oracle/bam_oracle.py).  The GPU kernels call the same functions
(tests/test_gpu_bgzf.py compares them with this emulation bit for bit)."""
import os
import zlib

import numpy as np
import pytest

import bgzf_cases as cases
from conftest import fixture_prefixes
from oracle import bam_oracle
from graphtyper_b200 import abi, bgzf, engine

SMALL = [p for p in fixture_prefixes(include_big=False)]
MODES = [(6, zlib.Z_DEFAULT_STRATEGY, 0xFF00), (0, zlib.Z_DEFAULT_STRATEGY, 0x8000), (6, zlib.Z_FIXED, 0xFF00),
         (9, zlib.Z_DEFAULT_STRATEGY, 3000), (1, zlib.Z_HUFFMAN_ONLY, 0xFF00), (4, zlib.Z_RLE, 20000)]


@pytest.mark.parametrize("level,strategy,block", MODES)
def test_inflate_equals_zlib_source(level, strategy, block):
    """Dynamic, fixed and stored DEFLATE blocks, match-free and run-length streams, small and full-size BGZF blocks; the
    CRC-32 of every block is verified through the combine identity the warp uses."""
    rd, bam = cases.fixture_batch(SMALL[2])
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end, block, level, strategy)
    files = cases.whole_file_segments(made)
    got, infl = engine.bgzf_host(files, bgzf.query(tid, beg, end), want_inflated=True)
    at = 0
    for raw, stream, blocks, hlen, s, g in made:
        first = [b for b in blocks if b[1] <= hlen < b[1] + b[2]][0]
        want = stream[first[1]:]
        assert bytes(infl[at:at + len(want)]) == want
        at = (at + len(want) + 15) // 16 * 16 + 64
    assert len(got) > 0


def test_corrupt_and_truncated_blocks_are_reported():
    rd, bam = cases.fixture_batch(SMALL[0])
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end)
    raw, stream, blocks, hlen, s, g = made[0]
    offs = [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]]
    u = bgzf.voffset_of(blocks, hlen)
    bad = bytearray(raw)
    bad[blocks[0][0] + 40] ^= 0x55  # somewhere in the first block's DEFLATE stream
    with pytest.raises(engine.GtbError) as e:
        engine.bgzf_host(bgzf.HostBgzfFiles([(bytes(bad), [(u, len(raw) << 16, True)], 0, 0)], offs), bgzf.query(tid, beg, end))
    assert e.value.code == -5
    crc = bytearray(raw)
    crc[blocks[0][0] + (blocks[1][0] if len(blocks) > 1 else len(raw) - 28) - blocks[0][0] - 8] ^= 1  # the stored CRC-32
    with pytest.raises(engine.GtbError) as e:
        engine.bgzf_host(bgzf.HostBgzfFiles([(bytes(crc), [(u, len(raw) << 16, True)], 0, 0)], offs), bgzf.query(tid, beg, end))
    assert e.value.code == -5 and "CRC" in str(e.value)
    engine.bgzf_host(bgzf.HostBgzfFiles([(bytes(crc), [(u, len(raw) << 16, True)], 0, 0)], offs),
                     bgzf.query(tid, beg, end, check_crc=False))  # not checked: decodes
    # bytes that end inside the chunk: the first block only, chunk end far behind it
    if len(blocks) > 1:
        cut = raw[:blocks[1][0]]
        files = bgzf.HostBgzfFiles([(cut, [(u, len(raw) << 16, False)], 0, 0)], [[blocks[0][0], blocks[1][0]]])
        files.files[0].segments[0].to_eof = 0
        with pytest.raises(engine.GtbError) as e:
            engine.bgzf_host(files, bgzf.query(tid, beg, end))
        assert e.value.code == -5 and "end before" in str(e.value)
    with pytest.raises(engine.GtbError) as e:  # not BGZF at all
        engine.bgzf_host(bgzf.HostBgzfFiles([(b"\0" * 100, [(0, 100 << 16, True)], 0, 0)], [[0, 100]]), bgzf.query(0, 0, 10))
    assert e.value.code == -5


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_whole_file_records_equal_python_iterator(pre):
    """One chunk per file (sample / read group), decoys around and between the records: the emulation's record batch equals
    the plain-Python iterator + filters + merge order, byte for byte."""
    rd, bam = cases.fixture_batch(pre)
    if len(bam) == 0:
        pytest.skip("no records")
    tid, beg, end = cases.region_of(bam)
    sv = "sv" in os.path.basename(pre)
    made = cases.pool_files(bam, tid, beg, end, block_size=0x4000)
    files = cases.whole_file_segments(made)
    got = engine.bgzf_host(files, bgzf.query(tid, beg, end, sv=sv))
    want = cases.expected_batch(made, cases.whole_file_chunks(made), tid, beg, end, sv=sv)
    cases.assert_batches_equal(got, want, pre, modulo_ties=len(made) > 1)
    # and it is the fixture's own record set (the decoys are all dropped): same multiset of (sample, pos, flag, name + rest)
    assert len(got) == len(bam)


@pytest.mark.parametrize("n_chunks,gap,block", [(2, 4, 0x1000), (4, 9, 0x800), (3, 1, 0xFF00)])
def test_chunked_files_follow_the_iterator_rules(n_chunks, gap, block):
    """Several non-adjacent chunks per file: a chunk is read until the offset BEHIND a record reaches its end (so the first
    record of a gap is still read), the next chunk starts with a seek."""
    rd, bam = cases.fixture_batch(SMALL[2])
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end, block_size=block)
    files, chunk_lists = cases.chunked_segments(made, n_chunks, gap)
    got = engine.bgzf_host(files, bgzf.query(tid, beg, end))
    want = cases.expected_batch(made, chunk_lists, tid, beg, end)
    cases.assert_batches_equal(got, want, "chunked", modulo_ties=True)
    assert 0 < len(got) < len(bam)  # the gaps are really skipped


def test_region_bounds_and_long_reads():
    rd, bam = cases.fixture_batch(SMALL[2])
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end)
    mid = (beg + end) // 2
    files = cases.whole_file_segments(made)
    got = engine.bgzf_host(files, bgzf.query(tid, mid, mid + 40))
    want = cases.expected_batch(made, cases.whole_file_chunks(made), tid, mid, mid + 40)
    cases.assert_batches_equal(got, want, "sub-region", modulo_ties=True)
    assert 0 < len(got) < len(bam)
    empty = engine.bgzf_host(files, bgzf.query(tid + 1, 1000, 2000))  # first record on another contig: nothing is read
    assert len(empty) == 0
    # a read longer than the device capacity inside the region: capacity error, before anything runs
    long_rec = bgzf.bam_record(tid, mid, 60, 99, b"long", [(200 << 4) | 0], bytes([0x11]) * 100, 200, bytes([30]) * 200, b"")
    stream = bgzf.bam_header(cases.REFS) + long_rec
    raw, blocks = bgzf.bgzf_compress(stream)
    f = bgzf.HostBgzfFiles([(raw, [(bgzf.voffset_of(blocks, len(stream) - len(long_rec)), len(raw) << 16, True)], 0, 0)],
                           [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]])
    with pytest.raises(engine.GtbError) as e:
        engine.bgzf_host(f, bgzf.query(tid, beg, end))
    assert e.value.code == -4


def test_replayed_merge_equals_device_order_when_nothing_is_ambiguous(monkeypatch):
    """Single file, no position with more than 16 records: the device order (exact duplicates in reverse file order) IS what
    the replay of std::sort + the heap gives."""
    rd, bam = cases.fixture_batch(SMALL[1])
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end)
    assert len(made) == 1
    files = cases.whole_file_segments(made)
    a = engine.bgzf_host(files, bgzf.query(tid, beg, end))
    monkeypatch.setenv("GTB_BGZF_FORCE_MERGE", "1")
    b = engine.bgzf_host(files, bgzf.query(tid, beg, end))
    cases.assert_batches_equal(a, b, "forced merge")


def _two_contig_records(order):
    recs = []
    for k, (tid, pos) in enumerate(order):
        seq = bytes([0x12 + (k % 7)] * 30)
        recs.append(bgzf.bam_record(tid, pos, 60, 0 if tid >= 0 else 4, b"w%03d" % k, [(60 << 4) | 0] if tid >= 0 else [], seq, 60,
                                    bytes([30]) * 60, b"ASC\x05", -1, -1, 0))
    return recs


def test_whole_file_reading_spans_contigs_and_needs_sorted_files():
    """whole_file = the reader without a region (sam_read1): every record of every contig, merged by (contig, position, length,
    sequence).  A file that is not in coordinate order -- unmapped reads behind the mapped ones included, contig -1 sorts first
    in the reference's heap -- is declined: only for sorted files is the heap's output the sorted order."""
    a = _two_contig_records([(0, 100), (0, 100), (0, 250), (1, 5), (1, 90)])
    b = _two_contig_records([(0, 90), (0, 250), (1, 5), (1, 400)])
    made = []
    for fi, recs in enumerate((a, b)):
        hdr = bgzf.bam_header(cases.REFS)
        stream = hdr + b"".join(recs)
        raw, blocks = bgzf.bgzf_compress(stream, 300)
        made.append((raw, stream, blocks, len(hdr), fi, fi))
    files = cases.whole_file_segments(made)
    got = engine.bgzf_host(files, bgzf.query(0, 0, 0, whole_file=True))
    assert len(got) == 9
    assert list(zip(got.core["tid"], got.core["pos"])) == sorted(zip(got.core["tid"], got.core["pos"]))
    per_file = [[bam_oracle.ParsedRecord(r[4:]) for r in recs] for recs in (a, b)]
    rows = bam_oracle.expected_pool_records(per_file, 3840, False)
    assert [(r.tid, r.pos) for _, r in rows] == list(zip(got.core["tid"], got.core["pos"]))
    assert [fi for fi, _ in rows if True][:1] == [int(got.sample[0])]
    # a region query on the same files stops at the first record of the next contig
    reg = engine.bgzf_host(files, bgzf.query(0, 0, 1000))
    assert len(reg) == 5 and set(reg.core["tid"]) == {0}
    for bad_order in ([(0, 100), (0, 50)], [(1, 5), (0, 90)], [(0, 100), (-1, -1)]):
        hdr = bgzf.bam_header(cases.REFS)
        stream = hdr + b"".join(_two_contig_records(bad_order))
        raw, blocks = bgzf.bgzf_compress(stream)
        f = cases.whole_file_segments([(raw, stream, blocks, len(hdr), 0, 0)])
        with pytest.raises(engine.GtbError) as e:
            engine.bgzf_host(f, bgzf.query(0, 0, 0, whole_file=True))
        assert e.value.code == -5 and "coordinate order" in str(e.value)


def _one_file(stream_records, block, level, strategy):
    hdr = bgzf.bam_header(cases.REFS)
    stream = hdr + b"".join(stream_records)
    raw, blocks = bgzf.bgzf_compress(stream, block, level, strategy)
    offs = [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]]
    u = bgzf.voffset_of(blocks, len(hdr))
    return raw, stream, blocks, len(hdr), offs, u


def test_inflate_fuzz_random_streams_and_corruptions():
    """Seeded fuzz of the decoder: (1) random payloads of every entropy (runs, text-like, noise) under random zlib settings
    inflate to their source; (2) random bit flips and truncations anywhere in the compressed bytes either decode to something
    the CRC / ISIZE check rejects or are reported as a stream error -- never a crash, a hang or a different record batch
    accepted silently (a flip may hit bytes that do not matter, e.g. the gzip MTIME field: then the batch must be unchanged)."""
    rng = np.random.default_rng(2024)
    recs = []
    for k in range(400):
        kind = k % 4
        l = int(rng.integers(40, 150))
        if kind == 0:
            seq = bytes([0x11]) * ((l + 1) // 2)
        elif kind == 1:
            seq = bytes(rng.choice([0x12, 0x48, 0x84, 0x21], (l + 1) // 2).astype(np.uint8))
        else:
            seq = bytes(rng.integers(0, 256, (l + 1) // 2, dtype=np.uint8))
        qual = bytes(rng.integers(2, 41, l, dtype=np.uint8)) if kind != 0 else bytes([30]) * l
        recs.append(bgzf.bam_record(0, 1000 + k // 2, 60, 0, b"fz%04d" % k, [(l << 4) | 0], seq, l, qual, b"ASC\x07XSC\x01NMC\x00"))
    q = bgzf.query(0, 0, 0, whole_file=True)
    want = None
    for trial in range(12):
        level = int(rng.integers(0, 10))
        strategy = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        block = int(rng.choice([700, 4096, 20000, 0xFF00]))
        raw, stream, blocks, hlen, offs, u = _one_file(recs, block, level, strategy)
        files = bgzf.HostBgzfFiles([(raw, [(u, len(raw) << 16, True)], 0, 0)], offs)
        got, infl = engine.bgzf_host(files, q, want_inflated=True)
        first = [b for b in blocks if b[1] <= hlen < b[1] + b[2]][0]
        assert bytes(infl[:len(stream) - first[1]]) == stream[first[1]:], (level, strategy, block)
        assert len(got) == len(recs)
        if want is None:
            want = got
        else:
            cases.assert_batches_equal(got, want, "same records under every compression")
    raw, stream, blocks, hlen, offs, u = _one_file(recs, 4096, 6, zlib.Z_DEFAULT_STRATEGY)
    n_rejected = 0
    for trial in range(300):
        bad = bytearray(raw)
        if trial % 5 == 4:
            cut = int(rng.integers(blocks[0][0] + 20, len(raw) - 30))
            bad = bad[:cut]
        else:
            for _ in range(int(rng.integers(1, 4))):
                at = int(rng.integers(blocks[0][0], len(raw) - 28))
                bad[at] ^= 1 << int(rng.integers(0, 8))
        files = bgzf.HostBgzfFiles([(bytes(bad), [(u, len(raw) << 16, True)], 0, 0)], offs)
        try:
            got = engine.bgzf_host(files, q)
        except engine.GtbError as e:
            assert e.code in (-5, -4, -1), e
            n_rejected += 1
            continue
        cases.assert_batches_equal(got, want, "a flip that was accepted must not have changed anything")
    assert n_rejected > 200


def test_submit_bgzf_has_no_cpu_fallback():
    """The emulation above is a parity tap: the product entry refuses to run without a device."""
    c = engine.Context(device=-1)
    try:
        rd, bam = cases.fixture_batch(SMALL[0])
        tid, beg, end = cases.region_of(bam)
        files = cases.whole_file_segments(cases.pool_files(bam, tid, beg, end))
        with pytest.raises(engine.GtbError) as e:
            c.submit_bgzf(1, files, bgzf.query(tid, beg, end))
        assert e.value.code == -2
    finally:
        c.close()
