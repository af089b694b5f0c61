"""gtb_submit_bgzf on the GPU: compressed BGZF segments in, genotyped pool out.  The record batch the kernels
build must equal the CPU run of the same source functions (gtb_debug_bgzf_host, pinned to zlib, to a plain-Python iterator and
to the reference's own reader by the CPU suite) byte for byte, and the pool's accumulators must equal the golden ones."""
import os
import shutil
import tempfile
import zlib

import numpy as np
import pytest

import bgzf_cases as cases
import compare
from conftest import fixture_prefixes
from graphtyper_b200 import abi, bgzf, engine, gtba

pytestmark = pytest.mark.gpu
ALL = fixture_prefixes(include_big=True)
PICK = [p for p in ALL if os.path.basename(p).split(".")[0] in
        ("tiny", "mini_stress", "mini_complex", "mini_sv", "mini_r100", "stress", "sv30k", "r60k")]


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(device=0)
    yield c
    c.close()


def n_samples_of(rd):
    return int(rd["sample"].max()) + 1 if len(rd["sample"]) else 1


@pytest.mark.parametrize("pre", PICK, ids=[os.path.basename(p) for p in PICK])
def test_bgzf_submit_matches_golden_accumulators(pre, ctx):
    """One BAM per (sample, read group) with decoy records, whole-file segments: device record batch == CPU emulation,
    accumulators == the reference's."""
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd, bam = cases.fixture_batch(pre)
    tid, beg, end = cases.region_of(bam)
    sv = g.is_sv_graph
    made = cases.pool_files(bam, tid, beg, end, block_size=0xFF00)
    files = cases.whole_file_segments(made)
    q = bgzf.query(tid, beg, end, sv=sv)
    want_batch = engine.bgzf_host(files, q)
    ctx.region_begin(40, g)
    try:
        ctx.pool_begin(40, n_samples_of(rd))
        st = ctx.submit_bgzf(40, files, q)
        assert st.n_records == len(bam)
        cases.assert_batches_equal(ctx.debug_bgzf_records(), want_batch, "device vs CPU emulation")
        assert ctx.debug_bgzf_stitched() == engine.bgzf_host_stitched()  # these files cut records anywhere: mostly serial walks
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), ctx.pool_finish(40).as_dict(), "bgzf")
    finally:
        ctx.region_end(40)


@pytest.mark.parametrize("level,strategy,block,n_chunks", [(0, zlib.Z_DEFAULT_STRATEGY, 0x8000, 1), (6, zlib.Z_FIXED, 0x3000, 3),
                                                           (9, zlib.Z_DEFAULT_STRATEGY, 0x800, 4), (1, zlib.Z_HUFFMAN_ONLY, 0xFF00, 2)])
def test_bgzf_block_kinds_and_chunked_files(ctx, level, strategy, block, n_chunks, monkeypatch):
    """Stored / fixed / dynamic / match-free DEFLATE blocks, tiny and full-size BGZF blocks, several non-adjacent chunks per
    file; and with the merge replay forced (GTB_BGZF_FORCE_MERGE) the batch must not change."""
    pre = [p for p in ALL if "mini_stress" in p][0]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd, bam = cases.fixture_batch(pre)
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end, block, level, strategy)
    if n_chunks == 1:
        files = cases.whole_file_segments(made)
    else:
        files, _ = cases.chunked_segments(made, n_chunks, gap=6)
    q = bgzf.query(tid, beg, end)
    want = engine.bgzf_host(files, q)
    ctx.region_begin(41, g)
    try:
        ctx.pool_begin(41, n_samples_of(rd))
        ctx.submit_bgzf(41, files, q)
        got = ctx.debug_bgzf_records()
        cases.assert_batches_equal(got, want, "device vs CPU emulation")
        monkeypatch.setenv("GTB_BGZF_FORCE_MERGE", "1")
        ctx.pool_reset(41)
        ctx.submit_bgzf(41, files, q)
        cases.assert_batches_equal(ctx.debug_bgzf_records(), want, "forced merge replay")
        # the same records through the record entry give the same accumulators
        acc = ctx.pool_finish(41).as_dict()
        ctx.pool_reset(41)
        ctx.submit_bam(41, want)
        compare.compare_accum(ctx.pool_finish(41).as_dict(), acc, "bgzf vs record entry")
    finally:
        ctx.region_end(41)


def test_bgzf_errors_leave_the_pool_untouched(ctx):
    pre = [p for p in ALL if "tiny" in p][0]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd, bam = cases.fixture_batch(pre)
    tid, beg, end = cases.region_of(bam)
    made = cases.pool_files(bam, tid, beg, end)
    raw, stream, blocks, hlen, s, gq = made[0]
    offs = [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]]
    u = bgzf.voffset_of(blocks, hlen)
    bad = bytearray(raw)
    bad[blocks[0][0] + 60] ^= 0x5A
    long_rec = bgzf.bam_record(tid, beg + 5, 60, 99, b"long", [(200 << 4) | 0], bytes([0x11]) * 100, 200, bytes([30]) * 200, b"")
    lstream = bgzf.bam_header(cases.REFS) + long_rec
    lraw, lblocks = bgzf.bgzf_compress(lstream)
    ctx.region_begin(42, g)
    try:
        ctx.pool_begin(42, 1)
        with pytest.raises(engine.GtbError) as e:
            ctx.submit_bgzf(42, bgzf.HostBgzfFiles([(bytes(bad), [(u, len(raw) << 16, True)], 0, 0)], offs), bgzf.query(tid, beg, end))
        assert e.value.code == -5
        with pytest.raises(engine.GtbError) as e:
            ctx.submit_bgzf(42, bgzf.HostBgzfFiles([(lraw, [(bgzf.voffset_of(lblocks, len(lstream) - len(long_rec)), len(lraw) << 16, True)], 0, 0)],
                                                   [[b[0] for b in lblocks] + [len(lraw) - 28, len(lraw)]]), bgzf.query(tid, beg, end))
        assert e.value.code == -4
        # nothing was added, the pool is not poisoned: the real submit gives the golden accumulators
        ctx.submit_bgzf(42, cases.whole_file_segments(made), bgzf.query(tid, beg, end))
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), ctx.pool_finish(42).as_dict(), "after errors")
        # an empty pool: files without a record in the region
        ctx.pool_reset(42)
        st = ctx.submit_bgzf(42, cases.whole_file_segments(made), bgzf.query(tid + 1, 0, 100))
        assert st.n_records == 0
    finally:
        ctx.region_end(42)


def test_bgzf_against_the_reference_reader_on_htslib_files(ctx):
    """BAM files written by the vendored htslib, read by the compiled reference (gt_probe) and by the device: same record batch
    in the same order (exact duplicates across files and a 45-record pile-up included), same accumulators."""
    import test_bgzf_vs_reference as R
    if not os.path.exists(os.path.join(R.BIN, "gt_probe")):
        pytest.skip("compiled reference not available")
    for name in ("pileup_and_flags", "six_files_many_ties"):
        tmp = tempfile.mkdtemp(prefix="gtb_bgzf_gpu_")
        try:
            pre, want, files = R.build_case(name, tmp)
            g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
            ctx.region_begin(43, g)
            try:
                ctx.pool_begin(43, int(want.sample.max()) + 1)
                ctx.submit_bgzf(43, files, bgzf.query(0, 0, 0, whole_file=True))  # as gt_probe read them: no region
                cases.assert_batches_equal(ctx.debug_bgzf_records(), want, name)
                assert ctx.debug_bgzf_stitched() == files.n_files  # record boundaries from the per-block walks
                compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), ctx.pool_finish(43).as_dict(), name)
            finally:
                ctx.region_end(43)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
