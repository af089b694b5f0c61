"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against (a) the golden vectors of the compiled
reference and (b) the oracle on the same inputs.  Bit-exact: seeds, paths, accumulators, PL/GT/GQ."""
import os

import numpy as np
import pytest

import compare
from conftest import fixture_prefixes
from graphtyper_b200 import abi, engine, gtba

ALL = fixture_prefixes(include_big=True)
pytestmark = pytest.mark.gpu


def n_samples_of(rd):
    return len(rd["sample_names"].tobytes().split(b"\n")) - 1


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(device=0)
    yield c
    c.close()


@pytest.mark.parametrize("pre", ALL, ids=[os.path.basename(p) for p in ALL])
def test_cuda_matches_reference(pre, ctx):
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ns = n_samples_of(rd)
    ctx.region_begin(7, g)
    try:
        compare.compare_index(gtba.load(pre + ".index.gtba"), ctx.index_export(7))
        ctx.pool_begin(7, ns)
        ctx.debug_enable(True)
        st = ctx.submit(7, b)
        assert st.n_capacity_overflow == 0
        assert st.kernel_launches == 11  # prep_flags, scan, prep_fill, probe, chain, chain_general, slow, score | slow, huge, score
        assert compare.compare_seeds(compare.probe_seeds(rd), ctx.debug_seeds(7), "cuda") > 0
        compare.compare_paths(compare.probe_paths(rd), ctx.debug_paths(7), "cuda")
        acc = ctx.pool_finish(7)
        pa = gtba.load(pre + ".accum.gtba")
        compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), "cuda")
        ph, gt, gq = ctx.calls(acc)
        assert np.array_equal(ph, pa["call_phred"])
        assert np.array_equal(gt, pa["call_gt"])
        assert np.array_equal(gq, pa["call_gq"])
        assert not acc.saturated.any()
    finally:
        ctx.debug_enable(False)
        ctx.region_end(7)


@pytest.mark.parametrize("pre", ALL, ids=[os.path.basename(p) for p in ALL])
def test_cuda_connections_match_reference(pre, ctx):
    """Phasing connections (HapSample::connections) built by score_kernel<true> and the derived `ph` map against the
    compiled reference; the accumulators must not change when connections are collected."""
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ns = n_samples_of(rd)
    pa = gtba.load(pre + ".accum.gtba")
    ctx.region_begin(9, g)
    ctx.set_connections(1)
    try:
        ctx.pool_begin(9, ns)
        ctx.submit(9, b)
        acc = ctx.pool_finish(9)
        conn = ctx.connections(9)
        compare.compare_connections(compare.probe_connections(pa), abi.connections_as_table(conn), "cuda")
        compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), "cuda+conn")
        if "ph_tuples" in pa:
            got = abi.phase_as_table(ctx.phase_support(acc, conn))
            assert np.array_equal(pa["ph_tuples"].reshape(-1, 5).astype(np.uint32), got)
        # reset clears the table; a second pass gives the same entries (idempotence of reset + submit)
        ctx.pool_reset(9)
        ctx.submit(9, b)
        assert np.array_equal(abi.connections_as_table(ctx.connections(9)), abi.connections_as_table(conn))
    finally:
        ctx.set_connections(0)
        ctx.region_end(9)


@pytest.mark.parametrize("pre", ALL, ids=[os.path.basename(p) for p in ALL])
def test_cuda_record_parsing_matches_reference(pre, ctx, oracle_lib):
    """gtb_submit_bam_records: raw htslib records (as the probe dumped them from the reference's own loop) are parsed,
    de-duplicated and paired on the device; the derived columns equal the reference's (score_diff, duplicates) and the
    oracle's (mates, leftovers), and the accumulators equal the golden ones."""
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    bam = abi.HostBamBatch.from_probe(rd)
    ns = n_samples_of(rd)
    want = oracle_lib.parse_bam(bam, is_sv=g.is_sv_graph)
    ctx.region_begin(11, g)
    try:
        ctx.pool_begin(11, ns)
        st = ctx.submit_bam(11, bam)
        assert st.n_records == len(bam) and st.n_capacity_overflow == 0
        assert st.n_alignments == int((rd["isdup"] == 0).sum())
        col = ctx.debug_bam_columns(len(bam))
        assert np.array_equal(col["score_diff"], rd["score_diff"])
        assert np.array_equal(col["dup_of"] >= 0, rd["isdup"] != 0)
        for k in ("seq4", "lseq", "flag", "mapq", "isize", "same_tid", "score_diff", "mate", "dup_of", "leftover"):
            assert np.array_equal(col[k], getattr(want, k)), k
        acc = ctx.pool_finish(11)
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "cuda(bam)")
    finally:
        ctx.region_end(11)


@pytest.mark.parametrize("bits", [3, 9])
def test_record_parsing_survives_name_hash_collisions(ctx, oracle_lib, bits, monkeypatch):
    """Different read names that share one 64-bit hash land in one run of the sorted hashes; the pairing kernel replays the
    read-name map once per distinct name of a run.  Forced here by keeping only `bits` bits of the hash (2^3 = 8 runs for
    thousands of names): mates, leftovers and accumulators must not change."""
    for pre in [p for p in ALL if os.path.basename(p).startswith(("mini_stress", "mini_sv"))]:
        g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
        rd = gtba.load(pre + ".reads.gtba")
        bam = abi.HostBamBatch.from_probe(rd)
        want = oracle_lib.parse_bam(bam, is_sv=g.is_sv_graph)
        monkeypatch.setenv("GTB_BAM_HASH_BITS", str(bits))
        ctx.region_begin(12, g)
        try:
            ctx.pool_begin(12, n_samples_of(rd))
            ctx.submit_bam(12, bam)
            col = ctx.debug_bam_columns(len(bam))
            for k in ("mate", "dup_of", "leftover"):
                assert np.array_equal(col[k], getattr(want, k)), (k, bits)
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), ctx.pool_finish(12).as_dict(), "collisions")
        finally:
            monkeypatch.delenv("GTB_BAM_HASH_BITS")
            ctx.region_end(12)


def test_record_parsing_of_several_pools_in_one_call(ctx):
    """gtb_submit_bam_records_multi: pairing and the duplicate shortcut stay inside each pool; same accumulators as one call
    per region.  Also: records built from the synthetic generator (abi.bam_batch_from_readsets) give the same accumulators as
    the column batch of the same pool (abi.batch_from_readsets)."""
    pres = [p for p in ALL if os.path.basename(p).startswith("mini_") or os.path.basename(p).startswith("tiny")]
    ids = list(range(400, 400 + len(pres)))
    graphs, bams, ns = [], [], []
    for pre in pres:
        graphs.append(abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba")))
        rd = gtba.load(pre + ".reads.gtba")
        bams.append(abi.HostBamBatch.from_probe(rd))
        ns.append(n_samples_of(rd))
    ctx.region_begin_multi(ids, graphs)
    try:
        for k, n in zip(ids, ns):
            ctx.pool_begin(k, n)
        st = ctx.submit_bam_multi(ids, bams)
        assert st.n_records == sum(len(b) for b in bams)
        for pre, acc in zip(pres, ctx.pool_finish_multi(ids)):
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "cuda(bam multi) " + pre)
    finally:
        for k in ids:
            ctx.region_end(k)
    from graphtyper_b200 import graph_build, synth
    ds = synth.make_dataset(length=6000, n_sites=60, n_samples=2, seed=77, coverage=10, err=0.005, unpaired_rate=0.05)
    g = graph_build.build_region_graph(ds.ref, ds.sites, 1, 6000, pad=0)
    cols = abi.batch_from_readsets(ds.reads)
    bam = abi.bam_batch_from_readsets(ds.reads)
    ctx.region_begin(450, g)
    try:
        ctx.pool_begin(450, 2)
        ctx.submit(450, cols)
        a = ctx.pool_finish(450).as_dict()
        ctx.pool_reset(450)
        ctx.submit_bam(450, bam)
        b = ctx.pool_finish(450).as_dict()
        compare.compare_accum(a, b, "columns vs records")
        assert a["read_strand"].sum() > 0
    finally:
        ctx.region_end(450)


def test_record_parsing_rejects_bad_records(ctx):
    _ref, g = _plain_graph()
    ctx.region_begin(12, g)
    try:
        ctx.pool_begin(12, 1)

        def one(l_qseq, data, l_qname=4):
            core = np.zeros(1, abi.BAM_CORE_DTYPE)
            core["l_qseq"], core["l_qname"] = l_qseq, l_qname
            return abi.HostBamBatch(core, np.frombuffer(data, np.uint8), np.array([0, len(data)], np.uint64), np.zeros(1, np.int32),
                                    np.zeros(1, np.int32))
        with pytest.raises(engine.GtbError) as e:   # sequence + qualities do not fit the data block
            ctx.submit_bam(12, one(100, b"r1\0\0" + b"\x11" * 20))
        assert e.value.code == -1
        with pytest.raises(engine.GtbError) as e:   # a failed submit leaves the pool undefined: refused until it is reset
            ctx.submit_bam(12, one(100, b"r1\0\0" + b"\x11" * 50 + b"I" * 100))
        assert e.value.code == -3
        ctx.pool_reset(12)
        with pytest.raises(engine.GtbError) as e:   # longer than the reference's MAX_READ_LENGTH
            ctx.submit_bam(12, one(200, b"r1\0\0" + b"\x11" * 100 + b"I" * 200))
        assert e.value.code == -4
        ctx.pool_reset(12)
        st = ctx.submit_bam(12, abi.HostBamBatch(np.zeros(0, abi.BAM_CORE_DTYPE), np.zeros(0, np.uint8), np.zeros(1, np.uint64),
                                                 np.zeros(0, np.int32), np.zeros(0, np.int32)))
        assert st.n_records == 0
    finally:
        ctx.region_end(12)


def test_connection_table_grows_and_replay_doubles(ctx):
    """A 2-slots-per-record budget forces table growth (re-insert on the device) between submits; the result is unchanged.
    Replaying the resident batch (the last shard) only ever adds counts (monotonicity)."""
    pre = [p for p in ALL if os.path.basename(p) in ("stress.r0", "mini_stress.r0")][-1]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ns = n_samples_of(rd)
    pa = gtba.load(pre + ".accum.gtba")
    want = compare.probe_connections(pa)
    ctx.region_begin(9, g)
    ctx.set_connections(2)
    try:
        ctx.pool_begin(9, ns)
        for shard in abi.shard_batch(b, 3):  # three submits into the same pool: the table grows between them
            ctx.submit(9, shard)
        got = abi.connections_as_table(ctx.connections(9))
        compare.compare_connections(want, got, "cuda small budget")
        ctx.replay()
        again = abi.connections_as_table(ctx.connections(9))
        assert np.array_equal(again[:, :5], got[:, :5]) and (again[:, 5] >= got[:, 5]).all() and again[:, 5].sum() > got[:, 5].sum()
    finally:
        ctx.set_connections(0)
        ctx.region_end(9)


def test_region_batched_submit_equals_separate(ctx):
    """Several regions in ONE launch give the same accumulators as one launch per region."""
    pres = ALL[:3]
    graphs, batches, ns = [], [], []
    for k, pre in enumerate(pres):
        graphs.append(abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba")))
        rd = gtba.load(pre + ".reads.gtba")
        batches.append(abi.batch_from_probe(rd))
        ns.append(n_samples_of(rd))
        ctx.region_begin(100 + k, graphs[-1])
        ctx.pool_begin(100 + k, ns[-1])
    try:
        st = ctx.submit_multi([100 + k for k in range(len(pres))], batches)
        assert st.n_records == sum(len(b) for b in batches)
        for k, pre in enumerate(pres):
            acc = ctx.pool_finish(100 + k)
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "cuda-multi")
    finally:
        for k in range(len(pres)):
            ctx.region_end(100 + k)


@pytest.mark.parametrize("n_chunks", [2, 3, 4, 8])
def test_chunked_pipeline_equals_golden(ctx, n_chunks):
    """The double-buffered chunk pipeline (copy stream + compute stream) gives the same accumulators."""
    pres = [p for p in ALL if "sv" not in os.path.basename(p)][:9]
    graphs, batches, ns = [], [], []
    ids = list(range(300, 300 + len(pres)))
    for k, pre in zip(ids, pres):
        graphs.append(abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba")))
        rd = gtba.load(pre + ".reads.gtba")
        batches.append(abi.batch_from_probe(rd))
        ns.append(n_samples_of(rd))
    ctx.region_begin_multi(ids, graphs)
    try:
        for k, n in zip(ids, ns):
            ctx.pool_begin(k, n)
        ctx.set_chunks(n_chunks)
        st = ctx.submit_multi(ids, batches)
        assert st.n_records == sum(len(b) for b in batches)
        accs = ctx.pool_finish_multi(ids)
        for pre, acc in zip(pres, accs):
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), f"chunks={n_chunks}")
        # replaying a chunked batch doubles the sums
        ctx.replay()
        a2 = ctx.pool_finish(ids[0]).as_dict()
        assert np.array_equal(a2["read_strand"].astype(np.int64), 2 * accs[0].as_dict()["read_strand"].astype(np.int64))
    finally:
        ctx.set_chunks(0)
        for k in ids:
            ctx.region_end(k)


def test_replay_accumulates_linearly(ctx):
    """Size-independent property: accumulators are additive -- replaying the resident batch doubles every sum."""
    pre = ALL[0]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ctx.region_begin(9, g)
    try:
        ctx.pool_begin(9, n_samples_of(rd))
        ctx.submit(9, b)
        a1 = ctx.pool_finish(9).as_dict()
        ctx.replay()
        a2 = ctx.pool_finish(9).as_dict()
        for k in ("log_score", "gt_coverage", "pa_mapq_squared", "read_strand", "vs_mapq_squared"):
            assert np.array_equal(a2[k].astype(np.int64), 2 * a1[k].astype(np.int64)), k
        ctx.pool_reset(9)
        ctx.replay()
        a3 = ctx.pool_finish(9).as_dict()
        for k in ("log_score", "gt_coverage", "read_strand"):
            assert np.array_equal(a3[k], a1[k]), k
    finally:
        ctx.region_end(9)


def test_full_config2_workload_matches_oracle(ctx, oracle_lib):
    """BASELINE configs[1] at full size (1 Mb, 10k sites, 2e5 records, 20 regions, one region-batched submit):
    accumulators of every region bit-identical to the oracle; PL/GT/GQ identical."""
    import bench
    ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
    ids = list(range(200, 200 + len(graphs)))
    for k, g in zip(ids, graphs):
        ctx.region_begin(k, g)
        ctx.pool_begin(k, 1)
    try:
        st = ctx.submit_multi(ids, batches)
        assert st.n_records == 200000 and st.n_capacity_overflow == 0
        accs = ctx.pool_finish_multi(ids)
        n_scored = 0
        for g, b, acc in zip(graphs, batches, accs):
            h = oracle_lib.index_build(g)
            r = oracle_lib.pool_run(g, h, 1, b, tap=False)
            ref_acc = oracle_lib.result_accum(r, 1)
            n_scored += oracle_lib.result_stats(r).n_pairs_scored
            compare.compare_accum({k: v for k, v in ref_acc.as_dict().items() if k != "saturated"}, acc.as_dict(), "cuda-1Mb")
            a, bb, cc = ctx.calls(acc)
            a2, b2, c2 = oracle_lib.calls(ref_acc)
            assert np.array_equal(a, a2) and np.array_equal(bb, b2) and np.array_equal(cc, c2)
            oracle_lib.result_free(r)
            oracle_lib.index_free(h)
        assert st.n_pairs_scored == n_scored
    finally:
        for k in ids:
            ctx.region_end(k)


def _large_region_case(length, n_sites, seed, coverage):
    from graphtyper_b200 import graph_build, synth
    ref = synth.make_reference(length, seed)
    sites = synth.make_sites(ref, n_sites, seed + 1)
    gts = synth.make_genotypes(n_sites, 1, seed + 2)
    rs = synth.simulate_reads(ref, sites, gts[0], "SAMP1", seed + 3, coverage=coverage)
    g = graph_build.build_region_graph(ref, sites, 1, length, pad=0)
    return g, abi.batch_from_readsets([rs])


@pytest.mark.parametrize("length,n_sites,expect_log2cap", [(150_000, 1500, 20), (420_000, 3000, 21)],
                         ids=["fold4_shared_filter", "global_bitmap"])
def test_probe_paths_of_large_regions_match_oracle(ctx, oracle_lib, length, n_sites, expect_log2cap):
    """probe_kernel stages a region's presence bitmap in shared memory folded 1x / 2x / 4x and falls back to the global
    bitmap beyond that (more than ~2.6e5 distinct k-mers, e.g. the 1.2 Mb regions of genotype_sv).  The fixtures only reach
    the 1x / 2x cases; here one region per remaining case, against the oracle."""
    g, b = _large_region_case(length, n_sites, seed=301 + expect_log2cap, coverage=4)
    h = oracle_lib.index_build(g)
    n_keys = len(oracle_lib.index_export(h)["keys"])
    assert (1 << (expect_log2cap - 1)) < 4 * n_keys + 2 <= (1 << expect_log2cap), n_keys   # table capacity = load <= 0.25
    r = oracle_lib.pool_run(g, h, 1, b, tap=False)
    want = {k: v for k, v in oracle_lib.result_accum(r, 1).as_dict().items() if k != "saturated"}
    ctx.region_begin(30, g)
    try:
        ctx.pool_begin(30, 1)
        st = ctx.submit(30, b)
        assert st.n_capacity_overflow == 0 and st.n_pairs_scored == oracle_lib.result_stats(r).n_pairs_scored
        compare.compare_accum(want, ctx.pool_finish(30).as_dict(), "large region")
    finally:
        ctx.region_end(30)
        oracle_lib.result_free(r)
        oracle_lib.index_free(h)


def test_probe_global_bitmap_path_on_fixtures(ctx, monkeypatch):
    """The same fixtures with the shared-memory filter switched off: every region probes the global bitmap."""
    monkeypatch.setenv("GTB_PROBE_FILTER_FOLD", "-1")
    for pre in ALL[:6]:
        g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
        rd = gtba.load(pre + ".reads.gtba")
        ctx.region_begin(31, g)
        try:
            ctx.pool_begin(31, n_samples_of(rd))
            ctx.submit(31, abi.batch_from_probe(rd))
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), ctx.pool_finish(31).as_dict(), "global bitmap")
        finally:
            ctx.region_end(31)


def test_pinned_inputs_take_the_direct_dma_path_with_equal_results(ctx):
    """Columns in page-locked memory (gtb_host_alloc) skip the staging copy; results must not change."""
    pre = ALL[1]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ns = n_samples_of(rd)
    pb, arena = engine.pin_batches([b])
    ctx.region_begin(11, g)
    try:
        ctx.pool_begin(11, ns)
        ctx.submit(11, pb[0])
        acc = ctx.pool_finish(11)
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "cuda-pinned")
    finally:
        ctx.region_end(11)
        arena.close()


@pytest.mark.parametrize("skew,n_chunks", [(0, 1), (1, 1), (3, 2), (5, 0)])
def test_zero_copy_columns_equal_staged_columns(ctx, skew, n_chunks, monkeypatch):
    """gather_columns_kernel: with every column page-locked the device reads lengths, flags, links ... straight from the
    caller's arrays (links rebased to chunk-global indices on the device).  All fixtures in ONE submit, so that record bases
    are odd and destinations unaligned; `skew` shifts every source column off its 16-byte alignment; SV fixtures carry the
    leftover column.  The submit must report the gather launch, and every pool must equal the golden accumulators."""
    pres = [p for p in ALL[:9] if n_chunks == 1 or "sv" not in os.path.basename(p)]
    ids = list(range(700, 700 + len(pres)))
    graphs, batches, ns = [], [], []
    for pre in pres:
        graphs.append(abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba")))
        rd = gtba.load(pre + ".reads.gtba")
        batches.append(abi.batch_from_probe(rd))
        ns.append(n_samples_of(rd))
    pinned, arena = engine.pin_batches(batches, skew=skew)
    monkeypatch.setenv("GTB_ZERO_COPY", "1")  # the default only with several ranks per host (LOCAL_WORLD_SIZE > 1)
    ctx.region_begin_multi(ids, graphs)
    try:
        for k, n in zip(ids, ns):
            ctx.pool_begin(k, n)
        ctx.set_chunks(n_chunks)
        st = ctx.submit_multi(ids, pinned)
        launches_zero_copy = st.kernel_launches
        assert st.n_records == sum(len(b) for b in batches)
        for pre, acc in zip(pres, ctx.pool_finish_multi(ids)):
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), f"zero-copy skew={skew}")
        # the same submit from pageable memory takes the staged path: fewer launches (no gather kernel), same results
        ctx.pool_reset_multi(ids)
        st = ctx.submit_multi(ids, batches)
        assert st.kernel_launches < launches_zero_copy
        for pre, acc in zip(pres, ctx.pool_finish_multi(ids)):
            compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "staged")
    finally:
        ctx.set_chunks(0)
        for k in ids:
            ctx.region_end(k)
        arena.close()


# ------------------------------------------------------------------------------------------------ edge cases
def _plain_graph(length=400, seed=3):
    """A region without any variant: a single ref node."""
    from graphtyper_b200 import graph_build, synth
    ref = synth.make_reference(length, seed)
    return ref, graph_build.build_region_graph(ref, [], 1, length, pad=0)


def _batch_from_seqs(seqs, flags, mates=None, lens=None):
    n = len(seqs)
    L = max(len(s) for s in seqs) if n else 0
    arr = np.full((n, max(L, 1)), ord("A"), np.uint8)
    for i, s in enumerate(seqs):
        arr[i, :len(s)] = np.frombuffer(s, np.uint8)
    seq4 = abi.pack_seq4(arr) if n else np.zeros((0, abi.SEQ_STRIDE), np.uint8)
    lseq = np.array([len(s) for s in seqs] if lens is None else lens, np.uint16)
    mate = np.full(n, -1, np.int32) if mates is None else np.array(mates, np.int32)
    return abi.HostBatch(seq4, lseq, np.array(flags, np.uint16), np.full(n, 60, np.uint8), np.zeros(n, np.int32),
                         np.ones(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32),
                         mate, np.full(n, -1, np.int32))


def test_empty_batch_and_variant_free_graph(ctx, oracle_lib):
    ref, g = _plain_graph()
    ctx.region_begin(21, g)
    try:
        ctx.pool_begin(21, 1)
        st = ctx.submit(21, _batch_from_seqs([], []))
        assert st.n_records == 0 and st.kernel_launches == 0
        # ragged lengths incl. reads below the 63-bp limit and at the 151/152 maximum; no bubbles -> nothing to score
        seqs = [bytes(ref[10:10 + n]) for n in (40, 62, 63, 100, 151, 152)]
        st = ctx.submit(21, _batch_from_seqs(seqs, [0] * len(seqs)))
        assert st.n_records == 6 and st.n_oriented == 4 and st.n_capacity_overflow == 0
        ctx.debug_enable(True)
        ctx.submit(21, _batch_from_seqs(seqs, [0] * len(seqs)))
        paths = ctx.debug_paths(21)
        ctx.debug_enable(False)
        assert paths["gp_npaths"].tolist() == [0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 1, 0]
        assert paths["gp_longest"][4::2].tolist() == [63, 100, 151, 152]
        acc = ctx.pool_finish(21)
        assert acc.n_bubbles == 0 and len(acc.log_score) == 0
    finally:
        ctx.region_end(21)


def test_input_errors_are_reported(ctx):
    ref, g = _plain_graph()
    ctx.region_begin(22, g)
    try:
        ctx.pool_begin(22, 1)
        s = bytes(ref[20:170])
        with pytest.raises(engine.GtbError) as e:  # both mates first-in-pair: the reference exits (hts_parallel_reader.cpp:306)
            ctx.submit(22, _batch_from_seqs([s, s], [65, 65], mates=[-1, 0]))
        assert e.value.code == -5
        with pytest.raises(engine.GtbError) as e:  # the pool is undefined after a failed submit: refused until it is reset
            ctx.submit(22, _batch_from_seqs([s], [0]))
        assert e.value.code == -3
        ctx.pool_reset(22)
        with pytest.raises(engine.GtbError) as e:  # longer than the 152-base device limit
            ctx.submit(22, _batch_from_seqs([s], [0], lens=[153]))
        assert e.value.code == -4
        ctx.pool_reset(22)
        with pytest.raises(engine.GtbError) as e:  # forward mate reference
            ctx.submit(22, _batch_from_seqs([s, s], [65, 129], mates=[1, -1]))
        assert e.value.code == -1
    finally:
        ctx.region_end(22)


def test_device_index_build_equals_host_builder_and_reference(ctx):
    """N2: the index built on the device (default) equals the host builder's and the reference's golden index -- keys,
    label lists and bucket order -- one region at a time and all regions of a call in one launch sequence."""
    pres = ALL
    graphs = [abi.HostGraph.from_gtba(gtba.load(p + ".graph.gtba")) for p in pres]
    ids = list(range(100, 100 + len(pres)))
    ctx.set_index_build(True)
    ctx.region_begin_multi(ids, graphs)
    dev = [ctx.index_export(i) for i in ids]
    for i in ids:
        ctx.region_end(i)
    ctx.set_index_build(False)
    try:
        ctx.region_begin_multi(ids, graphs)
        host = [ctx.index_export(i) for i in ids]
        for i in ids:
            ctx.region_end(i)
    finally:
        ctx.set_index_build(True)
    for pre, d, h in zip(pres, dev, host):
        compare.compare_index(gtba.load(pre + ".index.gtba"), d)
        for k in ("keys", "label_off", "labels"):
            assert np.array_equal(d[k], h[k]), (pre, k)


def test_host_built_index_still_drives_the_kernels(ctx):
    pre = ALL[0]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    pa = gtba.load(pre + ".accum.gtba")
    ctx.set_index_build(False)
    try:
        ctx.region_begin(7, g)
        ctx.pool_begin(7, n_samples_of(rd))
        ctx.submit(7, b)
        acc = ctx.pool_finish(7)
        compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), "cuda/host-index")
    finally:
        ctx.set_index_build(True)
        ctx.region_end(7)


def test_huge_tier_takes_what_the_shared_memory_state_cannot_hold(ctx):
    """Reads on 15-allele bubbles need more locations / paths than slow_kernel's 13 KB shared-memory state holds: they are
    re-queued for huge_kernel (global-memory state, the reference's own limits) instead of failing, and the result is
    still the reference's (test_cuda_matches_reference covers the equality on this fixture)."""
    pre = [p for p in ALL if os.path.basename(p).startswith("mini_complex")][0]
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    ctx.region_begin(7, g)
    try:
        ctx.pool_begin(7, n_samples_of(rd))
        st = ctx.submit(7, abi.batch_from_probe(rd))
        assert st.n_capacity_overflow == 0
        cnt = ctx.debug_counters()
        assert sum(cnt[12:24]) > 0, "expected slow_kernel -> huge_kernel re-queues on this fixture"
        acc = ctx.pool_finish(7)
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), "cuda/huge")
    finally:
        ctx.region_end(7)
