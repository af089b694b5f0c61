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
        assert st.kernel_launches == 4  # probe, chain, slow, score
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
