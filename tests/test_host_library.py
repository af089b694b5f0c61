"""CPU suite for the product library: it loads, exports every symbol include/gtb200.h declares, its host-side
index builder reproduces the reference's PHIndex (keys, labels, bucket order) on the golden fixtures, the host
finalisation reproduces PL/GT/GQ, and the compute path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import compare
from conftest import ROOT, fixture_prefixes
from graphtyper_b200 import abi, engine, gtba

SMALL = fixture_prefixes(include_big=False)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gtb200.h")).read()
    declared = set(re.findall(r"\b(gtb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = engine.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libgtb200.so does not export {name}"
    assert set(engine.EXPORTS) <= declared
    assert b"sm_100a" in lib.gtb_version()


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_host_index_matches_reference(pre):
    ctx = engine.Context(device=-1)
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    ctx.region_begin(0, g)
    compare.compare_index(gtba.load(pre + ".index.gtba"), ctx.index_export(0))
    ctx.region_end(0)
    ctx.close()


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_host_finalisation_matches_reference(pre):
    pa = gtba.load(pre + ".accum.gtba")
    ref = compare.probe_accum(pa)
    ns, nb = int(pa["meta"][0]), int(pa["meta"][1])
    acc = abi.HostAccumulators(nb, int(ref["score_off"][-1]), int(ref["cov_off"][-1]), ns)
    for k in ("n_alleles", "score_off", "cov_off", "log_score"):
        getattr(acc, k)[:] = ref[k]
    for k in ("gt_coverage", "ambiguous_depth", "ambiguous_depth_alt", "alt_proper_pair_depth"):
        getattr(acc, k)[:] = ref[k]
    ctx = engine.Context(device=-1)
    # the SampleCall of every bubble x sample: AD / MD / PP are the accumulators, RA-style totals derived (sample_call.cpp:34-61)
    assert np.array_equal(pa["call_cov"], ref["gt_coverage"])
    assert np.array_equal(pa["call_amb"], ref["ambiguous_depth"]) and np.array_equal(pa["call_altpp"], ref["alt_proper_pair_depth"])
    ref_total, alt_total = ctx.sample_depths(acc)
    assert np.array_equal(ref_total, pa["call_ref_total"]) and np.array_equal(alt_total, pa["call_alt_total"])
    ph, gt, gq = ctx.calls(acc)
    assert np.array_equal(ph, pa["call_phred"])
    assert np.array_equal(gt, pa["call_gt"])
    assert np.array_equal(gq, pa["call_gq"])
    if "ref_depth" not in pa:  # Variant::scan_calls only runs for non-SV graphs (hts_parallel_reader.cpp:1019-1023)
        var, allele, ratio = ctx.scan_calls(acc, ph)
        assert np.array_equal(var, pa["stats_var"])
        assert np.array_equal(allele, pa["stats_allele"])
        assert np.array_equal(ratio, pa["stats_allele_ratio"])
    ctx.close()


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_host_phase_support_matches_reference(pre):
    """gtb_phase_support (pure host function) on the reference's own connections + coverage = the reference's `ph` map."""
    pa = gtba.load(pre + ".accum.gtba")
    if "ph_tuples" not in pa:
        pytest.skip("SV fixture: the reference never writes haplotypes for SV graphs")
    ref = compare.probe_accum(pa)
    ns, nb = int(pa["meta"][0]), int(pa["meta"][1])
    acc = abi.HostAccumulators(nb, int(ref["score_off"][-1]), int(ref["cov_off"][-1]), ns)
    for k in ("bubble_id", "n_alleles", "score_off", "cov_off", "gt_coverage"):
        getattr(acc, k)[:] = ref[k]
    tab = compare.probe_connections(pa)
    conn = np.zeros(len(tab), abi.CONNECTION_DTYPE)
    for j, k in enumerate(("sample", "hap1", "allele1", "hap2", "allele2", "count")):
        conn[k] = tab[:, j]
    ctx = engine.Context(device=-1)
    got = abi.phase_as_table(ctx.phase_support(acc, conn))
    assert np.array_equal(pa["ph_tuples"].reshape(-1, 5).astype(np.uint32), got)
    ctx.close()


def test_sharded_connections_merge_to_unsharded(oracle_lib):
    """Additivity of the phasing connections under read sharding (mates and duplicate groups stay together,
    abi.shard_batch): the shards' lists, merged by gtb_merge_connections, equal the unsharded list -- and the reference's."""
    pre = [p for p in SMALL if "mini_stress" in p][0]
    O = oracle_lib
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
    h = O.index_build(g)
    ctx = engine.Context(device=-1)
    O.set_connections(True)
    try:
        merged = np.zeros(0, abi.CONNECTION_DTYPE)
        for shard in abi.shard_batch(abi.batch_from_probe(rd), 3):
            r = O.pool_run(g, h, ns, shard, tap=False)
            merged = ctx.merge_connections(merged, O.result_connections(r))
            O.result_free(r)
    finally:
        O.set_connections(False)
        O.index_free(h)
    want = compare.probe_connections(gtba.load(pre + ".accum.gtba"))
    compare.compare_connections(want, abi.connections_as_table(merged), "merged shards")
    # counters wrap at 2^16 like the reference's uint16: 0xFFFF + 1 disappears, 0xFFFF + 3 becomes 2
    one = np.zeros(2, abi.CONNECTION_DTYPE)
    one["hap2"], one["allele2"], one["count"] = [1, 2], [0, 1], [0xFFFF, 0xFFFF]
    two = one.copy()
    two["count"] = [1, 3]
    out = ctx.merge_connections(one, two)
    assert abi.connections_as_table(out).tolist() == [[0, 0, 0, 2, 1, 2]]
    ctx.close()


def test_no_cpu_fallback():
    ctx = engine.Context(device=-1)
    g = abi.HostGraph.from_gtba(gtba.load(SMALL[0] + ".graph.gtba"))
    ctx.region_begin(0, g)
    with pytest.raises(engine.GtbError) as e:
        ctx.pool_begin(0, 1)
    assert e.value.code == -2
    ctx.close()


def test_rejects_malformed_graph():
    d = dict(gtba.load(SMALL[0] + ".graph.gtba"))
    d["ref_var_off"] = d["ref_var_off"].copy()
    d["ref_var_off"][1] = 1  # a bubble with a single allele
    g = abi.HostGraph(d)
    ctx = engine.Context(device=-1)
    with pytest.raises(engine.GtbError):
        ctx.region_begin(0, g)
    ctx.close()


def test_scan_calls_multi_equals_per_region_calls(oracle_lib):
    """gtb_scan_calls_multi (several regions, rows back to back) = gtb_calls_from_accumulators + gtb_scan_calls per region."""
    from conftest import fixture_prefixes
    from graphtyper_b200 import abi, engine, gtba
    ctx = engine.Context(device=-1)
    accs = []
    for pre in fixture_prefixes(False)[:3]:
        g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
        rd = gtba.load(pre + ".reads.gtba")
        ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
        h = oracle_lib.index_build(g)
        accs.append(oracle_lib.result_accum(oracle_lib.pool_run(g, h, ns, abi.batch_from_probe(rd), tap=False), ns))
    V, A, R = ctx.scan_calls_multi(accs)
    ev, ea, er = [], [], []
    for a in accs:
        ph, _, _ = ctx.calls(a)
        v, al, ra = ctx.scan_calls(a, ph)
        ev.append(v)
        ea.append(al)
        er.append(ra)
    assert np.array_equal(V, np.concatenate(ev)) and np.array_equal(A, np.concatenate(ea)) and np.array_equal(R, np.concatenate(er))
