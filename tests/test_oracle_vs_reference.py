"""CPU suite: the oracle (oracle/gtb_oracle.cpp) against golden vectors produced by the compiled, unmodified
reference (tests/golden/make_golden.py).  Pins every stage: index contents and bucket order, per-seed label
lists, per-read GenotypePaths, per-bubble accumulators, PL/GT/GQ."""
import os

import numpy as np
import pytest

import compare
from conftest import fixture_prefixes
from graphtyper_b200 import abi, gtba

SMALL = fixture_prefixes(include_big=False)


def n_samples_of(rd):
    return len(rd["sample_names"].tobytes().split(b"\n")) - 1


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_oracle_matches_reference(pre, oracle_lib):
    O = oracle_lib
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    h = O.index_build(g)
    compare.compare_index(gtba.load(pre + ".index.gtba"), O.index_export(h))
    rd = gtba.load(pre + ".reads.gtba")
    b = abi.batch_from_probe(rd)
    ns = n_samples_of(rd)
    r = O.pool_run(g, h, ns, b)
    assert compare.compare_seeds(compare.probe_seeds(rd), O.result_seeds(r), "oracle") > 0
    compare.compare_paths(compare.probe_paths(rd), O.result_paths(r), "oracle")
    acc = O.result_accum(r, ns)
    pa = gtba.load(pre + ".accum.gtba")
    compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), "oracle")
    ph, gt, gq = O.calls(acc)
    assert np.array_equal(ph, pa["call_phred"])
    assert np.array_equal(gt, pa["call_gt"])
    assert np.array_equal(gq, pa["call_gq"])
    assert not acc.saturated.any()
    O.result_free(r)
    O.index_free(h)


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_oracle_connections_and_phase_support_match_reference(pre, oracle_lib):
    """HapSample::connections (vcf_writer.cpp:88-250,587-637) and the `ph` map the reference's own pool function
    derives with is_writing_hap (hts_parallel_reader.cpp:782-893)."""
    O = oracle_lib
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    h = O.index_build(g)
    rd = gtba.load(pre + ".reads.gtba")
    ns = n_samples_of(rd)
    O.set_connections(True)
    try:
        r = O.pool_run(g, h, ns, abi.batch_from_probe(rd), tap=False)
    finally:
        O.set_connections(False)
    pa = gtba.load(pre + ".accum.gtba")
    conn = O.result_connections(r)
    compare.compare_connections(compare.probe_connections(pa), abi.connections_as_table(conn), "oracle")
    acc = O.result_accum(r, ns)
    compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), "oracle+conn")
    if "ph_tuples" in pa:
        got = abi.phase_as_table(O.phase_support(acc, conn))
        assert np.array_equal(pa["ph_tuples"].reshape(-1, 5).astype(np.uint32), got)
    O.result_free(r)
    O.index_free(h)


@pytest.mark.parametrize("pre", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_oracle_record_parsing_matches_reference(pre, oracle_lib):
    """The per-record columns derived from raw records (htslib's core fields + bam1_t::data dumped by the probe) equal what
    the reference's own loop produced: AS-XS of update_paths / get_score_diff, the duplicate shortcut of the pool loop, and
    the mate links implied by its read-name maps (checked through the accumulators of a full oracle pool run)."""
    O = oracle_lib
    rd = gtba.load(pre + ".reads.gtba")
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    bam = abi.HostBamBatch.from_probe(rd)
    b = O.parse_bam(bam, is_sv=g.is_sv_graph)
    want = abi.batch_from_probe(rd)
    assert np.array_equal(b.score_diff, rd["score_diff"])
    assert np.array_equal(b.dup_of >= 0, rd["isdup"] != 0)
    for k in ("seq4", "lseq", "flag", "mapq", "isize", "same_tid", "dup_of", "mate", "leftover"):
        assert np.array_equal(getattr(b, k), getattr(want, k)), k
    ns = n_samples_of(rd)
    h = O.index_build(g)
    r = O.pool_run(g, h, ns, b, tap=False)
    compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), O.result_accum(r, ns).as_dict(), "oracle(bam)")
    O.result_free(r)
    O.index_free(h)


def test_record_parsing_corner_cases(oracle_lib):
    """Aux walks get_score_diff handles in its own way, and the read-name map's treatment of unpaired records."""
    O = oracle_lib

    def rec(name, flag, aux, seq=b"\x11" * 32, lq=64):
        qn = name + b"\0"
        qn += b"\0" * ((4 - len(qn) % 4) % 4)
        return dict(data=qn + seq + b"I" * lq + aux, l_qname=len(qn), flag=flag, lq=lq)

    recs = [
        rec(b"a", 0x41, b"ASC\x64XSC\x0a"),                       # AS 100, XS 10 -> 90
        rec(b"a", 0x00, b"NMC\x01ASc\xfb"),                       # unpaired, name waits -> pairs anyway; AS -5 (signed), no XS -> 0
        rec(b"b", 0x41, b"RGZgrp\0ASs\x2c\x01XSS\x05\x00"),       # Z string skipped; AS 300, XS 5 -> 255 (saturates)
        rec(b"b", 0x81, b"XSC\x05"),                               # no AS -> 0
        rec(b"c", 0x41, b"ASC\x32BXc\x01XSC\x28"),                 # other tags of integer type are stepped over; 50 - 40
        rec(b"c", 0x81, b"ZZB\x01ASC\x10"),                        # type 'B' ends the walk before AS -> 0
        rec(b"d", 0x41, b"ASC\x05XSC\x09"),                        # AS < XS -> 0; never gets a mate
    ]
    n = len(recs)
    core = np.zeros(n, abi.BAM_CORE_DTYPE)
    off = np.zeros(n + 1, np.uint64)
    data = b""
    for i, r in enumerate(recs):
        core["l_qseq"][i], core["flag"][i], core["l_qname"][i], core["pos"][i] = r["lq"], r["flag"], r["l_qname"], 100 + i
        data += r["data"]
        off[i + 1] = len(data)
    bam = abi.HostBamBatch(core, np.frombuffer(data, np.uint8), off, np.zeros(n, np.int32), np.zeros(n, np.int32))
    b = O.parse_bam(bam, is_sv=True)
    assert b.score_diff.tolist() == [90, 0, 255, 0, 0x32 - 0x28, 0, 0]
    assert b.mate.tolist() == [-1, 0, -1, 2, -1, 4, -1]
    assert b.leftover.tolist() == [0, 0, 0, 0, 0, 0, 1]
    assert (b.dup_of == -1).all()


def test_golden_fixtures_exercise_connections():
    n_conn = n_ph = 0
    for pre in SMALL:
        pa = gtba.load(pre + ".accum.gtba")
        n_conn += len(pa["conn_counts"])
        n_ph += len(pa.get("ph_tuples", [])) // 5
        if "ph_tuples" in pa and len(pa["ph_tuples"]):
            assert set(np.unique(pa["ph_tuples"].reshape(-1, 5)[:, 4])) <= {1, 2, 3}
    assert n_conn > 5000 and n_ph > 300


def test_kmer_codec_known_answers(oracle_lib):
    """Known answers of the reference's own unit tests for the k-mer codec
    (test/utilities/test_kmer_help_functions.cpp, test/utilities/test_utilities.cpp:123-162):
    first base in the most significant bits, A0 C1 G2 T3; 96 Hamming-1 neighbours = XOR of {1,2,3} << 2*bb."""
    # a graph with a single 32-base ref node indexes exactly one k-mer with the expected key
    seq = b"ACGT" * 8
    arrays = {
        "ref_order": np.array([1], np.uint32), "ref_seq_off": np.array([0, 32], np.uint64),
        "ref_var_off": np.array([0, 0], np.uint32), "var_order": np.zeros(0, np.uint32),
        "var_seq_off": np.array([32], np.uint64), "var_out_ref": np.zeros(0, np.uint32),
        "seq": np.frombuffer(seq, np.uint8), "var_ev_off": np.zeros(1, np.uint32), "var_ev": np.zeros(0, np.int64),
        "var_aev_off": np.zeros(1, np.uint32), "var_aev": np.zeros(0, np.int64),
        "actual_poses": np.zeros(0, np.uint32), "ref_reach_poses": np.zeros(0, np.uint32),
        "sp_keys": np.zeros(0, np.uint32), "sp_off": np.zeros(1, np.uint32), "sp_list": np.zeros(0, np.uint32)}
    g = abi.HostGraph(arrays)
    h = oracle_lib.index_build(g)
    ix = oracle_lib.index_export(h)
    key = 0
    for c in seq:
        key = (key << 2) | b"ACGT".index(c)
    assert ix["keys"].tolist() == [key]
    assert ix["labels"].tolist() == [1, 32, 0xFFFFFFFF]
    oracle_lib.index_free(h)
