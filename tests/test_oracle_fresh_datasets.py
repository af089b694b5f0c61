"""CPU: the oracle against the compiled reference on datasets generated on the spot (not only the committed fixtures):
fresh seeds and parameter mixes through oracle/_ref/bin/gt_probe, every stage compared.  Only where the reference was compiled."""
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest

import compare
import oracle
from graphtyper_b200 import abi, gtba

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

CASES = {
    "snp_indel_2samples": dict(length=9000, n_sites=120, n_samples=2, seed=1001, coverage=10, err=0.006, n_rate=0.001,
                               lowmapq_rate=0.05, unpaired_rate=0.03, improper_rate=0.05, flip_rate=0.3),
    "complex_3samples": dict(length=8000, n_sites=110, n_samples=3, seed=1002, coverage=8, err=0.004, complex_sites=True),
    "short_reads_dense": dict(length=5000, n_sites=200, n_samples=1, seed=1003, coverage=15, err=0.01, read_len=100, n_rate=0.003),
}


@pytest.mark.skipif(oracle.ref_binary("gt_probe") is None or oracle.ref_binary("bgzip") is None,
                    reason="compiled reference not present")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_on_fresh_dataset(name, oracle_lib):
    import make_golden
    O = oracle_lib
    tmp = tempfile.mkdtemp(prefix="gtb_fresh_")
    try:
        make_golden.run(name, CASES[name], 50000, tmp)
        pre = os.path.join(tmp, name + ".r0")
        g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
        h = O.index_build(g)
        compare.compare_index(gtba.load(pre + ".index.gtba"), O.index_export(h))
        rd = gtba.load(pre + ".reads.gtba")
        ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
        # raw records -> columns (N3), then the pool with connections on
        b = O.parse_bam(abi.HostBamBatch.from_probe(rd))
        assert np.array_equal(b.score_diff, rd["score_diff"]) and np.array_equal(b.dup_of >= 0, rd["isdup"] != 0)
        O.set_connections(True)
        try:
            r = O.pool_run(g, h, ns, b)
        finally:
            O.set_connections(False)
        assert compare.compare_seeds(compare.probe_seeds(rd), O.result_seeds(r), name) > 0
        compare.compare_paths(compare.probe_paths(rd), O.result_paths(r), name)
        pa = gtba.load(pre + ".accum.gtba")
        acc = O.result_accum(r, ns)
        compare.compare_accum(compare.probe_accum(pa), acc.as_dict(), name)
        conn = O.result_connections(r)
        compare.compare_connections(compare.probe_connections(pa), abi.connections_as_table(conn), name)
        assert np.array_equal(pa["ph_tuples"].reshape(-1, 5).astype(np.uint32), abi.phase_as_table(O.phase_support(acc, conn)))
        ph, gt, gq = O.calls(acc)
        assert np.array_equal(ph, pa["call_phred"]) and np.array_equal(gt, pa["call_gt"]) and np.array_equal(gq, pa["call_gq"])
        O.result_free(r)
        O.index_free(h)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
