"""The harness graph builder (graphtyper_b200/graph_build.py) against graphs dumped from the compiled reference's
construct_graph (tests/golden/make_golden.py): identical node arrays, sequences and special-position tables."""
import os

import numpy as np
import pytest

from conftest import fixture_prefixes
from graphtyper_b200 import abi, graph_build, gtba, synth
from golden.make_golden import BIG, SMALL


def _cases():
    out = []
    for pre in fixture_prefixes(include_big=True):
        name, reg = os.path.basename(pre).rsplit(".r", 1)
        spec = SMALL.get(name) or BIG.get(name)
        if spec and not spec[0].get("complex_sites"):  # the harness builder restates the simple non-overlapping case only
            out.append((pre, name, int(reg), spec))
    return out


CASES = _cases()


@pytest.mark.parametrize("pre,name,reg,spec", CASES, ids=[os.path.basename(c[0]) for c in CASES])
def test_graph_builder_matches_reference(pre, name, reg, spec):
    kw, region_size = spec
    ref = synth.make_reference(kw["length"], kw.get("seed", 11))
    sites = synth.make_sites(ref, kw["n_sites"], kw.get("seed", 11) + 1)
    b, e = synth.split_regions(kw["length"], region_size)[reg]
    g = graph_build.build_region_graph(ref, sites, b, e)
    d = gtba.load(pre + ".graph.gtba")
    for k in abi.HostGraph.FIELDS:
        assert np.array_equal(g.a[k], d[k].astype(g.a[k].dtype)), k
