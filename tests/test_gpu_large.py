"""GPU suite (-m gpu): the large-shape fixtures (BASELINE configs 3 and 4, tests/large_fixtures.py).

 * pool50: ONE 50 kb region, a pool of 50 samples at 30x (5.2*10^5 records) -- the [bubble][sample] accumulator indexing
   at the per-GPU pool size of configs 3 / 5, cross-sample duplicate shortcut and per-read-group mate maps at scale.
 * sv1m: the genotype_sv window of 1.2 Mb with 400 <DEL>/<INS>/<DUP>, 10 samples (2.4*10^6 records) -- SV graph on the
   global-bitmap probe path (the region's k-mer set does not fit the shared-memory filter), leftover mates, is_good_read
   filtered stream, per-base ReferenceDepth tracks of 1.2 Mb per sample.
The record stream is regenerated from seeds; graph, processed-record masks and golden accumulators come from the compiled
reference (tests/golden/make_golden_large.py).  Bit-exact on every accumulator array, then PL / GT / GQ."""
import os

import numpy as np
import pytest

import compare
import large_fixtures as lf
from graphtyper_b200 import abi, engine, gtba

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(lf.CONFIGS))
def test_large_shape_matches_reference(name):
    if not os.path.exists(os.path.join(lf.LARGE_DIR, f"{name}.accum.gtba.gz")):
        pytest.skip(f"tests/golden/large/{name}.* not generated")
    ds = lf.build(name)
    ns = len(ds["readsets"])
    b = lf.batch(ds, lf.kept_indices(name, ns))
    g = abi.HostGraph.from_gtba(gtba.load(os.path.join(lf.LARGE_DIR, f"{name}.graph.gtba")))
    golden = compare.probe_accum(gtba.load(os.path.join(lf.LARGE_DIR, f"{name}.accum.gtba")))
    ctx = engine.Context(device=0)
    try:
        ctx.region_begin(3, g)
        ctx.pool_begin(3, ns)
        st = ctx.submit(3, b)
        assert st.n_records == len(b) and st.n_capacity_overflow == 0
        acc = ctx.pool_finish(3)
        compare.compare_accum(golden, acc.as_dict(), name)
        assert not acc.saturated.any()
        ph, gt, gq = ctx.calls(acc)
        pa = gtba.load(os.path.join(lf.LARGE_DIR, f"{name}.accum.gtba"))
        assert np.array_equal(ph, pa["call_phred"]) and np.array_equal(gt, pa["call_gt"]) and np.array_equal(gq, pa["call_gq"])
    finally:
        ctx.close()
