"""GPU suite (-m gpu): several pool threads, one context each, sharing regions (gtb_region_attach).

The reference runs one pool per worker thread over ONE shared PHIndex (SURVEY.md section 8b "Threading",
src/typer/caller.cpp:272-391).  Here: context A builds the regions, contexts B and C attach them; three host threads
submit pools of the same regions concurrently, many times over; every thread's accumulators must equal the golden
vectors of the compiled reference bit for bit, whatever the interleaving on the device."""
import os
import threading

import numpy as np
import pytest

import compare
from conftest import fixture_prefixes
from graphtyper_b200 import abi, engine, gtba

pytestmark = pytest.mark.gpu
PRE = fixture_prefixes(include_big=False)


def _load(pre):
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
    return g, abi.batch_from_probe(rd), ns, compare.probe_accum(gtba.load(pre + ".accum.gtba"))


def test_attached_regions_three_threads():
    data = [_load(p) for p in PRE]
    ids = list(range(len(data)))
    owner = engine.Context(device=0)
    owner.region_begin_multi(ids, [d[0] for d in data])
    others = [engine.Context(device=0) for _ in range(2)]
    for c in others:
        for k in ids:
            c.region_attach(100 + k, owner, k)
    ctxs = [(owner, 0)] + [(c, 100) for c in others]
    for c, base in ctxs:
        for k in ids:
            c.pool_begin(base + k, data[k][2])
    errors = []
    start = threading.Barrier(len(ctxs))

    def work(c, base, order):
        try:
            start.wait()
            for rep in range(6):
                for k in order:
                    rid = base + k
                    c.pool_reset(rid)
                    st = c.submit(rid, data[k][1])
                    assert st.n_capacity_overflow == 0
                    acc = c.pool_finish(rid)
                    compare.compare_accum(data[k][3], acc.as_dict(), f"{os.path.basename(PRE[k])} rep {rep}")
                # all regions of this context in one region-batched call as well
                for k in ids:
                    c.pool_reset(base + k)
                c.submit_multi([base + k for k in ids], [data[k][1] for k in ids])
                for k, acc in zip(ids, c.pool_finish_multi([base + k for k in ids])):
                    compare.compare_accum(data[k][3], acc.as_dict(), f"{os.path.basename(PRE[k])} multi rep {rep}")
        except Exception as ex:  # surfaced in the main thread
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(c, base, ids[i:] + ids[:i])) for i, (c, base) in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for c in others:
        for k in ids:
            c.region_end(100 + k)
        c.close()
    owner.close()
    if errors:
        raise errors[0]


def test_attach_rejects_bad_arguments():
    pre = PRE[0]
    g, b, ns, _ = _load(pre)
    a, c = engine.Context(device=0), engine.Context(device=0)
    try:
        a.region_begin(1, g)
        with pytest.raises(engine.GtbError):
            c.region_attach(1, a, 99)  # unknown owner region
        with pytest.raises(engine.GtbError):
            a.region_attach(2, a, 1)  # same context
        c.region_attach(5, a, 1)
        with pytest.raises(engine.GtbError):
            c.region_attach(5, a, 1)  # id in use
        # the attached region's index is the owner's
        ia, ic = a.index_export(1), c.index_export(5)
        for k in ia:
            assert np.array_equal(ia[k], ic[k])
        c.region_end(5)
    finally:
        c.close()
        a.close()
