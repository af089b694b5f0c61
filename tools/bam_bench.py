"""End-to-end throughput of the record-parsing entry (gtb_submit_bam_records_multi) on the bench workload, next to the column
entry (gtb_submit_reads_multi) on the same records; checks that both give the same accumulators."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench, compare
from graphtyper_b200 import abi, engine, synth

ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
bams = [abi.bam_batch_from_readsets([rs], [synth.reads_for_region(rs, b, e)]) for (b, e) in regions]
n = sum(len(b) for b in bams)
ctx = engine.Context(0)
ids = list(range(len(graphs)))
ctx.region_begin_multi(ids, graphs)
for k in ids:
    ctx.pool_begin(k, 1)
pinned, _arena = engine.pin_batches(batches)
bufs = [ctx.alloc_accumulators(k) for k in ids]

def run(fn, reps=12):
    ts = []
    for it in range(reps):
        t0 = time.perf_counter()
        ctx.pool_reset_multi(ids)
        st = fn()
        accs = ctx.pool_finish_multi(ids, out=bufs)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts[2:])), st, [a.as_dict() for a in accs]

t_col, st_col, acc_col = run(lambda: ctx.submit_multi(ids, pinned))
acc_col = [{k: v.copy() for k, v in a.items()} for a in acc_col]
t_pageable, _, _ = run(lambda: ctx.submit_bam_multi(ids, bams), reps=6)
pbams, _arena2 = engine.pin_bam_batches(bams)
t_bam, st_bam, acc_bam = run(lambda: ctx.submit_bam_multi(ids, pbams))
for a, b in zip(acc_col, acc_bam):
    compare.compare_accum(a, b, "columns vs records")
raw = sum(b.data.nbytes + b.core.nbytes + b.data_off.nbytes + b.sample.nbytes + b.rg.nbytes for b in bams)
print(f"records {n}: column entry {t_col*1e3:.3f} ms ({n/t_col/1e6:.1f} M reads/s, H2D {sum(b.nbytes_h2d() for b in batches)/1e6:.1f} MB); "
      f"record entry {t_bam*1e3:.3f} ms ({n/t_bam/1e6:.1f} M reads/s, H2D {raw/1e6:.1f} MB, pinned host buffers; {t_pageable*1e3:.3f} ms from pageable ones, "
      f"{st_bam.kernel_launches} launches); units {st_bam.n_alignments} = {st_col.n_alignments}; accumulators identical")
print(ctx.last_kernel_timing(), ctx.last_timing())
