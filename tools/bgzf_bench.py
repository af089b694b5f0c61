"""Throughput of the BGZF entry (gtb_submit_bgzf: compressed BAM bytes in, genotyped pool out) on the bench workload, region by
region, next to the record entry (gtb_submit_bam_records) on the same records and to zlib inflating the same bytes on one
host core; checks that both entries give the same accumulators.  GTB_BGZF_SERIAL_WALK=1 times the serial record walk."""
import os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench, compare
from graphtyper_b200 import abi, bgzf, engine, synth

ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
n_regions = int(os.environ.get("BB_REGIONS", len(regions)))
regions, graphs = regions[:n_regions], graphs[:n_regions]
bams = [abi.bam_batch_from_readsets([rs], [synth.reads_for_region(rs, b, e)]) for (b, e) in regions]
n = sum(len(b) for b in bams)
hdr = bgzf.bam_header([("chr1", 250000000)])
pools, queries, comp_bytes = [], [], 0
t0 = time.perf_counter()
for bam in bams:
    recs = bgzf.records_from_batch(bam)
    raw, blocks = bgzf.bgzf_compress_records(hdr, recs)
    u = bgzf.voffset_of(blocks, len(hdr))
    pools.append((raw, bgzf.HostBgzfFiles([(raw, [(u, len(raw) << 16, True)], 0, 0)], [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]])))
    pos = bam.core["pos"].astype(np.int64)
    queries.append(bgzf.query(int(bam.core["tid"][0]), int(pos.min()), int(pos.max()) + 1))
    comp_bytes += len(raw)
print(f"{n} records in {n_regions} regions, {comp_bytes/1e6:.1f} MB of BGZF (built in {time.perf_counter()-t0:.1f} s)")
t0 = time.perf_counter()
for raw, _ in pools:
    bgzf.inflate_file(raw)
t_zlib = time.perf_counter() - t0
ctx = engine.Context(0)
ids = list(range(len(graphs)))
ctx.region_begin_multi(ids, graphs)
for k in ids:
    ctx.pool_begin(k, 1)
bufs = [ctx.alloc_accumulators(k) for k in ids]

def run(fn, reps=6):
    ts, dev = [], []
    for it in range(reps):
        ctx.pool_reset_multi(ids)
        t0 = time.perf_counter()
        d = 0.0
        for k in ids:
            fn(k)
            d += ctx.last_timing()[0]
        ts.append(time.perf_counter() - t0)
        dev.append(d)
    accs = ctx.pool_finish_multi(ids, out=bufs)
    return float(np.mean(ts[1:])), float(np.mean(dev[1:])), [dict((a, b.copy()) for a, b in x.as_dict().items()) for x in accs]

t_rec, d_rec, acc_rec = run(lambda k: ctx.submit_bam(k, bams[k]))
t_bgzf, d_bgzf, acc_bgzf = run(lambda k: ctx.submit_bgzf(k, pools[k][1], queries[k]))
for a, b in zip(acc_rec, acc_bgzf):
    compare.compare_accum(a, b, "records vs bgzf")
print(f"record entry {t_rec*1e3:.2f} ms ({n/t_rec/1e6:.2f} M reads/s; copy phase {d_rec:.2f} ms); "
      f"BGZF entry {t_bgzf*1e3:.2f} ms ({n/t_bgzf/1e6:.2f} M reads/s, {comp_bytes/t_bgzf/1e9:.2f} GB/s compressed; decode phase "
      f"{d_bgzf:.2f} ms, files stitched {ctx.debug_bgzf_stitched()}); zlib inflate alone on one host core {t_zlib*1e3:.1f} ms; accumulators identical")

# ---- several pools in flight: one context per pool thread on the shared regions, as the drop-in reader runs them
T = int(os.environ.get("BB_THREADS", 4))
if T > 1:
    extra = []
    for t in range(T):
        c = engine.Context(0)
        for k in ids:
            c.region_attach(k, ctx, k)
            c.pool_begin(k, 1)
        extra.append(c)

    def pool(t, s):
        k = s % len(ids)
        extra[t].pool_reset(k)
        extra[t].submit_bgzf(k, pools[k][1], queries[k])

    reps = 4 * len(ids)
    bench.run_pipelined(T, reps, pool)
    wall = bench.run_pipelined(T, reps, pool)
    per = wall / reps
    print(f"{T} pool threads: {per*1e3:.2f} ms per region pool ({n/len(ids)/per/1e6:.2f} M reads/s, {comp_bytes/len(ids)/per/1e9:.2f} GB/s compressed)")
