import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphtyper_b200 import engine
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
ctx = engine.Context(0)
ids = list(range(len(graphs)))
for rep in range(3):
    t0 = time.perf_counter(); ctx.region_begin_multi(ids, graphs); t1 = time.perf_counter()
    for k in ids: ctx.pool_begin(k, 1)
    t2 = time.perf_counter()
    print(f"rep {rep}: region_begin_multi {1e3*(t1-t0):.1f} ms  pool_begin x20 {1e3*(t2-t1):.1f} ms")
    for k in ids: ctx.region_end(k)
hctx = engine.Context(-1)
t0 = time.perf_counter(); hctx.region_begin_multi(ids, graphs); print(f"host-only index builds (parallel): {1e3*(time.perf_counter()-t0):.1f} ms")
for k in ids: hctx.region_end(k)
t0 = time.perf_counter()
for k, g in zip(ids, graphs): hctx.region_begin(k, g)
print(f"host-only index builds (serial): {1e3*(time.perf_counter()-t0):.1f} ms")
